#!/usr/bin/env python
"""bench.py — headline benchmark of the ViT encoder hot path (BASELINE.json configs[1]):

    ViT-B/16 224 px bf16 TRAIN step (forward + backward + SGD momentum 0.9), batch 128 per GPU,
    synthetic images / CIFAR-100-shaped labels, data-parallel over N B200s (weak scaling).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Prints ONE JSON line (rank 0).  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the
same through the public module API with pinned-host inputs (H2D inside the timed region, loss read back
every step); `roofline` = the dominant kernel (tcgen05 GEMM) timed live with CUDA events against the
measured dense-bf16 peak; `cpu_baseline` = the UNMODIFIED reference module (src/model.py, staged byte for byte into the
git-ignored baseline/_ref/ by __graft_entry__.build()) in the reference's own training step, timed on this box's host
cores on a bounded sample (the oracle port only if the staged copy is missing).  The oracle is only ever the checker /
baseline here — never the measured product path.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH = "b16"
IMG = 224
CLASSES = 100
BATCH_PER_GPU = 128
GFLOP_PER_IMG_TRAIN = 105.379   # 3 x 35.126 GFLOP (dense contractions only)
CPU_SAMPLE_BATCH = 8
LR, TRAIN_STEPS, WARMUP_STEPS = 0.03, 15000, 500   # src/config.py:39-42 defaults
WORKLOAD = "ViT-B/16 224px train step (fwd+bwd+SGD momentum 0.9, OneCycleLR), batch %d/GPU, C=100"

# arch geometry (src/config.py:57-104): patch, D, M, H, L
_GEO = {"b16": (16, 768, 3072, 12, 12), "l16": (16, 1024, 4096, 16, 24), "h14": (14, 1280, 5120, 16, 32)}


def vit_gflop(arch, image=224, classes=100):
    """dense contractions only, per image.  Returns (forward GFLOP of the reference algorithm, forward
    GFLOP this repo executes).  The two differ because the last block is evaluated on the class-token row only (the
    reference computes all N rows and reads row 0, src/model.py:155,210): its query / attention / output projection /
    MLP run on 1 row instead of N; K and V still need every token."""
    P, D, M, H, L = _GEO[arch]
    n_p = (image // P) ** 2
    N = n_p + 1
    qkv, attn, out, fc = 6.0 * N * D * D, 4.0 * N * N * D, 2.0 * N * D * D, 4.0 * N * D * M
    layer = qkv + attn + out + fc
    patch = 2.0 * n_p * 3 * P * P * D
    head = 2.0 * D * classes
    fwd = patch + L * layer + head
    last_exec = qkv * 2.0 / 3.0 + (qkv / 3.0 + attn + out + fc) / N
    return fwd / 1e9, (fwd - layer + last_exec) / 1e9


def resvit_train_gflop(N=197, D=768, M=3072, L=12, start=2, hdim=512, low_rank=256, r=8, active=0.4, P=16, classes=100):
    """FLOPs of one Res-ViT fine-tune step per image in the REFERENCE's semantics: the base weights are
    frozen, so every base GEMM costs forward + dgrad (2x) and no wgrad; in a dynamic layer the teacher path is forward
    only and the student path forward + dgrad; router, LoRA and approximators are trainable (3x), the approximators run
    on the skipped rows (1 - active)."""
    layer = 8.0 * N * D * D + 4.0 * N * N * D + 4.0 * N * D * M
    lora = 3 * 2.0 * N * (D * r + r * D)
    router = 2.0 * N * (D * hdim + 2 * hdim * hdim + hdim * (hdim // 2) + (hdim // 2) * 2)
    approx = 2.0 * N * (1.0 - active) * 2 * D * low_rank
    plain = start * (2 * layer + 3 * lora)
    dyn = (L - start) * (layer + 2 * layer + 3 * lora + 3 * router + 3 * approx)
    patch = 2.0 * (N - 1) * 3 * P * P * D
    head = 3 * 2.0 * D * classes
    return (patch + plain + dyn + head) / 1e9


CONFIGS = {   # BASELINE.json configs[1..4]
    "c2": dict(kind="vit_train", arch="b16", batch=128, classes=100,
               workload="ViT-B/16 224px train step (fwd+bwd+SGD momentum 0.9, OneCycleLR), batch %d/GPU, C=100"),
    "c3": dict(kind="vit_train", arch="l16", batch=64, classes=100,
               workload="ViT-L/16 224px train step (fwd+bwd+SGD momentum 0.9, OneCycleLR), batch %d/GPU, C=100"),
    "c4": dict(kind="vit_infer", arch="h14", batch=256, classes=1000,
               workload="ViT-H/14 224px inference (N=257 tokens, head_dim 80), batch %d/GPU, C=1000"),
    "c5": dict(kind="resvit_train", arch="b16", batch=128, classes=100,
               workload="Res-ViT B/16 224px fine-tune step (router target 0.4, LoRA rank 8, approximator rank 256, block_size 1; "
                        "CE + active + distill losses, AdamW + clip 1.0), batch %d/GPU, C=100"),
    # not BASELINE.json configs: the shapes the key-block attention backward exists for (profiles/)
    "h14train": dict(kind="vit_train", arch="h14", batch=32, classes=100,
                     workload="ViT-H/14 224px train step (N=257 tokens, head_dim 80; fwd+bwd+SGD), batch %d/GPU, C=100"),
    "b16_384": dict(kind="vit_train", arch="b16", batch=32, classes=100, image=384,
                    workload="ViT-B/16 384px train step (N=577 tokens; fwd+bwd+SGD), batch %d/GPU, C=100"),
}


_RESULT_FD = None


def emit(result):
    line = (json.dumps(result) + "\n").encode()
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, line)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def gemm_traffic():
    """Average DRAM bytes (read + write) per vitb_gemm_kernel launch of one training step, from the committed
    ncu capture profiles/gemm_traffic_r02.json (dram__bytes_read.sum + dram__bytes_write.sum); None if absent."""
    p = os.path.join(ROOT, "profiles", "gemm_traffic_r02.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get("dram_bytes_per_launch")


# ---------------------------------------------------------------------------------------------------
# NVLink byte counters of one GPU (NVML field values, KiB, summed over its links): evidence that the gradient exchange
# travels over NVLink, and how many bytes per step it moves
# ---------------------------------------------------------------------------------------------------
def nvlink_bytes(gpu_index):
    """(tx_bytes, rx_bytes) counted so far on all NVLinks of the GPU, or None when NVML has no such counters."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF),
                                                   (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal) * 1024)
        return tuple(out)
    except Exception:  # noqa: BLE001 - the counters are evidence, not a dependency
        return None


# ---------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self, t0=None, t1=None):
        """Median SM clock and throttle reasons over the samples taken inside [t0, t1] (epoch seconds)."""
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                    continue
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for n, v in zip(names, r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference module (baseline/_ref, staged by build()) on host cores;
# the oracle port only when the staged copy is missing
# ---------------------------------------------------------------------------------------------------
def _lr_schedule(n):
    probe = torch.optim.SGD([torch.zeros(1, requires_grad=True)], lr=LR, momentum=0.9)
    sched = torch.optim.lr_scheduler.OneCycleLR(probe, max_lr=LR, pct_start=WARMUP_STEPS / TRAIN_STEPS, total_steps=TRAIN_STEPS)
    out = []
    for _ in range(n):
        out.append(probe.param_groups[0]["lr"])
        probe.step()
        sched.step()
    return out


def reference_kind():
    from oracle import ref_loader
    return "reference" if ref_loader.available() else "port"


def cpu_train_steps(steps, warmup, batch=CPU_SAMPLE_BATCH, device="cpu", autocast=False):
    """fwd + bwd + SGD(momentum .9, lr .03, OneCycleLR) of ViT-B/16 — the reference's own training step
    (src/train.py:16-25,154-163) on the reference's own module (src/model.py:159-211, loaded unmodified from
    baseline/_ref or /root/reference), plain torch fp32 on the host cores (or `device`).  Without the staged reference:
    the oracle port (oracle/vit_oracle.py).  Returns (images/s, seconds per step, threads, kind)."""
    from oracle import ref_loader, vit_init, vit_oracle
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = vit_init.arch_cfg(ARCH, IMG, CLASSES)
    sd = vit_init.reference_state_dict(cfg, seed=0, scaled=True)
    g = torch.Generator().manual_seed(1234)
    img = torch.randn(batch, 3, IMG, IMG, generator=g).to(device)
    labels = torch.randint(0, CLASSES, (batch,), generator=g).to(device)
    dev_type = "cuda" if device != "cpu" else "cpu"
    times = []
    if ref_loader.available():
        ref = ref_loader.load_src_model()
        model = ref.VisionTransformer(attn_dropout_rate=0.0, dropout_rate=0.0, **cfg)
        model.load_state_dict(sd)
        model = model.to(device).train()
        opt = torch.optim.SGD(model.parameters(), lr=LR, weight_decay=0.0, momentum=0.9)        # src/train.py:154-158
        sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=LR, pct_start=WARMUP_STEPS / TRAIN_STEPS, total_steps=TRAIN_STEPS)
        crit = torch.nn.CrossEntropyLoss()
        for i in range(warmup + steps):
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            opt.zero_grad()
            with torch.autocast(device_type=dev_type, dtype=torch.bfloat16, enabled=autocast):
                loss = crit(model(img), labels)
            loss.backward()
            opt.step()
            sched.step()
            float(loss.detach())
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "reference"
    else:
        params = {k: v.clone().to(device) for k, v in sd.items()}
        bufs = {}
        lrs = _lr_schedule(warmup + steps)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            leaf = {k: v.detach().requires_grad_(True) for k, v in params.items()}
            with torch.autocast(device_type=dev_type, dtype=torch.bfloat16, enabled=autocast):
                loss = vit_oracle.vit_loss(img, labels, leaf)
            loss.backward()
            vit_oracle.sgd_momentum_step(params, {k: v.grad for k, v in leaf.items()}, bufs, lrs[i], 0.9, first=(i == 0))
            float(loss.detach())
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        kind = "port"
    sec = sum(times) / len(times)
    return batch / sec, sec, threads, kind


def run_reference(args, rank):
    if rank != 0:
        return
    if args.ref_device != "cpu":
        # comparator only (SURVEY 8d "stock PyTorch GPU"): the same unmodified module on the B200 through ATen / cuBLAS,
        # full batch, optionally under autocast(bf16).  Not the reference arm the driver runs.
        ips, sec, _, kind = cpu_train_steps(args.steps, args.warmup, batch=args.ref_batch or BATCH_PER_GPU, device=args.ref_device,
                                            autocast=args.ref_autocast)
        emit({"impl": "stock-torch-gpu", "metric": "train images/sec", "value": ips, "unit": "images/s", "n_gpus": 1,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
              "dtype": "bf16 autocast" if args.ref_autocast else "f32",
              "config": {"workload": WORKLOAD % (args.ref_batch or BATCH_PER_GPU),
                         "launch": "%s on %s" % ("unmodified reference src/model.py (baseline/_ref)" if kind == "reference"
                                                 else "oracle port (plain torch ops)", args.ref_device)}})
        return
    batch = args.ref_batch or CPU_SAMPLE_BATCH
    ips, sec, threads, kind = cpu_train_steps(args.steps, args.warmup, batch=batch)
    what = ("unmodified reference src/model.py + torch.optim.SGD + OneCycleLR (baseline/_ref)" if kind == "reference"
            else "oracle port (oracle/vit_oracle.py)")
    out = {
        "impl": "reference", "metric": "train images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % BATCH_PER_GPU, "parallelism": "dp%d" % args.gpus,
                   "global_batch": args.gpus * BATCH_PER_GPU, "launch": "torch CPU fp32 on the host cores (rank 0 only): " + what,
                   "sample": "each step is a bounded sample of the workload: batch %d instead of %d "
                             "(--ref-batch changes it)" % (batch, BATCH_PER_GPU),
                   "weights": "reference constructor, seed 0, attention/pos weights x0.02 (SURVEY F5)"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind,
                         "sample": "%d steps of batch %d (%s, torch CPU fp32)" % (args.steps, batch, what)},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ---------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vitb200", choices=["vitb200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS),
                    help="BASELINE.json config: c2 ViT-B/16 train bs128 (default, the headline), c3 ViT-L/16 train bs64, "
                         "c4 ViT-H/14 inference bs256, c5 Res-ViT B/16 fine-tune bs128")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of one CUDA graph")
    ap.add_argument("--ddp-mode", default="auto", choices=["auto", "graph1", "nvlink", "graph2", "overlap"],
                    help="N > 1: graph1 = the whole step incl. one gradient all-reduce as ONE CUDA graph (default, same launch "
                         "mode as N = 1); graph2 = fwd+bwd graph, one eager all-reduce, optimizer graph "
                         "(train.GraphedDataParallelStep); overlap = eager launches, per-block all-reduces overlapped with "
                         "the backward pass on a side stream (ddp.DataParallel)")
    ap.add_argument("--graph-ddp", action="store_true", help="alias of --ddp-mode graph2")
    ap.add_argument("--ref-batch", type=int, default=0, help="--impl reference only: batch of the reference step (default 8 on the CPU, 128 on a GPU)")
    ap.add_argument("--ref-device", default="cpu", help="--impl reference only: 'cuda' times the plain-torch port on the GPU (comparator)")
    ap.add_argument("--ref-autocast", action="store_true", help="--impl reference --ref-device cuda: under torch.autocast(bf16)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: anything libraries print there while the job runs (NCCL's
    # version banner at N > 1, for one) is sent to stderr instead
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # never hang a GPU box: abort the whole process if the run has not finished in time
    import signal
    signal.signal(signal.SIGALRM, lambda *_: os._exit(3))
    signal.alarm(int(os.environ.get("VITB_BENCH_TIMEOUT_S", "900")))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    import torch.distributed as dist
    import vitb200

    cfg = CONFIGS[args.config]
    kind, arch, classes = cfg["kind"], cfg["arch"], cfg["classes"]
    B = args.batch if args.batch > 0 else cfg["batch"]
    image = cfg.get("image", IMG)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vitb200.set_precision("bf16")
    torch.manual_seed(0)                     # same constructor RNG sequence as the reference (tests/test_oracle.py)
    train = kind != "vit_infer"
    # gradient exchange at N > 1: "graph1" = one NCCL all-reduce inside the step graph; "nvlink" = the same with
    # vitb_p2p_allreduce (our kernel over NVLink peer memory / NVSwitch multicast) in its place
    # "auto" (default) = "nvlink" when the symmetric gradient buffer can be set up on every rank, else "graph1"
    xchg = None
    if world > 1 and train and args.ddp_mode in ("auto", "nvlink") and not args.no_graph and not args.graph_ddp:
        xchg = vitb200.p2p.NvlinkExchange()
    elif args.ddp_mode == "auto":
        args.ddp_mode = "graph1"

    def make_optimizer(build):
        """build(grad_buffer_factory) -> optimizer.  With an NVLink exchange the gradient buffer goes into symmetric
        memory; if that fails on ANY rank, every rank falls back to an ordinary buffer and the NCCL all-reduce."""
        nonlocal xchg
        if xchg is None:
            return build(None)
        opt_, ok = None, 1
        try:
            opt_ = build(xchg.allocate)
        except Exception as exc:  # noqa: BLE001
            ok = 0
            print("bench: symmetric gradient buffer unavailable on rank %d (%r)" % (rank, exc), file=sys.stderr)
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag) == 1:
            args.ddp_mode = "nvlink"
            return opt_
        if args.ddp_mode == "nvlink":
            raise RuntimeError("--ddp-mode nvlink: the symmetric gradient buffer could not be set up on every rank")
        xchg, args.ddp_mode = None, "graph1"
        if rank == 0:
            print("bench: falling back to the NCCL all-reduce (--ddp-mode graph1)", file=sys.stderr)
        return build(None)
    if kind == "resvit_train":
        # res-vit/config.py presets + the fine-tune switches of BASELINE.json configs[4] (res-vit/ft_resvit.sh)
        from vitb200 import resvit
        margs = resvit.ModelArgs(use_lora=True, use_reslr=True, block_size=1, dynamic_active_target=0.4, lora_rank=8,
                                 num_classes=classes, device="cuda")
        model = resvit.Transformer(margs)
        with torch.no_grad():
            model.pos_embedding.pos_embedding.mul_(0.02)
        model = model.to(dev).train()
        # res-vit/train.py:272-277,65: AdamW(1e-4, wd 0.05) over the trainable set + clip_grad_norm_(1.0)
        opt = make_optimizer(lambda fac: vitb200.optim.FusedAdamW([q for q in model.parameters() if q.requires_grad], lr=1e-4,
                                                                  weight_decay=0.05, max_grad_norm=1.0, grad_buffer_factory=fac))
        resvit.bind_optimizer(model, opt)   # approximators whose key does not occur in a batch are skipped, as torch's AdamW does
        sched = None
        gflop_ref = gflop_exec = resvit_train_gflop(classes=classes)
        if world > 1:   # ActiveLoss is (batch mean - target)^2: make it the loss of the GLOBAL batch (one scalar all-reduce)
            model.criterion_active.sync_group = True

        def forward_loss(net, img, labels):        # res-vit/train.py:30,51-52: c_loss + a_loss + d_loss
            c, a, d, e, metric = net(img, labels)
            return c + a + d
    else:
        model = vitb200.build_vit(arch, image, classes)
        with torch.no_grad():                    # SURVEY.md F5 recipe: trained-like scale for attention / pos weights
            for k, v in model.state_dict().items():
                if k.endswith(("attn.query.weight", "attn.key.weight", "attn.value.weight", "attn.out.weight",
                               "pos_embedding.pos_embedding")):
                    v.mul_(0.02)
        model = model.to(dev)
        model.train(train)
        fwd_ref, fwd_exec = vit_gflop(arch, image, classes)
        gflop_ref, gflop_exec = (3 * fwd_ref, 3 * fwd_exec) if train else (fwd_ref, fwd_exec)
        opt = sched = None
        if train:
            opt = make_optimizer(lambda fac: vitb200.optim.FusedSGD(model.parameters(), lr=LR, momentum=0.9, grad_buffer_factory=fac))
            # the reference's schedule (src/train.py:159-163, config defaults src/config.py:39-42): the step starts at
            # max_lr / 25; lr AND the cycled momentum reach the kernels through device scalars, so they also drive the graph
            sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=LR, pct_start=WARMUP_STEPS / TRAIN_STEPS, total_steps=TRAIN_STEPS)

        def forward_loss(net, img, labels):
            return vitb200.functional.cross_entropy(net(img), labels)
    if args.graph_ddp:
        args.ddp_mode = "graph2"
    if args.no_graph and world > 1:
        args.ddp_mode = "overlap"
    if not train:
        args.ddp_mode = "replicas"           # inference: independent replicas, no exchange
    graph_ddp = world > 1 and args.ddp_mode == "graph2"
    graph_one = world > 1 and args.ddp_mode in ("graph1", "nvlink")
    net = vitb200.ddp.DataParallel(model, opt) if (world > 1 and args.ddp_mode == "overlap") else model
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    img_d = torch.randn(B, 3, image, image, generator=gen, device=dev)
    lab_d = torch.randint(0, classes, (B,), generator=gen, device=dev)
    img_h = img_d.cpu().pin_memory()
    lab_h = lab_d.cpu().pin_memory()

    def eager_step(img, labels):
        if not train:
            with torch.no_grad():
                return net(img)[:, 0].sum()          # a scalar of the result (read back in the e2e leg)
        opt.zero_grad()
        loss = forward_loss(net, img, labels)
        loss.backward()
        opt.step()
        if sched is not None:
            sched.step()
        return loss

    graphed = None
    # One CUDA graph per step: fwd + bwd (+ ONE gradient all-reduce on the capture stream at N > 1) + optimizer.
    if graph_ddp:
        graphed = vitb200.train.GraphedDataParallelStep(net, opt, img_d, lab_d)
    elif train and not args.no_graph and (world == 1 or graph_one):
        try:   # the whole step (fwd + bwd + all-reduce + optimizer) as one replayable CUDA graph
            graphed = vitb200.train.GraphedTrainStep(net, opt, img_d, lab_d, data_parallel=graph_one, exchange=xchg,
                                                     forward_loss=forward_loss if kind == "resvit_train" else None)
        except Exception as exc:  # noqa: BLE001 - report and measure eagerly rather than die
            if world > 1:
                raise     # the eager fallback below has no gradient exchange
            if rank == 0:
                print("bench: CUDA-graph capture failed (%r); timing the eager step" % (exc,), file=sys.stderr)
            graphed = None
            torch.cuda.synchronize()
    if graphed is not None:
        def step(img, labels):          # src/train.py:19-25: ... optimizer.step(); lr_scheduler.step()
            loss = graphed(img, labels)
            if sched is not None:
                sched.step()
            return loss
    else:
        step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # started before the warm-up so samples exist for short timed regions
    for _ in range(args.warmup):
        step(img_d, lab_d)
    barrier()
    t_wall0 = time.time()
    launches0 = vitb200._lib.LAUNCHES[0]
    nvl0 = nvlink_bytes(local) if (world > 1 and rank == 0) else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(img_d, lab_d)
    e1.record()
    torch.cuda.synchronize()
    nvl1 = nvlink_bytes(local) if nvl0 is not None else None
    ms = e0.elapsed_time(e1)
    launches = vitb200._lib.LAUNCHES[0] - launches0
    if graphed is not None:
        launches = graphed.launches_per_step * args.steps
    clocks = sampler.stop(t_wall0, time.time()) if rank == 0 else None
    barrier()
    # end to end through the public API: every step's inputs come from pinned host memory (H2D inside the timed
    # region, staged one batch ahead on a side stream by vitb200.train.InputPrefetcher) and the loss is read back
    into = (graphed.images, graphed.labels) if graphed is not None else None
    pre = vitb200.train.InputPrefetcher(img_d, lab_d, into=into)
    step(img_d, lab_d)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    last = 0.0
    pre.start(img_h, lab_h)
    for i in range(args.steps):
        img_e, lab_e = pre.get()
        if i + 1 < args.steps:
            pre.start(img_h, lab_h)        # the next step's H2D overlaps this step's kernels
        last = float(step(img_e, lab_e))   # D2H read of the loss every step
    f1.record()
    torch.cuda.synchronize()
    ms_e2e = f0.elapsed_time(f1)
    # dominant-kernel roofline: every tcgen05 GEMM launch of two steps timed with CUDA events
    vitb200.ops.PROFILE_GEMM = []
    eager_step(img_d, lab_d)
    eager_step(img_d, lab_d)
    torch.cuda.synchronize()
    recs = vitb200.ops.PROFILE_GEMM
    vitb200.ops.PROFILE_GEMM = None
    gemm_ms = sum(r[0].elapsed_time(r[1]) for r in recs)
    gemm_flop = sum(r[2] for r in recs)
    gemm_bytes = sum(r[3] for r in recs)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        pk = peaks()
        ips = world * B * args.steps / (ms / 1e3)
        ips_e2e = world * B * args.steps / (ms_e2e / 1e3)
        ach = gemm_flop / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        step_tf = ips / world * gflop_ref / 1e3
        step_tf_exec = ips / world * gflop_exec / 1e3
        out = {
            "metric": "train images/sec" if train else "inference images/sec", "value": ips, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["workload"] % B, "baseline_config": args.config,
                       "parallelism": ("dp%d" % world) if train else ("%d replicas" % world), "global_batch": world * B,
                       "launch": ("fwd+bwd graph, one eager all-reduce, optimizer graph" if graph_ddp else "one CUDA graph per step")
                       if graphed is not None else "eager (Python launches)",
                       "grad_exchange": (None if (world == 1 or not train) else
                                         {"graph1": "one NCCL all-reduce (AVG) of the flat fp32 gradient buffer, a node of the step graph",
                                          "nvlink": "vitb_p2p_allreduce (one kernel over NVLink peer memory, %s), a node of the step graph"
                                                    % (xchg.mode if xchg is not None else "-"),
                                          "graph2": "one eager NCCL all-reduce between two graphs",
                                          "overlap": "per-block NCCL all-reduces on a side stream, overlapped with backward"}[args.ddp_mode]),
                       "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2; no flush needed",
                       "weights": "reference constructor, seed 0, attention/pos weights x0.02 (SURVEY F5)"},
            "clocks": clocks,
            "nvlink": (None if nvl1 is None else
                       {"tx_bytes_per_step_rank0": (nvl1[0] - nvl0[0]) / args.steps, "rx_bytes_per_step_rank0": (nvl1[1] - nvl0[1]) / args.steps,
                        "grad_bytes": sum(p.numel() for p in net.parameters() if p.requires_grad) * 4,
                        "source": "NVML NVLINK_THROUGHPUT_DATA_TX/RX of rank 0's GPU over the timed steps"}),
            "e2e": {"value": ips_e2e, "unit": "images/s", "h2d_bytes_per_step": img_h.numel() * 4 + lab_h.numel() * 8,
                    "d2h_bytes_per_step": 4, "last_loss": last},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "tcgen05 GEMMs: vitb_gemm_kernel + vitb_wgrad_pair_kernel (cta_group::2)", "achieved": ach,
                         "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tflops_sustained"],
                         "peak_source": pk["source"] + " sustained cuBLAS bf16", "traffic": gemm_traffic(),
                         "traffic_source": "profiles/gemm_traffic_r02.json (ncu dram__bytes over the GEMM launches of one step of this "
                                           "command; null when the capture is missing)",
                         "algorithmic_bytes_per_launch": gemm_bytes / max(1, len(recs)),
                         "flop_per_launch": gemm_flop / max(1, len(recs)),
                         "gemm_share_of_step": (gemm_ms / 2) / (ms / args.steps), "launches_timed": len(recs)},
            "roofline_step": {"achieved": step_tf, "unit": "TFLOP/s per GPU at %.3f GFLOP/img (the reference algorithm's dense "
                                                             "contractions, SURVEY.md App. A)" % gflop_ref,
                              "frac_of_sustained": step_tf / pk["tflops_sustained"],
                              "frac_of_burst": step_tf / pk["tflops_burst"], "frac_of_spec_2250": step_tf / 2250.0,
                              "executed": {"achieved": step_tf_exec, "gflop_per_img": gflop_exec,
                                           "frac_of_sustained": step_tf_exec / pk["tflops_sustained"],
                                           "note": "FLOPs this repo executes: the last block runs on the class-token row only"}},
        }
        if not train:
            out["latency_ms_per_batch"] = ms / args.steps
        if world == 1 and not args.no_cpu_baseline and args.config == "c2":
            cips, csec, threads, kind_ = cpu_train_steps(2, 1)
            out["cpu_baseline"] = {"value": cips, "unit": "images/s", "cores": threads, "kind": kind_,
                                   "sample": "2 timed steps of batch %d after 1 warm-up (%s, torch CPU fp32)"
                                             % (CPU_SAMPLE_BATCH, "unmodified reference src/model.py, baseline/_ref" if kind_ == "reference"
                                                else "oracle port")}
        emit(out)
    if world > 1:
        # a process group whose collectives were captured into CUDA graphs does not always tear down cleanly (N = 2 run of
        # round 2: destroy_process_group() never returned after the result line): leave together and skip the teardown
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
