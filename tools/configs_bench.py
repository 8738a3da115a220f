"""Throughput of the other BASELINE.json configs on one B200 (diagnostic; the headline is bench.py):
  c3  ViT-L/16 224 bf16 train step, batch 64
  c4  ViT-H/14 224 bf16 inference, batch 256 (N = 257 tokens)
  c5  Res-ViT B/16 fine-tune step, router target 0.4, LoRA rank 8, bf16, batch 128 (and 32)
  e384  ViT-B/16 384 px inference, batch 64 (N = 577 tokens: the reference's default evaluation resolution)
Writes gpurun_out/configs_bench.txt."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402
from vitb200 import resvit  # noqa: E402

lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    lines.append(s)


def scale_attn(model):
    with torch.no_grad():
        for k, v in model.state_dict().items():
            if k.endswith(("attn.query.weight", "attn.key.weight", "attn.value.weight", "attn.out.weight",
                           "pos_embedding.pos_embedding")):
                v.mul_(0.02)


def timed(fn, steps, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


which = sys.argv[1:] or ["c3", "c4", "c5"]
vitb200.set_precision("bf16")
if "c3" in which:
    torch.manual_seed(0)
    m = vitb200.build_vit("l16", 224, 100)
    scale_attn(m)
    m = m.cuda().train()
    opt = vitb200.optim.FusedSGD(m.parameters(), lr=0.03, momentum=0.9)
    B = 64
    img = torch.randn(B, 3, 224, 224, device="cuda")
    lab = torch.randint(0, 100, (B,), device="cuda")

    def step():
        opt.zero_grad()
        loss = vitb200.functional.cross_entropy(m(img), lab)
        loss.backward()
        opt.step()
    ms = timed(step, 10, 3)
    log("c3 ViT-L/16 train bs64: %.2f ms/step, %.0f img/s, %.0f TFLOP/s (369.323 GFLOP/img)" % (ms, B / ms * 1e3, B / ms * 369.323))
    del m, opt
    torch.cuda.empty_cache()
if "c4" in which:
    torch.manual_seed(0)
    m = vitb200.build_vit("h14", 224, 1000)
    scale_attn(m)
    m = m.cuda().eval()
    B = 256
    img = torch.randn(B, 3, 224, 224, device="cuda")
    with torch.no_grad():
        ms = timed(lambda: m(img), 5, 2)
    log("c4 ViT-H/14 inference bs256 (N=257, head_dim 80, tcgen05 wide-head attention): %.2f ms/batch, %.0f img/s, %.0f TFLOP/s (334.588 GFLOP/img)"
        % (ms, B / ms * 1e3, B / ms * 334.588))
    del m
    torch.cuda.empty_cache()
if "e384" in which:
    torch.manual_seed(0)
    m = vitb200.build_vit("b16", 384, 1000)
    scale_attn(m)
    m = m.cuda().eval()
    B = 64
    img = torch.randn(B, 3, 384, 384, device="cuda")
    with torch.no_grad():
        ms = timed(lambda: m(img), 5, 2)
        out = m(img)
    log("e384 ViT-B/16 384px inference bs64 (N=577, multi-block tcgen05 attention): %.2f ms/batch, %.0f img/s, logits finite %s"
        % (ms, B / ms * 1e3, bool(torch.isfinite(out).all())))
    del m
    torch.cuda.empty_cache()
if "c5" in which:
    for B in (32, 128):
        torch.manual_seed(0)
        args = resvit.ModelArgs(use_lora=True, use_reslr=True, block_size=1, dynamic_active_target=0.4, lora_rank=8,
                                num_classes=100, device="cuda")
        m = resvit.Transformer(args)
        with torch.no_grad():
            m.pos_embedding.pos_embedding.mul_(0.02)
        m = m.cuda().train()
        opt = vitb200.optim.FusedAdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.05,
                                       max_grad_norm=1.0)
        img = torch.randn(B, 3, 224, 224, device="cuda")
        lab = torch.randint(0, 100, (B,), device="cuda")

        def step():
            opt.zero_grad()
            c, a, d, e, metric = m(img, lab)
            (c + a + d).backward()
            opt.step()
        ms = timed(step, 5, 2)
        log("c5 Res-ViT B/16 fine-tune (router 0.4, LoRA r8, block_size 1) bs%d eager: %.2f ms/step, %.0f img/s" % (B, ms, B / ms * 1e3))
        try:
            def fl(net, x, y):
                c, a, d, e, metric = net(x, y)
                return c + a + d
            gs = vitb200.train.GraphedTrainStep(m, opt, img, lab, warmup=1, forward_loss=fl)
            ms = timed(lambda: gs(img, lab), 10, 2)
            log("c5 Res-ViT B/16 fine-tune bs%d one CUDA graph per step (%d launches): %.2f ms/step, %.0f img/s"
                % (B, gs.launches_per_step, ms, B / ms * 1e3))
        except Exception as exc:  # noqa: BLE001
            log("c5 graph capture failed:", repr(exc)[:300])
        del m, opt
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/configs_bench.txt", "a") as fh:
    fh.write("\n".join(lines) + "\n")
