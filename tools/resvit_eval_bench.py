"""Res-ViT B/16 inference (BASELINE.json configs[4] geometry, batch 128) with and without device-side token compaction
(VITB_RESVIT_COMPACT): images/s of the dense-select path vs the path that runs the output projection and the MLP on the
active rows only, at the keep ratio the (randomised) routers produce.  Diagnostic; writes one line per setting."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402
from vitb200 import resvit  # noqa: E402


def timed(fn, steps=10, warm=3):
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


B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
vitb200.set_precision("bf16")
torch.manual_seed(0)
args = resvit.ModelArgs(use_lora=True, use_reslr=True, block_size=1, dynamic_active_target=0.4, lora_rank=8, num_classes=100,
                        device="cuda")
m = resvit.Transformer(args)
gen = torch.Generator().manual_seed(1)
with torch.no_grad():
    m.pos_embedding.pos_embedding.mul_(0.02)
    for name, p in m.named_parameters():          # routers that actually split the tokens (they keep everything at init)
        if name.endswith("router.out_conv.4.weight"):
            p.copy_(torch.randn(p.shape, generator=gen) * 0.5)
        elif name.endswith("router.out_conv.4.bias"):
            p.copy_(torch.tensor([0.1, -0.1]).repeat(p.numel() // 2))        # lean slightly towards "skip" (target keep ratio 0.4)
m = m.cuda().eval()
img = torch.randn(B, 3, 224, 224, device="cuda")
lab = torch.randint(0, 100, (B,), device="cuda")
out = {}
for flag in ("0", "1", "0", "1"):
    os.environ["VITB_RESVIT_COMPACT"] = flag
    with torch.no_grad():
        ms = timed(lambda: m(img, lab))
        c, a, d, e, metric = m(img, lab)
        torch.cuda.synchronize()
        out[flag] = m.logits.float().clone()
        # the same forward as ONE CUDA graph: nothing in either path reads the device (the compacted path keeps its row list
        # and row count in device memory), so eval is capturable; Python launch overhead (~1,000 launches) then drops out
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            m(img, lab)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            m(img, lab)
        ms_g = timed(g.replay)
    print("VITB_RESVIT_COMPACT=%s  Res-ViT B/16 eval bs%d: eager %.2f ms/batch (%.0f img/s), one CUDA graph %.2f ms/batch (%.0f img/s), keep ratio %.3f"
          % (flag, B, ms, B / ms * 1e3, ms_g, B / ms_g * 1e3, float(metric["non_low_rank_ratio"])), flush=True)
print("logits rel diff compact vs dense: %.2e" % float((out["1"] - out["0"]).norm() / out["0"].norm()))
