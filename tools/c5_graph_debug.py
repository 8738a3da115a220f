"""Diagnostic: capture the Res-ViT fine-tune step in a CUDA graph and print the full traceback on failure."""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402
from vitb200 import resvit  # noqa: E402

vitb200.set_precision("bf16")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
args = resvit.ModelArgs(use_lora=True, use_reslr=True, block_size=1, dynamic_active_target=0.4, lora_rank=8,
                        num_classes=100, device="cuda")
m = resvit.Transformer(args)
with torch.no_grad():
    m.pos_embedding.pos_embedding.mul_(0.02)
m = m.cuda().train()
opt = vitb200.optim.FusedAdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.05, max_grad_norm=1.0)
img = torch.randn(B, 3, 224, 224, device="cuda")
lab = torch.randint(0, 100, (B,), device="cuda")


def fl(net, x, y):
    c, a, d, e, metric = net(x, y)
    return c + a + d


try:
    gs = vitb200.train.GraphedTrainStep(m, opt, img, lab, warmup=2, forward_loss=fl)
    for _ in range(3):
        loss = gs(img, lab)
    torch.cuda.synchronize()
    print("captured OK: %d launches per step, loss %.4f" % (gs.launches_per_step, float(loss)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gs(img, lab)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("c5 Res-ViT B/16 fine-tune bs%d one CUDA graph per step: %.2f ms/step, %.0f img/s" % (B, ms, B / ms * 1e3))
except Exception:  # noqa: BLE001
    print("CAPTURE FAILED")
    traceback.print_exc()
