"""Runs the LayerNorm backward, the column sums and the input transform a few times at ViT-B/16 batch-128 size (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402

T, D = 25216, 768
x = torch.randn(T, D, device="cuda")
g, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
dy = torch.randn(T, D, device="cuda").to(torch.bfloat16)
dres = torch.randn(T, D, device="cuda")
_, _, _, mean, rstd = vitb200.ops.layernorm_fwd(x, g, b, 1e-5, want_bf16=False)
acc = [torch.zeros(D, device="cuda") for _ in range(3)]
for _ in range(4):
    vitb200.ops.layernorm_bwd(dy, x, mean, rstd, g, dres=dres, want_f32=True, want_bf16=True, dgamma=acc[0], dbeta=acc[1],
                              dcolsum=acc[2])
x3 = torch.randn(T, 3 * D, device="cuda").to(torch.bfloat16)
for _ in range(4):
    vitb200.ops.colsum3(x3, *acc)
u8 = torch.randint(0, 256, (128, 32, 32, 3), dtype=torch.uint8, device="cuda")
tf = vitb200.DeviceImageTransform((32, 32), 224, device="cuda")
for _ in range(4):
    tf(u8)
torch.cuda.synchronize()
print("done")
