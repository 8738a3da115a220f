"""One forward + backward of the tcgen05 attention at the ViT-B/16 batch-128 shape (ncu target; VITB_ATTN_WS selects the kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200  # noqa: E402

# ATTN_SHAPE="B,N,H,dh" picks another shape (e.g. 32,577,12,64: the key-block backward at 384 px)
B, N, H, dh = (int(t) for t in os.environ.get("ATTN_SHAPE", "128,197,12,64").split(","))
D = H * dh
qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
do = torch.randn(B, N, D, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    o, lse = vitb200.ops.attn_fwd(q, k, v, H)
    vitb200.ops.attn_bwd(do, q, k, v, o, lse, H, dq=dqkv[:, :, :D], dk=dqkv[:, :, D:2 * D], dv=dqkv[:, :, 2 * D:])
torch.cuda.synchronize()
print("ok")
