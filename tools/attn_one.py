import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitb200
B, N, H = 128, 197, 12
D = H * 64
qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
for _ in range(3):
    o, lse = vitb200.ops.attn_fwd(q, k, v, H)
do = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    vitb200.ops.attn_bwd(do, q, k, v, o, lse, H, dq=dqkv[:, :, :D], dk=dqkv[:, :, D:2 * D], dv=dqkv[:, :, 2 * D:])
torch.cuda.synchronize()
print("done")
