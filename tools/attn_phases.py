import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["VITB_ATTN_DBG"] = str(dbg.data_ptr())
import vitb200
B, N, H = 128, 197, 12
D = H * 64
qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
q, k, v = qkv[:, :, :D], qkv[:, :, D:2 * D], qkv[:, :, 2 * D:]
for _ in range(3):
    o, lse = vitb200.ops.attn_fwd(q, k, v, H)
torch.cuda.synchronize()
t = dbg.cpu().tolist()
names = ["start", "alloc+sync", "TMA q,k landed", "S mma done", "pass1 max", "pass2 exp+P", "PV mma done", "epilogue"]
for i in range(1, 8):
    print("%-16s +%6d cycles" % (names[i], t[i] - t[i - 1]))
print("total", t[7] - t[0])
