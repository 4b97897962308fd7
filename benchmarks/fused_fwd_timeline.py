"""Per-role clock64 timeline of block 0 of the fused dense forward kernel (tgpb200_debug_fused_timeline)."""
import sys, os, ctypes
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200"))
import torch, tgp_b200 as T
from tgp_b200 import _lib as L
lib = L.load(); lib.tgpb200_debug_fused_timeline.argtypes = [ctypes.c_void_p]
B, N, K, F = 512, 256, 64, 128
a = (torch.rand(B, N, N, device="cuda") < 0.05).float(); s = torch.softmax(torch.randn(B, N, K, device="cuda"), -1); x = torch.randn(B, N, F, device="cuda")
for _ in range(3): T.mincut_pool(x, a, s)
dbg = torch.zeros(160 * 8, dtype=torch.long, device="cuda")
lib.tgpb200_debug_fused_timeline(dbg.data_ptr())
T.mincut_pool(x, a, s); torch.cuda.synchronize()
lib.tgpb200_debug_fused_timeline(None)
d = dbg.cpu().view(160, 8)
t0 = int(d[0, 0])
print("kb   tma_start tma_done | mma_ready ring_wait_done mma_done | split_start split_done pass1_done   (cycles since first TMA; TMEM-operand kernel)")
for i in list(range(0, 24)) + list(range(40, 52)):
    r = [int(v) - t0 for v in d[i, :8]]
    print(f"{i:3d} {r[0]:9d} {r[1]:8d} | {r[2]:8d} {r[3]:7d} {r[4]:8d} | {r[5]:9d} {r[6]:9d} {r[7]:9d}")
