"""clock64 timeline of block 0 of the fp32 TMEM-operand GEMM engine (k_tc_gemm_ts) on the C2 backward products.
Columns per k-block: tma issue | split start (stage landed), A read done, ring wait done, arrive | mma ready, mma issued.
"""
import ctypes
import os
import sys

sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200"))
import torch  # noqa: E402
from tgp_b200 import _lib as L  # noqa: E402

lib = L.load()
lib.tgpb200_debug_engine_timeline.argtypes = [ctypes.c_void_p]
B, N, K, F = 512, 256, 64, 128
dev = "cuda"
a = (torch.rand(B, N, N, device=dev) < 0.05).float()
s = torch.softmax(torch.randn(B, N, K, device=dev), -1)
x = torch.randn(B, N, F, device=dev)
gx = torch.randn(B, K, F, device=dev)
st = torch.cuda.current_stream().cuda_stream


def run(which):
    if which == "W":   # W = A S : A K-major [N, N], B = S MN-major
        out = torch.empty(B, N, K, device=dev)
        args = (a.data_ptr(), s.data_ptr(), out.data_ptr(), B, N, K, N, N * N, N, 0, N * K, K, 1, N * K, K, 1, 0, 0, 1.0, 0, st)
    else:              # dX = S Gx : A = S K-major [N, K], B = Gx MN-major [K, F]
        out = torch.empty(B, N, F, device=dev)
        args = (s.data_ptr(), gx.data_ptr(), out.data_ptr(), B, N, F, K, N * K, K, 0, K * F, F, 1, N * F, F, 1, 0, 0, 1.0, 0, st)
    for _ in range(3):
        assert lib.tgpb200_tc_gemm(*args) == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.tgpb200_tc_gemm(*args)
    e1.record(); torch.cuda.synchronize()
    print(f"== {which}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us per launch")
    dbg = torch.zeros(160 * 8, dtype=torch.long, device=dev)
    lib.tgpb200_debug_engine_timeline(dbg.data_ptr())
    lib.tgpb200_tc_gemm(*args); torch.cuda.synchronize()
    lib.tgpb200_debug_engine_timeline(None)
    d = dbg.cpu().view(160, 8)
    t0 = int(d[0, 0])
    print(" kc  tma_issue | split_start A_read_done ring_wait_done arrive | mma_ready mma_issued")
    for i in range(0, 56):
        r = [int(v) - t0 for v in d[i, :8]]
        print(f"{i:3d} {r[0]:9d} | {r[1]:9d} {r[2]:9d} {r[3]:9d} {r[4]:9d} | {r[5]:9d} {r[6]:9d}")
    print(" item  epi_start epi_end")
    for i in range(0, 8):
        r = [int(v) - t0 for v in d[128 + i, :2]]
        print(f"{i:3d} {r[0]:9d} {r[1]:9d}")


for w in sys.argv[1:] or ["W", "dX"]:
    run(w)
