"""Micro-benchmark of the GEMM engine on the C3 level-1 (bf16) product shapes."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200"))
import torch
from tgp_b200 import _lib as L
def run(B, M, N, Kd, dt, a_mn, b_mn, iters=10, out_dt=None):
    a = torch.randn((B, Kd, M) if a_mn else (B, M, Kd), device="cuda").to(dt)
    b = torch.randn((B, Kd, N) if b_mn else (B, N, Kd), device="cuda").to(dt)
    out_dt = out_dt or torch.float32
    out = torch.empty(B, M, N, device="cuda", dtype=out_dt)
    args = (L.ptr(a), L.ptr(b), L.ptr(out), B, M, N, Kd, a.stride(0), a.stride(1), int(a_mn), b.stride(0), b.stride(1), int(b_mn), M*N, N, 1, L.dtype_code(dt), L.dtype_code(out_dt), 1.0, 0, L.stream())
    for _ in range(3): L.call("tgpb200_tc_gemm", *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): L.call("tgpb200_tc_gemm", *args)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bytes_ = (a.numel() + b.numel()) * a.element_size() + out.numel() * out.element_size()
    print(f"B={B} M={M} N={N} K={Kd} {str(dt)[6:]}->{str(out_dt)[6:]} a_mn={int(a_mn)} b_mn={int(b_mn)}: {ms*1e3:8.1f} us  {bytes_/ms/1e6:7.0f} GB/s  {2*B*M*N*Kd/ms/1e9:8.1f} TFLOP/s")
bf = torch.bfloat16
run(1024, 512, 256, 1024, bf, False, False, out_dt=bf)   # dS-like total K
run(1024, 512, 256, 512, bf, True, True, out_dt=bf)      # Tt-like
run(1024, 512, 256, 512, bf, False, True, out_dt=bf)     # W-like
run(1024, 512, 256, 256, bf, False, True, out_dt=bf)     # dX-like
run(1024, 256, 256, 512, bf, True, True)                 # Araw / M-like
run(64, 2048, 256, 2048, bf, False, False, out_dt=bf)
