"""Micro-benchmark of the tcgen05 batched GEMM engine (C2-shaped and large products). Run on a B200:
    python benchmarks/gemm_engine.py
"""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200"))
import torch
from tgp_b200 import _lib as L
def run(B, M, N, Kd, dt, a_mn, b_mn, iters=20):
    a = torch.randn((B, Kd, M) if a_mn else (B, M, Kd), device="cuda").to(dt)
    b = torch.randn((B, Kd, N) if b_mn else (B, N, Kd), device="cuda").to(dt)
    out = torch.empty(B, M, N, device="cuda")
    args = (L.ptr(a), L.ptr(b), L.ptr(out), B, M, N, Kd, a.stride(0), a.stride(1), int(a_mn), b.stride(0), b.stride(1), int(b_mn), M*N, N, 1, L.dtype_code(dt), 0, 1.0, 0, L.stream())
    for _ in range(3): L.call("tgpb200_tc_gemm", *args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): L.call("tgpb200_tc_gemm", *args)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    bytes_ = (a.numel() + b.numel()) * a.element_size() + out.numel() * 4
    print(f"B={B} M={M} N={N} K={Kd} {str(dt)[6:]} a_mn={int(a_mn)} b_mn={int(b_mn)}: {ms*1e3:8.1f} us  {bytes_/ms/1e6:7.0f} GB/s  {2*B*M*N*Kd/ms/1e9:8.1f} TFLOP/s")
for dt in (torch.float32, torch.bfloat16):
    run(512, 256, 64, 256, dt, True, True)     # Tt-like
    run(512, 256, 64, 256, dt, False, True)    # W-like
    run(512, 256, 64, 256, dt, False, False)
    run(512, 256, 128, 256, dt, False, False)
    run(1, 8192, 8192, 8192 if dt == torch.bfloat16 else 2048, dt, False, False, iters=5)
    run(64, 2048, 128, 2048, dt, False, False, iters=5)
