import sys, os, ctypes
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200")); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from test_gpu_tc_gemm import _run, _ref
for dt in (torch.float32, torch.bfloat16):
  for a_mn, b_mn in [(False, False), (True, False), (False, True)]:
    B, M, N, Kd = 1, 128, 64, 64
    g = torch.Generator().manual_seed(1)
    a = torch.randn((B, Kd, M) if a_mn else (B, M, Kd), generator=g).to(dt).cuda()
    b = torch.randn((B, Kd, N) if b_mn else (B, N, Kd), generator=g).to(dt).cuda()
    out = _run(a, b, a_mn, b_mn, M, N, Kd)
    ref = _ref(a, b, a_mn, b_mn)
    d = (out.double() - ref).abs()
    print(os.environ.get("TGPB200_DBG_MODE"), dt, a_mn, b_mn, "maxerr", round(d.max().item(), 4), "absmax out", round(out.abs().max().item(), 3), "nonzero frac", (out != 0).double().mean().item())
