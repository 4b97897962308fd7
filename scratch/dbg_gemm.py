import sys, os, ctypes
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200")); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from tgp_b200 import _lib as L
from test_gpu_tc_gemm import _run, _ref
lib = L.load()
lib.tgpb200_debug_set_dump.argtypes = [ctypes.c_void_p]
for a_mn, b_mn in [(True, False)]:
    B, M, N, Kd = 1, 128, 64, 32
    a = (torch.arange(Kd * M, dtype=torch.float32).view(1, Kd, M)).cuda()   # a[k, m] = k*128 + m
    b = torch.ones(B, N, Kd).cuda()
    dump = torch.full((6144,), -7.0, device="cuda")
    lib.tgpb200_debug_set_dump(dump.data_ptr())
    out = _run(a, b, a_mn, b_mn, M, N, Kd)
    lib.tgpb200_debug_set_dump(None)
    d = dump.cpu()
    print("A tile first 2 rows (32 floats each):")
    print(d[:32].tolist()); print(d[32:64].tolist()); print("row 8:", d[256:288].tolist())
    print("block1 row0:", d[1024:1056].tolist())
    print("B tile row0:", d[4096:4128].tolist())
    print("out row0", out[0, 0, :4].tolist(), "expected", float(a[0, :, 0].sum()))
