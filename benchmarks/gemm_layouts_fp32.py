"""3xTF32 engine throughput on one large product for the four operand layouts (K-major vs MN-major): isolates the
tensor core's shared-memory operand fetch rate per layout."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "benchmarks"))
import torch
from gemm_engine import run  # noqa: F401  (prints its own table first)
print("--- layouts, fp32 3xTF32, compute-bound shape")
for a_mn in (False, True):
    for b_mn in (False, True):
        run(1, 4096, 4096, 2048, torch.float32, a_mn, b_mn, iters=5)
print("--- layouts, bf16")
for a_mn in (False, True):
    for b_mn in (False, True):
        run(1, 4096, 4096, 4096, torch.bfloat16, a_mn, b_mn, iters=5)
