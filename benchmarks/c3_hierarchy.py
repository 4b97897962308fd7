"""C3: three chained DiffPool levels in bf16, (N, K) = (512, 256) -> (256, 64) -> (64, 16), F = 256.

Level l+1 consumes level l's post-processed A_pool and X_pool directly (no GNN layer in between: it is outside the
Reduce + Connect path).  Reports forward+backward time per level and for the chain.

    python benchmarks/c3_hierarchy.py [B]
"""
import json
import os
import sys

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200"))
import torch

from tgp_b200 import functional as F_

LEVELS = [(512, 256), (256, 64), (64, 16)]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    Fdim, dev, dt = 256, "cuda", torch.bfloat16
    g = torch.Generator().manual_seed(0)
    a = (torch.rand(B, 512, 512, generator=g) < 0.02).float()
    a = torch.triu(a, 1)
    a = (a + a.transpose(1, 2)).to(dev, dt)
    x = torch.randn(B, 512, Fdim, generator=g).to(dev, dt).requires_grad_(True)
    s = [torch.softmax(torch.randn(B, n, k, generator=g), -1).to(dev, dt).requires_grad_(True) for n, k in LEVELS]

    def level(x_, a_, s_, n):
        return F_.dense_pool(x_, a_, s_, remove_self_loops=True, degree_norm=True, adj_transpose=True,
                             loss_kind=F_.LOSS_DIFFPOOL, ent_div=float(B * n))

    def chain():
        for t in s:
            t.grad = None
        x.grad = None
        xl, al, total = x, a, 0.0
        for (n, k), sl in zip(LEVELS, s):
            xl, al, losses = level(xl, al, sl, n)
            total = total + losses[2] + losses[3]
        (xl.float().sum() + al.float().sum() + total).backward()

    def timed(fn, iters=20):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    out = {"B": B, "dtype": "bf16", "levels": []}
    # per level in isolation (inputs of level l produced once by the chain's forward)
    xl, al = x.detach(), a
    for (n, k), sl in zip(LEVELS, s):
        xi = xl.clone().requires_grad_(True)
        ai = al.clone()

        def one(xi=xi, ai=ai, sl=sl, n=n):
            sl.grad = None
            xi.grad = None
            xp, ap, losses = level(xi, ai, sl, n)
            (xp.float().sum() + ap.float().sum() + losses[2] + losses[3]).backward()

        ms = timed(one)
        out["levels"].append({"N": n, "K": k, "ms_fwd_bwd": round(ms, 4), "graphs_per_s": round(B / ms * 1e3)})
        with torch.no_grad():
            xl, al, _ = level(xi.detach(), ai, sl.detach(), n)
    ms = timed(chain)
    out["chain_ms_fwd_bwd"] = round(ms, 4)
    out["chain_graphs_per_s"] = round(B / ms * 1e3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
