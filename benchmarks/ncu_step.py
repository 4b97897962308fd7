"""A few EAGER steps of one bench workload (no CUDA graph, no timing loops): the thing to run under ncu.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python benchmarks/ncu_step.py c4
    ncu --set full --clock-control none --import-source on -k regex:k_bucket_tiles -c 1 -o rep python benchmarks/ncu_step.py c4
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "torch-geometric-pool_b200"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    if name in bench.DENSE:
        w = bench.DENSE[name]
        st = bench.DenseStep(w, w["B"], dev, 1000)
        for _ in range(steps):
            st.run()
    else:
        # reuse the bench's input construction and eager step; stop after the first measurement hook
        args = types.SimpleNamespace(steps=steps, warmup=1, workload=name)
        ctx = types.SimpleNamespace(args=args, dev=dev, world=1, rank=0, local_rank=0, dist=None)

        class Stop(Exception):
            pass

        calls = {"n": 0}

        def long_run(fn, **kw):
            for _ in range(steps):
                fn()
            torch.cuda.synchronize()
            raise Stop

        ctx.long_run = long_run
        try:
            bench.run_sparse(ctx, name)
        except Stop:
            pass
    torch.cuda.synchronize()
    print(f"ncu_step {name}: done")


if __name__ == "__main__":
    main()
