#!/usr/bin/env python
"""bench.py -- coarsened graphs/s of the Reduce + Connect hot path (forward + backward).

Workload at N=1 (BASELINE.json configs[1], "C2"): MinCutPooling dense, 64 clusters, batch 512 graphs x 256
nodes x 128 feats, fp32, S^T X + S^T A S + mincut/ortho losses + post-processing, fwd+bwd.
One "step" = one pass of the path over one batch of 512 graphs per GPU (weak scaling: every rank owns its own
512 graphs, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3l1|c4]

Prints ONE JSON line (rank 0).  `value` = graphs/s with inputs resident in HBM; `e2e` = the same metric through
the public API with pinned HOST buffers copied in every step and the losses read back; `roofline` = algorithmic
bytes of the dominant entry point / its CUDA-event duration / MEASURED_PEAKS.json; `cpu_baseline` = the CPU
oracle (restated reference path) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "torch-geometric-pool_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

DOMINANT_KERNEL = "k_dense_fwd_fused"

WORKLOADS = {
    # name: (B, N, K, F, dtype, adjacency density, pooler)
    "c2": dict(B=512, N=256, K=64, F=128, dtype="f32", p=0.05, pooler="mincut",
               desc="MinCutPooling dense, K=64, 512 graphs x 256 nodes x 128 feats, fp32, fwd+bwd"),
    "c3l1": dict(B=1024, N=512, K=256, F=256, dtype="bf16", p=0.02, pooler="diff",
                 desc="DiffPool dense level 1, K=256, 1024 graphs x 512 nodes x 256 feats, bf16, fwd+bwd"),
}


SPARSE_WORKLOADS = {
    "c1": dict(kind="topk", graphs=128, n_lo=24, n_hi=36, p=0.1, F=64,
               desc="TopK ratio 0.5 + sum reduce + kept-node connect, 128 ER graphs (~30 nodes, 64 feats), fwd+bwd"),
    "c4": dict(kind="cluster", N=1_000_000, E=20_000_000, F=128,
               desc="cluster connect (remap + coalesce sum + self-loop removal) + mean reduce, 1M nodes / 20M edges "
                    "power-law, 128 feats, matching-style cluster map (K ~ 0.55 N), fwd+bwd"),
    "c5": dict(kind="topk_big", N=16_000_000, E=400_000_000, F=128,
               desc="TopK 50% kept-node connect + sum reduce on one 16M-node / 400M-edge graph, 128 feats, fwd"),
}


def powerlaw_graph(n, e, device, seed):
    """Chung-Lu style power-law graph (exponent 2.3), symmetric, no self loops, sorted by (row, col)."""
    g = torch.Generator(device=device).manual_seed(seed)
    w = (torch.arange(1, n + 1, device=device, dtype=torch.float32)) ** (-1.0 / 1.3)
    cdf = torch.cumsum(w.double(), 0)
    cdf = (cdf / cdf[-1]).float()
    half = e // 2
    src = torch.searchsorted(cdf, torch.rand(half, device=device, generator=g)).clamp_(max=n - 1)
    dst = torch.randint(0, n, (half,), device=device, generator=g)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    perm = torch.randperm(n, device=device, generator=g)  # hubs are spread over the id range
    src, dst = perm[src], perm[dst]
    row = torch.cat([src, dst])
    col = torch.cat([dst, src])
    order = torch.argsort(row * n + col)
    return torch.stack([row[order], col[order]])


def run_sparse(args, name):
    import tgp_b200 as T
    from tgp_b200 import _lib

    w = SPARSE_WORKLOADS[name]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    if w["kind"] == "topk":
        gc = torch.Generator().manual_seed(0)
        xs, eis, bs, off = [], [], [], 0
        for gi in range(w["graphs"]):
            n = int(torch.randint(w["n_lo"], w["n_hi"] + 1, (1,), generator=gc))
            up = torch.triu(torch.rand(n, n, generator=gc) < w["p"], 1)
            r, c = up.nonzero(as_tuple=True)
            eis.append(torch.cat([torch.stack([r, c]), torch.stack([c, r])], 1) + off)
            bs.append(torch.full((n,), gi))
            off += n
        ei = torch.cat(eis, 1)
        ei = ei[:, torch.argsort(ei[0] * off + ei[1])].to(dev)
        batch = torch.cat(bs).to(dev)
        N, units = off, w["graphs"]
    else:
        N = w["N"]
        ei = powerlaw_graph(N, w["E"], dev, 0)
        batch = None
        units = 1
    E = ei.size(1)
    F = w["F"]
    x = torch.randn(N, F, device=dev, generator=g).requires_grad_(w["kind"] != "topk_big")
    ew = (torch.rand(E, device=dev, generator=g) + 0.5).requires_grad_(w["kind"] != "topk_big")
    if w["kind"] == "cluster":
        # matching-style map: a random pairing along a permutation, ~45% of the nodes paired -> K ~ 0.55 N
        perm = torch.randperm(N, device=dev, generator=g)
        paired = int(0.9 * N) // 2 * 2
        cl = torch.empty(N, dtype=torch.long, device=dev)
        cl[perm[:paired]] = torch.arange(paired // 2, device=dev).repeat_interleave(2)
        cl[perm[paired:]] = torch.arange(paired // 2, paired // 2 + N - paired, device=dev)
        K = paired // 2 + N - paired
        so = T.SelectOutput(cluster_index=cl, num_nodes=N, num_supernodes=K)
        reduce_op = "mean"
    else:
        score = torch.tanh(torch.randn(N, device=dev, generator=g))
        if batch is None:
            node_index = torch.topk(score, N // 2).indices
        else:  # per-graph top half (selection is upstream of the path and not timed)
            order = torch.argsort(batch * 4.0 - score)  # graph asc, score desc
            ptr = torch.zeros(int(batch.max()) + 2, dtype=torch.long, device=dev)
            ptr[1:] = torch.bincount(batch).cumsum(0)
            rank_in = torch.arange(N, device=dev) - ptr[batch[order]]
            keep = rank_in < ((ptr[1:] - ptr[:-1] + 1) // 2)[batch[order]]
            node_index = order[keep]
        K = node_index.numel()
        so = T.SelectOutput(node_index=node_index, num_nodes=N, cluster_index=torch.arange(K, device=dev),
                            num_supernodes=K, weight=score[node_index])
        reduce_op = "sum"
    T.functional.csr_of(so)  # the CSR is cached on the SelectOutput (pre-coarsened pipelines reuse it)
    train = w["kind"] != "topk_big"

    def step():
        x.grad = None
        ew.grad = None
        xp, eo, wo, _ = T.sparse_pool(x, ei, so, edge_weight=ew, batch=batch, reduce_op=reduce_op)
        if train:
            torch.autograd.backward([xp, wo], [torch.ones_like(xp), torch.ones_like(wo)])
        return eo.size(1)

    for _ in range(max(args.warmup, 3)):
        e_out = step()
    torch.cuda.synchronize()
    l0 = _lib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    hbm, _, src = peaks()
    # algorithmic bytes: edges in (2*8+4) + edges out (2*8+4) + index map + x in + x_pool out, x2 for bwd feats
    alg = 20 * E + 20 * e_out + 8 * so.node_index.numel() + 4 * F * (so.node_index.numel() + K)
    if train:
        alg += 4 * F * (N + K) + 4 * E + 4 * e_out
    line = {"metric": "coarsened graphs/s (Reduce+Connect fwd+bwd)", "value": units / (ms * 1e-3), "unit": "graphs/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "N": N, "E": E, "K": K, "E_out": e_out},
            "gpu_launches": _lib.kernel_launches() - l0,
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / hbm, "traffic": None, "kernel": "whole step",
                         "algorithmic_bytes": alg, "peak_source": src}}
    print(json.dumps(line), flush=True)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "measured"
    return 6650.0, 1400.0, "fallback"


def algorithmic_bytes(w, direction):
    B, N, K, F = w["B"], w["N"], w["K"], w["F"]
    es = 4 if w["dtype"] == "f32" else 2
    fwd = es * B * (N * N + N * K + N * F + K * F + K * K)
    # backward: re-read A, S, X and the two upstream gradients, write dS, dX
    bwd = es * B * (N * N + N * K + N * F + K * F + K * K + N * K + N * F)
    return {"fwd": fwd, "bwd": bwd, "step": fwd + bwd}[direction]


def make_inputs(w, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    B, N, K, F = w["B"], w["N"], w["K"], w["F"]
    dt = torch.float32 if w["dtype"] == "f32" else torch.bfloat16
    a = (torch.rand(B, N, N, generator=g) < w["p"]).float()
    a = torch.triu(a, 1)
    a = a + a.transpose(1, 2)
    s = torch.softmax(torch.randn(B, N, K, generator=g), -1)
    x = torch.randn(B, N, F, generator=g)
    return a.to(dt), s.to(dt), x.to(dt)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def oracle_step(w, a, s, x):
    """Reference CPU path (restated; PyG unavailable offline): fwd + bwd of the same step."""
    from oracle import ref_path as R

    s = s.clone().requires_grad_(True)
    x = x.clone().requires_grad_(True)
    if w["pooler"] == "mincut":
        xp, ap, loss = R.mincut_pool(x, a, s)
    else:
        xp, ap, loss = R.diff_pool(x, a, s)
    (xp.sum() + ap.sum() + sum(loss.values())).backward()
    return float(sum(v.detach() for v in loss.values()))


def cpu_baseline(w, budget_s=12.0, sample_graphs=32, warmup=1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ws = dict(w, B=sample_graphs)
    a, s, x = make_inputs(ws, "cpu", 123)
    a, s, x = a.float(), s.float(), x.float()
    for _ in range(warmup):
        oracle_step(ws, a, s, x)
    times = []
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end and len(times) < 30:
        t0 = time.perf_counter()
        oracle_step(ws, a, s, x)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": sample_graphs / med, "unit": "graphs/s", "cores": cores, "kind": "port",
            "sample": f"{sample_graphs} graphs of the workload shape, fp32, fwd+bwd, median of {len(times)} iterations; "
                      "reference CPU path restated in oracle/ (PyG unavailable offline)"}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = min(w["B"], 512)  # the full C2 batch per step (bounded for the larger workloads)
    ws = dict(w, B=sample)
    a, s, x = make_inputs(ws, "cpu", 123)
    a, s, x = a.float(), s.float(), x.float()
    for _ in range(args.warmup):
        oracle_step(ws, a, s, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(ws, a, s, x)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    line = {
        "impl": "reference", "metric": "coarsened graphs/s (Reduce+Connect fwd+bwd)", "value": value, "unit": "graphs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "sample_graphs_per_step": sample, "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": "graphs/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} graphs of the workload shape per step, fp32, fwd+bwd; reference CPU path "
                                   "restated in oracle/ (torch_geometric / torch_scatter are not installable offline)"},
        "e2e": {"value": value, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + sorted(SPARSE_WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload in SPARSE_WORKLOADS:
        run_sparse(args, args.workload)
        return
    w = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, w)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback); use --impl reference")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    import tgp_b200 as T
    from tgp_b200 import _lib

    a_h, s_h, x_h = make_inputs(w, "cpu", 1000 + rank)
    a_h, s_h, x_h = a_h.pin_memory(), s_h.pin_memory(), x_h.pin_memory()
    a = a_h.to(dev)
    s = s_h.to(dev).requires_grad_(True)
    x = x_h.to(dev).requires_grad_(True)
    B, N, K, F = w["B"], w["N"], w["K"], w["F"]
    dt = a.dtype
    g_xp = torch.ones(B, K, F, dtype=dt, device=dev)
    g_ap = torch.ones(B, K, K, dtype=dt, device=dev)
    g_l = torch.zeros(4, dtype=torch.float32, device=dev)
    pool = T.mincut_pool if w["pooler"] == "mincut" else T.diff_pool
    if w["pooler"] == "mincut":
        g_l[0] = 1.0
        g_l[1] = 1.0
    else:
        g_l[2] = 1.0
        g_l[3] = 1.0

    from tgp_b200 import functional as F_

    def step(a_, s_, x_):
        # loss = x_pool.sum() + adj_pool.sum() + aux losses  (examples/time_and_mem_test.py:380-383, plus the
        # pooled adjacency so the connect backward is exercised), expressed as explicit upstream gradients
        s_.grad = None
        x_.grad = None
        kind = F_.LOSS_MINCUT if w["pooler"] == "mincut" else F_.LOSS_DIFFPOOL
        xp, ap, losses = F_.dense_pool(x_, a_, s_, remove_self_loops=True, degree_norm=True, adj_transpose=True,
                                       loss_kind=kind, ent_div=float(B * N))
        torch.autograd.backward([xp, ap, losses], [g_xp, g_ap, g_l])
        return losses

    # --- event hooks on the two entry points (roofline of the dominant one)
    ev = {"tgpb200_dense_pool_fwd": [], "tgpb200_dense_pool_bwd": []}
    raw_call = _lib.call

    def timed_call(name, *cargs):
        if name in ev and ev["on"]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            raw_call(name, *cargs)
            e1.record()
            ev[name].append((e0, e1))
        else:
            raw_call(name, *cargs)

    ev["on"] = False
    _lib.call = timed_call
    F_.L.call = timed_call

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(a, s, x)
    barrier()

    # --- timed region 1: inputs resident in HBM.  nvidia-smi samples every 100 ms while a step takes < 1 ms, so the
    # sampler also covers an untimed soak of the same step right before the timed region (same load, same clocks).
    with ClockSampler(local_rank) as clk:
        t_soak = time.perf_counter() + 1.2
        while time.perf_counter() < t_soak:
            for _ in range(20):
                step(a, s, x)
            torch.cuda.synchronize()
        ev["on"] = True
        _lib.time_kernel(DOMINANT_KERNEL)
        launches0 = _lib.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            step(a, s, x)
        e1.record()
        barrier()
        launches = _lib.kernel_launches() - launches0
        ev["on"] = False
        dom_ms, dom_n = _lib.kernel_time_ms()
        _lib.time_kernel(None)
        t_soak = time.perf_counter() + 0.3
        while time.perf_counter() < t_soak:
            step(a, s, x)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_step = float(t_ms.item()) / args.steps
    fwd_ms = statistics.mean(p[0].elapsed_time(p[1]) for p in ev["tgpb200_dense_pool_fwd"])
    bwd_ms = statistics.mean(p[0].elapsed_time(p[1]) for p in ev["tgpb200_dense_pool_bwd"])

    # --- timed region 2: end to end through the public API from pinned host buffers.  Every step copies its own
    # inputs host -> device and reads its result back; the copy of step i+1 runs on a side stream while step i
    # computes (an ordinary double-buffered input pipeline), so the step time is max(copy, compute), not the sum.
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)

    # two resident device buffer sets (no allocator traffic inside the loop); a set is overwritten only after the
    # step that consumed it has finished (event recorded on the compute stream)
    bufs = [tuple(torch.empty_like(t, device=dev) for t in (a_h, s_h, x_h)) for _ in range(2)]
    consumed = [None, None]

    def stage_inputs(j):
        with torch.cuda.stream(copy_stream):
            if consumed[j] is not None:
                copy_stream.wait_event(consumed[j])
            for dst, src in zip(bufs[j], (a_h, s_h, x_h)):
                dst.copy_(src, non_blocking=True)
            done = torch.cuda.Event()
            done.record(copy_stream)
        return done

    def e2e_loop(n):
        ready = stage_inputs(0)
        for i in range(n):
            j = i & 1
            main_stream.wait_event(ready)
            if i + 1 < n:
                ready = stage_inputs(j ^ 1)
            a_d, s_d, x_d = bufs[j]
            losses = step(a_d, s_d.detach().requires_grad_(True), x_d.detach().requires_grad_(True))
            consumed[j] = torch.cuda.Event()
            consumed[j].record(main_stream)
            losses.cpu()  # device -> host read of the step's result

    e2e_loop(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 10))
    e0.record()
    e2e_loop(e2e_steps)
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item()) / e2e_steps
    h2d = a_h.numel() * a_h.element_size() + s_h.numel() * s_h.element_size() + x_h.numel() * x_h.element_size()

    if rank == 0:
        hbm, tf, src = peaks()
        fwd_bytes = algorithmic_bytes(w, "fwd")
        bwd_bytes = algorithmic_bytes(w, "bwd")
        # Dominant kernel = the fused forward main kernel (one pass over A, X, S per graph); its algorithmic bytes are
        # the forward's: 4 B (N^2 + NK + NF + KF + K^2).  Timed live with CUDA events on its launch
        # stream inside the timed region (tgpb200_debug_time_kernel).  When the shape does not take the fused kernel
        # the roofline falls back to the slower of the two entry points.
        if dom_n > 0:
            dom, dom_bytes = DOMINANT_KERNEL, fwd_bytes
        else:
            dom, dom_ms, dom_bytes = ("tgpb200_dense_pool_bwd", bwd_ms, bwd_bytes) if bwd_ms >= fwd_ms else (
                "tgpb200_dense_pool_fwd", fwd_ms, fwd_bytes)
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.isfile(tpath) and dom_n > 0 and args.workload == "c2":
            traffic = json.load(open(tpath)).get(DOMINANT_KERNEL)
        line = {
            "metric": "coarsened graphs/s (Reduce+Connect fwd+bwd)",
            "value": world * B / (ms_step * 1e-3),
            "unit": "graphs/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": w["dtype"],
            "data": "synthetic",
            "config": {"workload": w["desc"], "graphs_per_gpu": B, "loss": "x_pool.sum()+adj_pool.sum()+aux losses",
                       "l2_policy": f"inputs larger than L2 ({h2d / 1e6:.0f} MB per step vs 126 MB L2), no flush",
                       "parallelism": f"graph-batch sharding x{world}, no data-path collective"},
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "graphs/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 16, "ms_per_step": e2e_ms,
                    "pipeline": "H2D of step i+1 on a copy stream overlaps compute of step i"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "kernel": dom, "kernel_launches_timed": dom_n, "kernel_ms": dom_ms, "algorithmic_bytes": dom_bytes,
                         "peak_source": f"{src} (MEASURED_PEAKS.json hbm_gbs)",
                         "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
                         "step_frac": (fwd_bytes + bwd_bytes) / (ms_step * 1e-3) / 1e9 / hbm},
            "clocks": clk.summary(),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(w)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
