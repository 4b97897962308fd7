"""CPU, world_size 2, gloo: the sharding / routing logic of tgp_b200.distributed with the oracle's CPU operators
as the local kernels.  The sharded result (rank-order concatenation) must equal the single-process oracle:
bit-exact indices and edge order, rtol 1e-6 weights."""
import io
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyg_shim as pyg
from oracle import ref_path as R
from tgp_b200 import distributed as D

EPS = 1e-8


class OracleOps:
    @staticmethod
    def filter_relabel(edge_index, edge_weight, node_index, num_nodes, remove_self_loops):
        ei, w = pyg.subgraph(node_index, edge_index, edge_weight, relabel_nodes=True, num_nodes=num_nodes)
        return R.postprocess_adj_pool_sparse(ei, w, node_index.numel(), remove_self_loops=remove_self_loops)

    @staticmethod
    def coalesce(edge_index, edge_weight, cluster_index, num_nodes, num_clusters, reduce_op, remove_self_loops,
                 filter_tiny):
        ei, w = pyg.coalesce(cluster_index[edge_index], edge_weight, num_nodes=num_clusters, reduce=reduce_op)
        if remove_self_loops:
            ei, w = pyg.remove_self_loops(ei, w)
        if filter_tiny and w is not None and w.numel() > 0:
            m = w.abs() > EPS
            ei, w = ei[:, m], w[m]
        return ei, w

    @staticmethod
    def degree_accumulate(row, w, K, rows_sorted=False):
        w = torch.ones(row.numel()) if w is None else w
        return torch.zeros(max(K, 1)).scatter_add_(0, row, w)

    @staticmethod
    def degree_bwd_accumulate(edge_index, w, deg, grad_out, K, rows_sorted=False):
        # d loss / d dinv[v] = sum_{row = v} g w dinv[col] + sum_{col = v} g w dinv[row]   
        dinv = deg.clamp(min=EPS).pow(-0.5)
        t = grad_out * w
        return (torch.zeros(max(K, 1)).scatter_add_(0, edge_index[0], t * dinv[edge_index[1]])
                + torch.zeros(max(K, 1)).scatter_add_(0, edge_index[1], t * dinv[edge_index[0]]))

    @staticmethod
    def degree_bwd_apply(edge_index, deg, grad_out, grad_dinv, K):
        dinv = deg.clamp(min=EPS).pow(-0.5)
        gdeg = torch.where(deg >= EPS, -0.5 * grad_dinv * dinv ** 3, torch.zeros_like(deg))
        return grad_out * dinv[edge_index[0]] * dinv[edge_index[1]] + gdeg[edge_index[0]]

    @staticmethod
    def segment_sum(x, cluster_index, K):
        return torch.zeros(K, x.size(1)).index_add_(0, cluster_index, x)

    @staticmethod
    def degree_apply(edge_index, w, deg, K):
        w = torch.ones(edge_index.size(1)) if w is None else w
        dinv = deg.clamp(min=EPS).pow(-0.5)
        return w * dinv[edge_index[0]] * dinv[edge_index[1]]

    @staticmethod
    def max_accumulate(row, w, batch_pooled, G):
        return torch.zeros(max(G, 1)).scatter_reduce_(0, batch_pooled[row], w.abs(), "amax", include_self=True)

    @staticmethod
    def max_apply(row, w, batch_pooled, mx, G):
        mx = torch.where(mx == 0, torch.ones_like(mx), mx)
        return w / mx[batch_pooled[row]]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _graph(seed, n=400, e=3000):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    ei = ei[:, torch.argsort(ei[0] * n + ei[1], stable=True)]
    ew = torch.rand(e, generator=g) + 0.5
    ew[5] = 0.0
    return n, ei, ew, g


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, ei, ew, g = _graph(0)
        # ---- kept-node connect, all normalisations
        score = torch.randn(n, generator=g)
        batch = torch.sort(torch.randint(0, 5, (n,), generator=g))[0]
        so = R.topk_select(score, None, 0.5, batch)
        bp = R.reduce_batch(so, batch)
        ei_l, ew_l = D.shard_edges(ei, ew, rank, world)
        eo, wo, off, tot = D.sharded_kept_node_connect(
            ei_l, ew_l, so.node_index, n, degree_norm=True, edge_weight_norm=True, batch_pooled=bp, ops=OracleOps)
        outs = [None] * world
        dist.all_gather_object(outs, (eo, wo, off, tot))
        # ---- cluster connect
        cluster = torch.randint(0, 97, (n,), generator=g)
        ec, wc, rng = D.sharded_cluster_connect(ei_l, ew_l, cluster, 97, degree_norm=True, ops=OracleOps)
        outs2 = [None] * world
        dist.all_gather_object(outs2, (ec, wc, rng))
        # ---- unweighted cluster connect + loss combination
        ec2, wc2, _ = D.sharded_cluster_connect(ei_l, None, cluster, 97, ops=OracleOps)
        outs3 = [None] * world
        dist.all_gather_object(outs3, (ec2, wc2))
        losses = D.combine_losses({"cut_loss": torch.tensor(float(rank + 1)), "link_loss": torch.tensor(3.0 + rank)},
                                  local_graphs=2 + rank)
        # ---- mean coalesce, node-sharded feature reduce (reduce_scatter of [K, F] partials), gradients
        ec4, wc4, _ = D.sharded_cluster_connect(ei_l, ew_l, cluster, 97, reduce_op="mean", ops=OracleOps)
        outs4 = [None] * world
        dist.all_gather_object(outs4, (ec4, wc4))
        x = torch.randn(n, 6, generator=g)
        lo, hi = D.even_ranges(n, world)[rank]
        outs5 = [None] * world
        for op in ("sum", "mean"):
            xp, rows = D.sharded_cluster_reduce(x[lo:hi], cluster[lo:hi], 97, reduce_op=op, ops=OracleOps)
            got = [None] * world
            dist.all_gather_object(got, (xp, rows))
            outs5[0 if op == "sum" else 1] = got
        ew_g = ew_l.clone().requires_grad_(True)
        _, wg, _, _ = D.sharded_kept_node_connect(ei_l, ew_g, so.node_index, n, degree_norm=True, ops=OracleOps)
        (wg * torch.arange(1, wg.numel() + 1)).sum().backward()
        outs6 = [None] * world
        dist.all_gather_object(outs6, (ew_g.grad, wg.numel()))
        if rank == 0:
            buf = io.BytesIO()  # plain bytes: tensors in a SimpleQueue travel by fd passing, which dies with the worker
            torch.save((outs, outs2, outs3, losses, outs4, outs5, outs6), buf)
            q.put(buf.getvalue())
    finally:
        dist.destroy_process_group()


def test_edge_sharded_connect_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs, outs2, outs3, losses, outs4, outs5, outs6 = torch.load(io.BytesIO(q.get()), weights_only=False)
    for p in procs:
        p.join(600)
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"

    n, ei, ew, g = _graph(0)
    score = torch.randn(n, generator=g)
    batch = torch.sort(torch.randint(0, 5, (n,), generator=g))[0]
    so = R.topk_select(score, None, 0.5, batch)
    bp = R.reduce_batch(so, batch)
    e_ref, w_ref = R.sparse_connect_so(ei, so, edge_weight=ew, batch_pooled=bp, degree_norm=True, edge_weight_norm=True)
    e_cat = torch.cat([o[0] for o in outs], 1)
    w_cat = torch.cat([o[1] for o in outs])
    assert torch.equal(e_cat, e_ref)
    torch.testing.assert_close(w_cat, w_ref, rtol=1e-6, atol=1e-7)
    assert outs[0][2] == 0 and outs[1][2] == outs[0][0].size(1) and outs[0][3] == e_ref.size(1)

    cluster = torch.randint(0, 97, (n,), generator=g)
    so_c = R.OracleSelectOutput(cluster_index=cluster, num_supernodes=97)
    e_ref, w_ref = R.sparse_connect_so(ei, so_c, edge_weight=ew, degree_norm=True)
    e_cat = torch.cat([o[0] for o in outs2], 1)
    w_cat = torch.cat([o[1] for o in outs2])
    assert torch.equal(e_cat, e_ref)
    torch.testing.assert_close(w_cat, w_ref, rtol=1e-6, atol=1e-7)
    for (eo, _, (lo, hi)) in outs2:  # each rank owns a contiguous coarse-row range
        assert eo.numel() == 0 or (int(eo[0].min()) >= lo and int(eo[0].max()) < hi)
    e_ref2, w_ref2 = R.sparse_connect_so(ei, so_c)
    assert w_ref2 is None and all(o[1] is None for o in outs3)
    assert torch.equal(torch.cat([o[0] for o in outs3], 1), e_ref2)

    e_ref4, w_ref4 = R.sparse_connect_so(ei, so_c, edge_weight=ew, reduce_op="mean")
    assert torch.equal(torch.cat([o[0] for o in outs4], 1), e_ref4)
    torch.testing.assert_close(torch.cat([o[1] for o in outs4]), w_ref4, rtol=1e-6, atol=1e-7)
    x = torch.randn(n, 6, generator=g)
    ones = R.OracleSelectOutput(cluster_index=cluster, num_supernodes=97)
    for j, op in enumerate(("sum", "mean")):
        ref = R.base_reduce(x, ones)[0] if op == "sum" else R.aggr_reduce(x, ones, op="mean")[0]
        got = torch.cat([o[0] for o in outs5[j]])
        torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
        assert outs5[j][0][1][0] == 0 and outs5[j][-1][1][1] == 97
    # gradient of the sharded degree normalisation (all-reduce of the [K] partials in the backward)
    ew_full = ew.clone().requires_grad_(True)
    _, w_full = R.sparse_connect_so(ei, so, edge_weight=ew_full, degree_norm=True)
    off = 0
    coef = torch.zeros(w_full.numel())
    for gr, cnt in outs6:  # every rank weighted ITS output slice with 1..cnt
        coef[off:off + cnt] = torch.arange(1, cnt + 1).float()
        off += cnt
    (w_full * coef).sum().backward()
    # (fp32 sums of terms weighted up to ~700 with cancellation: the bound is relative to the largest gradient)
    torch.testing.assert_close(torch.cat([o[0] for o in outs6]), ew_full.grad, rtol=1e-4,
                               atol=1e-5 * float(ew_full.grad.abs().max()))

    # losses: weighted batch mean, and sqrt of the summed squares for the global Frobenius norm
    torch.testing.assert_close(losses["cut_loss"], torch.tensor((1.0 * 2 + 2.0 * 3) / 5))
    torch.testing.assert_close(losses["link_loss"], torch.tensor((9.0 + 16.0) ** 0.5))


def test_partition_helpers():
    assert D.even_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    x = torch.arange(12.0).view(6, 2)
    batch = torch.tensor([0, 0, 1, 1, 2, 2])
    ei = torch.tensor([[0, 1, 2, 3, 4, 5], [1, 0, 3, 2, 5, 4]])
    xs, es, bs, _, n_off, g_off = D.shard_graph_batch(x, ei, batch, rank=1, world=2)
    assert n_off == 4 and g_off == 2 and torch.equal(bs, torch.tensor([0, 0]))
    assert torch.equal(es, torch.tensor([[0, 1], [1, 0]])) and torch.equal(xs, x[4:])
