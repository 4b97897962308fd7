"""Multi-GPU forms of the path (one process per GPU, ``torch.distributed``; NCCL on the GPU box).

Two partitionings:

* **graph mini-batch** (C1-C3): graphs are independent units, so each rank runs the whole Reduce + Connect chain
  on its own contiguous range of graphs with NO data-path collective; only the scalar auxiliary losses are
  combined (``combine_losses``).
* **one large graph, edges sharded** (C4/C5): every rank owns a contiguous range of the edge list (global edge
  order = rank order).  The kept-node connect needs no exchange for the edges themselves (outputs stay sharded,
  global order is rank-order concatenation) and one all-reduce of the ``[K]`` degree / ``[G]`` max partials when a
  normalisation is on.  The cluster connect coalesces locally, routes each partial coarse edge to the rank that owns
  its coarse row (one all-to-all), and merges there, so duplicate keys are combined in rank order (deterministic).

The local operators are injected (``ops``): the default is the CUDA kernel set of this package; the CPU tests run the
same sharding / routing logic over gloo with the oracle's CPU operators.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

EPS = 1e-8


# --------------------------------------------------------------------------- #
# local operator sets
# --------------------------------------------------------------------------- #
class CudaOps:
    """Local operators backed by libtgp_b200.so."""

    @staticmethod
    def filter_relabel(edge_index, edge_weight, node_index, num_nodes, remove_self_loops):
        from . import functional as F_

        return F_.sparse_connect(edge_index, edge_weight, node_index=node_index, num_nodes=num_nodes,
                                 num_supernodes=node_index.numel(), remove_self_loops=remove_self_loops)

    @staticmethod
    def coalesce(edge_index, edge_weight, cluster_index, num_nodes, num_clusters, reduce_op, remove_self_loops,
                 filter_tiny):
        from . import _lib as L

        row, col = edge_index[0].contiguous(), edge_index[1].contiguous()
        E, dev = row.numel(), row.device
        w = None if edge_weight is None else edge_weight.to(torch.float32).contiguous()
        lib = L.load()
        ws = L.workspace(lib.tgpb200_remap_coalesce_workspace_bytes(E, num_clusters), dev)
        count = torch.empty(1, dtype=torch.long, device=dev)
        flags = L.REMOVE_SELF_LOOPS if remove_self_loops else 0
        eps = EPS if filter_tiny else -1.0  # |w| > -1 keeps everything (no tiny-weight filter before the merge)
        L.call("tgpb200_remap_coalesce_count", L.ptr(row), L.ptr(col), L.ptr(w), E, L.ptr(cluster_index.contiguous()),
               num_nodes, num_clusters, L.OPS[reduce_op], flags, eps, L.ptr(count), L.ptr(ws), ws.numel(), L.stream())
        n_out = int(count.item())
        ei = torch.empty((2, n_out), dtype=torch.long, device=dev)
        wo = None if w is None else torch.empty(n_out, dtype=torch.float32, device=dev)
        if n_out > 0:
            L.call("tgpb200_remap_coalesce_emit", E, num_clusters, int(w is not None), flags, eps, L.ptr(ei[0]),
                   L.ptr(ei[1]), L.ptr(wo), None, None, L.ptr(ws), ws.numel(), L.stream())
        return ei, wo

    @staticmethod
    def _norm_ws(E, K, dev):
        from . import _lib as L

        return L.workspace(L.load().tgpb200_edge_norm_workspace_bytes(E, K), dev)

    @staticmethod
    def degree_accumulate(row, w, num_clusters, rows_sorted=False):
        from . import _lib as L

        deg = torch.empty(max(num_clusters, 1), dtype=torch.float32, device=row.device)
        ws = CudaOps._norm_ws(row.numel(), num_clusters, row.device)
        L.call("tgpb200_degree_accumulate", L.ptr(row.contiguous()), L.ptr(w), row.numel(), None, num_clusters,
               int(rows_sorted), L.ptr(deg), L.ptr(ws), ws.numel(), L.stream())
        return deg

    @staticmethod
    def degree_apply(edge_index, w, deg, num_clusters):
        from . import _lib as L

        out = torch.empty(edge_index.size(1), dtype=torch.float32, device=edge_index.device)
        L.call("tgpb200_degree_apply", L.ptr(edge_index[0].contiguous()), L.ptr(edge_index[1].contiguous()), L.ptr(w),
               L.ptr(deg), edge_index.size(1), None, num_clusters, EPS, L.ptr(out), L.stream())
        return out

    @staticmethod
    def degree_bwd_accumulate(edge_index, w, deg, grad_out, num_clusters, rows_sorted=False):
        from . import _lib as L

        E, dev = edge_index.size(1), edge_index.device
        gd = torch.empty(max(num_clusters, 1), dtype=torch.float32, device=dev)
        ws = CudaOps._norm_ws(E, num_clusters, dev)
        L.call("tgpb200_degree_bwd_accumulate", L.ptr(edge_index[0].contiguous()), L.ptr(edge_index[1].contiguous()),
               L.ptr(w), L.ptr(deg), L.ptr(grad_out.contiguous()), E, None, num_clusters, EPS, int(rows_sorted),
               L.ptr(gd), L.ptr(ws), ws.numel(), L.stream())
        return gd

    @staticmethod
    def degree_bwd_apply(edge_index, deg, grad_out, grad_dinv, num_clusters):
        from . import _lib as L

        E = edge_index.size(1)
        gw = torch.empty(E, dtype=torch.float32, device=edge_index.device)
        L.call("tgpb200_degree_bwd_apply", L.ptr(edge_index[0].contiguous()), L.ptr(edge_index[1].contiguous()),
               L.ptr(deg), L.ptr(grad_out.contiguous()), L.ptr(grad_dinv), E, None, num_clusters, EPS, L.ptr(gw),
               L.stream())
        return gw

    @staticmethod
    def max_accumulate(row, w, batch_pooled, num_graphs):
        from . import _lib as L

        mx = torch.empty(max(num_graphs, 1), dtype=torch.float32, device=row.device)
        L.call("tgpb200_weight_max_accumulate", L.ptr(row.contiguous()), L.ptr(w), L.ptr(batch_pooled.contiguous()),
               row.numel(), None, num_graphs, L.ptr(mx), L.stream())
        return mx

    @staticmethod
    def max_apply(row, w, batch_pooled, mx, num_graphs):
        from . import _lib as L

        out = torch.empty_like(w)
        L.call("tgpb200_weight_max_apply", L.ptr(row.contiguous()), L.ptr(w), L.ptr(batch_pooled.contiguous()),
               L.ptr(mx), row.numel(), None, num_graphs, L.ptr(out), L.stream())
        return out

    @staticmethod
    def segment_sum(x, cluster_index, num_clusters):
        """[K, F] partial sums of the local node shard (unit weights)."""
        from . import functional as F_

        nodes = torch.arange(x.size(0), device=x.device)
        return F_.segment_reduce(x, nodes, cluster_index, None, num_clusters, "sum")


# --------------------------------------------------------------------------- #
# partitioning helpers
# --------------------------------------------------------------------------- #
def even_ranges(total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, order-preserving split of ``range(total)`` into ``world`` nearly equal ranges."""
    base, rem = divmod(total, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append((start, start + n))
        start += n
    return out


def shard_edges(edge_index: Tensor, edge_weight: Optional[Tensor], rank: int, world: int):
    lo, hi = even_ranges(edge_index.size(1), world)[rank]
    return edge_index[:, lo:hi].contiguous(), None if edge_weight is None else edge_weight[lo:hi].contiguous()


def shard_graph_batch(x: Tensor, edge_index: Tensor, batch: Tensor, rank: int, world: int,
                      edge_weight: Optional[Tensor] = None):
    """Split a PyG-style block-diagonal batch by contiguous graph ranges.  Returns the rank's
    ``(x, edge_index (re-based to local node ids), batch (re-based to local graph ids), edge_weight,
    node_offset, graph_offset)``; concatenating the ranks' outputs with these offsets reproduces
    ``Batch.from_data_list`` numbering."""
    num_graphs = int(batch.max().item()) + 1 if batch.numel() else 0
    g_lo, g_hi = even_ranges(num_graphs, world)[rank]
    node_mask = (batch >= g_lo) & (batch < g_hi)
    nodes = node_mask.nonzero().view(-1)
    n_lo = int(nodes[0].item()) if nodes.numel() else 0
    n_hi = int(nodes[-1].item()) + 1 if nodes.numel() else 0
    e_mask = (edge_index[0] >= n_lo) & (edge_index[0] < n_hi)
    ei = edge_index[:, e_mask] - n_lo
    ew = None if edge_weight is None else edge_weight[e_mask]
    return x[n_lo:n_hi], ei, batch[n_lo:n_hi] - g_lo, ew, n_lo, g_lo


def combine_losses(losses: Dict[str, Tensor], local_graphs: int, group=None, link_keys=("link_loss",)) -> Dict[str, Tensor]:
    """Batch-mean losses -> global batch mean (weighted by the local number of graphs); the DiffPool link loss
    is ONE Frobenius norm over the whole batch (tgp/utils/losses.py:674-676): square, sum across ranks, sqrt."""
    world = dist.get_world_size(group)
    out = {}
    n = torch.tensor([float(local_graphs)], device=next(iter(losses.values())).device)
    dist.all_reduce(n, group=group)
    for k, v in losses.items():
        if k in link_keys:
            t = v.detach().clone().square()
            dist.all_reduce(t, group=group)
            out[k] = t.sqrt()
        else:
            t = v.detach().clone() * local_graphs
            dist.all_reduce(t, group=group)
            out[k] = t / n.squeeze(0)
    return out if world > 0 else losses


# --------------------------------------------------------------------------- #
# edge-sharded single-graph connect
# --------------------------------------------------------------------------- #
class _ShardedDegreeNorm(torch.autograd.Function):
    """Degree normalisation of an edge-sharded coarse edge list (tgp/utils/ops.py:383-401): the ``[K]`` degree
    partials of the shards are summed with one all-reduce in the forward, the ``[K]`` d(loss)/d(dinv) partials with
    one all-reduce in the backward; everything else is local."""

    @staticmethod
    def forward(ctx, w, edge_index, K, rows_sorted, group, ops):
        deg = ops.degree_accumulate(edge_index[0], w, K, rows_sorted)
        dist.all_reduce(deg, group=group)
        ctx.save_for_backward(w, edge_index, deg)
        ctx.K, ctx.rows_sorted, ctx.group, ctx.ops = K, rows_sorted, group, ops
        return ops.degree_apply(edge_index, w, deg, K)

    @staticmethod
    def backward(ctx, g):
        w, edge_index, deg = ctx.saved_tensors
        if w is None:
            return (None,) * 6
        gd = ctx.ops.degree_bwd_accumulate(edge_index, w, deg, g, ctx.K, ctx.rows_sorted)
        dist.all_reduce(gd, group=ctx.group)
        return ctx.ops.degree_bwd_apply(edge_index, deg, g, gd, ctx.K), None, None, None, None, None


def sharded_kept_node_connect(
    edge_index_local: Tensor,
    edge_weight_local: Optional[Tensor],
    node_index: Tensor,
    num_nodes: int,
    *,
    remove_self_loops: bool = True,
    degree_norm: bool = False,
    edge_weight_norm: bool = False,
    batch_pooled: Optional[Tensor] = None,
    num_graphs: Optional[int] = None,
    rows_sorted: bool = False,
    group=None,
    ops=CudaOps,
):
    """Kept-node connect (tgp/connect/base_conn.py:79-82 + tgp/utils/ops.py:370-417) over an edge-sharded graph.

    Returns ``(edge_index_out_local, edge_weight_out_local, offset, total)``: the rank's slice of the global
    output, which is the rank-order concatenation of the slices (= the single-GPU output, bit-exact indices).
    Differentiable w.r.t. the local edge weights (the degree normalisation all-reduces its ``[K]`` partials in both
    directions).  ``rows_sorted``: the local edge list is sorted by row (sort-free deterministic degree sums).
    """
    K = node_index.numel()
    ei, w = ops.filter_relabel(edge_index_local, edge_weight_local, node_index, num_nodes, remove_self_loops)
    world = dist.get_world_size(group)
    counts = torch.zeros(world, dtype=torch.long, device=ei.device)
    counts[dist.get_rank(group)] = ei.size(1)
    dist.all_reduce(counts, group=group)  # = all-gather of one int64 per rank
    counts = counts.tolist()
    offset, total = int(sum(counts[: dist.get_rank(group)])), int(sum(counts))
    if degree_norm:
        w = _ShardedDegreeNorm.apply(w, ei, K, rows_sorted, group, ops)
    if edge_weight_norm and w is not None:
        if batch_pooled is None:
            raise AssertionError("edge_weight_norm=True but batch_pooled=None.")
        G = num_graphs if num_graphs is not None else int(batch_pooled.max().item()) + 1
        mx = ops.max_accumulate(ei[0], w.detach(), batch_pooled, G)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        w = ops.max_apply(ei[0], w, batch_pooled, mx, G)  # the arg-max gradient routing is not sharded (forward)
    return ei, w, offset, total


def _route_by_row(ei: Tensor, cols: List[Tensor], K: int, group) -> Tuple[Tensor, List[Tensor]]:
    """One exchange: every (row-sorted) partial coarse edge goes to the rank that owns its coarse row.  The per-pair
    counts travel in ONE all-gather (one host read), the edges and their float columns in ONE all-to-all of a packed
    int64 matrix.  Returns the received ``edge_index`` (grouped by source rank, each group row-sorted) and columns."""
    world = dist.get_world_size(group)
    dev = ei.device
    ranges = even_ranges(K, world)
    bounds = torch.tensor([lo for lo, _ in ranges] + [K], dtype=torch.long, device=dev)
    cut = torch.searchsorted(ei[0].contiguous(), bounds)  # rows are sorted -> contiguous slices per owner
    send = (cut[1:] - cut[:-1]).to(torch.long)
    allc = torch.empty(world * world, dtype=torch.long, device=dev)
    dist.all_gather_into_tensor(allc, send, group=group)
    allc = allc.view(world, world).tolist()  # [source][destination]
    rank = dist.get_rank(group)
    sc, rc = allc[rank], [allc[src][rank] for src in range(world)]
    width = 2 + len(cols)
    packed = torch.empty((ei.size(1), width), dtype=torch.long, device=dev)
    packed[:, 0], packed[:, 1] = ei[0], ei[1]
    for j, c in enumerate(cols):
        packed[:, 2 + j] = c.contiguous().view(torch.int32).to(torch.long)
    recv = torch.empty((int(sum(rc)), width), dtype=torch.long, device=dev)
    dist.all_to_all_single(recv, packed, rc, sc, group=group)
    out_cols = [recv[:, 2 + j].to(torch.int32).view(torch.float32) for j in range(len(cols))]
    return recv[:, :2].t().contiguous(), out_cols


def sharded_cluster_connect(
    edge_index_local: Tensor,
    edge_weight_local: Optional[Tensor],
    cluster_index: Tensor,
    num_clusters: int,
    *,
    reduce_op: str = "sum",
    remove_self_loops: bool = True,
    degree_norm: bool = False,
    group=None,
    ops=CudaOps,
):
    """Cluster connect (tgp/connect/base_conn.py:83-89) over an edge-sharded graph.

    1. local remap + coalesce (no filters) -> partial coarse edges, sorted by (row, col);
    2. one all-to-all: each partial edge goes to the rank owning its coarse row (contiguous row ranges);
    3. merge-coalesce of the received lists in rank order, then the self-loop / tiny-weight filters.
    ``mean`` carries (sum, count) partials through the exchange and divides after the merge.
    Returns ``(edge_index_out_local, edge_weight_out_local, (row_lo, row_hi))``: the rank's rows of the global
    output; rank-order concatenation equals the single-GPU result (bit-exact indices; weights equal up to the
    association order of the fp32 sums, which is fixed: shard by shard).
    """
    if reduce_op not in ("sum", "mean", "min", "max", "mul"):
        raise ValueError(f"unknown reduce_op '{reduce_op}'")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    N, K = cluster_index.numel(), num_clusters
    dev = edge_index_local.device
    mean = reduce_op == "mean" and edge_weight_local is not None
    op1 = "sum" if mean else reduce_op
    ei, w = ops.coalesce(edge_index_local, edge_weight_local, cluster_index, N, K, op1, False, False)
    cols = [] if w is None else [w]
    if mean:
        ones = torch.ones(edge_index_local.size(1), dtype=torch.float32, device=dev)
        _, cnt = ops.coalesce(edge_index_local, ones, cluster_index, N, K, "sum", False, False)
        cols.append(cnt)
    recv_ei, recv_cols = _route_by_row(ei, cols, K, group)
    ident = torch.arange(K, dtype=torch.long, device=dev)
    if mean:
        e_sum, w_sum = ops.coalesce(recv_ei, recv_cols[0], ident, K, K, "sum", False, False)
        _, w_cnt = ops.coalesce(recv_ei, recv_cols[1], ident, K, K, "sum", False, False)
        eo, wo = e_sum, w_sum / w_cnt
        keep = wo.abs() > EPS
        if remove_self_loops:
            keep &= eo[0] != eo[1]
        eo, wo = eo[:, keep], wo[keep]
    else:
        eo, wo = ops.coalesce(recv_ei, recv_cols[0] if recv_cols else None, ident, K, K, reduce_op, remove_self_loops,
                              True)
    if degree_norm:
        wo = _ShardedDegreeNorm.apply(wo, eo, K, True, group, ops)
    return eo, wo, even_ranges(K, world)[rank]


def sharded_cluster_reduce(
    x_local: Tensor,
    cluster_index_local: Tensor,
    num_clusters: int,
    *,
    reduce_op: str = "sum",
    group=None,
    ops=CudaOps,
):
    """Sparse reduce (tgp/reduce/base_reduce.py:141-155) of ONE large graph whose NODES are sharded: every rank
    reduces its node shard into a ``[K, F]`` partial (local segment reduce), the partials are combined with one
    ``reduce_scatter`` (sum) so that rank r ends up owning the coarse rows ``[r K/G, (r+1) K/G)`` -- the "NCCL reduce
    of coarse partials" of the single-large-graph mode.  ``mean`` reduce-scatters the member counts alongside.
    Returns ``(x_pool_rows, (row_lo, row_hi))``."""
    if reduce_op not in ("sum", "add", "mean"):
        raise ValueError("sharded_cluster_reduce supports sum / mean (max / min need an all-reduce(max) instead)")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    K, F = num_clusters, x_local.size(1)
    per = (K + world - 1) // world
    part = ops.segment_sum(x_local, cluster_index_local, K)
    mean = reduce_op == "mean"
    width = F + (1 if mean else 0)
    buf = torch.zeros((per * world, width), dtype=torch.float32, device=x_local.device)
    buf[:K, :F] = part
    if mean:
        buf[:K, F] = torch.bincount(cluster_index_local, minlength=K).to(torch.float32)
    out = torch.empty((per, width), dtype=torch.float32, device=x_local.device)
    dist.reduce_scatter_tensor(out, buf, group=group)
    lo, hi = rank * per, min((rank + 1) * per, K)
    out = out[: max(hi - lo, 0)]
    if mean:
        out = out[:, :F] / out[:, F:].clamp(min=1.0)
    return out, (lo, hi)
