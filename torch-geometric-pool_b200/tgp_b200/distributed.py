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
    def degree_accumulate(row, w, num_clusters):
        from . import _lib as L

        deg = torch.empty(max(num_clusters, 1), dtype=torch.float32, device=row.device)
        L.call("tgpb200_degree_accumulate", L.ptr(row.contiguous()), L.ptr(w), row.numel(), num_clusters, L.ptr(deg),
               L.stream())
        return deg

    @staticmethod
    def degree_apply(edge_index, w, deg, num_clusters):
        from . import _lib as L

        out = torch.empty(edge_index.size(1), dtype=torch.float32, device=edge_index.device)
        L.call("tgpb200_degree_apply", L.ptr(edge_index[0].contiguous()), L.ptr(edge_index[1].contiguous()), L.ptr(w),
               L.ptr(deg), edge_index.size(1), num_clusters, EPS, L.ptr(out), L.stream())
        return out

    @staticmethod
    def max_accumulate(row, w, batch_pooled, num_graphs):
        from . import _lib as L

        mx = torch.empty(max(num_graphs, 1), dtype=torch.float32, device=row.device)
        L.call("tgpb200_weight_max_accumulate", L.ptr(row.contiguous()), L.ptr(w), L.ptr(batch_pooled.contiguous()),
               row.numel(), num_graphs, L.ptr(mx), L.stream())
        return mx

    @staticmethod
    def max_apply(row, w, batch_pooled, mx, num_graphs):
        from . import _lib as L

        out = torch.empty_like(w)
        L.call("tgpb200_weight_max_apply", L.ptr(row.contiguous()), L.ptr(w), L.ptr(batch_pooled.contiguous()),
               L.ptr(mx), row.numel(), num_graphs, L.ptr(out), L.stream())
        return out


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
    group=None,
    ops=CudaOps,
):
    """Kept-node connect (tgp/connect/base_conn.py:79-82 + tgp/utils/ops.py:370-417) over an edge-sharded graph.

    Returns ``(edge_index_out_local, edge_weight_out_local, offset, total)``: the rank's slice of the global
    output, which is the rank-order concatenation of the slices (= the single-GPU output, bit-exact indices).
    """
    K = node_index.numel()
    ei, w = ops.filter_relabel(edge_index_local, edge_weight_local, node_index, num_nodes, remove_self_loops)
    world = dist.get_world_size(group)
    counts = torch.zeros(world, dtype=torch.long, device=ei.device)
    counts[dist.get_rank(group)] = ei.size(1)
    dist.all_reduce(counts, group=group)  # = all-gather of one int64 per rank
    offset = int(counts[: dist.get_rank(group)].sum().item())
    total = int(counts.sum().item())
    if degree_norm:
        deg = ops.degree_accumulate(ei[0], w, K)
        dist.all_reduce(deg, group=group)
        w = ops.degree_apply(ei, w, deg, K)
    if edge_weight_norm and w is not None:
        if batch_pooled is None:
            raise AssertionError("edge_weight_norm=True but batch_pooled=None.")
        G = num_graphs if num_graphs is not None else int(batch_pooled.max().item()) + 1
        mx = ops.max_accumulate(ei[0], w, batch_pooled, G)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        w = ops.max_apply(ei[0], w, batch_pooled, mx, G)
    return ei, w, offset, total


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
    2. all-to-all: each partial edge goes to the rank owning its coarse row (contiguous row ranges);
    3. merge-coalesce of the received lists in rank order, then the self-loop / tiny-weight filters.
    Returns ``(edge_index_out_local, edge_weight_out_local, (row_lo, row_hi))``: the rank's rows of the global
    output; rank-order concatenation equals the single-GPU result (bit-exact indices; weights equal up to the
    association order of the fp32 sums, which is fixed: shard by shard).
    """
    if reduce_op not in ("sum", "min", "max", "mul"):
        raise ValueError("sharded_cluster_connect supports sum / min / max / mul (mean needs the run lengths)")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    N, K = cluster_index.numel(), num_clusters
    dev = edge_index_local.device
    ei, w = ops.coalesce(edge_index_local, edge_weight_local, cluster_index, N, K, reduce_op, False, False)
    ranges = even_ranges(K, world)
    bounds = torch.tensor([lo for lo, _ in ranges] + [K], dtype=torch.long, device=dev)
    cut = torch.searchsorted(ei[0].contiguous(), bounds)  # rows are sorted -> contiguous slices per owner
    send_counts = (cut[1:] - cut[:-1]).to(torch.long)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    n_recv = int(sum(rc))
    rows = torch.empty(n_recv, dtype=torch.long, device=dev)
    cols = torch.empty(n_recv, dtype=torch.long, device=dev)
    dist.all_to_all_single(rows, ei[0].contiguous(), rc, sc, group=group)
    dist.all_to_all_single(cols, ei[1].contiguous(), rc, sc, group=group)
    wr = None
    if w is not None:
        wr = torch.empty(n_recv, dtype=torch.float32, device=dev)
        dist.all_to_all_single(wr, w.contiguous(), rc, sc, group=group)
    ident = torch.arange(K, dtype=torch.long, device=dev)
    eo, wo = ops.coalesce(torch.stack([rows, cols]), wr, ident, K, K, reduce_op, remove_self_loops, True)
    if degree_norm:
        deg = ops.degree_accumulate(eo[0], wo, K)
        dist.all_reduce(deg, group=group)
        wo = ops.degree_apply(eo, wo, deg, K)
    return eo, wo, ranges[rank]
