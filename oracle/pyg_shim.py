"""Restatement (torch, CPU) of the PyG / torch_scatter utilities the hot path calls.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference (tgp v1.0.1) delegates all sparse arithmetic to
``torch_geometric>=2.6,<3`` (pyproject.toml:38) and ``torch_scatter`` 2.1.2
(pre-requirements.txt:2).  Neither is vendored under /root/reference nor
installed in this image, so each utility is restated here from the published
semantics of that pinned range.  Call sites in the reference:

* ``scatter``            tgp/reduce/base_reduce.py:147, tgp/utils/ops.py:388,409
* ``coalesce``           tgp/connect/base_conn.py:87
* ``subgraph``           tgp/connect/base_conn.py:80
* ``maybe_num_nodes``    tgp/connect/base_conn.py:78
* ``remove_self_loops``  tgp/utils/ops.py:371
* ``unbatch`` / ``unbatch_edge_index``  tgp/reduce/base_reduce.py:171, tgp/connect/dense_conn.py:183-193
* ``to_dense_adj`` / ``to_dense_batch`` tgp/src.py:434,448
* ``topk`` / ``softmax`` tgp/select/topk_select.py:192,194
* PyG ``Aggregation`` (sum/mean/max/min) tgp/reduce/aggr_reduce.py:29

Every function keeps the argument names of the utility it restates.
"""

from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor


# --------------------------------------------------------------------------- #
# scatter
# --------------------------------------------------------------------------- #
def _expand_index(index: Tensor, src: Tensor, dim: int) -> Tensor:
    """Broadcast a 1-D ``index`` against ``src`` along ``dim``."""
    dim = src.dim() + dim if dim < 0 else dim
    shape = [1] * src.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(src)


def scatter(
    src: Tensor,
    index: Tensor,
    dim: int = 0,
    dim_size: Optional[int] = None,
    reduce: str = "sum",
) -> Tensor:
    """``torch_geometric.utils.scatter`` on the CPU code path.

    * ``sum``/``add``: zero-initialised ``scatter_add_`` (sequential in index order on CPU).
    * ``mean``: sum divided by ``clamp(count, min=1)``.
    * ``min``/``max``: zero-initialised ``scatter_reduce_(amin/amax, include_self=False)``
      -> untouched rows stay 0; the gradient is split evenly among ties.
    * ``mul``: one-initialised product.
    """
    if index.dim() != 1:
        raise ValueError("scatter: index must be one-dimensional")
    dim = src.dim() + dim if dim < 0 else dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.size())
    size[dim] = dim_size

    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(dim, _expand_index(index, src, dim), src)
    if reduce == "mean":
        count = src.new_zeros(dim_size)
        count.scatter_add_(0, index, src.new_ones(src.size(dim)))
        count = count.clamp(min=1)
        out = src.new_zeros(size).scatter_add_(dim, _expand_index(index, src, dim), src)
        shape = [1] * out.dim()
        shape[dim] = -1
        return out / count.view(shape)
    if reduce in ("min", "max", "amin", "amax"):
        return src.new_zeros(size).scatter_reduce_(
            dim, _expand_index(index, src, dim), src, reduce=f"a{reduce[-3:]}", include_self=False
        )
    if reduce == "mul":
        return src.new_ones(size).scatter_reduce_(
            dim, _expand_index(index, src, dim), src, reduce="prod", include_self=True
        )
    raise ValueError(f"scatter: unknown reduce '{reduce}'")


class _ScatterArgExtreme(torch.autograd.Function):
    """torch_scatter 2.1.2 ``scatter_max`` / ``scatter_min`` along dim 0 (CPU rule).

    Forward: extreme per segment, untouched rows 0.  The CPU kernel walks ``src`` in
    order and replaces the running value only on a *strict* improvement, so the arg
    is the FIRST position holding the extreme; backward routes the whole gradient to
    that single position.
    """

    @staticmethod
    def forward(ctx, src: Tensor, index: Tensor, dim_size: int, is_max: bool):
        n = src.size(0)
        size = [dim_size] + list(src.shape[1:])
        idx = _expand_index(index, src, 0)
        out = src.new_zeros(size).scatter_reduce_(0, idx, src, reduce="amax" if is_max else "amin", include_self=False)
        hit = src == out.gather(0, idx) if n > 0 else torch.zeros_like(src, dtype=torch.bool)
        pos = torch.arange(n, device=src.device).view([-1] + [1] * (src.dim() - 1)).expand_as(src)
        cand = torch.where(hit, pos, torch.full_like(pos, n))
        arg = torch.full(size, n, dtype=torch.long, device=src.device)
        arg.scatter_reduce_(0, idx, cand, reduce="amin", include_self=True)
        ctx.save_for_backward(arg)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        (arg,) = ctx.saved_tensors
        size = [ctx.n + 1] + list(grad_out.shape[1:])
        grad = grad_out.new_zeros(size).scatter_(0, arg, grad_out)
        return grad.narrow(0, 0, ctx.n), None, None, None


def torch_scatter_scatter(
    src: Tensor,
    index: Tensor,
    dim: int = 0,
    dim_size: Optional[int] = None,
    reduce: str = "sum",
) -> Tensor:
    """``torch_scatter.scatter`` (imported at tgp/utils/ops.py:19; used :388, :409).

    Forward values equal the PyG wrapper's.  It differs in the max/min *backward*:
    the gradient goes to a single arg element (first occurrence on CPU) instead of
    being split evenly among ties, and ``dim_size`` defaults to ``index.max()+1``.
    """
    if reduce in ("max", "min") and dim == 0:
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
        return _ScatterArgExtreme.apply(src, index, dim_size, reduce == "max")
    return scatter(src, index, dim=dim, dim_size=dim_size, reduce=reduce)


# --------------------------------------------------------------------------- #
# small helpers
# --------------------------------------------------------------------------- #
def maybe_num_nodes(edge_index: Tensor, num_nodes: Optional[int] = None) -> int:
    if num_nodes is not None:
        return num_nodes
    if isinstance(edge_index, Tensor):
        if edge_index.is_sparse:
            return max(edge_index.size(0), edge_index.size(1))
        return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
    raise NotImplementedError


def cumsum(x: Tensor, dim: int = 0) -> Tensor:
    """PyG ``cumsum``: exclusive-style prefix sum with a leading zero (length n+1)."""
    size = list(x.size())
    size[dim] += 1
    out = x.new_zeros(size)
    out.narrow(dim, 1, x.size(dim)).copy_(torch.cumsum(x, dim=dim))
    return out


def degree(index: Tensor, num_nodes: Optional[int] = None, dtype=None) -> Tensor:
    n = maybe_num_nodes(index, num_nodes)
    out = torch.zeros(n, dtype=dtype or torch.get_default_dtype(), device=index.device)
    return out.scatter_add_(0, index, out.new_ones(index.size(0)))


def index_sort(inputs: Tensor, max_value: Optional[int] = None, stable: bool = False):
    """PyG ``index_sort``: ascending sort; the CPU path is a *stable* sort."""
    return torch.sort(inputs, stable=True)


def index_to_mask(index: Tensor, size: Optional[int] = None) -> Tensor:
    index = index.view(-1)
    size = int(index.max()) + 1 if size is None else size
    mask = index.new_zeros(size, dtype=torch.bool)
    mask[index] = True
    return mask


def remove_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    if edge_attr is None:
        return edge_index, None
    return edge_index, edge_attr[mask]


# --------------------------------------------------------------------------- #
# coalesce / subgraph
# --------------------------------------------------------------------------- #
def coalesce(
    edge_index: Tensor,
    edge_attr: Optional[Tensor] = None,
    num_nodes: Optional[int] = None,
    reduce: str = "sum",
    is_sorted: bool = False,
    sort_by_row: bool = True,
):
    """PyG ``coalesce``: stable sort by ``row*num_nodes+col``, merge equal keys.

    Output order is lexicographic (row, col); duplicate attributes are combined with
    ``scatter(reduce)`` in stable-sorted (= original) order; ``edge_attr=None`` stays
    ``None`` (duplicates dropped).  If there are no duplicates the *sorted* inputs are
    returned unchanged.
    """
    nnz = edge_index.size(1)
    num_nodes = maybe_num_nodes(edge_index, num_nodes)

    key = edge_index.new_empty(nnz + 1)
    key[0] = -1
    key[1:] = edge_index[1 - int(sort_by_row)]
    key[1:].mul_(num_nodes).add_(edge_index[int(sort_by_row)])

    if not is_sorted:
        key[1:], perm = index_sort(key[1:], max_value=num_nodes * num_nodes)
        edge_index = edge_index[:, perm]
        if edge_attr is not None:
            edge_attr = edge_attr[perm]

    first = key[1:] > key[:-1]
    if bool(first.all()):
        return edge_index, edge_attr

    edge_index = edge_index[:, first]
    if edge_attr is None:
        return edge_index, None

    run = torch.arange(0, nnz, device=edge_index.device)
    run.sub_(first.logical_not().cumsum(dim=0))
    edge_attr = scatter(edge_attr, run, 0, edge_index.size(1), reduce)
    return edge_index, edge_attr


def subgraph(
    subset: Tensor,
    edge_index: Tensor,
    edge_attr: Optional[Tensor] = None,
    relabel_nodes: bool = False,
    num_nodes: Optional[int] = None,
):
    """PyG ``subgraph`` (index ``subset``): keep edges whose endpoints are both in
    ``subset`` (input order preserved); relabel endpoint -> position in ``subset``."""
    num_nodes = maybe_num_nodes(edge_index, num_nodes)
    if subset.dtype == torch.bool:
        node_mask = subset
        subset = node_mask.nonzero().view(-1)
    else:
        node_mask = index_to_mask(subset, size=num_nodes)

    edge_mask = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    edge_index = edge_index[:, edge_mask]
    edge_attr = edge_attr[edge_mask] if edge_attr is not None else None

    if relabel_nodes:
        table = torch.full((num_nodes,), -1, dtype=torch.long, device=edge_index.device)
        table[subset] = torch.arange(subset.numel(), device=edge_index.device)
        edge_index = table[edge_index.reshape(-1)].view(2, -1)
    return edge_index, edge_attr


# --------------------------------------------------------------------------- #
# batching helpers
# --------------------------------------------------------------------------- #
def unbatch(src: Tensor, batch: Tensor, dim: int = 0, batch_size: Optional[int] = None) -> List[Tensor]:
    sizes = degree(batch, batch_size, dtype=torch.long).tolist()
    return list(src.split(sizes, dim))


def unbatch_edge_index(edge_index: Tensor, batch: Tensor, batch_size: Optional[int] = None) -> List[Tensor]:
    deg = degree(batch, batch_size, dtype=torch.long)
    ptr = cumsum(deg)
    edge_batch = batch[edge_index[0]]
    edge_index = edge_index - ptr[edge_batch]
    sizes = degree(edge_batch, batch_size, dtype=torch.long).cpu().tolist()
    return list(edge_index.split(sizes, dim=1))


def to_dense_batch(
    x: Tensor,
    batch: Optional[Tensor] = None,
    fill_value: float = 0.0,
    max_num_nodes: Optional[int] = None,
    batch_size: Optional[int] = None,
) -> Tuple[Tensor, Tensor]:
    if batch is None and max_num_nodes is None:
        mask = torch.ones(1, x.size(0), dtype=torch.bool, device=x.device)
        return x.unsqueeze(0), mask
    if batch is None:
        batch = x.new_zeros(x.size(0), dtype=torch.long)
    if batch_size is None:
        batch_size = int(batch.max()) + 1 if batch.numel() > 0 else 1
    num_nodes = scatter(batch.new_ones(x.size(0)), batch, dim=0, dim_size=batch_size, reduce="sum")
    cum_nodes = cumsum(num_nodes)
    if max_num_nodes is None:
        max_num_nodes = int(num_nodes.max())
    tmp = torch.arange(batch.size(0), device=x.device) - cum_nodes[batch]
    idx = tmp + (batch * max_num_nodes)
    keep = tmp < max_num_nodes
    size = [batch_size * max_num_nodes] + list(x.size())[1:]
    out = torch.as_tensor(fill_value, device=x.device).to(x.dtype).repeat(size)
    out[idx[keep]] = x[keep]
    out = out.view([batch_size, max_num_nodes] + list(x.size())[1:])
    mask = torch.zeros(batch_size * max_num_nodes, dtype=torch.bool, device=x.device)
    mask[idx[keep]] = True
    return out, mask.view(batch_size, max_num_nodes)


def to_dense_adj(
    edge_index: Tensor,
    batch: Optional[Tensor] = None,
    edge_attr: Optional[Tensor] = None,
    max_num_nodes: Optional[int] = None,
    batch_size: Optional[int] = None,
) -> Tensor:
    """Duplicates are summed; missing ``edge_attr`` means weight 1.0."""
    if batch is None:
        n = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
        batch = edge_index.new_zeros(n)
    if batch_size is None:
        batch_size = int(batch.max()) + 1 if batch.numel() > 0 else 1
    one = batch.new_ones(batch.size(0))
    num_nodes = scatter(one, batch, dim=0, dim_size=batch_size, reduce="sum")
    cum_nodes = cumsum(num_nodes)
    idx0 = batch[edge_index[0]]
    idx1 = edge_index[0] - cum_nodes[batch][edge_index[0]]
    idx2 = edge_index[1] - cum_nodes[batch][edge_index[1]]
    if max_num_nodes is None:
        max_num_nodes = int(num_nodes.max()) if num_nodes.numel() > 0 else 0
    elif (idx1.numel() > 0 and idx1.max() >= max_num_nodes) or (idx2.numel() > 0 and idx2.max() >= max_num_nodes):
        keep = (idx1 < max_num_nodes) & (idx2 < max_num_nodes)
        idx0, idx1, idx2 = idx0[keep], idx1[keep], idx2[keep]
        edge_attr = None if edge_attr is None else edge_attr[keep]
    if edge_attr is None:
        edge_attr = torch.ones(idx0.numel(), device=edge_index.device)
    size = [batch_size, max_num_nodes, max_num_nodes] + list(edge_attr.size())[1:]
    flat = batch_size * max_num_nodes * max_num_nodes
    idx = idx0 * max_num_nodes * max_num_nodes + idx1 * max_num_nodes + idx2
    adj = scatter(edge_attr, idx, dim=0, dim_size=flat, reduce="sum")
    return adj.view(size)


# --------------------------------------------------------------------------- #
# selection helpers
# --------------------------------------------------------------------------- #
def softmax(src: Tensor, index: Tensor, num_nodes: Optional[int] = None) -> Tensor:
    n = maybe_num_nodes(index, num_nodes)
    src_max = scatter(src.detach(), index, 0, n, reduce="max")
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = scatter(out, index, 0, n, reduce="sum") + 1e-16
    return out / out_sum.index_select(0, index)


def topk(
    x: Tensor,
    ratio: Optional[float],
    batch: Tensor,
    min_score: Optional[float] = None,
    tol: float = 1e-7,
) -> Tensor:
    """PyG ``topk``: per-graph top ``ceil(ratio*n_g)`` by descending score.

    Result order = (graph ascending, score descending).  PyG sorts descending with an
    unstable ``torch.sort``; the tie order among equal scores is unpinned by the
    reference's tests, so the oracle FIXES it: lower node id first (stable sort).
    """
    if min_score is not None:
        scores_max = scatter(x, batch, reduce="max")[batch] - tol
        scores_min = scores_max.clamp(max=min_score)
        return (x > scores_min).nonzero().view(-1)

    if ratio is None:
        raise ValueError("At least one of the 'ratio' and 'min_score' parameters must be specified")

    num_nodes = scatter(batch.new_ones(x.size(0)), batch, reduce="sum")
    if ratio >= 1:
        k = num_nodes.new_full((num_nodes.size(0),), int(ratio))
    else:
        k = (float(ratio) * num_nodes.to(x.dtype)).ceil().to(torch.long)

    x_sorted, x_perm = torch.sort(x.view(-1), descending=True, stable=True)
    batch_sorted = batch[x_perm]
    batch_sorted, batch_perm = torch.sort(batch_sorted, descending=False, stable=True)

    arange = torch.arange(x.size(0), dtype=torch.long, device=x.device)
    ptr = cumsum(num_nodes)
    batched_arange = arange - ptr[batch_sorted]
    mask = batched_arange < k[batch_sorted]
    return x_perm[batch_perm[mask]]
