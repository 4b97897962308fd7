"""Unbatched dense mode (``batched=False``): sparse adjacency + dense ``[N, K]`` assignment, no ``[B, N, N]`` tensor.

Mirror of ``DenseConnect._dense_connect_unbatched`` (tgp/connect/dense_conn.py:141-208) and of the sparse loss
twins ``sparse_mincut_loss`` / ``unbatched_orthogonality_loss`` / ``sparse_link_pred_loss``
(tgp/utils/losses.py:126-215, 319-389, 711-777).  The reference loops over the graphs of the batch in Python
(one ``torch.sparse.mm`` + one matmul per graph); here the whole batch is

* one SpMM ``W = A S`` (``[N, K]``): the segment-reduce kernel keyed by edge row, gathering rows of ``S``
  (deterministic, edge order inside a row), and
* one batched tensor-core product ``S_g^T W_g`` over the graphs padded to ``[B, Nmax, K]`` -- the padding is
  ``O(B Nmax K)``, never ``O(B Nmax^2)``, which is what the unbatched mode exists to avoid.

The losses reuse ``W`` (``tr(S^T A S) = sum_i <S_i, W_i>``), so they cost no extra pass over the edges and no
``[E, K]`` temporaries.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L
from . import functional as F_

EPS = 1e-8


class _SpMM(torch.autograd.Function):
    """``out[i] = sum_{e: row_e = i} w_e x[col_e]`` (``A x`` for the COO adjacency); backward = SpMM with the
    transposed adjacency plus one row dot product per edge for the weights."""

    @staticmethod
    def forward(ctx, x, weight, row, col, num_nodes):
        E, F = row.numel(), x.size(1)
        order, ptr = F_.build_csr(row, num_nodes)
        out = torch.empty((num_nodes, F), dtype=x.dtype, device=x.device)
        dt = L.dtype_code(x.dtype)
        L.call("tgpb200_segment_reduce_fwd", L.ptr(x), L.ptr(col), L.ptr(weight), L.ptr(order), L.ptr(ptr), x.size(0), E,
               num_nodes, F, L.SUM, dt, dt, L.ptr(out), L.stream())
        ctx.save_for_backward(x, weight, row, col)
        ctx.N = num_nodes
        return out

    @staticmethod
    def backward(ctx, g):
        x, weight, row, col = ctx.saved_tensors
        E, F = row.numel(), x.size(1)
        g = g.contiguous()
        gx = None
        if ctx.needs_input_grad[0]:
            order, ptr = F_.build_csr(col, x.size(0))
            gx = torch.empty_like(x)
            dt = L.dtype_code(g.dtype)
            L.call("tgpb200_segment_reduce_fwd", L.ptr(g), L.ptr(row), L.ptr(weight), L.ptr(order), L.ptr(ptr), ctx.N, E,
                   x.size(0), F, L.SUM, dt, dt, L.ptr(gx), L.stream())
        gw = None
        if weight is not None and ctx.needs_input_grad[1]:
            gw = (g[row].float() * x[col].float()).sum(-1)
        return gx, gw, None, None, None


def spmm(edge_index: Tensor, edge_weight: Optional[Tensor], x: Tensor, num_nodes: Optional[int] = None) -> Tensor:
    """``A x`` for a COO adjacency (duplicates add up), differentiable in ``x`` and ``edge_weight``."""
    F_._require_cuda(edge_index, edge_weight, x)
    F_._validate_edge_index(edge_index)
    n = x.size(0) if num_nodes is None else num_nodes
    w = F_._as_f32_weight(edge_weight)
    ei = edge_index.contiguous()
    if ei.size(1) == 0:
        return x.new_zeros((n, x.size(1)))
    return _SpMM.apply(x.contiguous(), w, ei[0], ei[1], n)


class _BmmTN(torch.autograd.Function):
    """``a^T b`` per batch item: ``[B, N, K1]^T [B, N, K2] -> [B, K1, K2]`` (contraction over the nodes)."""

    @staticmethod
    def forward(ctx, a, b):
        B, N, K1 = a.shape
        K2 = b.size(-1)
        out = torch.empty((B, K1, K2), dtype=a.dtype, device=a.device)
        L.call("tgpb200_bmm", L.ptr(a), L.ptr(b), L.ptr(out), B, K1, K2, N, N * K1, K1, 1, N * K2, K2, 1,
               L.dtype_code(a.dtype), L.stream())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        B, N, K1 = a.shape
        K2 = b.size(-1)
        g = g.contiguous()
        dt = L.dtype_code(a.dtype)
        ga = torch.empty_like(a)  # b g^T  [N, K1]
        L.call("tgpb200_bmm", L.ptr(b), L.ptr(g), L.ptr(ga), B, N, K1, K2, N * K2, K2, 0, K1 * K2, K2, 0, dt, L.stream())
        gb = torch.empty_like(b)  # a g    [N, K2]
        L.call("tgpb200_bmm", L.ptr(a), L.ptr(g), L.ptr(gb), B, N, K2, K1, N * K1, K1, 0, K1 * K2, K2, 1, dt, L.stream())
        return ga, gb


def _graphs_of(batch: Optional[Tensor], num_nodes: int, num_graphs: Optional[int], dev) -> Tuple[Tensor, int]:
    if batch is None:
        return torch.zeros(num_nodes, dtype=torch.long, device=dev), 1
    if num_graphs is None:  # the reference's own sync (losses.py:187, dense_conn.py:329): pass num_graphs to avoid it
        num_graphs = int(batch.max().item()) + 1 if batch.numel() else 1
    return batch, num_graphs


def _per_graph_sum(v: Tensor, batch: Tensor, num_graphs: int) -> Tensor:
    """Deterministic per-graph sums of a node vector (segment-reduce kernel in readout mode)."""
    nodes = torch.arange(v.numel(), device=v.device)
    return F_.segment_reduce(v.view(-1, 1).float(), nodes, batch, None, num_graphs, "sum",
                             csr=F_.build_csr(batch, num_graphs)).view(-1)


def dense_connect_unbatched(edge_index: Tensor, edge_weight: Optional[Tensor], batch: Optional[Tensor], s: Tensor,
                            num_graphs: Optional[int] = None, return_aux: bool = False):
    """Raw ``S^T A S`` ``[B, K, K]`` from a sparse adjacency and ``S [N, K]`` (dense_conn.py:141-208)."""
    N, K = s.shape
    batch, B = _graphs_of(batch, N, num_graphs, s.device)
    w_as = spmm(edge_index, edge_weight, s, N)  # A S
    s3, _ = F_.to_dense_batch(s, batch, B)
    w3, _ = F_.to_dense_batch(w_as, batch, B, s3.size(1))
    adj_pool = _BmmTN.apply(s3.contiguous(), w3.contiguous())
    return (adj_pool, w_as, s3, batch, B) if return_aux else adj_pool


def sparse_mincut_loss(edge_index: Tensor, S: Tensor, edge_weight: Optional[Tensor] = None,
                       batch: Optional[Tensor] = None, batch_reduction: str = "mean",
                       num_graphs: Optional[int] = None, _as: Optional[Tensor] = None) -> Tensor:
    """tgp/utils/losses.py:126-215: ``-tr(S^T A S) / (sum_i d_i |S_i|^2 + eps)`` per graph."""
    N = S.size(0)
    batch, B = _graphs_of(batch, N, num_graphs, S.device)
    w_as = _as if _as is not None else spmm(edge_index, edge_weight, S, N)
    deg = spmm(edge_index, edge_weight, torch.ones(N, 1, dtype=S.dtype, device=S.device), N).view(-1)
    num = _per_graph_sum((S * w_as).sum(-1), batch, B)
    den = _per_graph_sum(deg * (S * S).sum(-1), batch, B)
    loss = -(num / (den + EPS))
    return loss.mean(0) if batch_reduction == "mean" else loss.sum(0)


def unbatched_orthogonality_loss(S: Tensor, batch: Optional[Tensor] = None, batch_reduction: str = "mean",
                                 num_graphs: Optional[int] = None, _s3: Optional[Tensor] = None) -> Tensor:
    """tgp/utils/losses.py:319-389: per graph ``|| S_g^T S_g / ||.||_F - I / sqrt(K) ||_F`` (one batched product
    instead of the reference's loop over graphs)."""
    K = S.size(1)
    batch, B = _graphs_of(batch, S.size(0), num_graphs, S.device)
    s3 = _s3 if _s3 is not None else F_.to_dense_batch(S, batch, B)[0].contiguous()
    sts = _BmmTN.apply(s3, s3).float()
    sts = sts / torch.linalg.matrix_norm(sts, keepdim=True)
    loss = torch.linalg.matrix_norm(sts - torch.eye(K, device=S.device) / math.sqrt(K))
    return loss.mean(0) if batch_reduction == "mean" else loss.sum(0)


def sparse_link_pred_loss(S: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None,
                          batch: Optional[Tensor] = None, normalize_loss: bool = True,
                          num_graphs: Optional[int] = None, _as: Optional[Tensor] = None,
                          _s3: Optional[Tensor] = None) -> Tensor:
    """tgp/utils/losses.py:711-777.  The reference's ``sum_e (w - ss)^2 - sum_e ss^2`` equals
    ``sum_e w^2 - 2 sum_i <S_i, (A S)_i>`` term by term, so the edges are visited once (inside the SpMM) and no
    ``[E, K]`` gather is made; ``sum_g ||S_g^T S_g||_F^2`` comes from one batched product."""
    N = S.size(0)
    batch, B = _graphs_of(batch, N, num_graphs, S.device)
    w_as = _as if _as is not None else spmm(edge_index, edge_weight, S, N)
    w = F_._as_f32_weight(edge_weight)
    sum_w2 = (w * w).sum() if w is not None else torch.tensor(float(edge_index.size(1)), device=S.device)
    s3 = _s3 if _s3 is not None else F_.to_dense_batch(S, batch, B)[0].contiguous()
    sts = _BmmTN.apply(s3, s3).float()
    sq = sum_w2 - 2.0 * (S.float() * w_as.float()).sum() + (sts * sts).sum()
    link = torch.sqrt(torch.clamp(sq, min=0.0))
    if normalize_loss:
        counts = torch.bincount(batch, minlength=B).to(link.dtype)
        link = link / (counts * counts).sum().clamp(min=1.0)
    return link


def mincut_pool_unbatched(x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor], s: Tensor,
                          batch: Optional[Tensor] = None, num_graphs: Optional[int] = None,
                          cut_loss_coeff: float = 1.0, ortho_loss_coeff: float = 1.0, remove_self_loops: bool = True,
                          degree_norm: bool = True, edge_weight_norm: bool = False) -> Tuple[Tensor, Tensor, Dict]:
    """Unbatched MinCutPooling after select (tgp/poolers/mincut.py:260-289): reduce, connect and both losses from ONE
    SpMM and two batched products.  Returns ``(x_pool [B, K, F], adj_pool [B, K, K], losses)``."""
    from .connect import postprocess_adj_pool_dense

    raw, w_as, s3, b, B = dense_connect_unbatched(edge_index, edge_weight, batch, s, num_graphs, return_aux=True)
    x3, _ = F_.to_dense_batch(x, b, B, s3.size(1))
    x_pool = _BmmTN.apply(s3.contiguous(), x3.contiguous())
    loss = {"cut_loss": sparse_mincut_loss(edge_index, s, edge_weight, b, num_graphs=B, _as=w_as) * cut_loss_coeff,
            "ortho_loss": unbatched_orthogonality_loss(s, b, num_graphs=B, _s3=s3.contiguous()) * ortho_loss_coeff}
    adj_pool = postprocess_adj_pool_dense(raw, remove_self_loops=remove_self_loops, degree_norm=degree_norm,
                                          adj_transpose=False, edge_weight_norm=edge_weight_norm)
    return x_pool, adj_pool, loss
