"""``lift`` operator on the B200 kernels: ``x_lift = S_inv^T-style matrix @ x_pool``.

Mirror of ``tgp.lift.BaseLift`` (tgp/lift/base_lift.py:85-240) for the cases that reuse the Reduce kernels:

* sparse S (``matrix_op`` "precomputed" with the default transposed ``s_inv``, or "transpose"):
  ``x_lift[n] = op_{i: node_index[i] = n} weight[i] * x_pool[cluster_index[i]]`` -- the segment reduction of
  ``tgp_b200.functional.segment_reduce`` with the two index roles swapped (base_lift.py:113-123);
* dense batched S ``[B, N, K]`` with ``x_pool [B, K, F]``: one tensor-core batched product ``S x_pool``.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor, nn

from . import _lib as L
from . import functional as F_


class _LiftSparse(torch.autograd.Function):
    """Forward = segment reduce keyed by node; backward = segment reduce keyed by cluster (the Reduce forward)
    plus one row dot product per entry for the weights."""

    @staticmethod
    def forward(ctx, x_pool, weight, node_index, cluster_index, num_nodes, op):
        K, F = x_pool.shape
        nnz = node_index.numel()
        order_n, ptr_n = F_.build_csr(node_index, num_nodes)
        out = torch.empty((num_nodes, F), dtype=x_pool.dtype, device=x_pool.device)
        L.call("tgpb200_segment_reduce_fwd", L.ptr(x_pool), L.ptr(cluster_index), L.ptr(weight), L.ptr(order_n),
               L.ptr(ptr_n), K, nnz, num_nodes, F, op, L.dtype_code(x_pool.dtype), L.dtype_code(x_pool.dtype),
               L.ptr(out), L.stream())
        ctx.save_for_backward(x_pool, weight, node_index, cluster_index)
        ctx.op, ctx.N = op, num_nodes
        return out

    @staticmethod
    def backward(ctx, g):
        x_pool, weight, node_index, cluster_index = ctx.saved_tensors
        if ctx.op != L.SUM:
            raise RuntimeError("tgp_b200 lift: backward is implemented for reduce_op='sum'")
        K, F = x_pool.shape
        nnz = node_index.numel()
        g = g.contiguous()
        order_c, ptr_c = F_.build_csr(cluster_index, K)
        gx = torch.empty_like(x_pool)
        L.call("tgpb200_segment_reduce_fwd", L.ptr(g), L.ptr(node_index), L.ptr(weight), L.ptr(order_c), L.ptr(ptr_c),
               ctx.N, nnz, K, F, L.SUM, L.dtype_code(g.dtype), L.dtype_code(g.dtype), L.ptr(gx), L.stream())
        gw = None
        if weight is not None and ctx.needs_input_grad[1]:
            gw = (x_pool[cluster_index].float() * g[node_index].float()).sum(-1)
        return gx, gw, None, None, None, None


class _LiftDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, x_pool):
        B, N, K = s.shape
        F = x_pool.size(-1)
        out = torch.empty((B, N, F), dtype=s.dtype, device=s.device)
        dt = L.dtype_code(s.dtype)
        L.call("tgpb200_bmm", L.ptr(s), L.ptr(x_pool), L.ptr(out), B, N, F, K, N * K, K, 0, K * F, F, 1, dt, L.stream())
        ctx.save_for_backward(s, x_pool)
        return out

    @staticmethod
    def backward(ctx, g):
        s, x_pool = ctx.saved_tensors
        B, N, K = s.shape
        F = x_pool.size(-1)
        g = g.contiguous()
        dt = L.dtype_code(s.dtype)
        gs = torch.empty_like(s)      # g x_pool^T  [N, K]
        L.call("tgpb200_bmm", L.ptr(g), L.ptr(x_pool), L.ptr(gs), B, N, K, F, N * F, F, 0, K * F, F, 0, dt, L.stream())
        gx = torch.empty_like(x_pool)  # S^T g  [K, F]
        L.call("tgpb200_bmm", L.ptr(s), L.ptr(g), L.ptr(gx), B, K, F, N, N * K, K, 1, N * F, F, 1, dt, L.stream())
        return gs, gx


class B200Lift(nn.Module):
    def __init__(self, matrix_op: str = "precomputed", reduce_op: str = "sum"):
        super().__init__()
        if matrix_op not in ("precomputed", "transpose"):
            raise RuntimeError(f"'matrix_op' must be 'precomputed' or 'transpose' on this backend ({matrix_op} given)")
        if reduce_op not in ("sum", "add"):
            # BaseLift's default; the other scatter reductions would need their own backward
            raise ValueError(f"tgp_b200 lift supports reduce_op='sum' only ({reduce_op} given)")
        self.matrix_op = matrix_op
        self.reduce_op = reduce_op

    def reset_parameters(self):
        pass

    def forward(self, x_pool: Tensor, so=None, batch: Optional[Tensor] = None, batch_pooled: Optional[Tensor] = None,
                **kwargs) -> Tensor:
        s = so.s
        if s.is_sparse:
            if self.reduce_op not in L.OPS or self.reduce_op == "mul":
                raise ValueError(f"unsupported reduce op '{self.reduce_op}'")
            w = so.weight
            w32 = None if w is None else w.to(torch.float32).contiguous()
            return _LiftSparse.apply(x_pool.contiguous(), w32, so.node_index.contiguous(), so.cluster_index.contiguous(),
                                     so.num_nodes, L.OPS[self.reduce_op])
        if s.dim() == 3 and x_pool.dim() == 3:
            return _LiftDense.apply(s.contiguous(), x_pool.contiguous())
        if s.dim() == 3 and x_pool.dim() == 2:
            B, _, K = s.shape
            return _LiftDense.apply(s.contiguous(), x_pool.reshape(B, K, -1).contiguous())
        if s.dim() == 2 and x_pool.dim() == 2 and x_pool.size(0) == s.size(-1):
            return _LiftDense.apply(s.unsqueeze(0).contiguous(), x_pool.unsqueeze(0).contiguous()).squeeze(0)
        if s.dim() == 2 and x_pool.dim() == 2 and batch is not None:
            # dense [N, K] assignment over a multi-graph batch (base_lift.py:170-190): pad to [B, Nmax, K], one
            # batched product, gather the valid rows back
            K = s.size(-1)
            B = x_pool.size(0) // K
            s3, mask = F_.to_dense_batch(s, batch, B)
            out = _LiftDense.apply(s3.contiguous(), x_pool.reshape(B, K, -1).contiguous())
            return out[mask]
        raise ValueError("tgp_b200 lift: dense [N, K] assignment with x_pool [B*K, F] needs `batch`")

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(matrix_op={self.matrix_op}, reduce_op={self.reduce_op})"
