"""Drop-in ``reduce`` operators (tgp/reduce/base_reduce.py, tgp/reduce/aggr_reduce.py).

``B200Reduce`` has the ``forward`` signature and return convention of ``BaseReduce`` /
``AggrReduce`` and is installed the way the reference itself swaps reducers
(``pooler.reducer = ...``, examples/classification_aggr_reduce.py:74).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import functional as F_


class Reduce(nn.Module):
    """Template of the reduce operator (tgp/reduce/base_reduce.py:11-85)."""

    @staticmethod
    def reduce_batch(select_output, batch: Optional[Tensor]) -> Optional[Tensor]:
        """Batch vector of the coarsened graph (base_reduce.py:15-53)."""
        if batch is None:
            return None
        if select_output.s.is_sparse:
            return F_.reduce_batch_sparse(select_output, batch)
        if batch.numel() == 0:
            return batch.new_empty((0,), dtype=batch.dtype)
        batch_size = int(batch.max().item()) + 1
        return torch.arange(batch_size, dtype=batch.dtype, device=batch.device).repeat_interleave(
            select_output.num_supernodes
        )

    def reset_parameters(self):
        pass

    def forward(self, x: Tensor, so, *, batch: Optional[Tensor] = None, **kwargs):
        raise NotImplementedError

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"


class B200Reduce(Reduce):
    r"""``S^T X`` on B200 kernels.

    * sparse S: deterministic segment reduction (``reduce_op`` in sum / mean / max / min;
      sum reproduces ``BaseReduce``, the others ``AggrReduce`` with the matching PyG aggregation);
    * dense batched S ``[B, N, K]``: batched GEMM ``S^T X``.
    """

    def __init__(self, reduce_op: str = "sum"):
        super().__init__()
        if reduce_op not in ("sum", "add", "mean", "max", "min"):
            raise ValueError(f"Unknown aggregator alias '{reduce_op}'")
        self.reduce_op = reduce_op

    def forward(
        self,
        x: Tensor,
        so=None,
        *,
        batch: Optional[Tensor] = None,
        return_batched: bool = False,
        size: Optional[int] = None,
        **kwargs,
    ) -> Tuple[Tensor, Optional[Tensor]]:
        if so is None:  # readout mode of AggrReduce (aggr_reduce.py:112-153)
            return self._readout(x, batch=batch, size=size)
        if batch is None and so.batch is not None:
            batch = so.batch

        if so.s.is_sparse:
            if return_batched:
                raise ValueError("return_batched=True is only supported for dense assignment matrices.")
            x_pool = F_.segment_reduce(
                x, so.node_index, so.cluster_index, so.weight, so.num_supernodes, self.reduce_op, csr=F_.csr_of(so)
            )
            return x_pool, self.reduce_batch(so, batch)

        if self.reduce_op not in ("sum", "add"):
            raise ValueError(
                "AggrReduce supports only sparse SelectOutput assignments. "
                "Dense assignments are not supported; use BaseReduce for dense/soft reductions."
            )
        if so.s.dim() == 3:
            x_pool, _, _ = F_.dense_pool(x, None, so.s)
            return x_pool, self.reduce_batch(so, batch)
        if so.s.dim() != 2:
            raise ValueError(f"Dense SelectOutput.s must be 2D [N, K] or 3D [B, N, K], got ndim={so.s.dim()}.")
        multi = batch is not None and batch.numel() > 0 and int(batch.min().item()) != int(batch.max().item())
        if multi:  # base_reduce.py:170-182: per-graph S_i^T X_i, here as one padded batched product
            B = int(batch.max().item()) + 1
            s3, _ = F_.to_dense_batch(so.s, batch, B)
            x3, _ = F_.to_dense_batch(x, batch, B, s3.size(1))
            x_pool, _, _ = F_.dense_pool(x3, None, s3)
            if not return_batched:
                x_pool = x_pool.reshape(B * so.num_supernodes, -1)
            return x_pool, self.reduce_batch(so, batch)
        x_pool, _, _ = F_.dense_pool(x.unsqueeze(0), None, so.s.unsqueeze(0))
        x_pool = x_pool if return_batched else x_pool.squeeze(0)
        return x_pool, self.reduce_batch(so, batch)

    def _readout(self, x: Tensor, *, batch: Optional[Tensor], size: Optional[int]):
        dev = x.device
        if x.dim() == 3:
            B, N, _ = x.shape
            k = size if size is not None else B
            idx = torch.arange(B, device=dev).repeat_interleave(N)
            nodes = torch.arange(B * N, device=dev)
            out = F_.segment_reduce(x.reshape(B * N, -1), nodes, idx, None, k, self.reduce_op)
            return out, torch.arange(k, device=dev)
        if x.dim() != 2:
            raise ValueError(f"Readout mode expects x to be 2D [N, F] or 3D [B, N, F], got ndim={x.dim()}.")
        nodes = torch.arange(x.size(0), device=dev)
        if batch is None:
            idx = torch.zeros(x.size(0), dtype=torch.long, device=dev)
            return F_.segment_reduce(x, nodes, idx, None, 1, self.reduce_op), None
        k = size if size is not None else (int(batch.max().item()) + 1 if batch.numel() > 0 else 1)
        return F_.segment_reduce(x, nodes, batch, None, k, self.reduce_op), torch.arange(k, device=dev)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(reduce_op={self.reduce_op})"
