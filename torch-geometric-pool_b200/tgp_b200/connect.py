"""Drop-in ``connect`` operators (tgp/connect/base_conn.py, tgp/connect/dense_conn.py)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from . import functional as F_
from .functional import sparse_connect  # noqa: F401  (functional form, base_conn.py:57-69)


class Connect(nn.Module):
    """Template of the connect operator (tgp/connect/base_conn.py:20-54)."""

    def reset_parameters(self):
        pass

    def forward(self, edge_index, so, *, edge_weight: Optional[Tensor] = None, **kwargs):
        raise NotImplementedError

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}()"


class B200SparseConnect(Connect):
    """``SparseConnect`` (base_conn.py:115-224) on B200 kernels: same ctor attributes, same
    ``forward(edge_index, so, *, edge_weight, batch_pooled)`` and the same errors."""

    def __init__(
        self,
        reduce_op: str = "sum",
        remove_self_loops: bool = True,
        edge_weight_norm: bool = False,
        degree_norm: bool = False,
    ):
        super().__init__()
        self.reduce_op = reduce_op
        self.remove_self_loops = remove_self_loops
        self.edge_weight_norm = edge_weight_norm
        self.degree_norm = degree_norm

    def forward(
        self,
        edge_index,
        so,
        *,
        edge_weight: Optional[Tensor] = None,
        batch_pooled: Optional[Tensor] = None,
        **kwargs,
    ) -> Tuple[Tensor, Optional[Tensor]]:
        if self.edge_weight_norm and batch_pooled is None:
            raise AssertionError(
                "edge_weight_norm=True but batch_pooled=None. "
                "batch_pooled parameter is required for per-graph normalization in SparseConnect."
            )
        return F_.sparse_connect(
            edge_index,
            edge_weight,
            node_index=so.node_index,
            cluster_index=so.cluster_index,
            num_nodes=so.num_nodes,
            num_supernodes=so.num_supernodes,
            remove_self_loops=self.remove_self_loops,
            reduce_op=self.reduce_op,
            edge_weight_norm=self.edge_weight_norm,
            batch_pooled=batch_pooled,
            degree_norm=self.degree_norm,
            csr=F_.csr_of(so) if (so.cluster_index is not None and len(so.cluster_index) == so.num_nodes) else None,
        )

    def __repr__(self) -> str:
        return (
            f"{self.__class__.__name__}(reduce_op={self.reduce_op}, remove_self_loops={self.remove_self_loops}, "
            f"edge_weight_norm={self.edge_weight_norm}, degree_norm={self.degree_norm})"
        )


def is_dense_adj(edge_index) -> bool:
    """tgp/utils/ops.py:267-279."""
    if not isinstance(edge_index, Tensor) or edge_index.is_sparse:
        return False
    if edge_index.dim() == 3:
        return True
    if edge_index.dim() == 2 and edge_index.size(0) == edge_index.size(1):
        return edge_index.is_floating_point()
    return False


def postprocess_adj_pool_dense(
    adj_pool: Tensor,
    remove_self_loops: bool = False,
    degree_norm: bool = False,
    adj_transpose: bool = False,
    edge_weight_norm: bool = False,
) -> Tensor:
    """tgp/utils/ops.py:282-335 as a stand-alone call (returns a new tensor; the reference zeroes the diagonal in
    place): the post-processing epilogue of the fused op, reached with S = I (``I^T A I = A``; differentiable)."""
    squeeze = adj_pool.dim() == 2
    a = adj_pool.unsqueeze(0) if squeeze else adj_pool
    if a.dim() != 3 or a.size(-1) != a.size(-2):
        raise ValueError("adj_pool must have shape [B, K, K] or [K, K].")
    eye = torch.eye(a.size(-1), dtype=a.dtype, device=a.device).expand(a.size(0), -1, -1).contiguous()
    _, out, _ = F_.dense_pool(None, a, eye, remove_self_loops=remove_self_loops, degree_norm=degree_norm,
                              adj_transpose=adj_transpose, edge_weight_norm=edge_weight_norm)
    return out.squeeze(0) if squeeze else out


class B200DenseConnect(Connect):
    """``DenseConnect`` (dense_conn.py:22-364) for batched dense inputs on B200 kernels."""

    def __init__(
        self,
        remove_self_loops: bool = True,
        degree_norm: bool = True,
        adj_transpose: bool = True,
        edge_weight_norm: bool = False,
        sparse_output: bool = False,
    ):
        super().__init__()
        if not isinstance(sparse_output, bool):
            raise TypeError("sparse_output must be a bool.")
        self.remove_self_loops = remove_self_loops
        self.degree_norm = degree_norm
        self.adj_transpose = adj_transpose
        self.edge_weight_norm = edge_weight_norm
        self.sparse_output = sparse_output

    @staticmethod
    def _prepare_batched_dense_inputs(s: Tensor, adj: Tensor) -> Tuple[Tensor, Tensor]:
        if s.dim() == 2:
            s = s.unsqueeze(0)
        if adj.dim() == 2:
            adj = adj.unsqueeze(0)
        if s.dim() != 3 or adj.dim() != 3:
            raise ValueError("Expected batched dense inputs with 3 dimensions.")
        if s.size(0) != adj.size(0):
            raise ValueError(
                "Assignment and adjacency batch sizes do not match: "
                f"got s.size(0)={s.size(0)} and adj.size(0)={adj.size(0)}."
            )
        return s, adj

    @staticmethod
    def _validate_select_output(so) -> Tensor:
        if so is None:
            raise ValueError("SelectOutput is required for DenseConnect.")
        s = so.s
        if not isinstance(s, Tensor):
            raise TypeError("SelectOutput.s must be a torch.Tensor.")
        if s.is_sparse:
            raise ValueError("DenseConnect expects a dense assignment matrix.")
        return s

    def dense_connect(self, adj: Tensor, s: Tensor) -> Tensor:
        """Raw ``S^T A S`` ``[B, K, K]`` (dense_conn.py:124-138)."""
        s, adj = self._prepare_batched_dense_inputs(s, adj)
        _, adj_pool, _ = F_.dense_pool(None, adj, s)
        return adj_pool

    def forward(
        self,
        edge_index,
        so,
        *,
        edge_weight: Optional[Tensor] = None,
        batch: Optional[Tensor] = None,
        batch_pooled: Optional[Tensor] = None,
        **kwargs,
    ):
        s = self._validate_select_output(so)
        if not is_dense_adj(edge_index):
            return self._forward_unbatched_inputs(edge_index, edge_weight, batch, s, batch_pooled)
        s, adj = self._prepare_batched_dense_inputs(s, edge_index)
        _, adj_pool, _ = F_.dense_pool(
            None,
            adj,
            s,
            remove_self_loops=self.remove_self_loops,
            degree_norm=self.degree_norm,
            adj_transpose=self.adj_transpose,
            edge_weight_norm=self.edge_weight_norm,
        )
        return adj_pool, None

    def _forward_unbatched_inputs(self, edge_index, edge_weight, batch, s, batch_pooled):
        """Sparse adjacency + dense ``[N, K]`` assignment (dense_conn.py:273-354).  The per-graph ``sparse.mm`` loop
        of the reference becomes one SpMM kernel + one batched tensor-core product (``tgp_b200.unbatched``); memory
        stays ``O(E + B Nmax K)``."""
        to_coo = isinstance(edge_index, Tensor) and edge_index.is_sparse
        if to_coo:
            coo = edge_index.coalesce()
            edge_index, edge_weight = coo.indices(), coo.values()
        if s.dim() == 3:
            if s.size(0) != 1:
                raise ValueError(
                    "[DenseConnect - unbatched]: SelectOutput.s must have shape "
                    f"[N, K] or [1, N, K], but got {s.size()}."
                )
            s = s.squeeze(0)
        elif s.dim() != 2:
            raise ValueError(
                "[DenseConnect - unbatched]: SelectOutput.s must have shape "
                f"[N, K] or [1, N, K], but got {s.size()}."
            )
        from .unbatched import dense_connect_unbatched

        num_nodes, K = s.size()
        B = 1 if batch is None else int(batch.max().item()) + 1  # the reference's own sync (dense_conn.py:329)
        # S^T (A S): one SpMM over the edge list + one batched product over [B, Nmax, K] -- no [B, Nmax, Nmax] tensor
        raw = dense_connect_unbatched(edge_index, edge_weight, batch, s, B)
        if not self.sparse_output:
            return postprocess_adj_pool_dense(raw, remove_self_loops=self.remove_self_loops,
                                              degree_norm=self.degree_norm, adj_transpose=False,
                                              edge_weight_norm=self.edge_weight_norm), None
        if self.edge_weight_norm and batch_pooled is None:
            raise AssertionError(
                "edge_weight_norm=True but batch_pooled=None. "
                "batch_pooled parameter is required for per-graph normalization in DenseConnect."
            )
        ei, ew = F_.dense_to_block_diag(raw)
        n_super = B * K
        flags = F_.L.REMOVE_SELF_LOOPS if self.remove_self_loops else 0
        ident = torch.arange(n_super, device=s.device)
        ew32 = ew.to(torch.float32)
        ei, ew, _, _ = F_.O.filter_relabel_edges(ei[0].contiguous(), ei[1].contiguous(), ew32, ident, n_super, flags,
                                                 F_.EPS, False, ew32.requires_grad)
        ew = F_.edge_postprocess(ei, ew, n_super, self.degree_norm, self.edge_weight_norm, batch_pooled,
                                 sorted_rows=True)  # block-diagonal order is row-major
        if to_coo:
            return torch.sparse_coo_tensor(ei, ew, (n_super, n_super)).coalesce(), None
        return ei, ew

    def __repr__(self) -> str:
        return (
            f"{self.__class__.__name__}(remove_self_loops={self.remove_self_loops}, degree_norm={self.degree_norm}, "
            f"adj_transpose={self.adj_transpose}, edge_weight_norm={self.edge_weight_norm}, "
            f"sparse_output={self.sparse_output})"
        )
