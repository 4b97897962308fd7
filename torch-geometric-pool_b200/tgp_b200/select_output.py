"""``SelectOutput``: the data contract between select and reduce/connect.

Host-side mirror of ``tgp.select.SelectOutput`` (tgp/select/base_select.py:75-296) holding
exactly what the Reduce + Connect path reads.  The reference's own ``SelectOutput`` objects are
accepted everywhere as well (duck typing on ``s``, ``node_index``, ``cluster_index``, ``weight``,
``num_nodes``, ``num_supernodes``, ``batch``): this class exists so the path can run where
``tgp`` / PyG are not installed.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor


def cluster_to_s(
    cluster_index: Tensor,
    node_index: Optional[Tensor] = None,
    weight: Optional[Tensor] = None,
    num_nodes: Optional[int] = None,
    num_supernodes: Optional[int] = None,
) -> Tensor:
    """Sparse assignment ``[N, K]`` from index vectors (base_select.py:19-71): ``node_index`` is
    sorted ascending and ``cluster_index`` / ``weight`` are permuted along; missing weights
    become fp32 ones."""
    if num_nodes is None:
        num_nodes = cluster_index.size(0)
    if num_supernodes is None:
        num_supernodes = int(cluster_index.max().item()) + 1
    if node_index is None:
        node_index = torch.arange(num_nodes, dtype=torch.long, device=cluster_index.device)
    node_index, perm = torch.sort(node_index)
    cluster_index = cluster_index[perm]
    indices = torch.stack([node_index, cluster_index], dim=0)
    values = weight[perm] if weight is not None else torch.ones(indices.size(1), device=indices.device)
    return torch.sparse_coo_tensor(indices, values, (num_nodes, num_supernodes), is_coalesced=True, check_invariants=False)


class SelectOutput:
    def __init__(
        self,
        s: Optional[Tensor] = None,
        s_inv: Optional[Tensor] = None,
        node_index: Optional[Tensor] = None,
        num_nodes: Optional[int] = None,
        cluster_index: Optional[Tensor] = None,
        num_supernodes: Optional[int] = None,
        weight: Optional[Tensor] = None,
        batch: Optional[Tensor] = None,
        in_mask: Optional[Tensor] = None,
        s_inv_op: str = "transpose",
        _trusted: bool = False,
        **extra_args,
    ):
        if s_inv_op != "transpose":
            raise ValueError(f"tgp_b200.SelectOutput supports s_inv_op='transpose' only, got '{s_inv_op}'")
        if isinstance(s, Tensor):
            if s.is_sparse:
                assert cluster_index is None and node_index is None
                s = s.coalesce()
                if weight is not None or num_nodes is not None or num_supernodes is not None:
                    # rebuild with the new values / size (tgp/select/base_select.py:122-141)
                    size = (num_nodes if num_nodes is not None else s.size(0),
                            num_supernodes if num_supernodes is not None else s.size(1))
                    s = torch.sparse_coo_tensor(s.indices(), weight if weight is not None else s.values(), size,
                                                is_coalesced=True, check_invariants=False)
                if s.is_cuda and not _trusted and s.indices().size(1) > 1:
                    # a tensor flagged coalesced is taken at its word by torch; the kernels rely on ascending node ids
                    from .functional import is_sorted

                    if not is_sorted(s.indices()[0]):
                        s = cluster_to_s(s.indices()[1], s.indices()[0], s.values(), s.size(0), s.size(1))
            else:
                assert cluster_index is None and node_index is None and weight is None
        elif s is None:
            assert cluster_index is not None, "'cluster_index' cannot be None if 's' is None"
            s = cluster_to_s(cluster_index, node_index, weight, num_nodes, num_supernodes)
        else:
            raise ValueError("Either 's' or 'cluster_index' must be provided.")
        self.s = s
        self.s_inv = s_inv if s_inv is not None else (s.t() if s.is_sparse else s.transpose(-1, -2))
        self.batch = batch
        if in_mask is not None:
            if in_mask.dim() != 2 or s.is_sparse or s.dim() != 3 or in_mask.shape != s.shape[:2]:
                raise ValueError("SelectOutput.in_mask must be [B, N] for a batched dense assignment [B, N, K].")
            in_mask = in_mask.to(torch.bool)
        self.in_mask = in_mask
        for k, v in extra_args.items():
            setattr(self, k, v)

    @property
    def is_sparse(self) -> bool:
        return self.s.is_sparse

    @property
    def is_dense(self) -> bool:
        return not self.s.is_sparse

    @property
    def num_nodes(self) -> int:
        return self.s.size(-2)

    @property
    def num_supernodes(self) -> int:
        return self.s.size(-1)

    @property
    def node_index(self) -> Optional[Tensor]:
        return self.s.indices()[0] if self.is_sparse else None

    @property
    def cluster_index(self) -> Optional[Tensor]:
        return self.s.indices()[1] if self.is_sparse else None

    @property
    def weight(self) -> Optional[Tensor]:
        return self.s.values() if self.is_sparse else None

    @property
    def out_mask(self) -> Optional[Tensor]:
        """[B, K] validity of pooled supernodes for batched dense S (tgp/utils/ops.py:85-132)."""
        if self.s.is_sparse:
            return None
        if self.s.dim() == 3:
            return self.s.sum(dim=-2) > 0
        if self.batch is None:
            return (self.s.sum(dim=-2) > 0).unsqueeze(0)
        return None

    def to(self, device) -> "SelectOutput":
        out = SelectOutput(s=self.s.to(device), s_inv=None if self.s_inv is None else self.s_inv.to(device),
                           batch=None if self.batch is None else self.batch.to(device),
                           in_mask=None if self.in_mask is None else self.in_mask.to(device), _trusted=True)
        skip = ("s", "s_inv", "batch", "in_mask")
        for k, v in self.__dict__.items():  # extra attributes travel along (tgp/select/base_select.py:226-244)
            if k not in skip and not k.startswith("_"):
                setattr(out, k, v.to(device) if isinstance(v, Tensor) else v)
        return out

    def __repr__(self) -> str:
        kind = "sparse" if self.is_sparse else "dense"
        return f"SelectOutput({kind}, num_nodes={self.num_nodes}, num_supernodes={self.num_supernodes})"
