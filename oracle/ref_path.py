"""CPU restatement of tgp's Reduce + Connect path (torch CPU ops, autograd for backward).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites the
reference lines (relative to /root/reference) it follows.  Gradients come from
``torch.autograd`` over these forwards, exactly as in the reference, which defines
no custom backward.
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import pyg_shim as pyg

EPS = 1e-8  # tgp/__init__.py:6


# --------------------------------------------------------------------------- #
# SelectOutput (sparse + dense views)            tgp/select/base_select.py:19-296
# --------------------------------------------------------------------------- #
class OracleSelectOutput:
    """The slice of ``SelectOutput`` the path reads.

    Sparse S is held as a coalesced COO ``[N, K]`` whose ``node_index`` is sorted
    ascending with ``cluster_index`` / ``weight`` permuted along
    (base_select.py:56-71); missing weights become fp32 ones (:61-65).
    """

    def __init__(
        self,
        s: Optional[Tensor] = None,
        node_index: Optional[Tensor] = None,
        num_nodes: Optional[int] = None,
        cluster_index: Optional[Tensor] = None,
        num_supernodes: Optional[int] = None,
        weight: Optional[Tensor] = None,
        batch: Optional[Tensor] = None,
        in_mask: Optional[Tensor] = None,
    ):
        if s is None:
            assert cluster_index is not None
            if num_nodes is None:
                num_nodes = cluster_index.size(0)
            if num_supernodes is None:
                num_supernodes = int(cluster_index.max()) + 1
            if node_index is None:
                node_index = torch.arange(num_nodes, dtype=torch.long)
            node_index, perm = torch.sort(node_index)
            cluster_index = cluster_index[perm]
            values = weight[perm] if weight is not None else torch.ones(node_index.numel())
            s = torch.sparse_coo_tensor(
                torch.stack([node_index, cluster_index]), values, (num_nodes, num_supernodes), is_coalesced=True, check_invariants=False
            )
        elif s.is_sparse:
            s = s.coalesce()
        self.s = s
        self.batch = batch
        self.in_mask = in_mask

    @property
    def is_sparse(self) -> bool:
        return self.s.is_sparse

    @property
    def num_nodes(self) -> int:
        return self.s.size(-2)

    @property
    def num_supernodes(self) -> int:
        return self.s.size(-1)

    @property
    def node_index(self):
        return self.s.indices()[0] if self.is_sparse else None

    @property
    def cluster_index(self):
        return self.s.indices()[1] if self.is_sparse else None

    @property
    def weight(self):
        return self.s.values() if self.is_sparse else None


# --------------------------------------------------------------------------- #
# Reduce                                          tgp/reduce/base_reduce.py
# --------------------------------------------------------------------------- #
def reduce_batch(so: OracleSelectOutput, batch: Optional[Tensor]) -> Optional[Tensor]:
    """base_reduce.py:15-53."""
    if batch is None:
        return None
    if so.is_sparse:
        out = torch.arange(so.num_supernodes, device=batch.device)
        return out.scatter_(0, so.cluster_index, batch[so.node_index])
    if batch.numel() == 0:
        return batch.new_empty((0,))
    batch_size = int(batch.max()) + 1
    # build_pooled_batch, tgp/utils/ops.py:152-169
    return torch.arange(batch_size, dtype=batch.dtype).repeat_interleave(so.num_supernodes)


def base_reduce(
    x: Tensor, so: OracleSelectOutput, batch: Optional[Tensor] = None, return_batched: bool = False
) -> Tuple[Tensor, Optional[Tensor]]:
    """BaseReduce.forward, base_reduce.py:108-190 (S^T X)."""
    if batch is None and so.batch is not None:
        batch = so.batch
    if so.is_sparse:  # :141-155
        if return_batched:
            raise ValueError("return_batched=True is only supported for dense assignment matrices.")
        src = x[so.node_index] * so.weight.view(-1, 1)
        x_pool = pyg.scatter(src, so.cluster_index, dim=0, dim_size=so.num_supernodes, reduce="sum")
        return x_pool, reduce_batch(so, batch)
    if so.s.dim() == 3:  # :158-161
        return so.s.transpose(-2, -1).matmul(x), reduce_batch(so, batch)
    if so.s.dim() != 2:
        raise ValueError(f"Dense SelectOutput.s must be 2D [N, K] or 3D [B, N, K], got ndim={so.s.dim()}.")
    multi = batch is not None and batch.numel() > 0 and int(batch.min()) != int(batch.max())  # ops.py:135-149
    if multi:  # :170-182
        parts = [s_i.t().matmul(x_i) for s_i, x_i in zip(pyg.unbatch(so.s, batch), pyg.unbatch(x, batch))]
        x_pool = torch.stack(parts, 0) if return_batched else torch.cat(parts, 0)
        return x_pool, reduce_batch(so, batch)
    x_pool = so.s.transpose(-2, -1).matmul(x)  # :185-190
    if return_batched:
        x_pool = x_pool.unsqueeze(0)
    return x_pool, reduce_batch(so, batch)


def aggr_reduce(
    x: Tensor, so: OracleSelectOutput, op: str = "sum", batch: Optional[Tensor] = None
) -> Tuple[Tensor, Optional[Tensor]]:
    """AggrReduce.forward sparse path, tgp/reduce/aggr_reduce.py:13-29,99-105.

    Stable sort by cluster id, then the PyG Sum/Mean/Max/Min aggregation, which on
    CPU is ``scatter(reduce=op)`` over the sorted rows.
    """
    if batch is None and so.batch is not None:
        batch = so.batch
    if not so.is_sparse:
        raise ValueError("AggrReduce supports only sparse SelectOutput assignments.")
    src = x[so.node_index] * so.weight.view(-1, 1)
    index_sorted, perm = torch.sort(so.cluster_index, stable=True)
    x_pool = pyg.scatter(src[perm], index_sorted, dim=0, dim_size=so.num_supernodes, reduce=op)
    return x_pool, reduce_batch(so, batch)


def readout(x: Tensor, op: str = "sum", batch: Optional[Tensor] = None, size: Optional[int] = None):
    """AggrReduce readout mode (so=None), aggr_reduce.py:112-153."""
    if x.dim() == 3:
        B, N, _ = x.shape
        k = size if size is not None else B
        idx = torch.arange(B).repeat_interleave(N)
        return pyg.scatter(x.reshape(-1, x.size(-1)), idx, 0, k, op), torch.arange(k)
    if batch is None:
        return pyg.scatter(x, torch.zeros(x.size(0), dtype=torch.long), 0, 1, op), None
    k = size if size is not None else (int(batch.max()) + 1 if batch.numel() > 0 else 1)
    return pyg.scatter(x, batch, 0, k, op), torch.arange(k)


# --------------------------------------------------------------------------- #
# Sparse connect                                  tgp/connect/base_conn.py, tgp/utils/ops.py
# --------------------------------------------------------------------------- #
def check_and_filter_edge_weights(edge_weight: Optional[Tensor]) -> Optional[Tensor]:
    """ops.py:1043-1058."""
    if edge_weight is not None and edge_weight.ndim > 1:
        if edge_weight.ndim == 2 and edge_weight.size(-1) == 1:
            return edge_weight.flatten()
        raise RuntimeError(f"Edge weights must be of shape [E] or [E, 1], but got {edge_weight.shape}.")
    return edge_weight


def validate_edge_index(edge_index: Tensor) -> None:
    """The dense-tensor branch of connectivity_to_edge_index, ops.py:455-476."""
    if edge_index.dim() == 3 or (edge_index.dim() == 2 and edge_index.size(0) != 2):
        raise ValueError("Dense adjacency matrices are not supported by connectivity_to_edge_index().")
    if edge_index.dim() != 2:
        raise ValueError("connectivity_to_edge_index() expected edge_index with shape [2, E]")
    if edge_index.dtype != torch.int64:
        raise ValueError("connectivity_to_edge_index() expected edge_index indices to be an integer tensor")


def postprocess_adj_pool_sparse(
    edge_index: Tensor,
    edge_weight: Optional[Tensor],
    num_nodes: int,
    remove_self_loops: bool = False,
    degree_norm: bool = False,
    edge_weight_norm: bool = False,
    batch_pooled: Optional[Tensor] = None,
) -> Tuple[Tensor, Optional[Tensor]]:
    """ops.py:338-419 -- order: self loops, |w|<=eps filter, degree norm, max norm."""
    if remove_self_loops:
        edge_index, edge_weight = pyg.remove_self_loops(edge_index, edge_weight)
    if edge_weight is not None:
        edge_weight = edge_weight.view(-1)
        if edge_weight.numel() > 0:
            mask = edge_weight.abs() > EPS
            if not bool(torch.all(mask)):
                edge_index = edge_index[:, mask]
                edge_weight = edge_weight[mask]
    if degree_norm:
        if edge_weight is None:
            edge_weight = torch.ones(edge_index.size(1))
        deg = pyg.torch_scatter_scatter(edge_weight, edge_index[0], dim=0, dim_size=num_nodes, reduce="sum")
        dinv = deg.clamp(min=EPS).pow(-0.5)
        edge_weight = edge_weight * dinv[edge_index[0]] * dinv[edge_index[1]]
    if edge_weight_norm and edge_weight is not None:
        edge_batch = batch_pooled[edge_index[0]]
        mx = pyg.torch_scatter_scatter(edge_weight.abs(), edge_batch, dim=0, reduce="max")
        mx = torch.where(mx == 0, torch.ones_like(mx), mx)
        edge_weight = edge_weight / mx[edge_batch]
    return edge_index, edge_weight


def sparse_connect(
    edge_index: Tensor,
    edge_weight: Optional[Tensor] = None,
    node_index: Optional[Tensor] = None,
    cluster_index: Optional[Tensor] = None,
    num_nodes: Optional[int] = None,
    num_supernodes: Optional[int] = None,
    remove_self_loops: bool = True,
    reduce_op: str = "sum",
    edge_weight_norm: bool = False,
    batch_pooled: Optional[Tensor] = None,
    degree_norm: bool = False,
) -> Tuple[Tensor, Optional[Tensor]]:
    """base_conn.py:57-112 for ``[2,E]`` / torch-COO inputs."""
    to_coo = edge_index.is_sparse
    if to_coo:  # ops.py:449-454
        coo = edge_index
        edge_index, edge_weight = coo.indices().clone(), coo.values().clone()
    else:
        validate_edge_index(edge_index)
        edge_weight = check_and_filter_edge_weights(edge_weight)
    num_nodes = pyg.maybe_num_nodes(edge_index, num_nodes)
    if node_index is not None and len(node_index) < num_nodes:  # :79-82 kept-node path
        edge_index, edge_weight = pyg.subgraph(
            node_index, edge_index, edge_weight, relabel_nodes=True, num_nodes=num_nodes
        )
    elif cluster_index is not None and len(cluster_index) == num_nodes:  # :83-89 cluster path
        edge_index = cluster_index[edge_index]
        edge_index, edge_weight = pyg.coalesce(edge_index, edge_weight, num_nodes=num_supernodes, reduce=reduce_op)
    else:
        raise RuntimeError
    edge_index, edge_weight = postprocess_adj_pool_sparse(
        edge_index,
        edge_weight,
        num_nodes=num_supernodes,
        remove_self_loops=remove_self_loops,
        degree_norm=degree_norm,
        edge_weight_norm=edge_weight_norm,
        batch_pooled=batch_pooled,
    )
    if to_coo:  # :107-110 -> connectivity_to_torch_coo, ops.py:540-550
        if edge_weight is None:
            edge_weight = torch.ones(edge_index.size(1))
        edge_index = torch.sparse_coo_tensor(edge_index, edge_weight, (num_supernodes, num_supernodes)).coalesce()
        edge_weight = None
    return edge_index, edge_weight


def sparse_connect_so(edge_index, so: OracleSelectOutput, edge_weight=None, batch_pooled=None, **flags):
    """SparseConnect.forward, base_conn.py:162-215."""
    if flags.get("edge_weight_norm", False) and batch_pooled is None:
        raise AssertionError("edge_weight_norm=True but batch_pooled=None.")
    return sparse_connect(
        edge_index,
        edge_weight,
        node_index=so.node_index,
        cluster_index=so.cluster_index,
        num_nodes=so.num_nodes,
        num_supernodes=so.num_supernodes,
        batch_pooled=batch_pooled,
        **flags,
    )


# --------------------------------------------------------------------------- #
# Dense connect                                   tgp/connect/dense_conn.py, tgp/utils/ops.py
# --------------------------------------------------------------------------- #
def prepare_batched_dense_inputs(s: Tensor, adj: Tensor) -> Tuple[Tensor, Tensor]:
    """dense_conn.py:86-98."""
    if s.dim() == 2:
        s = s.unsqueeze(0)
    if adj.dim() == 2:
        adj = adj.unsqueeze(0)
    if s.dim() != 3 or adj.dim() != 3:
        raise ValueError("Expected batched dense inputs with 3 dimensions.")
    if s.size(0) != adj.size(0):
        raise ValueError("Assignment and adjacency batch sizes do not match")
    return s, adj


def dense_connect(adj: Tensor, s: Tensor) -> Tensor:
    """dense_conn.py:112-138: (S^T A) S with that association."""
    s, adj = prepare_batched_dense_inputs(s, adj)
    return torch.matmul(torch.matmul(s.transpose(-2, -1), adj), s)


def postprocess_adj_pool_dense(
    adj_pool: Tensor,
    remove_self_loops: bool = False,
    degree_norm: bool = False,
    adj_transpose: bool = False,
    edge_weight_norm: bool = False,
) -> Tensor:
    """ops.py:282-335.  The reference zeroes the diagonal IN PLACE (:308); the oracle
    clones first so callers can keep the raw tensor (values are identical)."""
    if remove_self_loops:
        adj_pool = adj_pool.clone()
        torch.diagonal(adj_pool, dim1=-2, dim2=-1)[:] = 0
    if degree_norm:
        d = adj_pool.sum(-2 if adj_transpose else -1, keepdim=True)
        d = torch.sqrt(d.clamp(min=EPS))
        adj_pool = (adj_pool / d) / d.transpose(-2, -1)
    if edge_weight_norm:
        B = adj_pool.size(0)
        mx = adj_pool.reshape(B, -1).abs().max(dim=1, keepdim=True)[0].unsqueeze(-1)
        mx = torch.where(mx == 0, torch.ones_like(mx), mx)
        adj_pool = adj_pool / mx
    return adj_pool


def dense_connect_forward(adj: Tensor, s: Tensor, **flags) -> Tensor:
    """DenseConnect._forward_batched_inputs, dense_conn.py:257-271."""
    return postprocess_adj_pool_dense(dense_connect(adj, s), **flags)


def dense_to_block_diag(adj_pool: Tensor) -> Tuple[Tensor, Tensor]:
    """ops.py:53-82."""
    if adj_pool.dim() == 2:
        adj_pool = adj_pool.unsqueeze(0)
    K = adj_pool.size(1)
    mask = adj_pool.abs() > EPS
    if not bool(mask.any()):
        return torch.empty((2, 0), dtype=torch.long), torch.empty((0,), dtype=adj_pool.dtype)
    b, r, c = mask.nonzero(as_tuple=True)
    off = b * K
    return torch.stack([r + off, c + off], 0), adj_pool[b, r, c]


def dense_connect_unbatched(
    edge_index: Tensor, edge_weight: Optional[Tensor], batch: Optional[Tensor], s: Tensor, batch_size: int
) -> Tensor:
    """DenseConnect._dense_connect_unbatched, dense_conn.py:141-208 (sparse A, S [N,K])."""
    N, K = s.size()
    if edge_weight is None:
        edge_weight = torch.ones(edge_index.size(1))
    edge_weight = edge_weight.view(-1)
    if batch_size == 1:
        if edge_index.size(1) == 0:
            return s.new_zeros((1, K, K))
        a = torch.sparse_coo_tensor(edge_index, edge_weight, (N, N)).coalesce()
        return s.t().matmul(torch.sparse.mm(a, s)).unsqueeze(0)
    s_list = pyg.unbatch(s, batch)
    if edge_index.size(1) == 0:
        return torch.stack([u.new_zeros((K, K)) for u in s_list], 0)
    adj_list = pyg.unbatch_edge_index(edge_index, batch)
    w_list = pyg.unbatch(edge_weight, batch[edge_index[0]])
    out = []
    for ei, u, w in zip(adj_list, s_list, w_list):
        a = torch.sparse_coo_tensor(ei, w, (u.size(0), u.size(0))).coalesce()
        out.append(u.t().matmul(torch.sparse.mm(a, u)))
    return torch.stack(out, 0)


# --------------------------------------------------------------------------- #
# Auxiliary losses                                tgp/utils/losses.py
# --------------------------------------------------------------------------- #
def _batch_reduce(loss: Tensor, how: str) -> Tensor:
    if how == "mean":
        return loss.mean(0)
    if how == "sum":
        return loss.sum(0)
    raise ValueError(f"Batch reduction {how} not allowed, must be one of ['mean', 'sum'].")


def mincut_loss(adj: Tensor, S: Tensor, adj_pooled: Tensor, batch_reduction: str = "mean") -> Tensor:
    """losses.py:39-84."""
    num = torch.einsum("ijj->i", adj_pooled)
    d = torch.diag_embed(adj.sum(-1))
    den = torch.einsum("ijj->i", torch.matmul(torch.matmul(S.transpose(-2, -1), d), S))
    return _batch_reduce(-(num / (den + EPS)), batch_reduction)


def orthogonality_loss(S: Tensor, batch_reduction: str = "mean") -> Tensor:
    """losses.py:87-123."""
    sts = torch.matmul(S.transpose(-2, -1), S)
    sts = sts / torch.norm(sts, dim=(-2, -1), keepdim=True)
    k = S.size(-1)
    eye = torch.eye(k, dtype=S.dtype) / math.sqrt(k)
    return _batch_reduce(torch.norm(sts - eye, dim=(-2, -1)), batch_reduction)


def link_pred_loss(S: Tensor, adj: Tensor, normalize_loss: bool = True) -> Tensor:
    """losses.py:644-679 -- ONE Frobenius norm over the whole batch tensor."""
    loss = torch.norm(adj - torch.matmul(S, S.transpose(1, 2)), p=2)
    if normalize_loss is True:
        loss = loss / adj.numel()
    return loss


def entropy_loss(S: Tensor, num_nodes: int) -> Tensor:
    """losses.py:682-708 -> unbatched_entropy_loss :476-500."""
    S2 = S.reshape(-1, S.size(-1))
    return (-(S2 * torch.log(S2 + EPS)).sum(dim=-1)).sum() / num_nodes


def sparse_mincut_loss(edge_index: Tensor, S: Tensor, edge_weight: Optional[Tensor] = None,
                       batch: Optional[Tensor] = None, batch_reduction: str = "mean") -> Tensor:
    """tgp/utils/losses.py:126-215 (unbatched MinCutPooling): -sum_e w_e <S_i, S_j> / (sum_i d_i |S_i|^2 + eps) per
    graph, d = row sums of the sparse adjacency."""
    n = S.size(0)
    w = torch.ones(edge_index.size(1), dtype=S.dtype) if edge_weight is None else edge_weight.view(-1)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long)
    B = int(batch.max()) + 1
    deg = pyg.scatter(w, edge_index[0], dim=0, dim_size=n, reduce="sum")
    den = pyg.scatter(deg * (S * S).sum(-1), batch, dim=0, dim_size=B, reduce="sum")
    contrib = w * (S[edge_index[0]] * S[edge_index[1]]).sum(-1)
    num = pyg.scatter(contrib, batch[edge_index[0]], dim=0, dim_size=B, reduce="sum")
    return _batch_reduce(-(num / (den + EPS)), batch_reduction)


def unbatched_orthogonality_loss(S: Tensor, batch: Optional[Tensor] = None, batch_reduction: str = "mean") -> Tensor:
    """tgp/utils/losses.py:319-389: per graph || S_g^T S_g / ||S_g^T S_g||_F - I / sqrt(K) ||_F."""
    K = S.size(1)
    if batch is None:
        batch = torch.zeros(S.size(0), dtype=torch.long)
    id_k = torch.eye(K, dtype=S.dtype) / math.sqrt(K)
    out = []
    for g in range(int(batch.max()) + 1):
        sg = S[batch == g]
        sts = sg.t() @ sg
        out.append(torch.norm(sts / torch.norm(sts) - id_k))
    return _batch_reduce(torch.stack(out), batch_reduction)


def sparse_link_pred_loss(S: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None,
                          batch: Optional[Tensor] = None, normalize_loss: bool = True) -> Tensor:
    """tgp/utils/losses.py:711-777: sqrt(sum_e (w - ss)^2 + sum_g ||S_g^T S_g||_F^2 - sum_e ss^2), ss = <S_i, S_j>."""
    w = torch.ones(edge_index.size(1), dtype=S.dtype) if edge_weight is None else edge_weight.view(-1).to(S.dtype)
    if batch is None:
        batch = torch.zeros(S.size(0), dtype=torch.long)
    ss = (S[edge_index[0]] * S[edge_index[1]]).sum(-1)
    total, numel = S.new_zeros(()), 0
    for g in range(int(batch.max()) + 1):
        sg = S[batch == g]
        sts = sg.t() @ sg
        total = total + (sts * sts).sum()
        numel += sg.size(0) ** 2
    link = torch.sqrt(torch.clamp(((w - ss) ** 2).sum() + total - (ss ** 2).sum(), min=0.0))
    return link / numel if (normalize_loss and numel > 0) else link



# --------------------------------------------------------------------------- #
# Orchestration (call order of the poolers)
# --------------------------------------------------------------------------- #
def topk_select(
    x: Tensor,
    weight: Optional[Tensor],
    ratio: float = 0.5,
    batch: Optional[Tensor] = None,
    act=torch.tanh,
    min_score: Optional[float] = None,
) -> OracleSelectOutput:
    """TopkSelect.forward, tgp/select/topk_select.py:163-203."""
    if batch is None:
        batch = x.new_zeros(x.size(0), dtype=torch.long)
    if weight is None:
        score = x if x.dim() == 1 else x.view(-1)
    else:
        xx = x.view(-1, 1) if x.dim() == 1 else x
        score = (xx * weight).sum(dim=-1)
        if min_score is None:
            score = score / weight.norm(p=2, dim=-1)
    score = act(score) if min_score is None else pyg.softmax(score, batch)
    node_index = pyg.topk(score, ratio, batch, min_score)
    return OracleSelectOutput(
        node_index=node_index,
        num_nodes=x.size(0),
        cluster_index=torch.arange(node_index.size(0)),
        num_supernodes=node_index.size(0),
        weight=score[node_index],
    )


def topk_pool(x, edge_index, edge_weight, so: OracleSelectOutput, batch=None, multiplier: float = 1.0, **flags):
    """TopkPooling.forward after select, tgp/poolers/topk.py:171-190."""
    x_pool, batch_pool = base_reduce(x, so, batch=batch)
    if multiplier != 1:
        x_pool = multiplier * x_pool
    ei, ew = sparse_connect_so(edge_index, so, edge_weight=edge_weight, batch_pooled=batch_pool, **flags)
    return x_pool, ei, ew, batch_pool


def mincut_pool(
    x: Tensor,
    adj: Tensor,
    s: Tensor,
    cut_loss_coeff: float = 1.0,
    ortho_loss_coeff: float = 1.0,
    remove_self_loops: bool = True,
    degree_norm: bool = True,
    adj_transpose: bool = True,
    edge_weight_norm: bool = False,
) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """MinCutPooling.forward batched path after select, tgp/poolers/mincut.py:219-237,291-310.

    Loss is computed from the RAW ``S^T A S`` before the diagonal is zeroed."""
    so = OracleSelectOutput(s=s)
    x_pool, _ = base_reduce(x, so)
    adj_pool = dense_connect(adj, s)
    loss = {
        "cut_loss": mincut_loss(adj, s, adj_pool, "mean") * cut_loss_coeff,
        "ortho_loss": orthogonality_loss(s, "mean") * ortho_loss_coeff,
    }
    adj_pool = postprocess_adj_pool_dense(
        adj_pool,
        remove_self_loops=remove_self_loops,
        degree_norm=degree_norm,
        adj_transpose=adj_transpose,
        edge_weight_norm=edge_weight_norm,
    )
    return x_pool, adj_pool, loss


def diff_pool(
    x: Tensor,
    adj: Tensor,
    s: Tensor,
    num_nodes: Optional[int] = None,
    link_loss_coeff: float = 1.0,
    ent_loss_coeff: float = 1.0,
    normalize_loss: bool = False,
    remove_self_loops: bool = True,
    degree_norm: bool = True,
    adj_transpose: bool = True,
    edge_weight_norm: bool = False,
) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """DiffPool.forward batched path after select, tgp/poolers/diffpool.py:208-218,262-284."""
    so = OracleSelectOutput(s=s)
    x_pool, _ = base_reduce(x, so)
    adj_pool = dense_connect_forward(
        adj,
        s,
        remove_self_loops=remove_self_loops,
        degree_norm=degree_norm,
        adj_transpose=adj_transpose,
        edge_weight_norm=edge_weight_norm,
    )
    if num_nodes is None:
        num_nodes = s.size(0) * s.size(1)
    loss = {
        "link_loss": link_pred_loss(s, adj, normalize_loss=normalize_loss) * link_loss_coeff,
        "entropy_loss": entropy_loss(s, num_nodes) * ent_loss_coeff,
    }
    return x_pool, adj_pool, loss


# --------------------------------------------------------------------------- #
# Sparse output of the dense poolers              tgp/src.py:500-557, tgp/utils/ops.py:85-132
# --------------------------------------------------------------------------- #
def out_mask_from_dense_s(s: Tensor) -> Tensor:
    """get_mask_from_dense_s for [B, N, K] (ops.py:119-121)."""
    return s.sum(dim=-2) > 0


def finalize_sparse_output(x_pool: Tensor, adj_pool: Tensor, batch, batch_pooled, out_mask):
    """DenseSRCPooling._finalize_sparse_output, src.py:500-557."""
    B, K = adj_pool.size(0), adj_pool.size(1)
    x_flat = x_pool.reshape(-1, x_pool.size(-1))
    if batch_pooled is None and batch is not None:
        batch_pooled = torch.arange(int(batch.max()) + 1).repeat_interleave(K)
    if batch_pooled is None and B > 1:
        batch_pooled = torch.arange(B).repeat_interleave(K)
    if batch_pooled is None and out_mask is not None:
        batch_pooled = torch.zeros(B * K, dtype=torch.long)
    if out_mask is not None:
        valid_flat = out_mask.reshape(-1)
        valid_indices = valid_flat.nonzero(as_tuple=True)[0]
        x_out = x_flat[valid_indices]
        batch_pooled = batch_pooled[valid_flat]
        adj_masked = adj_pool * out_mask.unsqueeze(-1).to(adj_pool.dtype) * out_mask.unsqueeze(-2).to(adj_pool.dtype)
        ei, ew = dense_to_block_diag(adj_masked)
        old_to_new = torch.full((B * K,), -1, dtype=torch.long)
        old_to_new[valid_indices] = torch.arange(valid_indices.numel())
        keep = (old_to_new[ei[0]] >= 0) & (old_to_new[ei[1]] >= 0)
        ei = torch.stack([old_to_new[ei[0][keep]], old_to_new[ei[1][keep]]], 0)
        ew = ew[keep]
    else:
        ei, ew = dense_to_block_diag(adj_pool)
        x_out = x_flat
    return x_out, ei, ew, batch_pooled


def dense_connect_forward_unbatched(edge_index, edge_weight, batch, s, batch_pooled=None, remove_self_loops=True,
                                    degree_norm=True, edge_weight_norm=False, sparse_output=False):
    """DenseConnect._forward_unbatched_inputs, dense_conn.py:273-354."""
    batch_size = 1 if batch is None else int(batch.max()) + 1
    K = s.size(-1)
    raw = dense_connect_unbatched(edge_index, edge_weight, batch, s, batch_size)
    if not sparse_output:
        return postprocess_adj_pool_dense(raw, remove_self_loops=remove_self_loops, degree_norm=degree_norm,
                                          adj_transpose=False, edge_weight_norm=edge_weight_norm), None
    ei, ew = dense_to_block_diag(raw)
    return postprocess_adj_pool_sparse(ei, ew, num_nodes=batch_size * K, remove_self_loops=remove_self_loops,
                                       degree_norm=degree_norm, edge_weight_norm=edge_weight_norm,
                                       batch_pooled=batch_pooled)
