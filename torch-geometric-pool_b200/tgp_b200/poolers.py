"""Pooler-level orchestration of the path (what the pooler ``forward``s do after ``select``).

``mincut_pool`` / ``diff_pool`` run reduce + connect + auxiliary losses + post-processing as ONE
fused forward and ONE fused backward; ``topk_pool`` / ``cluster_pool`` chain the sparse reduce and
the sparse connect.  ``patch_pooler`` installs the B200 operators on a reference pooler object.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import functional as F_
from .connect import B200DenseConnect, B200SparseConnect
from .reduce import B200Reduce


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
    """MinCutPooling.forward batched path after select (tgp/poolers/mincut.py:219-237, 291-310)."""
    x_pool, adj_pool, losses = F_.dense_pool(
        x, adj, s, remove_self_loops=remove_self_loops, degree_norm=degree_norm, adj_transpose=adj_transpose,
        edge_weight_norm=edge_weight_norm, loss_kind=F_.LOSS_MINCUT,
    )
    loss = {"cut_loss": losses[0] * cut_loss_coeff, "ortho_loss": losses[1] * ortho_loss_coeff}
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
    """DiffPool.forward batched path after select (tgp/poolers/diffpool.py:208-218, 262-284)."""
    if num_nodes is None:
        num_nodes = s.size(0) * s.size(1)
    x_pool, adj_pool, losses = F_.dense_pool(
        x, adj, s, remove_self_loops=remove_self_loops, degree_norm=degree_norm, adj_transpose=adj_transpose,
        edge_weight_norm=edge_weight_norm, loss_kind=F_.LOSS_DIFFPOOL,
        link_div=float(adj.numel()) if normalize_loss else 1.0, ent_div=float(num_nodes),
    )
    loss = {"link_loss": losses[2] * link_loss_coeff, "entropy_loss": losses[3] * ent_loss_coeff}
    return x_pool, adj_pool, loss


def sparse_pool(
    x: Tensor,
    edge_index: Tensor,
    so,
    edge_weight: Optional[Tensor] = None,
    batch: Optional[Tensor] = None,
    reduce_op: str = "sum",
    multiplier: float = 1.0,
    connect_op: str = "sum",
    remove_self_loops: bool = True,
    degree_norm: bool = False,
    edge_weight_norm: bool = False,
):
    """TopkPooling / GraclusPooling forward after select (tgp/poolers/topk.py:171-190,
    tgp/poolers/graclus.py:91-156): reduce, then connect."""
    x_pool, batch_pool = B200Reduce(reduce_op)(x, so, batch=batch)
    if multiplier != 1:
        x_pool = multiplier * x_pool
    conn = B200SparseConnect(connect_op, remove_self_loops, edge_weight_norm, degree_norm)
    ei, ew = conn(edge_index, so, edge_weight=edge_weight, batch_pooled=batch_pool)
    return x_pool, ei, ew, batch_pool


def patch_pooler(pooler, reduce_op: str = "sum"):
    """Swap ``pooler.reducer`` / ``pooler.connector`` of a reference ``tgp`` pooler for the B200
    operators, preserving the connector's ctor attributes (they are read by the poolers,
    tgp/poolers/mincut.py:233-236)."""
    conn = getattr(pooler, "connector", None)
    if conn is not None:
        name = type(conn).__name__
        if name == "SparseConnect":
            pooler.connector = B200SparseConnect(
                conn.reduce_op, conn.remove_self_loops, conn.edge_weight_norm, conn.degree_norm
            )
        elif name == "DenseConnect":
            pooler.connector = B200DenseConnect(
                conn.remove_self_loops, conn.degree_norm, conn.adj_transpose, conn.edge_weight_norm, conn.sparse_output
            )
    if getattr(pooler, "reducer", None) is not None:
        pooler.reducer = B200Reduce(reduce_op)
    return pooler
