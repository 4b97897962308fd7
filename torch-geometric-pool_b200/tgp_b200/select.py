"""TopK selection on B200 kernels (tgp/select/topk_select.py:163-203).

The score projection stays in torch (it is differentiable and feeds ``so.weight`` -> the projection vector);
the selection itself (per-graph top-k with the reference tie rule, plus the node-sorted SelectOutput layout)
runs as one radix sort + two compactions in ``libtgp_b200.so``.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple, Union

import torch
from torch import Tensor

from . import _lib as L
from .select_output import SelectOutput


def topk(score: Tensor, ratio: float, batch: Optional[Tensor] = None, num_graphs: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """``(node_index ascending, cluster_index)`` of the per-graph top ``ceil(ratio * n_g)`` scores."""
    if not score.is_cuda:
        raise RuntimeError("tgp_b200 runs on CUDA tensors only (no CPU fallback by design)")
    score = score.detach().to(torch.float32).contiguous().view(-1)
    N, dev = score.numel(), score.device
    if batch is not None:
        batch = batch.contiguous()
        if num_graphs is None:
            num_graphs = int(batch.max().item()) + 1 if N > 0 else 1
    else:
        num_graphs = 1
    node_index = torch.empty(max(N, 1), dtype=torch.long, device=dev)
    cluster_index = torch.empty(max(N, 1), dtype=torch.long, device=dev)
    count = torch.empty(1, dtype=torch.long, device=dev)
    ws = L.workspace(L.load().tgpb200_topk_select_workspace_bytes(N, num_graphs), dev)
    L.call("tgpb200_topk_select", L.ptr(score), L.ptr(batch), N, num_graphs, float(ratio), L.ptr(node_index),
           L.ptr(cluster_index), L.ptr(count), L.ptr(ws), ws.numel(), L.stream())
    k = int(count.item())
    return node_index[:k], cluster_index[:k]


def topk_select(
    x: Tensor,
    weight: Optional[Tensor] = None,
    ratio: float = 0.5,
    batch: Optional[Tensor] = None,
    act: Union[str, Callable] = "tanh",
    num_graphs: Optional[int] = None,
) -> SelectOutput:
    """TopkSelect.forward (ratio mode): score = act((x . p) / ||p||), per-graph top-k, SelectOutput with
    ``weight = score[node_index]`` (differentiable through torch)."""
    if weight is None:
        score = x if x.dim() == 1 else x.view(-1)
    else:
        xx = x.view(-1, 1) if x.dim() == 1 else x
        score = (xx * weight).sum(dim=-1) / weight.norm(p=2, dim=-1)
    if isinstance(act, str):
        act = {"tanh": torch.tanh, "sigmoid": torch.sigmoid, "relu": torch.relu, "linear": lambda v: v,
               "identity": lambda v: v}[act]
    score = act(score)
    node_index, cluster_index = topk(score, ratio, batch, num_graphs)
    s = torch.sparse_coo_tensor(torch.stack([node_index, cluster_index]), score[node_index],
                                (x.size(0), node_index.numel()), is_coalesced=True, check_invariants=False)
    return SelectOutput(s=s, _trusted=True)  # node_index ascending by construction
