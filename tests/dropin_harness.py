"""A stand-in for the reference's SRC pooler classes (torch_geometric is not installable offline, so ``tgp`` itself
cannot be imported): same class names, same ctor attributes, and the exact keyword forwarding of
tgp/src.py:189-229 (``reduce(**kw) -> self.reducer(**kw)``, ``connect(**kw) -> self.connector(**kw)``),
tgp/poolers/topk.py:166-190 and tgp/poolers/mincut.py:210-258 / diffpool.py:200-237.  ``tgp_b200.patch_pooler``
is exercised on these objects the way it would be on the real ones (it dispatches on class names and attributes)."""
import torch
from torch import nn


class SparseConnect(nn.Module):  # attributes of tgp/connect/base_conn.py:139-152
    def __init__(self, reduce_op="sum", remove_self_loops=True, edge_weight_norm=False, degree_norm=False):
        super().__init__()
        self.reduce_op, self.remove_self_loops = reduce_op, remove_self_loops
        self.edge_weight_norm, self.degree_norm = edge_weight_norm, degree_norm

    def forward(self, **kw):
        raise AssertionError("the reference connector must have been replaced")


class DenseConnect(nn.Module):  # attributes of tgp/connect/dense_conn.py:64-84
    def __init__(self, remove_self_loops=True, degree_norm=True, adj_transpose=True, edge_weight_norm=False,
                 sparse_output=False):
        super().__init__()
        self.remove_self_loops, self.degree_norm, self.adj_transpose = remove_self_loops, degree_norm, adj_transpose
        self.edge_weight_norm, self.sparse_output = edge_weight_norm, sparse_output

    def forward(self, **kw):
        raise AssertionError("the reference connector must have been replaced")


class BaseReduce(nn.Module):
    def forward(self, **kw):
        raise AssertionError("the reference reducer must have been replaced")


class MeanAggregation(nn.Module):
    pass


class LSTMAggregation(nn.Module):
    pass


class AggrReduce(nn.Module):  # tgp/reduce/aggr_reduce.py:53-62
    def __init__(self, aggr):
        super().__init__()
        self.aggr = aggr

    def forward(self, **kw):
        return "reference AggrReduce ran"


class SRCPooling(nn.Module):  # tgp/src.py:150-229
    def __init__(self, selector=None, reducer=None, lifter=None, connector=None):
        super().__init__()
        self.selector, self.reducer, self.lifter, self.connector = selector, reducer, lifter, connector

    def select(self, **kwargs):
        return self.selector(**kwargs)

    def reduce(self, **kwargs):
        return self.reducer(**kwargs)

    def connect(self, **kwargs):
        return self.connector(**kwargs)


class TopkPooling(SRCPooling):  # tgp/poolers/topk.py:166-190
    multiplier = 1.0

    def forward(self, x, adj=None, edge_weight=None, so=None, batch=None, attn=None, lifting=False, **kwargs):
        so = self.select(x=x if attn is None else attn, batch=batch)
        x_pooled, batch_pooled = self.reduce(x=x, so=so, batch=batch)
        x_pooled = self.multiplier * x_pooled if self.multiplier != 1 else x_pooled
        edge_index_pooled, edge_weight_pooled = self.connect(so=so, edge_index=adj, edge_weight=edge_weight,
                                                             batch_pooled=batch_pooled)
        return x_pooled, edge_index_pooled, edge_weight_pooled, batch_pooled, so


class _DenseBase(SRCPooling):
    batched, sparse_output, cache_preprocessing = True, False, False

    def _ensure_batched_inputs(self, x, edge_index, edge_weight, batch, mask, use_cache=None):  # tgp/src.py:454-491
        x = x.unsqueeze(0) if x.dim() == 2 else x
        if mask is None:
            mask = x.new_ones(x.size(0), x.size(1), dtype=torch.bool)
        return x, edge_index, mask

    def forward(self, **kw):
        raise AssertionError("the reference forward must have been replaced by the fused one")


class MinCutPooling(_DenseBase):
    cut_loss_coeff, ortho_loss_coeff = 0.7, 1.3


class DiffPool(_DenseBase):
    link_loss_coeff, ent_loss_coeff, normalize_loss = 0.5, 2.0, False
