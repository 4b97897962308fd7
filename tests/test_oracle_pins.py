"""CPU: pin the oracle against the hand values the reference's own tests hold.

Each test names the reference test (relative to /root/reference) it re-types.
"""
import math

import pytest
import torch

from oracle import pyg_shim as pyg
from oracle import ref_path as R


def test_dense_connect_hand_matrix():
    # tests/connect/test_dense_conn.py:210-232
    s = torch.tensor([[1.0, 0.0], [0.0, 1.0], [1.0, 0.0]])
    adj = torch.tensor([[0.0, 1.0, 2.0], [1.0, 0.0, 3.0], [2.0, 3.0, 0.0]])
    out = R.dense_connect(adj, s)
    assert out.shape == (1, 2, 2)
    assert torch.equal(out[0], torch.tensor([[4.0, 4.0], [4.0, 0.0]]))


def test_dense_postprocess_diag_and_maxnorm():
    # tests/connect/test_dense_conn.py:39-107
    torch.manual_seed(0)
    a = torch.rand(2, 5, 5)
    a = a + a.transpose(1, 2)
    s = torch.softmax(torch.randn(2, 5, 3), -1)
    raw = R.dense_connect(a, s)
    out = R.postprocess_adj_pool_dense(raw, remove_self_loops=True)
    assert torch.all(torch.diagonal(out, dim1=-2, dim2=-1) == 0)
    out = R.postprocess_adj_pool_dense(raw, edge_weight_norm=True)
    torch.testing.assert_close(out, raw / raw.reshape(2, -1).abs().max(1)[0].view(2, 1, 1), atol=1e-6, rtol=0)


@pytest.mark.parametrize("op", ["sum", "mean", "max", "min"])
def test_readout_hand_values(op):
    # tests/reduce/test_global_reduce.py:9-37
    x = torch.tensor([[[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]], [[-1.0, 0.0], [0.0, 1.0], [2.0, -2.0]]])
    out, _ = R.readout(x, op)
    exp = {
        "sum": [[9.0, 12.0], [1.0, -1.0]],
        "mean": [[9.0 / 3.0, 12.0 / 3.0], [1.0 / 3.0, -1.0 / 3.0]],
        "max": [[5.0, 6.0], [2.0, 1.0]],
        "min": [[1.0, 2.0], [-1.0, -2.0]],
    }[op]
    assert torch.equal(out, torch.tensor(exp))


def test_aggr_sum_equals_base_reduce():
    # tests/reduce/test_aggr_reduce.py:42-84
    torch.manual_seed(0)
    x = torch.randn(9, 4)
    cluster = torch.tensor([2, 0, 1, 1, 0, 2, 2, 1, 0])
    so = R.OracleSelectOutput(cluster_index=cluster, num_supernodes=3, weight=torch.rand(9))
    a, _ = R.base_reduce(x, so)
    b, _ = R.aggr_reduce(x, so, "sum")
    torch.testing.assert_close(a, b)


def test_tiny_weight_filter():
    # tests/utils/test_ops.py:254-269
    ei = torch.tensor([[0, 1], [1, 0]])
    ew = torch.tensor([0.0, 1.0])
    oi, ow = R.postprocess_adj_pool_sparse(ei, ew, num_nodes=2)
    assert oi.shape == (2, 1) and torch.equal(ow, torch.tensor([1.0]))


def test_topk_selection_set_and_k():
    # tests/poolers/test_topk.py:22-34, 60-63
    x = torch.arange(1.0, 6.0).unsqueeze(-1)
    so = R.topk_select(x, None, ratio=0.5, act=lambda v: v)
    assert torch.equal(so.node_index.sort(descending=True)[0], torch.tensor([4, 3, 2]))
    torch.manual_seed(0)
    so = R.topk_select(torch.randn(6, 4), torch.randn(1, 4), ratio=0.5)
    assert so.num_supernodes == 3 == math.ceil(0.5 * 6)


def test_select_output_from_cluster_index():
    # tests/selection/test_base_select.py:26-67
    cluster = torch.tensor([0, 1, 0, 2, 1])
    so = R.OracleSelectOutput(cluster_index=cluster)
    assert torch.equal(so.node_index, torch.arange(5))
    assert torch.equal(so.cluster_index, cluster)
    assert torch.equal(so.weight, torch.ones(5))
    assert so.num_supernodes == 3


def test_identity_pooling_keeps_edge_set():
    # tests/poolers/test_nopool.py:99-135
    torch.manual_seed(0)
    n = 7
    ei = torch.tensor([[0, 1, 1, 2, 3, 4, 5, 6, 2], [1, 0, 2, 1, 4, 3, 6, 5, 5]])
    x = torch.randn(n, 3)
    so = R.OracleSelectOutput(cluster_index=torch.arange(n))
    xp, _ = R.base_reduce(x, so)
    assert torch.equal(xp, x)
    eo, wo = R.sparse_connect_so(ei, so, remove_self_loops=False)
    assert wo is None
    key_in = (ei[0] * n + ei[1]).sort()[0]
    assert torch.equal(eo[0] * n + eo[1], key_in)  # coalesce returns them sorted


def test_cluster_path_equals_dense_stas():
    # cross-check: cluster path densified == S^T A S with one-hot S
    torch.manual_seed(1)
    n, k, e = 20, 6, 80
    ei = torch.randint(0, n, (2, e))
    ew = torch.rand(e) + 0.1
    cluster = torch.randint(0, k, (n,))
    so = R.OracleSelectOutput(cluster_index=cluster, num_supernodes=k)
    eo, wo = R.sparse_connect_so(ei, so, edge_weight=ew, remove_self_loops=False)
    dense = torch.zeros(k, k).index_put_((eo[0], eo[1]), wo)
    a = torch.zeros(n, n).index_put_((ei[0], ei[1]), ew, accumulate=True)
    s = torch.nn.functional.one_hot(cluster, k).float()
    torch.testing.assert_close(dense, R.dense_connect(a, s)[0], rtol=1e-5, atol=1e-6)
    key = eo[0] * k + eo[1]
    assert torch.all(key[1:] > key[:-1])  # lexicographic, unique


def test_kept_node_path_equals_submatrix():
    # cross-check: kept-node path densified == A[idx][:, idx]
    torch.manual_seed(2)
    n, e = 15, 60
    ei = torch.randint(0, n, (2, e))
    ei = ei[:, ei[0] != ei[1]]
    ei = torch.unique(ei, dim=1)
    ew = torch.rand(ei.size(1)) + 0.1
    keep = torch.tensor([11, 2, 7, 3, 9, 0])
    so = R.OracleSelectOutput(node_index=keep, num_nodes=n, cluster_index=torch.arange(6), num_supernodes=6)
    eo, wo = R.sparse_connect_so(ei, so, edge_weight=ew, remove_self_loops=False)
    a = torch.zeros(n, n).index_put_((ei[0], ei[1]), ew)
    idx = keep.sort()[0]
    dense = torch.zeros(6, 6).index_put_((eo[0], eo[1]), wo)
    assert torch.equal(dense, a[idx][:, idx])


def test_topk_quirk_row_numbering_vs_edge_numbering():
    # x_pool rows follow cluster_index (score rank), edges follow
    # position in the ascending-sorted node_index.
    x = torch.tensor([[0.1], [0.9], [0.2], [0.95], [0.3], [0.5]])
    so = R.topk_select(x, None, ratio=0.5, act=lambda v: v)
    # topk order: nodes 3, 1, 5 -> clusters 0,1,2; sorted node_index [1,3,5] -> clusters [1,0,2]
    assert torch.equal(so.node_index, torch.tensor([1, 3, 5]))
    assert torch.equal(so.cluster_index, torch.tensor([1, 0, 2]))
    xp, _ = R.base_reduce(x, so)
    torch.testing.assert_close(xp.view(-1), torch.tensor([0.95 * 0.95, 0.9 * 0.9, 0.25]))
    ei = torch.tensor([[1, 3], [3, 5]])
    eo, _ = R.sparse_connect_so(ei, so)
    assert torch.equal(eo, torch.tensor([[0, 1], [1, 2]]))


def test_coalesce_rules():
    ei = torch.tensor([[1, 0, 1, 0], [0, 1, 0, 1]])
    w = torch.tensor([1.0, 2.0, 3.0, 4.0])
    e, ww = pyg.coalesce(ei, w, num_nodes=2, reduce="sum")
    assert torch.equal(e, torch.tensor([[0, 1], [1, 0]])) and torch.equal(ww, torch.tensor([6.0, 4.0]))
    e, ww = pyg.coalesce(ei, w, num_nodes=2, reduce="mean")
    assert torch.equal(ww, torch.tensor([3.0, 2.0]))
    e, ww = pyg.coalesce(ei, None, num_nodes=2)
    assert ww is None and e.size(1) == 2


def test_errors():
    so = R.OracleSelectOutput(cluster_index=torch.tensor([0, 1, 0]))
    with pytest.raises(ValueError):
        R.base_reduce(torch.randn(3, 2), so, return_batched=True)
    with pytest.raises(ValueError):
        R.sparse_connect_so(torch.zeros(2, 3, dtype=torch.int32), so)
    with pytest.raises(RuntimeError):
        R.sparse_connect_so(torch.zeros(2, 3, dtype=torch.long), so, edge_weight=torch.ones(3, 2))
    with pytest.raises(AssertionError):
        R.sparse_connect_so(torch.zeros(2, 3, dtype=torch.long), so, edge_weight_norm=True)
    with pytest.raises(ValueError):
        R.dense_connect(torch.randn(3, 4, 4), torch.randn(2, 4, 2))
    with pytest.raises(RuntimeError):  # neither branch applies (base_conn.py:90-91)
        R.sparse_connect(torch.tensor([[0, 2], [2, 0]]), None, node_index=None,
                         cluster_index=torch.tensor([0, 1]), num_nodes=3, num_supernodes=2)
