"""CPU: the oracle restatement must reproduce the reference-generated golden vectors.

Fixtures come from the reference's own source files (tests/golden/make_golden.py).
Integer outputs are compared bit-exactly; floats at rtol 1e-6.
"""
import pytest
import torch

from oracle import ref_path as R

RT = dict(rtol=1e-6, atol=1e-7)


def _names(golden, prefix):
    return sorted(k for k in golden if k.startswith(prefix))


def _flags(name):
    parts = {p[:-1]: bool(int(p[-1])) for p in name.split("_") if p[:-1] in ("dn", "ewn", "rsl", "t")}
    return parts


def test_golden_has_all_families(golden):
    assert len(_names(golden, "topk_")) == 16
    assert len(_names(golden, "cluster_")) == 40
    assert len(_names(golden, "dense_")) == 16


def test_topk_cases(golden):
    for name in _names(golden, "topk_"):
        c = golden[name]
        f = _flags(name)
        x = c["x"].clone().requires_grad_(True)
        p = c["p"].clone().requires_grad_(True)
        w = None if c["edge_weight"] is None else c["edge_weight"].clone().requires_grad_(True)
        so = R.topk_select(x, p, ratio=0.5, batch=c["batch"])
        assert torch.equal(so.node_index, c["node_index"]), name
        assert torch.equal(so.cluster_index, c["cluster_index"]), name
        torch.testing.assert_close(so.weight, c["weight"], **RT)
        xp, ei, ew, bp = R.topk_pool(
            x, c["edge_index"], w, so, batch=c["batch"],
            remove_self_loops=f["rsl"], degree_norm=f["dn"], edge_weight_norm=f["ewn"],
        )
        assert torch.equal(ei, c["edge_index_out"]), name
        assert torch.equal(bp, c["batch_pool"]), name
        torch.testing.assert_close(xp, c["x_pool"], **RT)
        if c["edge_weight_out"] is None:
            assert ew is None
        else:
            torch.testing.assert_close(ew, c["edge_weight_out"], **RT)
        loss = xp.square().sum()
        if ew is not None and ew.requires_grad:
            loss = loss + (ew * torch.arange(1, ew.numel() + 1)).sum()
        loss.backward()
        torch.testing.assert_close(x.grad, c["grad_x"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(p.grad, c["grad_p"], rtol=1e-5, atol=1e-6)
        if c["grad_w"] is not None:
            torch.testing.assert_close(w.grad, c["grad_w"], rtol=1e-5, atol=1e-6)


def test_cluster_cases(golden):
    for name in _names(golden, "cluster_"):
        c = golden[name]
        f = _flags(name)
        op = name.split("_")[1]
        x = c["x"].clone().requires_grad_(True)
        w = None if c["edge_weight"] is None else c["edge_weight"].clone().requires_grad_(True)
        so = R.OracleSelectOutput(cluster_index=c["cluster"], num_nodes=x.size(0), num_supernodes=c["num_supernodes"])
        xp, bp = R.base_reduce(x, so, batch=c["batch"])
        ei, ew = R.sparse_connect_so(
            c["edge_index"], so, edge_weight=w, batch_pooled=c["batch_pooled"], reduce_op=op,
            remove_self_loops=f["rsl"], degree_norm=f["dn"], edge_weight_norm=f["ewn"],
        )
        assert torch.equal(ei, c["edge_index_out"]), name
        assert torch.equal(bp, c["batch_pool"]), name
        torch.testing.assert_close(xp, c["x_pool"], **RT)
        if c["edge_weight_out"] is None:
            assert ew is None
        else:
            torch.testing.assert_close(ew, c["edge_weight_out"], **RT)
        loss = xp.square().sum()
        if ew is not None and ew.requires_grad:
            loss = loss + (ew * torch.arange(1, ew.numel() + 1)).sum()
        loss.backward()
        torch.testing.assert_close(x.grad, c["grad_x"], rtol=1e-5, atol=1e-6)
        if c["grad_w"] is not None:
            torch.testing.assert_close(w.grad, c["grad_w"], rtol=1e-5, atol=1e-6)


def test_dense_cases(golden):
    for name in _names(golden, "dense_"):
        c = golden[name]
        f = _flags(name)
        sr = c["s_raw"].clone().requires_grad_(True)
        x = c["x"].clone().requires_grad_(True)
        a = c["adj"].clone().requires_grad_(True)
        s = torch.softmax(sr, -1) * c["mask"][..., None]
        xp, _ = R.base_reduce(x, R.OracleSelectOutput(s=s))
        raw = R.dense_connect(a, s)
        cut = R.mincut_loss(a, s, raw)
        ortho = R.orthogonality_loss(s)
        link = R.link_pred_loss(s, a, normalize_loss=False)
        link_n = R.link_pred_loss(s, a, normalize_loss=True)
        ent = R.entropy_loss(s, int(c["mask"].sum()))
        post = R.postprocess_adj_pool_dense(
            raw, remove_self_loops=f["rsl"], degree_norm=f["dn"], adj_transpose=f["t"], edge_weight_norm=f["ewn"]
        )
        for got, key in ((xp, "x_pool"), (raw, "adj_pool_raw"), (post, "adj_pool"), (cut, "cut"), (ortho, "ortho"),
                         (link, "link"), (link_n, "link_norm"), (ent, "ent")):
            torch.testing.assert_close(got, c[key], **RT)
        wts = torch.arange(1, post.numel() + 1, dtype=torch.float).view_as(post) / post.numel()
        (xp.square().sum() + (post * wts).sum() + cut + ortho + 0.5 * link + 0.25 * ent).backward()
        torch.testing.assert_close(sr.grad, c["grad_s_raw"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(x.grad, c["grad_x"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(a.grad, c["grad_adj"], rtol=1e-5, atol=1e-6)


def test_oracle_unbatched_losses_and_connect_match_reference_golden():
    """tests/golden/ref_unbatched.pt was produced by the reference's own DenseConnect (unbatched) and sparse loss
    functions (make_golden_unbatched.py); the oracle restatements must reproduce it."""
    import os

    import torch

    from oracle import ref_path as R

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_unbatched.pt")
    cases = torch.load(path, weights_only=False)
    assert set(cases) == {"ragged_w", "ragged_now", "single_w"}
    for name, c in cases.items():
        s = torch.softmax(c["s_raw"], -1)
        ei, ew, b = c["edge_index"], c["edge_weight"], c["batch"]
        assert torch.equal(R.sparse_mincut_loss(ei, s, ew, b), c["cut"]), name
        assert torch.equal(R.unbatched_orthogonality_loss(s, b), c["ortho"]), name
        assert torch.equal(R.sparse_link_pred_loss(s, ei, ew, b, normalize_loss=False), c["link"]), name
        assert torch.equal(R.sparse_link_pred_loss(s, ei, ew, b, normalize_loss=True), c["link_norm"]), name
        K = s.size(1)
        nb = 1 if b is None else int(b.max()) + 1
        bp = torch.arange(nb).repeat_interleave(K)
        for so_flag in (False, True):
            for dn in (False, True):
                a_ref, w_ref = c[f"adj_so{int(so_flag)}_dn{int(dn)}"]
                a, w = R.dense_connect_forward_unbatched(ei, ew, b, s, bp, remove_self_loops=True, degree_norm=dn,
                                                         edge_weight_norm=False, sparse_output=so_flag)
                torch.testing.assert_close(a, a_ref, rtol=1e-6, atol=1e-7)
                if w_ref is not None:
                    torch.testing.assert_close(w, w_ref, rtol=1e-6, atol=1e-7)
