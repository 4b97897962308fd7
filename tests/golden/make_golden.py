"""Generate ``tests/golden/ref_vectors.pt`` by running the reference's OWN files.

Run in the build container only (``/root/reference`` must exist):

    python tests/golden/make_golden.py

The reference's ``BaseReduce``, ``SparseConnect``, ``DenseConnect``, ``SelectOutput``,
``TopkSelect``, ``postprocess_adj_pool_*`` and the four dense losses are imported from
/root/reference through ``oracle/ref_import.py`` (stub ``torch_geometric`` /
``torch_scatter`` exposing the restated primitives only) and executed on seeded
inputs; inputs, outputs and autograd gradients are stored.  The fixtures are what
``tests/test_oracle_golden.py`` (CPU) and the ``-m gpu`` parity tests compare against.
"""

from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402


def er_batch(gen, num_graphs, n_lo, n_hi, p, feats, weighted):
    xs, eis, batch = [], [], []
    off = 0
    for g in range(num_graphs):
        n = int(torch.randint(n_lo, n_hi + 1, (1,), generator=gen))
        upper = torch.triu(torch.rand(n, n, generator=gen) < p, diagonal=1)
        r, c = upper.nonzero(as_tuple=True)
        ei = torch.cat([torch.stack([r, c]), torch.stack([c, r])], 1) + off
        eis.append(ei)
        xs.append(torch.randn(n, feats, generator=gen))
        batch.append(torch.full((n,), g, dtype=torch.long))
        off += n
    ei = torch.cat(eis, 1)
    # PyG datasets store edges sorted by (row, col)
    order = torch.argsort(ei[0] * off + ei[1], stable=True)
    ei = ei[:, order]
    ew = torch.rand(ei.size(1), generator=gen) + 0.5 if weighted else None
    return torch.cat(xs), ei, ew, torch.cat(batch)


def main() -> None:
    ref = ref_import.load_reference()
    gen = torch.Generator().manual_seed(0)
    cases = {}

    # ---- kept-node path (TopK): reduce + connect, all post-processing flag combos
    x, ei, ew, batch = er_batch(gen, 6, 8, 14, 0.3, 8, weighted=True)
    sel = ref.topk_select.TopkSelect(in_channels=8, ratio=0.5)
    with torch.no_grad():
        sel.weight.copy_(torch.randn(1, 8, generator=gen))
    for tag, w in (("w", ew), ("now", None)):
        for dn in (False, True):
            for ewn in (False, True):
                for rsl in (False, True):
                    xx = x.clone().requires_grad_(True)
                    ww = None if w is None else w.clone().requires_grad_(True)
                    so = sel(x=xx, batch=batch)
                    xp, bp = ref.base_reduce.BaseReduce()(x=xx, so=so, batch=batch)
                    conn = ref.base_conn.SparseConnect(remove_self_loops=rsl, edge_weight_norm=ewn, degree_norm=dn)
                    eo, wo = conn(ei, so, edge_weight=ww, batch_pooled=bp)
                    loss = xp.square().sum()
                    if wo is not None and wo.requires_grad:
                        loss = loss + (wo * torch.arange(1, wo.numel() + 1)).sum()
                    loss.backward()
                    cases[f"topk_{tag}_dn{int(dn)}_ewn{int(ewn)}_rsl{int(rsl)}"] = dict(
                        x=x, edge_index=ei, edge_weight=w, batch=batch, p=sel.weight.detach().clone(),
                        node_index=so.node_index, cluster_index=so.cluster_index, weight=so.weight.detach(),
                        num_nodes=so.num_nodes, num_supernodes=so.num_supernodes,
                        x_pool=xp.detach(), batch_pool=bp, edge_index_out=eo,
                        edge_weight_out=None if wo is None else wo.detach(),
                        grad_x=xx.grad, grad_w=None if ww is None else ww.grad,
                        grad_p=sel.weight.grad.clone(),
                    )
                    sel.weight.grad = None

    # ---- cluster path: random cluster maps with duplicates / self loops / empty clusters
    x, ei, ew, batch = er_batch(gen, 4, 10, 16, 0.35, 6, weighted=True)
    n = x.size(0)
    loops = torch.arange(0, n, 5)
    ei = torch.cat([ei, torch.stack([loops, loops]), ei[:, :7]], 1)  # self loops + exact duplicates
    ew = torch.cat([ew, torch.rand(loops.numel(), generator=gen), torch.rand(7, generator=gen)])
    ew[3] = 0.0  # tiny weight
    k_sup = n // 2 + 3  # a few empty clusters at the end
    cluster = torch.randint(0, n // 2, (n,), generator=gen)
    bp_cluster = torch.zeros(k_sup, dtype=torch.long).scatter_(0, cluster, batch)
    for op in ("sum", "mean", "max", "min", "mul"):
        for tag, w in (("w", ew), ("now", None)):
            for dn, ewn, rsl in ((False, False, True), (True, False, True), (True, True, False), (False, True, True)):
                xx = x.clone().requires_grad_(True)
                ww = None if w is None else w.clone().requires_grad_(True)
                so = ref.SelectOutput(cluster_index=cluster, num_nodes=n, num_supernodes=k_sup)
                xp, bp = ref.base_reduce.BaseReduce()(x=xx, so=so, batch=batch)
                conn = ref.base_conn.SparseConnect(reduce_op=op, remove_self_loops=rsl, edge_weight_norm=ewn, degree_norm=dn)
                eo, wo = conn(ei, so, edge_weight=ww, batch_pooled=bp_cluster)
                loss = xp.square().sum()
                if wo is not None and wo.requires_grad:
                    loss = loss + (wo * torch.arange(1, wo.numel() + 1)).sum()
                loss.backward()
                cases[f"cluster_{op}_{tag}_dn{int(dn)}_ewn{int(ewn)}_rsl{int(rsl)}"] = dict(
                    x=x, edge_index=ei, edge_weight=w, batch=batch, cluster=cluster, num_supernodes=k_sup,
                    batch_pooled=bp_cluster, x_pool=xp.detach(), batch_pool=bp, edge_index_out=eo,
                    edge_weight_out=None if wo is None else wo.detach(),
                    grad_x=xx.grad, grad_w=None if ww is None else ww.grad,
                )

    # ---- dense MinCut / DiffPool pieces
    B, N, K, F = 3, 12, 4, 5
    a = (torch.rand(B, N, N, generator=gen) < 0.3).float()
    a = torch.triu(a, 1)
    a = a + a.transpose(1, 2)
    a[2, 9:, :] = 0  # zero-padded smaller graph
    a[2, :, 9:] = 0
    s_raw = torch.randn(B, N, K, generator=gen)
    mask = torch.ones(B, N, dtype=torch.bool)
    mask[2, 9:] = False
    xd = torch.randn(B, N, F, generator=gen) * mask[..., None]
    for dn in (False, True):
        for adjt in (False, True):
            for ewn in (False, True):
                for rsl in (False, True):
                    sr = s_raw.clone().requires_grad_(True)
                    xx = xd.clone().requires_grad_(True)
                    aa = a.clone().requires_grad_(True)
                    s = torch.softmax(sr, -1) * mask[..., None]
                    so = ref.SelectOutput(s=s)
                    xp, _ = ref.base_reduce.BaseReduce()(x=xx, so=so)
                    conn = ref.dense_conn.DenseConnect(remove_self_loops=rsl, degree_norm=dn, adj_transpose=adjt, edge_weight_norm=ewn)
                    raw = conn.dense_connect(adj=aa, s=s)
                    cut = ref.losses.mincut_loss(aa, s, raw, batch_reduction="mean")
                    ortho = ref.losses.orthogonality_loss(s, batch_reduction="mean")
                    link = ref.losses.link_pred_loss(s, aa, normalize_loss=False)
                    link_n = ref.losses.link_pred_loss(s, aa, normalize_loss=True)
                    ent = ref.losses.entropy_loss(s, int(mask.sum()))
                    post = ref.ops.postprocess_adj_pool_dense(
                        raw.clone(), remove_self_loops=rsl, degree_norm=dn, adj_transpose=adjt, edge_weight_norm=ewn
                    )
                    wts = torch.arange(1, post.numel() + 1, dtype=torch.float).view_as(post) / post.numel()
                    total = xp.square().sum() + (post * wts).sum() + cut + ortho + 0.5 * link + 0.25 * ent
                    total.backward()
                    cases[f"dense_dn{int(dn)}_t{int(adjt)}_ewn{int(ewn)}_rsl{int(rsl)}"] = dict(
                        adj=a, s_raw=s_raw, mask=mask, x=xd, s=s.detach(), x_pool=xp.detach(), adj_pool_raw=raw.detach(),
                        adj_pool=post.detach(), cut=cut.detach(), ortho=ortho.detach(), link=link.detach(),
                        link_norm=link_n.detach(), ent=ent.detach(), grad_s_raw=sr.grad, grad_x=xx.grad, grad_adj=aa.grad,
                    )

    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.pt")
    torch.save(cases, out)
    print(f"wrote {len(cases)} cases -> {out} ({os.path.getsize(out)/1024:.0f} KiB)")


if __name__ == "__main__":
    main()
