"""Golden vectors for the unbatched dense mode (sparse adjacency + dense [N, K] assignment): produced by the
reference's OWN ``DenseConnect`` and sparse loss functions (tgp/connect/dense_conn.py:141-208,273-354;
tgp/utils/losses.py:126-215,319-389,711-777) through oracle/ref_import.py.  Run in the build container:

    python tests/golden/make_golden_unbatched.py     ->  tests/golden/ref_unbatched.pt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle import ref_path as R  # noqa: E402


def ragged_batch(gen, sizes, p, weighted):
    eis, off = [], 0
    for n in sizes:
        up = torch.triu(torch.rand(n, n, generator=gen) < p, 1)
        r, c = up.nonzero(as_tuple=True)
        eis.append(torch.cat([torch.stack([r, c]), torch.stack([c, r])], 1) + off)
        off += n
    ei = torch.cat(eis, 1)
    ei = ei[:, torch.argsort(ei[0] * off + ei[1])]
    ew = (torch.rand(ei.size(1), generator=gen) + 0.5) if weighted else None
    batch = torch.cat([torch.full((n,), i) for i, n in enumerate(sizes)])
    return ei, ew, batch


def main():
    ref = ref_import.load_reference()
    gen = torch.Generator().manual_seed(0)
    cases = {}
    for name, sizes, K, weighted in (("ragged_w", [12, 20, 7, 15], 5, True), ("ragged_now", [9, 4, 11], 4, False),
                                     ("single_w", [23], 6, True)):
        ei, ew, batch = ragged_batch(gen, sizes, 0.35, weighted)
        N = batch.numel()
        s_raw = torch.randn(N, K, generator=gen)
        b = batch if len(sizes) > 1 else None
        sr = s_raw.clone().requires_grad_(True)
        ww = None if ew is None else ew.clone().requires_grad_(True)
        s = torch.softmax(sr, -1)
        cut = ref.losses.sparse_mincut_loss(ei, s, ww, b, batch_reduction="mean")
        ortho = ref.losses.unbatched_orthogonality_loss(s, b, batch_reduction="mean")
        link = ref.losses.sparse_link_pred_loss(s, ei, ww, b, normalize_loss=False)
        link_n = ref.losses.sparse_link_pred_loss(s, ei, ww, b, normalize_loss=True)
        (cut + 0.5 * ortho + 0.25 * link).backward()
        case = dict(edge_index=ei, edge_weight=ew, batch=b, s_raw=s_raw, cut=cut.detach(), ortho=ortho.detach(),
                    link=link.detach(), link_norm=link_n.detach(), grad_s_raw=sr.grad.clone(),
                    grad_w=None if ww is None else ww.grad.clone())
        bp = torch.arange(len(sizes)).repeat_interleave(K)
        for sparse_output in (False, True):
            for dn in (False, True):
                conn = ref.dense_conn.DenseConnect(remove_self_loops=True, degree_norm=dn, adj_transpose=False,
                                                   edge_weight_norm=False, sparse_output=sparse_output)
                so = ref.base_select.SelectOutput(s=torch.softmax(s_raw, -1), batch=b)
                a, w = conn(ei, so, edge_weight=ew, batch=b, batch_pooled=bp)
                case[f"adj_so{int(sparse_output)}_dn{int(dn)}"] = (a.detach(), None if w is None else w.detach())
        # the oracle restatement must reproduce the reference bit for bit on the same inputs
        s0 = torch.softmax(s_raw, -1)
        assert torch.equal(R.sparse_mincut_loss(ei, s0, ew, b), cut.detach())
        assert torch.equal(R.unbatched_orthogonality_loss(s0, b), ortho.detach())
        assert torch.equal(R.sparse_link_pred_loss(s0, ei, ew, b, normalize_loss=False), link.detach())
        cases[name] = case
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_unbatched.pt")
    torch.save(cases, out)
    print(f"wrote {len(cases)} cases -> {out} ({os.path.getsize(out) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
