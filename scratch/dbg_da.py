import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200"))
import torch, tgp_b200 as T
from oracle import ref_path as R
golden = torch.load("tests/golden/ref_vectors.pt", weights_only=False)
c = golden["dense_dn0_t0_ewn1_rsl0"]
kw = dict(remove_self_loops=False, degree_norm=False, adj_transpose=False, edge_weight_norm=True)
def run(mod, dev, which):
    sr = c["s_raw"].clone().to(dev).requires_grad_(True); x = c["x"].clone().to(dev).requires_grad_(True); a = c["adj"].clone().to(dev).requires_grad_(True)
    s = torch.softmax(sr, -1) * c["mask"].to(dev)[..., None]
    if which == "mincut_post":
        xp, post, l = mod.mincut_pool(x, a, s, **kw); tot = (post * torch.arange(post.numel(), device=dev).view_as(post) / post.numel()).sum()
    elif which == "mincut_loss":
        xp, post, l = mod.mincut_pool(x, a, s, **kw); tot = l["cut_loss"] + l["ortho_loss"]
    elif which == "diff_loss":
        xp, post, l = mod.diff_pool(x, a, s, num_nodes=int(c["mask"].sum()), **kw); tot = 0.5 * l["link_loss"] + 0.25 * l["entropy_loss"]
    tot.backward()
    return a.grad.cpu(), sr.grad.cpu()
for which in ("mincut_post", "mincut_loss", "diff_loss"):
    ea, es = run(R, "cpu", which); ga, gs = run(T, "cuda", which)
    print(which, "dA maxdiff", (ea - ga).abs().max().item(), "dS maxdiff", (es - gs).abs().max().item(), "|dA|max", ea.abs().max().item())
