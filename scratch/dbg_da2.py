import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200")); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, tgp_b200 as T
from tgp_b200 import _lib as L
from oracle import ref_path as R
golden = torch.load("tests/golden/ref_vectors.pt", weights_only=False)
c = golden["dense_dn0_t0_ewn1_rsl0"]
s = (torch.softmax(c["s_raw"], -1) * c["mask"][..., None])
a = c["adj"]
raw = R.dense_connect(a, s).detach().requires_grad_(True)
post = R.postprocess_adj_pool_dense(raw, edge_weight_norm=True)
(post * torch.arange(post.numel()).view_as(post) / post.numel()).sum().backward()
G = raw.grad  # [3,4,4]
print("G", G[0])
B, N, K = s.shape
sg, Gg = s.cuda().contiguous(), G.cuda().contiguous()
U = torch.zeros(B, N, K, device="cuda")
L.call("tgpb200_tc_gemm", L.ptr(sg), L.ptr(Gg), L.ptr(U), B, N, K, K, N*K, K, 0, K*K, K, 1, N*K, K, 1, 0, 0, 1.0, 0, L.stream())
dA = torch.zeros(B, N, N, device="cuda")
L.call("tgpb200_tc_gemm", L.ptr(U), L.ptr(sg), L.ptr(dA), B, N, N, K, N*K, K, 0, N*K, K, 0, N*N, N, 1, 0, 0, 1.0, 0, L.stream())
torch.cuda.synchronize()
Ur = s @ G; dAr = Ur @ s.transpose(1, 2)
print("U maxdiff", (U.cpu() - Ur).abs().max().item(), "dA maxdiff", (dA.cpu() - dAr).abs().max().item())
print((dA.cpu() - dAr)[0])
