import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200"))
import torch, tgp_b200 as T
from oracle import ref_path as R
g = torch.Generator().manual_seed(3*256+64)
B,N,K,F=3,256,64,128
a = (torch.rand(B, N, N, generator=g) < 0.1).float(); a = torch.triu(a, 1); a = a + a.transpose(1, 2)
s = torch.softmax(torch.randn(B, N, K, generator=g), -1); x = torch.randn(B, N, F, generator=g)
e = R.diff_pool(x, a, s)[2]
o = T.diff_pool(x.cuda(), a.cuda(), s.cuda())[2]
print({k: float(v) for k, v in e.items()}, {k: float(v) for k, v in o.items()})
ss = s @ s.transpose(1,2)
print("a2", float((a*a).sum()), "num", float(torch.einsum('bii->', R.dense_connect(a,s))), "m2", float(((s.transpose(1,2)@s)**2).sum()), "ss2", float((ss*ss).sum()), "q", float(((a-ss)**2).sum()), float(((a.double()-ss.double())**2).sum()))
