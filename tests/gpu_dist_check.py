"""Multi-GPU check (run under torchrun, NCCL): the edge-sharded connect with the CUDA operators must equal the
single-GPU result bit-exactly in indices / order and to rtol 1e-5 in weights.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/gpu_dist_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "torch-geometric-pool_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist

import tgp_b200 as T
from tgp_b200 import distributed as D


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(0)
    n, e = 200_000, 2_000_000
    ei = torch.randint(0, n, (2, e), generator=g)
    ei = ei[:, torch.argsort(ei[0] * n + ei[1], stable=True)].to(dev)
    ew = (torch.rand(e, generator=g) + 0.5).to(dev)
    score = torch.randn(n, generator=g).to(dev)
    node_index = torch.sort(torch.topk(score, n // 2).indices)[0]
    batch_pooled = torch.zeros(n // 2, dtype=torch.long, device=dev)
    so = T.SelectOutput(node_index=node_index, num_nodes=n, cluster_index=torch.arange(n // 2, device=dev),
                        num_supernodes=n // 2)
    ref_e, ref_w = T.B200SparseConnect(degree_norm=True, edge_weight_norm=True)(ei, so, edge_weight=ew,
                                                                                batch_pooled=batch_pooled)
    ei_l, ew_l = D.shard_edges(ei, ew, rank, world)
    eo, wo, off, tot = D.sharded_kept_node_connect(ei_l, ew_l, node_index, n, degree_norm=True, edge_weight_norm=True,
                                                   batch_pooled=batch_pooled, num_graphs=1)
    assert tot == ref_e.size(1)
    assert torch.equal(eo, ref_e[:, off:off + eo.size(1)])
    torch.testing.assert_close(wo, ref_w[off:off + eo.size(1)], rtol=1e-5, atol=1e-7)

    K = 60_000
    cluster = torch.randint(0, K, (n,), generator=g).to(dev)
    so_c = T.SelectOutput(cluster_index=cluster, num_supernodes=K)
    ref_e, ref_w = T.B200SparseConnect(degree_norm=True)(ei, so_c, edge_weight=ew)
    ec, wc, (lo, hi) = D.sharded_cluster_connect(ei_l, ew_l, cluster, K, degree_norm=True)
    sel = (ref_e[0] >= lo) & (ref_e[0] < hi)
    assert torch.equal(ec, ref_e[:, sel])
    torch.testing.assert_close(wc, ref_w[sel], rtol=1e-5, atol=1e-7)
    dist.barrier()
    if rank == 0:
        print(f"gpu_dist_check ok: world={world}, kept-node {tot} edges, cluster rows [{lo},{hi}) {ec.size(1)} edges")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
