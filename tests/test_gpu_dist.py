"""Multi-GPU (needs >= 2 GPUs, NCCL): the edge-sharded connect and the node-sharded reduce with the CUDA operators as
the local kernels must equal the single-GPU result -- bit-exact indices and edge order, rtol 1e-5 weights and
gradients.  Self-spawning (one process per GPU); skipped on a single-GPU box.

    gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu -q
"""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fail):
    import torch.distributed as dist

    import tgp_b200 as T
    from tgp_b200 import distributed as D

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        g = torch.Generator().manual_seed(0)
        n, e = 200_000, 2_000_000
        ei = torch.randint(0, n, (2, e), generator=g)
        ei = ei[:, torch.argsort(ei[0] * n + ei[1], stable=True)].to(dev)
        ew = (torch.rand(e, generator=g) + 0.5).to(dev)
        score = torch.randn(n, generator=g).to(dev)
        node_index = torch.sort(torch.topk(score, n // 2).indices)[0]
        K = n // 2
        batch_pooled = torch.zeros(K, dtype=torch.long, device=dev)
        so = T.SelectOutput(node_index=node_index, num_nodes=n, cluster_index=torch.arange(K, device=dev),
                            num_supernodes=K)
        # ---- kept-node connect, both normalisations, forward
        ref_e, ref_w = T.B200SparseConnect(degree_norm=True, edge_weight_norm=True)(ei, so, edge_weight=ew,
                                                                                    batch_pooled=batch_pooled)
        ei_l, ew_l = D.shard_edges(ei, ew, rank, world)
        eo, wo, off, tot = D.sharded_kept_node_connect(ei_l, ew_l, node_index, n, degree_norm=True,
                                                       edge_weight_norm=True, batch_pooled=batch_pooled, num_graphs=1,
                                                       rows_sorted=True)
        assert tot == ref_e.size(1)
        assert torch.equal(eo, ref_e[:, off:off + eo.size(1)])
        torch.testing.assert_close(wo, ref_w[off:off + eo.size(1)], rtol=1e-5, atol=1e-7)
        # ---- ... and the gradient through the sharded degree normalisation (all-reduce in the backward)
        w_full = ew.clone().requires_grad_(True)
        _, rw = T.B200SparseConnect(degree_norm=True)(ei, so, edge_weight=w_full)
        coef = torch.rand(rw.numel(), generator=torch.Generator(device=dev).manual_seed(1), device=dev)
        (rw * coef).sum().backward()
        w_loc = ew_l.clone().requires_grad_(True)
        eo, wo, off, tot = D.sharded_kept_node_connect(ei_l, w_loc, node_index, n, degree_norm=True, rows_sorted=True)
        (wo * coef[off:off + wo.numel()]).sum().backward()
        lo, hi = D.even_ranges(e, world)[rank]
        torch.testing.assert_close(w_loc.grad, w_full.grad[lo:hi], rtol=1e-4, atol=1e-5 * float(w_full.grad.abs().max()))
        # ---- cluster connect: sum and mean, degree normalisation
        Kc = 60_000
        cluster = torch.randint(0, Kc, (n,), generator=g).to(dev)
        so_c = T.SelectOutput(cluster_index=cluster, num_supernodes=Kc)
        for op in ("sum", "mean"):
            ref_e, ref_w = T.B200SparseConnect(op, degree_norm=True)(ei, so_c, edge_weight=ew)
            ec, wc, (r_lo, r_hi) = D.sharded_cluster_connect(ei_l, ew_l, cluster, Kc, reduce_op=op, degree_norm=True)
            sel = (ref_e[0] >= r_lo) & (ref_e[0] < r_hi)
            assert torch.equal(ec, ref_e[:, sel]), op
            torch.testing.assert_close(wc, ref_w[sel], rtol=1e-5, atol=1e-7)
        # ---- node-sharded feature reduce: reduce_scatter of the [K, F] partials
        x = torch.randn(n, 32, generator=g).to(dev)
        n_lo, n_hi = D.even_ranges(n, world)[rank]
        for op in ("sum", "mean"):
            ref, _ = T.B200Reduce(op)(x, so_c)
            got, (r_lo, r_hi) = D.sharded_cluster_reduce(x[n_lo:n_hi].contiguous(), cluster[n_lo:n_hi].contiguous(), Kc,
                                                         reduce_op=op)
            torch.testing.assert_close(got, ref[r_lo:r_hi], rtol=1e-5, atol=1e-5)
        dist.barrier()
    except Exception:  # noqa: BLE001
        import traceback

        traceback.print_exc()
        fail.value = 1
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_connect_and_reduce_on_gpus(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    fail = ctx.Value("i", 0)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fail)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(900)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    assert fail.value == 0
