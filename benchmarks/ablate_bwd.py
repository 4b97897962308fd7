"""Times tgpb200_dense_pool_bwd (C2 shape: k_graph_bwd + k_dense_bwd_fused) for every library variant given on the
command line (benchmarks/ablate_bwd.sh builds them).  Results of ablated variants are wrong by construction."""
import ctypes
import os
import sys

sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200"))
import torch  # noqa: E402
from tgp_b200 import _lib  # noqa: E402

B, N, K, F = 512, 256, 64, 128
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
a = (torch.rand(B, N, N, device=dev, generator=g) < 0.05).float()
a = torch.triu(a, 1); a = (a + a.transpose(1, 2)).contiguous()
s = torch.softmax(torch.randn(B, N, K, device=dev, generator=g), -1)
x = torch.randn(B, N, F, device=dev, generator=g)
gxp, gap = torch.ones(B, K, F, device=dev), torch.ones(B, K, K, device=dev)
gl = torch.tensor([1.0, 1.0, 0.0, 0.0], device=dev)
st = torch.cuda.current_stream().cuda_stream
for path in sys.argv[1:]:
    lib = ctypes.CDLL(path)
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name); fn.restype, fn.argtypes = res, args
    saved = torch.empty(lib.tgpb200_dense_pool_saved_bytes(B, N, K), dtype=torch.uint8, device=dev)
    ws = torch.empty(lib.tgpb200_dense_pool_bwd_workspace_bytes(B, N, K, 0), dtype=torch.uint8, device=dev)
    xp, ap = torch.empty(B, K, F, device=dev), torch.empty(B, K, K, device=dev)
    losses = torch.zeros(4, device=dev)
    gs, gx = torch.empty_like(s), torch.empty_like(x)
    assert lib.tgpb200_dense_pool_fwd(a.data_ptr(), s.data_ptr(), x.data_ptr(), B, N, K, F, 0, 7, 1, 1e-8, 1.0, 1.0,
                                      xp.data_ptr(), ap.data_ptr(), losses.data_ptr(), saved.data_ptr(), saved.numel(), st) == 0

    def bwd(stream):
        assert lib.tgpb200_dense_pool_bwd(a.data_ptr(), s.data_ptr(), x.data_ptr(), gxp.data_ptr(), gap.data_ptr(), gl.data_ptr(),
                                          B, N, K, F, 0, 7, 1, 1e-8, 1.0, 1.0, gs.data_ptr(), gx.data_ptr(), None,
                                          saved.data_ptr(), saved.numel(), ws.data_ptr(), ws.numel(), stream) == 0
    for _ in range(5):
        bwd(st)
    torch.cuda.synchronize()
    # one CUDA graph of 10 backward calls: the eager call is CPU-bound (~60 us of tensor-map encodes per call)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cs = torch.cuda.current_stream().cuda_stream
        for _ in range(10):
            bwd(cs)
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{os.path.basename(path):18s} backward {e0.elapsed_time(e1) / 100 * 1000:7.1f} us (graph replay)", flush=True)
