"""Times tgpb200_dense_pool_fwd (C2 shape: fused forward + A_raw + per-graph epilogue + finalize) as a CUDA-graph replay
for every library variant given on the command line (variants: -DTGPB200_ABLF=<bits>, see dense_fused_ts.cu)."""
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
for path in sys.argv[1:]:
    lib = ctypes.CDLL(path)
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name); fn.restype, fn.argtypes = res, args
    saved = torch.empty(lib.tgpb200_dense_pool_saved_bytes(B, N, K), dtype=torch.uint8, device=dev)
    xp, ap = torch.empty(B, K, F, device=dev), torch.empty(B, K, K, device=dev)
    losses = torch.zeros(4, device=dev)

    def fwd(stream):
        assert lib.tgpb200_dense_pool_fwd(a.data_ptr(), s.data_ptr(), x.data_ptr(), B, N, K, F, 0, 7, 1, 1e-8, 1.0, 1.0,
                                          xp.data_ptr(), ap.data_ptr(), losses.data_ptr(), saved.data_ptr(), saved.numel(),
                                          stream) == 0
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        fwd(st)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        cs = torch.cuda.current_stream().cuda_stream
        for _ in range(10):
            fwd(cs)
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{os.path.basename(path):18s} forward chain {e0.elapsed_time(e1) / 100 * 1000:7.1f} us (graph replay)", flush=True)
