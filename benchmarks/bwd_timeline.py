"""clock64 timeline of block 0 of the fused dense backward (k_dense_bwd_fused, C2 shape) + its launch time.
Columns per k-block: split start (stage landed), A read done, ring wait done, arrive | mma ready, mma issued."""
import ctypes
import os
import sys

sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "torch-geometric-pool_b200"))
import torch  # noqa: E402
from tgp_b200 import _lib as L  # noqa: E402

lib = L.load()
lib.tgpb200_debug_engine_timeline.argtypes = [ctypes.c_void_p]
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
saved = torch.empty(lib.tgpb200_dense_pool_saved_bytes(B, N, K), dtype=torch.uint8, device=dev)
ws = torch.empty(lib.tgpb200_dense_pool_bwd_workspace_bytes(B, N, K, 0), dtype=torch.uint8, device=dev)
xp, ap = torch.empty(B, K, F, device=dev), torch.empty(B, K, K, device=dev)
losses = torch.zeros(4, device=dev)
gs, gx = torch.empty_like(s), torch.empty_like(x)


def fwd():
    assert lib.tgpb200_dense_pool_fwd(a.data_ptr(), s.data_ptr(), x.data_ptr(), B, N, K, F, 0, 7, 1, 1e-8, 1.0, 1.0,
                                      xp.data_ptr(), ap.data_ptr(), losses.data_ptr(), saved.data_ptr(), saved.numel(), st) == 0


def bwd():
    assert lib.tgpb200_dense_pool_bwd(a.data_ptr(), s.data_ptr(), x.data_ptr(), gxp.data_ptr(), gap.data_ptr(), gl.data_ptr(),
                                      B, N, K, F, 0, 7, 1, 1e-8, 1.0, 1.0, gs.data_ptr(), gx.data_ptr(), None,
                                      saved.data_ptr(), saved.numel(), ws.data_ptr(), ws.numel(), st) == 0


fwd()
for _ in range(5):
    bwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    bwd()
e1.record(); torch.cuda.synchronize()
print(f"backward (k_graph_bwd + fused or 3 products): {e0.elapsed_time(e1) / 50 * 1000:.1f} us")
dbg = torch.zeros(360 * 8, dtype=torch.long, device=dev)
lib.tgpb200_debug_engine_timeline(dbg.data_ptr())
bwd(); torch.cuda.synchronize()
lib.tgpb200_debug_engine_timeline(None)
d = dbg.cpu().view(360, 8)
t0 = int(d[0, 1])
print(" kc | arrive_q0 arrive_q1 arrive_q2 arrive_q3 | wait_start mma_ready mmas_issued committed   (22 k-blocks per item: 8 W, 4 X, 2 T, 2 SP, 2 WG, 4 dX)")
for i in range(0, 70):
    r = [int(v) - t0 for v in d[i, :8]]
    print(f"{i:3d} | {r[1]:9d} {r[2]:9d} {r[3]:9d} {r[4]:9d} | {r[7]:9d} {r[5]:9d} {r[0]:9d} {r[6]:9d}")
print(" item  epi_start epi_end")
for i in range(0, 8):
    r = [int(v) - t0 for v in d[128 + i, :2]]
    print(f"{i:3d} {r[0]:9d} {r[1]:9d}")

# per-CTA start / end (global timer, ns)
import statistics
rows = [(int(d[200 + i, 0]), int(d[200 + i, 1]), int(d[200 + i, 2])) for i in range(148) if int(d[200 + i, 0]) > 0]
g0 = min(r[0] for r in rows)
starts = sorted(r[0] - g0 for r in rows); ends = sorted(r[1] - g0 for r in rows); durs = sorted(r[1] - r[0] for r in rows)
print(f"CTAs: {len(rows)}  start ns min/med/max {starts[0]}/{statistics.median(starts)}/{starts[-1]}  end ns min/med/max {ends[0]}/{statistics.median(ends)}/{ends[-1]}")
print(f"duration ns min/med/max {durs[0]}/{statistics.median(durs)}/{durs[-1]}")
slow = sorted(rows, key=lambda r: r[0] - r[1])[:12]
print("slowest CTAs (block, smid, dur ns):", [(rows.index(r), r[2], r[1] - r[0]) for r in slow])
fast = sorted(rows, key=lambda r: r[1] - r[0])[:12]
print("fastest CTAs (block, smid, dur ns):", [(rows.index(r), r[2], r[1] - r[0]) for r in fast])
