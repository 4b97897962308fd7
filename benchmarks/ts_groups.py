"""Library-variant sweep (compile-time switches such as TGPB200_TS_GROUPS, TGPB200_FWD_*): times the C2 dense step
(tgpb200_dense_pool_fwd + _bwd through ctypes; FWD_ONLY=1: forward only) for every library given on the command line.

    for g in 2 3 4; do TGPB200_OUT=$PWD/gpurun_out/lib_g$g.so TGPB200_OBJ_DIR=/tmp/obj_g$g TGPB200_SKIP_OPS=1 \
        TGPB200_EXTRA_FLAGS=-DTGPB200_TS_GROUPS=$g bash torch-geometric-pool_b200/csrc/build.sh; done
    python benchmarks/ts_groups.py gpurun_out/lib_g2.so gpurun_out/lib_g3.so gpurun_out/lib_g4.so
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-geometric-pool_b200"))
import torch  # noqa: E402

from tgp_b200 import _lib  # noqa: E402


def bind(path):
    lib = ctypes.CDLL(path)
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def main():
    B, N, K, F = 512, 256, 64, 128
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    a = (torch.rand(B, N, N, device=dev, generator=g) < 0.05).float()
    a = torch.triu(a, 1)
    a = (a + a.transpose(1, 2)).contiguous()
    s = torch.softmax(torch.randn(B, N, K, device=dev, generator=g), -1)
    x = torch.randn(B, N, F, device=dev, generator=g)
    gxp, gap = torch.ones(B, K, F, device=dev), torch.ones(B, K, K, device=dev)
    gl = torch.tensor([1.0, 1.0, 0.0, 0.0], device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ref = None
    for path in sys.argv[1:]:
        lib = bind(path)
        saved = torch.empty(lib.tgpb200_dense_pool_saved_bytes(B, N, K), dtype=torch.uint8, device=dev)
        ws = torch.empty(lib.tgpb200_dense_pool_bwd_workspace_bytes(B, N, K, 0), dtype=torch.uint8, device=dev)
        xp, ap = torch.empty(B, K, F, device=dev), torch.empty(B, K, K, device=dev)
        losses = torch.zeros(4, device=dev)
        gs, gx = torch.empty_like(s), torch.empty_like(x)

        fwd_only = os.environ.get("FWD_ONLY") == "1"

        def step():
            rc = lib.tgpb200_dense_pool_fwd(a.data_ptr(), s.data_ptr(), x.data_ptr(), B, N, K, F, 0, 7, 1, 1e-8, 1.0, 1.0,
                                            xp.data_ptr(), ap.data_ptr(), losses.data_ptr(), saved.data_ptr(),
                                            saved.numel(), st)
            assert rc == 0, rc
            if fwd_only:
                return
            rc = lib.tgpb200_dense_pool_bwd(a.data_ptr(), s.data_ptr(), x.data_ptr(), gxp.data_ptr(), gap.data_ptr(),
                                            gl.data_ptr(), B, N, K, F, 0, 7, 1, 1e-8, 1.0, 1.0, gs.data_ptr(),
                                            gx.data_ptr(), None, saved.data_ptr(), saved.numel(), ws.data_ptr(),
                                            ws.numel(), st)
            assert rc == 0, rc

        for _ in range(10):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            step()
        e1.record()
        torch.cuda.synchronize()
        out = (ap.clone(), xp.clone(), losses.clone()) if fwd_only else (gs.clone(), gx.clone(), ap.clone())
        ok = True if ref is None else all(torch.equal(p, q) for p, q in zip(out, ref))
        ref = ref or out
        print(f"{os.path.basename(path)}: {e0.elapsed_time(e1) / 200:.4f} ms/step ({"fwd" if fwd_only else "fwd+bwd"}, C2, eager C ABI)  same_as_first={ok}")


if __name__ == "__main__":
    main()
