#!/usr/bin/env bash
# Ablation sweep of k_dense_bwd_fused (TGPB200_ABL bits: 1 no B split, 2 no proxy fence, 4 no A read, 8 no TMEM store,
# 16 no epilogue, 32 no A loads, 64 no TMA loads at all, 128 no MMAs).  Builds one library per variant into
# benchmarks/build/ (run here), then on the GPU box: python benchmarks/ablate_bwd.py benchmarks/build/lib_abl*.so
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
for v in "$@"; do
  rm -rf /tmp/obj_abl$v; mkdir -p /tmp/obj_abl$v
  cp torch-geometric-pool_b200/csrc/build/*.o /tmp/obj_abl$v/; rm -f /tmp/obj_abl$v/dense_bwd_fused.o
  TGPB200_OUT=$PWD/benchmarks/build/lib_abl$v.so TGPB200_OBJ_DIR=/tmp/obj_abl$v TGPB200_SKIP_OPS=1 \
    TGPB200_EXTRA_FLAGS="-DTGPB200_ABL=$v" bash torch-geometric-pool_b200/csrc/build.sh | tail -1
done
