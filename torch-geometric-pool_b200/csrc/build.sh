#!/usr/bin/env bash
# Build libtgp_b200.so (sm_100a only) in-tree next to the python package.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${TGPB200_OUT:-${HERE}/../tgp_b200/libtgp_b200.so}"  # TGPB200_OUT / TGPB200_OBJ_DIR / TGPB200_EXTRA_FLAGS: tuning variants
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC ${TGPB200_EXTRA_FLAGS:-})
OBJ="${TGPB200_OBJ_DIR:-${HERE}/build}"
mkdir -p "${OBJ}"
pids=()
for src in "${HERE}"/*.cu; do
  o="${OBJ}/$(basename "${src}" .cu).o"
  if [[ ! -f "${o}" || "${src}" -nt "${o}" || -n "$(find "${HERE}" "${HERE}/../../include" -name '*.h' -newer "${o}" -o -name '*.cuh' -newer "${o}" | head -1)" ]]; then
    "${NVCC}" "${FLAGS[@]}" -c "${src}" -o "${o}" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "${p}" ]] && wait "${p}"; done

"${NVCC}" -shared -gencode arch=compute_100a,code=sm_100a -o "${OUT}" "${OBJ}"/*.o
echo "built ${OUT}"
