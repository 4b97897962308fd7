#!/usr/bin/env bash
# Build libtgp_b200.so (sm_100a only) in-tree next to the python package.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../tgp_b200/libtgp_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC)
OBJ="${HERE}/build"
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
