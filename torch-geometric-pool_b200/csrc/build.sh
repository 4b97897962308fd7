#!/usr/bin/env bash
# Build libtgp_b200.so (sm_100a kernels + C ABI) and libtgp_b200_ops.so (torch custom-op shim) in-tree next to the
# python package.
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

# torch custom-op shim: host-only C++ (no device code), linked against libtgp_b200.so through $ORIGIN
OPS_OUT="$(dirname "${OUT}")/libtgp_b200_ops.so"
OPS_SRC="${HERE}/torch_ops.cpp"
if [[ -z "${TGPB200_SKIP_OPS:-}" && ( ! -f "${OPS_OUT}" || "${OPS_SRC}" -nt "${OPS_OUT}" || "${HERE}/../../include/tgp_b200.h" -nt "${OPS_OUT}" ) ]]; then
  PY="${PYTHON:-python}"
  readarray -t TF < <("${PY}" - <<'PYEOF'
import os, torch
from torch.utils import cpp_extension as ce
print(" ".join("-I" + p for p in ce.include_paths()))
print(os.path.join(os.path.dirname(torch.__file__), "lib"))
print(int(torch._C._GLIBCXX_USE_CXX11_ABI))
PYEOF
  )
  g++ -O2 -std=c++17 -fPIC -shared -D_GLIBCXX_USE_CXX11_ABI="${TF[2]}" ${TF[0]} -I/usr/local/cuda/include \
    "${OPS_SRC}" -o "${OPS_OUT}" -L"$(dirname "${OUT}")" -ltgp_b200 -L"${TF[1]}" -ltorch -ltorch_cpu -lc10 -lc10_cuda \
    -ltorch_cuda -Wl,-rpath,'$ORIGIN' -Wl,-rpath,"${TF[1]}"
  echo "built ${OPS_OUT}"
fi
