#!/usr/bin/env bash
# Builds the C-ABI shared library in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${RESNMTF_OUT:-$here/../libresnmtf_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2 -shared)
if [[ "${RESNMTF_VERBOSE_PTXAS:-0}" == "1" ]]; then FLAGS+=(-Xptxas -v); fi
if [[ -n "${RESNMTF_DEFS:-}" ]]; then FLAGS+=(${RESNMTF_DEFS}); fi
if [[ "${RESNMTF_WITH_NCCL:-0}" == "1" && -f /usr/include/nccl.h ]]; then
  FLAGS+=(-DRESNMTF_WITH_NCCL -lnccl)
fi
"$NVCC" "${FLAGS[@]}" -o "$out" "$here/resnmtf_capi.cu"
echo "built $out"
