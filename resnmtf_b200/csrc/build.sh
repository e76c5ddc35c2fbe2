#!/usr/bin/env bash
# Builds the C-ABI shared library in-tree for sm_100a (nvcc cross-compiles without a GPU).  Every *.cu of this
# directory is one translation unit; they are compiled in parallel and linked into one libresnmtf_b200.so.
# RESNMTF_OBJ_CACHE=<dir>: keep the objects there and recompile only the units whose source (or any header of this
# directory) is newer than their object -- for the edit / build loop; the default is a clean build in a temporary directory.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${RESNMTF_OUT:-$here/../libresnmtf_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2)
if [[ "${RESNMTF_VERBOSE_PTXAS:-0}" == "1" ]]; then FLAGS+=(-Xptxas -v); fi
if [[ -n "${RESNMTF_DEFS:-}" ]]; then FLAGS+=(${RESNMTF_DEFS}); fi
if [[ -n "${RESNMTF_OBJ_CACHE:-}" ]]; then
  obj="$RESNMTF_OBJ_CACHE"
  mkdir -p "$obj"
else
  obj="$(mktemp -d "${TMPDIR:-/tmp}/resnmtf_obj.XXXXXX")"
  trap 'rm -rf "$obj"' EXIT
fi
newest_header="$(ls -t "$here"/*.cuh "$here"/*.h "$here"/../../include/*.h | head -1)"
pids=()
objs=()
for src in "$here"/*.cu; do
  o="$obj/$(basename "${src%.cu}").o"
  objs+=("$o")
  if [[ -n "${RESNMTF_OBJ_CACHE:-}" && -f "$o" && "$o" -nt "$src" && "$o" -nt "$newest_header" ]]; then continue; fi
  "$NVCC" "${FLAGS[@]}" -c -o "$o" "$src" &
  pids+=($!)
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
# linked beside the target and moved into place: a reader (a running test, a snapshot of the tree) never sees a partial file
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o "$out.tmp.$$" "${objs[@]}" -lpthread
mv -f "$out.tmp.$$" "$out"
echo "built $out"
