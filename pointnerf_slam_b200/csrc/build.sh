#!/bin/bash
# Build libpnslam.so for sm_100a in-tree (the .so travels to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I"$ROOT/include" -I"$HERE" ${PN_NVCC_EXTRA:-})
mkdir -p "$HERE/obj"
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/obj/$(basename "${f%.cu}").o"
  stale=0
  for dep in "$f" "$HERE"/*.cuh "$ROOT/include/pnslam.h"; do
    if [[ ! -f "$o" || "$dep" -nt "$o" ]]; then stale=1; fi
  done
  if [[ $stale == 1 ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$HERE/../libpnslam.so" "$HERE"/obj/*.o
echo "built $HERE/../libpnslam.so"
