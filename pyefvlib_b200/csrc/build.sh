#!/usr/bin/env bash
# Build libefv_b200.so for sm_100a (B200) in-tree.  Usage: build.sh [extra nvcc flags]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libefv_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall)
mkdir -p "$HERE/_obj"
pids=()
for f in runtime mesh geometry csr assemble krylov amg comm api; do
  "$NVCC" "${FLAGS[@]}" "$@" -c "$HERE/$f.cu" -o "$HERE/_obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$HERE"/_obj/{runtime,mesh,geometry,csr,assemble,krylov,amg,comm,api}.o -lcudart -ldl
echo "built $OUT"
