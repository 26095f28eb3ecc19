#!/usr/bin/env bash
# Builds libmpo_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).  Stale objects are rebuilt in parallel.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
OUT=../libmpo_b200.so
mkdir -p build
objs=()
pids=()
names=()
for f in *.cu; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ] || [ ../../include/mpo_b200.h -nt "$o" ]; then
    echo "[nvcc] $f"
    ( $NVCC $FLAGS -c "$f" -o "$o.tmp" 2> build/${f%.cu}.ptxas.log && mv "$o.tmp" "$o" ) &
    pids+=($!)
    names+=("$f")
  fi
  objs+=("$o")
done
fail=0
for i in "${!pids[@]}"; do
  if ! wait "${pids[$i]}"; then
    echo "nvcc failed on ${names[$i]}:"; cat "build/${names[$i]%.cu}.ptxas.log"; fail=1
  fi
done
[ "$fail" = 0 ] || exit 1
$NVCC -shared -o $OUT "${objs[@]}"
echo "built $(realpath $OUT)"
