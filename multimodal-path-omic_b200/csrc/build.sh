#!/usr/bin/env bash
# Builds libmpo_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
OUT=../libmpo_b200.so
mkdir -p build
objs=()
for f in *.cu; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ] || [ ../../include/mpo_b200.h -nt "$o" ]; then
    echo "[nvcc] $f"
    $NVCC $FLAGS -c "$f" -o "$o" 2> build/${f%.cu}.ptxas.log || { cat build/${f%.cu}.ptxas.log; exit 1; }
  fi
  objs+=("$o")
done
$NVCC -shared -o $OUT "${objs[@]}"
echo "built $(realpath $OUT)"
