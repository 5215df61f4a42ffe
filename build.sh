#!/bin/bash
# Build libsat_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/show-attend-and-tell-pytorch-lightning_b200/csrc"
OUT="$HERE/show-attend-and-tell-pytorch-lightning_b200/libsat_b200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v"
mkdir -p "$HERE/build"
for f in "$SRC"/*.cu; do
  o="$HERE/build/$(basename "${f%.cu}").o"
  if [ ! -f "$o" ] || [ -n "$(find "$SRC" "$HERE/include" -newer "$o" -type f | head -1)" ]; then
    echo "nvcc $f"
    $NVCC $FLAGS -c "$f" -o "$o" 2> "$o.log" || { cat "$o.log"; exit 1; }
  fi
done
$NVCC -shared -o "$OUT" "$HERE"/build/*.o
echo "built $OUT"
