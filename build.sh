#!/bin/bash
# Build libsat_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).  Translation units compile in parallel.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/show-attend-and-tell-pytorch-lightning_b200/csrc"
OUT="$HERE/show-attend-and-tell-pytorch-lightning_b200/libsat_b200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v"
mkdir -p "$HERE/build"
pids=()
objs=()
for f in "$SRC"/*.cu; do
  o="$HERE/build/$(basename "${f%.cu}").o"
  objs+=("$o")
  if [ ! -f "$o" ] || [ -n "$(find "$SRC" "$HERE/include" -newer "$o" -type f | head -1)" ]; then
    echo "nvcc $f"
    ( $NVCC $FLAGS -c "$f" -o "$o.tmp" 2> "$o.log" && mv "$o.tmp" "$o" ) &
    pids+=("$!:$o")
  fi
done
fail=0
for e in "${pids[@]}"; do
  pid="${e%%:*}"; o="${e#*:}"
  if ! wait "$pid"; then
    echo "FAILED: $o"; grep -v "deprecated-gpu-targets" "$o.log" | grep -B2 -A6 "error" | head -60
    rm -f "$o" "$o.tmp"
    fail=1
  fi
done
[ "$fail" = 0 ] || exit 1
$NVCC -shared -o "$OUT" "${objs[@]}"
echo "built $OUT"
