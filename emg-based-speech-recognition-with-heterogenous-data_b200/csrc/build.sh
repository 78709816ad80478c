#!/bin/bash
# Builds libsst.so (all CUDA kernels + the C ABI) for sm_100a, in-tree.
set -e
cd "$(dirname "$0")"
OUT=../libsst.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v"
mkdir -p build
objs=""
pids=""
for f in *.cu; do
  o=build/${f%.cu}.o
  objs="$objs $o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer "$o")" ] || [ ../../include/sst.h -nt "$o" ]; then
    ( $NVCC $FLAGS -c "$f" -o "$o" > build/${f%.cu}.log 2>&1 || { cat build/${f%.cu}.log; exit 1; } ) &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $objs -lcudart -ldl
echo "built $(realpath $OUT)"
