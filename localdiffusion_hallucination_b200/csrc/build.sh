#!/bin/bash
# Build libld_sampler.so in-tree for sm_100a.  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=../libld_sampler.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
mkdir -p build
pids=()
for f in ld_kernels_simt.cu ld_conv_tc.cu ld_linattn_tc.cu ld_attn_tc.cu ld_conv7_tc.cu ld_producers.cu ld_knn_tc.cu ld_engine.cu; do
  o=build/${f%.cu}.o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$o")" ] || [ ../../include/ld_sampler.h -nt "$o" ]; then
    $NVCC $FLAGS "$@" -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT build/ld_kernels_simt.o build/ld_conv_tc.o build/ld_linattn_tc.o build/ld_attn_tc.o build/ld_conv7_tc.o build/ld_producers.o build/ld_knn_tc.o build/ld_engine.o -cudart static
echo "built $OUT"
