#!/bin/bash
# side builds of the library with -DLD_EXP=<mask> (timing experiments on the conv epilogue chain): tools/_exp/libld_exp<mask>.so
set -e
cd "$(dirname "$0")/../localdiffusion_hallucination_b200/csrc"
mkdir -p ../../tools/_exp
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
for m in "$@"; do
  nvcc $FLAGS -DLD_EXP=$m -c ld_conv_tc.cu -o ../../tools/_exp/conv_exp$m.o &
done
wait
for m in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_exp/libld_exp$m.so build/ld_kernels_simt.o ../../tools/_exp/conv_exp$m.o build/ld_linattn_tc.o build/ld_attn_tc.o build/ld_conv7_tc.o build/ld_producers.o build/ld_knn_tc.o build/ld_engine.o -cudart static
done
ls -la ../../tools/_exp/*.so
