#!/bin/bash
# LD_CONV_DBG sweep (1 no activation loads, 2 no MMAs, 4 no output stores) for a side-built library (default: tools/_mx)
lib=${1:-tools/_mx/libld_sampler_mx.so}
for shape in "32 0 256 32 3 0 32" "64 0 256 32 3 0 32"; do
  for d in 0 1 2 4 7; do
    LD_SAMPLER_LIB=$lib LD_CONV_DBG=$d python tools/gpu_conv_one.py $shape 2>&1 | tail -1
  done
done
