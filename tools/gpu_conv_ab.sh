#!/bin/bash
# A/B of the conv pipeline depth knobs on the four in-situ variants of the dominant convolution (and two neighbours)
for cfg in "4 2" "5 2" "6 2" "4 4" "6 4"; do
  set -- $cfg
  echo "== LD_CONV_SA=$1 LD_CONV_NACC=$2"
  for v in 0 1 2 3; do LD_CONV_SA=$1 LD_CONV_NACC=$2 timeout 120 python tools/gpu_conv_variant_one.py $v 2>&1 | tail -1; done
  LD_CONV_SA=$1 LD_CONV_NACC=$2 timeout 120 python tools/gpu_conv_one.py 64 0 128 64 3 0 32 2>&1 | tail -1
  LD_CONV_SA=$1 LD_CONV_NACC=$2 timeout 120 python tools/gpu_conv_one.py 32 0 256 32 1 0 32 2>&1 | tail -1
done
