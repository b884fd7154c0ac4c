#!/bin/bash
# timing sweep of LD_EXP side builds on the normalise-on-load + statistics variant of the 32-channel conv (and the statistics variant)
for v in 2 1; do python tools/gpu_conv_variant_one.py $v; done
for m in "$@"; do
  echo "== LD_EXP=$m"
  for v in 2 1; do LD_SAMPLER_LIB=tools/_exp/libld_exp$m.so timeout 120 python tools/gpu_conv_variant_one.py $v; done
done
for d in 1 2 4; do echo "== LD_CONV_DBG=$d"; LD_CONV_DBG=$d python tools/gpu_conv_variant_one.py 2; done
