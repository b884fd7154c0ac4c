#!/bin/bash
# timing sweep of the LD_EXP side builds (tools/build_exp.sh) x LD_CONV_DBG knock-outs on the dominant conv launch
for d in 0 1 2 4 3 5 6 7; do
  LD_CONV_DBG=$d timeout 120 python tools/gpu_conv_one.py 32 0 256 32 3 0 32 2>&1 | tail -1
done
for m in 1 2 3 4 8 16 31; do
  for d in 0 7; do
    LD_SAMPLER_LIB=tools/_exp/libld_exp$m.so LD_CONV_DBG=$d timeout 120 python tools/gpu_conv_one.py 32 0 256 32 3 0 32 2>&1 | tail -1
  done
done
