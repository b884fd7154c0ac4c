#!/bin/bash
# knock-outs of the 256 -> 256 convolution at 32 x 32 (N images): LD_CONV_DBG 1 no activation loads, 2 no MMAs, 4 no stores; LD_CONV_MT=1 no tile pairs
N=${1:-32}
for e in "LD_X=1" "LD_CONV_DBG=1" "LD_CONV_DBG=2" "LD_CONV_DBG=4" "LD_CONV_DBG=6" "LD_CONV_DBG=7" "LD_CONV_MT=1" "LD_CONV_MT=1 LD_CONV_DBG=6"; do
  env $e python tools/gpu_conv_one.py 256 0 32 256 3 0 $N | sed "s/^/$e  /"
done
for n in 16 24 37 64; do python tools/gpu_conv_one.py 256 0 32 256 3 0 $n; done
python tools/gpu_conv_one.py 128 0 64 128 3 0 32; python tools/gpu_conv_one.py 64 0 128 64 3 0 32
