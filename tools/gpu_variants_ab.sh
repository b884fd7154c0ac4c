#!/bin/bash
# usage: gpu_variants_ab.sh "ENV1=a ENV2=b" "ENV1=c" ... : the four in-situ variants of the 32-channel conv under each environment
for e in "$@"; do
  echo "== $e"
  for v in 0 1 2 3; do env $e python tools/gpu_conv_variant_one.py $v; done
done
