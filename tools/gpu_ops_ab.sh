#!/bin/bash
# usage: gpu_ops_ab.sh "ENV=a" "ENV=b" ... : total per-op forward time (mri, 32 x 256 x 256) under each environment, two rounds
for r in 1 2; do for e in "$@"; do
  echo "== $e: $(env $e LD_PROFILE_OPS=60 python tools/gpu_profile_ops.py 32 256 mri 3 2>&1 | grep 'LDPROF total' | tail -2 | tr '\n' ' ')"
done; done
