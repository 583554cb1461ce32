#!/bin/bash
# quick experiment: parity + phase times for each h-multicast cluster size
mkdir -p gpurun_out
for c in 4 8 2 1; do
  echo "=== cluster $c"
  IE_STEP_CLUSTER=$c timeout 300 python tools/gpu_check.py --only tiny,r4_small --timeout 200 --log gpurun_out/exp_check_c$c.log 2>&1 | grep -E "status|min_cosine" | cut -c1-400
  IE_STEP_CLUSTER=$c timeout 120 python tools/profile_step.py --B 256 --T 512 --iters 3 2>&1 | tail -1
done
