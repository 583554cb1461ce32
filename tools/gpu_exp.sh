#!/bin/bash
# quick experiment: parity + phase times, persistent (IE_SEQ=1) vs per-step (IE_SEQ=0)
mkdir -p gpurun_out
for sq in 1; do
  echo "=== IE_SEQ=$sq"
  IE_SEQ=$sq timeout 400 python tools/gpu_check.py --only tiny,r4_small,n3 --timeout 120 --log gpurun_out/exp_check_seq$sq.log 2>&1 | grep -E "status|min_cosine|rror|TIMEOUT" | cut -c1-600
  IE_SEQ=$sq timeout 120 python tools/profile_step.py --B 256 --T 512 --iters 3 2>&1 | tail -2
  IE_SEQ=$sq timeout 120 python tools/profile_step.py --B 256 --T 64 --iters 3 2>&1 | tail -2
done
