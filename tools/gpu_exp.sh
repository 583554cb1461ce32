#!/bin/bash
mkdir -p gpurun_out
for fm in 0 1; do
  echo "=== IE_FAST_MATH=$fm"
  IE_FAST_MATH=$fm timeout 400 python tools/gpu_check.py --only r4_small,n3b --timeout 150 --log gpurun_out/exp_check_fm$fm.log 2>&1 | grep -E "status|rror|TIMEOUT" | cut -c1-300
  grep -o '"min_cosine": [0-9.]*\|"rel_l2": [0-9.e-]*\|"max_abs": [0-9.e-]*\|"launches": [0-9]*' gpurun_out/exp_check_fm$fm.log | tr '\n' ' '; echo
  IE_FAST_MATH=$fm timeout 120 python tools/profile_step.py --B 256 --T 512 --iters 3 2>&1 | tail -1
  IE_FAST_MATH=$fm timeout 120 python tools/trace_seq.py --T 96 --layer 1 2>&1 | sed -n 2,12p
done
IE_FAST_MATH=0 timeout 120 python tools/trace_seq.py --T 96 --layer 3 2>&1 | sed -n 2,12p
