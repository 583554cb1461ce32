#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/profile_step.py --B 768 --T 512 --iters 3 2>&1 | tail -1
timeout 400 python tools/gpu_check.py --only wide --timeout 250 --log gpurun_out/exp_check.log 2>&1 | grep -E "status|rror|TIMEOUT" | cut -c1-400
grep -o '"min_cosine": [0-9.]*\|"rel_l2": [0-9.e-]*' gpurun_out/exp_check.log | tr '\n' ' '; echo
