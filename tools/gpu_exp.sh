#!/bin/bash
mkdir -p gpurun_out
export IE_FAST_MATH=1
timeout 120 python tools/umma_rate.py 2>&1 | tail -12
timeout 500 python tools/gpu_check.py --only gemm,dual,r4_small --timeout 200 --log gpurun_out/exp_check.log 2>&1 | grep -E "status|rror|TIMEOUT|max_abs" | cut -c1-1500
grep -o '"min_cosine": [0-9.]*\|"rel_l2": [0-9.e-]*\|"launches": [0-9]*' gpurun_out/exp_check.log | tr '\n' ' '; echo
timeout 120 python tools/profile_step.py --B 512 --T 512 --iters 3 2>&1 | tail -1
IE_GEMM_PAIR=0 timeout 120 python tools/profile_step.py --B 512 --T 512 --iters 3 2>&1 | tail -1
