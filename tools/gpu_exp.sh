#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_exp.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_exp.log
timeout 120 python tools/profile_step.py --B 512 --T 512 --iters 3 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-900
