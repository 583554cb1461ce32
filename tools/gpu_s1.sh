#!/bin/bash
# GPU session: full GPU test-suite on the consolidated kernel + power probe of the main variants
set -u
mkdir -p gpurun_out
O=gpurun_out; TAG=${1:-s1}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/pytest_gpu_$TAG.log 2>&1
echo "rc=$?"; tail -25 $O/pytest_gpu_$TAG.log
echo "== power probe"
timeout 400 python tools/power_probe.py --seconds 3 --what "cublas,enc,enc:IE_GX_BF16=0,enc:IE_EMB_PROJ=0,enc:IE_BATCHES=3,enc:IE_BATCHES=6,enc:IE_BATCHES=8,enc256" > $O/power_$TAG.jsonl 2> $O/power_$TAG.err
echo "rc=$?"; cut -c1-700 $O/power_$TAG.jsonl; tail -3 $O/power_$TAG.err
