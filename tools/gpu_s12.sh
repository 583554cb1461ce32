#!/bin/bash
# GPU session: full GPU test-suite, A/B of the x-tile L2 prefetch in the fused last layer, trace
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s12}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread > $O/pytest_gpu_$TAG.log 2>&1
echo "rc=$?"; tail -5 $O/pytest_gpu_$TAG.log | cut -c1-300
echo "== A/B in sustained runs"
timeout 600 python tools/power_probe.py --seconds 3 --what "enc,enc:IE_FUSE_PREFETCH=0,enc,enc:IE_FUSE_PREFETCH=0" > $O/power_$TAG.jsonl 2> $O/power_$TAG.err
echo "rc=$?"; cut -c1-900 $O/power_$TAG.jsonl; tail -3 $O/power_$TAG.err
echo "== trace of the fused last layer"
timeout 120 python tools/trace_layer.py --B 1280 --T 128 --layer 3 > $O/trace_l3_$TAG.log 2>&1; echo "rc=$?"; sed -n 2,14p $O/trace_l3_$TAG.log
