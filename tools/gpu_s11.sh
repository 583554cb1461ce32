#!/bin/bash
# GPU session: last layer with its input projection fused into the recurrent K loop (IE_FUSE_LAST, default on):
# parity tests of the new path, A/B against the hoisted form in sustained runs, per-item timeline of the fused layer
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s11}
echo "== tests (fused path, bit identity, goldens)"
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread -k "fused_last or every_path or golden or raw_features or very_long or bulk_equals" > $O/pytest_gpu_$TAG.log 2>&1
echo "rc=$?"; tail -15 $O/pytest_gpu_$TAG.log | cut -c1-300
echo "== A/B in sustained runs"
timeout 600 python tools/power_probe.py --seconds 3 --what "enc,enc:IE_FUSE_LAST=0,enc,enc:IE_FUSE_LAST=0" > $O/power_$TAG.jsonl 2> $O/power_$TAG.err
echo "rc=$?"; cut -c1-900 $O/power_$TAG.jsonl; tail -3 $O/power_$TAG.err
echo "== trace of the fused last layer"
timeout 120 python tools/trace_layer.py --B 1280 --T 128 --layer 3 > $O/trace_l3_$TAG.log 2>&1; echo "rc=$?"; sed -n 2,14p $O/trace_l3_$TAG.log
