#!/bin/bash
# GPU session: L2 residency experiments for the recurrent kernel's operand streams
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s6}
echo "== parity under the knobs"
IE_L2_PERSIST=1 IE_H_EVICT_FIRST=1 timeout 600 python -m pytest tests -m gpu -q --timeout 300 --timeout-method=thread -k "golden_r4_bench or golden_small or edge_cases" 2>&1 | tail -3
echo "== power probe A/B"
timeout 500 python tools/power_probe.py --seconds 3 --what "enc,enc:IE_L2_PERSIST=1,enc:IE_H_EVICT_FIRST=1,enc:IE_L2_PERSIST=1+IE_H_EVICT_FIRST=1,enc" > $O/power_$TAG.jsonl 2> $O/power_$TAG.err
echo "rc=$?"; cut -c1-720 $O/power_$TAG.jsonl; tail -3 $O/power_$TAG.err
