#!/bin/bash
# GPU session: h-tile multicast (IE_MC=1) parity + A/B, GEMM panel sweep
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s4}
echo "== parity under IE_MC=1"
timeout 600 python -m pytest tests -m gpu -v --timeout 200 --timeout-method=thread -k "every_path or golden_small or wait_timeout" > $O/pytest_mc_$TAG.log 2>&1; echo "rc=$?"; grep -E "passed|failed" $O/pytest_mc_$TAG.log | tail -3; grep -E "FAILED|Timeout" $O/pytest_mc_$TAG.log | head
IE_MC=1 timeout 900 python -m pytest tests -m gpu -v --timeout 300 --timeout-method=thread -k "golden_r4 or full_size or golden_n3 or edge_cases or bulk_equals" > $O/pytest_mc_env_$TAG.log 2>&1; echo "rc=$?"; grep -E "passed|failed" $O/pytest_mc_env_$TAG.log | tail -3; grep -E "FAILED|Timeout" $O/pytest_mc_env_$TAG.log | head
echo "== power probe A/B"
timeout 500 python tools/power_probe.py --seconds 3 --what "enc,enc:IE_MC=1,enc:IE_GEMM_PANEL=37,enc:IE_GEMM_PANEL=16,enc:IE_GEMM_PANEL=4,enc:IE_MC=1+IE_BATCHES=6,enc" > $O/power_$TAG.jsonl 2> $O/power_$TAG.err
echo "rc=$?"; cut -c1-720 $O/power_$TAG.jsonl; tail -3 $O/power_$TAG.err
echo "== trace IE_MC=1"
IE_MC=1 timeout 120 python tools/trace_layer.py --B 1280 --T 128 --layer 1 > $O/trace_l1_mc_$TAG.log 2>&1; echo "rc=$?"; head -14 $O/trace_l1_mc_$TAG.log
