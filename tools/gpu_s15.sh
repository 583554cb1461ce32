#!/bin/bash
# GPU session: h-tile multicast (IE_MC=1) revisited after the fabric-bound finding: sustained A/B and per-item timelines
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s15}
timeout 600 python tools/power_probe.py --seconds 3 --what "enc,enc:IE_MC=1,enc:IE_MC=1+IE_BATCHES=6,enc:IE_MC=1+IE_BATCHES=8" > $O/power_$TAG.jsonl 2> $O/power_$TAG.err
echo "rc=$?"; cut -c1-900 $O/power_$TAG.jsonl; tail -3 $O/power_$TAG.err
echo "== trace layer 1, default"
timeout 120 python tools/trace_layer.py --B 1280 --T 128 --layer 1 > $O/trace_l1_$TAG.log 2>&1; echo "rc=$?"; sed -n 1,14p $O/trace_l1_$TAG.log
echo "== trace layer 1, IE_MC=1"
IE_MC=1 timeout 120 python tools/trace_layer.py --B 1280 --T 128 --layer 1 > $O/trace_l1_mc_$TAG.log 2>&1; echo "rc=$?"; sed -n 1,14p $O/trace_l1_mc_$TAG.log
