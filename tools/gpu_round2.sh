#!/bin/bash
# First GPU session of round 2 (tools/README.md checklist, items 0-4) in one gpurun call; every step under its own
# timeout, outputs in gpurun_out/.  Roughly 8 GPU-minutes.
#   gpurun --timeout 900 -- 'bash tools/gpu_round2.sh r8'
set -u
mkdir -p gpurun_out
TAG=${1:-r8}
O=gpurun_out
echo "== 0. opt-in knobs reproduce the default bits"
IE_TEST_EXPERIMENTAL=1 timeout 240 python -m pytest tests -m gpu -x -q -k "experimental or rotating" > $O/pytest_experimental_$TAG.log 2>&1
echo "rc=$?"; tail -3 $O/pytest_experimental_$TAG.log
echo "== 1. R4-size equality + timing of every variant"
timeout 300 python tools/gpu_rot.py --skip-small --proj --poolraw --rotvar --log $O/variants_$TAG.jsonl > $O/variants_$TAG.log 2>&1
echo "rc=$?"; grep -c '"equal": true' $O/variants_$TAG.log; grep '"equal": false' $O/variants_$TAG.log | head -5
grep '"ms"' $O/variants_$TAG.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('%-28s %7.1f issues/s  %6.2f ms/256  gemm %s  steps %s' % (d['path'], d['issues_per_s'], d['ms_per_256'], d['gemm'], d['steps']))"
echo "== 2. power / clock / TFLOP/s next to cuBLAS"
timeout 120 python tools/power_probe.py --what cublas,wide,rot --seconds 4 > $O/power_$TAG.jsonl 2> $O/power_$TAG.err
echo "rc=$?"; cut -c1-300 $O/power_$TAG.jsonl
echo "== 3. last-layer timelines"
timeout 90 python tools/trace_seq.py --B 768 --T 128 --layer 3 > $O/trace_seq_l3_$TAG.log 2>&1; echo "rc=$?"; head -12 $O/trace_seq_l3_$TAG.log
timeout 90 python tools/trace_rot.py --B 1280 --T 128 --layer 3 > $O/trace_rot_l3_$TAG.log 2>&1; echo "rc=$?"; head -14 $O/trace_rot_l3_$TAG.log
echo "== 4. sustained bench: default, IE_EMB_PROJ, IE_ROT (+ both)"
for cfg in "" "IE_EMB_PROJ=1" "IE_ROT=1" "IE_ROT=1 IE_EMB_PROJ=1 IE_POOL_RAW=1"; do
  name=$(echo "default $cfg" | tr ' =' '__')
  env $cfg timeout 200 python bench.py --no-cpu-baseline --steps 15 > $O/bench_${TAG}_$name.json 2> $O/bench_${TAG}_$name.err
  echo "[$cfg] rc=$?"
  python -c "
import json,sys
try:
    d=json.loads(open('$O/bench_${TAG}_$name.json').read().strip().splitlines()[-1])
    print('   value %.0f  e2e %.0f  ms/step %.2f  frac %.3f  clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']))
except Exception as e: print('   no line:', e)"
done
ls -la $O | tail -20
