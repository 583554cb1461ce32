#!/bin/bash
# GPU session: full GPU test-suite, bench (both arms), layer traces
set -u
mkdir -p gpurun_out
O=gpurun_out; TAG=${1:-s2}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -v -s --timeout 400 --timeout-method=thread > $O/pytest_gpu_$TAG.log 2>&1
echo "rc=$?"; grep -E "passed|failed|error" $O/pytest_gpu_$TAG.log | tail -5; grep -E "FAILED|ERROR|Timeout" $O/pytest_gpu_$TAG.log | head -20
grep -E "^(r4|encoder_|n3|fp32|varlen|full-size)" $O/pytest_gpu_$TAG.log | cut -c1-400
echo "== bench"
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "rc=$?"; tail -3 $O/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$TAG.json').read().strip().splitlines()[-1])
    print('value %.0f  e2e %.0f  single %.0f  ms/step %.2f  frac %.3f whole %.3f' % (d['value'], d['e2e']['value'], d['single_batch']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac']))
    print('clocks', d['clocks']); print('phases', d['roofline']['phase_ms_last_call']); print('mhz', d['roofline']['phase_sm_mhz'])
    print('extra', json.dumps(d['extra'])[:1500]); print('cpu', d.get('cpu_baseline'))
except Exception as e: print('no line', e)
PY
echo "== traces"
timeout 120 python tools/trace_layer.py --B 1280 --T 128 --layer 1 > $O/trace_l1_$TAG.log 2>&1; echo "rc=$?"; head -16 $O/trace_l1_$TAG.log
timeout 120 python tools/trace_layer.py --B 1280 --T 128 --layer 3 > $O/trace_l3_$TAG.log 2>&1; echo "rc=$?"; head -16 $O/trace_l3_$TAG.log
echo "== power probe A/B"
timeout 300 python tools/power_probe.py --seconds 3 --what "cublas,enc,enc:IE_GEMM_TMA_STORE=0,enc:IE_GX_BF16=0" > $O/power_$TAG.jsonl 2> $O/power_$TAG.err
echo "rc=$?"; cut -c1-700 $O/power_$TAG.jsonl; tail -3 $O/power_$TAG.err
