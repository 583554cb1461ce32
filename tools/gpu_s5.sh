#!/bin/bash
# GPU session: full test-suite, smoke(), bench, traces (after the full-sector store epilogue + MLP tile change)
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s5}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -v -s --timeout 400 --timeout-method=thread > $O/pytest_gpu_$TAG.log 2>&1
echo "rc=$?"; grep -E "passed|failed" $O/pytest_gpu_$TAG.log | tail -3; grep -E "FAILED|Timeout" $O/pytest_gpu_$TAG.log | head
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "rc=$?"; tail -4 $O/smoke_$TAG.log
echo "== bench"
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "rc=$?"; tail -3 $O/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$TAG.json').read().strip().splitlines()[-1])
    print('value %.0f  e2e %.0f  single %.0f  ms/step %.2f  frac %.3f whole %.3f' % (d['value'], d['e2e']['value'], d['single_batch']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac']))
    print('clocks', d['clocks']); print('phases', d['roofline']['phase_ms_last_call']); print('mhz', d['roofline']['phase_sm_mhz'])
    print('e2e', d['e2e']); print('extra', json.dumps(d['extra'])[:2500]); print('cpu', d.get('cpu_baseline'))
except Exception as e: print('no line', e)
PY
echo "== bench --impl reference (short)"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "rc=$?"; cut -c1-400 $O/bench_ref_$TAG.json
echo "== traces"
timeout 120 python tools/trace_layer.py --B 1280 --T 128 --layer 1 > $O/trace_l1_$TAG.log 2>&1; echo "rc=$?"; sed -n 2,12p $O/trace_l1_$TAG.log
timeout 120 python tools/trace_layer.py --B 1280 --T 128 --layer 3 > $O/trace_l3_$TAG.log 2>&1; echo "rc=$?"; sed -n 2,12p $O/trace_l3_$TAG.log
