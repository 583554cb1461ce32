#!/bin/bash
# GPU session: full GPU test-suite, smoke(), bench (driver form, both arms), seq_len sweep with the final kernels
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s16}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -v -s --timeout 400 --timeout-method=thread > $O/pytest_gpu_$TAG.log 2>&1
echo "rc=$?"; grep -E "passed|failed" $O/pytest_gpu_$TAG.log | tail -3; grep -E "FAILED|Timeout" $O/pytest_gpu_$TAG.log | head
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "rc=$?"; tail -4 $O/smoke_$TAG.log
echo "== bench --impl reference (driver form)"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref_$TAG.json 2> $O/bench_ref_$TAG.err; echo "rc=$?"; cut -c1-500 $O/bench_ref_$TAG.json
echo "== bench (driver form)"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "rc=$?"; tail -3 $O/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$TAG.json').read().strip().splitlines()[-1])
    print('value %.0f  e2e %.0f  single %.0f  ms/step %.2f  frac %.3f whole %.3f' % (d['value'], d['e2e']['value'], d['single_batch']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac']))
    print('clocks', d['clocks']); print('phases', d['roofline']['phase_ms_last_call']); print('mhz', d['roofline']['phase_sm_mhz'])
    print('extra', json.dumps(d['extra'])[:2600]); print('cpu', d.get('cpu_baseline'))
except Exception as e: print('no line', e)
PY
if [ "${SKIP_SWEEP:-0}" = "1" ]; then exit 0; fi
echo "== seq_len sweep (configs[2])"
timeout 900 python tools/sweep_seq_len.py > $O/sweep_$TAG.jsonl 2> $O/sweep_$TAG.err; echo "rc=$?"; cat $O/sweep_$TAG.jsonl | cut -c1-420
