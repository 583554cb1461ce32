#!/bin/bash
# GPU session: full test-suite, per-bucket ncu tensor-pipe numbers, GEMM capture after the L2 blocking, bench
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s8}
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -v -s --timeout 400 --timeout-method=thread > $O/pytest_gpu_$TAG.log 2>&1
echo "rc=$?"; grep -E "passed|failed" $O/pytest_gpu_$TAG.log | tail -3; grep -E "FAILED|Timeout" $O/pytest_gpu_$TAG.log | head
echo "== ncu per bucket"
bash tools/ncu_buckets.sh $TAG
echo "== ncu --set full: pair GEMM (layer 1 of a 1280 x 512 encode)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_pair -s 4 -c 1 -o $O/prof_gemm_$TAG -f python tools/profile_step.py --B 1280 --T 512 --warm 1 --iters 1 > $O/ncu_gemm_$TAG.log 2>&1
echo "rc=$?"
ncu -i $O/prof_gemm_$TAG.ncu-rep --page raw --csv > $O/ncu_full_gemm_$TAG.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('$O/ncu_full_gemm_$TAG.csv')))
if len(rows)>=3:
    d=dict(zip(rows[0],rows[2]))
    for k in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','derived__lts__lts2xbar_bytes.sum.per_second'):
        print(k, d.get(k))
    for k,v in d.items():
        if k.endswith('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'): print(k, v)
PY
echo "== bench"
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "rc=$?"; tail -3 $O/bench_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$TAG.json').read().strip().splitlines()[-1])
    print('value %.0f  e2e %.0f  single %.0f  ms/step %.2f  frac %.3f whole %.3f' % (d['value'], d['e2e']['value'], d['single_batch']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac']))
    print('clocks', d['clocks']); print('phases', d['roofline']['phase_ms_last_call']); print('e2e', d['e2e'])
except Exception as e: print('no line', e)
PY
