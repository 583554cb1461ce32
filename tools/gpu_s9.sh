#!/bin/bash
# GPU session: ncu evidence on the final kernels: --set full capture of lstm_layer_kernel (kernel replay), launch list,
# per-bucket tensor-pipe numbers (application replay), quick test subset
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s9}
echo "== quick tests"
timeout 900 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread -k "every_path or golden_small or golden_r4_bench or wait_timeout or shape_sweep" 2>&1 | tail -3
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$TAG.csv python tools/profile_step.py --B 1280 --T 512 --warm 1 --iters 1 > $O/ncu_list_$TAG.log 2>&1
echo "rc=$?"; tail -1 $O/ncu_list_$TAG.log | cut -c1-200
echo "== ncu --set full lstm_layer_kernel (kernel replay)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_layer_kernel -s 5 -c 1 -o $O/prof_layer_$TAG -f python tools/profile_step.py --B 1280 --T 512 --warm 1 --iters 1 > $O/ncu_layer_$TAG.log 2>&1
echo "rc=$?"; grep -E "ERROR|LaunchFailed" $O/ncu_layer_$TAG.log | head -3
ncu -i $O/prof_layer_$TAG.ncu-rep --page raw --csv > $O/ncu_full_layer_$TAG.csv 2>/dev/null
python - <<PY
import csv
try:
    rows=list(csv.reader(open('$O/ncu_full_layer_$TAG.csv')))
    d=dict(zip(rows[0],rows[2]))
    for k in ('Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','derived__lts__lts2xbar_bytes.sum.per_second'):
        print(k, d.get(k,'')[:100])
    for k,v in d.items():
        if k.endswith('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed'): print(k, v)
except Exception as e: print('no capture', e)
PY
echo "== ncu per bucket (application replay)"
bash tools/ncu_buckets.sh $TAG
tail -2 $O/ncu_bucket_512_$TAG.out | cut -c1-300
echo "== batches per launch A/B"
timeout 600 python tools/power_probe.py --seconds 3 --what "enc,enc:IE_BATCHES=8,enc:IE_BATCHES=10,enc:IE_BATCHES=12,enc" > $O/power_$TAG.jsonl 2> $O/power_$TAG.err
echo "rc=$?"; cut -c1-720 $O/power_$TAG.jsonl; tail -3 $O/power_$TAG.err
