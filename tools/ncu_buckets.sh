#!/bin/bash
# BASELINE configs[2]: per-bucket ncu tensor-pipe activity (one 1280-issue encode per seq_len bucket), aggregated per kernel
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-nb}
for T in 64 128 256 512 1024 2048; do
  # application replay: the cooperative cluster launch of the persistent kernel does not survive ncu's kernel replay
  # when more than one pass is needed
  timeout 900 ncu --replay-mode application --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    -k regex:"lstm_layer|gemm_bf16_pair" --clock-control none --csv --log-file $O/ncu_bucket_${T}_$TAG.csv python tools/profile_step.py --B 1280 --T $T --warm 1 --iters 1 > $O/ncu_bucket_${T}_$TAG.out 2>&1
done
python - <<PY
import csv, json
out = {}
for T in (64, 128, 256, 512, 1024, 2048):
    try:
        rows = [r for r in csv.reader(open('$O/ncu_bucket_%d_$TAG.csv' % T)) if len(r) > 10]
    except Exception as e:
        continue
    h = rows[0]; ki = h.index('Kernel Name'); mi = h.index('Metric Name'); vi = h.index('Metric Value'); ii = h.index('ID')
    ui = h.index('Metric Unit')
    def fl(x):
        try: return float(x.replace(',', ''))
        except Exception: return 0.0
    per = {}
    for r in rows[1:]:
        v = fl(r[vi])
        if r[mi] == 'gpu__time_duration.sum': v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(r[ui], 1.0)          # -> ns
        if r[mi].startswith('dram__bytes'): v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[ui], 1.0)  # -> bytes
        per.setdefault(r[ii], {'k': r[ki]})[r[mi]] = v
    # the second encode only (skip the warm-up call): launches are in order; take the last half of the ie:: launches
    ls = [v for k, v in sorted(per.items(), key=lambda x: int(x[0])) if 'ie::' in v['k'] and 'convert_rows' not in v['k']]
    ls = ls[len(ls) // 2:]
    agg = {}
    for v in ls:
        name = 'lstm_layer' if 'lstm_layer' in v['k'] else ('gemm' if 'gemm' in v['k'] else 'other')
        a = agg.setdefault(name, [0.0, 0.0, 0.0])
        t = v.get('gpu__time_duration.sum', 0.0)
        a[0] += t; a[1] += t * v.get('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 0.0)
        a[2] += v.get('dram__bytes_read.sum', 0.0) + v.get('dram__bytes_write.sum', 0.0)
    tot = sum(a[0] for a in agg.values())
    out[T] = {k: {'ms': a[0] / 1e6, 'tensor_pipe_active_pct': (a[1] / a[0] if a[0] else 0.0), 'dram_gb': a[2] / 1e9} for k, a in agg.items()}
    out[T]['all'] = {'ms': tot / 1e6, 'tensor_pipe_active_pct': sum(a[1] for a in agg.values()) / tot if tot else 0.0}
    print(T, json.dumps(out[T]))
json.dump(out, open('$O/ncu_buckets_$TAG.json', 'w'), indent=1)
PY
