#!/bin/bash
# BASELINE configs[2]: per seq_len bucket, one `ncu --set full` capture (kernel replay) of a wide-layer launch of
# lstm_layer_kernel and of the input-projection GEMM inside a 1280-issue encode -> tensor-pipe activity, L2 -> SM rate, DRAM bytes
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-nbf}
for T in 64 128 256 512 1024 2048; do
  for K in layer gemm; do
    if [ $K = layer ]; then RX="regex:lstm_layer_kernel"; SK=4; else RX="regex:gemm_bf16_pair"; SK=3; fi
    timeout 600 ncu --set full --clock-control none -k $RX -s $SK -c 1 -o $O/nbf_${K}_${T}_$TAG -f python tools/profile_step.py --B 1280 --T $T --warm 1 --iters 1 > $O/nbf_${K}_${T}_$TAG.log 2>&1
    ncu -i $O/nbf_${K}_${T}_$TAG.ncu-rep --page raw --csv > $O/nbf_${K}_${T}_$TAG.csv 2>/dev/null
    rm -f $O/nbf_${K}_${T}_$TAG.ncu-rep
  done
done
python - <<PY
import csv, json
out = {}
for T in (64, 128, 256, 512, 1024, 2048):
    out[T] = {}
    for K in ('layer', 'gemm'):
        try:
            rows = list(csv.reader(open('$O/nbf_%s_%d_$TAG.csv' % (K, T))))
            d = dict(zip(rows[0], rows[2]))
            tp = [v for k, v in d.items() if k.endswith('sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed')]
            out[T][K] = {'kernel': d.get('Kernel Name', '')[:60], 'ms': float(d['gpu__time_duration.sum']),
                         'tensor_pipe_active_pct': float(tp[0]) if tp else None,
                         'l2_to_sm_tbps': float(d['l1tex__m_xbar2l1tex_read_bytes.sum.per_second']),
                         'dram_read_gb': float(d['dram__bytes_read.sum']), 'dram_write_gb': float(d['dram__bytes_write.sum']),
                         'sm_ghz': float(d['gpc__cycles_elapsed.max.per_second'])}
        except Exception as e:
            out[T][K] = {'error': repr(e)}
    print(T, json.dumps(out[T]))
out['note'] = ('one launch per bucket and kernel, ncu --set full --clock-control none, kernel replay, inside a 1280-issue encode with lengths = T; '
               'units as ncu prints them (ms, %, Tbyte/s, Gbyte, GHz)')
json.dump(out, open('$O/ncu_buckets_full_$TAG.json', 'w'), indent=1)
PY
