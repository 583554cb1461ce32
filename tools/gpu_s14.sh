#!/bin/bash
# GPU session: MLP head -- device-resident timing under knobs (N tile, rows per pass), its ncu launch list, MLP tests
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s14}
echo "== MLP tests"
timeout 600 python -m pytest tests -m gpu -q --timeout 400 --timeout-method=thread -k "mlp or threshold or gemm" 2>&1 | tail -3
echo "== MLP probe"
timeout 600 python tools/mlp_probe.py --what "default,IE_MLP_CHUNK=65536,default" > $O/mlp_probe_$TAG.jsonl 2> $O/mlp_probe_$TAG.err
echo "rc=$?"; cat $O/mlp_probe_$TAG.jsonl; tail -3 $O/mlp_probe_$TAG.err
echo "== ncu launch list of the MLP (n = 2^18)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_mlp_$TAG.csv python tools/mlp_probe.py --what default --n 262144 --iters 1 > $O/ncu_mlp_$TAG.log 2>&1
echo "rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open('$O/launches_mlp_$TAG.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    if v > 20000: print('%-70s %10.3f ms' % (r[ki].split('(')[0][:70], v/1e6))
PY
