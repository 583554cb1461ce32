#!/bin/bash
# Multi-GPU session (gpurun --gpus N): two devices from one process, then bench.py under torchrun at N GPUs
set -u
mkdir -p gpurun_out
O=gpurun_out; N=${1:-2}; TAG=${2:-m$N}
echo "== devices"; nvidia-smi -L | head -8
echo "== two devices from one process"
timeout 600 python -m pytest tests -m gpu -q -k "two_devices" > $O/pytest_two_devices_$TAG.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_two_devices_$TAG.log
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $O/bench_${N}gpu_$TAG.json 2> $O/bench_${N}gpu_$TAG.err
echo "rc=$?"; tail -3 $O/bench_${N}gpu_$TAG.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/bench_${N}gpu_$TAG.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('n_gpus %d value %.0f  e2e %.0f  ms/step %.2f  frac %.3f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac']))
    print('bulk_varlen', d['extra'].get('bulk_varlen'))
except Exception as e: print('no line', e)
PY
