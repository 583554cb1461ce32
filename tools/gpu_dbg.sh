#!/bin/bash
set -u
mkdir -p gpurun_out; O=gpurun_out
export IE_SPIN_LIMIT_MS=2000
for args in "512 1" "2048 1" "4096 1" "4097 1" "5000 1" "8193 1" "20000 1" "20000 300"; do
  echo "== $args"; timeout 90 python tools/debug_long.py $args 2>&1 | tail -3; echo "rc=$?"
done
echo "== IE_CHUNK_T=1024 20000"
IE_CHUNK_T=1024 timeout 90 python tools/debug_long.py 20000 1 2>&1 | tail -3
echo "== IE_SEQ=0 6000"
IE_SEQ=0 timeout 120 python tools/debug_long.py 6000 1 2>&1 | tail -3
echo "== remaining tests"
timeout 900 python -m pytest tests -m gpu -v --timeout 300 --timeout-method=thread -k "full_size or oom or inference_wrapper or mlp or threshold" 2>&1 | tail -25 | cut -c1-250
