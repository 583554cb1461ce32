#!/bin/bash
set -u
mkdir -p gpurun_out; O=gpurun_out; TAG=${1:-s9}
echo "== ncu per bucket (application replay)"
bash tools/ncu_buckets.sh $TAG
tail -3 $O/ncu_bucket_512_$TAG.out | cut -c1-300
