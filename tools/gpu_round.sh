#!/bin/bash
# One GPU session: tests, smoke, bench (both arms), ncu launch list + full captures.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:lstm_|gemm_bf16|embed_gather|pool_finalize' -s 10 -c 10 --csv --log-file gpurun_out/launches_$TAG.csv python tools/profile_step.py --B 768 --T 512 --iters 1 --warm 1 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_wide -s 4 -c 1 -o gpurun_out/prof_seq_$TAG -f python tools/profile_step.py --B 768 --T 128 --iters 1 --warm 1 > gpurun_out/ncu_seq_$TAG.log 2>&1; echo "ncu seq rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_pair -s 5 -c 1 -o gpurun_out/prof_gemm_$TAG -f python tools/profile_step.py --B 768 --T 128 --iters 1 --warm 1 > gpurun_out/ncu_gemm_$TAG.log 2>&1; echo "ncu gemm rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$TAG.json
ls -la gpurun_out | tail -12
