#!/bin/bash
# One GPU session: parity tests, GEMM micro-bench, bench line, ncu launch lists. Outputs under gpurun_out/<tag>/.
TAG=${1:-r01}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python tools/gemm_bench.py 512 5 > $O/gemm_bench.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_dense.json 2> $O/bench_dense.err
timeout 600 python tools/profile_step.py --cull 0 > $O/ps_dense.log 2>&1
timeout 600 python tools/profile_step.py --cull 80 > $O/ps_cull.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_dense.csv python tools/profile_step.py --cull 0 > $O/ncu_dense.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_cull.csv python tools/profile_step.py --cull 80 > $O/ncu_cull.log 2>&1
tail -3 $O/pytest_gpu.log; cat $O/gemm_bench.log; cat $O/bench_dense.json; cat $O/ps_dense.log $O/ps_cull.log
