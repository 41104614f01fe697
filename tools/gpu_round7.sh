#!/bin/bash
TAG=${1:-r01j}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 python tools/profile_step.py --cull 0 > $O/ps_dense.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_dense.csv python tools/profile_step.py --cull 0 > $O/ncu_dense.log 2>&1
timeout 600 python tools/profile_step.py --cull 80 > $O/ps_cull.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_cull.csv python tools/profile_step.py --cull 80 > $O/ncu_cull.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dgemm_sl -s 6 -c 3 -o $O/sl_step_prof python tools/profile_step.py --cull 0 > $O/ncu_sl.log 2>&1
tail -3 $O/pytest_gpu.log; cat $O/bench.json $O/bench_ref.json $O/ps_dense.log $O/ps_cull.log; tail -2 $O/bench.err
