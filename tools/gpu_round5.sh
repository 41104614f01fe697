#!/bin/bash
TAG=${1:-r01f}
O=gpurun_out/$TAG
mkdir -p $O
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "small_left" > $O/pytest_sl.log 2>&1; rc=$?; echo "pytest sl rc=$rc" >> $O/pytest_sl.log
tail -15 $O/pytest_sl.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python tools/gemm_bench.py 512 5 > $O/gemm_bench.log 2>&1
timeout 600 python tools/profile_step.py --cull 0 > $O/ps_dense.log 2>&1
timeout 600 python tools/profile_step.py --cull 80 > $O/ps_cull.log 2>&1
timeout 600 python tools/profile_step.py --cull 80 --mode 0 > $O/ps_frozen.log 2>&1
tail -5 $O/pytest_gpu.log; cat $O/gemm_bench.log $O/ps_dense.log $O/ps_cull.log $O/ps_frozen.log
