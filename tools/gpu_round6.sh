#!/bin/bash
TAG=${1:-r01h}
O=gpurun_out/$TAG
mkdir -p $O
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "small_left" > $O/pytest_sl.log 2>&1; rc=$?; echo "pytest sl rc=$rc" >> $O/pytest_sl.log
tail -5 $O/pytest_sl.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python tools/gemm_bench.py 512 5 > $O/gemm_bench.log 2>&1
cat $O/gemm_bench.log
