#!/bin/bash
TAG=${1:-r01d}
O=gpurun_out/$TAG
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python tools/profile_step.py --cull 0 > $O/ps_dense.log 2>&1
timeout 600 python tools/profile_step.py --cull 80 > $O/ps_cull.log 2>&1
timeout 300 python tools/gemm_bench.py 512 3 > $O/gemm_bench.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_sym -c 4 -o $O/sym_prof python tools/gemm_bench.py 512 1 > $O/sym_ncu.log 2>&1
tail -5 $O/pytest_gpu.log; cat $O/gemm_bench.log $O/ps_dense.log $O/ps_cull.log
