#!/bin/bash
TAG=${1:-r01g}
O=gpurun_out/$TAG
mkdir -p $O
timeout 300 python tools/gemm_bench.py 512 1 > $O/gemm_bench.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_sl -c 4 -o $O/sl_prof python tools/gemm_bench.py 512 1 > $O/sl_ncu.log 2>&1
cat $O/gemm_bench.log; tail -3 $O/sl_ncu.log
