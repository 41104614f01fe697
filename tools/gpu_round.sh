#!/bin/bash
# One GPU session under gpurun: `bash tools/gpu_round.sh <tag> [steps...]`, outputs under gpurun_out/<tag>/.
#   tests    python -m pytest tests -m gpu
#   gemm     GEMM micro-benchmark of the three DMMA kernels on the contraction shapes of one dense chunk
#   steps    one evaluation at the bench shape: dense, culled, frozen (tools/profile_step.py)
#   bench    bench.py (ours + reference arm)
#   metrics  ncu pass over one evaluation: time, DRAM bytes and FP64 instruction counts of every launch (tools/summarize_metrics.py)
#   restarts bench.py --mode restarts at every shape (one replica per GPU)
#   launches ncu launch lists of one dense and one culled evaluation
#   ncu-sl / ncu-sym / ncu-axx   ncu --set full of the named kernel inside one evaluation (CULL=746 exact-zero windows by default, NCAP launches)
TAG=${1:-run}; shift
STEPS=${@:-tests gemm steps}
O=gpurun_out/$TAG
mkdir -p $O
for S in $STEPS; do
  case $S in
    tests) timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log;;
    gemm) timeout 300 python tools/gemm_bench.py 512 5 > $O/gemm_bench.log 2>&1; cat $O/gemm_bench.log;;
    steps) timeout 600 python tools/profile_step.py --cull 0 > $O/ps_dense.log 2>&1
           timeout 600 python tools/profile_step.py --cull 80 > $O/ps_cull.log 2>&1
           timeout 600 python tools/profile_step.py --cull 80 --mode 0 > $O/ps_frozen.log 2>&1
           cat $O/ps_dense.log $O/ps_cull.log $O/ps_frozen.log;;
    bench) timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err
           timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
           cat $O/bench.json $O/bench_ref.json; tail -2 $O/bench.err;;
    launches) timeout 600 python tools/profile_step.py --cull 0 > $O/ps_dense.log 2>&1 && \
           timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_dense.csv python tools/profile_step.py --cull 0 > $O/ncu_dense.log 2>&1
           timeout 600 python tools/profile_step.py --cull 80 > $O/ps_cull.log 2>&1 && \
           timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_cull.csv python tools/profile_step.py --cull 80 > $O/ncu_cull.log 2>&1;;
    metrics) timeout 600 python tools/profile_step.py --cull ${CULL:-746} > $O/ps_step.log 2>&1 && \
           timeout 1800 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed_pipe_tensor_op_dmma.sum,sm__pipe_fp64_cycles_active.avg,sm__cycles_active.avg --clock-control none --profile-from-start off --csv --log-file $O/step_metrics.csv python tools/profile_step.py --cull ${CULL:-746} > $O/ncu_metrics.log 2>&1; tail -2 $O/ncu_metrics.log;;
    restarts) timeout 900 python bench.py --mode restarts --shape all --steps 5 --warmup 3 > $O/bench_restarts.json 2> $O/bench_restarts.err; cat $O/bench_restarts.json; tail -2 $O/bench_restarts.err;;
    ncu-sl|ncu-sym|ncu-axx)
           K=${S#ncu-}; R=dgemm_sl; [ $K = sym ] && R=dgemm_sym; [ $K = axx ] && R=axx_sum
           SK=6; [ $K = axx ] && SK=0
           timeout 600 python tools/profile_step.py --cull ${CULL:-746} > $O/ps_step.log 2>&1 && \
           timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$R -s $SK -c ${NCAP:-3} -o $O/${K}_prof python tools/profile_step.py --cull ${CULL:-746} > $O/ncu_$K.log 2>&1; tail -2 $O/ncu_$K.log;;
  esac
done
