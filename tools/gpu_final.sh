#!/bin/bash
# Evidence run of the current state on one GPU: `bash tools/gpu_final.sh <tag>`, outputs under gpurun_out/<tag>/.
#   tests, bench (ours + reference arm), launch list and metrics pass (DRAM bytes, FP64 instruction counts) of one
#   evaluation at the bench shape, `ncu --set full` captures of the hot kernels, the GEMM shapes micro-benchmark,
#   the binary128 report, named shapes.
TAG=${1:-final}; O=gpurun_out/$TAG; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; cut -c1-400 $O/bench.json; tail -2 $O/bench.err
timeout 300 python tools/gemm_shapes.py > $O/gemm_shapes.log 2>&1; cat $O/gemm_shapes.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_746.csv python tools/profile_step.py --cull 746 > $O/ncu_launches.log 2>&1
timeout 1800 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed_pipe_tensor_op_dmma.sum,sm__pipe_fp64_cycles_active.avg,sm__cycles_active.avg --clock-control none --profile-from-start off --csv --log-file $O/step_metrics.csv python tools/profile_step.py --cull 746 > $O/ncu_metrics.log 2>&1
for K in "dgemm_sl:6:3:sl" "dgemm_sym:6:3:sym" "ahx_gen_sep:4:1:gen" "ahx_dot_sep:4:1:dot" "axx_sum:0:1:axx"; do
  IFS=: read R SK NC NM <<< "$K"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$R -s $SK -c $NC -o $O/${NM}_prof python tools/profile_step.py --cull 746 > $O/ncu_$NM.log 2>&1; tail -1 $O/ncu_$NM.log
  # gpurun copies back at most 64 MiB: keep the raw and source pages as CSV, drop the report
  ncu -i $O/${NM}_prof.ncu-rep --page raw --csv > $O/${NM}_raw.csv 2>/dev/null
  ncu -i $O/${NM}_prof.ncu-rep --page source --csv > $O/${NM}_src.csv 2>/dev/null
  rm -f $O/${NM}_prof.ncu-rep
done
timeout 600 python tools/quad_report.py > $O/quad_truth.jsonl 2> $O/quad.err
timeout 600 python tools/named_shapes.py > $O/named.jsonl 2> $O/named.err
ls -la $O; du -sh $O
