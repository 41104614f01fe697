#!/bin/bash
# Reduced evidence run (one GPU, ~4 min): bench (ours + reference arm), launch list, metrics pass, ncu of dgemm_sl, GEMM
# shapes, binary128 report.  `bash tools/gpu_final_short.sh <tag>`
TAG=${1:-final}; O=gpurun_out/$TAG; mkdir -p $O
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; cut -c1-300 $O/bench.json; tail -2 $O/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_746.csv python tools/profile_step.py --cull 746 > $O/ncu_launches.log 2>&1
timeout 1800 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed_pipe_tensor_op_dmma.sum,sm__pipe_fp64_cycles_active.avg,sm__cycles_active.avg --clock-control none --profile-from-start off --csv --log-file $O/step_metrics.csv python tools/profile_step.py --cull 746 > $O/ncu_metrics.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dgemm_sl -s 6 -c 4 -o $O/sl_prof python tools/profile_step.py --cull 746 > $O/ncu_sl.log 2>&1
ncu -i $O/sl_prof.ncu-rep --page raw --csv > $O/sl_raw.csv 2>/dev/null; rm -f $O/sl_prof.ncu-rep
timeout 600 python tools/quad_report.py > $O/quad_truth.jsonl 2> $O/quad.err
du -sh $O
