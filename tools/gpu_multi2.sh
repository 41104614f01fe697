#!/bin/bash
# Multi-GPU evidence on G GPUs of one box: `bash tools/gpu_multi2.sh G TAG [steps]` (steps: test bench sweep restarts)
G=${1:-2}; TAG=${2:-multi}; shift; shift; STEPS=${@:-test bench sweep}
O=gpurun_out/$TAG; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
for S in $STEPS; do
  case $S in
    test) timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi_$G.log 2>&1; tail -2 $O/pytest_multi_$G.log;;
    bench) timeout 900 $RUN --master-port 29512 bench.py --gpus $G --steps 10 --warmup 3 > $O/bench_$G.json 2> $O/bench_$G.err; echo "bench rc=$?"
           python - <<PY
import json
try:
    d = json.loads([l for l in open('$O/bench_$G.json') if l.startswith('{')][-1])
    print('  evals/s %.3f  ms %.2f  e2e %.3f  gemm frac %.3f  whole %.3f  vs_n1 %s / %s  balance %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac'], d.get('elbo_rel_diff_vs_n1'), d.get('grad_rel_diff_vs_n1'), d['config'].get('shard_balance_max_over_mean_before_each_round')))
except Exception as e:
    print('  failed', e)
PY
           tail -2 $O/bench_$G.err;;
    sweep) timeout 900 $RUN --master-port 29513 tools/multi_sweep.py > $O/sweep_$G.jsonl 2> $O/sweep_$G.err; echo "sweep rc=$?"; cut -c1-200 $O/sweep_$G.jsonl; tail -2 $O/sweep_$G.err;;
    restarts) timeout 900 $RUN --master-port 29514 bench.py --mode restarts --gpus $G --shape all --steps 5 --warmup 3 > $O/restarts_$G.json 2> $O/restarts_$G.err; echo "restarts rc=$?"; cut -c1-600 $O/restarts_$G.json; tail -2 $O/restarts_$G.err;;
  esac
done
