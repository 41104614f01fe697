#!/bin/bash
# Throughput of every named shape and of the scaling sweep at G GPUs of one box (BASELINE.json: "at 1, 2, 4 and 8 GPUs").
#   bash tools/multi_shapes.sh G TAG
# named shapes (toy / ou / hrir / crude, n = 400..600): independent restarts, one replica per GPU (bench.py --mode restarts);
# scaling sweep: one evaluation sharded over the G GPUs (bench.py --n N --m M), exact-zero windows.
G=${1:-2}; TAG=${2:-r02ms}; O=gpurun_out/$TAG; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
[ "$G" = 1 ] && RUN="python"
P=29600
timeout 600 $RUN $([ "$G" != 1 ] && echo --master-port $P) bench.py --mode restarts --gpus $G --shape all --steps 5 --warmup 3 > $O/restarts_$G.json 2> $O/restarts_$G.err
echo "restarts rc=$?"; cut -c1-300 $O/restarts_$G.json
for NM in "10000 50" "10000 200" "100000 50" "100000 200" "100000 400" "1000000 200" "1000000 400"; do
  set -- $NM; P=$((P+1))
  timeout 900 $RUN $([ "$G" != 1 ] && echo --master-port $P) bench.py --gpus $G --n $1 --m $2 --steps 3 --warmup 3 --no-cpu-baseline > $O/sweep_${G}_n$1_m$2.json 2> $O/sweep_${G}_n$1_m$2.err
  echo "sweep n=$1 m=$2 rc=$?"; python - <<PY
import json
try:
    d = json.loads(open('$O/sweep_${G}_n$1_m$2.json').read().strip().splitlines()[-1])
    print('  evals/s %.3f  ms %.2f  gemm frac %.3f  vs_n1 %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'] or 0, d.get('elbo_rel_diff_vs_n1')))
except Exception as e:
    print('  failed', e)
PY
done
