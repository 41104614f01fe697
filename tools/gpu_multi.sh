#!/bin/bash
G=${1:-2}
TAG=${2:-r01m}
O=gpurun_out/$TAG
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 tools/check_multi.py 20000 96 > $O/check_multi_$G.log 2>&1; echo "rc=$?" >> $O/check_multi_$G.log
grep -v "^\*\*\|OMP_NUM\|^$" $O/check_multi_$G.log | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 5 --warmup 3 > $O/bench_$G.json 2> $O/bench_$G.err; echo "bench rc=$?"
grep '"metric"' $O/bench_$G.json | cut -c1-700
tail -3 $O/bench_$G.err
