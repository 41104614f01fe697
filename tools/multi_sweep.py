"""The scaling sweep of BASELINE.json (configs[4]: N = 1e4 .. 1e6, M = 50 .. 400) at G GPUs of one box in ONE process
group: every point is one full-regime ELBO + gradient evaluation sharded over the ranks (strong scaling, exact-zero
windows, cost-balanced shards, one packed ncclAllReduce per sweep), timed on the device (CUDA events inside the
library, max over ranks).  One JSON line per point on rank 0.

    python tools/multi_sweep.py                                           # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/multi_sweep.py
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import cgpcm_b200
from cgpcm_b200.cgpcm import rebalance_costs, shard_bounds, window_costs, window_radius
from tests.workload import sweep_workload

ap = argparse.ArgumentParser()
ap.add_argument('--points', default='10000x50,10000x200,100000x50,100000x200,100000x400,1000000x200,1000000x400')
ap.add_argument('--cull', type=float, default=746.0)
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--peak', type=float, default=37.16, help='measured DMMA peak, TFLOP/s (profiles/fp64_peaks_r01.json)')
a = ap.parse_args()
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

for pt in a.points.split(','):
    n, m = [int(v) for v in pt.split('x')]
    wl = sweep_workload(n, m)
    cost = window_costs(wl['t'], wl['tx'], m, window_radius(*wl['hyp'], a.cull)) if world > 1 else None
    lo, hi = shard_bounds(n, rank, world, cost)
    eng = cgpcm_b200.Engine(m, m, causal=True, device=local_rank)
    if world > 1:
        box = [cgpcm_b200.Engine.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        eng.comm_init(box[0], rank, world)
    eng.set_option('cull', a.cull)
    eng.set_data(wl['t'][lo:hi], wl['y'][lo:hi], wl['th'], wl['tx'])
    for _ in range(a.warmup):
        out = eng.elbo_grad(wl['params'], reg=wl['reg'])
    if world > 1:                       # shards of equal measured time (as bench.py does during its warm-up)
        for _ in range(2):
            tm = eng.last_timing()
            mine = torch.tensor([tm['own_sweeps_ms']], dtype=torch.float64, device='cuda')
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            times = np.array([float(v.item()) for v in every])
            bounds = [shard_bounds(n, r, world, cost) for r in range(world)]
            cost = rebalance_costs(cost, bounds, times)
            lo, hi = shard_bounds(n, rank, world, cost)
            eng.set_data(wl['t'][lo:hi], wl['y'][lo:hi], wl['th'], wl['tx'])
            out = eng.elbo_grad(wl['params'], reg=wl['reg'])
            out = eng.elbo_grad(wl['params'], reg=wl['reg'])
    ms, flops = [], 0.0
    for _ in range(a.steps):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        out = eng.elbo_grad(wl['params'], reg=wl['reg'])
        tm = eng.last_timing()
        ms.append(tm['total_ms'])
        flops = tm['gemm_flops']
    dev = torch.tensor(ms + [flops], dtype=torch.float64, device='cuda')
    mx, sm = dev.clone(), dev.clone()
    if dist is not None:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    step_ms = float(mx[:-1].sum().item()) / a.steps
    total_flops = float(sm[-1].item())
    if rank == 0:
        print(json.dumps({'n': n, 'nh': m, 'nx': m, 'n_gpus': world, 'cull': a.cull, 'evals_per_s': 1e3 / step_ms,
                          'ms_per_step': step_ms, 'gemm_flops_per_step_all_ranks': total_flops,
                          'whole_step_frac_of_dmma_peak_per_gpu': total_flops / world / (step_ms * 1e-3) / 1e12 / a.peak, 'elbo': out[0]}), flush=True)
    # every rank closes its handle at the same point (ncclCommDestroy waits for the other ranks); the sweep stores of
    # one point must be gone before the next point allocates its own
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    eng.close()
    del eng
    torch.cuda.empty_cache()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
