"""Summarise an ncu metrics pass over ONE evaluation (tools/gpu_round.sh step `metrics`) by kernel: time, DRAM bytes,
FP64 work executed.  Writes a JSON summary (and profiles/traffic.json with --traffic) from the CSV launch list.

    python tools/summarize_metrics.py gpurun_out/<tag>/step_metrics.csv [--traffic profiles/traffic.json]

FP64 work: DFMA = 2 flop, DMUL / DADD = 1 flop per thread-instruction; one DMMA.8x8x4 warp-instruction = 512 flop
(8 x 8 x 4 multiply-adds).  DMMA and the scalar FP64 instructions issue to the same pipe (profiles/fp64_peaks_r01.json:
15.75 + 15.75 TFLOP/s when mixed), so their sum over the step time is the utilisation of that pipe.
"""
import csv
import json
import re
import sys
from collections import defaultdict

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
per = defaultdict(lambda: defaultdict(float))
launch_ids = defaultdict(set)
for r in rows:
    name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '')
    m = r['Metric Name']
    try:
        v = float(r['Metric Value'].replace(',', ''))
    except ValueError:
        continue
    unit = r.get('Metric Unit', '')
    if m == 'gpu__time_duration.sum':
        v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
    if m.startswith('dram__bytes'):
        v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)
    per[name][m] += v
    launch_ids[name].add(r['ID'])
out = {'kernels': {}, 'source': path}
tot = defaultdict(float)
for name, d in per.items():
    dfma = d.get('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 0.0)
    dmul = d.get('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 0.0)
    dadd = d.get('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 0.0)
    dmma = d.get('smsp__inst_executed_pipe_tensor_op_dmma.sum', d.get('sm__inst_executed_pipe_tensor_op_dmma.sum', 0.0))
    flops_scalar = 2 * dfma + dmul + dadd
    flops_dmma = 512.0 * dmma
    k = {'launches': len(launch_ids[name]), 'ms': d.get('gpu__time_duration.sum', 0.0),
         'dram_read_bytes': d.get('dram__bytes_read.sum', 0.0), 'dram_write_bytes': d.get('dram__bytes_write.sum', 0.0),
         'fp64_scalar_flops': flops_scalar, 'fp64_dmma_flops': flops_dmma}
    out['kernels'][name] = k
    for key in ('ms', 'dram_read_bytes', 'dram_write_bytes', 'fp64_scalar_flops', 'fp64_dmma_flops'):
        tot[key] += k[key]
    tot['launches'] += k['launches']
out['total'] = dict(tot)
out['total']['dram_bytes'] = tot['dram_read_bytes'] + tot['dram_write_bytes']
out['total']['fp64_flops'] = tot['fp64_scalar_flops'] + tot['fp64_dmma_flops']
out['kernels'] = dict(sorted(out['kernels'].items(), key=lambda kv: -kv[1]['ms']))
print(json.dumps(out, indent=1))
if '--traffic' in sys.argv:
    tpath = sys.argv[sys.argv.index('--traffic') + 1]
    sl = {n: k for n, k in out['kernels'].items() if 'dgemm_sl_kernel<13, 12, 0' in n or 'dgemm_sl_kernel<13,12,0' in n}
    n, k = next(iter(sl.items())) if sl else (None, None)
    traffic = {'dram_bytes_per_step': out['total']['dram_bytes'],
               'fp64_flops_executed_per_step': out['total']['fp64_flops'],
               'fp64_dmma_flops_executed_per_step': tot['fp64_dmma_flops'],
               'fp64_scalar_flops_executed_per_step': tot['fp64_scalar_flops'],
               'kernel_ms_sum': tot['ms'], 'source': path.replace('gpurun_out/', 'profiles/(from) gpurun_out/'),
               'note': 'ncu metrics pass over one evaluation at the bench shape (N = 1e5, M = 200, exact-zero windows): '
                       'dram__bytes_read.sum + dram__bytes_write.sum summed over every launch of the step; per launch of '
                       'the dominant kernel (dgemm_sl_kernel<13,12,0>, T1 = H A) in dgemm_sl_kernel_T1_bytes_per_launch'}
    if k:
        traffic['dgemm_sl_kernel_T1_bytes_per_launch'] = (k['dram_read_bytes'] + k['dram_write_bytes']) / k['launches']
    with open(tpath, 'w') as f:
        json.dump(traffic, f, indent=1)
