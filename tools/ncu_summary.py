"""Text summary of an ncu report (`ncu -i X.ncu-rep --page raw --csv`): per launch, the metrics the roofline
discussion in DESIGN.md cites.  Usage: python tools/ncu_summary.py report.ncu-rep | raw_page.csv > profiles/name.txt"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'smsp__pipe_tensor_subpipe_dmma_cycles_active.avg', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
if sys.argv[1].endswith('.csv'):      # an exported raw page (tools/gpu_final.sh keeps the CSV pages, not the reports)
    out = open(sys.argv[1]).read()
else:
    out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
print('# ncu --set full --clock-control none, report %s' % sys.argv[1].split('/')[-1])
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('\n== %s  grid %s block %s' % (d.get('Kernel Name'), d.get('Grid Size'), d.get('Block Size')))
    for i, h in enumerate(hdr):
        short = h.split('TriageCompute.')[-1]
        if short in WANT or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            v = r[i]
            if 'stalled' in h:
                try:
                    if float(v.replace(',', '')) < 0.2:
                        continue
                except ValueError:
                    pass
            print('  %-95s %s %s' % (short, v, units[i]))
