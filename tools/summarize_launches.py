"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = r['Kernel Name']
    name = re.sub(r'\(.*', '', name)
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print('kernel,launches,total_ms,share')
for k in sorted(tot, key=lambda k: -tot[k]):
    print('%s,%d,%.3f,%.4f' % (k, cnt[k], tot[k], tot[k] / total))
print('TOTAL,%d,%.3f,1.0' % (sum(cnt.values()), total))
