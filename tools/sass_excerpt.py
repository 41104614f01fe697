"""SASS evidence for the hot kernels: per kernel of libcgpcm_b200.so the instruction count and the counts of the
mnemonics that identify the data path -- DMMA.8x8x4 (FP64 tensor), UBLKCP (1-D bulk copy, cp.async.bulk), SYNCS
(mbarrier arrive / try_wait), LDGSTS (cp.async), DFMA / DMUL / DADD -- plus the first lines around the first DMMA /
UBLKCP of the three contraction kernels.   python tools/sass_excerpt.py > profiles/r02_sass_hot_kernels.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'cgpcm_b200', 'lib', 'libcgpcm_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(['c++filt', s], capture_output=True, text=True).stdout.strip()
funcs = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.search(r'/\*[0-9a-f]{4}\*/', line):
        funcs[cur].append(line)
HOT = ['dgemm_sl_kernel', 'dgemm_sym16_kernel', 'dgemm_sym_kernel', 'dgemm_dmma_kernel', 'axx_sum_kernel',
       'ahx_gen_sep_kernel', 'ahx_dot_sep_kernel', 'ahx_gen_kernel', 'ahx_dot_kernel', 'potrf_panel_kernel']
KEYS = ['DMMA', 'UBLKCP', 'SYNCS', 'LDGSTS', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'BAR.SYNC']
print('# cuobjdump -sass cgpcm_b200/lib/libcgpcm_b200.so (sm_100a), mnemonic counts per kernel instantiation')
print('# %-64s %7s ' % ('kernel', 'instrs') + ' '.join('%8s' % k for k in KEYS))
shown = set()
for name, lines in funcs.items():
    d = demangle(name)
    if not any(h in d for h in HOT):
        continue
    short = re.sub(r'\(.*', '', d).replace('void ', '').replace('cg::', '')
    ops = Counter()
    for l in lines:
        m = re.search(r'\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    ops[k] += 1
    print('  %-64s %7d ' % (short[:64], len(lines)) + ' '.join('%8d' % ops[k] for k in KEYS))
for want, key in (('dgemm_sl_kernel<13, 12, false>', 'UBLKCP'), ('dgemm_sl_kernel<13, 12, false>', 'DMMA'),
                  ('dgemm_sym_kernel<true, 1, 5, 5>', 'DMMA')):
    for name, lines in funcs.items():
        d = demangle(name)
        if want in d and (want, key) not in shown:
            shown.add((want, key))
            idx = next((i for i, l in enumerate(lines) if key in l), None)
            if idx is None:
                continue
            print('\n# %s : first %s and its neighbourhood' % (want, key))
            for l in lines[max(0, idx - 6):idx + 10]:
                print('   ' + re.sub(r'\s+/\* 0x[0-9a-f]+ \*/\s*$', '', l.rstrip()))
