"""ELBO of the scaling-sweep workload at a small shape under the `tri` and `cull` options (full precision, for the
comparison with the binary128 arbiter): python tools/tri_probe.py N M"""
import sys
sys.path.insert(0, '.')
import numpy as np
import cgpcm_b200
from tests.workload import sweep_workload
n, m = int(sys.argv[1]), int(sys.argv[2])
wl = sweep_workload(n, m)
for cull in (80.0, 746.0, 0.0):
    for tri in (0, 1):
        for lo, hi in ((0, n),):
            eng = cgpcm_b200.Engine(m, m)
            eng.set_option('cull', cull)
            eng.set_option('tri', tri)
            eng.set_data(wl['t'][lo:hi], wl['y'][lo:hi], wl['th'], wl['tx'])
            e, terms, g = eng.elbo_grad(wl['params'], reg=wl['reg'])
            print('cull %g tri %d obs [%d,%d) elbo %.17e terms %s gmax %.6e' % (cull, tri, lo, hi, e, ['%.12e' % x for x in terms], np.abs(g).max()), flush=True)
            eng.close()
