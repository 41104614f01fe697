"""Subprocess front-end of oracle/ref_run.py: the reference's own code (oracle/_ref) evaluated on given inputs.
TEST INFRASTRUCTURE ONLY — only tests/, tools/ fixture generators and bench.py's reference arm call this."""
import os
import subprocess
import sys
import tempfile

import numpy as np

from . import build_ref

HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    """True when oracle/_ref exists or can be built here (needs /root/reference)."""
    return os.path.isdir(build_ref.DST) or os.path.isdir(build_ref.DEFAULT_SRC)


def ensure_built():
    if not os.path.isdir(build_ref.DST):
        build_ref.build(quiet=True)


def call(timeout=3600, threads=None, **inp):
    """Run the reference on `inp` (see oracle/ref_run.py) and return its outputs as a dict of numpy values."""
    ensure_built()
    env = dict(os.environ)
    if threads is not None:
        env['OMP_NUM_THREADS'] = env['MKL_NUM_THREADS'] = str(int(threads))
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, 'in.npz'), os.path.join(d, 'out.npz')
        np.savez(src, **{k: np.asarray(v) for k, v in inp.items()})
        p = subprocess.run([sys.executable, os.path.join(HERE, 'ref_run.py'), src, dst], env=env, timeout=timeout,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        if p.returncode != 0:
            raise RuntimeError('oracle/ref_run.py failed:\n' + p.stderr[-4000:])
        with np.load(dst, allow_pickle=False) as z:
            return {k: (z[k][()] if z[k].ndim == 0 else z[k]) for k in z.files}
