"""Build `oracle/_ref/`: the reference's OWN sources for the hot path, made runnable in this image.
TEST INFRASTRUCTURE ONLY.

The reference (wesselb/cgpcm) is Python 2 + TensorFlow 1.x + an external `bvn-cdf` build; none of them exists here
and `lib2to3` is not in the image.  This recipe

  1. copies the files of the path from `/root/reference/src` into `oracle/_ref/` (git-ignored: the reference's
     sources never enter the history; they do travel to the GPU box with the snapshot),
  2. applies the Python-3 fixes below -- every one is an exact, counted textual substitution of a Python-2 idiom
     (`reduce`, `zip(...)[0]`, `imp.load_source`, list-valued `range` / `map`, the `print` statement, a tuple
     parameter, `__div__`, `/` on integer indices); no arithmetic is touched,
  3. leaves TensorFlow and `bvn_cdf` to `oracle/tfshim/` (a deferred-graph stand-in on torch-CPU float64, and the
     oracle's Genz BVND where `$BVN_CDF_REPO/bvn_cdf.py` is expected).

`oracle/ref_run.py` then drives `VCGPCM` through the reference's own API.  Run:  python oracle/build_ref.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SRC = '/root/reference/src'
DST = os.path.join(HERE, '_ref')

FILES = ['config.py'] + ['core/' + f for f in (
    '__init__.py', 'tf_util.py', 'exponentiated_quadratic.py', 'exponentiated_quadratic_test.py', 'cgpcm.py',
    'kernel.py', 'distribution.py', 'parametrisable.py', 'util.py', 'learn.py', 'sample.py', 'out.py', 'data.py')]

FUNCTOOLS = 'from functools import reduce\n'

# (file, old, new, expected count)
PATCHES = [
    # tf_util.py: `imp` is gone from Python 3.12; range objects do not support item assignment
    ('core/tf_util.py', 'import imp\n', 'import importlib.util\n', 1),
    ('core/tf_util.py',
     "bvn_cdf = imp.load_source('bvn_cdf',\n"
     "                          os.path.join(os.environ['BVN_CDF_REPO'],\n"
     "                                       'bvn_cdf.py')).bvn_cdf\n",
     "_spec = importlib.util.spec_from_file_location(\n"
     "    'bvn_cdf', os.path.join(os.environ['BVN_CDF_REPO'], 'bvn_cdf.py'))\n"
     "_mod = importlib.util.module_from_spec(_spec)\n"
     "_spec.loader.exec_module(_mod)\n"
     "bvn_cdf = _mod.bvn_cdf\n", 1),
    ('core/tf_util.py', 'perm = range(len(shape(x)))', 'perm = list(range(len(shape(x))))', 1),
    ('core/tf_util.py', 'indices=zip(*np.tril_indices(m))', 'indices=list(zip(*np.tril_indices(m)))', 1),
    ('core/tf_util.py', 'tf.gather_nd(x, zip(*np.tril_indices(n)))', 'tf.gather_nd(x, list(zip(*np.tril_indices(n))))', 1),
    # exponentiated_quadratic.py
    ('core/exponentiated_quadratic.py', 'import operator\n', 'import operator\n' + FUNCTOOLS, 1),
    ('core/exponentiated_quadratic.py', '*zip(*vars_and_lims)[0]', '*list(zip(*vars_and_lims))[0]', 1),
    # cgpcm.py
    ('core/cgpcm.py', 'from operator import add\n', 'from operator import add\n' + FUNCTOOLS, 1),
    # learn.py: map() is lazy in Python 3, the callers concatenate / index the result
    ('core/learn.py', 'return map(mapping_fun, xs)', 'return list(map(mapping_fun, xs))', 1),
    # sample.py: tuple parameter, print statement
    ('core/sample.py', 'def _draw(self, (theta_l, theta_u), u, attempts=1):',
     'def _draw(self, theta_lu, u, attempts=1):', 1),
    ('core/sample.py', '        self._draw_proposal(theta_l, theta_u)\n',
     '        theta_l, theta_u = theta_lu\n        self._draw_proposal(theta_l, theta_u)\n', 1),
    ('core/sample.py', "print 'warning: theta violation'", "print('warning: theta violation')", 1),
    # util.py
    ('core/util.py', 'inverse_perm = range(n)', 'inverse_perm = list(range(n))', 1),
    ('core/util.py', 'return map(lambda x: colorsys.hsv_to_rgb(*x), hsvs)',
     'return list(map(lambda x: colorsys.hsv_to_rgb(*x), hsvs))', 1),
    # data.py: Python-2 division protocol, integer indices
    ('core/data.py', 'import operator\n', 'import operator\n' + FUNCTOOLS, 1),
    ('core/data.py', '    def __mul__(self, other):\n        return Data(self.x, self.y * self._to_y(other))\n',
     '    __truediv__ = __div__\n    __rtruediv__ = __rdiv__\n\n'
     '    def __mul__(self, other):\n        return Data(self.x, self.y * self._to_y(other))\n', 1),
    ('core/data.py', 'x -= x[self.n / 2]', 'x -= x[self.n // 2]', 1),
    ('core/data.py', '(factor - 1) * (self.n / 2)', '(factor - 1) * (self.n // 2)', 1),
    ('core/data.py', 'inds = [range(s, s + l) for s, l in zip(start, length)]',
     'inds = [list(range(s, s + l)) for s, l in zip(start, length)]', 1),
]


def build(src=DEFAULT_SRC, dst=DST, quiet=False):
    if not os.path.isdir(src):
        raise FileNotFoundError('reference sources not found at {} (oracle/_ref can only be built where '
                                '/root/reference exists; the GPU box uses the prebuilt copy)'.format(src))
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(os.path.join(dst, 'core'))
    texts = {}
    for f in FILES:
        with open(os.path.join(src, f)) as fh:
            texts[f] = fh.read()
    for f, old, new, count in PATCHES:
        have = texts[f].count(old)
        if have != count:
            raise RuntimeError('patch does not apply to {}: {!r} found {} times, expected {}'.format(f, old, have, count))
        texts[f] = texts[f].replace(old, new)
    for f, s in texts.items():
        with open(os.path.join(dst, f), 'w') as fh:
            fh.write(s)
    with open(os.path.join(dst, 'README'), 'w') as fh:
        fh.write('Generated by oracle/build_ref.py from {} (py3-patched copies of the reference; do not commit).\n'.format(src))
    if not quiet:
        print('oracle/_ref: {} files, {} substitutions'.format(len(FILES), len(PATCHES)))


def paths():
    """sys.path entries + environment a process needs to import the built reference (see oracle/ref_run.py)."""
    shim = os.path.join(HERE, 'tfshim')
    return [shim, DST, os.path.join(DST, 'core')], {'BVN_CDF_REPO': os.path.join(shim, 'bvn_cdf_repo')}


if __name__ == '__main__':
    build(*(sys.argv[1:2] or [DEFAULT_SRC]))
