"""``learn.minimise_lbfgs`` with the reference's signature (``src/core/learn.py:102-133``).

The reference wraps ``tf.contrib.opt.ScipyOptimizerInterface``: variables are packed into one flat
float64 vector and SciPy's L-BFGS-B calls a closure that runs ``sess.run([loss, packed_grad] +
fetches)``.  Here the closure is one ``cgpcm_elbo_grad`` call through the C-ABI.
"""
import sys
import time

import numpy as np
from scipy.optimize import minimize


class Progress(object):
    """Console progress display (``src/core/learn.py:10-99``), plain line output."""

    def __init__(self, name, iters=None, fetches_config=None, stream=None, quiet=False):
        self._started = False
        self._iters = iters
        self._name = name
        self._fetches_config = fetches_config or []
        self._fetches_cache = [None] * len(self._fetches_config)
        self._stream = stream or sys.stdout
        self._quiet = quiet

    def __call__(self, fetches=None, step=True):
        if not self._started:
            self._started = True
            self._start_time = time.time()
            self._iter = 1
        elif step:
            self._iter += 1
        if fetches is None:
            fetches = self._fetches_cache
        else:
            self._fetches_cache = fetches
        if self._quiet or not step:
            return
        status = '{}/{}'.format(self._iter, self._iters) if self._iters is not None else str(self._iter)
        parts = ['%s: iteration %s' % (self._name, status), 'elapsed %.1fs' % (time.time() - self._start_time)]
        for conf, fetch in zip(self._fetches_config, fetches):
            if conf is not None and fetch is not None:
                parts.append(('{}={:' + conf.get('modifier', '') + '}{}').format(conf['name'], fetch,
                                                                                 conf.get('unit', '')))
        self._stream.write(', '.join(parts) + '\n')


def minimise_lbfgs(sess, objective, vars, iters, fetches_config=None, name='minimisation using L-BFGS',
                   quiet=False):
    """Minimise some objective using SciPy's L-BFGS-B.

    :param sess: session
    :param objective: objective (``-elbo``)
    :param vars: list of variables to optimise
    :param iters: number of iterations
    :param fetches_config: fetches as specified in ``Progress`` with an additional key ``tensor``
    :param name: name of minimisation
    :return: the ``scipy.optimize.OptimizeResult`` (the reference returns ``None``)
    """
    if iters == 0:
        return None
    if fetches_config is None:
        fetches_config = []
    progress = Progress(name=name, iters=iters, fetches_config=fetches_config, quiet=quiet)
    shapes = [v.value.shape for v in vars]
    sizes = [v.value.size for v in vars]

    def unpack(x):
        off = 0
        for v, shp, sz in zip(vars, shapes, sizes):
            v.value = np.array(x[off:off + sz], dtype=np.float64).reshape(shp)
            off += sz

    def fun(x):
        unpack(x)
        f, g = objective.value_and_grad(vars)
        progress(sess.run([c['tensor'] for c in fetches_config]), step=False)
        return f, g

    x0 = np.concatenate([v.value.ravel() for v in vars])
    # SciPy's fmin_l_bfgs_b tends to perform two extra iterations (src/core/learn.py:121-125)
    res = minimize(fun, x0, jac=True, method='L-BFGS-B', callback=lambda x: progress(step=True),
                   options={'maxiter': max(iters - 2, 0)})
    unpack(res.x)
    return res


def map_progress(f, xs, name):
    progress = Progress(name=name, iters=len(xs), quiet=True)

    def mapping_fun(x):
        progress()
        return f(x)

    return list(map(mapping_fun, xs))
