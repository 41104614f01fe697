"""The mean-field training schedule of the reference's ``experiment.train`` (``src/core/experiment.py:192-262``) on the
reference-facing API: precompute, L-BFGS on q(u), on q(u) + (s2_f, s2), undo-precompute and L-BFGS on everything,
precompute, fixed-point iterations.  Only the schedule lives here; every evaluation is one C-ABI call."""
import numpy as np

from . import config, learn
from .cgpcm import VCGPCM


def train(sess, e, nx, nh, tau_w, tau_f, causal=True, reg=None, noise_init=1e-4, iters_fpi_pre=0, iters_pre=0,
          iters=0, iters_post=0, iters_fpi_post=0, fix_alpha=False, quiet=True, tx_range=None):
    """Build a ``VCGPCM`` by recipe and run the three-phase schedule.  Returns ``(mod, report)`` with ``report`` the
    ELBO after every phase and the number of objective evaluations of every L-BFGS run."""
    if reg is not None:
        config.reg = reg
    mod = VCGPCM.from_recipe(sess=sess, e=e, nx=nx, nh=nh, tau_w=tau_w, tau_f=tau_f, causal=causal,
                             noise_init=noise_init, tx_range=tx_range)
    report = {'elbo': {}, 'evals': {}}
    mod.precompute()
    elbo, terms = mod.elbo()
    report['elbo']['start'] = sess.run(elbo)
    if iters_fpi_pre:
        mod.fpi(iters_fpi_pre)
        report['elbo']['fpi_pre'] = sess.run(elbo)
    V = mod.vars
    fetches = [{'name': 'ELBO', 'tensor': elbo, 'modifier': '.2e'}]

    def lbfgs(name, names, n_it, fetch):
        res = learn.minimise_lbfgs(sess, -fetch[0]['tensor'], vars=[V[k] for k in names], iters=n_it,
                                   fetches_config=fetch + terms, name=name, quiet=quiet)
        report['evals'][name] = int(res.nfev) if res is not None else 0
        report['elbo'][name] = sess.run(fetch[0]['tensor'])

    lbfgs('pretraining', ['mu_u', 'var_u'], iters_pre, fetches)
    lbfgs('training', ['mu_u', 'var_u', 's2_f', 's2'], iters, fetches)
    if iters_post > 0:
        mod.undo_precompute()
        elbo, terms = mod.elbo()
        fetches = [{'name': 'ELBO', 'tensor': elbo, 'modifier': '.2e'}]
        lbfgs('posttraining', ['mu_u', 'var_u', 's2_f', 's2', 'gamma', 'omega'] + ([] if fix_alpha else ['alpha']),
              iters_post, fetches)
        mod.precompute()
    if iters_fpi_post:
        elbo = mod.elbo()[0]
        mod.fpi(iters_fpi_post)
        report['elbo']['fpi_post'] = sess.run(elbo)
    report['elbo']['final'] = sess.run(mod.elbo()[0])
    return mod, report
