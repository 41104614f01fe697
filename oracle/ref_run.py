"""Runs the reference's OWN code (oracle/_ref: py3-patched copies of /root/reference/src, built by
oracle/build_ref.py, on the TensorFlow stand-in of oracle/tfshim) through its own public API, in a process of its
own.  TEST INFRASTRUCTURE ONLY.

    python oracle/ref_run.py in.npz out.npz

`in.npz`: t, y, nx, nh, tau_w, tau_f, causal, causal_id, reg, params  (parameter vector of SURVEY.md 8b) and
optionally params_frozen, fpi_num / fpi_high_reg, t_star + samples_h (+ smf), t_k, time (repetitions to time).
`out.npz`: what `VCGPCM.from_recipe(...)`, `mod.elbo()`, `tf.gradients`, `mod.precompute()`, `mod.fpi()`,
`mod.convert()`, `mod.predict_f()` of src/core/cgpcm.py return for those inputs.

The module is a script on purpose: the reference's modules have top-level names such as `config`, `util`, `data`
and a `tensorflow` stand-in must be importable, none of which may leak into the test or bench process.
"""
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import build_ref  # noqa: E402


def _setup():
    if not os.path.isdir(build_ref.DST):
        build_ref.build(quiet=True)
    paths, env = build_ref.paths()
    sys.path[:0] = paths
    os.environ.update(env)
    warnings.filterwarnings('ignore', category=SyntaxWarning)


def run(inp):
    _setup()
    import tensorflow as tf
    import config
    config.reg = float(inp['reg'])                       # tasks set config.reg before building the model
    import cgpcm as ref_cgpcm
    import data as ref_data
    import out as ref_out
    ref_out._out = lambda msg: None                     # the progress display writes escape codes to stdout

    t = np.asarray(inp['t'], dtype=np.float64)
    y = np.asarray(inp['y'], dtype=np.float64)
    nh, nx = int(inp['nh']), int(inp['nx'])
    causal, causal_id = bool(inp['causal']), bool(inp.get('causal_id', False))
    params = np.asarray(inp['params'], dtype=np.float64)
    if 'seed' in inp:
        np.random.seed(int(inp['seed']))

    sess = tf.Session()
    e = ref_data.Data(t, y)
    tx_range = tuple(float(v) for v in inp['tx_range']) if 'tx_range' in inp else None
    mod = ref_cgpcm.VCGPCM.from_recipe(sess=sess, e=e, nx=nx, nh=nh, tau_w=float(inp['tau_w']),
                                       tau_f=float(inp['tau_f']), causal=causal, causal_id=causal_id,
                                       tx_range=tx_range)
    nh = int(mod.nh)                                     # acausal: from_recipe makes nh odd
    res = {'th': sess.run(mod.th), 'tx': sess.run(mod.tx), 'nh': nh,
           'recipe_vars': np.array([float(sess.run(mod.vars[k])) for k in ('s2', 's2_f', 'alpha', 'gamma', 'omega')])}

    names = ['s2', 's2_f', 'alpha', 'gamma', 'omega', 'mu_u', 'var_u']
    var_list = [mod.vars[k] for k in names]

    def assign(p):
        sess.run([mod.vars[k].assign(p[i]) for i, k in enumerate(names[:5])]
                 + [mod.vars['mu_u'].assign(p[5:5 + nh].reshape(nh, 1)), mod.vars['var_u'].assign(p[5 + nh:])])

    def pack(gs):
        return np.concatenate([np.asarray(g, dtype=np.float64).ravel() for g in gs])

    assign(params)
    elbo, terms = mod.elbo()
    grads = tf.gradients(elbo, var_list)
    fetch = [elbo, [tm['tensor'] for tm in terms], grads]
    t0 = time.time()
    ev, tv, gv = sess.run(fetch)
    res['seconds_first'] = time.time() - t0
    res.update(elbo=ev, terms=np.array(tv, dtype=np.float64), grad=pack(gv),
               term_names=np.array([tm['name'] for tm in terms]))
    reps = int(inp.get('time', 0))
    if reps > 0:
        ts = []
        for _ in range(reps):
            t0 = time.time()
            sess.run(fetch)
            ts.append(time.time() - t0)
        res['seconds'] = np.array(ts)

    if inp.get('want_mats', True):
        keys = ['a', 'Ahh', 'sum_Axx', 'sum_Ahx_y', 'sum_b', 'sum_Bxx', 'sum_Bhh', 'sum_a', 'sum_Ahh']
        if inp.get('want_per_n', True):
            keys += ['Ahx', 'Axx']
        vals = sess.run([mod.mats[k] for k in keys] + [mod.Kh, mod.Kx, mod.iKh, mod.iKx, mod.h.m2])
        for k, v in zip(keys + ['Kh', 'Kx', 'iKh', 'iKx', 'm2_u'], vals):
            res['mat_' + k] = v
        lam, P = sess.run(list(mod._optimal_q(mod.h.mean, mod.h.m2, z=True)))
        res['optq_lam'], res['optq_P'] = lam, P

    if 'params_frozen' in inp:
        # precomputed regime (src/core/cgpcm.py:270-292): `mats` become numpy constants, then new variable values
        mod.precompute()
        assign(np.asarray(inp['params_frozen'], dtype=np.float64))
        elbo_f, terms_f = mod.elbo()
        grads_f = tf.gradients(elbo_f, var_list)
        ev, tv, gv = sess.run([elbo_f, [tm['tensor'] for tm in terms_f], grads_f])
        res.update(elbo_frozen=ev, terms_frozen=np.array(tv, dtype=np.float64), grad_frozen=pack(gv))
        fpi_ok = True
        if 'fpi_num' in inp:
            try:
                mod.fpi(num=int(inp['fpi_num']), z=True, high_reg=bool(inp.get('fpi_high_reg', False)))
                mod.convert(z=True)
                for k in ('mu_u', 'var_u', 'mu_z', 'var_z'):
                    res['fpi_' + k] = np.asarray(sess.run(mod.vars[k])).ravel()
                res['fpi_elbo'] = sess.run(mod.elbo()[0])
            except tf.InvalidArgumentError as exc:
                # the reference's own iteration can leave the positive-definite cone (seen with causal_id=True, where
                # sum_Bhh is not positive semi-definite): recorded, and the dependent outputs are left out
                fpi_ok = False
                res['fpi_error'] = str(exc)
        if 'qz_mu' in inp:
            # the z = False variants from a given q(z): elbo(z=False), then fpi(num, z=False) + convert(z=False)
            sess.run([mod.vars['mu_z'].assign(np.asarray(inp['qz_mu'], dtype=np.float64).reshape(-1, 1)),
                      mod.vars['var_z'].assign(np.asarray(inp['qz_var'], dtype=np.float64))])
            elbo_z, terms_z = mod.elbo(z=False)
            try:
                ev, tv = sess.run([elbo_z, [tm['tensor'] for tm in terms_z]])
                res.update(qz_elbo=ev, qz_terms=np.array(tv, dtype=np.float64))
            except tf.InvalidArgumentError as exc:
                res['qz_error'] = str(exc)
            snapshot = {k: np.asarray(sess.run(mod.vars[k])).copy() for k in ('mu_u', 'var_u')}
            try:
                mod.fpi(num=int(inp.get('qz_fpi_num', 2)), z=False)
                mod.convert(z=False)
                for k in ('mu_u', 'var_u', 'mu_z', 'var_z'):
                    res['qz_fpi_' + k] = np.asarray(sess.run(mod.vars[k])).ravel()
            except tf.InvalidArgumentError as exc:
                res['qz_fpi_error'] = str(exc)
            sess.run([mod.vars[k].assign(v) for k, v in snapshot.items()])
        if 't_star' in inp and fpi_ok:
            if not bool(inp.get('smf', False)):
                # the reference takes the non-SMF branch only when it draws the filter samples itself
                # (`is_numeric(samples_h)`, cgpcm.py:800-805): replay its draws (chol(var) eps + mean, one
                # standard-normal vector per sample from the stand-in's numpy generator) and hand them back
                seed, B = int(inp.get('seed', 0)), int(inp['num_samples'])
                Lq, mq = sess.run([tf.cholesky(mod.h.var), mod.h.mean])
                np.random.seed(seed)
                res['pred_samples'] = np.stack([(Lq @ np.random.standard_normal([nh, 1]) + mq).ravel()
                                                for _ in range(B)])
                np.random.seed(seed)
                pred = mod.predict_f(np.asarray(inp['t_star'], dtype=np.float64), samples_h=B)
            else:
                samples = [s.reshape(nh, 1) for s in np.asarray(inp['samples_h'], dtype=np.float64)]
                pred = mod.predict_f(np.asarray(inp['t_star'], dtype=np.float64), samples_h=samples)
            res['pred_mean'], res['pred_std'] = pred.mean.y, pred.std.y
        mod.undo_precompute()
    return res


if __name__ == '__main__':
    src, dst = sys.argv[1], sys.argv[2]
    with np.load(src, allow_pickle=False) as z:
        inp = {k: z[k] for k in z.files}
    for k in list(inp):
        if inp[k].ndim == 0:
            inp[k] = inp[k][()]
    out = run(inp)
    np.savez_compressed(dst, **out)
