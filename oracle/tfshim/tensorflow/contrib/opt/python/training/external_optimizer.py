"""tf.contrib.opt.ScipyOptimizerInterface stand-in: SciPy's minimiser driven by shim gradients, with the calling
convention the reference uses (src/core/learn.py:119-133).  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import scipy.optimize

import tensorflow as tf


class ScipyOptimizerInterface(object):
    def __init__(self, loss, var_list=None, method='L-BFGS-B', options=None, **kw_args):
        self._loss = loss
        self._vars = list(var_list)
        self._method = method
        self._options = dict(options or {})
        self._grads = tf.gradients(loss, self._vars)

    def _pack(self, arrays):
        return np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in arrays])

    def _assign(self, sess, x):
        k = 0
        for v in self._vars:
            shp = v.get_shape()
            n = int(np.prod(shp)) if len(shp) else 1
            sess.run(v.assign(x[k:k + n].reshape(shp)))
            k += n

    def minimize(self, session=None, feed_dict=None, fetches=None, step_callback=None, loss_callback=None, **kw):
        sess = session
        fetches = list(fetches or [])
        x0 = self._pack(sess.run(self._vars))

        def fun(x):
            self._assign(sess, x)
            out = sess.run([self._loss, self._grads] + fetches, feed_dict=feed_dict)
            if loss_callback is not None:
                loss_callback(*out[2:])
            return float(out[0]), self._pack(out[1])

        res = scipy.optimize.minimize(fun, x0, jac=True, method=self._method, callback=step_callback,
                                      options=self._options)
        self._assign(sess, res.x)
        return res
