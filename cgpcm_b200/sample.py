"""Elliptical slice sampling (Murray, Adams & MacKay 2010) with the interface of the reference's ``ESS``
(``src/core/sample.py:8-130``): ``ESS(log_lik, sample_prior, x_init)``, ``move``, ``update``, ``sample(num)``.

``log_lik`` returns the log of a quantity proportional to the likelihood of the target; ``sample_prior`` draws from
the zero-mean Gaussian prior.  Every state is a point on the ellipse through the current state and a fresh prior
draw; the bracket of the angle shrinks towards 0 until the proposal lies above the slice."""
import numpy as np


class ESS(object):
    _min_theta = 1e-10

    def __init__(self, log_lik, sample_prior, x_init=None, rng=None):
        self._log_lik = log_lik
        self._sample_prior = sample_prior
        self._rng = rng if rng is not None else np.random
        self._x = sample_prior() if x_init is None else x_init
        self._log_lik_x = log_lik(self._x)
        self.attempts = []

    def move(self, x, log_lik=None):
        """Move the sampler to a new state."""
        self._x = x
        self._log_lik_x = self._log_lik(x) if log_lik is None else log_lik

    def update(self, log_lik):
        """Replace the log-likelihood function."""
        self._log_lik = log_lik

    def _step(self):
        u = self._log_lik_x - self._rng.exponential(1.0)          # slice height
        nu = self._sample_prior()                                  # the ellipse
        theta = self._rng.uniform(0, 2 * np.pi)
        lo, hi = theta - 2 * np.pi, theta
        attempts = 0
        while True:
            attempts += 1
            theta = self._rng.uniform(lo, hi)
            x = np.cos(theta) * self._x + np.sin(theta) * nu
            ll = self._log_lik(x)
            if ll > u or abs(theta) < self._min_theta:
                self._x, self._log_lik_x = x, ll
                return attempts
            if theta > 0:
                hi = theta
            else:
                lo = theta

    def sample(self, num=1):
        """``num`` successive states (a list; a single state if ``num == 1``)."""
        out = []
        for _ in range(num):
            self.attempts.append(self._step())
            out.append(self._x)
        return out if len(out) > 1 else out[0]
