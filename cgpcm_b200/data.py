"""Minimal ``Data`` container: what the hot path reads of the reference's ``core.data.Data``
(``src/core/data.py:25-34``): observation inputs ``x`` and outputs ``y`` as float64 vectors."""
from collections import namedtuple

import numpy as np


class Data(object):
    def __init__(self, x, y=None):
        self.x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
        if y is None:
            y = np.zeros_like(self.x)
        self.y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
        if self.x.shape != self.y.shape:
            raise ValueError('x and y must have the same length')

    @property
    def n(self):
        return self.x.shape[0]

    def __len__(self):
        return self.n

    def __getitem__(self, sl):
        return Data(self.x[sl], self.y[sl])

    # -- the part of the reference's Data that ``load_akm`` uses (``src/core/data.py:109-125,354-387,430-446``)
    def _to_y(self, other):
        return other.y if isinstance(other, Data) else other

    def __sub__(self, other):
        return Data(self.x, self.y - self._to_y(other))

    def __mul__(self, other):
        return Data(self.x, self.y * self._to_y(other))

    __rmul__ = __mul__

    def __truediv__(self, other):
        return Data(self.x, self.y / self._to_y(other))

    __div__ = __truediv__

    def make_noisy(self, var):
        """Add white noise of variance ``var``."""
        return Data(self.x, self.y + var ** .5 * np.random.randn(*self.y.shape))

    def positive_part(self):
        """The part of the data at non-negative inputs."""
        keep = self.x >= 0
        return Data(self.x[keep], self.y[keep])

    @property
    def mean(self):
        return np.nanmean(self.y)

    @property
    def std(self):
        return np.nanstd(self.y, ddof=1)

    @property
    def max(self):
        return self.y.max()

    @property
    def energy(self):
        trap = getattr(np, 'trapezoid', None) or np.trapz
        return trap(self.y ** 2, self.x)


# Named tuple for bundling predictions (``src/core/data.py:481-482``)
UncertainData = namedtuple('UncertainData', 'mean lower upper std')


def load_akm(sess, causal, n=250, nh=31, tau_w=.1, tau_f=.05, resample=0):
    """Sample from the AKM (``src/core/data.py:594-641``): the series, kernel and filter of the toy experiment
    (``src/tasks/toy.py:57-66``), normalised like the reference (zero mean / unit std, unit energy, unit maximum).

    :return: data for function, kernel, and filter
    """
    from .cgpcm import AKM
    k_stretch = 8
    e = Data(np.linspace(0, 1, n), None)
    akm = AKM.from_recipe(sess=sess, e=e, nx=0, nh=nh, tau_w=tau_w, tau_f=tau_f, causal=causal)
    akm.sample(e.x)
    for _ in range(resample):
        akm.sample_f(e.x)
    f = akm.f()
    tk = np.linspace(-k_stretch * tau_w, k_stretch * tau_w, 301)
    k = akm.k(tk)
    h = akm.h(tk)
    if causal:
        h = h.positive_part()
    f = f - f.mean
    f = f / f.std
    h = h / h.energy ** .5
    k = k / k.max
    return f, k, h
