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


# Named tuple for bundling predictions (``src/core/data.py:481-482``)
UncertainData = namedtuple('UncertainData', 'mean lower upper std')
