"""The few helpers of ``src/core/util.py`` that the hot path uses."""
import numpy as np

inf = np.inf


def length_scale(ls):
    """Constant of an EQ kernel ``exp(-c r^2)`` with length scale ``ls`` (``src/core/util.py:29-36``)."""
    return (.5 * np.pi) * (.5 / ls ** 2)


def to_float(v):
    """``tf.cast(tf.to_float(v), float64)`` (``src/core/tf_util.py:108-115``): a float32 round trip."""
    return float(np.float32(v))


def is_inf(x):
    return np.isinf(x)


def tril_to_vec(x):
    """``src/core/tf_util.py:436-447``: row-major lower triangle (``np.tril_indices`` order)."""
    x = np.asarray(x)
    return x[np.tril_indices(x.shape[-1])]


def vec_to_tril(v):
    """``src/core/tf_util.py:419-433``."""
    v = np.asarray(v)
    m = int(((1 + 8 * v.shape[0]) ** .5 - 1) / 2)
    out = np.zeros((m, m), dtype=v.dtype)
    out[np.tril_indices(m)] = v
    return out
