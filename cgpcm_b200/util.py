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


# Percentiles of the +-2 sigma band used for Monte-Carlo predictions (``src/core/util.py:11-12``)
lower_perc = 2.275013194817921    # 100 * Phi(-2)
upper_perc = 97.72498680518208    # 100 * Phi(2)


def fft_spectrum(t, y, zero_pad=2000):
    """``Data(t, y).fft()`` of the reference (``src/core/data.py:184-209``, ``util.py:67-93,104-121``): zero pad by
    ``zero_pad`` on both sides, FFT along axis 0, ``fftshift``, scale by the spacing.  Returns ``(freq, spectrum)``."""
    t = np.asarray(t, dtype=np.float64)
    d = np.diff(t)
    if t.shape[0] < 2 or np.abs(d - d[0]).max() > 1e-8 * max(abs(d[0]), 1e-300):
        raise AssertionError('data must be evenly spaced')
    dx = t[1] - t[0]
    y = np.asarray(y)
    pad = np.zeros((zero_pad,) + y.shape[1:], dtype=y.dtype)
    spec = dx * np.fft.fftshift(np.fft.fft(np.concatenate((pad, y, pad), axis=0), axis=0), axes=0)
    return np.fft.fftshift(np.fft.fftfreq(spec.shape[0])) / dx, spec


def minimum_phase(y):
    """``Data.minimum_phase`` (``src/core/data.py:293-303``): the minimum-phase signal with the magnitude spectrum
    of ``y`` (homomorphic method: Hilbert transform of the log magnitude)."""
    from scipy import signal
    mag = np.abs(np.fft.fft(y))
    spec = np.exp(signal.hilbert(np.log(mag)).conj())
    return np.real(np.fft.ifft(spec))


def zero_phase(x, y):
    """``Data.zero_phase`` (``src/core/data.py:265-276``): the zero-phase signal with the magnitude spectrum of ``y``,
    centred; returns ``(x', y')`` with ``x' = dx (arange(n) - n // 2)``."""
    x = np.asarray(x, dtype=np.float64)
    d = np.diff(x)
    if x.shape[0] < 2 or np.abs(d - d[0]).max() > 1e-8 * max(abs(d[0]), 1e-300):
        raise AssertionError('data must be evenly spaced')
    n = x.shape[0]
    yz = np.real(np.fft.fftshift(np.fft.ifft(np.abs(np.fft.fft(y)))))
    xz = d[0] * np.arange(n)
    return xz - xz[n // 2], yz


def energy(x, y):
    """``Data.energy`` (``src/core/data.py:354-359``): trapezoidal integral of ``y^2``."""
    trap = getattr(np, 'trapezoid', None) or np.trapz
    return trap(np.asarray(y) ** 2, x)


def autocorrelation(x, y, normalise=False):
    """``Data.autocorrelation`` (``src/core/data.py:136-171``, biased estimate): ``(lags, ac)``."""
    y = np.where(np.isnan(y), 0.0, y)
    n = y.shape[0]
    ac = np.convolve(y[::-1], y) / n
    lag = x[-1] - x[0]
    if normalise:
        ac = ac / ac.max()
    return np.linspace(-lag, lag, 2 * n - 1), ac
