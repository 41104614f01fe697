"""Coefficients of csrc/cgmath.cuh: Chebyshev interpolants of exp on |r| <= ln2/2 (degree 12) and of
g(t) = (1 + 2x) erfcx(x), x = 4 (1 + t) / (1 - t), on [-1, 1] (degree 26), computed with mpmath at 60 digits, converted to
the monomial basis exactly and rounded to double (printed as hex floats)."""
import mpmath as mp, numpy as np
mp.mp.dps = 60
def cheb_fit(f, deg):
    # interpolate f on [-1,1] at Chebyshev nodes, return monomial coefficients (mp)
    N = deg + 1
    nodes = [mp.cos(mp.pi * (2*k + 1) / (2*N)) for k in range(N)]
    fv = [f(x) for x in nodes]
    # chebyshev coefficients
    a = []
    for j in range(N):
        s = mp.mpf(0)
        for k in range(N):
            s += fv[k] * mp.cos(mp.pi * j * (2*k + 1) / (2*N))
        a.append(2 * s / N)
    a[0] /= 2
    # convert to monomial: T_0=1, T_1=x, T_{n+1}=2xT_n - T_{n-1}
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for n in range(1, deg):
        t = [mp.mpf(0)] + [2*c for c in T[n]]
        for i, c in enumerate(T[n-1]): t[i] -= c
        T.append(t)
    mono = [mp.mpf(0)] * N
    for j in range(N):
        for i, c in enumerate(T[j]): mono[i] += a[j] * c
    return mono, a
# ---- erfc: g(t) = (1 + 2x) erfcx(x), x = 4 (1 + t) / (1 - t)
def g(t):
    if t >= 1: return 2 / mp.sqrt(mp.pi)
    x = 4 * (1 + t) / (1 - t)
    return (1 + 2*x) * mp.exp(x*x) * mp.erfc(x)
for deg in (20, 22, 24, 26):
    mono, a = cheb_fit(g, deg)
    print(deg, 'last cheb coefs', [mp.nstr(abs(c), 3) for c in a[-3:]], 'sum|c|', mp.nstr(sum(abs(c) for c in mono), 5))

print('---- exp')
h = mp.log(2) / 2
for deg in (11, 12, 13):
    mono_s, a = cheb_fit(lambda s: mp.exp(s * h), deg)
    mono = [c / h**i for i, c in enumerate(mono_s)]     # in r
    print(deg, 'last cheb', [mp.nstr(abs(c), 3) for c in a[-2:]])
# ---- emit tables
def emit(name, coefs):
    print('static const double %s[%d] = {' % (name, len(coefs)))
    for c in coefs:
        print('    %s,' % float(c).hex())
    print('};')
mono_g, _ = cheb_fit(g, 26)
mono_s, _ = cheb_fit(lambda s: mp.exp(s * h), 12)
mono_e = [c / h**i for i, c in enumerate(mono_s)]
emit('ERFC_G', mono_g)
emit('EXP_P', mono_e)
import pickle
pass
print('ln2_hi/lo', float(mp.log(2)).hex())
