// Bivariate normal CDF on the device (FP64, registers only) — the sm_100a replacement of the
// reference's external CPU custom op `bvn_cdf` (loaded at src/core/tf_util.py:9-13, called once at
// src/core/exponentiated_quadratic.py:552).  Algorithm: Genz (2004) BVND, the published routine that
// the reference's dependency wraps (doc/paper.tex:341); restated from the literature, not from code.
//
// Two entry points:
//   bvnd_tab(h, k, T)      correlation fixed per launch.  In the CGPCM rho = gamma/(alpha+gamma+omega)
//                          is one number for the whole model, so sin(), asin(),
//                          sqrt() and every reciprocal of the quadrature nodes are hoisted to the host
//                          (BvnTab travels as a kernel parameter = constant bank, uniform access);
//                          what remains per element is exp() and FMAs, and the Genz branch is uniform.
//   bvnd_general(h, k, r)  per-element correlation, for the stand-alone cgpcm_bvn_cdf export.
#pragma once
#include "cgmath.cuh"
#include <cuda_runtime.h>
#include <math.h>

namespace cg {

struct BvnTab {
  int ng;          // Gauss-Legendre order: 6 / 12 / 20
  int high;        // |rho| >= 0.925
  double rho;
  double asr_4pi;  // asin(rho) / (4 pi)                                  (low branch)
  double as_, a_;  // 1 - rho^2, sqrt(1 - rho^2)                          (high branch)
  double w[20];    // low: GL weight            | high: (a/2) * GL weight
  double c0[20];   // low: sin(asr (x+1)/2)     | high: 1 / xs ,  xs = (a/2 (x+1))^2
  double c1[20];   // low: 1 / (1 - sn^2)       | high: xs / (2 (1 + rs)^2),  rs = sqrt(1 - xs)
  double c2[20];   // low: unused               | high: 1 / rs
  double c3[20];   // low: unused               | high: xs
  // partials
  double inv_s;      // 1 / sqrt(1 - rho^2)
  double inv_2om;    // 1 / (2 (1 - rho^2))
  double inv_2pis;   // 1 / (2 pi sqrt(1 - rho^2))
};

__device__ __constant__ double c_gl6_x[6] = {
    -9.32469514203151939e-01, -6.61209386466264482e-01, -2.38619186083196905e-01,
    2.38619186083196905e-01,  6.61209386466264482e-01,  9.32469514203151939e-01};
__device__ __constant__ double c_gl6_w[6] = {
    1.71324492379170273e-01, 3.60761573048138717e-01, 4.67913934572691037e-01,
    4.67913934572691037e-01, 3.60761573048138717e-01, 1.71324492379170273e-01};
__device__ __constant__ double c_gl12_x[12] = {
    -9.81560634246719244e-01, -9.04117256370474798e-01, -7.69902674194304693e-01,
    -5.87317954286617483e-01, -3.67831498998180184e-01, -1.25233408511468913e-01,
    1.25233408511468913e-01,  3.67831498998180184e-01,  5.87317954286617483e-01,
    7.69902674194304693e-01,  9.04117256370474798e-01,  9.81560634246719244e-01};
__device__ __constant__ double c_gl12_w[12] = {
    4.71753363865114114e-02, 1.06939325995319065e-01, 1.60078328543346415e-01,
    2.03167426723065730e-01, 2.33492536538354611e-01, 2.49147045813402690e-01,
    2.49147045813402690e-01, 2.33492536538354611e-01, 2.03167426723065730e-01,
    1.60078328543346415e-01, 1.06939325995319065e-01, 4.71753363865114114e-02};
__device__ __constant__ double c_gl20_x[20] = {
    -9.93128599185094996e-01, -9.63971927277913809e-01, -9.12234428251325946e-01,
    -8.39116971822218782e-01, -7.46331906460150796e-01, -6.36053680726515025e-01,
    -5.10867001950827126e-01, -3.73706088715419549e-01, -2.27785851141645068e-01,
    -7.65265211334973383e-02, 7.65265211334973383e-02,  2.27785851141645068e-01,
    3.73706088715419549e-01,  5.10867001950827126e-01,  6.36053680726515025e-01,
    7.46331906460150796e-01,  8.39116971822218782e-01,  9.12234428251325946e-01,
    9.63971927277913809e-01,  9.93128599185094996e-01};
__device__ __constant__ double c_gl20_w[20] = {
    1.76140071391508932e-02, 4.06014298003864460e-02, 6.26720483341087903e-02,
    8.32767415767047131e-02, 1.01930119817240705e-01, 1.18194531961518606e-01,
    1.31688638449176887e-01, 1.42096109318382402e-01, 1.49172986472604241e-01,
    1.52753387130726281e-01, 1.52753387130726281e-01, 1.49172986472604241e-01,
    1.42096109318382402e-01, 1.31688638449176887e-01, 1.18194531961518606e-01,
    1.01930119817240705e-01, 8.32767415767047131e-02, 6.26720483341087903e-02,
    4.06014298003864460e-02, 1.76140071391508932e-02};

static const double h_gl6_x[6] = {
    -9.32469514203151939e-01, -6.61209386466264482e-01, -2.38619186083196905e-01,
    2.38619186083196905e-01,  6.61209386466264482e-01,  9.32469514203151939e-01};
static const double h_gl6_w[6] = {
    1.71324492379170273e-01, 3.60761573048138717e-01, 4.67913934572691037e-01,
    4.67913934572691037e-01, 3.60761573048138717e-01, 1.71324492379170273e-01};
static const double h_gl12_x[12] = {
    -9.81560634246719244e-01, -9.04117256370474798e-01, -7.69902674194304693e-01,
    -5.87317954286617483e-01, -3.67831498998180184e-01, -1.25233408511468913e-01,
    1.25233408511468913e-01,  3.67831498998180184e-01,  5.87317954286617483e-01,
    7.69902674194304693e-01,  9.04117256370474798e-01,  9.81560634246719244e-01};
static const double h_gl12_w[12] = {
    4.71753363865114114e-02, 1.06939325995319065e-01, 1.60078328543346415e-01,
    2.03167426723065730e-01, 2.33492536538354611e-01, 2.49147045813402690e-01,
    2.49147045813402690e-01, 2.33492536538354611e-01, 2.03167426723065730e-01,
    1.60078328543346415e-01, 1.06939325995319065e-01, 4.71753363865114114e-02};
static const double h_gl20_x[20] = {
    -9.93128599185094996e-01, -9.63971927277913809e-01, -9.12234428251325946e-01,
    -8.39116971822218782e-01, -7.46331906460150796e-01, -6.36053680726515025e-01,
    -5.10867001950827126e-01, -3.73706088715419549e-01, -2.27785851141645068e-01,
    -7.65265211334973383e-02, 7.65265211334973383e-02,  2.27785851141645068e-01,
    3.73706088715419549e-01,  5.10867001950827126e-01,  6.36053680726515025e-01,
    7.46331906460150796e-01,  8.39116971822218782e-01,  9.12234428251325946e-01,
    9.63971927277913809e-01,  9.93128599185094996e-01};
static const double h_gl20_w[20] = {
    1.76140071391508932e-02, 4.06014298003864460e-02, 6.26720483341087903e-02,
    8.32767415767047131e-02, 1.01930119817240705e-01, 1.18194531961518606e-01,
    1.31688638449176887e-01, 1.42096109318382402e-01, 1.49172986472604241e-01,
    1.52753387130726281e-01, 1.52753387130726281e-01, 1.49172986472604241e-01,
    1.42096109318382402e-01, 1.31688638449176887e-01, 1.18194531961518606e-01,
    1.01930119817240705e-01, 8.32767415767047131e-02, 6.26720483341087903e-02,
    4.06014298003864460e-02, 1.76140071391508932e-02};

#define CG_TWO_PI 6.283185307179586476925286766559
#define CG_SQRT_TWO_PI 2.506628274631000502415765284811

// Host: hoist everything that depends on rho only.
inline void bvn_make_tab(double rho, BvnTab* T) {
  double ar = fabs(rho);
  const double *x, *w;
  if (ar < 0.3) { T->ng = 6; x = h_gl6_x; w = h_gl6_w; }
  else if (ar < 0.75) { T->ng = 12; x = h_gl12_x; w = h_gl12_w; }
  else { T->ng = 20; x = h_gl20_x; w = h_gl20_w; }
  T->rho = rho;
  T->high = ar >= 0.925;
  for (int j = 0; j < 20; ++j) T->w[j] = T->c0[j] = T->c1[j] = T->c2[j] = T->c3[j] = 0.0;
  double om = (1.0 - rho) * (1.0 + rho);
  T->as_ = om;
  T->a_ = sqrt(om);
  T->asr_4pi = 0.0;
  if (!T->high) {
    double asr = asin(rho);
    T->asr_4pi = asr / (2.0 * CG_TWO_PI);
    for (int j = 0; j < T->ng; ++j) {
      double sn = sin(asr * (x[j] + 1.0) / 2.0);
      T->w[j] = w[j];
      T->c0[j] = sn;
      T->c1[j] = 1.0 / (1.0 - sn * sn);
    }
  } else if (ar < 1.0) {
    double a = T->a_ / 2.0;
    for (int j = 0; j < T->ng; ++j) {
      double xs = (a * (x[j] + 1.0)) * (a * (x[j] + 1.0));
      double rs = sqrt(1.0 - xs);
      T->w[j] = a * w[j];
      T->c0[j] = 1.0 / xs;
      T->c1[j] = xs / (2.0 * (1.0 + rs) * (1.0 + rs));
      T->c2[j] = 1.0 / rs;
      T->c3[j] = xs;
    }
  }
  T->inv_s = om > 0 ? 1.0 / sqrt(om) : 0.0;
  T->inv_2om = om > 0 ? 1.0 / (2.0 * om) : 0.0;
  T->inv_2pis = om > 0 ? 1.0 / (CG_TWO_PI * sqrt(om)) : 0.0;
}

__device__ __forceinline__ double phid(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }

// P(X > h, Y > k), correlation from the table.
__device__ __forceinline__ double bvnd_tab(double h, double k, const BvnTab& T) {
  double hk = h * k;
  if (!T.high) {
    double bvn = 0.0;
    if (T.rho != 0.0) {
      const double hs = 0.5 * (h * h + k * k);
      for (int j = 0; j < T.ng; ++j) bvn += T.w[j] * exp((T.c0[j] * hk - hs) * T.c1[j]);
      bvn *= T.asr_4pi;
    }
    return bvn + phid(-h) * phid(-k);
  }
  if (T.rho < 0.0) { k = -k; hk = -hk; }
  double bvn = 0.0;
  if (T.as_ > 0.0) {
    const double as = T.as_, a = T.a_;
    const double bs = (h - k) * (h - k);
    const double c = (4.0 - hk) / 8.0, d = (12.0 - hk) / 16.0;
    double asr = -(bs / as + hk) / 2.0;
    if (asr > -100.0)
      bvn = a * exp(asr) * (1.0 - c * (bs - as) * (1.0 - d * bs / 5.0) / 3.0 + c * d * as * as / 5.0);
    if (-hk < 100.0) {
      double b = sqrt(bs);
      bvn -= exp(-hk / 2.0) * CG_SQRT_TWO_PI * phid(-b / a) * b * (1.0 - c * bs * (1.0 - d * bs / 5.0) / 3.0);
    }
    for (int j = 0; j < T.ng; ++j) {
      asr = -(bs * T.c0[j] + hk) / 2.0;
      if (asr > -100.0) {
        const double xs = T.c3[j];
        bvn += T.w[j] * exp(asr) * (exp(-hk * T.c1[j]) * T.c2[j] - (1.0 + c * xs * (1.0 + d * xs)));
      }
    }
    bvn = -bvn / CG_TWO_PI;
  }
  if (T.rho > 0.0) return bvn + phid(-fmax(h, k));
  bvn = -bvn;
  if (k > h) bvn += (h < 0.0) ? (phid(k) - phid(h)) : (phid(-h) - phid(-k));
  return bvn;
}


// ---- pair-hoisted high-correlation branch ---------------------------------------------------------
// In the CGPCM  h - k = -(x1 - x2) = (p - q)(tx_k - tx_l)  does not depend on the observation: for a fixed
// pair (k, l) Genz's  bs = (h - k)^2  is a constant.  Of the two exp() per quadrature node of the
// |rho| >= 0.925 branch one, exp(-bs / (2 xs_j)), is therefore hoisted out of the loop over observations
// (together with sqrt(bs) and Phi(-b / a)), and every term carries the common factor exp(-hk / 2).
// What remains of the 20-node sum is a function of the single scalar u = hk,
//
//     f(u) = sum_j P_j / rs_j * exp(-c1_j u)  -  (K0 + c K1 + c d K2),      P_j = (a/2) w_j exp(-bs / (2 xs_j)),
//
// a positive combination of 20 slowly varying exponentials (c1_j u stays inside +-4 on the interval
// |u| <= 200 that Genz's own cut-offs leave).  It is evaluated as a Chebyshev series in u / 200 whose
// coefficients are per-pair linear combinations a_m = sum_j (P_j / rs_j) B[m][j] of the exact Chebyshev
// coefficients B[m][j] = (2 - [m = 0]) (-1)^m I_m(200 c1_j) of the node exponentials (host, Bessel series):
// ~2 FMAs per degree (18 at rho = 0.97) in place of 20 exp().  Truncation: coefficients below 1e-18 of the
// leading one are dropped -- the series reproduces the node sum to rounding (tests/test_gpu_psi.py).
// Genz's skip tests (asr > -100, -hk < 100) only drop terms below 4e-44 and are not needed in the factored
// form; outside |hk| <= 200 the tail term alone (hk > 200) or the un-hoisted routine (hk < -200) is used.
// Valid for rho > 0 (always true here: rho = gamma / (alpha + gamma + omega)).
constexpr int BVN_PAIR_THREADS = 256;
constexpr int BVN_CHEB_MAXDEG = 40;
constexpr double BVN_CHEB_U = 200.0;

struct BvnPair {
  double bs, P0, Pb, K0, K1, K2;
};

// Host: B[m][j] for m = 0..deg (row-major, 20 per row) and the degree.  I_m by its power series (200 c1_j < 5).
inline int bvn_make_cheb(const BvnTab& T, double* B /* (BVN_CHEB_MAXDEG + 1) * 20 */) {
  double mx[BVN_CHEB_MAXDEG + 1];
  for (int m = 0; m <= BVN_CHEB_MAXDEG; ++m) {
    mx[m] = 0.0;
    for (int j = 0; j < 20; ++j) {
      const long double half = 0.5L * (long double)(BVN_CHEB_U * T.c1[j]);
      // I_m(2 half) = sum_k half^(2k + m) / (k! (k + m)!)
      long double term = 1.0L;
      for (int i = 1; i <= m; ++i) term *= half / i;
      long double sum = term;
      for (int k = 1; k < 80; ++k) {
        term *= half * half / ((long double)k * (k + m));
        sum += term;
        if (term < 1e-25L * sum) break;
      }
      double v = (double)((m == 0 ? 1.0L : 2.0L) * ((m & 1) ? -sum : sum));
      B[m * 20 + j] = v;
      mx[m] = fmax(mx[m], fabs(v));
    }
  }
  int deg = 0;
  for (int m = 0; m <= BVN_CHEB_MAXDEG; ++m)
    if (mx[m] > 1e-18 * mx[0]) deg = m;
  return deg;
}

// Per-pair set-up.  sA: shared [deg + 1][BVN_PAIR_THREADS] (column threadIdx.x belongs to this thread);
// sB: shared copy of B (already filled, (deg + 1) * 20).
__device__ __forceinline__ void bvn_pair_init(double hmk, const BvnTab& T, BvnPair& R, double* sA, const double* sB,
                                              int deg) {
  const double bs = hmk * hmk;
  R.bs = bs;
  R.P0 = T.a_ * exp(-0.5 * bs / T.as_);
  const double b = fabs(hmk);
  R.Pb = CG_SQRT_TWO_PI * phid(-b / T.a_) * b;
  double pc[20];
  double k0 = 0.0, k1 = 0.0, k2 = 0.0;
#pragma unroll
  for (int j = 0; j < 20; ++j) {
    const double P = T.w[j] * exp(-0.5 * bs * T.c0[j]);     // (a/2) w_j exp(-bs / (2 xs_j))
    const double xs = T.c3[j];
    k0 += P;
    k1 += P * xs;
    k2 += P * xs * xs;
    pc[j] = P * T.c2[j];
  }
  R.K0 = k0; R.K1 = k1; R.K2 = k2;
  for (int m = 0; m <= deg; ++m) {
    double a = 0.0;
#pragma unroll
    for (int j = 0; j < 20; ++j) a += pc[j] * sB[m * 20 + j];
    sA[m * BVN_PAIR_THREADS + threadIdx.x] = a;
  }
}

// Phi_2(x1, x2; rho) for the pair whose constants are R / sA  (same value as bvnd_tab(-x1, -x2, T)).
__device__ __forceinline__ double bvn_cdf_pair(double x1, double x2, const BvnTab& T, const BvnPair& R,
                                               const double* sA, int deg) {
  const double hk = x1 * x2;
  const double tail = phid(fmin(x1, x2));
  if (hk > BVN_CHEB_U) return tail;                      // every remaining term is below exp(-100) (Genz's own cut)
  if (hk < -BVN_CHEB_U) return bvnd_tab(-x1, -x2, T);    // outside the Chebyshev interval: un-hoisted routine
  const double E1 = exp(-0.5 * hk);
  const double c = (4.0 - hk) * 0.125, d = (12.0 - hk) * 0.0625;
  const double bs = R.bs, as = T.as_;
  const double t5 = 1.0 - d * bs * 0.2;
  double s = R.P0 * (1.0 - c * (bs - as) * t5 * (1.0 / 3.0) + c * d * as * as * 0.2) -
             R.Pb * (1.0 - c * bs * t5 * (1.0 / 3.0)) - (R.K0 + c * (R.K1 + d * R.K2));
  // Clenshaw:  sum_m a_m T_m(u / U)
  const double ut = hk * (1.0 / BVN_CHEB_U), t2 = ut + ut;
  const double* a = sA + threadIdx.x;
  double b1 = 0.0, b2 = 0.0;
#pragma unroll 4
  for (int m = deg; m >= 1; --m) {
    const double b0 = fma(t2, b1, a[m * BVN_PAIR_THREADS]) - b2;
    b2 = b1;
    b1 = b0;
  }
  s += fma(ut, b1, a[0]) - b2;
  return tail - E1 * s * (1.0 / CG_TWO_PI);
}

// Phi_2(x1, x2; rho) and its three partial derivatives.
__device__ __forceinline__ void bvn_partials_tab(double x1, double x2, const BvnTab& T, double& d1, double& d2,
                                                 double& dr) {
  const double inv_sqrt_2pi = 0.39894228040143267794;
  d1 = inv_sqrt_2pi * exp(-0.5 * x1 * x1) * phid((x2 - T.rho * x1) * T.inv_s);
  d2 = inv_sqrt_2pi * exp(-0.5 * x2 * x2) * phid((x1 - T.rho * x2) * T.inv_s);
  dr = exp(-(x1 * x1 - 2.0 * T.rho * x1 * x2 + x2 * x2) * T.inv_2om) * T.inv_2pis;
}
// The pair-hoisted value and the three partials of one element with all of its transcendentals evaluated in lock-step
// (cgmath.cuh): five exp -- the caller's envelope exp(G), exp(-hk / 2), the two marginal densities and the bivariate
// density -- and three erfc -- Phi(min(x1, x2)) and the two conditional Phis.  Same formulas as bvn_cdf_pair +
// bvn_partials_tab; one materialised polynomial coefficient feeds five (three) FMAs and the eight chains overlap.
// Returns false when hk < -U (outside the Chebyshev interval): the caller takes the un-hoisted routine.
__device__ __forceinline__ bool bvn_pair_all(double G, double x1, double x2, const BvnTab& T, const BvnPair& R,
                                             const double* sA, int deg, double& envexp, double& cdf, double& d1,
                                             double& d2, double& dr) {
  const double hk = x1 * x2;
  if (hk < -BVN_CHEB_U) return false;
  const double q = x1 * x1 - 2.0 * T.rho * x1 * x2 + x2 * x2;
  const double ea[5] = {G, -0.5 * hk, -0.5 * x1 * x1, -0.5 * x2 * x2, -q * T.inv_2om};
  const double is2 = 0.70710678118654752440;
  const double ca[3] = {-fmin(x1, x2) * is2, -((x2 - T.rho * x1) * T.inv_s) * is2, -((x1 - T.rho * x2) * T.inv_s) * is2};
  double ev[5], cx[3];
  cg_exp_neg<5, true>(ea, ev);
  // The Gaussian factor of every erfc(z) = exp(-z^2) erfcx(|z|) here is one of the five exponentials above:
  //   Phi(m), m = min(x1, x2):             exp(-z^2) = exp(-m^2 / 2) = ev[2] or ev[3]
  //   phi(x1) Phi((x2 - rho x1) / s):      exp(-x1^2 / 2 - z^2) = exp(-q / (2 (1 - rho^2))) = ev[4]   (likewise for x2)
  // so only the rational part erfcx(|z|) is evaluated (cgmath.cuh), and erfc(z) = 2 - erfc(-z) for z < 0.
  cg_erfcx_abs<3>(ca, cx);
  envexp = ev[0];
  const double inv_sqrt_2pi = 0.39894228040143267794;
  const double rb = ev[4] * cx[1], rc = ev[4] * cx[2];
  d1 = (0.5 * inv_sqrt_2pi) * (ca[1] < 0.0 ? fma(2.0, ev[2], -rb) : rb);
  d2 = (0.5 * inv_sqrt_2pi) * (ca[2] < 0.0 ? fma(2.0, ev[3], -rc) : rc);
  dr = ev[4] * T.inv_2pis;
  const double ra = (x1 <= x2 ? ev[2] : ev[3]) * cx[0];
  const double tail = 0.5 * (ca[0] < 0.0 ? 2.0 - ra : ra);
  if (hk > BVN_CHEB_U) { cdf = tail; return true; }      // every remaining term is below exp(-100) (Genz's own cut)
  const double E1 = ev[1];
  const double c = (4.0 - hk) * 0.125, d = (12.0 - hk) * 0.0625;
  const double bs = R.bs, as = T.as_;
  const double t5 = 1.0 - d * bs * 0.2;
  double s = R.P0 * (1.0 - c * (bs - as) * t5 * (1.0 / 3.0) + c * d * as * as * 0.2) -
             R.Pb * (1.0 - c * bs * t5 * (1.0 / 3.0)) - (R.K0 + c * (R.K1 + d * R.K2));
  const double ut = hk * (1.0 / BVN_CHEB_U), t2 = ut + ut;
  const double* a = sA + threadIdx.x;
  double b1 = 0.0, b2 = 0.0;
#pragma unroll 4
  for (int m = deg; m >= 1; --m) {
    const double b0 = fma(t2, b1, a[m * BVN_PAIR_THREADS]) - b2;
    b2 = b1;
    b1 = b0;
  }
  s += fma(ut, b1, a[0]) - b2;
  cdf = tail - E1 * s * (1.0 / CG_TWO_PI);
  return true;
}

__device__ __forceinline__ void bvn_cdf_grad_tab(double x1, double x2, const BvnTab& T, double& cdf,
                                                 double& d1, double& d2, double& dr) {
  cdf = bvnd_tab(-x1, -x2, T);
  const double inv_sqrt_2pi = 0.39894228040143267794;
  d1 = inv_sqrt_2pi * exp(-0.5 * x1 * x1) * phid((x2 - T.rho * x1) * T.inv_s);
  d2 = inv_sqrt_2pi * exp(-0.5 * x2 * x2) * phid((x1 - T.rho * x2) * T.inv_s);
  dr = exp(-(x1 * x1 - 2.0 * T.rho * x1 * x2 + x2 * x2) * T.inv_2om) * T.inv_2pis;
}

// Per-element correlation (stand-alone op).
__device__ inline double bvnd_general(double h, double k, double r) {
  const double ar = fabs(r);
  int ng;
  const double *x, *w;
  if (ar < 0.3) { ng = 6; x = c_gl6_x; w = c_gl6_w; }
  else if (ar < 0.75) { ng = 12; x = c_gl12_x; w = c_gl12_w; }
  else { ng = 20; x = c_gl20_x; w = c_gl20_w; }
  double hk = h * k;
  if (ar < 0.925) {
    double bvn = 0.0;
    if (ar > 0.0) {
      const double hs = 0.5 * (h * h + k * k), asr = asin(r);
      for (int j = 0; j < ng; ++j) {
        double sn = sin(asr * (x[j] + 1.0) / 2.0);
        bvn += w[j] * exp((sn * hk - hs) / (1.0 - sn * sn));
      }
      bvn *= asr / (2.0 * CG_TWO_PI);
    }
    return bvn + phid(-h) * phid(-k);
  }
  if (r < 0.0) { k = -k; hk = -hk; }
  double bvn = 0.0;
  if (ar < 1.0) {
    const double as = (1.0 - r) * (1.0 + r);
    double a = sqrt(as);
    const double bs = (h - k) * (h - k);
    const double c = (4.0 - hk) / 8.0, d = (12.0 - hk) / 16.0;
    double asr = -(bs / as + hk) / 2.0;
    if (asr > -100.0)
      bvn = a * exp(asr) * (1.0 - c * (bs - as) * (1.0 - d * bs / 5.0) / 3.0 + c * d * as * as / 5.0);
    if (-hk < 100.0) {
      double b = sqrt(bs);
      bvn -= exp(-hk / 2.0) * CG_SQRT_TWO_PI * phid(-b / a) * b * (1.0 - c * bs * (1.0 - d * bs / 5.0) / 3.0);
    }
    a = a / 2.0;
    for (int j = 0; j < ng; ++j) {
      double xs = (a * (x[j] + 1.0)) * (a * (x[j] + 1.0));
      double rs = sqrt(1.0 - xs);
      asr = -(bs / xs + hk) / 2.0;
      if (asr > -100.0)
        bvn += a * w[j] * exp(asr) * (exp(-hk * xs / (2.0 * (1.0 + rs) * (1.0 + rs))) / rs - (1.0 + c * xs * (1.0 + d * xs)));
    }
    bvn = -bvn / CG_TWO_PI;
  }
  if (r > 0.0) return bvn + phid(-fmax(h, k));
  bvn = -bvn;
  if (k > h) bvn += (h < 0.0) ? (phid(k) - phid(h)) : (phid(-h) - phid(-k));
  return bvn;
}

__global__ void bvn_cdf_kernel(const double* __restrict__ x1, const double* __restrict__ x2,
                               const double* __restrict__ rho, double* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = bvnd_general(-x1[i], -x2[i], rho[i]);
}

}  // namespace cg
