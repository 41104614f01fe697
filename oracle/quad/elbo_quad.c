/* Quad-precision (IEEE binary128, libquadmath) evaluation of the VCGPCM ELBO, its 7 terms and directional
 * derivatives.  TEST INFRASTRUCTURE ONLY (oracle/): it arbitrates between two FP64 implementations -- the reference's
 * arithmetic (oracle/_ref, oracle/model.py) and the CUDA path -- when they differ by more than BASELINE.json's 1e-9 at
 * trained points (s2 ~ 1e-3, cond(Kh) ~ 1 / reg), where every FP64 evaluation carries rounding noise amplified by the
 * conditioning.  113 mantissa bits leave ~1e-25 after the same amplification.
 *
 * What it restates (reference src/core/cgpcm.py:111-268, 458-477, 518-575,
 * src/core/distribution.py:60-76, src/core/kernel.py:43-46): prior kernels with jitter, the closed forms of the Psi
 * statistics a, Ahh, Ahx, Axx (erfc / bivariate normal CDF), the sums over observations, the optimal q(z), the 7 terms.
 * The bivariate normal CDF is NOT Genz's 6/12/20-node rule (1e-15 is not enough here) but composite Gauss-Legendre on
 * Plackett's integral  Phi2(x, y; rho) = Phi(x) Phi(y) + 1/(2 pi) int_0^asin(rho) exp(-(x^2 + y^2 - 2 x y sin u) / (2 cos^2 u)) du
 * with nodes computed in binary128 and enough panels for the integrand's decay.
 * Elements whose Gaussian envelope is below exp(-110) (4e-48 of an O(1) prefactor) are skipped.
 *
 * Gradients: central differences along a direction v in binary128, with the Richardson pair (h, 2h) so that the caller
 * sees the truncation error.   Build: oracle/quad/build.sh  (gcc -O2 -fopenmp ... -lquadmath).
 */
#include <quadmath.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef __float128 Q;

#define QPI M_PIq
#define NQ 24          /* Gauss-Legendre nodes per panel */

static Q gl_x[NQ], gl_w[NQ];
static int gl_ready = 0;

static void gl_init(void) {
  if (gl_ready) return;
  for (int i = 0; i < NQ; ++i) {
    Q x = cosq(QPI * (i + 0.75Q) / (NQ + 0.5Q));
    Q dp = 1;
    for (int it = 0; it < 100; ++it) {
      Q p0 = 1, p1 = x;
      for (int k = 2; k <= NQ; ++k) {
        Q p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0 = p1;
        p1 = p2;
      }
      dp = NQ * (x * p1 - p0) / (x * x - 1);
      Q dx = p1 / dp;
      x -= dx;
      if (fabsq(dx) < 1e-33Q) break;
    }
    {
      Q p0 = 1, p1 = x;
      for (int k = 2; k <= NQ; ++k) {
        Q p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        p0 = p1;
        p1 = p2;
      }
      dp = NQ * (x * p1 - p0) / (x * x - 1);
    }
    gl_x[i] = x;
    gl_w[i] = 2 / ((1 - x * x) * dp * dp);
  }
  gl_ready = 1;
}

static Q phi_cdf(Q x) { return 0.5Q * erfcq(-x * M_SQRT1_2q); }

/* sin / cos at the nodes of the single-panel rule for the last rho seen (rho is one constant per model) */
static Q tab_rho = -2, tab_sn[NQ], tab_cs2[NQ];
static void bvn_table(Q rho) {
  if (rho == tab_rho) return;
#pragma omp critical(bvn_tab)
  {
    if (rho != tab_rho) {
      Q as = asinq(rho);
      for (int i = 0; i < NQ; ++i) {
        Q u = 0.5Q * as * (1 + gl_x[i]);
        tab_sn[i] = sinq(u);
        Q c = cosq(u);
        tab_cs2[i] = 2 * c * c;
      }
      tab_rho = rho;
    }
  }
}

/* Phi2(x, y; rho), |rho| < 1 */
static Q bvn_cdf_q(Q x, Q y, Q rho) {
  Q base = phi_cdf(x) * phi_cdf(y);
  if (rho == 0) return base;
  Q as = asinq(rho);
  Q s2 = x * x + y * y, xy = x * y;
  /* exponent g(u) = (s2 - 2 xy sin u) / (2 cos^2 u) at both ends; the integrand is exp(-g) */
  Q g0 = 0.5Q * s2;
  Q ge = (s2 - 2 * xy * rho) / (2 * (1 - rho * rho));
  Q gmin = g0 < ge ? g0 : ge;
  /* g is monotone or has one interior extremum; a lower bound of min g over the interval: */
  if (xy != 0) {
    /* stationary point of g: sin u* = (s2 - sqrt(s2^2 - 4 xy^2)) / (2 xy)  (when inside the interval) */
    Q disc = s2 * s2 - 4 * xy * xy;
    if (disc >= 0) {
      Q su = (s2 - sqrtq(disc)) / (2 * xy);
      if ((as > 0 && su > 0 && su < rho) || (as < 0 && su < 0 && su > rho)) {
        Q gs = (s2 - 2 * xy * su) / (2 * (1 - su * su));
        if (gs < gmin) gmin = gs;
      }
    }
  }
  if (gmin > 120) return base;                 /* integral < exp(-120): below 1e-52 */
  Q range = fabsq(g0 - ge);
  int panels = 1 + (int)(range / 6);
  if (panels > 64) panels = 64;
  Q sum = 0;
  Q hw = as / panels;
  if (panels == 1 && rho == tab_rho) {
    Q acc = 0;
    for (int i = 0; i < NQ; ++i) acc += gl_w[i] * expq(-(s2 - 2 * xy * tab_sn[i]) / tab_cs2[i]);
    return base + acc * 0.5Q * as / (2 * QPI);
  }
  for (int p = 0; p < panels; ++p) {
    Q c = hw * (p + 0.5Q), r = 0.5Q * hw;
    Q acc = 0;
    for (int i = 0; i < NQ; ++i) {
      Q u = c + r * gl_x[i];
      Q sn = sinq(u), cs = cosq(u);
      acc += gl_w[i] * expq(-(s2 - 2 * xy * sn) / (2 * cs * cs));
    }
    sum += acc * r;
  }
  return base + sum / (2 * QPI);
}

/* in-place lower Cholesky; returns 0 or the failing pivot + 1 */
static int chol_q(Q* A, int n) {
  for (int j = 0; j < n; ++j) {
    Q d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0)) return j + 1;
    d = sqrtq(d);
    A[j * n + j] = d;
#pragma omp parallel for schedule(static)
    for (int i = j + 1; i < n; ++i) {
      Q s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
    for (int i = 0; i < j; ++i) A[i * n + j] = 0;
  }
  return 0;
}

/* X = L^-1 (lower) */
static void tri_inv_q(const Q* L, Q* X, int n) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int c = 0; c < n; ++c) {
    for (int r = 0; r < n; ++r) X[r * n + c] = 0;
    for (int r = c; r < n; ++r) {
      Q v = (r == c) ? 1 : 0;
      for (int m = c; m < r; ++m) v -= L[r * n + m] * X[m * n + c];
      X[r * n + c] = v / L[r * n + r];
    }
  }
}

/* inv = (L L^T)^-1 = X^T X */
static void chol_inverse_q(const Q* L, Q* inv, int n) {
  Q* X = (Q*)malloc(sizeof(Q) * n * n);
  tri_inv_q(L, X, n);
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      Q s = 0;
      for (int k = i; k < n; ++k) s += X[k * n + i] * X[k * n + j];
      inv[i * n + j] = inv[j * n + i] = s;
    }
  free(X);
}

static Q logdet_q(const Q* L, int n) {
  Q s = 0;
  for (int i = 0; i < n; ++i) s += logq(L[i * n + i]);
  return 2 * s;
}

/* The ELBO and its 7 terms at params (binary128).  Returns 0, or 10 * which + (pivot != 0) on a failed Cholesky. */
static int elbo_core(int n, const double* t, const double* y, int nh, const double* th, int nx, const double* tx,
                     const Q* params, double reg_d, int causal, Q* elbo, Q terms[7]) {
  gl_init();
  const Q reg = (Q)reg_d;
  const Q s2 = expq(params[0]), s2f = expq(params[1]);
  const Q alpha = expq(params[2]), gamma = expq(params[3]), omega = expq(params[4]);
  const Q r = s2f / s2, c0 = sqrtq(s2f) / s2;
  const Q* mu = params + 5;
  const Q* varu = params + 5 + nh;
  int rc = 0;

  Q* Kh = (Q*)malloc(sizeof(Q) * nh * nh);
  Q* iKh = (Q*)malloc(sizeof(Q) * nh * nh);
  Q* Kx = (Q*)malloc(sizeof(Q) * nx * nx);
  Q* Lx = (Q*)malloc(sizeof(Q) * nx * nx);
  Q* iKx = (Q*)malloc(sizeof(Q) * nx * nx);
  Q* Ahh = (Q*)malloc(sizeof(Q) * nh * nh);
  Q* m2 = (Q*)malloc(sizeof(Q) * nh * nh);
  Q* var = (Q*)malloc(sizeof(Q) * nh * nh);
  Q* H = (Q*)malloc(sizeof(Q) * nh * nh);
  Q* sAxx = (Q*)calloc((size_t)nx * nx, sizeof(Q));
  Q* C1 = (Q*)calloc((size_t)nx * nx, sizeof(Q));
  Q* Qm = (Q*)calloc((size_t)nh * nh, sizeof(Q));
  Q* Y = (Q*)calloc((size_t)nh * nx, sizeof(Q));

  /* prior kernels (cgpcm.py:214-229, kernel.py:43-46) */
  for (int i = 0; i < nh; ++i)
    for (int j = 0; j < nh; ++j) {
      Q ti = th[i], tj = th[j];
      Kh[i * nh + j] = expq(-alpha * (ti * ti + tj * tj) - gamma * (ti - tj) * (ti - tj)) + (i == j ? reg : 0);
    }
  for (int k = 0; k < nx; ++k)
    for (int l = 0; l < nx; ++l) {
      Q d = (Q)tx[k] - (Q)tx[l];
      Kx[k * nx + l] = sqrtq(0.5Q * QPI / omega) * expq(-0.5Q * omega * d * d) + (k == l ? reg : 0);
    }
  if ((rc = chol_q(Kh, nh))) { rc = 10; goto done; }
  chol_inverse_q(Kh, iKh, nh);
  memcpy(Lx, Kx, sizeof(Q) * nx * nx);
  if ((rc = chol_q(Lx, nx))) { rc = 20; goto done; }
  chol_inverse_q(Lx, iKx, nx);
  /* Kx (jittered) is needed again for P: rebuild the upper triangle lost by nothing -- Kx itself is untouched */

  /* q(u) (cgpcm.py:444-445, distribution.py:35-42) */
  {
    Q* Lq = (Q*)calloc((size_t)nh * nh, sizeof(Q));
    long e = 0;
    for (int i = 0; i < nh; ++i)
      for (int j = 0; j <= i; ++j) Lq[i * nh + j] = varu[e++];
    for (int i = 0; i < nh; ++i)
      for (int j = 0; j < nh; ++j) {
        Q s = 0;
        int m = i < j ? i : j;
        for (int k = 0; k <= m; ++k) s += Lq[i * nh + k] * Lq[j * nh + k];
        var[i * nh + j] = s + (i == j ? reg : 0);
        m2[i * nh + j] = var[i * nh + j] + mu[i] * mu[j];
        H[i * nh + j] = m2[i * nh + j] - iKh[i * nh + j];
      }
    free(Lq);
  }

  /* Psi statistics that do not depend on n */
  const Q A = alpha + gamma + omega;
  const Q a_stat = (causal ? 0.5Q : 1) * sqrtq(QPI / (2 * alpha));
  {
    const Q B = alpha + gamma;
    for (int i = 0; i < nh; ++i)
      for (int j = 0; j < nh; ++j) {
        Q ti = th[i], tj = th[j];
        Q b = -2 * gamma * (ti + tj), c = -B * (ti * ti + tj * tj);
        Q v = sqrtq(QPI / (2 * B)) * expq(c + b * b / (8 * B));
        if (causal) v *= 0.5Q * erfcq(b / (2 * sqrtq(2 * B)));
        Ahh[i * nh + j] = v;
      }
  }
  const Q det = 4 * (A * A - gamma * gamma);
  const Q S11 = 2 * A / det, S12 = 2 * gamma / det, rho = gamma / A;
  const Q g1 = omega * (1 - 2 * omega * S11), g2 = 4 * omega * omega * S12;
  const Q pp = 2 * omega * sqrtq(S11), qq = 2 * omega * S12 / sqrtq(S11);
  const Q pref_xx = 2 * QPI / sqrtq(det);
  const Q pref_hx = (causal ? 0.5Q : 1) * sqrtq(QPI / A);
  const Q CUT = 110;
  bvn_table(rho);

  /* sums over observations */
#pragma omp parallel
  {
    Q* lA = (Q*)malloc(sizeof(Q) * nh * nx);      /* A_n restricted to its window [k0, k1) */
    Q* lT = (Q*)malloc(sizeof(Q) * nh * nx);
    Q* pAxx = (Q*)calloc((size_t)nx * nx, sizeof(Q));
    Q* pC1 = (Q*)calloc((size_t)nx * nx, sizeof(Q));
    Q* pQ = (Q*)calloc((size_t)nh * nh, sizeof(Q));
    Q* pY = (Q*)calloc((size_t)nh * nx, sizeof(Q));
    int* win = (int*)malloc(sizeof(int) * nx);
#pragma omp for schedule(dynamic, 1)
    for (int o = 0; o < n; ++o) {
      const Q tn = t[o], yn = y[o];
      /* window of inducing inputs: the envelope of Ahx in d is exp(-lam d^2), lam = e_dd - e_hd^2 / (4 e_hh) */
      const Q e_hh = ((alpha + gamma) * A - gamma * gamma) / A, e_dd = omega * (alpha + gamma) / A;
      const Q e_hd = 2 * gamma * omega / A;
      const Q lam = e_dd - e_hd * e_hd / (4 * e_hh);
      int kw = 0;
      for (int k = 0; k < nx; ++k) {
        Q d = tn - (Q)tx[k];
        if (lam * d * d <= CUT) win[kw++] = k;
      }
      /* Ahx */
      for (int i = 0; i < nh; ++i) {
        Q ti = th[i];
        for (int w = 0; w < kw; ++w) {
          Q d = tn - (Q)tx[win[w]];
          Q b = -2 * gamma * ti - 2 * omega * d;
          Q E = -(alpha + gamma) * ti * ti - omega * d * d + b * b / (4 * A);
          Q v = 0;
          if (E > -CUT - 40) {
            v = pref_hx * expq(E);
            if (causal) v *= erfcq(b / (2 * sqrtq(A)));
          }
          lA[i * kw + w] = v;
          pY[i * nx + win[w]] += yn * v;
        }
      }
      /* T = H A ; C1 += A^T T (lower) ; V = A iKx_win ; Q += V A^T (lower) */
      for (int i = 0; i < nh; ++i)
        for (int w = 0; w < kw; ++w) {
          Q s = 0;
          for (int j = 0; j < nh; ++j) s += H[i * nh + j] * lA[j * kw + w];
          lT[i * kw + w] = s;
        }
      for (int w = 0; w < kw; ++w)
        for (int v = 0; v <= w; ++v) {
          Q s = 0;
          for (int i = 0; i < nh; ++i) s += lA[i * kw + w] * lT[i * kw + v];
          pC1[win[w] * nx + win[v]] += s;
        }
      for (int i = 0; i < nh; ++i)
        for (int w = 0; w < kw; ++w) {
          Q s = 0;
          for (int v = 0; v < kw; ++v) s += lA[i * kw + v] * iKx[win[v] * nx + win[w]];
          lT[i * kw + w] = s;
        }
      for (int i = 0; i < nh; ++i)
        for (int j = 0; j <= i; ++j) {
          Q s = 0;
          for (int w = 0; w < kw; ++w) s += lT[i * kw + w] * lA[j * kw + w];
          pQ[i * nh + j] += s;
        }
      /* Axx (lower) */
      for (int k = 0; k < nx; ++k) {
        Q dk = tn - (Q)tx[k];
        for (int l = 0; l <= k; ++l) {
          Q dl = tn - (Q)tx[l];
          Q G = -g1 * (dk * dk + dl * dl) + g2 * dk * dl;
          if (G < -CUT) continue;
          Q v = pref_xx * expq(G);
          if (causal) v *= bvn_cdf_q(pp * dk + qq * dl, qq * dk + pp * dl, rho);
          pAxx[k * nx + l] += v;
        }
      }
    }
#pragma omp critical
    {
      for (long e = 0; e < (long)nx * nx; ++e) { sAxx[e] += pAxx[e]; C1[e] += pC1[e]; }
      for (long e = 0; e < (long)nh * nh; ++e) Qm[e] += pQ[e];
      for (long e = 0; e < (long)nh * nx; ++e) Y[e] += pY[e];
    }
    free(lA); free(lT); free(pAxx); free(pC1); free(pQ); free(pY); free(win);
  }
  for (int k = 0; k < nx; ++k)
    for (int l = 0; l < k; ++l) { sAxx[l * nx + k] = sAxx[k * nx + l]; C1[l * nx + k] = C1[k * nx + l]; }
  for (int i = 0; i < nh; ++i)
    for (int j = 0; j < i; ++j) Qm[j * nh + i] = Qm[i * nh + j];

  /* M x M algebra (cgpcm.py:458-477, 518-575) */
  {
    Q sum_y2 = 0;
    for (int o = 0; o < n; ++o) sum_y2 += (Q)y[o] * (Q)y[o];
    Q* P = (Q*)malloc(sizeof(Q) * nx * nx);
    for (long e = 0; e < (long)nx * nx; ++e) P[e] = Kx[e] + r * (sAxx[e] + C1[e]);
    for (int k = 0; k < nx; ++k) P[k * nx + k] += reg;
    Q logdet_kx = logdet_q(Lx, nx);
    if ((rc = chol_q(P, nx))) { rc = 30; free(P); goto done; }
    Q logdet_p = logdet_q(P, nx);
    Q* lamv = (Q*)malloc(sizeof(Q) * nx);
    for (int k = 0; k < nx; ++k) {
      Q s = 0;
      for (int i = 0; i < nh; ++i) s += Y[i * nx + k] * mu[i];
      lamv[k] = c0 * s;
    }
    Q fit = 0;
    for (int k = 0; k < nx; ++k) {          /* forward substitution: w = Lp^-1 lam */
      Q s = lamv[k];
      for (int m = 0; m < k; ++m) s -= P[k * nx + m] * lamv[m];
      lamv[k] = s / P[k * nx + k];
      fit += lamv[k] * lamv[k];
    }
    Q tr_ikh_ahh = 0, tr_ikx_axx = 0, tr_ikh_q = 0, tr_bhh_m2 = 0;
    for (long e = 0; e < (long)nh * nh; ++e) {
      tr_ikh_ahh += iKh[e] * Ahh[e];
      tr_ikh_q += iKh[e] * Qm[e];
      tr_bhh_m2 += (n * Ahh[e] - Qm[e]) * m2[e];
    }
    for (long e = 0; e < (long)nx * nx; ++e) tr_ikx_axx += iKx[e] * sAxx[e];
    Q sum_b = n * a_stat - n * tr_ikh_ahh - tr_ikx_axx + tr_ikh_q;
    /* KL(N(mu, var) || N(0, iKh + reg I))  (distribution.py:60-76) */
    Q* So = (Q*)malloc(sizeof(Q) * nh * nh);
    Q* iSo = (Q*)malloc(sizeof(Q) * nh * nh);
    Q* Lv = (Q*)malloc(sizeof(Q) * nh * nh);
    memcpy(So, iKh, sizeof(Q) * nh * nh);
    for (int i = 0; i < nh; ++i) So[i * nh + i] += reg;
    memcpy(Lv, var, sizeof(Q) * nh * nh);
    int r1 = chol_q(So, nh), r2 = chol_q(Lv, nh);
    if (r1 || r2) { rc = r1 ? 40 : 50; free(P); free(lamv); free(So); free(iSo); free(Lv); goto done; }
    chol_inverse_q(So, iSo, nh);
    Q tr = 0, quad = 0;
    for (int i = 0; i < nh; ++i)
      for (int j = 0; j < nh; ++j) {
        tr += iSo[i * nh + j] * var[i * nh + j];
        quad += mu[i] * iSo[i * nh + j] * mu[j];
      }
    Q KL = 0.5Q * (tr + quad - nh + logdet_q(So, nh) - logdet_q(Lv, nh));
    terms[0] = -0.5Q * n * logq(2 * QPI * s2) - 0.5Q * sum_y2 / s2;
    terms[1] = 0.5Q * logdet_kx;
    terms[2] = -0.5Q * logdet_p;
    terms[3] = 0.5Q * fit;
    terms[4] = -0.5Q * r * sum_b;
    terms[5] = -0.5Q * r * tr_bhh_m2;
    terms[6] = -KL;
    Q e = 0;
    for (int i = 0; i < 7; ++i) e += terms[i];
    *elbo = e;
    free(P); free(lamv); free(So); free(iSo); free(Lv);
  }
done:
  free(Kh); free(iKh); free(Kx); free(Lx); free(iKx); free(Ahh); free(m2); free(var); free(H);
  free(sAxx); free(C1); free(Qm); free(Y);
  return rc;
}

static void split(Q v, double* hi, double* lo) {
  *hi = (double)v;
  *lo = (double)(v - (Q)*hi);
}

/* ELBO and terms at params (doubles, exact in binary128).  Outputs as (hi, lo) double pairs: value = hi + lo. */
int elbo_quad(int n, const double* t, const double* y, int nh, const double* th, int nx, const double* tx,
              const double* params, double reg, int causal, double elbo[2], double terms[14]) {
  const long np = 5 + nh + (long)nh * (nh + 1) / 2;
  Q* p = (Q*)malloc(sizeof(Q) * np);
  for (long i = 0; i < np; ++i) p[i] = params[i];
  Q e, tm[7];
  int rc = elbo_core(n, t, y, nh, th, nx, tx, p, reg, causal, &e, tm);
  free(p);
  if (rc) return rc;
  split(e, elbo, elbo + 1);
  for (int i = 0; i < 7; ++i) split(tm[i], terms + 2 * i, terms + 2 * i + 1);
  return 0;
}

/* Directional derivative d/ds ELBO(params + s dir) at s = 0 by central differences in binary128 with steps h and 2h:
 * out[0] = Richardson-extrapolated value (4 D(h) - D(2h)) / 3, out[1] = D(h), out[2] = D(2h). */
int elbo_quad_dderiv(int n, const double* t, const double* y, int nh, const double* th, int nx, const double* tx,
                     const double* params, const double* dir, double h, double reg, int causal, int richardson,
                     double out[3]) {
  const long np = 5 + nh + (long)nh * (nh + 1) / 2;
  Q* p = (Q*)malloc(sizeof(Q) * np);
  Q vals[4], tm[7];
  const Q steps[4] = {(Q)h, -(Q)h, 2 * (Q)h, -2 * (Q)h};
  const int nsteps = richardson ? 4 : 2;
  vals[2] = vals[3] = 0;
  for (int s = 0; s < nsteps; ++s) {
    for (long i = 0; i < np; ++i) p[i] = (Q)params[i] + steps[s] * (Q)dir[i];
    int rc = elbo_core(n, t, y, nh, th, nx, tx, p, reg, causal, &vals[s], tm);
    if (rc) { free(p); return rc; }
  }
  free(p);
  Q d1 = (vals[0] - vals[1]) / (2 * (Q)h), d2 = (vals[2] - vals[3]) / (4 * (Q)h);
  out[0] = richardson ? (double)((4 * d1 - d2) / 3) : (double)d1;
  out[1] = (double)d1;
  out[2] = richardson ? (double)d2 : (double)d1;
  return 0;
}

/* element-wise checks for the tests */
void bvn_quad(const double* x, const double* y, const double* rho, long n, double* out) {
  gl_init();
  for (long i = 0; i < n; ++i) out[i] = (double)bvn_cdf_q(x[i], y[i], rho[i]);
}
