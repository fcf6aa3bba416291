// K4 — winner refit + acceptance (sm_100a).  One warp per problem.
//
// Replaces, for the winning hypothesis only: Calibrator._match_bijective
// (calibrator.py:317-380), the inlier selection (:640-643),
// models.robust_polyfit (models.py:70-123) and the clipped-residual RMS /
// minimum_fit_error test (:664-679).
//
// robust_polyfit minimises the Huber loss (f_scale = 1) of the residuals of a
// polynomial in x/std(x) against y/std(y) with SciPy's Trust Region Reflective
// solver from x0 = ones on a forward-difference Jacobian.  Two modes:
//   mode 0 (default, what the facade uses): the same iteration, step for step
//       (scipy 1.18 optimize/_lsq/trf.py::trf_no_bounds, tr_solver='exact',
//       common.py::solve_lsq_trust_region / update_tr_radius /
//       check_termination, _numdiff '2-point' with rel_step 1e-5), with a
//       one-sided Jacobi SVD in place of LAPACK gesdd.  TRF stops 1e-8..1e-5
//       (relative, coefficients) short of the optimum and that distance is
//       chaotic in the last bits of the iterates (FD-Jacobian noise 1e-11 x
//       cond(J) 1e3..1e5), so agreement with a given SciPy build is at that
//       level too - see DESIGN.md; what is reproduced is the algorithm, its
//       number of function evaluations and the size of its inexactness (which
//       the reference's minimum_fit_error test on exact data relies on).
//   mode 1: the exact Huber optimum (Householder QR least squares + IRLS when a
//       normalised residual leaves the quadratic zone).
#include "common.cuh"
#include <stdio.h>

namespace rb2 {

constexpr int NCF = RB2_MAX_DEG + 1;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// least squares min || diag(sw) (V a - y) ||, V[i][j] = x[i]^j ; A is m x (n+1) scratch (row stride n+1)
__device__ void warp_qr_solve(const double* x, const double* y, const double* sw, int m, int n, double* A, double* a) {
  const int lane = lane_id();
  const int ld = n + 1;
  for (int i = lane; i < m; i += 32) {
    const double w = sw ? sw[i] : 1.0;
    double pw = w;
    for (int j = 0; j < n; ++j) { A[i * ld + j] = pw; pw *= x[i]; }
    A[i * ld + n] = w * y[i];
  }
  __syncwarp();
  for (int k = 0; k < n; ++k) {
    double part = 0.0;
    for (int i = k + 1 + lane; i < m; i += 32) part = fma(A[i * ld + k], A[i * ld + k], part);
    const double tail2 = warp_sum(part);
    const double akk = A[k * ld + k];
    const double nrm = sqrt(fma(akk, akk, tail2));
    if (nrm == 0.0) continue;
    const double alpha = (akk > 0.0) ? -nrm : nrm;
    const double v0 = akk - alpha;
    const double vv = fma(v0, v0, tail2);
    if (vv == 0.0) continue;
    const double beta = 2.0 / vv;
    for (int j = k + 1; j <= n; ++j) {
      double p = 0.0;
      for (int i = k + 1 + lane; i < m; i += 32) p = fma(A[i * ld + k], A[i * ld + j], p);
      const double s = (warp_sum(p) + v0 * A[k * ld + j]) * beta;
      __syncwarp();
      if (lane == 0) A[k * ld + j] -= s * v0;
      for (int i = k + 1 + lane; i < m; i += 32) A[i * ld + j] = fma(-s, A[i * ld + k], A[i * ld + j]);
      __syncwarp();
    }
    if (lane == 0) A[k * ld + k] = alpha;
    __syncwarp();
  }
  // back substitution (every lane computes the same values)
  for (int k = n - 1; k >= 0; --k) {
    double s = A[k * ld + n];
    for (int j = k + 1; j < n; ++j) s = fma(-A[k * ld + j], a[j], s);
    a[k] = s / A[k * ld + k];
  }
  __syncwarp();
}


// ---- SciPy TRF (no bounds, exact tr_solver, Huber loss), one warp ------------
constexpr double TRF_EPS = 2.220446049250313e-16;

struct TrfWs {        // shared-memory work space of one problem (m points, n params)
  double *f_true, *f, *f_new, *js;   // [m]
  double *J, *U;                     // [m*n] column-major
};

// models.py:10-54: y - (a[n-1] + sum_i a[n-1-i] * x**i), each power formed separately
// (NT > 0: the number of coefficients is a compile-time constant - the loops unroll and the small vectors of the
//  solver live in registers instead of local memory; NT = 0: run-time n, degrees 6 and 7)
template <int NT>
__device__ __forceinline__ double trf_residual(const double* a, int n_, double x, double y) {
  const int n = NT > 0 ? NT : n_;
  double t = a[n - 1];
  double pw = x;
#pragma unroll
  for (int i = 1; i < n; ++i) {
    t = add(t, mul(a[n - 1 - i], pw));
    pw = mul(pw, x);
  }
  return sub(y, t);
}

__device__ __forceinline__ void huber_rho(double f, double& r0, double& r1, double& r2) {
  const double z = f * f;
  if (z <= 1.0) { r0 = z; r1 = 1.0; r2 = 0.0; }
  else { const double sq = sqrt(z); r0 = 2.0 * sq - 1.0; r1 = 1.0 / sq; r2 = -0.5 / (z * sq); }
}

// f_true (filled) -> J (2-point), robust scaling of J and f, returns cost; g = J^T f
template <int NT>
__device__ double trf_linearise(const double* a, int n_, int m, const double* xn, const double* yn, TrfWs& w, double* g) {
  const int n = NT > 0 ? NT : n_;
  const int lane = lane_id();
#pragma unroll
  for (int j = 0; j < n; ++j) {  // _numdiff._compute_absolute_step / _dense_difference
    const double sign = (a[j] >= 0.0) ? 1.0 : -1.0;
    double h = 1e-5 * sign * fabs(a[j]);
    if (sub(add(a[j], h), a[j]) == 0.0) h = sqrt(TRF_EPS) * sign * fmax(1.0, fabs(a[j]));
    const double dx = sub(add(a[j], h), a[j]);
    double a1[NCF];
#pragma unroll
    for (int k = 0; k < n; ++k) a1[k] = (k == j) ? add(a[j], h) : a[k];
    for (int i = lane; i < m; i += 32) w.J[j * m + i] = sub(trf_residual<NT>(a1, n, xn[i], yn[i]), w.f_true[i]) / dx;
  }
  double c = 0.0;
  for (int i = lane; i < m; i += 32) {  // common.scale_for_robust_loss_function
    const double f = w.f_true[i];
    double r0, r1, r2;
    huber_rho(f, r0, r1, r2);
    double s = r1 + 2.0 * r2 * f * f;
    if (s < TRF_EPS) s = TRF_EPS;
    s = sqrt(s);
    w.js[i] = s;
    w.f[i] = f * (r1 / s);
    c += r0;
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < n; ++j) {
    double gj = 0.0;
    for (int i = lane; i < m; i += 32) {
      const double v = w.J[j * m + i] * w.js[i];
      w.J[j * m + i] = v;
      gj = fma(v, w.f[i], gj);
    }
    g[j] = warp_sum(gj);
  }
  __syncwarp();
  return 0.5 * warp_sum(c);
}

// one-sided Jacobi SVD of w.J (copied to w.U): U columns normalised in place, s descending, V (n x n,
// shared memory).  A pair step is latency-bound on one warp, so it is kept short: the three inner
// products share one interleaved butterfly, the rotation comes from one sqrt, one division and one
// rsqrt (t = sign * |h| / (|d| + sqrt(d^2 + h^2)), d = be - al, h = 2 ga), lanes update the rows of U
// and lane k row k of V.
#ifdef RB2_K4_PROFILE
__device__ int g_k4_rot = 0, g_k4_sweeps = 0;
#endif
// `warm`: V holds the right singular vectors of the previous Jacobian of the same problem; the
// iteration then starts from U = J V (columns already nearly orthogonal: the Jacobian changes little
// between trust-region steps), which saves most sweeps.  U S V^T is the same decomposition either way.
template <int NT>
__device__ void trf_svd(int n_, int m, TrfWs& w, double* s, double (*V)[NCF], bool warm) {
  const int n = NT > 0 ? NT : n_;
  const int lane = lane_id();
  if (!warm) {
    for (int i = lane; i < m * n; i += 32) w.U[i] = w.J[i];
    for (int i = lane; i < n * n; i += 32) V[i / n][i % n] = (i / n == i % n) ? 1.0 : 0.0;
  } else {
    for (int i = lane; i < m; i += 32) {
#pragma unroll
      for (int j = 0; j < n; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < n; ++k) acc = fma(w.J[k * m + i], V[k][j], acc);
        w.U[j * m + i] = acc;
      }
    }
  }
  __syncwarp();
  // (the pair loops are NOT unrolled: nothing in a pair step needs p or q at compile time, and ten copies of
  //  its ~300 instructions - the kernel is one warp - left it waiting for instruction fetches 44 % of the time)
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
#pragma unroll 1
    for (int p = 0; p < n - 1; ++p) {
#pragma unroll 1
      for (int q = p + 1; q < n; ++q) {
        double al = 0.0, be = 0.0, ga = 0.0;
        for (int i = lane; i < m; i += 32) {
          const double up = w.U[p * m + i], uq = w.U[q * m + i];
          al = fma(up, up, al); be = fma(uq, uq, be); ga = fma(up, uq, ga);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          al += __shfl_xor_sync(0xffffffffu, al, o);
          be += __shfl_xor_sync(0xffffffffu, be, o);
          ga += __shfl_xor_sync(0xffffffffu, ga, o);
        }
        // converged pair: |ga| <= eps * sqrt(al * be), tested on the squares
        if (fabs(ga) <= 1e-300 || ga * ga <= (TRF_EPS * TRF_EPS) * (al * be)) continue;
        rotated = true;
#ifdef RB2_K4_PROFILE
        ++g_k4_rot;
#endif
        const double dd = be - al, hh = 2.0 * ga;
        const double t = copysign(fabs(hh) / (fabs(dd) + sqrt(fma(dd, dd, hh * hh))),
                                  (dd == 0.0 || (dd > 0.0) == (hh > 0.0)) ? 1.0 : -1.0);
        const double c = rsqrt(fma(t, t, 1.0)), sn = c * t;
        for (int i = lane; i < m; i += 32) {
          const double up = w.U[p * m + i], uq = w.U[q * m + i];
          w.U[p * m + i] = c * up - sn * uq;
          w.U[q * m + i] = sn * up + c * uq;
        }
        if (lane < n) {
          const double vp = V[lane][p], vq = V[lane][q];
          V[lane][p] = c * vp - sn * vq;
          V[lane][q] = sn * vp + c * vq;
        }
        __syncwarp();
      }
    }
#ifdef RB2_K4_PROFILE
    ++g_k4_sweeps;
#endif
    if (!rotated) break;
  }
#pragma unroll
  for (int j = 0; j < n; ++j) {
    double ss = 0.0;
    for (int i = lane; i < m; i += 32) ss = fma(w.U[j * m + i], w.U[j * m + i], ss);
    s[j] = sqrt(warp_sum(ss));
  }
  // selection sort, descending (stable enough: n <= 8), permuting U columns and V columns
#pragma unroll
  for (int a = 0; a < n - 1; ++a) {
    int b = a;
    double sb = s[a];
#pragma unroll
    for (int k = 0; k < NCF; ++k) if (k > a && k < n && s[k] > sb) { b = k; sb = s[k]; }
    if (b != a) {
#pragma unroll
      for (int k = 0; k < NCF; ++k) if (k == b) s[k] = s[a];
      s[a] = sb;
      for (int i = lane; i < m; i += 32) { const double tu = w.U[a * m + i]; w.U[a * m + i] = w.U[b * m + i]; w.U[b * m + i] = tu; }
      if (lane < n) { const double tv = V[lane][a]; V[lane][a] = V[lane][b]; V[lane][b] = tv; }
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < n; ++j) {
    const double inv = (s[j] > 0.0) ? 1.0 / s[j] : 0.0;
    for (int i = lane; i < m; i += 32) w.U[j * m + i] *= inv;
  }
  __syncwarp();
}

template <int NT = 0>
__device__ __forceinline__ double vnorm(const double* v, int n_) {
  const int n = NT > 0 ? NT : n_;
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < n; ++i) s = fma(v[i], v[i], s);
  return sqrt(s);
}

// common.solve_lsq_trust_region (scalar: every lane computes the same values)
template <int NT>
__device__ void trf_solve_tr(int n_, int m, const double* uf, const double* s, const double (*V)[NCF], double Delta,
                             double& alpha, double* p) {
  const int n = NT > 0 ? NT : n_;
  double suf[NCF];
#pragma unroll
  for (int i = 0; i < n; ++i) suf[i] = s[i] * uf[i];
  auto phi_and_derivative = [&](double al, double& phi, double& phi_prime) {
    double pn = 0.0, sm = 0.0;
#pragma unroll
    for (int i = 0; i < NCF; ++i) {
      if (i < n) {
        const double r = 1.0 / (s[i] * s[i] + al);  // one division per term: q = suf / den, suf^2 / den^3 = q^2 / den
        const double q = suf[i] * r;
        pn = fma(q, q, pn);
        sm = fma(q * q, r, sm);
      }
    }
    pn = sqrt(pn);
    phi = pn - Delta;
    phi_prime = -sm / pn;
  };
  const bool full_rank = (m >= n) && (s[n - 1] > TRF_EPS * m * s[0]);
  if (full_rank) {
    double gn[NCF];  // Gauss-Newton step in the singular basis
#pragma unroll
    for (int i = 0; i < n; ++i) gn[i] = uf[i] / s[i];
#pragma unroll
    for (int k = 0; k < n; ++k) {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < n; ++i) acc = fma(V[k][i], gn[i], acc);
      p[k] = -acc;
    }
    if (vnorm<NT>(p, n) <= Delta) { alpha = 0.0; return; }
  }
  double alpha_upper = vnorm<NT>(suf, n) / Delta, alpha_lower = 0.0;
  if (full_rank) {
    double phi, phip;
    phi_and_derivative(0.0, phi, phip);
    alpha_lower = -phi / phip;
  }
  if (!full_rank && alpha == 0.0) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
  for (int it = 0; it < 10; ++it) {
    if (alpha < alpha_lower || alpha > alpha_upper) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    double phi, phip;
    phi_and_derivative(alpha, phi, phip);
    if (phi < 0.0) alpha_upper = alpha;
    const double ratio = phi / phip;
    alpha_lower = fmax(alpha_lower, alpha - ratio);
    alpha -= (phi + Delta) * ratio / Delta;
    if (fabs(phi) < 0.01 * Delta) break;
  }
  double lm[NCF];  // Levenberg-Marquardt step in the singular basis
#pragma unroll
  for (int i = 0; i < n; ++i) lm[i] = suf[i] / (s[i] * s[i] + alpha);
#pragma unroll
  for (int k = 0; k < n; ++k) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) acc = fma(V[k][i], lm[i], acc);
    p[k] = -acc;
  }
  const double sc = Delta / vnorm<NT>(p, n);
#pragma unroll
  for (int k = 0; k < n; ++k) p[k] *= sc;
}

// trf.trf_no_bounds; a[] = solution (high order first, as models.polynomial expects)
template <int NT>
__device__ void trf_huber_warp(int n_, int m, const double* xn, const double* yn, TrfWs& w, double* a, int& nfev_out,
                               int& status_out) {
  const int n = NT > 0 ? NT : n_;
  const int lane = lane_id();
  const double ftol = 1e-8, xtol = 1e-8, gtol = 1e-8;
  _Pragma("unroll") for (int k = 0; k < n; ++k) a[k] = 1.0;
  for (int i = lane; i < m; i += 32) w.f_true[i] = trf_residual<NT>(a, n, xn[i], yn[i]);
  __syncwarp();
  int nfev = 1;
  double g[NCF];
  double cost = trf_linearise<NT>(a, n, m, xn, yn, w, g);
  double Delta = vnorm<NT>(a, n);
  if (Delta == 0.0) Delta = 1.0;
  const int max_nfev = 100 * n;
  double alpha = 0.0;
  int status = -1;
  double s[NCF], uf[NCF], step[NCF], a_new[NCF];
  int svd_count = 0;
  __shared__ double V[NCF][NCF];  // one warp per CTA in every kernel that runs the solver
#ifdef RB2_K4_PROFILE
  long long t_svd = 0, t_lin = 0, t_inner = 0, t0c = 0;
  int n_outer = 0, n_inner = 0;
#define K4_TIC() t0c = clock64()
#define K4_TOC(acc) acc += clock64() - t0c
#else
#define K4_TIC()
#define K4_TOC(acc)
#endif
  for (;;) {
    double g_norm = 0.0;
    _Pragma("unroll") for (int k = 0; k < n; ++k) g_norm = fmax(g_norm, fabs(g[k]));
    if (g_norm < gtol) status = 1;
    if (status >= 0 || nfev == max_nfev) break;
    K4_TIC();
    trf_svd<NT>(n, m, w, s, V, svd_count++ > 0);
    _Pragma("unroll") for (int j = 0; j < n; ++j) {
      double acc = 0.0;
      for (int i = lane; i < m; i += 32) acc = fma(w.U[j * m + i], w.f[i], acc);
      uf[j] = warp_sum(acc);
    }
    K4_TOC(t_svd);
#ifdef RB2_K4_PROFILE
    ++n_outer;
#endif
    K4_TIC();
    double actual_reduction = -1.0, cost_new = cost;
    while (actual_reduction <= 0.0 && nfev < max_nfev) {
      trf_solve_tr<NT>(n, m, uf, s, V, Delta, alpha, step);
      double q = 0.0;  // evaluate_quadratic
      for (int i = lane; i < m; i += 32) {
        double js = 0.0;
        _Pragma("unroll") for (int k = 0; k < n; ++k) js = fma(w.J[k * m + i], step[k], js);
        q = fma(js, js, q);
      }
      q = warp_sum(q);
      double l = 0.0;
      _Pragma("unroll") for (int k = 0; k < n; ++k) l = fma(step[k], g[k], l);
      const double predicted_reduction = -(0.5 * q + l);
      _Pragma("unroll") for (int k = 0; k < n; ++k) a_new[k] = a[k] + step[k];
      double c = 0.0;
      int bad = 0;
      for (int i = lane; i < m; i += 32) {
        const double fn = trf_residual<NT>(a_new, n, xn[i], yn[i]);
        w.f_new[i] = fn;
        bad |= !isfinite(fn);
        double r0, r1, r2;
        huber_rho(fn, r0, r1, r2);
        c += r0;
      }
      ++nfev;
      const double step_norm = vnorm<NT>(step, n);
      if (__any_sync(0xffffffffu, bad)) { Delta = 0.25 * step_norm; continue; }
      cost_new = 0.5 * warp_sum(c);
      actual_reduction = cost - cost_new;
      double ratio;
      if (predicted_reduction > 0.0) ratio = actual_reduction / predicted_reduction;
      else if (predicted_reduction == 0.0 && actual_reduction == 0.0) ratio = 1.0;
      else ratio = 0.0;
      double Delta_new = Delta;
      if (ratio < 0.25) Delta_new = 0.25 * step_norm;
      else if (ratio > 0.75 && step_norm > 0.95 * Delta) Delta_new = Delta * 2.0;
      const bool ftol_ok = (actual_reduction < ftol * cost) && (ratio > 0.25);
      const bool xtol_ok = step_norm < xtol * (xtol + vnorm<NT>(a, n));
      if (ftol_ok && xtol_ok) status = 4; else if (ftol_ok) status = 2; else if (xtol_ok) status = 3;
      if (status >= 0) break;
      alpha *= Delta / Delta_new;
      Delta = Delta_new;
    }
    K4_TOC(t_inner);
    if (actual_reduction > 0.0) {
      _Pragma("unroll") for (int k = 0; k < n; ++k) a[k] = a_new[k];
      __syncwarp();
      for (int i = lane; i < m; i += 32) w.f_true[i] = w.f_new[i];
      __syncwarp();
      cost = cost_new;
      K4_TIC();
      trf_linearise<NT>(a, n, m, xn, yn, w, g);
      K4_TOC(t_lin);
    }
  }
#ifdef RB2_K4_PROFILE
  if (lane == 0) printf("K4 profile: m %d n %d nfev %d outer %d  cycles svd %lld inner %lld linearise %lld  sweeps %d rotations %d (all lanes x32)\n", m, n, nfev, n_outer, t_svd, t_inner, t_lin, g_k4_sweeps, g_k4_rot);
#endif
  nfev_out = nfev;
  status_out = status < 0 ? 0 : status;
}

// models.robust_polyfit (models.py:70-123) on m points (ix, iy) in shared memory, one warp.
// xn, yn, sw: [m] scratch; A: [(2n+3)*m] scratch.  Fills r.coeffs (low order first),
// r.huber_iters (nfev / IRLS passes) and r.pad (SciPy status).
__device__ void robust_polyfit_warp(int m, int n, const double* ix, const double* iy, double* xn, double* yn, double* sw,
                                    double* A, int mode, rb2_refit_result& r) {
  const int lane = lane_id();
  // models.normalise_input: population standard deviations
  double sx = 0.0, sy = 0.0;
  for (int i = lane; i < m; i += 32) { sx += ix[i]; sy += iy[i]; }
  const double mxm = warp_sum(sx) / m, mym = warp_sum(sy) / m;
  double vx = 0.0, vy = 0.0;
  for (int i = lane; i < m; i += 32) {
    const double dx = ix[i] - mxm, dy = iy[i] - mym;
    vx = fma(dx, dx, vx); vy = fma(dy, dy, vy);
  }
  const double xs = sqrt(warp_sum(vx) / m), ys = sqrt(warp_sum(vy) / m);
  for (int i = lane; i < m; i += 32) { xn[i] = ix[i] / xs; yn[i] = iy[i] / ys; }
  __syncwarp();
  double a[NCF];
  for (int i = 0; i < NCF; ++i) a[i] = 0.0;
  int iters = 0;
  r.pad = 0;
  if (mode == 0) {
    TrfWs ws;
    ws.J = A; ws.U = A + (size_t)m * n; ws.f_true = A + (size_t)2 * m * n;
    ws.f = ws.f_true + m; ws.f_new = ws.f + m; ws.js = sw;
    double ah[NCF];
    int nfev = 0, status = 0;
    switch (n) {  // compile-time coefficient counts for the degrees in use (1..5)
      case 2: trf_huber_warp<2>(n, m, xn, yn, ws, ah, nfev, status); break;
      case 3: trf_huber_warp<3>(n, m, xn, yn, ws, ah, nfev, status); break;
      case 4: trf_huber_warp<4>(n, m, xn, yn, ws, ah, nfev, status); break;
      case 5: trf_huber_warp<5>(n, m, xn, yn, ws, ah, nfev, status); break;
      case 6: trf_huber_warp<6>(n, m, xn, yn, ws, ah, nfev, status); break;
      default: trf_huber_warp<0>(n, m, xn, yn, ws, ah, nfev, status); break;
    }
    for (int j = 0; j < n; ++j) a[j] = ah[n - 1 - j];  // low order first
    iters = nfev;
    r.pad = status;
  } else {
    warp_qr_solve(xn, yn, nullptr, m, n, A, a);
    // Huber (f_scale = 1): leave the quadratic zone only if some |r| > 1
    for (; iters < 100; ++iters) {
      int nout = 0;
      for (int i = lane; i < m; i += 32) {
        double f = a[n - 1];
        for (int j = n - 2; j >= 0; --j) f = fma(f, xn[i], a[j]);
        const double rr = fabs(yn[i] - f);
        sw[i] = (rr <= 1.0) ? 1.0 : sqrt(1.0 / rr);
        nout += (rr > 1.0);
      }
      nout = __reduce_add_sync(0xffffffffu, nout);
      __syncwarp();
      if (nout == 0 && iters == 0) break;
      double prev[NCF];
      for (int i = 0; i < n; ++i) prev[i] = a[i];
      warp_qr_solve(xn, yn, sw, m, n, A, a);
      double dmax = 0.0, amax = 0.0;
      for (int i = 0; i < n; ++i) { dmax = fmax(dmax, fabs(a[i] - prev[i])); amax = fmax(amax, fabs(a[i])); }
      if (dmax <= 1e-15 * amax) { ++iters; break; }
    }
  }
  r.huber_iters = iters;
  // un-normalise (models.py:117-123), low order first
  for (int j = 0; j < n; ++j) r.coeffs[j] = (j == 0) ? a[j] * ys : (a[j] * ys) / pow(xs, (double)j);
}

__global__ void __launch_bounds__(32)
refit_kernel(const rb2_ransac_desc* __restrict__ descs, const rb2_ransac_result* __restrict__ winners,
             double min_fit_error, int mode, rb2_refit_result* __restrict__ out, double* __restrict__ mx_out,
             double* __restrict__ my_out, double* __restrict__ res_out, int max_up) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = blockIdx.x, lane = threadIdx.x;
  const rb2_ransac_desc& d = descs[p];
  const rb2_ransac_result& w = winners[p];
  const int n_u = d.n_upeaks, n_w = d.n_wids, deg = d.deg, n = deg + 1;
  rb2_refit_result r;
  for (int i = 0; i < NCF; ++i) r.coeffs[i] = 0.0;
  r.rms = 0.0; r.n_inliers = 0; r.accepted = 0; r.huber_iters = 0; r.pad = 0;
  if (w.hyp_id < 0) {
    if (lane == 0) out[p] = r;
    return;
  }
  unsigned long long* berr = reinterpret_cast<unsigned long long*>(smem_raw);  // [n_w]
  double* bwave = reinterpret_cast<double*>(berr + n_w);                         // [n_w]
  double* ix = bwave + n_w;                                                      // [n_u] inlier x
  double* iy = ix + n_u;                                                         // [n_u] inlier y
  double* xn = iy + n_u;                                                         // [n_u]
  double* yn = xn + n_u;                                                         // [n_u]
  double* sw = yn + n_u;                                                         // [n_u]
  double* A = sw + n_u;                                                          // [n_u*(n+1)] (QR) / J, U, f_true, f, f_new (TRF)
  int* bpeak = reinterpret_cast<int*>(A + (size_t)n_u * (2 * n + 3));            // [n_w]

  double c[NCF];
  for (int i = 0; i < NCF; ++i) c[i] = w.coeffs[i];
  for (int i = lane; i < n_w; i += 32) { berr[i] = 0x7ff0000000000000ull; bpeak[i] = 0x7fffffff; }
  __syncwarp();
  // nearest candidate per peak (first minimum), best error per wavelength id
  for (int u = lane; u < n_u; u += 32) {
    const double fit = horner_unfused(c, n, d.d_peaks[u]);
    double be = CUDART_INF; int bw = -1;
    for (int k = d.d_cand_off[u]; k < d.d_cand_off[u + 1]; ++k) {
      const double e = fabs(sub(fit, d.d_cand_wave[k]));
      if (e < be) { be = e; bw = d.d_cand_wid[k]; }
    }
    if (bw >= 0) atomicMin(&berr[bw], (unsigned long long)__double_as_longlong(be));
  }
  __syncwarp();
  // ties inside a wavelength group: first peak in ascending order (np.argmin)
  for (int u = lane; u < n_u; u += 32) {
    const double fit = horner_unfused(c, n, d.d_peaks[u]);
    double be = CUDART_INF; int bw = -1, bk = -1;
    for (int k = d.d_cand_off[u]; k < d.d_cand_off[u + 1]; ++k) {
      const double e = fabs(sub(fit, d.d_cand_wave[k]));
      if (e < be) { be = e; bw = d.d_cand_wid[k]; bk = k; }
    }
    if (bw >= 0 && (unsigned long long)__double_as_longlong(be) == berr[bw]) {
      atomicMin(&bpeak[bw], u);
      bwave[bw] = d.d_cand_wave[bk];  // same value for every writer (same wavelength id)
    }
  }
  __syncwarp();
  // inliers (err < tol), ascending wavelength = ascending id, order-preserving compaction
  int m = 0;
  for (int i0 = 0; i0 < n_w; i0 += 32) {
    const int i = i0 + lane;
    bool in = false;
    if (i < n_w) {
      const unsigned long long bits = berr[i];
      in = (bits != 0x7ff0000000000000ull) && (__longlong_as_double((long long)bits) < d.tol);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    if (in) {
      const int pos = m + __popc(bal & ((1u << lane) - 1u));
      ix[pos] = d.d_peaks[bpeak[i]];
      iy[pos] = bwave[i];
    }
    m += __popc(bal);
  }
  __syncwarp();
  r.n_inliers = m;
  if (m <= deg) {  // calibrator.py:645-649
    if (lane == 0) out[p] = r;
    return;
  }
  robust_polyfit_warp(m, n, ix, iy, xn, yn, sw, A, mode, r);
  // residual = polyval(matched_peaks, coeffs) - matched_atlas, clipped (:664-671)
  double ss = 0.0;
  for (int i = lane; i < m; i += 32) {
    double res = sub(basis_val(d.basis, r.coeffs, n, ix[i]), iy[i]);  // self.polyval on the refit coefficients (:664)
    if (fabs(res) > d.tol) res = d.tol;
    ss = fma(res, res, ss);
    mx_out[(size_t)p * max_up + i] = ix[i];
    my_out[(size_t)p * max_up + i] = iy[i];
    res_out[(size_t)p * max_up + i] = res;
  }
  r.rms = sqrt(warp_sum(ss) / m);
  const double mfe = (min_fit_error == min_fit_error) ? min_fit_error : d.min_fit_error;  // NaN: per problem
  r.accepted = (r.rms < mfe) ? 0 : 1;
  if (lane == 0) out[p] = r;
}

// models.robust_polyfit for arbitrary point sets (Calibrator.match_peaks :1915-1920, manual_refit):
// problem p owns points [off[p], off[p+1]).
__global__ void __launch_bounds__(32)
robust_polyfit_kernel(const int* __restrict__ off, const double* __restrict__ x, const double* __restrict__ y, int deg,
                      int mode, int max_m, rb2_refit_result* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = blockIdx.x, lane = threadIdx.x;
  const int o0 = off[p], m = off[p + 1] - o0, n = deg + 1;
  double* ix = reinterpret_cast<double*>(smem_raw);
  double* iy = ix + max_m;
  double* xn = iy + max_m;
  double* yn = xn + max_m;
  double* sw = yn + max_m;
  double* A = sw + max_m;
  rb2_refit_result r;
  for (int i = 0; i < NCF; ++i) r.coeffs[i] = 0.0;
  r.rms = 0.0; r.n_inliers = m; r.accepted = 0; r.huber_iters = 0; r.pad = 0;
  if (m <= deg) {
    if (lane == 0) out[p] = r;
    return;
  }
  for (int i = lane; i < m; i += 32) { ix[i] = x[o0 + i]; iy[i] = y[o0 + i]; }
  __syncwarp();
  robust_polyfit_warp(m, n, ix, iy, xn, yn, sw, A, mode, r);
  double ss = 0.0;
  for (int i = lane; i < m; i += 32) {
    const double res = sub(iy[i], horner_unfused(r.coeffs, n, ix[i]));
    ss = fma(res, res, ss);
  }
  r.rms = sqrt(warp_sum(ss) / m);
  r.accepted = 1;
  if (lane == 0) out[p] = r;
}

// K5 - Calibrator.match_peaks (calibrator.py:1824-1956, refine=False), one warp per problem.
__global__ void __launch_bounds__(32)
match_peaks_kernel(const rb2_match_desc* __restrict__ descs, int mode, int max_p, int max_n,
                   rb2_match_result* __restrict__ out, double* __restrict__ mx_out, double* __restrict__ my_out,
                   double* __restrict__ res_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int p = blockIdx.x, lane = threadIdx.x;
  const rb2_match_desc& d = descs[p];
  const int n_p = d.n_peaks, n_a = d.n_atlas, n = d.fit_deg + 1;
  double* ix = reinterpret_cast<double*>(smem_raw);
  double* iy = ix + max_p;
  double* xn = iy + max_p;
  double* yn = xn + max_p;
  double* sw = yn + max_p;
  double* A = sw + max_p;  // (2 max_n + 3) * max_p
  (void)max_n;
  rb2_match_result r;
  r.n_matched = 0; r.n_ambiguous = 0; r.polyfit_err = 0.0; r.polyfit_rms = 0.0;
  for (int i = 0; i < NCF; ++i) { r.polyfit_coeffs[i] = 0.0; r.refit.coeffs[i] = 0.0; }
  r.refit.rms = 0.0; r.refit.n_inliers = 0; r.refit.accepted = 0; r.refit.huber_iters = 0; r.refit.pad = 0;
  double cin[NCF];
  for (int i = 0; i < NCF; ++i) cin[i] = (i < d.n_coef) ? (d.d_coeffs ? d.d_coeffs[i] : d.coeffs[i]) : 0.0;
  // -- association (:1830-1838): lanes take peaks, every lane scans the atlas in order --
  int m = 0, n_amb = 0;
  for (int base = 0; base < n_p; base += 32) {
    const int u = base + lane;
    int hits = 0;
    double bd = CUDART_INF, bw = 0.0, px = 0.0;
    if (u < n_p) {
      px = d.d_peaks[u];
      const double x = basis_val(d.basis, cin, d.n_coef, px);
      for (int k = 0; k < n_a; ++k) {
        const double w = d.d_atlas[k];
        const double ad = fabs(sub(w, x));
        if (ad < d.tolerance) {
          ++hits;
          if (ad < bd) { bd = ad; bw = w; }  // nearest, first on ties
        }
      }
    }
    const unsigned hm = __ballot_sync(0xffffffffu, hits > 0);
    n_amb += __popc(__ballot_sync(0xffffffffu, hits > 1));
    if (hits > 0) {
      const int slot = m + __popc(hm & ((1u << lane) - 1u));
      ix[slot] = px; iy[slot] = bw;
    }
    m += __popc(hm);
  }
  __syncwarp();
  r.n_matched = m; r.n_ambiguous = n_amb;
  for (int i = lane; i < m; i += 32) { mx_out[(size_t)p * max_p + i] = ix[i]; my_out[(size_t)p * max_p + i] = iy[i]; }
  if (m < n) {  // np.polyfit of fewer points than coefficients is rank deficient: reported, not fitted
    if (lane == 0) out[p] = r;
    return;
  }
  // -- polyfit of the association (:1893): least squares in the normalised variables --
  {
    double sx = 0.0, sy = 0.0;
    for (int i = lane; i < m; i += 32) { sx += ix[i]; sy += iy[i]; }
    const double mxm = warp_sum(sx) / m, mym = warp_sum(sy) / m;
    double vx = 0.0, vy = 0.0;
    for (int i = lane; i < m; i += 32) {
      const double dx = ix[i] - mxm, dy = iy[i] - mym;
      vx = fma(dx, dx, vx); vy = fma(dy, dy, vy);
    }
    double xs = sqrt(warp_sum(vx) / m), ys = sqrt(warp_sum(vy) / m);
    if (!(xs > 0.0)) xs = 1.0;
    if (!(ys > 0.0)) ys = 1.0;
    for (int i = lane; i < m; i += 32) { xn[i] = ix[i] / xs; yn[i] = iy[i] / ys; }
    __syncwarp();
    double a[NCF];
    for (int i = 0; i < NCF; ++i) a[i] = 0.0;
    warp_qr_solve(xn, yn, nullptr, m, n, A, a);
    double sc = ys;
    for (int j = 0; j < n; ++j) { r.polyfit_coeffs[j] = a[j] * sc; sc /= xs; }
    double se = 0.0, ss = 0.0;
    for (int i = lane; i < m; i += 32) {
      const double res = fabs(sub(iy[i], horner_unfused(r.polyfit_coeffs, n, ix[i])));
      if (!d.robust_refit) res_out[(size_t)p * max_p + i] = res;
      se += res; ss = fma(res, res, ss);
    }
    r.polyfit_err = warp_sum(se);
    r.polyfit_rms = sqrt(warp_sum(ss) / m);
  }
  if (d.robust_refit) {  // :1915-1941
    __syncwarp();
    robust_polyfit_warp(m, n, ix, iy, xn, yn, sw, A, mode, r.refit);
    double ss = 0.0;
    for (int i = lane; i < m; i += 32) {
      const double res = sub(iy[i], basis_val(d.basis, r.refit.coeffs, n, ix[i]));
      res_out[(size_t)p * max_p + i] = res;
      ss = fma(res, res, ss);
    }
    r.refit.rms = sqrt(warp_sum(ss) / m);
    r.refit.n_inliers = m;
    r.refit.accepted = 1;
  }
  if (lane == 0) out[p] = r;
}

// Calibrator._adjust_polyfit (calibrator.py:754-820), the objective of match_peaks(refine=True): one warp per
// trial shift `delta` of the first n_delta coefficients.
__global__ void __launch_bounds__(32)
adjust_polyfit_kernel(const double* __restrict__ deltas, int n_delta, const double* __restrict__ fit, int n_coef, int basis,
                      const double* __restrict__ peaks, int n_peaks, const double* __restrict__ atlas, int n_atlas,
                      const double* __restrict__ pix_sorted, int n_pix, double tolerance, double min_frac,
                      double* __restrict__ out) {
  const int e = blockIdx.x, lane = threadIdx.x;
  double c[NCF];
  for (int i = 0; i < NCF; ++i) c[i] = (i < n_coef) ? fit[i] : 0.0;
  for (int i = 0; i < n_delta && i < n_coef; ++i) c[i] = add(c[i], deltas[(size_t)e * n_delta + i]);  // fit_new[i] += d (:786-787)
  // nearest atlas line per peak (np.argmin: first minimum), kept if closer than the tolerance (:789-797)
  int n_m = 0;
  double ss = 0.0;
  for (int u = lane; u < n_peaks; u += 32) {
    const double p = peaks[u];
    const double x = basis_val(basis, c, n_coef, p);
    double best = CUDART_INF, bw = 0.0;
    for (int k = 0; k < n_atlas; ++k) {
      const double da = fabs(sub(atlas[k], x));
      if (da < best) { best = da; bw = atlas[k]; }
    }
    if (best < tolerance) {
      ++n_m;
      const double r = sub(bw, x);  // y_matched - polyval(x_matched, fit_new) (:816)
      ss = add(ss, mul(r, r));
    }
  }
  n_m = __reduce_add_sync(0xffffffffu, n_m);
  ss = warp_sum(ss);
  // monotonic over the sorted pixel list (:809-813)
  int bad = 0;
  for (int i = lane; i + 1 < n_pix; i += 32)
    bad |= !(sub(basis_val(basis, c, n_coef, pix_sorted[i + 1]), basis_val(basis, c, n_coef, pix_sorted[i])) > 0.0);
  bad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    const int dof = n_m - n_coef - 1;  // :802
    double v;
    if (dof < 1 || (double)n_m < (double)n_peaks * min_frac || bad) v = CUDART_INF;
    else v = ss / (double)dof;
    out[e] = v;
  }
}

}  // namespace rb2

extern "C" int rb2_adjust_polyfit(int n_eval, const double* d_deltas, int n_delta, const double* d_fit, int n_coef, int basis,
                                  const double* d_peaks, int n_peaks, const double* d_atlas, int n_atlas,
                                  const double* d_pix_sorted, int n_pix, double tolerance, double min_frac, double* d_out,
                                  void* stream_) {
  using namespace rb2;
  if (n_eval <= 0 || !d_deltas || !d_fit || !d_out || n_delta < 0 || n_coef < 1 || n_coef > RB2_MAX_DEG + 1 || basis < 0 ||
      basis > 2 || n_peaks < 0 || n_atlas < 0 || n_pix < 0 || (n_peaks && !d_peaks) || (n_atlas && !d_atlas) ||
      (n_pix && !d_pix_sorted))
    return RB2_ERR_ARG;
  adjust_polyfit_kernel<<<n_eval, 32, 0, (cudaStream_t)stream_>>>(d_deltas, n_delta, d_fit, n_coef, basis, d_peaks, n_peaks,
                                                                 d_atlas, n_atlas, d_pix_sorted, n_pix, tolerance, min_frac,
                                                                 d_out);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

extern "C" int rb2_refit_winners(int n_problems, const rb2_ransac_desc* d_desc, const rb2_ransac_desc* h_desc,
                                 const rb2_ransac_result* d_winner, double minimum_fit_error, int mode,
                                 rb2_refit_result* d_out, double* d_matched_x, double* d_matched_y,
                                 double* d_residual, int max_upeaks, void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_desc || !h_desc || !d_winner || !d_out || mode < 0 || mode > 1) return RB2_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  size_t smem = 0;
  for (int i = 0; i < n_problems; ++i) {
    const rb2_ransac_desc& h = h_desc[i];
    if (h.n_upeaks > max_upeaks) return RB2_ERR_ARG;
    const size_t n = h.deg + 1;
    size_t b = (size_t)h.n_wids * 16 + (size_t)h.n_upeaks * 8 * 5 + (size_t)h.n_upeaks * (2 * n + 3) * 8 + (size_t)h.n_wids * 4;
    smem = max(smem, b);
  }
  if (smem > 200 * 1024) return RB2_ERR_ARG;
  RB2_CHECK(cudaFuncSetAttribute(refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  refit_kernel<<<n_problems, 32, smem, stream>>>(d_desc, d_winner, minimum_fit_error, mode, d_out, d_matched_x,
                                                 d_matched_y, d_residual, max_upeaks);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

extern "C" int rb2_robust_polyfit(int n_problems, const int32_t* d_off, const int32_t* h_off, const double* d_x,
                                  const double* d_y, int deg, int mode, rb2_refit_result* d_out, void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_off || !h_off || !d_x || !d_y || !d_out || deg < 1 || deg > RB2_MAX_DEG || mode < 0 || mode > 1)
    return RB2_ERR_ARG;
  int max_m = 1;
  for (int i = 0; i < n_problems; ++i) {
    if (h_off[i + 1] < h_off[i]) return RB2_ERR_ARG;
    max_m = max(max_m, h_off[i + 1] - h_off[i]);
  }
  const size_t smem = (size_t)max_m * 8 * (5 + 2 * (deg + 1) + 3);
  if (smem > 200 * 1024) return RB2_ERR_ARG;
  RB2_CHECK(cudaFuncSetAttribute(robust_polyfit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  robust_polyfit_kernel<<<n_problems, 32, smem, (cudaStream_t)stream_>>>(d_off, d_x, d_y, deg, mode, max_m, d_out);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

extern "C" int rb2_match_peaks(int n_problems, const rb2_match_desc* d_desc, const rb2_match_desc* h_desc, int mode,
                               rb2_match_result* d_out, double* d_matched_x, double* d_matched_y, double* d_residual,
                               int max_peaks, void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_desc || !h_desc || !d_out || !d_matched_x || !d_matched_y || !d_residual || max_peaks <= 0 ||
      mode < 0 || mode > 1)
    return RB2_ERR_ARG;
  int max_n = 2;
  for (int i = 0; i < n_problems; ++i) {
    const rb2_match_desc& h = h_desc[i];
    if (h.n_peaks < 0 || h.n_peaks > max_peaks || h.n_atlas < 0 || h.n_coef < 1 || h.n_coef > RB2_MAX_DEG + 1 ||
        h.fit_deg < 1 || h.fit_deg > RB2_MAX_DEG || !(h.tolerance >= 0.0) || h.basis < 0 || h.basis > 2)
      return RB2_ERR_ARG;
    if ((h.n_peaks && !h.d_peaks) || (h.n_atlas && !h.d_atlas)) return RB2_ERR_ARG;
    max_n = max(max_n, h.fit_deg + 1);
  }
  const size_t smem = (size_t)max_peaks * 8 * (5 + 2 * max_n + 3);
  if (smem > 200 * 1024) return RB2_ERR_ARG;
  RB2_CHECK(cudaFuncSetAttribute(match_peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  match_peaks_kernel<<<n_problems, 32, smem, (cudaStream_t)stream_>>>(d_desc, mode, max_peaks, max_n, d_out, d_matched_x,
                                                                       d_matched_y, d_residual);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}
