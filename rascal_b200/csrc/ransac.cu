// K3 — batched RANSAC hypothesis draw / fit / test / score / reduce (sm_100a).
//
// Replaces the loop body of Calibrator._solve_candidate_ransac
// (calibrator.py:525-735).  One hypothesis = one minimal sample drawn and
// fitted (= one polyfit call of the reference).
//
// Mapping.  The three very different cost classes of a hypothesis never share a warp:
// they are separate small kernels connected by compacted survivor queues (see "The staged
// pipeline" below): A draw + fit + intercept test (one thread per hypothesis; ~85 % end
// here), B monotonicity test (one thread per survivor: instead of the reference's
// G ~ 1.4 * ptp(peaks) point grid, locate the critical points of q(k) = p(t_k + 1) - p(t_k)
// and evaluate the reference's own Horner arithmetic only at the grid points around them),
// C score (one warp per survivor: lanes take peaks, shared-memory atomicMin de-duplicates by
// wavelength id; M-SAC cost; Hough-density weight by locating the threshold crossings of the
// tangent-intercept polynomial with warp-parallel 32-ary searches over the pixel list,
// SURVEY.md App. A.5, instead of evaluating a bicubic spline at every detector pixel).
// Selection rule (App. A.4): minimal cost among eligible hypotheses, the LATEST hypothesis id
// among equal costs - a 128-bit atomic min on (sortable cost, ~id) per problem.
#include "common.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace rb2 {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int NCOEF = RB2_MAX_DEG + 1;

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (id_lo, id_hi, block, 0) ----
struct Philox {
  unsigned k0, k1, c0, c1, blk;
  unsigned w[4];
  int pos;
  __device__ __forceinline__ void init(unsigned long long seed, unsigned long long id) {
    k0 = (unsigned)seed; k1 = (unsigned)(seed >> 32);
    c0 = (unsigned)id; c1 = (unsigned)(id >> 32);
    blk = 0; pos = 4;
  }
  __device__ __forceinline__ void gen() {
    unsigned x0 = c0, x1 = c1, x2 = blk, x3 = 0u, a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const unsigned hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
      const unsigned hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
      const unsigned y0 = hi1 ^ x1 ^ a, y2 = hi0 ^ x3 ^ b;
      x0 = y0; x1 = lo1; x2 = y2; x3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    w[0] = x0; w[1] = x1; w[2] = x2; w[3] = x3;
    ++blk; pos = 0;
  }
  __device__ __forceinline__ unsigned next() {
    if (pos == 4) gen();
    // static indexing keeps w[] in registers
    const unsigned r = pos == 0 ? w[0] : pos == 1 ? w[1] : pos == 2 ? w[2] : w[3];
    ++pos;
    return r;
  }
};

// ---- compile-time-degree helpers ---------------------------------------------------------
// DEG > 0: the degree is a template constant (loops unroll, coefficients live in registers);
// DEG == 0: run-time degree `deg` (generic path).
template <int DEG>
struct Deg {
  int rt;
  __device__ __forceinline__ explicit Deg(int deg) : rt(deg) {}
  __device__ __forceinline__ int get() const { return DEG > 0 ? DEG : rt; }
};
__host__ __device__ constexpr int deg_cap(int DEG) { return DEG > 0 ? DEG : RB2_MAX_DEG; }

// Horner in the reference's arithmetic (np.polynomial.polynomial.polyval: separate mul and add)
template <int CAP>
__device__ __forceinline__ double horner_ref(const double (&c)[CAP + 1], int n, double x) {
  double r = 0.0;
#pragma unroll
  for (int i = CAP; i >= 0; --i) {
    if (i == n - 1) r = c[i];
    else if (i < n - 1) r = add(c[i], mul(r, x));
  }
  return r;
}
template <int CAP>
__device__ __forceinline__ double horner_fast(const double (&a)[CAP + 1], int n, double x) {  // degree n
  double r = 0.0;
#pragma unroll
  for (int i = CAP; i >= 0; --i) {
    if (i == n) r = a[i];
    else if (i < n) r = fma(r, x, a[i]);
  }
  return r;
}

// ---- real roots inside (lo, hi) ------------------------------------------------------------
// Accuracy target is a fraction of a grid step: the callers re-check the grid points around
// every root with the reference's arithmetic.
__device__ __forceinline__ int quad_roots(double a, double b, double c, double lo, double hi, double* out) {
  // a x^2 + b x + c
  int n = 0;
  if (a == 0.0) {
    if (b != 0.0) { double r = -c / b; if (r > lo && r < hi) out[n++] = r; }
    return n;
  }
  const double disc = fma(b, b, -4.0 * a * c);
  if (!(disc >= 0.0)) return 0;
  const double sq = sqrt(disc);
  const double q = -0.5 * (b + copysign(sq, b));
  double r1 = q / a, r2 = (q != 0.0) ? c / q : r1;
  if (r1 > r2) { double t = r1; r1 = r2; r2 = t; }
  if (r1 > lo && r1 < hi) out[n++] = r1;
  if (r2 > lo && r2 < hi && r2 != r1) out[n++] = r2;
  return n;
}

// roots of sum_{i<=n} a[i] x^i (n <= CAP) in (lo, hi), ascending: between consecutive roots of
// f' the function is monotone -> bisection.  Recursion on the compile-time capacity.
template <int CAP>
__device__ __forceinline__ int real_roots(const double (&a)[CAP + 1], int n, double lo, double hi, double (&out)[CAP > 0 ? CAP : 1]) {
  if constexpr (CAP <= 0) {
    return 0;
  } else if constexpr (CAP <= 2) {
    const double a2 = (CAP >= 2 && n >= 2) ? a[CAP >= 2 ? 2 : 0] : 0.0;
    const double a1 = (n >= 1) ? a[1] : 0.0;
    if (n <= 0) return 0;
    return quad_roots(a2, a1, a[0], lo, hi, out);
  } else {
    while (n > 0 && a[n] == 0.0) --n;
    if (n <= 0) return 0;
    if (n <= 2) return quad_roots(n == 2 ? a[2] : 0.0, a[1], a[0], lo, hi, out);
    double da[CAP];           // derivative, degree n - 1 <= CAP - 1
#pragma unroll
    for (int i = 0; i < CAP; ++i) da[i] = (i < n) ? (double)(i + 1) * a[i + 1] : 0.0;
    double crit[CAP - 1 > 0 ? CAP - 1 : 1];
    const int nc = real_roots<CAP - 1>(da, n - 1, lo, hi, crit);
    int nout = 0;
    double left = lo, fl = horner_fast<CAP>(a, n, left);
#pragma unroll
    for (int s = 0; s < CAP; ++s) {
      if (s <= nc) {
        const double right = (s < nc) ? crit[s < CAP - 1 ? s : 0] : hi;
        const double fr = horner_fast<CAP>(a, n, right);
        if ((fl < 0.0) != (fr < 0.0) && right > left) {
          double xa = left, xb = right, fa = fl;
          for (int it = 0; it < 40 && (xb - xa) > 0.125; ++it) {
            const double xm = 0.5 * (xa + xb);
            const double fm = horner_fast<CAP>(a, n, xm);
            if ((fm < 0.0) == (fa < 0.0)) { xa = xm; fa = fm; } else { xb = xm; }
          }
          out[nout++] = 0.5 * (xa + xb);
        }
        left = right; fl = fr;
      }
    }
    return nout;
  }
}

// ---- monotonicity test (calibrator.py:585-600) -------------------------------
// all(diff(polyval(arange(pix_min, pix_max, 1), c)) > 0), in two parts so that stage B can run the second one
// on a compacted list: mono_ends() tests the first and the last difference (most non-monotonic hypotheses fail
// there), mono_interior() the grid points around every critical point of the difference polynomial.
template <int DEG>
struct MonoGrid {
  static constexpr int CAP = deg_cap(DEG);
  long long kmax;  // D_k for k in [0, kmax]; < 0: fewer than two grid points
  double pix_min;
  __device__ __forceinline__ MonoGrid(double pmin, double pmax) : pix_min(pmin) {
    const double gl = ceil((pmax - pmin) / 1.0);  // numpy arange length
    kmax = (gl >= 2.0) ? (long long)gl - 2 : -1;  // diff of <2 points is empty
  }
  __device__ __forceinline__ double D(const double (&c)[CAP + 1], int deg, long long k) const {
    const double t0 = add(pix_min, (double)k), t1 = add(pix_min, (double)(k + 1));
    return sub(horner_ref<CAP>(c, deg + 1, t1), horner_ref<CAP>(c, deg + 1, t0));
  }
};

template <int DEG>
__device__ __forceinline__ bool mono_ends(const double (&c)[deg_cap(DEG) + 1], int deg_rt, double pix_min, double pix_max) {
  const int deg = DEG > 0 ? DEG : deg_rt;
  const MonoGrid<DEG> g(pix_min, pix_max);
  if (g.kmax < 0) return true;
  return (g.D(c, deg, 0) > 0.0) && (g.D(c, deg, g.kmax) > 0.0);
}

template <int DEG>
__device__ __forceinline__ bool mono_interior(const double (&c)[deg_cap(DEG) + 1], int deg_rt, double pix_min, double pix_max) {
  constexpr int CAP = deg_cap(DEG);
  const int deg = DEG > 0 ? DEG : deg_rt;
  const MonoGrid<DEG> g(pix_min, pix_max);
  if (g.kmax < 0 || deg <= 2) return true;  // q is linear: extremes at the ends
  const long long kmax = g.kmax;
  // p~(tau) = p(pix_min + tau): Taylor shift
  double ps[CAP + 1];
#pragma unroll
  for (int i = 0; i <= CAP; ++i) ps[i] = (i <= deg) ? c[i] : 0.0;
#pragma unroll
  for (int i = 0; i < CAP; ++i)
#pragma unroll
    for (int j = CAP - 1; j >= 0; --j)
      if (j >= i && j < deg) ps[j] = fma(pix_min, ps[j + 1], ps[j]);
  // q(tau) = p~(tau+1) - p~(tau) = sum_i tau^i sum_{j>i} C(j,i) p~_j ; need q' (degree deg-2)
  double dq[(CAP - 2 > 0 ? CAP - 2 : 0) + 1];
#pragma unroll
  for (int i = 1; i < CAP; ++i) {
    double s = 0.0, binom = 1.0;  // C(j,i) for j = i.. : C(i,i)=1
#pragma unroll
    for (int j = 2; j <= CAP; ++j) {
      if (j > i) {
        binom = binom * (double)j / (double)(j - i);  // C(j,i)
        if (j <= deg) s = fma(binom, ps[j], s);
      }
    }
    if (i - 1 <= CAP - 2) dq[i - 1 <= CAP - 2 ? i - 1 : 0] = (i < deg) ? (double)i * s : 0.0;
  }
  double roots[(CAP - 2 > 0 ? CAP - 2 : 1)];
  const int nr = real_roots<(CAP - 2 > 0 ? CAP - 2 : 0)>(dq, deg - 2, 0.0, (double)kmax, roots);
  for (int r = 0; r < nr; ++r) {
    const long long kf = (long long)floor(roots[r]);
    for (long long k = kf - 1; k <= kf + 2; ++k) {
      if (k < 0 || k > kmax) continue;
      if (!(g.D(c, deg, k) > 0.0)) return false;
    }
  }
  return true;
}

// ---- Hough-density weight (calibrator.py:497-506, 614-627) ---------------------
template <int DEG>
struct WeightCtx {
  static constexpr int CAP = deg_cap(DEG);
  double c[CAP + 1];   // coefficients (registers)
  double dc[CAP + 1];  // derivative coefficients i*c[i] (degree deg-1), padded
  int deg;
  const double* pix;   // global or NULL
  __device__ __forceinline__ double px(int i) const { return pix ? pix[i] : (double)i; }
  // tangent intercept t(x) = p(x) - x p'(x), the reference's arithmetic (:614-618)
  __device__ __forceinline__ double t_at(int i) const {
    const double x = px(i);
    const double wave = horner_ref<CAP>(c, deg + 1, x);
    const double grad = (deg >= 1) ? horner_ref<CAP>(dc, deg, x) : 0.0;
    return sub(wave, mul(grad, x));
  }
};

enum { CMP_GT = 0, CMP_GE = 1, CMP_LT = 2, CMP_LE = 3 };

__device__ __noinline__ double pp_eval(const rb2_ransac_desc& d, double t) {
  // 1-D cubic spline along the first intercept row, pp-form
  int a = 0, b = d.n_pp - 1;
  while (a < b) {
    const int m = (a + b + 1) >> 1;
    if (t >= d.d_pp_breaks[m]) a = m; else b = m - 1;
  }
  const double dt = t - d.d_pp_breaks[a];
  const double* k = d.d_pp_coef + 4 * a;
  return fma(fma(fma(k[3], dt, k[2]), dt, k[1]), dt, k[0]);
}

// Hough-density weight of one hypothesis, ONE THREAD (the work is scalar: roots, a few class
// tests, bisections).  Returns false when the hypothesis leaves the corner-clamped regime
// (gradient above the first intercept-bin centre somewhere on the detector).  need_value =
// false: only the regime is established (*out_weight = NaN unless it comes for free).
template <int DEG>
__device__ __forceinline__ bool hough_weight_thread(const rb2_ransac_desc& d, const double* c, int deg_rt, bool need_value,
                                                    double* out_weight) {
  constexpr int CAP = deg_cap(DEG);
  const int deg = DEG > 0 ? DEG : deg_rt;
  const int np = d.n_pix;
  if (np <= 0) { *out_weight = 0.0; return true; }
  WeightCtx<DEG> w;
  w.deg = deg; w.pix = d.d_pix;
#pragma unroll
  for (int i = 0; i <= CAP; ++i) w.c[i] = (i <= deg) ? c[i] : 0.0;
#pragma unroll
  for (int i = 0; i <= CAP; ++i) w.dc[i] = (i < deg) ? mul((double)(i + 1), w.c[i + 1 <= CAP ? i + 1 : CAP]) : 0.0;
  const double x0 = w.px(0), x1 = w.px(np - 1);
  const double th_hi = d.gc_max, th_lo = d.gc_min;
  // breakpoints of t: t'(x) = -x p''(x)
  constexpr int P2 = CAP - 2 > 0 ? CAP - 2 : 0;   // capacity (degree) of p''
  double brk[P2 + 2];
  int nb = 0;
  if (deg >= 3) {
    double p2[P2 + 1];
#pragma unroll
    for (int i = 0; i <= P2; ++i) p2[i] = (i + 2 <= deg) ? (double)((i + 2) * (i + 1)) * w.c[i + 2 <= CAP ? i + 2 : CAP] : 0.0;
    double rts[P2 > 0 ? P2 : 1];
    nb = real_roots<P2>(p2, deg - 2, x0, x1, rts);
#pragma unroll
    for (int i = 0; i < (P2 > 0 ? P2 : 1); ++i) if (i < nb) brk[i] = rts[i];
  }
  // regime check: max of p'(x) over [x0, x1] must stay <= first intercept-bin centre
  {
    double gmax = 0.0;
    if (deg >= 1) {
      gmax = fmax(horner_fast<CAP>(w.dc, deg - 1, x0), horner_fast<CAP>(w.dc, deg - 1, x1));
      for (int r = 0; r < nb; ++r) gmax = fmax(gmax, horner_fast<CAP>(w.dc, deg - 1, brk[r]));
    }
    if (!(gmax <= d.ic_min)) return false;
  }
  if (!need_value) { *out_weight = CUDART_NAN; return true; }  // regime established; the value cannot matter
  if (x0 < 0.0 && x1 > 0.0) {  // x = 0 is a breakpoint too
    int k = nb;
    while (k > 0 && brk[k - 1] > 0.0) { brk[k] = brk[k - 1]; --k; }
    brk[k] = 0.0;
    ++nb;
  }
  auto cls = [&](int i) { const double t = w.t_at(i); return (t >= th_hi) ? 2 : ((t <= th_lo) ? 0 : 1); };
  long long n_hi = 0, n_lo = 0;
  double mid = 0.0;
  int ia = 0;
  for (int pc = 0; pc <= nb; ++pc) {
    int ib;
    if (pc < nb) {
      // last index with px <= brk[pc]
      const double r = brk[pc];
      if (d.d_pix) {
        int a = ia - 1, b = np - 1;
        while (a < b) { const int m = (a + b + 1) >> 1; if (d.d_pix[m] <= r) a = m; else b = m - 1; }
        ib = a;
      } else {
        ib = (int)fmin(fmax(floor(r), -1.0), (double)(np - 1));
      }
    } else {
      ib = np - 1;
    }
    if (ib < ia) continue;
    const int ca = cls(ia), cb = (ib > ia) ? cls(ib) : ca;
    int i_mid0, i_mid1;  // MID pixels are [i_mid0, i_mid1)
    if (ca == cb) {
      if (ca == 2) n_hi += ib - ia + 1; else if (ca == 0) n_lo += ib - ia + 1;
      i_mid0 = ia; i_mid1 = (ca == 1) ? ib + 1 : ia;
    } else {
      // increasing piece: LO.. MID.. HI, decreasing piece: HI.. MID.. LO.  Two bisections over the
      // pixel index (first pixel past the first class, first pixel of the last class), the
      // reference's arithmetic at every probe; ONE inlined copy of the loop, coefficients in registers.
      const bool inc = ca < cb;
      int f[2];
#pragma unroll 1
      for (int sidx = 0; sidx < 2; ++sidx) {
        const bool first = (sidx == 0);
        const bool needed = first ? (ca != 1) : (cb != 1);
        int lo = first ? ia : f[0], hi = ib + 1;
        if (!needed) { lo = first ? ia : ib + 1; hi = lo; }
        const double th = (first == inc) ? th_lo : th_hi;
        while (lo < hi) {
          const int m = (lo + hi) >> 1;
          const double t = w.t_at(m);
          // inc: first  t > th_lo, then  t >= th_hi ; dec: first  t < th_hi, then  t <= th_lo
          const bool pr = inc ? (first ? (t > th) : (t >= th)) : (first ? (t < th) : (t <= th));
          if (pr) hi = m; else lo = m + 1;
        }
        if (first) f[0] = lo; else f[1] = lo;
      }
      if (inc) { n_lo += f[0] - ia; n_hi += ib + 1 - f[1]; } else { n_hi += f[0] - ia; n_lo += ib + 1 - f[1]; }
      i_mid0 = f[0]; i_mid1 = f[1];
    }
    for (int i = i_mid0; i < i_mid1; ++i) {  // class re-checked per pixel: never extrapolate the pp-form
      const double t = w.t_at(i);
      mid += (t >= th_hi) ? d.h_hi : ((t <= th_lo) ? d.h_lo : pp_eval(d, t));
    }
    ia = ib + 1;
  }
  *out_weight = d.hough_weight * ((double)n_hi * d.h_hi + (double)n_lo * d.h_lo + mid);
  return true;
}

// ---- least squares for the generic path: Householder QR on the Vandermonde in
// a centred, scaled variable; returns monomial coefficients in x.
__device__ void lsq_polyfit(const double* xs, const double* ys, int m, int deg, double xc, double xsc, double* coef) {
  const int n = deg + 1;
  double A[RB2_MAX_SAMPLE][NCOEF];
  double b[RB2_MAX_SAMPLE];
  for (int i = 0; i < m; ++i) {
    const double t = (xs[i] - xc) * xsc;
    double pw = 1.0;
    for (int j = 0; j < n; ++j) { A[i][j] = pw; pw *= t; }
    b[i] = ys[i];
  }
  for (int k = 0; k < n; ++k) {
    double nrm = 0.0;
    for (int i = k; i < m; ++i) nrm = fma(A[i][k], A[i][k], nrm);
    nrm = sqrt(nrm);
    if (nrm == 0.0) continue;
    const double alpha = (A[k][k] > 0.0) ? -nrm : nrm;
    const double v0 = A[k][k] - alpha;
    // v = (v0, A[k+1..][k]); beta = 2 / (v.v)
    double vv = v0 * v0;
    for (int i = k + 1; i < m; ++i) vv = fma(A[i][k], A[i][k], vv);
    if (vv == 0.0) continue;
    const double beta = 2.0 / vv;
    for (int j = k + 1; j < n; ++j) {
      double s = v0 * A[k][j];
      for (int i = k + 1; i < m; ++i) s = fma(A[i][k], A[i][j], s);
      s *= beta;
      A[k][j] -= s * v0;
      for (int i = k + 1; i < m; ++i) A[i][j] = fma(-s, A[i][k], A[i][j]);
    }
    {
      double s = v0 * b[k];
      for (int i = k + 1; i < m; ++i) s = fma(A[i][k], b[i], s);
      s *= beta;
      b[k] -= s * v0;
      for (int i = k + 1; i < m; ++i) b[i] = fma(-s, A[i][k], b[i]);
    }
    A[k][k] = alpha;
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s = fma(-A[k][j], coef[j], s);
    coef[k] = s / A[k][k];
  }
  // t = (x - xc) * xsc  ->  monomials in x
  double sc = 1.0;
  for (int j = 0; j < n; ++j) { coef[j] *= sc; sc *= xsc; }
  for (int i = 0; i < n; ++i)
    for (int j = n - 2; j >= i; --j) coef[j] = fma(-xc, coef[j + 1], coef[j]);
}

// Over-determined fit (sample_size > deg + 1 and / or known pairs, calibrator.py:569-575) with the degree
// known at compile time: QR least squares by Givens rotations, one sample point at a time, on the
// Vandermonde in the centred, scaled pixel - the (deg+1)^2/2 entries of R and Q^T y live in registers,
// a rotation costs one rsqrt.  Backward stable like the Householder QR of the generic path (normal
// equations or an orthogonal-polynomial recurrence lose digits the parity tests can see), whose
// m x n matrix sits in local memory and makes it ~3x slower.
template <int DEG>
__device__ __forceinline__ void lsq_polyfit_givens(const double* xs, const double* ys, int m, double xc, double xsc,
                                                   double* coef) {
  constexpr int N = DEG + 1;
  double R[N][N], q[N];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    q[j] = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) R[j][k] = 0.0;
  }
  for (int i = 0; i < m; ++i) {
    const double t = (xs[i] - xc) * xsc;
    double row[N], y = ys[i];
    row[0] = 1.0;
#pragma unroll
    for (int j = 1; j < N; ++j) row[j] = row[j - 1] * t;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double ra = R[j][j], rb = row[j];
      const double h2 = fma(ra, ra, rb * rb);
      if (h2 > 0.0 && rb != 0.0) {
        const double rinv = rsqrt(h2);
        const double c = ra * rinv, sn = rb * rinv;
        R[j][j] = h2 * rinv;
#pragma unroll
        for (int k = j + 1; k < N; ++k) {
          const double rk = R[j][k], wk = row[k];
          R[j][k] = fma(c, rk, sn * wk);
          row[k] = fma(c, wk, -sn * rk);
        }
        const double qj = q[j];
        q[j] = fma(c, qj, sn * y);
        y = fma(c, y, -sn * qj);
      }
    }
  }
  double c[N];
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    double sk = q[k];
#pragma unroll
    for (int j = k + 1; j < N; ++j) sk = fma(-R[k][j], c[j], sk);
    c[k] = sk / R[k][k];
  }
  // t = (x - xc) * xsc  ->  monomials in x
  double sc = 1.0;
#pragma unroll
  for (int j = 0; j < N; ++j) { c[j] *= sc; sc *= xsc; }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = N - 2; j >= i; --j) c[j] = fma(-xc, c[j + 1], c[j]);
#pragma unroll
  for (int j = 0; j < N; ++j) coef[j] = c[j];
}

// exact interpolation through DEG+1 points: Newton divided differences with a
// single FP64 division (all pairwise reciprocals from one product inverse).
template <int DEG>
__device__ __forceinline__ void newton_fit(const double (&x)[DEG + 1], const double (&y)[DEG + 1], double (&c)[DEG + 1]) {
  constexpr int ND = DEG * (DEG + 1) / 2;
  double df[ND > 0 ? ND : 1];
  {
    int t = 0;
#pragma unroll
    for (int j = 1; j <= DEG; ++j)
#pragma unroll
      for (int k = j; k <= DEG; ++k) df[t++] = x[k] - x[k - j];
  }
  // reciprocals via prefix products and one division
  double pre[ND > 0 ? ND : 1];
  double run = 1.0;
#pragma unroll
  for (int t = 0; t < ND; ++t) { pre[t] = run; run *= df[t]; }
  double inv = 1.0 / run;
#pragma unroll
  for (int t = ND - 1; t >= 0; --t) {
    const double r = inv * pre[t];
    inv *= df[t];
    df[t] = r;  // now 1/df[t]
  }
  double dd[DEG + 1];
#pragma unroll
  for (int k = 0; k <= DEG; ++k) dd[k] = y[k];
  {
    int t = 0;
#pragma unroll
    for (int j = 1; j <= DEG; ++j) {
      // dd_new[k] = (dd[k] - dd[k-1]) * rinv(k, k-j) for k = j..DEG, using old values: go downwards
      double nd[DEG + 1];
#pragma unroll
      for (int k = j; k <= DEG; ++k) nd[k] = (dd[k] - dd[k - 1]) * df[t + (k - j)];
#pragma unroll
      for (int k = j; k <= DEG; ++k) dd[k] = nd[k];
      t += DEG + 1 - j;
    }
  }
  // Newton form -> monomials
#pragma unroll
  for (int k = 0; k <= DEG; ++k) c[k] = 0.0;
  c[0] = dd[DEG];
#pragma unroll
  for (int k = DEG - 1; k >= 0; --k) {
#pragma unroll
    for (int j = DEG - k; j >= 1; --j) c[j] = fma(-x[k], c[j], c[j - 1]);
    c[0] = fma(-x[k], c[0], dd[k]);
  }
}

// ---------------------------------------------------------------------------------------
// The staged pipeline.  A chunk of hypothesis ids of every problem goes through four small
// kernels connected by compacted survivor queues in global memory (L2-resident):
//
//   A  draw_fit     one thread per hypothesis: Philox draw (or replay) -> fit -> intercept test
//                   survivors (~16 %) -> queue A          [int + FP64, no divergence by cost class]
//   B  monotonic    one thread per queue-A entry: monotonicity test; survivors (~2 %) -> queue B
//   C  score        one warp per queue-B entry: bijective match, M-SAC cost, Hough weight;
//                   per-problem running best by a 128-bit atomic min on (cost key, ~id)
//   D  record       one thread per queue-B entry: the entry that owns the running best writes
//                   the winner record
//
// Queue entry = qs doubles: coefficients c[0..deg] + a tag (problem << 40 | local id).
// ---------------------------------------------------------------------------------------

struct alignas(16) BestKey { unsigned long long lo, hi; };  // hi = sortable cost, lo = ~global id

__device__ __forceinline__ unsigned long long cost_key(double c) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(c);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ bool key_better(const BestKey& a, const BestKey& b) {
  return (a.hi < b.hi) || (a.hi == b.hi && a.lo < b.lo);
}

constexpr unsigned long long TAG_ID_MASK = (1ull << 40) - 1ull;

struct Rng {  // 32-bit words for the peak draws, 16-bit halves for the candidate picks
  Philox p;
  unsigned half;
  bool has_half;
  __device__ __forceinline__ void init(unsigned long long seed, unsigned long long id) { p.init(seed, id); has_half = false; half = 0; }
  __device__ __forceinline__ unsigned next32() { return p.next(); }
  __device__ __forceinline__ unsigned next16() {
    if (has_half) { has_half = false; return half; }
    const unsigned w = p.next();
    half = w >> 16; has_half = true;
    return w & 0xffffu;
  }
  // exactly uniform in [0, K), K <= 65536, from the 16-bit stream (Lemire's multiply-shift with
  // the rare rejection that removes the bias; no division on the common path)
  __device__ __forceinline__ int uniform(int K) {
    unsigned x = next16() * (unsigned)K;
    if ((x & 0xffffu) < (unsigned)K) {
      const unsigned t = 65536u % (unsigned)K;
      while ((x & 0xffffu) < t) x = next16() * (unsigned)K;
    }
    return (int)(x >> 16);
  }
};

// ---- stage-A tables ---------------------------------------------------------------------------
// Stage A is bound by shared-memory wavefronts (every lane reads its own random table entries: ~3 wavefronts
// per 32-bit load, ~5.5 per 64-bit, ~9 per 128-bit), so the tables are as small as the draw allows:
//   PeakRec  16 B, one LDS.128: the peak's interval [lo, lo + wd) of the 32-bit sampling CDF
//            (calibrator.py:521-523; the last interval ends at 2^32, all arithmetic is mod 2^32) and its pixel;
//   okt      4 B: start of the peak's candidate list | K << 16 | (65536 mod K) << 24 - only read when the
//            peaks do not all have the same number K of candidates (then the three are warp-uniform constants
//            and the list of peak i starts at i * K);
//   cwave    8 B: the candidate wavelength.  "Already used" is tested on the wavelength itself (equal ids ==
//            equal values), the ids are not needed here.
struct alignas(16) PeakRec { unsigned lo, wd; double x; };

struct TablesA {  // shared-memory views of one problem (stage A)
  const PeakRec* prec; const unsigned* okt; const double* cwave;
  const unsigned short* lut;
  int n_u;
  unsigned uni;  // != 0: every peak has (uni & 0xff) candidates; uni = K | (65536 mod K) << 8
  __device__ __forceinline__ void cand_list(int i, unsigned& off, unsigned& K, unsigned& thr) const {
    if (uni) { K = uni & 0xffu; thr = uni >> 8; off = (unsigned)i * K; }
    else { const unsigned v = okt[i]; off = v & 0xffffu; K = (v >> 16) & 0xffu; thr = v >> 24; }
  }
};
constexpr int A_MAX_K = 255, A_MAX_CAND = 65535;  // limits of the packed okt word (else: RB2_ERR_ARG)

// index i of the interval that holds u (== searchsorted(cdf, u, 'right'), clamped) and its record: LUT over
// the top LUT_BITS bits.  A cell stores the index of its first value; bit 15 marks the few cells (about n_u
// of 2^LUT_BITS) that a CDF boundary may cross - only those walk the records, so a warp usually takes the
// two-load path with all lanes.
#ifndef RB2_LUT_BITS
#define RB2_LUT_BITS 12
#endif
constexpr int LUT_BITS = RB2_LUT_BITS;
constexpr unsigned short LUT_CHECK = 0x8000u;
__device__ __forceinline__ int find_peak(const TablesA& t, unsigned u, PeakRec& r) {
  const unsigned short e = t.lut[u >> (32 - LUT_BITS)];
  int i = e & 0x7fff;
  r = t.prec[i];
  if (e & LUT_CHECK) {
    const int last = t.n_u - 1;
    while (i < last && (u - r.lo) >= r.wd) { ++i; r = t.prec[i]; }
  }
  return i;
}

// One Philox4x32-10 block, stateless: counter = (id_lo, id_hi, blk, 0).
__device__ __forceinline__ void philox_block(unsigned k0, unsigned k1, unsigned c0, unsigned c1, unsigned blk,
                                             unsigned (&out)[4]) {
  unsigned x0 = c0, x1 = c1, x2 = blk, x3 = 0u, a = k0, b = k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * x0;
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * x2;
    const unsigned y0 = (unsigned)(p1 >> 32) ^ x1 ^ a, y2 = (unsigned)(p0 >> 32) ^ x3 ^ b;
    x0 = y0; x1 = (unsigned)p1; x2 = y2; x3 = (unsigned)p0;
    a += 0x9E3779B9u; b += 0xBB67AE85u;
  }
  out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}
// The round keys depend on the seed only; as a kernel parameter they sit in the constant bank and the
// key XOR of every round takes them as an immediate operand (no per-thread key schedule).
struct PhiloxKey { unsigned a[10], b[10]; };
static PhiloxKey philox_key(unsigned long long seed) {
  PhiloxKey k;
  unsigned a = (unsigned)seed, b = (unsigned)(seed >> 32);
  for (int r = 0; r < 10; ++r) { k.a[r] = a; k.b[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
  return k;
}
__device__ __forceinline__ void philox_block(const PhiloxKey& key, unsigned c0, unsigned c1, unsigned blk,
                                             unsigned (&out)[4]) {
  unsigned x0 = c0, x1 = c1, x2 = blk, x3 = 0u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * x0;
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * x2;
    const unsigned y0 = (unsigned)(p1 >> 32) ^ x1 ^ key.a[r], y2 = (unsigned)(p0 >> 32) ^ x3 ^ key.b[r];
    x0 = y0; x1 = (unsigned)p1; x2 = y2; x3 = (unsigned)p0;
  }
  out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}
// word `n` of the stream beyond the eager blocks (rare path; one copy of the code)
__device__ __noinline__ unsigned philox_word(unsigned k0, unsigned k1, unsigned c0, unsigned c1, unsigned n) {
  unsigned w[4];
  philox_block(k0, k1, c0, c1, n >> 2, w);
  const unsigned j = n & 3u;
  return j == 0 ? w[0] : j == 1 ? w[1] : j == 2 ? w[2] : w[3];
}

// Sampler of the exactly determined path (sample size S = deg + 1 known at compile time), SURVEY.md
// App. A.8.  Every lane generates the same NB Philox blocks (no divergence in the generator): word k
// draws peak k, half-word k of the following words is the first try of candidate k.
//
// Peaks - np.random.choice(peaks, S, replace=False, p=prob) (:538-540) is successive sampling: the next
// peak is drawn from the law renormalised over the peaks not yet chosen.  No rejection loop: word k is
// scaled onto the REMAINING mass (u = mulhi(w, 2^32 - chosen mass): a monotone map, so every interval keeps
// its probability to 2^-32, the CDF's own resolution) and then shifted past the chosen intervals in
// ascending order (<= S-1 compare+add, the chosen list is kept sorted in registers).
//
// Candidates - uniform over the peak's candidate ARRAY (duplicates count, App. A.3), redrawn while the
// wavelength is already used (:544-566).  The fast path takes the first try of every pick (Lemire's
// multiply-shift on 16 bits); a hypothesis that needs a redraw (a Lemire rejection or a repeated
// wavelength, ~5 %) is reported back and finished by draw_candidates_slow() in a warp-wide round of
// such hypotheses (stage A's deferred queue), which continues with the Philox blocks after the eager
// ones; an impossible draw aborts there (the try is consumed, :557-566).
template <int S>
struct SamplerDims {
  static constexpr int NWC = (S + 1) / 2;       // words holding the candidate half-words
  static constexpr int NB = (S + NWC + 3) / 4;  // eager blocks
  static constexpr int NW = 4 * NB;
};

template <int S>
__device__ __forceinline__ void draw_peaks(const TablesA& t, const unsigned (&w)[SamplerDims<S>::NW], int (&pi)[S],
                                           double (&xs)[S]) {
  unsigned lo_s[S > 1 ? S - 1 : 1], wd_s[S > 1 ? S - 1 : 1];  // chosen intervals, ascending lo
  unsigned mass = 0u;                                         // their total width (mod 2^32)
#pragma unroll
  for (int k = 0; k < S; ++k) {
    unsigned u = (k == 0) ? w[0] : __umulhi(w[k], 0u - mass);
#pragma unroll
    for (int j = 0; j < S - 1; ++j)
      if (j < k) u += (u >= lo_s[j]) ? wd_s[j] : 0u;
    PeakRec r;
    pi[k] = find_peak(t, u, r);
    xs[k] = r.x;
    if (k < S - 1) {
      unsigned cl = r.lo, cw = r.wd;
#pragma unroll
      for (int j = 0; j < S - 1; ++j) {
        if (j < k) {
          const bool sw = cl < lo_s[j];
          const unsigned tl = sw ? lo_s[j] : cl, tw = sw ? wd_s[j] : cw;
          lo_s[j] = sw ? cl : lo_s[j]; wd_s[j] = sw ? cw : wd_s[j];
          cl = tl; cw = tw;
        }
      }
      lo_s[k] = cl; wd_s[k] = cw;
      mass += r.wd;
    }
  }
}

template <int S, class KEY>
__device__ __forceinline__ void philox_eager(const KEY& key, unsigned c0, unsigned c1, unsigned (&w)[SamplerDims<S>::NW]) {
#pragma unroll
  for (int b = 0; b < SamplerDims<S>::NB; ++b) {
    unsigned o[4];
    if constexpr (sizeof(KEY) == sizeof(PhiloxKey)) philox_block(key, c0, c1, (unsigned)b, o);
    else philox_block((unsigned)key, (unsigned)(key >> 32), c0, c1, (unsigned)b, o);
    w[4 * b] = o[0]; w[4 * b + 1] = o[1]; w[4 * b + 2] = o[2]; w[4 * b + 3] = o[3];
  }
}

__device__ __forceinline__ bool same_wave(double a, double b) {  // bit equality on the integer pipe
  return __double_as_longlong(a) == __double_as_longlong(b);
}

// Candidate half-words: half j lives in word S + j / 2.  Halves 0..S-1 are the first tries of the S picks;
// the halves after them - first the ones left in the eager words, then the words of the following Philox
// blocks - form ONE retry stream that the picks consume in order.
template <int S>
struct CandStream {
  static constexpr int NW = SamplerDims<S>::NW;
  static constexpr int EAGER = 2 * (NW - S);                                  // halves held by the eager words
  static constexpr int FAST_RETRIES = (EAGER - S) < 2 ? (EAGER - S) : 2;      // retries the fast path may take
  __device__ static __forceinline__ unsigned eager_half(const unsigned (&w)[NW], int j) {
    return (j & 1) ? (w[S + j / 2] >> 16) : (w[S + j / 2] & 0xffffu);
  }
};

// returns false when the draw needs more retries than the eager words hold (-> slow round); pi / ci / xs / ys
// are complete when true.  Most redraws (a repeated wavelength, ~7 % of the C3 hypotheses) are settled here
// with the spare eager halves; the slow round sees ~1 %.
template <int S>
__device__ __forceinline__ bool draw_sample_fast(const TablesA& t, const PhiloxKey& key, unsigned long long gid,
                                                 int (&pi)[S], int (&ci)[S], double (&xs)[S], double (&ys)[S]) {
  using CS = CandStream<S>;
  unsigned w[SamplerDims<S>::NW];
  philox_eager<S>(key, (unsigned)gid, (unsigned)(gid >> 32), w);
  draw_peaks<S>(t, w, pi, xs);
  bool ok = true;
  int nr = 0;  // retry halves consumed
#pragma unroll
  for (int k = 0; k < S; ++k) {
    unsigned off, K, thr;
    t.cand_list(pi[k], off, K, thr);
    unsigned h = CS::eager_half(w, k);
    bool good;
    int c;
    double wv;
    {
      const unsigned x = h * K;
      good = (x & 0xffffu) >= thr;
      c = (int)(off + (x >> 16));
      wv = t.cwave[c];
#pragma unroll
      for (int q = 0; q < S; ++q) if (q < k) good &= !same_wave(ys[q], wv);
    }
#pragma unroll
    for (int r = 0; r < CS::FAST_RETRIES; ++r) {
      if (ok && !good && nr < CS::FAST_RETRIES) {
        h = (nr == 0) ? CS::eager_half(w, S) : CS::eager_half(w, S + (CS::FAST_RETRIES > 1 ? 1 : 0));
        ++nr;
        const unsigned x = h * K;
        good = (x & 0xffffu) >= thr;
        c = (int)(off + (x >> 16));
        wv = t.cwave[c];
#pragma unroll
        for (int q = 0; q < S; ++q) if (q < k) good &= !same_wave(ys[q], wv);
      }
    }
    ok &= good;
    ci[k] = c; ys[k] = wv;
  }
  return ok;
}

constexpr int DRAW_MAX_TRIES = 1 << 14;  // per pick; a draw that still fails then is counted as aborted

// the same hypothesis with unbounded redraw loops over the same retry stream; false = impossible draw (aborted)
template <int S>
__device__ __noinline__ bool draw_sample_slow(const TablesA& t, unsigned long long seed, unsigned long long gid,
                                              int (&pi)[S], int (&ci)[S], double (&xs)[S], double (&ys)[S]) {
  using CS = CandStream<S>;
  constexpr int NW = SamplerDims<S>::NW;
  const unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32), c0 = (unsigned)gid, c1 = (unsigned)(gid >> 32);
  unsigned w[NW];
  philox_eager<S>(seed, c0, c1, w);
  draw_peaks<S>(t, w, pi, xs);
  int n_retry = 0;     // position in the retry stream
  unsigned cur = 0u;   // current word beyond the eager blocks
  for (int k = 0; k < S; ++k) {
    unsigned off, K, thr;
    t.cand_list(pi[k], off, K, thr);
    unsigned h = CS::eager_half(w, k);
    bool checked = false;
    int n_dup = 0;
    for (int tries = 0;; ++tries) {
      if (tries >= DRAW_MAX_TRIES) return false;
      const unsigned x = h * K;
      if ((x & 0xffffu) >= thr) {
        const int c = (int)(off + (x >> 16));
        const double wv = t.cwave[c];
        bool dup = false;
        for (int q = 0; q < k; ++q) dup |= same_wave(ys[q], wv);
        if (!dup) { ci[k] = c; ys[k] = wv; break; }
        if (!checked && ++n_dup >= 3) {  // after a few duplicates: is any unused wavelength left for this peak?
          bool possible = false;
          for (unsigned jj = 0; jj < K; ++jj) {
            const double ww = t.cwave[off + jj];
            bool used = false;
            for (int q = 0; q < k; ++q) used |= same_wave(ys[q], ww);
            possible |= !used;
          }
          if (!possible) return false;  // impossible draw: the try is consumed (:557-566)
          checked = true;
        }
      }
      const int j = S + n_retry;  // next half of the retry stream
      ++n_retry;
      if (j < CS::EAGER) {
        h = 0u;
#pragma unroll
        for (int e = S; e < CS::EAGER; ++e) if (e == j) h = CS::eager_half(w, e);
      } else {
        const int z = j - CS::EAGER;
        if ((z & 1) == 0) cur = philox_word(k0, k1, c0, c1, (unsigned)NW + (unsigned)(z >> 1));
        h = (z & 1) ? (cur >> 16) : (cur & 0xffffu);
      }
    }
  }
  return true;
}

// generic (run-time sample size) variant for the least-squares path: rejection sampling with the
// sequential Philox stream (same law; the streams differ)
__device__ bool draw_sample_dyn(const TablesA& t, unsigned long long seed, unsigned long long gid, int s, int* pi, int* ci,
                                double* xs, double* ys) {
  Rng rng;
  rng.init(seed, gid);
  for (int k = 0; k < s; ++k) {
    int idx;
    bool dup;
    int tries = 0;
    PeakRec r;
    do {
      if (++tries > DRAW_MAX_TRIES) return false;  // degenerate sampling law: never spin
      idx = find_peak(t, rng.next32(), r);
      dup = false;
      for (int j = 0; j < k; ++j) dup |= (pi[j] == idx);
    } while (dup);
    pi[k] = idx; xs[k] = r.x;
  }
  for (int k = 0; k < s; ++k) {
    unsigned off, K, thr;
    t.cand_list(pi[k], off, K, thr);
    bool checked = false;
    for (int tries = 0;; ++tries) {
      if (tries >= DRAW_MAX_TRIES) return false;
      const int j = rng.uniform((int)K);
      const double wv = t.cwave[off + j];
      bool dup = false;
      for (int q = 0; q < k; ++q) dup |= same_wave(ys[q], wv);
      if (!dup) { ys[k] = wv; ci[k] = (int)off + j; break; }
      if (!checked) {
        bool possible = false;
        for (unsigned jj = 0; jj < K; ++jj) {
          const double ww = t.cwave[off + jj];
          bool used = false;
          for (int q = 0; q < k; ++q) used |= same_wave(ys[q], ww);
          possible |= !used;
        }
        if (!possible) return false;
        checked = true;
      }
    }
  }
  return true;
}

struct Pipe {  // device pointers of the staged pipeline (one per rb2_ransac_batch call)
  double* qa; double* qb;           // queues, `cap` entries of qs doubles
  unsigned* counts;                 // [0] = |queue A|, [1] = |queue B|
  BestKey* best;                    // [n_problems] running best over the stage-C2 costs (tree-summed errors)
  BestKey* best_exact;              // [n_problems] running best over the exact costs of the contenders (stage C3)
  double* res_cost; double* res_esum; int2* res_cnt;  // per queue-B entry
  double* res_w;                    // per queue-B entry: Hough-density weight (when the cost was evaluated)
  unsigned char* res_flag;          // per queue-B entry: RES_* bits
  rb2_ransac_result* win;           // [n_problems] winner records written by THIS lane's stage D
  unsigned cap; int qs;
  unsigned cap_a;                   // capacity of queue A: cap + the slack of the paged reservation (stage A)
};
enum : unsigned char { RES_ELIGIBLE = 1, RES_NEAR_FLOOR = 2, RES_EXACT = 4, RES_GENERAL = 8 };
// two costs computed from the same coefficients differ only by the summation order of <= n_peaks clipped
// errors (a few ulp); anything within this relative margin of the running best is re-summed exactly
constexpr double COST_MARGIN = 1e-12;

constexpr unsigned QA_PAGE = 32u;                    // slots a warp of stage A reserves at a time
constexpr unsigned QA_SLACK = 1u << 20;              // extra queue-A slots for the unfilled ends of the pages (<= 2 pages per warp)
constexpr unsigned long long QA_HOLE = ~0ull;        // tag of a reserved slot that holds no hypothesis

constexpr int PIPE_MAX_LANES = 4;
struct PipeSet { Pipe p[PIPE_MAX_LANES]; int n; };

__global__ void pipe_init_kernel(PipeSet ps, rb2_ransac_result* __restrict__ result, int n_problems) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2) for (int l = 0; l < ps.n; ++l) ps.p[l].counts[i] = 0u;
  if (i >= n_problems) return;
  ps.p[0].best[i].hi = ~0ull; ps.p[0].best[i].lo = ~0ull;  // the running bests are shared by all lanes
  ps.p[0].best_exact[i].hi = ~0ull; ps.p[0].best_exact[i].lo = ~0ull;
  rb2_ransac_result r;
  r.cost = CUDART_INF; r.hyp_id = -1; r.n_match = 0; r.n_inliers = 0;
  for (int k = 0; k < NCOEF; ++k) r.coeffs[k] = 0.0;
  for (int k = 0; k < 8; ++k) r.counters[k] = 0ull;
  result[i] = r;
  for (int l = 0; l < ps.n; ++l) ps.p[l].win[i] = r;
}

// after all lanes have drained: the best of the lane winners (lower cost, later id on ties,
// calibrator.py:636) becomes the problem's record; the counters already live in `result`
__global__ void pipe_join_kernel(PipeSet ps, rb2_ransac_result* __restrict__ result, int n_problems) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_problems) return;
  int best = 0;
  for (int l = 1; l < ps.n; ++l) {
    const rb2_ransac_result& a = ps.p[best].win[i];
    const rb2_ransac_result& b = ps.p[l].win[i];
    if (b.hyp_id >= 0 && (a.hyp_id < 0 || b.cost < a.cost || (b.cost == a.cost && b.hyp_id > a.hyp_id))) best = l;
  }
  const rb2_ransac_result& w = ps.p[best].win[i];
  rb2_ransac_result* r = result + i;
  r->cost = w.cost; r->hyp_id = w.hyp_id; r->n_match = w.n_match; r->n_inliers = w.n_inliers;
  for (int k = 0; k < NCOEF; ++k) r->coeffs[k] = w.coeffs[k];
}

__global__ void pipe_reset_counts_kernel(Pipe p) { if (threadIdx.x < 2) p.counts[threadIdx.x] = 0u; }

__device__ __forceinline__ void add_counter(rb2_ransac_result* r, int k, unsigned v) {
  if (v) atomicAdd(reinterpret_cast<unsigned long long*>(&r->counters[k]), (unsigned long long)v);
}

// ---- stage A ---------------------------------------------------------------------------
#ifndef RB2_A_CTAS
#define RB2_A_CTAS 4
#endif
// LSQ = false: sample_size == deg + 1, exact interpolation (newton_fit); LSQ = true (DEG_T > 0): run-time
// sample size and known pairs, least squares with the compile-time degree (lsq_polyfit_givens).
//
// Work items of a warp are either the next 32 hypothesis ids of its stream (fast sampler) or 32
// hypotheses from its deferred queue (those whose fast draw needs a redraw: slow sampler, all lanes
// busy again); draw, fit, intercept test and the queue push are the same code for both.
// PLAIN = the production launch: no replayed draws, no per-hypothesis outputs, monomial basis - the tests of
// those switches and the code behind them are compiled out (they were ~5 % of the instructions of a launch that
// uses none of them).
template <int DEG_T, bool LSQ = false, bool PLAIN = false>
__global__ void __launch_bounds__(RS_THREADS, DEG_T > 0 && DEG_T <= 4 && !LSQ ? RB2_A_CTAS : 2)
ransac_draw_fit_kernel(const rb2_ransac_desc* __restrict__ descs, Pipe pipe, long long chunk_begin, long long chunk_len,
                       rb2_ransac_result* __restrict__ result, const PhiloxKey pkey, const int paged) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ rb2_ransac_desc s_desc;
  __shared__ unsigned s_cnt[3];
  __shared__ unsigned s_defer[RS_WARPS][64];
  constexpr bool FAST = DEG_T > 0 && !LSQ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int problem = blockIdx.y;
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(descs + problem);
    unsigned* dst = reinterpret_cast<unsigned*>(&s_desc);
    for (int i = tid; i < (int)(sizeof(rb2_ransac_desc) / 4); i += RS_THREADS) dst[i] = src[i];
    if (tid < 3) s_cnt[tid] = 0u;
  }
  __syncthreads();
  const rb2_ransac_desc& d = s_desc;
  const int deg = DEG_T > 0 ? DEG_T : d.deg;
  const int ncoef = deg + 1;
  const int n_u = d.n_upeaks, n_c = d.n_cand;

  __shared__ int s_kmin, s_kmax;
  PeakRec* s_prec = reinterpret_cast<PeakRec*>(smem_raw);
  double* s_cwave = reinterpret_cast<double*>(s_prec + n_u);
  unsigned* s_okt = reinterpret_cast<unsigned*>(s_cwave + n_c);
  unsigned* s_cdf = s_okt + n_u;
  unsigned short* s_lut = reinterpret_cast<unsigned short*>(s_cdf + n_u);
  if (tid == 0) { s_kmin = 0x7fffffff; s_kmax = 0; }
  __syncthreads();
  for (int i = tid; i < n_u; i += RS_THREADS) {
    const unsigned hi = d.d_cdf32 ? d.d_cdf32[i] : 0u, lo = (i > 0 && d.d_cdf32) ? d.d_cdf32[i - 1] : 0u;
    const int o0 = d.d_cand_off[i], K = d.d_cand_off[i + 1] - o0;
    PeakRec r;
    r.lo = lo; r.wd = (i == n_u - 1) ? 0u - lo : hi - lo;  // the last interval ends at 2^32
    r.x = d.d_peaks[i];
    s_prec[i] = r;
    s_okt[i] = (unsigned)o0 | ((unsigned)K << 16) | ((K > 0 ? 65536u % (unsigned)K : 0u) << 24);
    s_cdf[i] = hi;
    atomicMin(&s_kmin, K); atomicMax(&s_kmax, K);
  }
  for (int i = tid; i < n_c; i += RS_THREADS) s_cwave[i] = d.d_cand_wave[i];
  __syncthreads();
  {
    // lut[b] = #{i : cdf32[i] <= b << (32 - LUT_BITS)}, clamped: a thread fills a run of consecutive cells
    // (one binary search, then a monotone walk); a cell is marked for checking when the next cell
    // starts at another index (or the table cannot hold the index in 15 bits)
    constexpr int NC = 1 << LUT_BITS, PER = NC / RS_THREADS;
    const int b0 = tid * PER;
    int lo = 0, hi = n_u - 1;
    const unsigned r0 = (unsigned)b0 << (32 - LUT_BITS);
    while (lo < hi) { const int m = (lo + hi) >> 1; if (r0 < s_cdf[m]) hi = m; else lo = m + 1; }
    for (int b = b0; b < b0 + PER; ++b) {
      const unsigned r = (unsigned)b << (32 - LUT_BITS);
      while (lo < n_u - 1 && s_cdf[lo] <= r) ++lo;
      s_lut[b] = (unsigned short)min(lo, 0x7fff);
    }
    __syncthreads();
    const bool wide = n_u > 0x7fff;
    for (int b = b0; b < b0 + PER; ++b) {
      const int here = s_lut[b] & 0x7fff;
      const int next = (b + 1 < NC) ? (s_lut[b + 1] & 0x7fff) : (n_u - 1);
      if (wide || next != here) s_lut[b] = (unsigned short)(here | LUT_CHECK);
    }
  }
  __syncthreads();
  TablesA tb;
  tb.prec = s_prec; tb.okt = s_okt; tb.cwave = s_cwave; tb.lut = s_lut; tb.n_u = n_u;
  tb.uni = (s_kmin == s_kmax && s_kmax > 0) ? ((unsigned)s_kmax | ((65536u % (unsigned)s_kmax) << 8)) : 0u;

  const bool replay = !PLAIN && d.d_replay_x != nullptr;
  const int s = FAST ? DEG_T + 1 : d.sample_size;
  const double min_i = d.min_intercept, max_i = d.max_intercept;
  const long long n_hyp = d.n_hyp;
  const double lsq_xc = 0.5 * (d.pix_min + d.pix_max);
  const double lsq_xsc = (d.pix_max > d.pix_min) ? 2.0 / (d.pix_max - d.pix_min) : 1.0;
  const int qs = FAST ? DEG_T + 2 : pipe.qs;
  unsigned c_drawn = 0, c_abort = 0, c_icpt = 0, w_pass = 0;  // w_pass: survivors of this warp (warp-uniform)
  const bool diag = !PLAIN && (d.d_out_status != nullptr || d.d_out_coeffs != nullptr || d.d_out_sample != nullptr);  // parity / diagnostics launches only
  const bool out_sample = !PLAIN && d.d_out_sample != nullptr;
  const bool other_basis = !PLAIN && d.basis != 0;

  const long long stride = (long long)gridDim.x * RS_THREADS;
  // ids of this chunk that exist (a chunk is at most 2^26 ids: 32-bit arithmetic in the loop)
  const unsigned limit = (unsigned)max(0ll, min(chunk_len, n_hyp - chunk_begin));
  const unsigned stride32 = (unsigned)stride;
  unsigned base = (unsigned)(blockIdx.x * RS_THREADS + (tid - lane));
  unsigned qn = 0u;  // hypotheses in this warp's deferred queue (warp-uniform)
  unsigned* const s_q = s_defer[warp];
  // Queue-A slots.  One returning atomic per warp iteration on ONE address was the largest single stall of this
  // kernel (99.6 % of the iterations have a survivor; the warp waited ~1 us for its slot).  Paged: a warp owns a
  // page of QA_PAGE slots and holds the NEXT page already - its atomic was issued a page earlier, so the reply is
  // there when it is needed.  Slots of the last two pages that stay empty are tagged QA_HOLE (stage B skips them).
  unsigned pg_base = 0u, pg_used = QA_PAGE, pg_next = 0u;  // pg_next: lane 0 only (raw reply of the prefetch)
  bool pg_have = false;
  if (paged && lane == 0) pg_next = atomicAdd(&pipe.counts[0], QA_PAGE);
  for (;;) {
    const bool stream_left = base < limit;
    bool slow = false, active = false;
    unsigned rel = 0u;  // hypothesis index inside the chunk
    if (FAST && (qn >= 32u || (!stream_left && qn > 0u))) {
      const unsigned take = min(qn, 32u);
      slow = true;
      active = (unsigned)lane < take;
      if (active) rel = s_q[qn - take + lane];
      qn -= take;
      __syncwarp();
    } else if (stream_left) {
      rel = base + (unsigned)lane;
      active = rel < limit;
      base += stride32;
    } else {
      break;
    }
    const long long id = chunk_begin + (long long)rel;  // index inside this call
    bool pass = false, aborted = false, deferred = false;
    double cf[NCOEF];
#pragma unroll
    for (int i = 0; i < NCOEF; ++i) cf[i] = 0.0;
    if constexpr (FAST) {
      int pi[DEG_T + 1], ci[DEG_T + 1];
      double xs[DEG_T + 1], ys[DEG_T + 1];
      if (active) {
        if (replay) {
#pragma unroll
          for (int k = 0; k <= DEG_T; ++k) {
            pi[k] = d.d_replay_x[id * (DEG_T + 1) + k];
            ci[k] = (int)(tb.okt[pi[k]] & 0xffffu) + d.d_replay_y[id * (DEG_T + 1) + k];
            xs[k] = tb.prec[pi[k]].x; ys[k] = tb.cwave[ci[k]];
          }
        } else if (!slow) {
          deferred = !draw_sample_fast<DEG_T + 1>(tb, pkey, (unsigned long long)(d.hyp_begin + id), pi, ci, xs, ys);
        } else {
          // own arrays: the call is not inlined, its outputs live in local memory - only here
          int pi2[DEG_T + 1], ci2[DEG_T + 1];
          double xs2[DEG_T + 1], ys2[DEG_T + 1];
          aborted = !draw_sample_slow<DEG_T + 1>(tb, d.seed, (unsigned long long)(d.hyp_begin + id), pi2, ci2, xs2, ys2);
#pragma unroll
          for (int k = 0; k <= DEG_T; ++k) { pi[k] = pi2[k]; ci[k] = ci2[k]; xs[k] = xs2[k]; ys[k] = ys2[k]; }
        }
      }
      if (!slow && !replay) {  // park the hypotheses that need a redraw
        const unsigned dm = __ballot_sync(0xffffffffu, deferred);
        if (dm) {
          if (deferred) s_q[qn + __popc(dm & ((1u << lane) - 1u))] = rel;
          qn += __popc(dm);
          __syncwarp();
        }
        active = active && !deferred;
      }
      if (active && !aborted) {
        if (replay && d.d_replay_coeffs) {
#pragma unroll
          for (int k = 0; k <= DEG_T; ++k) cf[k] = d.d_replay_coeffs[id * (DEG_T + 1) + k];
        } else {
          double cc[DEG_T + 1];
          newton_fit<DEG_T>(xs, ys, cc);
#pragma unroll
          for (int k = 0; k <= DEG_T; ++k) cf[k] = cc[k];
        }
      }
      if (active && out_sample) {
#pragma unroll
        for (int k = 0; k <= DEG_T; ++k) {
          d.d_out_sample[(id * (DEG_T + 1) + k) * 2] = aborted ? -1 : pi[k];
          d.d_out_sample[(id * (DEG_T + 1) + k) * 2 + 1] = aborted ? -1 : ci[k] - (int)(tb.okt[pi[k]] & 0xffffu);
        }
      }
    } else if (active) {
      int pi[RB2_MAX_SAMPLE], ci[RB2_MAX_SAMPLE];
      double xs[RB2_MAX_SAMPLE], ys[RB2_MAX_SAMPLE];
      if (replay) {
        for (int k = 0; k < s; ++k) {
          pi[k] = d.d_replay_x[id * s + k];
          ci[k] = (int)(tb.okt[pi[k]] & 0xffffu) + d.d_replay_y[id * s + k];
          xs[k] = tb.prec[pi[k]].x; ys[k] = tb.cwave[ci[k]];
        }
      } else {
        aborted = !draw_sample_dyn(tb, d.seed, (unsigned long long)(d.hyp_begin + id), s, pi, ci, xs, ys);
      }
      if (!aborted) {
        if (replay && d.d_replay_coeffs) {
          for (int k = 0; k < ncoef; ++k) cf[k] = d.d_replay_coeffs[id * ncoef + k];
        } else {
          int m = s;
          for (int k = 0; k < d.n_known && m < RB2_MAX_SAMPLE; ++k) { xs[m] = d.d_known_x[k]; ys[m] = d.d_known_y[k]; ++m; }
          if constexpr (LSQ) {
            lsq_polyfit_givens<DEG_T>(xs, ys, m, lsq_xc, lsq_xsc, cf);
          } else {
            lsq_polyfit(xs, ys, m, deg, lsq_xc, lsq_xsc, cf);
          }
        }
      }
      if (out_sample)
        for (int k = 0; k < s; ++k) {
          d.d_out_sample[(id * s + k) * 2] = aborted ? -1 : pi[k];
          d.d_out_sample[(id * s + k) * 2 + 1] = aborted ? -1 : ci[k] - (int)(tb.okt[pi[k]] & 0xffffu);
        }
    }
    if (active) {
      if (!aborted) {
        double c0 = cf[0];
        if (other_basis) {  // the first coefficient of the BASIS the reference fitted in (legfit / chebfit)
          c0 = 0.0;
          for (int k = 0; k < ncoef; ++k) c0 = fma(d.d_basis_c0[k], cf[k], c0);
        }
        pass = !((c0 < min_i) || (c0 > max_i));  // :578-582
      }
      if (diag) {
        if (!aborted) {
          if (d.d_out_coeffs)
            for (int i = 0; i < ncoef; ++i) d.d_out_coeffs[id * ncoef + i] = cf[i];
          if (!pass && d.d_out_status) d.d_out_status[id] = (uint8_t)RB2_HYP_INTERCEPT;
        } else if (d.d_out_status) {
          d.d_out_status[id] = (uint8_t)RB2_HYP_ABORT;
        }
      }
    }
    const unsigned pm = __ballot_sync(0xffffffffu, pass);
    if constexpr (FAST && PLAIN) {  // drawn / intercept follow from the ids this warp owns (closed form below)
      if (slow) c_abort += active && aborted;
      w_pass += (unsigned)__popc(pm);
    } else {
      c_drawn += active && !aborted;
      c_abort += active && aborted;
      c_icpt += active && !aborted && !pass;
    }
    if (pm) {  // warp-aggregated push
      const unsigned n_push = (unsigned)__popc(pm), rank = (unsigned)__popc(pm & ((1u << lane) - 1u));
      unsigned slot;
      if (paged) {
        const unsigned room = QA_PAGE - pg_used;
        unsigned nb = 0u;
        if (n_push > room) {  // warp-uniform: the held page becomes the current one, the next is requested
          nb = __shfl_sync(0xffffffffu, pg_next, 0);
          if (lane == 0) pg_next = atomicAdd(&pipe.counts[0], QA_PAGE);
        }
        slot = (rank < room) ? pg_base + pg_used + rank : nb + (rank - room);
        if (n_push > room) { pg_base = nb; pg_used = n_push - room; pg_have = true; }
        else pg_used += n_push;
      } else {
        unsigned slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(&pipe.counts[0], n_push);
        slot = __shfl_sync(0xffffffffu, slot0, 0) + rank;
      }
      if (pass && slot < pipe.cap_a) {
        double* q = pipe.qa + (size_t)slot * qs;
#pragma unroll
        for (int i = 0; i < NCOEF; ++i) if (i < ncoef) q[i] = cf[i];
        q[ncoef] = __longlong_as_double((long long)(((unsigned long long)problem << 40) | (unsigned long long)id));
      }
    }
  }
  if (paged) {  // unused slots of the current page and the whole held page
    const unsigned nb = __shfl_sync(0xffffffffu, pg_next, 0);
    for (unsigned k = lane; k < 2u * QA_PAGE; k += 32u) {
      const bool cur = k < QA_PAGE;
      const unsigned off = cur ? k : k - QA_PAGE;
      if (cur && (!pg_have || off < pg_used)) continue;
      const unsigned slot = (cur ? pg_base : nb) + off;
      if (slot < pipe.cap_a) pipe.qa[(size_t)slot * qs + ncoef] = __longlong_as_double((long long)QA_HOLE);
    }
  }
  // counters: warp -> block -> global
  c_abort = __reduce_add_sync(0xffffffffu, c_abort);
  if constexpr (FAST && PLAIN) {
    // every id of the warp's strided ranges below the end of the chunk / of the call is handled exactly once
    // (in a stream round, or in a deferred round): 32 per full round + the one partial round
    long long L = min(chunk_len, n_hyp - chunk_begin);
    const long long b0 = (long long)blockIdx.x * RS_THREADS + (tid - lane);
    long long owned = 0;
    if (L > b0) {
      const long long n_full = (L - b0 >= 32) ? (L - b0 - 32) / stride + 1 : 0;
      const long long start = b0 + n_full * stride;
      owned = 32 * n_full + ((start < L) ? min(32ll, L - start) : 0ll);
    }
    c_drawn = (unsigned)owned - c_abort;
    c_icpt = c_drawn - w_pass;
  } else {
    c_drawn = __reduce_add_sync(0xffffffffu, c_drawn);
    c_icpt = __reduce_add_sync(0xffffffffu, c_icpt);
  }
  if (lane == 0) { atomicAdd(&s_cnt[0], c_drawn); atomicAdd(&s_cnt[1], c_abort); atomicAdd(&s_cnt[2], c_icpt); }
  __syncthreads();
  if (tid == 0) {
    add_counter(result + problem, 0, s_cnt[0]);
    add_counter(result + problem, 1, s_cnt[1]);
    add_counter(result + problem, 2, s_cnt[2]);
  }
}

// ---- stage B ---------------------------------------------------------------------------
// Thread per queue-A entry for the end-point test (86 % of the entries stop there); the entries that pass it
// are parked in a per-warp list and their interior test - roots of the difference polynomial's derivative,
// data-dependent and ~10x the work - runs on 32 parked entries at a time, so that its lanes are all busy
// (as one fused test it ran with 9 of 32 lanes).
template <int DEG_T, bool PLAIN = false>
__global__ void __launch_bounds__(256)
ransac_monotonic_kernel(const rb2_ransac_desc* __restrict__ descs, Pipe pipe, rb2_ransac_result* __restrict__ result) {
  __shared__ unsigned s_pend[256 / 32][64];
  const unsigned n = min(pipe.counts[0], pipe.cap_a);
  const int lane = threadIdx.x & 31;
  const int qs = DEG_T > 0 ? DEG_T + 2 : pipe.qs, ncoef = qs - 1;
  const unsigned total = gridDim.x * blockDim.x;
  unsigned* const pend = s_pend[threadIdx.x >> 5];
  unsigned np = 0u;  // parked entries (warp-uniform)
  int acc_problem = -1;
  unsigned acc_n = 0;
  auto reject = [&](int problem, const rb2_ransac_desc* d, unsigned long long tag) {
    if (!PLAIN && d->d_out_status) d->d_out_status[tag & TAG_ID_MASK] = (uint8_t)RB2_HYP_MONOTONIC;
    if (problem != acc_problem) {  // per-thread run-length counter, flushed when the problem changes
      if (acc_n) add_counter(result + acc_problem, 3, acc_n);
      acc_problem = problem; acc_n = 0;
    }
    ++acc_n;
  };
  auto load = [&](unsigned i, double (&c)[deg_cap(DEG_T) + 1]) {
    const double* q = pipe.qa + (size_t)i * qs;
#pragma unroll
    for (int k = 0; k <= deg_cap(DEG_T); ++k) c[k] = (k < ncoef) ? q[k] : 0.0;
  };
  auto interior = [&](unsigned count) {  // the last `count` (<= 32) parked entries
    const bool have = (unsigned)lane < count;
    bool mono = false;
    double c[deg_cap(DEG_T) + 1];
    unsigned long long tag = 0;
    if (have) {
      const unsigned i = pend[np - count + lane];
      load(i, c);
      tag = (unsigned long long)__double_as_longlong(pipe.qa[(size_t)i * qs + ncoef]);
      const int problem = (int)(tag >> 40);
      const rb2_ransac_desc* d = descs + problem;
      mono = mono_interior<DEG_T>(c, DEG_T > 0 ? DEG_T : d->deg, d->pix_min, d->pix_max);
      if (!mono) reject(problem, d, tag);
    }
    np -= count;
    const unsigned pm = __ballot_sync(0xffffffffu, mono);
    if (pm) {
      unsigned slot0 = 0;
      if (lane == 0) slot0 = atomicAdd(&pipe.counts[1], (unsigned)__popc(pm));
      slot0 = __shfl_sync(0xffffffffu, slot0, 0);
      if (mono) {
        const unsigned slot = slot0 + __popc(pm & ((1u << lane) - 1u));
        double* q = pipe.qb + (size_t)slot * qs;
#pragma unroll
        for (int k = 0; k <= deg_cap(DEG_T); ++k) if (k < ncoef) q[k] = c[k];
        q[ncoef] = __longlong_as_double((long long)tag);
      }
    }
    __syncwarp();
  };
  for (unsigned i0 = blockIdx.x * blockDim.x + (threadIdx.x - lane); i0 < n; i0 += total) {
    const unsigned i = i0 + lane;
    bool have = i < n, park = false;
    if (have) {
      const unsigned long long tag = (unsigned long long)__double_as_longlong(pipe.qa[(size_t)i * qs + ncoef]);
      have = tag != QA_HOLE;  // an unfilled slot of stage A's paged reservation
      if (have) {
        double c[deg_cap(DEG_T) + 1];
        load(i, c);
        const int problem = (int)(tag >> 40);
        const rb2_ransac_desc* d = descs + problem;
        park = mono_ends<DEG_T>(c, DEG_T > 0 ? DEG_T : d->deg, d->pix_min, d->pix_max);
        if (!park) reject(problem, d, tag);
      }
    }
    const unsigned km = __ballot_sync(0xffffffffu, park);
    if (park) pend[np + __popc(km & ((1u << lane) - 1u))] = i;
    np += __popc(km);
    __syncwarp();
    if (np >= 32u) interior(32u);
  }
  if (np > 0u) interior(np);
  // final flush, aggregated over the lanes that hold the same problem
  const unsigned holders = __ballot_sync(0xffffffffu, acc_n > 0);
  if (acc_n > 0) {
    const unsigned peers = __match_any_sync(holders, acc_problem);
    const unsigned sum = __reduce_add_sync(peers, acc_n);
    if ((peers & ((1u << lane) - 1u)) == 0) add_counter(result + acc_problem, 3, sum);
  }
}

// ---- stage C1 --------------------------------------------------------------------------
// Bijective match, one warp per queue-B entry; dynamic smem per warp:
//   u64 berr[max_wids] (all +inf between entries); double be[max_up]; int bw[max_up]
// De-duplication by wavelength (calibrator.py:356-376): every peak posts its nearest-candidate
// error with an atomicMin on the wavelength's slot, then the peak whose error IS the minimum
// claims the slot with a CAS that also resets it (equal errors: either peak, same cost).
// Output per entry: sum of the clipped errors, match count, inlier count.
constexpr int MATCH_UNROLL = 5;  // the reference's default top_n_candidate
#ifndef RB2_C1_MIN_CTAS
#define RB2_C1_MIN_CTAS 4
#endif
template <int DEG_T, bool SINGLE = false>  // SINGLE: one problem in the call, its tables staged in shared memory
__global__ void __launch_bounds__(RS_THREADS, RB2_C1_MIN_CTAS)
ransac_match_kernel(const rb2_ransac_desc* __restrict__ descs, Pipe pipe, int max_wids, int max_up) {
  constexpr bool single = SINGLE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr unsigned long long INF_BITS = 0x7ff0000000000000ull;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned n = pipe.counts[1];
  const int qs = DEG_T > 0 ? DEG_T + 2 : pipe.qs, ncoef = DEG_T > 0 ? DEG_T + 1 : qs - 1;
  const size_t per_warp = (size_t)max_wids + max_up + (max_up + 1) / 2;  // in doubles
  unsigned long long* berr = reinterpret_cast<unsigned long long*>(smem_raw) + (size_t)warp * per_warp;
  double* s_be = reinterpret_cast<double*>(berr + max_wids);
  int* s_bw = reinterpret_cast<int*>(s_be + max_up);
  for (int i = lane; i < max_wids; i += 32) berr[i] = INF_BITS;  // +inf = empty
  // one problem in the call (the many-hypotheses case): its tables are staged in shared memory once
  // per CTA; a many-spectra batch reads them through L1 instead
  int n_u = 0;
  double tol = 0.0;
  const double* t_peaks = nullptr; const int* t_coff = nullptr; const double* t_cwave = nullptr; const int* t_cwid = nullptr;
  if (single) {
    const rb2_ransac_desc& d = descs[0];
    n_u = d.n_upeaks; tol = d.tol;
    const int n_c = d.n_cand;
    double* sp = reinterpret_cast<double*>(reinterpret_cast<unsigned long long*>(smem_raw) + (size_t)RS_WARPS * per_warp);
    double* sw = sp + n_u;
    int* so = reinterpret_cast<int*>(sw + n_c);
    int* si = so + (n_u + 1);
    for (int i = threadIdx.x; i < n_u; i += RS_THREADS) sp[i] = d.d_peaks[i];
    for (int i = threadIdx.x; i <= n_u; i += RS_THREADS) so[i] = d.d_cand_off[i];
    for (int i = threadIdx.x; i < n_c; i += RS_THREADS) { sw[i] = d.d_cand_wave[i]; si[i] = d.d_cand_wid[i]; }
    t_peaks = sp; t_coff = so; t_cwave = sw; t_cwid = si;
    __syncthreads();
  }
  __syncwarp();
  const unsigned total_warps = gridDim.x * RS_WARPS;
  if constexpr (SINGLE) {
    // One problem with at most 64 peaks (every single-spectrum fit): a lane always handles the same two peaks,
    // so their pixels and the first MATCH_UNROLL entries of their candidate lists stay in REGISTERS for the whole
    // kernel - an entry costs no table loads and no per-peak staging, only the polynomial, the comparisons and the
    // two 64-bit minima (half the instructions of the general loop below).  Same arithmetic, same order of the
    // per-lane error sums.
    if (n_u <= 64) {
      unsigned* const t_hi = reinterpret_cast<unsigned*>(berr);  // [max_wids] high words, 0xffffffff = empty
      unsigned* const t_lo = t_hi + max_wids;                    // [max_wids] low words
      for (int i = lane; i < 2 * max_wids; i += 32) t_hi[i] = 0xffffffffu;
      __syncwarp();
      const int u0 = lane, u1 = lane + 32;
      const bool v0 = u0 < n_u, v1 = u1 < n_u;
      const double x0 = v0 ? t_peaks[u0] : 0.0, x1 = v1 ? t_peaks[u1] : 0.0;
      const int ob0 = v0 ? t_coff[u0] : 0, oe0 = v0 ? t_coff[u0 + 1] : 0;
      const int ob1 = v1 ? t_coff[u1] : 0, oe1 = v1 ? t_coff[u1 + 1] : 0;
      double cw0[MATCH_UNROLL], cw1[MATCH_UNROLL];
      int wi0[MATCH_UNROLL], wi1[MATCH_UNROLL];
#pragma unroll
      for (int j = 0; j < MATCH_UNROLL; ++j) {  // a missing entry: +inf never is a (strict) minimum
        cw0[j] = (ob0 + j < oe0) ? t_cwave[ob0 + j] : CUDART_INF; wi0[j] = (ob0 + j < oe0) ? t_cwid[ob0 + j] : -1;
        cw1[j] = (ob1 + j < oe1) ? t_cwave[ob1 + j] : CUDART_INF; wi1[j] = (ob1 + j < oe1) ? t_cwid[ob1 + j] : -1;
      }
      auto nearest = [&](double fit, const double (&cw)[MATCH_UNROLL], const int (&wi)[MATCH_UNROLL], int ob, int oe,
                         double& be, int& bw) {
        be = CUDART_INF; bw = -1;
#pragma unroll
        for (int j = 0; j < MATCH_UNROLL; ++j) {
          const double er = fabs(sub(fit, cw[j]));
          if (er < be) { be = er; bw = wi[j]; }  // first minimum wins (calibrator.py:341-354)
        }
        for (int k = ob + MATCH_UNROLL; k < oe; ++k) {  // lists longer than the default top-N
          const double er = fabs(sub(fit, t_cwave[k]));
          if (er < be) { be = er; bw = t_cwid[k]; }
        }
      };
      double cn[deg_cap(DEG_T) + 1];  // the NEXT entry's coefficients: loaded one entry ahead (queue B is in L2 / DRAM)
      {
        const unsigned e0 = blockIdx.x * RS_WARPS + warp;
#pragma unroll
        for (int k = 0; k <= deg_cap(DEG_T); ++k) cn[k] = (k < ncoef && e0 < n) ? pipe.qb[(size_t)e0 * qs + k] : 0.0;
      }
      for (unsigned e = blockIdx.x * RS_WARPS + warp; e < n; e += total_warps) {
        double cr[deg_cap(DEG_T) + 1];
#pragma unroll
        for (int k = 0; k <= deg_cap(DEG_T); ++k) cr[k] = cn[k];
        {
          const unsigned en = e + total_warps;
          if (en < n) {
#pragma unroll
            for (int k = 0; k <= deg_cap(DEG_T); ++k) if (k < ncoef) cn[k] = pipe.qb[(size_t)en * qs + k];
          }
        }
        double be0, be1;
        int bw0, bw1;
        nearest(horner_ref<deg_cap(DEG_T)>(cr, ncoef, x0), cw0, wi0, ob0, oe0, be0, bw0);
        nearest(horner_ref<deg_cap(DEG_T)>(cr, ncoef, x1), cw1, wi1, ob1, oe1, be1, bw1);
        // 64-bit minimum per wavelength from NATIVE 32-bit shared atomics (a 64-bit shared atomicMin is a
        // compare-and-swap loop: it was a quarter of this path's instructions): errors are non-negative doubles, so
        // their bit patterns order like the values - minimum of the high words, then, among the peaks that hold it,
        // minimum of the low words; the peak that holds both claims the line with a CAS on the high word that also
        // resets the slot (equal errors: either peak, same cost).
        const unsigned hi0 = (unsigned)__double2hiint(be0), lo0 = (unsigned)__double2loint(be0);
        const unsigned hi1 = (unsigned)__double2hiint(be1), lo1 = (unsigned)__double2loint(be1);
        if (bw0 >= 0) atomicMin(&t_hi[bw0], hi0);
        if (bw1 >= 0) atomicMin(&t_hi[bw1], hi1);
        __syncwarp();
        const bool f0 = bw0 >= 0 && t_hi[bw0] == hi0, f1 = bw1 >= 0 && t_hi[bw1] == hi1;
        if (f0) atomicMin(&t_lo[bw0], lo0);
        if (f1) atomicMin(&t_lo[bw1], lo1);
        __syncwarp();
        double esum = 0.0;
        int nm = 0, ni = 0;
        if (f0 && t_lo[bw0] == lo0 && atomicCAS(&t_hi[bw0], hi0, 0xffffffffu) == hi0) {
          t_lo[bw0] = 0xffffffffu;
          double er = be0;
          if (er > tol) er = tol;  // M-SAC clip (:611)
          esum += er; ++nm; ni += (er < tol);
        }
        if (f1 && t_lo[bw1] == lo1 && atomicCAS(&t_hi[bw1], hi1, 0xffffffffu) == hi1) {
          t_lo[bw1] = 0xffffffffu;
          double er = be1;
          if (er > tol) er = tol;
          esum += er; ++nm; ni += (er < tol);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
        nm = __reduce_add_sync(0xffffffffu, nm);
        ni = __reduce_add_sync(0xffffffffu, ni);
        if (lane == 0) { pipe.res_esum[e] = esum; pipe.res_cnt[e] = make_int2(nm, ni); }
        __syncwarp();
      }
      return;
    }
  }
  for (unsigned e = blockIdx.x * RS_WARPS + warp; e < n; e += total_warps) {
    const double* q = pipe.qb + (size_t)e * qs;
    double cr[deg_cap(DEG_T) + 1];
#pragma unroll
    for (int k = 0; k <= deg_cap(DEG_T); ++k) cr[k] = (k < ncoef) ? q[k] : 0.0;
    if (!single) {
      const unsigned long long tag = (unsigned long long)__double_as_longlong(q[ncoef]);
      const rb2_ransac_desc& d = descs[(int)(tag >> 40)];
      n_u = d.n_upeaks; tol = d.tol;
      t_peaks = d.d_peaks; t_coff = d.d_cand_off; t_cwave = d.d_cand_wave; t_cwid = d.d_cand_wid;
    }
    const double* __restrict__ g_peaks = t_peaks;
    const int* __restrict__ g_coff = t_coff;
    const double* __restrict__ g_cwave = t_cwave;
    const int* __restrict__ g_cwid = t_cwid;
    // -- nearest candidate per peak (calibrator.py:341-354) --
    for (int u = lane; u < n_u; u += 32) {
      const double fit = horner_ref<deg_cap(DEG_T)>(cr, ncoef, g_peaks[u]);
      const int o0 = g_coff[u], o1 = g_coff[u + 1];
      double be = CUDART_INF;
      int bk = -1;
      // the first MATCH_UNROLL candidates as straight-line predicated code (top-N lists are that short: the loop
      // control was a sixth of this kernel's instructions), the rest in a loop
#pragma unroll
      for (int j = 0; j < MATCH_UNROLL; ++j) {
        const int k = o0 + j;
        if (k < o1) {
          const double er = fabs(sub(fit, g_cwave[k]));
          if (er < be) { be = er; bk = k; }  // first minimum wins
        }
      }
      for (int k = o0 + MATCH_UNROLL; k < o1; ++k) {
        const double er = fabs(sub(fit, g_cwave[k]));
        if (er < be) { be = er; bk = k; }
      }
      const int bw = (bk >= 0) ? g_cwid[bk] : -1;
      s_be[u] = be; s_bw[u] = bw;
      if (bw >= 0) atomicMin(&berr[bw], (unsigned long long)__double_as_longlong(be));
    }
    __syncwarp();
    double esum = 0.0;
    int nm = 0, ni = 0;
    for (int u = lane; u < n_u; u += 32) {
      const int bw = s_bw[u];
      if (bw < 0) continue;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(s_be[u]);
      if (berr[bw] != bits) continue;
      if (atomicCAS(&berr[bw], bits, INF_BITS) != bits) continue;  // claimed by an equal error
      double er = __longlong_as_double((long long)bits);
      if (er > tol) er = tol;  // M-SAC clip (:611)
      esum += er;
      ++nm;
      ni += (er < tol);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
    nm = __reduce_add_sync(0xffffffffu, nm);
    ni = __reduce_add_sync(0xffffffffu, ni);
    if (lane == 0) { pipe.res_esum[e] = esum; pipe.res_cnt[e] = make_int2(nm, ni); }
    __syncwarp();
  }
}

// ---- stage C2 --------------------------------------------------------------------------
// M-SAC cost of every matched entry, one THREAD per queue-B entry: regime test, Hough-density
// weight, cost, eligibility, per-problem running best (128-bit atomic min).
// Early skip: cost >= sum(err) / (m - deg) / (w_upper + 1e-9); when that bound already exceeds the
// problem's running best (or the hypothesis lacks the inliers to be eligible) the weight cannot
// change the outcome and only its regime test is run.  Selection and counters are unaffected
// (the `eligible` counter counts the inlier criteria only).
// The cost here uses stage C1's tree-summed errors; it ranks the hypotheses up to COST_MARGIN.  The
// final choice is made by stage C3 on exactly summed costs of the few that are that close to the best.
__device__ __forceinline__ double key_cost(unsigned long long k) {  // inverse of cost_key
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
__device__ __forceinline__ void best_update(BestKey* slot, const BestKey& mine) {
  BestKey cur;
  cur.hi = *reinterpret_cast<volatile unsigned long long*>(&slot->hi);
  cur.lo = *reinterpret_cast<volatile unsigned long long*>(&slot->lo);
  while (key_better(mine, cur)) {
    const BestKey old = atomicCAS(slot, cur, mine);
    if (old.hi == cur.hi && old.lo == cur.lo) break;
    cur = old;
  }
}

// status / counts of a scored hypothesis (parity outputs)
template <bool PLAIN = false>
__device__ __forceinline__ void cost_report(const rb2_ransac_desc& d, long long id, int nm, int ni) {
  if (PLAIN) return;
  if (d.d_out_status) d.d_out_status[id] = (uint8_t)RB2_HYP_SCORED;
  if (d.d_out_counts) { d.d_out_counts[2 * id] = nm; d.d_out_counts[2 * id + 1] = ni; }
}
// cost of a scored hypothesis from its (tree-summed) errors and its weight, eligibility, running best;
// returns the entry's RES_* flag
template <bool PLAIN = false>
__device__ __forceinline__ unsigned char cost_tail(const rb2_ransac_desc& d, Pipe& pipe, unsigned e, long long id,
                                                   int problem, int nm, int ni, double esum, double weight,
                                                   bool counts_ok, int ncoef) {
  cost_report<PLAIN>(d, id, nm, ni);
  const long long gid = d.hyp_begin + id;
  const double denom = (double)(nm - ncoef + 1);
  const double cost = esum / denom / (weight + 1e-9);  // :629-633
  bool elig = counts_ok && (cost <= d.cost_ceiling);
  bool near_floor = false;
  if (elig && d.floor_id >= 0) {  // fall-through: only what ranks after (floor_cost, floor_id), decided exactly in C3
    const double slack = COST_MARGIN * fabs(d.floor_cost);
    near_floor = fabs(cost - d.floor_cost) <= slack;
    elig = near_floor || cost > d.floor_cost;
  }
  if (!PLAIN && d.d_out_cost) d.d_out_cost[id] = cost;
  if (!PLAIN && d.d_out_weight) d.d_out_weight[id] = weight;
  if (!elig) return 0;
  pipe.res_cost[e] = cost;
  pipe.res_w[e] = weight;
  if (!near_floor) {
    BestKey mine;
    mine.hi = cost_key(cost); mine.lo = ~(unsigned long long)gid;
    best_update(pipe.best + problem, mine);
  }
  return near_floor ? RES_NEAR_FLOOR : RES_ELIGIBLE;
}

// ---- general Hough-density weight (calibrator.py:497-506, 614-627) -------------------
// twoditp(intercept, gradient) of the reference: RectBivariateSpline(gradient centres, intercept centres, hist)
// with both arguments clamped to its knot range (FITPACK bispev), evaluated at the tangent intercept and the
// gradient of the hypothesis at one pixel.  B-spline form: 4 x 4 coefficients times the non-zero cubic
// B-splines of either axis, each a stored polynomial on its knot interval.
__device__ __forceinline__ int bs_interval(const double* __restrict__ t, int n, double x) {
  int lo = 0, hi = n - 4;  // interval m: t[m + 3] <= x < t[m + 4]; the last one is closed
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (x >= t[mid + 3]) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ double bs_eval(const rb2_ransac_desc& d, double x, double y) {
  const int nx = d.bs_nx, ny = d.bs_ny;
  x = fmin(fmax(x, d.d_bs_tx[3]), d.d_bs_tx[nx]);
  y = fmin(fmax(y, d.d_bs_ty[3]), d.d_bs_ty[ny]);
  const int m = bs_interval(d.d_bs_tx, nx, x), l = bs_interval(d.d_bs_ty, ny, y);
  const double u = x - d.d_bs_tx[m + 3], v = y - d.d_bs_ty[l + 3];
  double by[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const double* k = d.d_bs_ppy + ((size_t)l * 4 + b) * 4;
    by[b] = fma(fma(fma(k[3], v, k[2]), v, k[1]), v, k[0]);
  }
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const double* k = d.d_bs_ppx + ((size_t)m * 4 + a) * 4;
    const double bx = fma(fma(fma(k[3], u, k[2]), u, k[1]), u, k[0]);
    const double* c = d.d_bs_c + (size_t)(m + a) * ny + l;
    s = fma(bx, fma(c[0], by[0], fma(c[1], by[1], fma(c[2], by[2], c[3] * by[3]))), s);
  }
  return s;
}

// one warp per queue-B entry flagged RES_GENERAL by stage C2: weight = hough_weight * sum over the pixels
template <int DEG_T>
__global__ void __launch_bounds__(256)
ransac_weight_general_kernel(const rb2_ransac_desc* __restrict__ descs, Pipe pipe) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned n = pipe.counts[1];
  const int qs = DEG_T > 0 ? DEG_T + 2 : pipe.qs, ncoef = DEG_T > 0 ? DEG_T + 1 : qs - 1, deg = ncoef - 1;
  const unsigned total_warps = gridDim.x * (blockDim.x >> 5);
  for (unsigned g = blockIdx.x * (blockDim.x >> 5) + warp; g * 32u < n; g += total_warps) {
    const unsigned e = g * 32u + lane;
    unsigned todo = __ballot_sync(0xffffffffu, e < n && (pipe.res_flag[e] & RES_GENERAL));
    while (todo) {
      const int b = __ffs(todo) - 1;
      todo &= todo - 1;
      const unsigned eb = g * 32u + b;
      const double* q = pipe.qb + (size_t)eb * qs;
      WeightCtx<DEG_T> w;
      constexpr int CAP = deg_cap(DEG_T);
#pragma unroll
      for (int k = 0; k <= CAP; ++k) w.c[k] = (k < ncoef) ? q[k] : 0.0;
#pragma unroll
      for (int k = 0; k <= CAP; ++k) w.dc[k] = (k < deg) ? mul((double)(k + 1), w.c[k + 1 <= CAP ? k + 1 : CAP]) : 0.0;
      const unsigned long long tag = (unsigned long long)__double_as_longlong(q[ncoef]);
      const int problem = (int)(tag >> 40);
      const long long id = (long long)(tag & TAG_ID_MASK);
      const rb2_ransac_desc& d = descs[problem];
      w.deg = deg; w.pix = d.d_pix;
      if (d.basis) {  // "gradient" = the monomial derivative rule on the basis coefficients (:614-618)
#pragma unroll
        for (int k = 0; k <= CAP; ++k) {
          double g = 0.0;
          if (k < deg)
            for (int j = 0; j < ncoef; ++j) g = fma(d.d_basis_grad[k * ncoef + j], w.c[j <= CAP ? j : CAP], g);
          w.dc[k] = g;
        }
      }
      double acc = 0.0;
      for (int i = lane; i < d.n_pix; i += 32) {
        const double x = w.px(i);
        const double grad = (deg >= 1) ? horner_ref<CAP>(w.dc, deg, x) : 0.0;
        acc += bs_eval(d, w.t_at(i), grad);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        const int2 cnt = pipe.res_cnt[eb];
        const bool counts_ok = (cnt.y >= d.min_inliers) && ((double)cnt.y >= d.min_inliers_frac);
        pipe.res_flag[eb] = cost_tail(d, pipe, eb, id, problem, cnt.x, cnt.y, pipe.res_esum[eb], d.hough_weight * acc,
                                      counts_ok, ncoef);
      }
    }
  }
}

template <int DEG_T, bool PLAIN = false>
__global__ void __launch_bounds__(128, 4)
ransac_cost_kernel(const rb2_ransac_desc* __restrict__ descs, Pipe pipe, rb2_ransac_result* __restrict__ result) {
  const unsigned n = pipe.counts[1];
  const int lane = threadIdx.x & 31;
  const int qs = DEG_T > 0 ? DEG_T + 2 : pipe.qs, ncoef = DEG_T > 0 ? DEG_T + 1 : qs - 1, deg = ncoef - 1;
  const unsigned total = gridDim.x * blockDim.x;
  for (unsigned e0 = blockIdx.x * blockDim.x + (threadIdx.x - lane); e0 < n; e0 += total) {
    const unsigned e = e0 + lane;
    int kind = -1;  // 0 empty, 1 scored, 2 scored + enough inliers, 3 unsupported
    int problem = -1;
    if (e < n) {
      const double* q = pipe.qb + (size_t)e * qs;
      double cr[deg_cap(DEG_T) + 1];
#pragma unroll
      for (int k = 0; k <= deg_cap(DEG_T); ++k) cr[k] = (k < ncoef) ? q[k] : 0.0;
      const unsigned long long tag = (unsigned long long)__double_as_longlong(q[ncoef]);
      problem = (int)(tag >> 40);
      const long long id = (long long)(tag & TAG_ID_MASK);
      const rb2_ransac_desc& d = descs[problem];
      const int2 cnt = pipe.res_cnt[e];
      const int nm = cnt.x, ni = cnt.y;
      const double esum = pipe.res_esum[e];
      unsigned char flag = 0;
      if (nm == 0) {  // :607-608
        kind = 0;
        if (!PLAIN && d.d_out_status) d.d_out_status[id] = (uint8_t)RB2_HYP_EMPTY;
      } else {
        const long long gid = d.hyp_begin + id;
        const bool counts_ok = (ni >= d.min_inliers) && ((double)ni >= d.min_inliers_frac);
        const bool wants_all = !PLAIN && (d.d_out_cost || d.d_out_weight);  // parity / diagnostics: every cost is reported
        bool need_value = wants_all || counts_ok;
        const double denom = (double)(nm - ncoef + 1);
        if (need_value && !wants_all && d.w_upper > 0.0) {
          const double lb = esum / denom / (d.w_upper + 1e-9) * (1.0 - COST_MARGIN);
          const unsigned long long best_hi = *reinterpret_cast<volatile unsigned long long*>(&pipe.best[problem].hi);
          if (cost_key(lb) > best_hi || lb > d.cost_ceiling) need_value = false;
        }
        double weight = 0.0;
        const bool ok = (!PLAIN && d.basis) ? false : hough_weight_thread<DEG_T>(d, cr, deg, need_value, &weight);
        if (!ok && d.d_bs_c) {
          // outside the corner regime: the weight is the clamped bicubic spline summed over the pixels -
          // warp-sized work, done by ransac_weight_general_kernel (which also finishes the cost)
          kind = counts_ok ? 2 : 1;
          flag = RES_GENERAL;
        } else if (!ok) {
          kind = 3;
          if (!PLAIN && d.d_out_status) d.d_out_status[id] = (uint8_t)RB2_HYP_UNSUPPORTED;
        } else {
          kind = counts_ok ? 2 : 1;
          if (need_value) flag = cost_tail<PLAIN>(d, pipe, e, id, problem, nm, ni, esum, weight, counts_ok, ncoef);
          else cost_report<PLAIN>(d, id, nm, ni);
        }
      }
      pipe.res_flag[e] = flag;
    }
    // counters (empty, scored, eligible, unsupported), aggregated over the lanes of one problem
    const unsigned have = __ballot_sync(0xffffffffu, kind >= 0);
    if (kind >= 0) {
      const unsigned peers = __match_any_sync(have, problem);
      const unsigned m_empty = __ballot_sync(peers, kind == 0), m_sc = __ballot_sync(peers, kind == 1 || kind == 2);
      const unsigned m_el = __ballot_sync(peers, kind == 2), m_un = __ballot_sync(peers, kind == 3);
      if ((peers & ((1u << lane) - 1u)) == 0) {
        add_counter(result + problem, 4, __popc(m_empty)); add_counter(result + problem, 5, __popc(m_sc));
        add_counter(result + problem, 6, __popc(m_el)); add_counter(result + problem, 7, __popc(m_un));
      }
    }
  }
}

// ---- stage C3 --------------------------------------------------------------------------
// The reference sums the clipped errors with Python's sequential sum() in ascending-wavelength order
// (calibrator.py:629); stage C1 sums them per lane and by a butterfly.  The few
// hypotheses whose C2 cost is within COST_MARGIN of the running best (or of the fall-through floor) are
// re-matched here, one warp each, and their errors added one by one in that order; the exact costs feed
// a second running best, which is the one stage D records.  Dynamic smem: u64 berr[max_wids] per warp.
template <int DEG_T>
__global__ void __launch_bounds__(RS_THREADS)
ransac_exact_kernel(const rb2_ransac_desc* __restrict__ descs, Pipe pipe, int max_wids) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr unsigned long long INF_BITS = 0x7ff0000000000000ull;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned n = pipe.counts[1];
  const int qs = DEG_T > 0 ? DEG_T + 2 : pipe.qs, ncoef = DEG_T > 0 ? DEG_T + 1 : qs - 1;
  unsigned long long* berr = reinterpret_cast<unsigned long long*>(smem_raw) + (size_t)warp * max_wids;
  for (int i = lane; i < max_wids; i += 32) berr[i] = INF_BITS;
  __syncwarp();
  const unsigned total_warps = gridDim.x * RS_WARPS;
  for (unsigned g = blockIdx.x * RS_WARPS + warp; g * 32u < n; g += total_warps) {
    const unsigned e = g * 32u + lane;
    bool contender = false;
    if (e < n) {
      const unsigned char flag = pipe.res_flag[e];
      if (flag & RES_NEAR_FLOOR) {
        contender = true;
      } else if (flag & RES_ELIGIBLE) {
        const unsigned long long tag = (unsigned long long)__double_as_longlong(pipe.qb[(size_t)e * qs + ncoef]);
        const unsigned long long bh = *reinterpret_cast<volatile unsigned long long*>(&pipe.best[(int)(tag >> 40)].hi);
        const double best = key_cost(bh);
        contender = pipe.res_cost[e] <= best + COST_MARGIN * fabs(best);
      }
    }
    unsigned todo = __ballot_sync(0xffffffffu, contender);
    while (todo) {
      const int b = __ffs(todo) - 1;
      todo &= todo - 1;
      const unsigned eb = g * 32u + b;
      const double* q = pipe.qb + (size_t)eb * qs;
      double cr[deg_cap(DEG_T) + 1];
#pragma unroll
      for (int k = 0; k <= deg_cap(DEG_T); ++k) cr[k] = (k < ncoef) ? q[k] : 0.0;
      const unsigned long long tag = (unsigned long long)__double_as_longlong(q[ncoef]);
      const int problem = (int)(tag >> 40);
      const rb2_ransac_desc& d = descs[problem];
      const int n_u = d.n_upeaks, n_w = d.n_wids;
      const double tol = d.tol;
      // nearest candidate per peak, best error per wavelength (calibrator.py:341-376)
      for (int u = lane; u < n_u; u += 32) {
        const double fit = horner_ref<deg_cap(DEG_T)>(cr, ncoef, d.d_peaks[u]);
        const int o0 = d.d_cand_off[u], o1 = d.d_cand_off[u + 1];
        double be = CUDART_INF;
        int bk = -1;
        for (int k = o0; k < o1; ++k) {
          const double er = fabs(sub(fit, d.d_cand_wave[k]));
          if (er < be) { be = er; bk = k; }
        }
        if (bk >= 0) atomicMin(&berr[d.d_cand_wid[bk]], (unsigned long long)__double_as_longlong(be));
      }
      __syncwarp();
      // sum(err) in ascending wavelength (= ascending id) order, one addition after the other
      double esum = 0.0;
      int nm = 0;
      for (int i0 = 0; i0 < n_w; i0 += 32) {
        const int i = i0 + lane;
        double er = 0.0;
        bool has = false;
        if (i < n_w) {
          const unsigned long long bits = berr[i];
          has = bits != INF_BITS;
          er = __longlong_as_double((long long)bits);
          if (er > tol) er = tol;  // M-SAC clip (:611)
          berr[i] = INF_BITS;
        }
        unsigned hm = __ballot_sync(0xffffffffu, has);
        nm += __popc(hm);
        while (hm) {
          const int src = __ffs(hm) - 1;
          hm &= hm - 1;
          esum = add(esum, __shfl_sync(0xffffffffu, er, src));
        }
      }
      __syncwarp();
      if (lane == 0) {
        const double cost = esum / (double)(nm - ncoef + 1) / (pipe.res_w[eb] + 1e-9);  // :629-633
        const long long gid = d.hyp_begin + (long long)(tag & TAG_ID_MASK);
        bool elig = cost <= d.cost_ceiling;
        if (elig && d.floor_id >= 0) elig = (cost > d.floor_cost) || (cost == d.floor_cost && gid < d.floor_id);
        if (d.d_out_cost) d.d_out_cost[tag & TAG_ID_MASK] = cost;
        if (elig) {
          pipe.res_cost[eb] = cost;
          pipe.res_flag[eb] = (unsigned char)(pipe.res_flag[eb] | RES_EXACT);
          BestKey mine;
          mine.hi = cost_key(cost); mine.lo = ~(unsigned long long)gid;
          best_update(pipe.best_exact + problem, mine);
        }
      }
    }
  }
}

// ---- stage D ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ransac_record_kernel(const rb2_ransac_desc* __restrict__ descs, Pipe pipe) {
  const unsigned n = pipe.counts[1];
  const int qs = pipe.qs, ncoef = qs - 1;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    if (!(pipe.res_flag[e] & RES_EXACT)) continue;
    const double cost = pipe.res_cost[e];
    const double* q = pipe.qb + (size_t)e * qs;
    const unsigned long long tag = (unsigned long long)__double_as_longlong(q[ncoef]);
    const int problem = (int)(tag >> 40);
    const long long gid = descs[problem].hyp_begin + (long long)(tag & TAG_ID_MASK);
    const BestKey b = pipe.best_exact[problem];
    if (b.hi != cost_key(cost) || b.lo != ~(unsigned long long)gid) continue;
    rb2_ransac_result* r = pipe.win + problem;
    r->cost = cost; r->hyp_id = gid;
    const int2 c = pipe.res_cnt[e];
    r->n_match = c.x; r->n_inliers = c.y;
    for (int k = 0; k < NCOEF; ++k) r->coeffs[k] = (k < ncoef) ? q[k] : 0.0;
  }
}

// one warp per problem: merge `count` records b[i * stride], b = records + p * base_mul
// (per-CTA records: base_mul = count, stride = 1; per-rank record sets [set][problem]:
//  base_mul = 1, stride = n_problems)
__global__ void ransac_finalize_kernel(const rb2_ransac_result* __restrict__ records, int blocks_per_problem,
                                       int base_mul, int stride, rb2_ransac_result* __restrict__ result) {
  const int p = blockIdx.x, lane = threadIdx.x;
  const rb2_ransac_result* b0 = records + (size_t)p * base_mul;
  auto rec = [&](int i) -> const rb2_ransac_result& { return b0[(size_t)i * stride]; };
  double bc = CUDART_INF;
  long long bid = -1;
  int bi = -1;
  unsigned long long cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = lane; i < blocks_per_problem; i += 32) {
    const double c = rec(i).cost;
    const long long id = rec(i).hyp_id;
    if (id >= 0 && ((c < bc) || (c == bc && id > bid) || bid < 0)) { bc = c; bid = id; bi = i; }
    for (int k = 0; k < 8; ++k) cnt[k] += rec(i).counters[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oc = __shfl_xor_sync(0xffffffffu, bc, o);
    const long long oid = __shfl_xor_sync(0xffffffffu, bid, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oid >= 0 && (bid < 0 || (oc < bc) || (oc == bc && oid > bid))) { bc = oc; bid = oid; bi = oi; }
    for (int k = 0; k < 8; ++k) cnt[k] += __shfl_xor_sync(0xffffffffu, cnt[k], o);
  }
  if (lane == 0) {
    rb2_ransac_result r;
    if (bi >= 0) {
      r = rec(bi);
    } else {
      r.cost = CUDART_INF; r.hyp_id = -1; r.n_match = 0; r.n_inliers = 0;
      for (int i = 0; i < NCOEF; ++i) r.coeffs[i] = 0.0;
    }
    for (int k = 0; k < 8; ++k) r.counters[k] = cnt[k];
    result[p] = r;
  }
}


}  // namespace rb2

namespace rb2 {

// hypotheses per chunk == queue capacity (worst case: all survive); RB2_PIPE_CHUNK_LOG2 overrides (tuning)
static unsigned pipe_chunk() {
  static const unsigned n = [] {
    const char* e = getenv("RB2_PIPE_CHUNK_LOG2");
    // measured on B200 (C3, 1e8 hypotheses, K3 alone): 2^23 4.03, 2^24 3.84, 2^25 3.61, 2^26 3.65 ms - a 1e8-hypothesis
    // call is then ONE chunk of 3.3e7 ids per lane (the queues of the three lanes take 18 GB of the 180 GB)
    const int v = e ? atoi(e) : 25;
    return 1u << (v < 12 ? 12 : (v > 26 ? 26 : v));
  }();
  return n;
}
#define PIPE_CHUNK (rb2::pipe_chunk())

// Two LANES: consecutive chunks alternate between two independent sets of queues on two streams,
// so that the issue-bound draw/fit stage of one chunk overlaps the latency-bound match / cost
// stages of the other.  The running best is shared; each lane records its own winners.
static int pipe_lanes() {
  static const int n = [] {
    const char* e = getenv("RB2_PIPE_LANES");
    const int v = e ? atoi(e) : 3;  // measured on B200 (C3, 1e8 hypotheses): 1 lane 1.37e10, 2: 1.52e10, 3: 1.60e10, 4: 1.58e10 hyp/s
    return v < 1 ? 1 : (v > PIPE_MAX_LANES ? PIPE_MAX_LANES : v);
  }();
  return n;
}
struct PipeLayout { size_t best, best_exact, lane0, lane_bytes, counts, win, qa, qb, res_cost, res_esum, res_cnt, res_w, res_flag, total; };

static PipeLayout pipe_layout(int n_problems) {
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  PipeLayout L;
  size_t o = 0;
  L.best = o; o = up(o + (size_t)n_problems * sizeof(BestKey));
  L.best_exact = o; o = up(o + (size_t)n_problems * sizeof(BestKey));
  const size_t lane0 = o;
  L.lane0 = lane0;
  L.counts = o - lane0; o = up(o + 16);
  L.win = o - lane0; o = up(o + (size_t)n_problems * sizeof(rb2_ransac_result));
  L.qa = o - lane0; o = up(o + ((size_t)PIPE_CHUNK + QA_SLACK) * (NCOEF + 1) * 8);
  L.qb = o - lane0; o = up(o + (size_t)PIPE_CHUNK * (NCOEF + 1) * 8);
  L.res_cost = o - lane0; o = up(o + (size_t)PIPE_CHUNK * 8);
  L.res_esum = o - lane0; o = up(o + (size_t)PIPE_CHUNK * 8);
  L.res_cnt = o - lane0; o = up(o + (size_t)PIPE_CHUNK * 8);
  L.res_w = o - lane0; o = up(o + (size_t)PIPE_CHUNK * 8);
  L.res_flag = o - lane0; o = up(o + (size_t)PIPE_CHUNK);
  L.lane_bytes = o - lane0;
  L.total = lane0 + (size_t)pipe_lanes() * L.lane_bytes;
  return L;
}

// extra streams + fork / join events, created once per device (one host thread per device drives
// rb2_ransac_batch: the lanes of concurrent calls on the same device would share these)
struct LaneStreams {
  cudaStream_t aux[PIPE_MAX_LANES - 1] = {};
  cudaEvent_t fork = nullptr, join[PIPE_MAX_LANES - 1] = {};
  bool ok = false;
};
static LaneStreams& lane_streams() {
  constexpr int MAX_DEV = 64;
  static LaneStreams per_dev[MAX_DEV];
  static LaneStreams none;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return none;
  LaneStreams& ls = per_dev[dev];
  if (!ls.ok) {
    bool good = cudaEventCreateWithFlags(&ls.fork, cudaEventDisableTiming) == cudaSuccess;
    for (int l = 0; l < PIPE_MAX_LANES - 1 && good; ++l)
      good = cudaStreamCreateWithFlags(&ls.aux[l], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&ls.join[l], cudaEventDisableTiming) == cudaSuccess;
    ls.ok = good;
  }
  return ls;
}

static size_t stage_a_smem(const rb2_ransac_desc& h) {
  return (size_t)h.n_upeaks * (sizeof(PeakRec) + 4 + 4) + (size_t)h.n_cand * 8 + (size_t)(1 << LUT_BITS) * 2 + 16;
}

static int knob(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : dflt;
  return v < 1 ? dflt : v;
}

// hypothesis ids per problem and chunk (see rb2_ransac_batch)
static long long chunk_len(int n_problems, long long max_hyp) {
  long long per_problem = max(1ll, (long long)(PIPE_CHUNK / (unsigned)n_problems));
  static const int per_lane = knob("RB2_CHUNKS_PER_LANE", 1), floor_log2 = knob("RB2_CHUNK_FLOOR_LOG2", 23);
  const long long want = (max_hyp + per_lane * pipe_lanes() - 1) / (per_lane * pipe_lanes());
  const long long floor_pp = max(1ll, (long long)((1u << floor_log2) / (unsigned)n_problems));
  return min(per_problem, max(want, floor_pp));
}

struct ChunkArgs {
  dim3 grid_a; size_t smem_a, smem_c, smem_x; int max_wids, max_up, single, sms; cudaStream_t stream;
  const rb2_ransac_desc* d_desc; Pipe pipe; long long chunk_begin, chunk_len; rb2_ransac_result* d_result;
  PhiloxKey key;
  int paged;  // stage A reserves queue slots page-wise (the grid's warps x 2 pages fit QA_SLACK)
  int plain;  // no problem replays draws, asks for per-hypothesis outputs or fits in another basis
  bool general_weight;  // some problem carries the spline tables of the general Hough-density weight
};

// Dynamic shared memory each kernel has been opted in for, per device.  The attribute is a CAP once set - a
// later launch that asks for more fails with "invalid argument" even below the 48 KB a kernel may use without any
// opt-in - so it only ever grows and never sits below 48 KB, and the record is kept per KERNEL: the match and the
// exact kernel are shared by the exactly determined and the least-squares instantiation of launch_chunk (a call
// sequence exact / over-determined / exact at one degree used to lower the cap under the third call's request).
template <int DEG_T> struct SmemOptIn { static size_t c[64], x[64]; };
template <int DEG_T> size_t SmemOptIn<DEG_T>::c[64] = {};
template <int DEG_T> size_t SmemOptIn<DEG_T>::x[64] = {};
constexpr size_t SMEM_NO_OPT_IN = 48 * 1024;

template <int DEG_T, bool LSQ = false>
static int launch_chunk(const ChunkArgs& a) {
  static size_t cfg_a[64] = {};  // ransac_draw_fit_kernel<DEG_T, LSQ, *>: kernels of this instantiation only
  size_t* const cfg_c = SmemOptIn<DEG_T>::c;
  size_t* const cfg_x = SmemOptIn<DEG_T>::x;
  int dev = 0;
  RB2_CHECK(cudaGetDevice(&dev));
  dev = (dev >= 0 && dev < 64) ? dev : 0;
  if (a.smem_a > cfg_a[dev]) {
    const size_t cap = max(a.smem_a, SMEM_NO_OPT_IN);
    RB2_CHECK((cudaFuncSetAttribute(ransac_draw_fit_kernel<DEG_T, LSQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap)));
    if (!LSQ && DEG_T > 0)
      RB2_CHECK((cudaFuncSetAttribute(ransac_draw_fit_kernel<DEG_T, LSQ, (!LSQ && DEG_T > 0)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap)));
    cfg_a[dev] = cap;
  }
  if (a.smem_c > cfg_c[dev]) {
    const size_t cap = max(a.smem_c, SMEM_NO_OPT_IN);
    RB2_CHECK(cudaFuncSetAttribute(ransac_match_kernel<DEG_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
    RB2_CHECK((cudaFuncSetAttribute(ransac_match_kernel<DEG_T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap)));
    cfg_c[dev] = cap;
  }
  if (a.smem_x > cfg_x[dev] && a.smem_x > SMEM_NO_OPT_IN) {
    RB2_CHECK(cudaFuncSetAttribute(ransac_exact_kernel<DEG_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a.smem_x));
    cfg_x[dev] = a.smem_x;
  }
  static const bool dbg = getenv("RB2_DEBUG_SYNC") != nullptr;  // diagnostics: name the stage that faults
  cudaStream_t stream = a.stream;
  auto stage_ok = [&](const char* name) {
    if (!dbg) return true;
    const cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { fprintf(stderr, "rb2: stage %s failed: %s\n", name, cudaGetErrorString(e)); return false; }
    return true;
  };
  const int sms = a.sms;
  // CTAs per SM of the small stages (tuning knobs; the defaults are the measured optimum)
  static const int nb = knob("RB2_B_CTAS_PER_SM", 8), nc1 = knob("RB2_C1_CTAS_PER_SM", 4), nc2 = knob("RB2_C2_CTAS_PER_SM", 8);
  pipe_reset_counts_kernel<<<1, 32, 0, stream>>>(a.pipe);
  if (!LSQ && DEG_T > 0 && a.plain)
    ransac_draw_fit_kernel<DEG_T, LSQ, (!LSQ && DEG_T > 0)><<<a.grid_a, RS_THREADS, a.smem_a, stream>>>(
        a.d_desc, a.pipe, a.chunk_begin, a.chunk_len, a.d_result, a.key, a.paged);
  else
    ransac_draw_fit_kernel<DEG_T, LSQ><<<a.grid_a, RS_THREADS, a.smem_a, stream>>>(a.d_desc, a.pipe, a.chunk_begin, a.chunk_len,
                                                                                   a.d_result, a.key, a.paged);
  if (!stage_ok("A draw_fit")) return RB2_ERR_CUDA;
  if (a.plain) ransac_monotonic_kernel<DEG_T, true><<<sms * nb, 256, 0, stream>>>(a.d_desc, a.pipe, a.d_result);
  else ransac_monotonic_kernel<DEG_T><<<sms * nb, 256, 0, stream>>>(a.d_desc, a.pipe, a.d_result);
  if (!stage_ok("B monotonic")) return RB2_ERR_CUDA;
  if (a.single) ransac_match_kernel<DEG_T, true><<<sms * nc1, RS_THREADS, a.smem_c, stream>>>(a.d_desc, a.pipe, a.max_wids, a.max_up);
  else ransac_match_kernel<DEG_T><<<sms * nc1, RS_THREADS, a.smem_c, stream>>>(a.d_desc, a.pipe, a.max_wids, a.max_up);
  if (!stage_ok("C1 match")) return RB2_ERR_CUDA;
  if (a.plain) ransac_cost_kernel<DEG_T, true><<<sms * nc2, 128, 0, stream>>>(a.d_desc, a.pipe, a.d_result);
  else ransac_cost_kernel<DEG_T><<<sms * nc2, 128, 0, stream>>>(a.d_desc, a.pipe, a.d_result);
  if (!stage_ok("C2 cost")) return RB2_ERR_CUDA;
  if (a.general_weight) {
    ransac_weight_general_kernel<DEG_T><<<sms * 4, 256, 0, stream>>>(a.d_desc, a.pipe);
    if (!stage_ok("C2g general weight")) return RB2_ERR_CUDA;
  }
  ransac_exact_kernel<DEG_T><<<sms * 2, RS_THREADS, a.smem_x, stream>>>(a.d_desc, a.pipe, a.max_wids);
  if (!stage_ok("C3 exact")) return RB2_ERR_CUDA;
  ransac_record_kernel<<<sms * 2, 256, 0, stream>>>(a.d_desc, a.pipe);
  if (!stage_ok("D record")) return RB2_ERR_CUDA;
  return RB2_OK;
}

}  // namespace rb2

extern "C" size_t rb2_ransac_scratch_bytes(int n_problems, int blocks_per_problem) {
  (void)blocks_per_problem;
  if (n_problems <= 0) return 0;
  return rb2::pipe_layout(n_problems).total;
}

extern "C" int rb2_ransac_chunk_plan(int n_problems, long long max_hyp, long long* per_problem, int* n_chunks, int* n_lanes) {
  if (n_problems <= 0 || max_hyp < 0 || (unsigned)n_problems > PIPE_CHUNK) return RB2_ERR_ARG;
  const long long pp = rb2::chunk_len(n_problems, max_hyp);
  const long long nc = (max_hyp + pp - 1) / pp;
  if (per_problem) *per_problem = pp;
  if (n_chunks) *n_chunks = (int)nc;
  if (n_lanes) *n_lanes = (int)(nc > 1 ? (nc < rb2::pipe_lanes() ? nc : rb2::pipe_lanes()) : 1);
  return RB2_OK;
}

extern "C" int rb2_ransac_batch(int n_problems, const rb2_ransac_desc* d_desc, const rb2_ransac_desc* h_desc,
                                rb2_ransac_result* d_result, void* d_block_scratch, int blocks_per_problem,
                                void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_desc || !h_desc || !d_result || !d_block_scratch) return RB2_ERR_ARG;
  if ((unsigned)n_problems > PIPE_CHUNK || n_problems >= (1 << 23)) return RB2_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  int sms = 148;
  if (rb2_sm_count(&sms) != RB2_OK) return RB2_ERR_NO_DEVICE;
  // all problems of one call share degree / sample size (one kernel instantiation) and the Philox key
  const int deg = h_desc[0].deg, s = h_desc[0].sample_size;
  bool fast = true, general_weight = false, plain = true;
  size_t smem_a = 0;
  int max_wids = 1, max_up = 1;
  long long max_hyp = 0;
  for (int i = 0; i < n_problems; ++i) {
    const rb2_ransac_desc& h = h_desc[i];
    if (h.deg != deg || h.sample_size != s || h.seed != h_desc[0].seed) return RB2_ERR_ARG;
    if (h.deg < 1 || h.deg > RB2_MAX_DEG || h.sample_size < h.deg + 1) return RB2_ERR_ARG;
    if (h.sample_size + h.n_known > RB2_MAX_SAMPLE) return RB2_ERR_ARG;
    if (h.n_upeaks < h.sample_size && !h.d_replay_x) return RB2_ERR_ARG;
    if (h.sample_size != h.deg + 1 || h.n_known != 0) fast = false;
    if (h.n_cand < h.n_upeaks || h.n_wids < 1 || h.n_upeaks > 65535 || h.n_cand > A_MAX_CAND) return RB2_ERR_ARG;
    if (h.n_hyp < 0 || h.n_hyp >= (1ll << 40)) return RB2_ERR_ARG;
    if (h.basis && (h.basis < 0 || h.basis > 2 || !h.d_bs_c || !h.d_basis_c0 || !h.d_basis_grad)) return RB2_ERR_ARG;
    if (h.d_bs_c) {
      if (h.bs_nx < 4 || h.bs_ny < 4 || !h.d_bs_tx || !h.d_bs_ty || !h.d_bs_ppx || !h.d_bs_ppy) return RB2_ERR_ARG;
      general_weight = true;
    }
    if (h.d_replay_x || h.d_out_status || h.d_out_coeffs || h.d_out_sample || h.d_out_cost || h.d_out_weight || h.d_out_counts ||
        h.basis) plain = false;
    smem_a = max(smem_a, stage_a_smem(h));
    max_wids = max(max_wids, h.n_wids);
    max_up = max(max_up, h.n_upeaks);
    max_hyp = max(max_hyp, (long long)h.n_hyp);
  }
  const int single = (n_problems == 1);
  const size_t smem_x = (size_t)RS_WARPS * max_wids * 8;
  const PhiloxKey key = philox_key(h_desc[0].seed);
  size_t smem_c = (size_t)RS_WARPS * ((size_t)max_wids + max_up + (max_up + 1) / 2) * 8;
  if (single) smem_c += (size_t)h_desc[0].n_upeaks * 8 + (size_t)h_desc[0].n_cand * 12 + (size_t)(h_desc[0].n_upeaks + 1) * 4 + 16;
  if (smem_a > 200 * 1024 || smem_c > 200 * 1024 || smem_x > 200 * 1024) return RB2_ERR_ARG;

  const PipeLayout L = pipe_layout(n_problems);
  unsigned char* base = (unsigned char*)d_block_scratch;
  PipeSet ps;
  ps.n = pipe_lanes();
  for (int l = 0; l < ps.n; ++l) {
    unsigned char* lb = base + L.lane0 + (size_t)l * L.lane_bytes;
    Pipe& pipe = ps.p[l];
    pipe.best = (BestKey*)(base + L.best);
    pipe.best_exact = (BestKey*)(base + L.best_exact);
    pipe.res_w = (double*)(lb + L.res_w);
    pipe.res_flag = (unsigned char*)(lb + L.res_flag);
    pipe.counts = (unsigned*)(lb + L.counts);
    pipe.win = (rb2_ransac_result*)(lb + L.win);
    pipe.qa = (double*)(lb + L.qa);
    pipe.qb = (double*)(lb + L.qb);
    pipe.res_cost = (double*)(lb + L.res_cost);
    pipe.res_esum = (double*)(lb + L.res_esum);
    pipe.res_cnt = (int2*)(lb + L.res_cnt);
    pipe.cap = PIPE_CHUNK;
    pipe.cap_a = PIPE_CHUNK + QA_SLACK;
    pipe.qs = deg + 2;
  }
  pipe_init_kernel<<<(n_problems + 255) / 256, 256, 0, stream>>>(ps, d_result, n_problems);

  // chunk = hypothesis ids per problem and launch.  Large calls use the full queue capacity (fewer, longer
  // kernels); a call that would fit in a few chunks (a rank's share of a sharded fit, a batch of spectra with
  // 1e4 hypotheses each) is cut into one chunk per lane, but not below 2^23 hypotheses in total: below that the
  // ramp-up, tail and per-CTA table set-up of the seven kernels of a chunk cost more than the overlap of
  // consecutive chunks gains (measured on B200, C3, K3 alone: 1.25e7 hypotheses 0.79 -> 0.68 ms, 2.5e7 1.32 -> 1.23 ms,
  // 5e7 2.39 -> 2.34 ms against two chunks per lane with a 2^21 floor)
  const long long per_problem = rb2::chunk_len(n_problems, max_hyp);
  const long long n_chunks = (max_hyp + per_problem - 1) / per_problem;
  LaneStreams& ls = lane_streams();
  const int n_lanes = (ls.ok && n_chunks > 1) ? (int)min((long long)ps.n, n_chunks) : 1;
  if (n_lanes > 1) {  // fork: the extra streams start after everything already queued on the caller's stream
    RB2_CHECK(cudaEventRecord(ls.fork, stream));
    for (int l = 1; l < n_lanes; ++l) RB2_CHECK(cudaStreamWaitEvent(ls.aux[l - 1], ls.fork, 0));
  }
  long long chunk_index = 0;
  for (long long c0 = 0; c0 < max_hyp; c0 += per_problem, ++chunk_index) {
    const long long len = min(per_problem, max_hyp - c0);
    const int lane_id_ = (int)(chunk_index % n_lanes);
    const Pipe& pipe = ps.p[lane_id_];
    cudaStream_t cstream = lane_id_ ? ls.aux[lane_id_ - 1] : stream;
    int bpp = blocks_per_problem;
    if (bpp <= 0) {
      const long long want = (len + RS_THREADS * 8 - 1) / (RS_THREADS * 8);   // >= 8 passes per CTA
      static const int ctas_per_sm = knob("RB2_A_CTAS_PER_SM", 3);  // 3 resident CTAs leave registers for the small stages of the other lanes (measured: 2: 1.71, 3: 1.78, 4: 1.74, 8: 1.68 e10 hyp/s)
      const long long fill = max(1ll, (long long)(ctas_per_sm * sms) / n_problems);  // CTAs per SM overall
      bpp = (int)max(1ll, min(want, fill));
    }
    ChunkArgs ca;
    ca.grid_a = dim3(bpp, n_problems); ca.smem_a = smem_a; ca.smem_c = smem_c; ca.smem_x = smem_x;
    ca.max_wids = max_wids; ca.max_up = max_up; ca.single = single; ca.sms = sms; ca.stream = cstream;
    static const bool no_pages = getenv("RB2_A_UNPAGED") != nullptr;  // A/B knob
    // paged only where the pages' empty ends (<= 2 pages per warp) are few next to the survivors (~16 % of the chunk):
    // a rank's share of a sharded fit and batches of small problems keep the per-iteration atomic
    const unsigned long long page_slots = (unsigned long long)bpp * n_problems * RS_WARPS * 2ull * QA_PAGE;
    ca.paged = (!no_pages && page_slots <= QA_SLACK && (unsigned long long)len * n_problems >= 48ull * page_slots) ? 1 : 0;
    ca.d_desc = d_desc; ca.pipe = pipe; ca.chunk_begin = c0; ca.chunk_len = len; ca.d_result = d_result; ca.key = key; ca.general_weight = general_weight; ca.plain = plain ? 1 : 0;
    int rc;
    switch ((fast ? 0 : 8) + (deg <= 5 ? deg : 0)) {
      case 1: rc = launch_chunk<1>(ca); break;
      case 2: rc = launch_chunk<2>(ca); break;
      case 3: rc = launch_chunk<3>(ca); break;
      case 4: rc = launch_chunk<4>(ca); break;
      case 5: rc = launch_chunk<5>(ca); break;
      // over-determined samples / known pairs: compile-time degree everywhere, least squares in stage A
      case 9: rc = launch_chunk<1, true>(ca); break;
      case 10: rc = launch_chunk<2, true>(ca); break;
      case 11: rc = launch_chunk<3, true>(ca); break;
      case 12: rc = launch_chunk<4, true>(ca); break;
      case 13: rc = launch_chunk<5, true>(ca); break;
      default: rc = launch_chunk<0>(ca); break;
    }
    if (rc != RB2_OK) return rc;
  }
  for (int l = 1; l < n_lanes; ++l) {  // join: the caller's stream continues after every lane
    RB2_CHECK(cudaEventRecord(ls.join[l - 1], ls.aux[l - 1]));
    RB2_CHECK(cudaStreamWaitEvent(stream, ls.join[l - 1], 0));
  }
  pipe_join_kernel<<<(n_problems + 255) / 256, 256, 0, stream>>>(ps, d_result, n_problems);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

// Merge n_sets record sets (layout [set][problem], e.g. the all-gathered per-rank winners or
// the winners of successive launches) with the same rule: lower cost, later id on ties;
// counters are summed.  d_out may alias set 0 only if n_sets == 1.
extern "C" int rb2_merge_winners(int n_problems, int n_sets, const rb2_ransac_result* d_sets,
                                 rb2_ransac_result* d_out, void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || n_sets <= 0 || !d_sets || !d_out) return RB2_ERR_ARG;
  ransac_finalize_kernel<<<n_problems, 32, 0, (cudaStream_t)stream_>>>(d_sets, n_sets, 1, n_problems, d_out);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}
