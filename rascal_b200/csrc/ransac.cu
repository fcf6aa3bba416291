// K3 — batched RANSAC hypothesis draw / fit / test / score / reduce (sm_100a).
//
// Replaces the loop body of Calibrator._solve_candidate_ransac
// (calibrator.py:525-735).  One hypothesis = one minimal sample drawn and
// fitted (= one polyfit call of the reference).
//
// Mapping.  A CTA serves one problem (one spectrum); its read-only tables
// (<= a few KB: peaks, candidate lists, sampling CDF) live in shared memory.
// Work is staged so that the three very different cost classes never diverge
// inside a warp:
//   stage A  (all 32 lanes, one hypothesis per lane)  counter-based Philox
//            draw -> exact interpolation (Newton divided differences, one
//            FP64 division per hypothesis) or least squares (Householder QR)
//            -> intercept test.  ~85 % of hypotheses end here.  Survivors
//            are pushed into a per-warp shared-memory ring.
//   stage B  (32 lanes, one queued hypothesis per lane, run when >= 32 are
//            queued) monotonicity test: instead of the reference's G ~ 1.4 *
//            ptp(peaks) point grid, locate the critical points of
//            q(k) = p(t_k + 1) - p(t_k) and evaluate the reference's own
//            Horner arithmetic only at the grid points around them.
//   stage C  (whole warp, one hypothesis at a time) bijective match: lanes
//            take peaks, shared-memory atomicMin de-duplicates by wavelength
//            id; M-SAC cost; Hough-density weight by locating the threshold
//            crossings of the tangent-intercept polynomial with warp-parallel
//            32-ary searches over the pixel list instead
//            of evaluating a bicubic spline at every detector pixel.
// Selection rule (App. A.4): minimal cost among eligible hypotheses, the
// LATEST hypothesis id among equal costs; per-warp -> per-CTA -> per-problem.
#include "common.cuh"
#include <stdio.h>

namespace rb2 {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int Q_CAP = 64;  // ring entries per warp
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

struct Tables {  // shared-memory views of one problem
  const double* peaks;     // [n_u]
  const double* cwave;     // [n_c]
  const int* coff;         // [n_u+1]
  const int* cwid;         // [n_c]
  const unsigned* cdf32;   // [n_u]
  int n_u, n_c, n_wids;
};

// first i with r < cdf32[i]  (== searchsorted(cdf, u, 'right')), clamped
__device__ __forceinline__ int draw_peak(const Tables& t, unsigned r) {
  int a = 0, b = t.n_u - 1;
  while (a < b) {
    const int m = (a + b) >> 1;
    if (r < t.cdf32[m]) b = m; else a = m + 1;
  }
  return a;
}

// ---- real roots of a polynomial of degree n <= 6 inside [lo, hi] -------------
// Accuracy target is a fraction of a grid step: the callers re-check the grid
// points around every root with the reference's arithmetic.
__device__ __forceinline__ double horner_fma(const double* a, int n, double x) {
  double r = a[n];
  for (int i = n - 1; i >= 0; --i) r = fma(r, x, a[i]);
  return r;
}

__device__ int quad_roots(double a, double b, double c, double lo, double hi, double* out) {
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

// roots of sum_{i<=n} a[i] x^i in (lo, hi), ascending, via the derivative chain:
// between consecutive roots of f' the function is monotone -> bisection.
__device__ int real_roots(const double* a, int n, double lo, double hi, double* out) {
  while (n > 0 && a[n] == 0.0) --n;
  if (n <= 0) return 0;
  if (n <= 2) return quad_roots(n == 2 ? a[2] : 0.0, a[1], a[0], lo, hi, out);
  // derivative-chain polynomials: D_k = (n-k)-th derivative scaled, degree k
  double prev[RB2_MAX_DEG], cur[RB2_MAX_DEG];
  int nprev;
  {
    // degree-2 member of the chain: d^(n-2) f / dx^(n-2)
    double c2[3];
    for (int i = 0; i <= 2; ++i) {
      double f = 1.0;
      for (int j = 0; j < n - 2; ++j) f *= (double)(i + n - 2 - j);
      c2[i] = a[i + n - 2] * f;
    }
    nprev = quad_roots(c2[2], c2[1], c2[0], lo, hi, prev);
  }
  for (int k = 3; k <= n; ++k) {
    // coefficients of the (n-k)-th derivative, degree k
    double ck[RB2_MAX_DEG + 1];
    for (int i = 0; i <= k; ++i) {
      double f = 1.0;
      for (int j = 0; j < n - k; ++j) f *= (double)(i + n - k - j);
      ck[i] = a[i + n - k] * f;
    }
    int ncur = 0;
    double left = lo, fl = horner_fma(ck, k, left);
    for (int s = 0; s <= nprev; ++s) {
      const double right = (s < nprev) ? prev[s] : hi;
      const double fr = horner_fma(ck, k, right);
      if ((fl < 0.0) != (fr < 0.0) && right > left) {
        double xa = left, xb = right, fa = fl;
        for (int it = 0; it < 40 && (xb - xa) > 0.125; ++it) {
          const double xm = 0.5 * (xa + xb);
          const double fm = horner_fma(ck, k, xm);
          if ((fm < 0.0) == (fa < 0.0)) { xa = xm; fa = fm; } else { xb = xm; }
        }
        cur[ncur++] = 0.5 * (xa + xb);
      }
      left = right; fl = fr;
    }
    nprev = ncur;
    for (int i = 0; i < ncur; ++i) prev[i] = cur[i];
  }
  for (int i = 0; i < nprev; ++i) out[i] = prev[i];
  return nprev;
}

// ---- monotonicity test (calibrator.py:585-600) -------------------------------
// all(diff(polyval(arange(pix_min, pix_max, 1), c)) > 0)
__device__ bool monotonic_test(const double* c, int deg, double pix_min, double pix_max) {
  const double gl = ceil((pix_max - pix_min) / 1.0);  // numpy arange length
  if (!(gl >= 2.0)) return true;                      // diff of <2 points is empty
  const long long G = (long long)gl;
  const long long kmax = G - 2;                       // D_k for k in [0, kmax]
  auto D = [&](long long k) {
    const double t0 = add(pix_min, (double)k), t1 = add(pix_min, (double)(k + 1));
    return sub(horner_unfused(c, deg + 1, t1), horner_unfused(c, deg + 1, t0));
  };
  if (!(D(0) > 0.0) || !(D(kmax) > 0.0)) return false;
  if (deg <= 2) return true;  // q is linear: extremes at the ends
  // p~(tau) = p(pix_min + tau): Taylor shift
  double ps[NCOEF];
  for (int i = 0; i <= deg; ++i) ps[i] = c[i];
  for (int i = 0; i < deg; ++i)
    for (int j = deg - 1; j >= i; --j) ps[j] = fma(pix_min, ps[j + 1], ps[j]);
  // q(tau) = p~(tau+1) - p~(tau) = sum_i tau^i sum_{j>i} C(j,i) p~_j ; need q' (degree deg-2)
  double q[NCOEF];
  for (int i = 0; i < deg; ++i) {
    double s = 0.0, binom = 1.0;  // C(j,i) for j = i.. : C(i,i)=1
    for (int j = i + 1; j <= deg; ++j) {
      binom = binom * (double)j / (double)(j - i);  // C(j,i)
      s = fma(binom, ps[j], s);
    }
    q[i] = s;
  }
  double dq[NCOEF];
  for (int i = 1; i < deg; ++i) dq[i - 1] = (double)i * q[i];
  double roots[RB2_MAX_DEG];
  const int nr = real_roots(dq, deg - 2, 0.0, (double)kmax, roots);
  for (int r = 0; r < nr; ++r) {
    const long long kf = (long long)floor(roots[r]);
    for (long long k = kf - 1; k <= kf + 2; ++k) {
      if (k < 0 || k > kmax) continue;
      if (!(D(k) > 0.0)) return false;
    }
  }
  return true;
}

// ---- warp-parallel search: first index in [a, b] where pred(i) is true ------
// pred must be monotone (false...false true...true); returns b+1 if none.
template <class Pred>
__device__ __forceinline__ int warp_first_true(int a, int b, Pred&& pred) {
  const int lane = lane_id();
  int lo = a, hi = b + 1;  // answer in [lo, hi]
  while (hi > lo) {
    const int width = hi - lo;
    if (width <= 32) {
      const int i = lo + lane;
      const bool p = (i < hi) ? pred(i) : true;
      const unsigned bal = __ballot_sync(0xffffffffu, p);
      return bal ? lo + (__ffs(bal) - 1) : hi;  // lanes >= width vote true -> hi; width == 32 and none true -> hi
    }
    const int step = (width + 31) / 32;
    int i = lo + (lane + 1) * step - 1;
    if (i >= hi) i = hi - 1;
    const bool p = pred(i);
    const unsigned bal = __ballot_sync(0xffffffffu, p);
    if (bal == 0u) return hi;  // even the last probe (hi-1) is false
    const int L = __ffs(bal) - 1;
    const int probeL = min(lo + (L + 1) * step - 1, hi - 1);
    const int probePrev = (L > 0) ? min(lo + L * step - 1, hi - 1) : lo - 1;
    hi = probeL;  // pred(probeL) true -> answer <= probeL
    lo = probePrev + 1;
  }
  return lo;
}

struct WeightCtx {
  const double* c;   // coefficients (smem)
  double dc[NCOEF];  // derivative coefficients i*c[i]
  int deg;
  const double* pix; // global or NULL
  __device__ __forceinline__ double px(int i) const { return pix ? pix[i] : (double)i; }
  // tangent intercept t(x) = p(x) - x p'(x), the reference's arithmetic (:614-618)
  __device__ __forceinline__ double t_at(int i) const {
    const double x = px(i);
    const double wave = horner_unfused(c, deg + 1, x);
    const double grad = (deg >= 1) ? horner_unfused(dc, deg, x) : 0.0;
    return sub(wave, mul(grad, x));
  }
};

__device__ __forceinline__ double pp_eval(const rb2_ransac_desc& d, double t) {
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

// Hough-density weight of one hypothesis (whole warp).  Returns false when the
// hypothesis leaves the corner-clamped regime (gradient above the first
// intercept-bin centre somewhere on the detector).
__device__ bool hough_weight_warp(const rb2_ransac_desc& d, const double* c, int deg, double* out_weight) {
  const int lane = lane_id();
  const int np = d.n_pix;
  if (np <= 0) { *out_weight = 0.0; return true; }
  WeightCtx w;
  w.c = c; w.deg = deg; w.pix = d.d_pix;
  for (int i = 1; i <= deg; ++i) w.dc[i - 1] = mul((double)i, c[i]);
  const double x0 = w.px(0), x1 = w.px(np - 1);
  // breakpoints of t: t'(x) = -x p''(x)
  double brk[RB2_MAX_DEG + 1];
  int nb = 0;
  double p2[NCOEF];
  for (int i = 2; i <= deg; ++i) p2[i - 2] = (double)(i * (i - 1)) * c[i];
  if (deg >= 3) nb = real_roots(p2, deg - 2, x0, x1, brk);
  // regime check: max of p'(x) over [x0, x1] must stay <= first intercept-bin centre
  {
    double gmax = fmax(horner_fma(w.dc, deg - 1 >= 0 ? deg - 1 : 0, x0), horner_fma(w.dc, deg - 1 >= 0 ? deg - 1 : 0, x1));
    if (deg == 0) gmax = 0.0;
    for (int r = 0; r < nb; ++r) gmax = fmax(gmax, horner_fma(w.dc, deg - 1, brk[r]));
    if (!(gmax <= d.ic_min)) return false;
  }
  if (x0 < 0.0 && x1 > 0.0) {  // x = 0 is a breakpoint too
    int k = nb;
    while (k > 0 && brk[k - 1] > 0.0) { brk[k] = brk[k - 1]; --k; }
    brk[k] = 0.0;
    ++nb;
  }
  const double th_hi = d.gc_max, th_lo = d.gc_min;
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
    const int ca = cls(ia), cb = cls(ib);
    int i_mid0, i_mid1;  // MID pixels are [i_mid0, i_mid1)
    if (ca == cb) {
      if (ca == 2) n_hi += ib - ia + 1; else if (ca == 0) n_lo += ib - ia + 1;
      i_mid0 = ia; i_mid1 = (ca == 1) ? ib + 1 : ia;
    } else if (ca < cb) {  // increasing: LO.. MID.. HI
      const int f1 = warp_first_true(ia, ib, [&](int i) { return cls(i) >= 1; });
      const int f2 = warp_first_true(f1, ib, [&](int i) { return cls(i) >= 2; });
      n_lo += f1 - ia; n_hi += ib + 1 - f2;
      i_mid0 = f1; i_mid1 = f2;
    } else {               // decreasing: HI.. MID.. LO
      const int f1 = warp_first_true(ia, ib, [&](int i) { return cls(i) <= 1; });
      const int f2 = warp_first_true(f1, ib, [&](int i) { return cls(i) <= 0; });
      n_hi += f1 - ia; n_lo += ib + 1 - f2;
      i_mid0 = f1; i_mid1 = f2;
    }
#ifdef RB2_DEBUG_WEIGHT
    if (lane == 0) printf("piece %d: ia %d ib %d ca %d cb %d mid [%d,%d) n_hi %lld n_lo %lld nb %d brk0 %g\n", pc, ia, ib, ca, cb, i_mid0, i_mid1, n_hi, n_lo, nb, nb ? brk[0] : -1.0);
#endif
    for (int i = i_mid0 + lane; i < i_mid1; i += 32) {  // class re-checked per pixel: never extrapolate the pp-form
      const double t = w.t_at(i);
      mid += (t >= th_hi) ? d.h_hi : ((t <= th_lo) ? d.h_lo : pp_eval(d, t));
    }
    ia = ib + 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mid += __shfl_xor_sync(0xffffffffu, mid, o);
  *out_weight = d.hough_weight * ((double)n_hi * d.h_hi + (double)n_lo * d.h_lo + mid);
  return true;
}

struct Best {
  double cost;
  long long id;
  int n_match, n_inl;
  __device__ __forceinline__ bool worse_than(double c, long long i) const {
    return (c < cost) || (c == cost && i > id);
  }
};

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

template <int DEG_T>
__global__ void __launch_bounds__(RS_THREADS)
ransac_kernel(const rb2_ransac_desc* __restrict__ descs, rb2_ransac_result* __restrict__ block_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ rb2_ransac_desc s_desc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    static_assert(sizeof(rb2_ransac_desc) % 4 == 0, "descriptor must be word-sized");
    const unsigned* src = reinterpret_cast<const unsigned*>(descs + blockIdx.y);
    unsigned* dst = reinterpret_cast<unsigned*>(&s_desc);
    for (int i = tid; i < (int)(sizeof(rb2_ransac_desc) / 4); i += RS_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const rb2_ransac_desc& d = s_desc;
  const int deg = DEG_T > 0 ? DEG_T : d.deg;
  const int ncoef = deg + 1;
  const int s = DEG_T > 0 ? DEG_T + 1 : d.sample_size;
  const int n_u = d.n_upeaks, n_wids = d.n_wids;
  const int n_c = d.n_cand;

  // ---- shared memory carve-up ----
  double* s_peaks = reinterpret_cast<double*>(smem_raw);
  double* s_cwave = s_peaks + n_u;
  double* s_queue = s_cwave + n_c;                                    // [RS_WARPS][Q_CAP][ncoef+1]
  unsigned long long* s_berr = reinterpret_cast<unsigned long long*>(s_queue + (size_t)RS_WARPS * Q_CAP * (ncoef + 1));  // [RS_WARPS][n_wids]
  int* s_coff = reinterpret_cast<int*>(s_berr + (size_t)RS_WARPS * n_wids);
  int* s_cwid = s_coff + (n_u + 1);
  unsigned* s_cdf = reinterpret_cast<unsigned*>(s_cwid + n_c);
  __shared__ unsigned long long s_cnt[8];
  __shared__ Best s_best[RS_WARPS];
  __shared__ double s_bcoef[RS_WARPS][NCOEF];

  for (int i = tid; i < n_u; i += RS_THREADS) { s_peaks[i] = d.d_peaks[i]; s_cdf[i] = d.d_cdf32 ? d.d_cdf32[i] : 0u; }
  for (int i = tid; i <= n_u; i += RS_THREADS) s_coff[i] = d.d_cand_off[i];
  for (int i = tid; i < n_c; i += RS_THREADS) { s_cwave[i] = d.d_cand_wave[i]; s_cwid[i] = d.d_cand_wid[i]; }
  if (tid < 8) s_cnt[tid] = 0ull;
  __syncthreads();

  Tables tb;
  tb.peaks = s_peaks; tb.cwave = s_cwave; tb.coff = s_coff; tb.cwid = s_cwid; tb.cdf32 = s_cdf;
  tb.n_u = n_u; tb.n_c = n_c; tb.n_wids = n_wids;

  const int qstride = ncoef + 1;
  double* q = s_queue + (size_t)warp * Q_CAP * qstride;
  unsigned long long* berr = s_berr + (size_t)warp * n_wids;
  int q_head = 0, q_count = 0;  // warp-uniform

  Best best;
  best.cost = CUDART_INF; best.id = -1; best.n_match = 0; best.n_inl = 0;
  double best_coef[NCOEF];
#pragma unroll
  for (int i = 0; i < NCOEF; ++i) best_coef[i] = 0.0;
  unsigned c_drawn = 0, c_abort = 0, c_icpt = 0, c_mono = 0, c_empty = 0, c_scored = 0, c_elig = 0, c_unsup = 0;

  const bool replay = d.d_replay_x != nullptr;
  const double min_i = d.min_intercept, max_i = d.max_intercept;
  const double tol = d.tol;
  const long long n_hyp = d.n_hyp;
  const double lsq_xc = 0.5 * (d.pix_min + d.pix_max);
  const double lsq_xsc = (d.pix_max > d.pix_min) ? 2.0 / (d.pix_max - d.pix_min) : 1.0;

  // ---- stage B + C: drain up to 32 queued hypotheses ----
  auto drain = [&]() {
    const int take = min(q_count, 32);
    const bool have = lane < take;
    const int slot = (q_head + lane) % Q_CAP;
    const double* qc = q + (size_t)slot * qstride;
    bool mono = false;
    long long hid = -1;
    if (have) {
      hid = __double_as_longlong(qc[ncoef]);
      mono = monotonic_test(qc, deg, d.pix_min, d.pix_max);
      if (!mono && d.d_out_status) d.d_out_status[hid] = (uint8_t)RB2_HYP_MONOTONIC;
    }
    unsigned pass = __ballot_sync(0xffffffffu, have && mono);
    c_mono += __popc(__ballot_sync(0xffffffffu, have && !mono));
    while (pass) {
      const int src = __ffs(pass) - 1;
      pass &= pass - 1;
      const int sl = (q_head + src) % Q_CAP;
      const double* cc = q + (size_t)sl * qstride;
      const long long id = __shfl_sync(0xffffffffu, hid, src);
      // -- bijective match (calibrator.py:341-380) --
      for (int i = lane; i < n_wids; i += 32) berr[i] = 0x7ff0000000000000ull;  // +inf = empty
      __syncwarp();
      for (int u = lane; u < n_u; u += 32) {
        const double fit = horner_unfused(cc, ncoef, s_peaks[u]);
        const int o0 = s_coff[u], o1 = s_coff[u + 1];
        double be = CUDART_INF;
        int bw = -1;
        for (int k = o0; k < o1; ++k) {
          const double e = fabs(sub(fit, s_cwave[k]));
          if (e < be) { be = e; bw = s_cwid[k]; }  // first minimum wins
        }
        if (bw >= 0) atomicMin(&berr[bw], (unsigned long long)__double_as_longlong(be));
      }
      __syncwarp();
      double esum = 0.0;
      int nm = 0, ni = 0;
      for (int i = lane; i < n_wids; i += 32) {
        const unsigned long long bits = berr[i];
        if (bits != 0x7ff0000000000000ull) {
          double e = __longlong_as_double((long long)bits);
          if (e > tol) e = tol;  // M-SAC clip (:611)
          esum += e;
          ++nm;
          ni += (e < tol);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        esum += __shfl_xor_sync(0xffffffffu, esum, o);
        nm += __shfl_xor_sync(0xffffffffu, nm, o);
        ni += __shfl_xor_sync(0xffffffffu, ni, o);
      }
      __syncwarp();
      if (nm == 0) {  // :607-608
        ++c_empty;
        if (lane == 0 && d.d_out_status) d.d_out_status[id] = (uint8_t)RB2_HYP_EMPTY;
        continue;
      }
      double weight;
      const bool ok = hough_weight_warp(d, cc, deg, &weight);
      if (!ok) {
        ++c_unsup;
        if (lane == 0 && d.d_out_status) d.d_out_status[id] = (uint8_t)RB2_HYP_UNSUPPORTED;
        continue;
      }
      const double cost = esum / (double)(nm - ncoef + 1) / (weight + 1e-9);  // :629-633
      ++c_scored;
      if (lane == 0) {
        if (d.d_out_status) d.d_out_status[id] = (uint8_t)RB2_HYP_SCORED;
        if (d.d_out_cost) d.d_out_cost[id] = cost;
        if (d.d_out_weight) d.d_out_weight[id] = weight;
        if (d.d_out_counts) { d.d_out_counts[2 * id] = nm; d.d_out_counts[2 * id + 1] = ni; }
      }
      const long long gid = d.hyp_begin + id;
      bool elig = (ni >= d.min_inliers) && ((double)ni >= d.min_inliers_frac) && (cost <= d.cost_ceiling);
      if (elig && d.floor_id >= 0) elig = (cost > d.floor_cost) || (cost == d.floor_cost && gid < d.floor_id);
      if (elig) {
        ++c_elig;
        if (best.worse_than(cost, gid)) {
          best.cost = cost; best.id = gid; best.n_match = nm; best.n_inl = ni;
#pragma unroll
          for (int i = 0; i < NCOEF; ++i) best_coef[i] = (i < ncoef) ? cc[i] : 0.0;
        }
      }
    }
    q_head = (q_head + take) % Q_CAP;
    q_count -= take;
  };

  // ---- stage A over this warp's hypotheses ----
  const long long stride = (long long)gridDim.x * RS_THREADS;
  for (long long base = (long long)blockIdx.x * RS_THREADS + warp * 32; base < n_hyp; base += stride) {
    const long long id = base + lane;  // index inside this call
    const bool active = id < n_hyp;
    bool pass = false, aborted = false;
    double cf[NCOEF];
#pragma unroll
    for (int i = 0; i < NCOEF; ++i) cf[i] = 0.0;
    if (active) {
      int pi[RB2_MAX_SAMPLE];
      int wi[RB2_MAX_SAMPLE];
      int ci[RB2_MAX_SAMPLE];
      if (replay) {
        for (int k = 0; k < s; ++k) {
          pi[k] = d.d_replay_x[id * s + k];
          ci[k] = tb.coff[pi[k]] + d.d_replay_y[id * s + k];
        }
      } else {
        Philox rng;
        rng.init(d.seed, (unsigned long long)(d.hyp_begin + id));
        // peaks: weighted, without replacement (np.random.choice(..., replace=False, p=prob), :538-540)
        for (int k = 0; k < s; ++k) {
          int idx;
          bool dup;
          do {
            idx = draw_peak(tb, rng.next());
            dup = false;
            for (int j = 0; j < k; ++j) dup |= (pi[j] == idx);
          } while (dup);
          pi[k] = idx;
        }
        // one candidate wavelength per peak, no wavelength twice (:544-566)
        for (int k = 0; k < s && !aborted; ++k) {
          const int o0 = tb.coff[pi[k]], K = tb.coff[pi[k] + 1] - o0;
          bool checked = false;
          for (;;) {
            const int j = (int)__umulhi(rng.next(), (unsigned)K);
            const int w = tb.cwid[o0 + j];
            bool dup = false;
            for (int t = 0; t < k; ++t) dup |= (wi[t] == w);
            if (!dup) { wi[k] = w; ci[k] = o0 + j; break; }
            if (!checked) {  // is any unused wavelength left for this peak?
              bool possible = false;
              for (int jj = 0; jj < K; ++jj) {
                const int ww = tb.cwid[o0 + jj];
                bool used = false;
                for (int t = 0; t < k; ++t) used |= (wi[t] == ww);
                possible |= !used;
              }
              if (!possible) { aborted = true; break; }
              checked = true;
            }
          }
        }
      }
      if (!aborted) {
        if constexpr (DEG_T > 0) {
          double xs[DEG_T + 1], ys[DEG_T + 1], cc[DEG_T + 1];
#pragma unroll
          for (int k = 0; k <= DEG_T; ++k) { xs[k] = tb.peaks[pi[k]]; ys[k] = tb.cwave[ci[k]]; }
          newton_fit<DEG_T>(xs, ys, cc);
#pragma unroll
          for (int k = 0; k <= DEG_T; ++k) cf[k] = cc[k];
        } else {
          double xs[RB2_MAX_SAMPLE], ys[RB2_MAX_SAMPLE];
          int m = 0;
          for (int k = 0; k < s; ++k) { xs[m] = tb.peaks[pi[k]]; ys[m] = tb.cwave[ci[k]]; ++m; }
          for (int k = 0; k < d.n_known && m < RB2_MAX_SAMPLE; ++k) { xs[m] = d.d_known_x[k]; ys[m] = d.d_known_y[k]; ++m; }
          lsq_polyfit(xs, ys, m, deg, lsq_xc, lsq_xsc, cf);
        }
        pass = !((cf[0] < min_i) || (cf[0] > max_i));  // :578-582
        if (d.d_out_coeffs)
          for (int i = 0; i < ncoef; ++i) d.d_out_coeffs[id * ncoef + i] = cf[i];
        if (!pass && d.d_out_status) d.d_out_status[id] = (uint8_t)RB2_HYP_INTERCEPT;
      } else if (d.d_out_status) {
        d.d_out_status[id] = (uint8_t)RB2_HYP_ABORT;
      }
    }
    const unsigned pm = __ballot_sync(0xffffffffu, pass);
    c_drawn += __popc(__ballot_sync(0xffffffffu, active && !aborted));
    c_abort += __popc(__ballot_sync(0xffffffffu, aborted));
    c_icpt += __popc(__ballot_sync(0xffffffffu, active && !aborted && !pass));
    const int npass = __popc(pm);
    if (q_count + npass > Q_CAP) drain();
    if (pass) {
      const int slot = (q_head + q_count + __popc(pm & ((1u << lane) - 1u))) % Q_CAP;
      double* qc = q + (size_t)slot * qstride;
      for (int i = 0; i < ncoef; ++i) qc[i] = cf[i];
      qc[ncoef] = __longlong_as_double(id);
    }
    q_count += npass;
    __syncwarp();
    if (q_count >= 32) drain();
  }
  while (q_count > 0) drain();

  // ---- per-warp -> per-CTA best, counters ----
  if (lane == 0) {
    s_best[warp] = best;
#pragma unroll
    for (int i = 0; i < NCOEF; ++i) s_bcoef[warp][i] = best_coef[i];
    atomicAdd(&s_cnt[0], (unsigned long long)c_drawn);
    atomicAdd(&s_cnt[1], (unsigned long long)c_abort);
    atomicAdd(&s_cnt[2], (unsigned long long)c_icpt);
    atomicAdd(&s_cnt[3], (unsigned long long)c_mono);
    atomicAdd(&s_cnt[4], (unsigned long long)c_empty);
    atomicAdd(&s_cnt[5], (unsigned long long)c_scored);
    atomicAdd(&s_cnt[6], (unsigned long long)c_elig);
    atomicAdd(&s_cnt[7], (unsigned long long)c_unsup);
  }
  __syncthreads();
  if (tid == 0) {
    int bw = 0;
    for (int w = 1; w < RS_WARPS; ++w)
      if (s_best[bw].worse_than(s_best[w].cost, s_best[w].id) && s_best[w].id >= 0) bw = w;
    rb2_ransac_result r;
    r.cost = s_best[bw].cost;
    r.hyp_id = s_best[bw].id;
    r.n_match = s_best[bw].n_match;
    r.n_inliers = s_best[bw].n_inl;
    for (int i = 0; i < NCOEF; ++i) r.coeffs[i] = s_bcoef[bw][i];
    for (int i = 0; i < 8; ++i) r.counters[i] = s_cnt[i];
    block_out[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = r;
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

static size_t ransac_smem_bytes(const rb2_ransac_desc& h, int n_c) {
  const int ncoef = h.deg + 1;
  size_t b = 0;
  b += (size_t)h.n_upeaks * 8 + (size_t)n_c * 8;
  b += (size_t)RS_WARPS * Q_CAP * (ncoef + 1) * 8;
  b += (size_t)RS_WARPS * h.n_wids * 8;
  b += (size_t)(h.n_upeaks + 1) * 4 + (size_t)n_c * 4 + (size_t)h.n_upeaks * 4;
  return b;
}

}  // namespace rb2

extern "C" size_t rb2_ransac_scratch_bytes(int n_problems, int blocks_per_problem) {
  int sms = 148;
  rb2_sm_count(&sms);
  if (blocks_per_problem <= 0) blocks_per_problem = 4 * sms;
  return (size_t)n_problems * blocks_per_problem * sizeof(rb2_ransac_result);
}

template <int DEG_T>
static int launch_ransac(dim3 grid, size_t smem, cudaStream_t stream, const rb2_ransac_desc* d_desc,
                         rb2_ransac_result* scratch) {
  using namespace rb2;
  RB2_CHECK(cudaFuncSetAttribute(ransac_kernel<DEG_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ransac_kernel<DEG_T><<<grid, RS_THREADS, smem, stream>>>(d_desc, scratch);
  return RB2_OK;
}

extern "C" int rb2_ransac_batch(int n_problems, const rb2_ransac_desc* d_desc, const rb2_ransac_desc* h_desc,
                                rb2_ransac_result* d_result, void* d_block_scratch, int blocks_per_problem,
                                void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_desc || !h_desc || !d_result || !d_block_scratch) return RB2_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  int sms = 148;
  if (rb2_sm_count(&sms) != RB2_OK) return RB2_ERR_NO_DEVICE;
  if (blocks_per_problem <= 0) blocks_per_problem = 4 * sms;
  // all problems of one call share degree / sample size (one kernel instantiation)
  const int deg = h_desc[0].deg, s = h_desc[0].sample_size;
  bool fast = true;
  size_t smem = 0;
  for (int i = 0; i < n_problems; ++i) {
    const rb2_ransac_desc& h = h_desc[i];
    if (h.deg != deg || h.sample_size != s) return RB2_ERR_ARG;
    if (h.deg < 1 || h.deg > RB2_MAX_DEG || h.sample_size < h.deg + 1) return RB2_ERR_ARG;
    if (h.sample_size + h.n_known > RB2_MAX_SAMPLE) return RB2_ERR_ARG;
    if (h.n_upeaks < h.sample_size && !h.d_replay_x) return RB2_ERR_ARG;
    if (h.sample_size != h.deg + 1 || h.n_known != 0) fast = false;
    if (h.n_cand < h.n_upeaks || h.n_wids < 1) return RB2_ERR_ARG;
    smem = max(smem, ransac_smem_bytes(h, h.n_cand));
  }
  if (smem > 200 * 1024) return RB2_ERR_ARG;
  dim3 grid(blocks_per_problem, n_problems);
  rb2_ransac_result* scratch = (rb2_ransac_result*)d_block_scratch;
  int rc;
  if (fast && deg == 1) rc = launch_ransac<1>(grid, smem, stream, d_desc, scratch);
  else if (fast && deg == 2) rc = launch_ransac<2>(grid, smem, stream, d_desc, scratch);
  else if (fast && deg == 3) rc = launch_ransac<3>(grid, smem, stream, d_desc, scratch);
  else if (fast && deg == 4) rc = launch_ransac<4>(grid, smem, stream, d_desc, scratch);
  else if (fast && deg == 5) rc = launch_ransac<5>(grid, smem, stream, d_desc, scratch);
  else rc = launch_ransac<0>(grid, smem, stream, d_desc, scratch);
  if (rc != RB2_OK) return rc;
  ransac_finalize_kernel<<<n_problems, 32, 0, stream>>>(scratch, blocks_per_problem, blocks_per_problem, 1, d_result);
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
