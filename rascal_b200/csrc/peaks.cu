// K6: peak finding and sub-pixel peak refinement for batches of arc spectra - the step
// before the calibration path.
//
//   find_peaks_kernel    scipy.signal.find_peaks(height, distance, prominence): one CTA per
//                        spectrum; candidates, the distance sort and the keep flags live in
//                        shared memory.
//   refine_windows_kernel / refine_fit_kernel / refine_collect_kernel
//                        rascal util.refine_peaks (util.py:450-523): the in-place window
//                        normalisation is sequential per spectrum (one warp walks the peaks),
//                        the Gaussian fits are independent (MINPACK's lmdif, n = 3, as
//                        scipy.optimize.curve_fit runs it: persistent one-warp CTAs, one peak per
//                        lane, idle lanes refilled from a queue over the whole batch, the lanes of
//                        a warp kept in step at every outer iteration), the final range /
//                        np.isclose filter is sequential again.
//
// The Levenberg-Marquardt code below is written from the published MINPACK-1 algorithm (More,
// Garbow, Hillstrom 1980: lmdif, lmpar, qrfac, qrsolv, enorm, fdjac2) specialised to three
// parameters; oracle/peaks.py holds the same algorithm in NumPy, pinned against SciPy.
#include "common.cuh"
#include <float.h>
#include <stdlib.h>

namespace rb2 {

// ------------------------------------------------------------------------------------------
// find_peaks
// ------------------------------------------------------------------------------------------
constexpr int FP_THREADS = 256;

// exclusive prefix sum of one int per thread over the CTA; returns the total through `total`
__device__ int block_exclusive_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // s_warp may still be read from a previous call
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  int base = 0, tot = 0;
  for (int i = 0; i < FP_THREADS / 32; ++i) {
    const int c = s_warp[i];
    if (i < w) base += c;
    tot += c;
  }
  *total = tot;
  return base + inc - v;
}

// ordered in-place compaction of (idx, val) by keep[]; tmp_* are scratch of the same size
__device__ int block_compact(int n, int* idx, double* val, const unsigned char* keep, int* tmp_i, double* tmp_v,
                             int* s_warp) {
  const int chunk = (n + FP_THREADS - 1) / FP_THREADS;
  const int b = min(n, (int)threadIdx.x * chunk), e = min(n, b + chunk);
  int c = 0;
  for (int i = b; i < e; ++i) c += keep[i] ? 1 : 0;
  int total;
  int o = block_exclusive_scan(c, s_warp, &total);
  for (int i = b; i < e; ++i)
    if (keep[i]) {
      tmp_i[o] = idx[i];
      tmp_v[o] = val[i];
      ++o;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += FP_THREADS) {
    idx[i] = tmp_i[i];
    val[i] = tmp_v[i];
  }
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(FP_THREADS) find_peaks_kernel(const rb2_findpeaks_desc* __restrict__ descs, int cap) {
  extern __shared__ double fp_smem[];
  double* val = fp_smem;              // [cap] height, later prominence, of candidate k
  double* key = val + cap;            // [cap] sort key / compaction scratch
  int* idx = (int*)(key + cap);       // [cap] sample index of candidate k (ascending)
  int* ord = idx + cap;               // [cap] candidates in ascending priority / compaction scratch
  unsigned char* keep = (unsigned char*)(ord + cap);
  __shared__ int s_warp[FP_THREADS / 32];

  const rb2_findpeaks_desc d = descs[blockIdx.x];
  const double* __restrict__ x = d.d_spectrum;
  const int n = d.n_pix, tid = threadIdx.x;
  const bool has_h = !isnan(d.height), has_d = !isnan(d.distance), has_p = !isnan(d.prominence);

  // (1) local maxima: a rising edge at i followed by a flat run that ends in a fall; the
  //     middle of the run is the peak.  Every maximum has exactly one rising edge, so the
  //     samples can be examined independently; contiguous chunks keep the output ordered.
  const int chunk = (n + FP_THREADS - 1) / FP_THREADS;
  const int b = max(1, tid * chunk), e = min(n - 1, (tid + 1) * chunk);
  int nc = 0;
  int o = 0;
  for (int pass = 0; pass < 2; ++pass) {
    int c = 0;
    for (int i = b; i < e; ++i) {
      const double v = x[i];
      if (!(x[i - 1] < v)) continue;
      int a = i + 1;
      while (a < n - 1 && x[a] == v) ++a;
      if (!(x[a] < v)) continue;
      if (has_h && !(v >= d.height)) continue;
      if (pass) {
        idx[o + c] = (i + a - 1) >> 1;
        val[o + c] = v;
      }
      ++c;
    }
    if (pass == 0) o = block_exclusive_scan(c, s_warp, &nc);
  }
  __syncthreads();

  // (2) distance: visit the candidates from the highest down (equal heights: later sample
  //     first) and strike every unvisited neighbour closer than ceil(distance) samples.
  if (has_d && nc > 1) {
    int P = 1;
    while (P < nc) P <<= 1;
    for (int i = tid; i < P; i += FP_THREADS) {
      key[i] = i < nc ? val[i] : CUDART_INF;
      ord[i] = i < nc ? i : 0x7fffffff;
      if (i < nc) keep[i] = 1;
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < P; i += FP_THREADS) {
          const int l = i ^ j;
          if (l > i) {
            const double ki = key[i], kl = key[l];
            const int oi = ord[i], ol = ord[l];
            const bool gt = ki > kl || (ki == kl && oi > ol);
            if (((i & k) == 0) == gt) {
              key[i] = kl; key[l] = ki;
              ord[i] = ol; ord[l] = oi;
            }
          }
        }
        __syncthreads();
      }
    if (tid == 0) {
      const double dist = ceil(d.distance);
      for (int i = nc - 1; i >= 0; --i) {
        const int j = ord[i];
        if (!keep[j]) continue;
        const int pj = idx[j];
        for (int k = j - 1; k >= 0 && (double)(pj - idx[k]) < dist; --k) keep[k] = 0;
        for (int k = j + 1; k < nc && (double)(idx[k] - pj) < dist; ++k) keep[k] = 0;
      }
    }
    __syncthreads();
    nc = block_compact(nc, idx, val, keep, ord, key, s_warp);
  }

  // (3) prominence (wlen = None): walk outwards to the next strictly higher sample on each
  //     side, remembering the lowest sample passed.
  if (has_p) {
    // one warp per peak, 32 samples per step: the first lane that sees a higher sample ends the walk
    const int lane = tid & 31;
    for (int k = tid >> 5; k < nc; k += FP_THREADS / 32) {
      const int p = idx[k];
      const double xp = x[p];
      double side_min[2];
      for (int dir = 0; dir < 2; ++dir) {
        double mn = xp;
        for (int base = p;; base += dir ? 32 : -32) {
          const int i = dir ? base + lane : base - lane;
          const bool valid = i >= 0 && i < n;
          const double v = valid ? x[i] : xp;
          const unsigned stop = __ballot_sync(0xffffffffu, !valid || !(v <= xp));
          const int first = stop ? __ffs(stop) - 1 : 32;
          double c = lane < first ? v : xp;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) c = fmin(c, __shfl_xor_sync(0xffffffffu, c, o));
          mn = fmin(mn, c);
          if (stop) break;
        }
        side_min[dir] = mn;
      }
      if (lane == 0) {
        const double prom = xp - fmax(side_min[0], side_min[1]);
        val[k] = prom;  // carried through the compaction (heights are re-read from the spectrum)
        keep[k] = prom >= d.prominence;
      }
    }
    __syncthreads();
    nc = block_compact(nc, idx, val, keep, ord, key, s_warp);
  }

  for (int k = tid; k < min(nc, d.max_peaks); k += FP_THREADS) {
    const int p = idx[k];
    d.d_peaks[k] = p;
    if (d.d_peaks_f64) d.d_peaks_f64[k] = (double)p;
    if (d.d_heights) d.d_heights[k] = x[p];
    if (d.d_prominences) d.d_prominences[k] = has_p ? val[k] : CUDART_NAN;
  }
  if (tid == 0) d.d_n_peaks[0] = nc;
}

// ------------------------------------------------------------------------------------------
// MINPACK lmdif for the three-parameter Gaussian of util.gauss
// ------------------------------------------------------------------------------------------
constexpr int LM_N = 3;
constexpr int LM_WMAX = 2 * RB2_REFINE_MAX_WINDOW;
constexpr double LM_EPSMCH = DBL_EPSILON;
constexpr double LM_DWARF = DBL_MIN;
constexpr double LM_TOL = 1.49012e-8;  // scipy.optimize.leastsq: ftol = xtol
constexpr int LM_MAXFEV = 200 * (LM_N + 1);
#ifndef RB2_LM_UNROLL
#define RB2_LM_UNROLL 1
#endif
constexpr int LM_UNROLL = RB2_LM_UNROLL;  // unrolling of the loops over the independent points of a window

// Euclidean norm with MINPACK's three accumulators (small / intermediate / large components)
__device__ __noinline__ double enorm(int n, const double* v) {
  const double rdwarf = 3.834e-20, rgiant = 1.304e19;
  const double agiant = rgiant / (double)n;
  {  // every component zero or of intermediate size (the normal case): MINPACK's result is sqrt(sum x^2)
    double q = 0.0;
    bool mid = true;
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
      const double xabs = fabs(v[i]);
      mid &= (xabs > rdwarf && xabs < agiant) || xabs == 0.0;
      q += xabs * xabs;
    }
    if (mid) return sqrt(q);
  }
  double s1 = 0, s2 = 0, s3 = 0, x1max = 0, x3max = 0;
  for (int i = 0; i < n; ++i) {
    const double xabs = fabs(v[i]);
    if (xabs > rdwarf && xabs < agiant) {
      s2 += xabs * xabs;
    } else if (xabs <= rdwarf) {
      if (xabs > x3max) {
        const double r = x3max / xabs;
        s3 = 1.0 + s3 * r * r;
        x3max = xabs;
      } else if (xabs != 0.0) {
        const double r = xabs / x3max;
        s3 += r * r;
      }
    } else {
      if (xabs > x1max) {
        const double r = x1max / xabs;
        s1 = 1.0 + s1 * r * r;
        x1max = xabs;
      } else {
        const double r = xabs / x1max;
        s1 += r * r;
      }
    }
  }
  if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
  if (s2 != 0.0) {
    if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
    return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
  }
  return x3max * sqrt(s3);
}

// residuals of util.gauss (util.py:447) on x = 0..m-1, each NumPy ufunc rounded on its own
__device__ __noinline__ void gauss_residuals(const double* p, const double* __restrict__ y, int m, double* __restrict__ f) {
  const double den = add(mul(2.0, mul(p[2], p[2])), 1e-9);
#pragma unroll LM_UNROLL
  for (int i = 0; i < m; ++i) {  // independent points: unrolled so that one lane overlaps their div / exp chains
    const double dx = sub((double)i, p[1]);
    f[i] = sub(mul(p[0], exp(-mul(dx, dx) / den)), y[i]);
  }
}

// Solve R z = Q^T b augmented by the diagonal `diag` (Givens eliminations); r: upper triangle
// in, strict lower triangle used as work space, diagonal restored.
__device__ __noinline__ void qrsolv3(double r[LM_N][LM_N], const int* ipvt, const double* diag, const double* qtb, double* x,
                        double* sdiag) {
  double wa[LM_N];
  for (int j = 0; j < LM_N; ++j) {
    for (int i = j; i < LM_N; ++i) r[i][j] = r[j][i];
    x[j] = r[j][j];
    wa[j] = qtb[j];
  }
  for (int j = 0; j < LM_N; ++j) {
    const int l = ipvt[j];
    if (diag[l] != 0.0) {
      for (int k = j; k < LM_N; ++k) sdiag[k] = 0.0;
      sdiag[j] = diag[l];
      double qtbpj = 0.0;
      for (int k = j; k < LM_N; ++k) {
        if (sdiag[k] == 0.0) continue;
        double sn, cs;
        if (fabs(r[k][k]) < fabs(sdiag[k])) {
          const double cotan = r[k][k] / sdiag[k];
          sn = 0.5 / sqrt(0.25 + 0.25 * cotan * cotan);
          cs = sn * cotan;
        } else {
          const double tn = sdiag[k] / r[k][k];
          cs = 0.5 / sqrt(0.25 + 0.25 * tn * tn);
          sn = cs * tn;
        }
        r[k][k] = cs * r[k][k] + sn * sdiag[k];
        const double t = cs * wa[k] + sn * qtbpj;
        qtbpj = -sn * wa[k] + cs * qtbpj;
        wa[k] = t;
        for (int i = k + 1; i < LM_N; ++i) {
          const double u = cs * r[i][k] + sn * sdiag[i];
          sdiag[i] = -sn * r[i][k] + cs * sdiag[i];
          r[i][k] = u;
        }
      }
    }
    sdiag[j] = r[j][j];
    r[j][j] = x[j];
  }
  int nsing = LM_N;
  for (int j = 0; j < LM_N; ++j) {
    if (sdiag[j] == 0.0 && nsing == LM_N) nsing = j;
    if (nsing < LM_N) wa[j] = 0.0;
  }
  for (int j = nsing - 1; j >= 0; --j) {
    double s = 0.0;
    for (int i = j + 1; i < nsing; ++i) s += r[i][j] * wa[i];
    wa[j] = (wa[j] - s) / sdiag[j];
  }
  for (int j = 0; j < LM_N; ++j) x[ipvt[j]] = wa[j];
}

// Levenberg-Marquardt parameter for the trust region ||D x|| <= delta
__device__ __noinline__ double lmpar3(double r[LM_N][LM_N], const int* ipvt, const double* diag, const double* qtb, double delta,
                         double par, double* x, double* sdiag) {
  double wa1[LM_N], wa2[LM_N];
  int nsing = LM_N;
  for (int j = 0; j < LM_N; ++j) {
    wa1[j] = qtb[j];
    if (r[j][j] == 0.0 && nsing == LM_N) nsing = j;
    if (nsing < LM_N) wa1[j] = 0.0;
  }
  for (int j = nsing - 1; j >= 0; --j) {
    wa1[j] /= r[j][j];
    const double t = wa1[j];
    for (int i = 0; i < j; ++i) wa1[i] -= r[i][j] * t;
  }
  for (int j = 0; j < LM_N; ++j) x[ipvt[j]] = wa1[j];
  for (int j = 0; j < LM_N; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = enorm(LM_N, wa2);
  double fp = dxnorm - delta;
  if (fp <= 0.1 * delta) return 0.0;
  double parl = 0.0;
  if (nsing >= LM_N) {
    for (int j = 0; j < LM_N; ++j) {
      const int l = ipvt[j];
      wa1[j] = diag[l] * (wa2[l] / dxnorm);
    }
    for (int j = 0; j < LM_N; ++j) {
      double s = 0.0;
      for (int i = 0; i < j; ++i) s += r[i][j] * wa1[i];
      wa1[j] = (wa1[j] - s) / r[j][j];
    }
    const double t = enorm(LM_N, wa1);
    parl = ((fp / delta) / t) / t;
  }
  for (int j = 0; j < LM_N; ++j) {
    double s = 0.0;
    for (int i = 0; i <= j; ++i) s += r[i][j] * qtb[i];
    wa1[j] = s / diag[ipvt[j]];
  }
  const double gnorm = enorm(LM_N, wa1);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = LM_DWARF / fmin(delta, 0.1);
  par = fmax(par, parl);
  par = fmin(par, paru);
  if (par == 0.0) par = gnorm / dxnorm;
  for (int it = 1;; ++it) {
    if (par == 0.0) par = fmax(LM_DWARF, 0.001 * paru);
    double t = sqrt(par);
    for (int j = 0; j < LM_N; ++j) wa1[j] = t * diag[j];
    qrsolv3(r, ipvt, wa1, qtb, x, sdiag);
    for (int j = 0; j < LM_N; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = enorm(LM_N, wa2);
    t = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= t && t < 0.0) || it == 10) break;
    for (int j = 0; j < LM_N; ++j) {
      const int l = ipvt[j];
      wa1[j] = diag[l] * (wa2[l] / dxnorm);
    }
    for (int j = 0; j < LM_N; ++j) {
      wa1[j] /= sdiag[j];
      const double u = wa1[j];
      for (int i = j + 1; i < LM_N; ++i) wa1[i] -= r[i][j] * u;
    }
    t = enorm(LM_N, wa1);
    const double parc = ((fp / delta) / t) / t;
    if (fp > 0.0) parl = fmax(parl, par);
    if (fp < 0.0) paru = fmin(paru, par);
    par = fmax(parl, par + parc);
  }
  return par;
}

// lmdif, mode 1, n = 3, cut at its outer loop so that the lanes of a warp can walk it in step:
// lm_start() evaluates the start, lm_iterate() runs ONE outer iteration (Jacobian, QR, the inner
// trust-region loop) and returns MINPACK's info (0 = not finished).
struct LMState {
  double y[LM_WMAX], fvec[LM_WMAX], wa4[LM_WMAX];
  double fj[LM_N][LM_WMAX];  // Jacobian by columns; after qrfac: R above the diagonal, Householder vectors below
  double x[LM_N], diag[LM_N];
  double fnorm, par, xnorm, delta;
  int m, nfev, iter;
};

__device__ __noinline__ void lm_start(LMState& S) {
  gauss_residuals(S.x, S.y, S.m, S.fvec);
  S.nfev = 1;
  S.iter = 1;
  S.fnorm = enorm(S.m, S.fvec);
  S.par = S.xnorm = S.delta = 0.0;
}

__device__ __noinline__ int lm_iterate(LMState& S) {
  double qtf[LM_N], wa1[LM_N], wa2[LM_N], wa3[LM_N], acn[LM_N], rdiag[LM_N], wq[LM_N], sdiag[LM_N];
  double r[LM_N][LM_N];
  int ipvt[LM_N];
  const int m = S.m;
  double* x = S.x;
  double* diag = S.diag;
  double* fvec = S.fvec;
  double* wa4 = S.wa4;
  double(*fj)[LM_WMAX] = S.fj;
  const double eps = sqrt(LM_EPSMCH);
  int info = 0;
  // forward-difference Jacobian
  for (int j = 0; j < LM_N; ++j) {
    const double t = x[j];
    double h = eps * fabs(t);
    if (h == 0.0) h = eps;
    x[j] = t + h;
    gauss_residuals(x, S.y, m, wa4);
    x[j] = t;
    const double rh = 1.0 / h;  // one division per column: the differences carry ~1e-8 of rounding noise anyway
#pragma unroll LM_UNROLL
    for (int i = 0; i < m; ++i) fj[j][i] = (wa4[i] - fvec[i]) * rh;
  }
  S.nfev += LM_N;
  // Householder QR with column pivoting
  for (int j = 0; j < LM_N; ++j) {
    acn[j] = enorm(m, fj[j]);
    rdiag[j] = acn[j];
    wq[j] = rdiag[j];
    ipvt[j] = j;
  }
  const int minmn = m < LM_N ? m : LM_N;
#pragma unroll 1
  for (int j = 0; j < minmn; ++j) {
    int kmax = j;
    for (int k = j; k < LM_N; ++k)
      if (rdiag[k] > rdiag[kmax]) kmax = k;
    if (kmax != j) {
#pragma unroll 1
      for (int i = 0; i < m; ++i) {
        const double t = fj[j][i];
        fj[j][i] = fj[kmax][i];
        fj[kmax][i] = t;
      }
      rdiag[kmax] = rdiag[j];
      wq[kmax] = wq[j];
      const int t = ipvt[j];
      ipvt[j] = ipvt[kmax];
      ipvt[kmax] = t;
    }
    double ajnorm = enorm(m - j, fj[j] + j);
    if (ajnorm != 0.0) {
      if (fj[j][j] < 0.0) ajnorm = -ajnorm;
      const double rn = 1.0 / ajnorm;
#pragma unroll LM_UNROLL
      for (int i = j; i < m; ++i) fj[j][i] *= rn;
      fj[j][j] += 1.0;
#pragma unroll 1
      for (int k = j + 1; k < LM_N; ++k) {
        double s = 0.0;
#pragma unroll 1
        for (int i = j; i < m; ++i) s += fj[j][i] * fj[k][i];
        const double t = s / fj[j][j];
#pragma unroll LM_UNROLL
        for (int i = j; i < m; ++i) fj[k][i] -= t * fj[j][i];
        if (rdiag[k] != 0.0) {
          const double u = fj[k][j] / rdiag[k];
          rdiag[k] *= sqrt(fmax(0.0, 1.0 - u * u));
          const double q = rdiag[k] / wq[k];
          if (0.05 * q * q <= LM_EPSMCH) {
            rdiag[k] = enorm(m - j - 1, fj[k] + j + 1);
            wq[k] = rdiag[k];
          }
        }
      }
    }
    rdiag[j] = -ajnorm;
  }
  if (S.iter == 1) {
    for (int j = 0; j < LM_N; ++j) {
      diag[j] = acn[j] == 0.0 ? 1.0 : acn[j];
      wa3[j] = diag[j] * x[j];
    }
    S.xnorm = enorm(LM_N, wa3);
    S.delta = 100.0 * S.xnorm;
    if (S.delta == 0.0) S.delta = 100.0;
  }
  // first n components of Q^T fvec
#pragma unroll 1
  for (int i = 0; i < m; ++i) wa4[i] = fvec[i];
#pragma unroll 1
  for (int j = 0; j < LM_N; ++j) {
    if (j < m && fj[j][j] != 0.0) {
      double s = 0.0;
#pragma unroll 1
      for (int i = j; i < m; ++i) s += fj[j][i] * wa4[i];
      const double t = -s / fj[j][j];
#pragma unroll LM_UNROLL
      for (int i = j; i < m; ++i) wa4[i] += fj[j][i] * t;
    }
    if (j < m) fj[j][j] = rdiag[j];
    qtf[j] = j < m ? wa4[j] : 0.0;
  }
  for (int j = 0; j < LM_N; ++j)
    for (int i = 0; i < LM_N; ++i) r[i][j] = (i <= j && i < m) ? fj[j][i] : 0.0;
  // norm of the scaled gradient
  double gnorm = 0.0;
  if (S.fnorm != 0.0) {
    for (int j = 0; j < LM_N; ++j) {
      const int l = ipvt[j];
      if (acn[l] != 0.0) {
        double s = 0.0;
        for (int i = 0; i <= j; ++i) s += r[i][j] * (qtf[i] / S.fnorm);
        const double g = fabs(s / acn[l]);
        if (g > gnorm) gnorm = g;  // max(gnorm, NaN) keeps gnorm, as the Fortran does
      }
    }
  }
  if (gnorm <= 0.0) return 4;  // gtol = 0
  for (int j = 0; j < LM_N; ++j) diag[j] = fmax(diag[j], acn[j]);
#pragma unroll 1
  for (;;) {
    S.par = lmpar3(r, ipvt, diag, qtf, S.delta, S.par, wa1, sdiag);
    for (int j = 0; j < LM_N; ++j) {
      wa1[j] = -wa1[j];
      wa2[j] = x[j] + wa1[j];
      wa3[j] = diag[j] * wa1[j];
    }
    const double pnorm = enorm(LM_N, wa3);
    if (S.iter == 1) S.delta = fmin(S.delta, pnorm);
    gauss_residuals(wa2, S.y, m, wa4);
    ++S.nfev;
    const double fnorm1 = enorm(m, wa4);
    double actred = -1.0;
    if (0.1 * fnorm1 < S.fnorm) {
      const double q = fnorm1 / S.fnorm;
      actred = 1.0 - q * q;
    }
    for (int j = 0; j < LM_N; ++j) wa3[j] = 0.0;
    for (int j = 0; j < LM_N; ++j) {
      const double t = wa1[ipvt[j]];
      for (int i = 0; i <= j; ++i) wa3[i] += r[i][j] * t;
    }
    const double temp1 = enorm(LM_N, wa3) / S.fnorm, temp2 = (sqrt(S.par) * pnorm) / S.fnorm;
    const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
    const double dirder = -(temp1 * temp1 + temp2 * temp2);
    const double ratio = prered != 0.0 ? actred / prered : 0.0;
    if (ratio <= 0.25) {
      double t = actred >= 0.0 ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
      if (0.1 * fnorm1 >= S.fnorm || t < 0.1) t = 0.1;
      S.delta = t * fmin(S.delta, pnorm / 0.1);
      S.par /= t;
    } else if (S.par == 0.0 || ratio >= 0.75) {
      S.delta = pnorm / 0.5;
      S.par *= 0.5;
    }
    if (ratio >= 1e-4) {
      for (int j = 0; j < LM_N; ++j) {
        x[j] = wa2[j];
        wa2[j] = diag[j] * x[j];
      }
#pragma unroll 1
      for (int i = 0; i < m; ++i) fvec[i] = wa4[i];
      S.xnorm = enorm(LM_N, wa2);
      S.fnorm = fnorm1;
      ++S.iter;
    }
    const bool fconv = fabs(actred) <= LM_TOL && prered <= LM_TOL && 0.5 * ratio <= 1.0;
    if (fconv) info = 1;
    if (S.delta <= LM_TOL * S.xnorm) info = 2;
    if (fconv && info == 2) info = 3;
    if (info != 0) break;
    if (S.nfev >= LM_MAXFEV) info = 5;
    if (fabs(actred) <= LM_EPSMCH && prered <= LM_EPSMCH && 0.5 * ratio <= 1.0) info = 6;
    if (S.delta <= LM_EPSMCH * S.xnorm) info = 7;
    if (gnorm <= LM_EPSMCH) info = 8;
    if (info != 0) break;
    if (ratio >= 1e-4) break;
  }
  return info;
}

// ------------------------------------------------------------------------------------------
// refine_peaks
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int refine_count(const rb2_refine_desc& d) {
  int np = d.n_peaks;
  if (d.d_n_peaks) np = max(0, min(np, d.d_n_peaks[0]));
  return np;
}

// util.py:471-485: working copy, then per peak (in order) the window divided by its nanmax in place.
__global__ void __launch_bounds__(128) refine_windows_kernel(const rb2_refine_desc* __restrict__ descs,
                                                             int* __restrict__ item_off) {
  const rb2_refine_desc d = descs[blockIdx.x];
  const int n = d.n_pix;
  if (threadIdx.x == 0) item_off[blockIdx.x] = refine_count(d);  // scanned by refine_queue_kernel
  for (int i = threadIdx.x; i < n; i += blockDim.x) d.d_work[i] = d.d_spectrum[i];
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x, np = refine_count(d), ww = d.window_width, W = 2 * ww;
  for (int k = 0; k < np; ++k) {
    const double peak = d.d_peaks[k];
    const int ip = (int)peak;  // int(): towards zero
    const int lo = max(0, ip - ww), hi = min(ip + ww, n), m = hi - lo;
    int status = RB2_PEAK_OK;
    if (m <= 0) {
      status = RB2_PEAK_EMPTY_WINDOW;
    } else {
      double mx = CUDART_NAN;
      for (int i = lane; i < m; i += 32) mx = fmax(mx, d.d_work[lo + i]);  // fmax drops NaN: np.nanmax
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      int nfin = 0;
      for (int i = lane; i < m; i += 32) {
        const double y = d.d_work[lo + i] / mx;
        d.d_work[lo + i] = y;
        d.d_window[(size_t)k * W + i] = y;
        nfin += isfinite(y) ? 1 : 0;
      }
      nfin = __reduce_add_sync(0xffffffffu, nfin);
      if (nfin == 0) status = RB2_PEAK_NO_DATA;
      else if (nfin < m) status = RB2_PEAK_NOT_FINITE;
      else if (m < LM_N) status = RB2_PEAK_TOO_FEW;
    }
    if (lane == 0) d.d_status[k] = status;
    __syncwarp();  // the next window may overlap this one
  }
}

// np.sum of a short float64 vector: NumPy's pairwise kernel below its block size (8 running
// sums over the multiples of 8, combined as a tree, the tail added in order)
template <typename F>
__device__ double numpy_sum(int n, F term) {
  if (n < 8) {
    double s = 0.0;  // -0.0 + ... is irrelevant here
    for (int i = 0; i < n; ++i) s = i == 0 ? term(0) : add(s, term(i));
    return s;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = term(j);
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = add(r[j], term(i + j));
  double s = add(add(add(r[0], r[1]), add(r[2], r[3])), add(add(r[4], r[5]), add(r[6], r[7])));
  for (; i < n; ++i) s = add(s, term(i));
  return s;
}

// util.py:487-502: start values, the fit, the per-peak rejections.  Every lane fits one peak at a time
// and takes the batch's next unfitted peak when it is done (most fits need ~10 outer iterations, one in
// a hundred needs 50-200: a fixed peak-to-lane map leaves 26 of 32 lanes idle); the lanes of a warp meet
// at every outer Levenberg-Marquardt iteration (the __any_sync), so the iteration's code runs once for
// all of them.
__device__ __noinline__ bool refine_fit_begin(const rb2_refine_desc& d, int k, LMState& S) {
  double* fit = d.d_fit + (size_t)k * 4;
  fit[0] = fit[1] = fit[2] = CUDART_NAN;
  fit[3] = 0.0;
  if (d.d_status[k] != RB2_PEAK_OK) return false;
  const int ww = d.window_width;
  const int ip = (int)d.d_peaks[k];
  const int lo = max(0, ip - ww), hi = min(ip + ww, d.n_pix), m = hi - lo;
  const double* __restrict__ w = d.d_window + (size_t)k * 2 * ww;
  double* y = S.y;
#pragma unroll 1
  for (int i = 0; i < m; ++i) y[i] = w[i];
  const double cnt = (double)m;
  const double mean = numpy_sum(m, [&](int i) { return mul((double)i, y[i]); }) / cnt;
  const double sigma = numpy_sum(m, [&](int i) {
                         const double dx = sub((double)i, mean);
                         return mul(y[i], mul(dx, dx));
                       }) / cnt;
  S.m = m;
  S.x[0] = 1.0;
  S.x[1] = mean;
  S.x[2] = sigma;
  lm_start(S);
  return true;
}

__device__ __noinline__ void refine_fit_end(const rb2_refine_desc& d, int k, const LMState& S, int ier) {
  double* fit = d.d_fit + (size_t)k * 4;
  fit[0] = S.x[0];
  fit[1] = S.x[1];
  fit[2] = S.x[2];
  fit[3] = (double)S.nfev;
  int status = RB2_PEAK_OK;
  if (ier < 1 || ier > 4) status = RB2_PEAK_FIT_FAILED;
  else if (S.x[0] < 0.0) status = RB2_PEAK_NEGATIVE;
  else if (S.x[1] > (double)d.n_pix || S.x[1] < 0.0) status = RB2_PEAK_OUTSIDE;
  d.d_status[k] = status | (ier << 8);
}

// Work queue of the fits: item_off[0..n] = exclusive prefix of the per-spectrum peak counts (in
// place over the counts refine_windows_kernel left), item_off[n + 1] = the queue head.
__global__ void __launch_bounds__(1024) refine_queue_kernel(int* __restrict__ item_off, int n) {
  __shared__ int s_part[1024];
  const int t = threadIdx.x, chunk = (n + 1023) / 1024;
  const int b = min(n, t * chunk), e = min(n, b + chunk);
  int sum = 0;
  for (int i = b; i < e; ++i) sum += item_off[i];
  s_part[t] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = t >= o ? s_part[t - o] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  int run = s_part[t] - sum;
  for (int i = b; i < e; ++i) {
    const int c = item_off[i];
    item_off[i] = run;
    run += c;
  }
  if (t == 1023) {
    item_off[n] = s_part[1023];
    item_off[n + 1] = 0;
  }
}

// One-warp CTAs, persistent: an idle lane takes the next (spectrum, peak) item of the whole batch.
__global__ void __launch_bounds__(32) refine_fit_kernel(const rb2_refine_desc* __restrict__ descs, int n_spectra,
                                                        int* __restrict__ item_off) {
  const int total = item_off[n_spectra];
  const unsigned lane = threadIdx.x, lt = (1u << lane) - 1u;
  LMState S;
  const rb2_refine_desc* dp = nullptr;
  int k = -1;
  bool run = false, drained = false;  // drained: warp-uniform
  for (;;) {
    const unsigned idle = __ballot_sync(0xffffffffu, !run);
    if (idle && !drained) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&item_off[n_spectra + 1], __popc(idle));
      base = __shfl_sync(0xffffffffu, base, 0);
      drained = base + __popc(idle) >= total;
      const int item = base + __popc(idle & lt);
      if (!run && item < total) {
        int lo = 0, hi = n_spectra - 1;  // last s with item_off[s] <= item
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (item_off[mid] <= item) lo = mid;
          else hi = mid - 1;
        }
        dp = descs + lo;
        k = item - item_off[lo];
        run = refine_fit_begin(*dp, k, S);
      }
    }
    if (!__any_sync(0xffffffffu, run)) {
      if (drained) break;
      continue;
    }
    if (run) {
      const int ier = lm_iterate(S);
      if (ier != 0) {
        refine_fit_end(*dp, k, S, ier);
        run = false;
      }
    }
  }
}

// util.py:499, 509-517: peak - window_width + centre, the (0, n_pix) mask and the np.isclose
// comparison with the previous entry of the list (rtol 1e-5, atol 1e-8).
__global__ void __launch_bounds__(64) refine_collect_kernel(const rb2_refine_desc* __restrict__ descs, int n_spectra) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_spectra) return;
  const rb2_refine_desc d = descs[s];
  const int np = refine_count(d);
  const double len = (double)d.n_pix, ww = (double)d.window_width;
  int out = 0;
  bool have_prev = false;
  double prev = 0.0;
  for (int k = 0; k < np; ++k) {
    const int st = d.d_status[k] & 0xff;
    double v;
    if (st == RB2_PEAK_OK) v = add(sub(d.d_peaks[k], ww), d.d_fit[(size_t)k * 4 + 1]);
    else if (st == RB2_PEAK_NOT_FINITE) v = CUDART_NAN;
    else continue;
    const bool close = have_prev && (prev == v || fabs(prev - v) <= 1e-8 + 1e-5 * fabs(v));
    if (v > 0.0 && v < len && !close) d.d_refined[out++] = v;
    prev = v;
    have_prev = true;
  }
  d.d_n_refined[0] = out;
}

}  // namespace rb2

extern "C" int rb2_find_peaks(int n_spectra, const rb2_findpeaks_desc* d_desc, const rb2_findpeaks_desc* h_desc,
                              void* stream_) {
  using namespace rb2;
  if (n_spectra <= 0 || !d_desc || !h_desc) return RB2_ERR_ARG;
  int max_pix = 0;
  for (int i = 0; i < n_spectra; ++i) {
    const rb2_findpeaks_desc& h = h_desc[i];
    if (h.n_pix < 0 || (h.n_pix && !h.d_spectrum) || h.max_peaks < 0 || !h.d_n_peaks || (h.max_peaks && !h.d_peaks))
      return RB2_ERR_ARG;
    if (!isnan(h.distance) && !(h.distance >= 1.0)) return RB2_ERR_ARG;  // SciPy: `distance` must be >= 1
    if (h.d_prominences && isnan(h.prominence)) return RB2_ERR_ARG;
    max_pix = max(max_pix, h.n_pix);
  }
  int cap = 32;
  while (cap < max_pix / 2 + 1) cap <<= 1;
  const size_t smem = (size_t)cap * (8 + 8 + 4 + 4 + 1);
  if (smem > 220 * 1024) return RB2_ERR_CAPACITY;  // n_pix <= RB2_FIND_MAX_PIX
  RB2_CHECK(cudaFuncSetAttribute(find_peaks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  find_peaks_kernel<<<n_spectra, FP_THREADS, smem, (cudaStream_t)stream_>>>(d_desc, cap);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

extern "C" size_t rb2_refine_scratch_bytes(int n_spectra) { return sizeof(int) * ((size_t)max(n_spectra, 0) + 2); }

extern "C" int rb2_refine_peaks(int n_spectra, const rb2_refine_desc* d_desc, const rb2_refine_desc* h_desc,
                                void* d_scratch, void* stream_) {
  using namespace rb2;
  if (n_spectra <= 0 || !d_desc || !h_desc || !d_scratch) return RB2_ERR_ARG;
  long long items = 0;
  for (int i = 0; i < n_spectra; ++i) {
    const rb2_refine_desc& h = h_desc[i];
    if (h.n_pix < 0 || h.n_peaks < 0 || h.window_width < 1 || h.window_width > RB2_REFINE_MAX_WINDOW) return RB2_ERR_ARG;
    if ((h.n_pix && (!h.d_spectrum || !h.d_work)) || !h.d_n_refined) return RB2_ERR_ARG;
    if (h.n_peaks && (!h.d_peaks || !h.d_window || !h.d_fit || !h.d_status || !h.d_refined)) return RB2_ERR_ARG;
    items += h.n_peaks;
  }
  if (items > 0x7fffffffLL) return RB2_ERR_CAPACITY;
  cudaStream_t st = (cudaStream_t)stream_;
  int* item_off = (int*)d_scratch;
  refine_windows_kernel<<<n_spectra, 128, 0, st>>>(d_desc, item_off);
  RB2_CHECK(cudaGetLastError());
  if (items > 0) {
    int sms = 0;
    const int rc = rb2_sm_count(&sms);
    if (rc != RB2_OK) return rc;
    refine_queue_kernel<<<1, 1024, 0, st>>>(item_off, n_spectra);
    RB2_CHECK(cudaGetLastError());
    // `items` is an upper bound of the queue length
    // Persistent warps per SM (RB2_FIT_WARPS_PER_SM overrides): the registers allow 16, but the kernel is
    // bound by its longest fits, not by throughput, and 4-8 warps keep the instruction cache and L1
    // (local-memory state) warmer: 22.5 ms against 24-28 ms for 4096 arcs, 11.6 against 12.9 ms for 512.
    static int per_sm = 0;
    if (!per_sm) {
      const char* e = getenv("RB2_FIT_WARPS_PER_SM");
      per_sm = e ? max(1, min(32, atoi(e))) : 8;
    }
    const int warps = (int)min((long long)sms * per_sm, (items + 31) / 32);
    refine_fit_kernel<<<warps, 32, 0, st>>>(d_desc, n_spectra, item_off);
    RB2_CHECK(cudaGetLastError());
  }
  refine_collect_kernel<<<(n_spectra + 63) / 64, 64, 0, st>>>(d_desc, n_spectra);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}
