// K4 — winner refit + acceptance (sm_100a).  One warp per problem.
//
// Replaces, for the winning hypothesis only: Calibrator._match_bijective
// (calibrator.py:317-380), the inlier selection (:640-643),
// models.robust_polyfit (models.py:70-123) and the clipped-residual RMS /
// minimum_fit_error test (:664-679).
//
// robust_polyfit minimises the Huber loss (f_scale = 1) of the residuals of a
// polynomial in x/std(x) against y/std(y) with SciPy's TRF.  The loss is convex;
// its minimiser is computed here directly: Householder QR least squares
// (exact when every normalised residual is inside Huber's quadratic zone,
// |r| <= 1 - always the case for inliers of a wavelength solution), followed by
// IRLS with Huber weights only if some |r| > 1.  See DESIGN.md for why the
// reference's own TRF iterate differs from this optimum by its stopping
// tolerance and finite-difference Jacobian noise.
#include "common.cuh"

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

__global__ void __launch_bounds__(32)
refit_kernel(const rb2_ransac_desc* __restrict__ descs, const rb2_ransac_result* __restrict__ winners,
             double min_fit_error, rb2_refit_result* __restrict__ out, double* __restrict__ mx_out,
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
  double* A = sw + n_u;                                                          // [n_u*(n+1)]
  int* bpeak = reinterpret_cast<int*>(A + (size_t)n_u * (n + 1));                // [n_w]

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
  warp_qr_solve(xn, yn, nullptr, m, n, A, a);
  // Huber (f_scale = 1): leave the quadratic zone only if some |r| > 1
  int iters = 0;
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
  r.huber_iters = iters;
  // un-normalise (models.py:117-123), low order first
  double xp = 1.0;
  for (int j = 0; j < n; ++j) { r.coeffs[j] = a[j] * ys / xp; xp *= xs; }
  // residual = polyval(matched_peaks, coeffs) - matched_atlas, clipped (:664-671)
  double ss = 0.0;
  for (int i = lane; i < m; i += 32) {
    double res = sub(horner_unfused(r.coeffs, n, ix[i]), iy[i]);
    if (fabs(res) > d.tol) res = d.tol;
    ss = fma(res, res, ss);
    mx_out[(size_t)p * max_up + i] = ix[i];
    my_out[(size_t)p * max_up + i] = iy[i];
    res_out[(size_t)p * max_up + i] = res;
  }
  r.rms = sqrt(warp_sum(ss) / m);
  r.accepted = (r.rms < min_fit_error) ? 0 : 1;
  if (lane == 0) out[p] = r;
}

}  // namespace rb2

extern "C" int rb2_refit_winners(int n_problems, const rb2_ransac_desc* d_desc, const rb2_ransac_desc* h_desc,
                                 const rb2_ransac_result* d_winner, double minimum_fit_error,
                                 rb2_refit_result* d_out, double* d_matched_x, double* d_matched_y,
                                 double* d_residual, int max_upeaks, void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_desc || !h_desc || !d_winner || !d_out) return RB2_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  size_t smem = 0;
  for (int i = 0; i < n_problems; ++i) {
    const rb2_ransac_desc& h = h_desc[i];
    if (h.n_upeaks > max_upeaks) return RB2_ERR_ARG;
    const size_t n = h.deg + 1;
    size_t b = (size_t)h.n_wids * 16 + (size_t)h.n_upeaks * 8 * 5 + (size_t)h.n_upeaks * (n + 1) * 8 + (size_t)h.n_wids * 4;
    smem = max(smem, b);
  }
  if (smem > 200 * 1024) return RB2_ERR_ARG;
  RB2_CHECK(cudaFuncSetAttribute(refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  refit_kernel<<<n_problems, 32, smem, stream>>>(d_desc, d_winner, minimum_fit_error, d_out, d_matched_x,
                                                 d_matched_y, d_residual, max_upeaks);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}
