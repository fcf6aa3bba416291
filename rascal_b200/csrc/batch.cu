// Device-resident glue of the many-spectra path (sm_100a): the host work the reference does
// between its stages, moved onto the device so that a batch of spectra goes
//   peaks/lines -> pairs -> K1 -> K2 -> [prepare] -> K3 -> K4 -> K5
// without a host round trip.
//
//   rb2_expand_pairs     Calibrator._generate_pairs (calibrator.py:86-107, no polygon mask):
//                        peak-major Cartesian product, plus the per-peak CSR view K2 reads.
//   rb2_ransac_prepare   the set-up at the top of _solve_candidate_ransac (calibrator.py:441-523):
//                        peaks that have candidates, candidate CSR, wavelength ids, the sampling
//                        law as a 32-bit CDF (:521-523 incl. its index -1 wrap-around), the
//                        monotonicity range (:585-586) and the Hough-density weight tables
//                        (:497-506) - written straight into K3's descriptors.
#include "common.cuh"

namespace rb2 {

// ---- pairs -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
expand_pairs_kernel(const rb2_pairs_desc* __restrict__ descs) {
  const rb2_pairs_desc d = descs[blockIdx.y];
  const int n_p = d.n_peaks, n_a = d.n_lines;
  const long long P = (long long)n_p * n_a;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i / n_a), k = (int)(i - (long long)u * n_a);
    d.d_pair_x[i] = d.d_peaks[u];
    d.d_pair_y[i] = d.d_lines[k];
    if (d.d_pair_wid) d.d_pair_wid[i] = k;
  }
  if (d.d_pair_off)
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u <= n_p; u += gridDim.x * blockDim.x) d.d_pair_off[u] = u * n_a;
}

// ---- K3 set-up ---------------------------------------------------------------------------
constexpr int PREP_THREADS = 256;

// one CTA per spectrum
__global__ void __launch_bounds__(PREP_THREADS)
ransac_prepare_kernel(const rb2_prepare_desc* __restrict__ descs, rb2_ransac_desc* __restrict__ out_descs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_nu, s_nc, s_nw;
  const rb2_prepare_desc d = descs[blockIdx.x];
  rb2_ransac_desc& o = out_descs[blockIdx.x];
  const int tid = threadIdx.x;
  const int n_p = d.n_peaks, top_n = d.top_n;
  // the tables this kernel fills live where K3's descriptor already points
  double* peaks = const_cast<double*>(o.d_peaks);
  int* coff = const_cast<int*>(o.d_cand_off);
  double* cwave = const_cast<double*>(o.d_cand_wave);
  int* cwid = const_cast<int*>(o.d_cand_wid);
  unsigned* cdf32 = const_cast<unsigned*>(o.d_cdf32);
  double* ppb = const_cast<double*>(o.d_pp_breaks);
  double* ppc = const_cast<double*>(o.d_pp_coef);
  int* first = reinterpret_cast<int*>(smem_raw);  // [n_p * top_n] 1 = first occurrence of its wavelength

  // 1. peaks with candidates -> compacted peaks + CSR (warp 0, 32 peaks per step)
  if (tid < 32) {
    int nu = 0, nc = 0;
    for (int base = 0; base < n_p; base += 32) {
      const int u = base + tid;
      const int c = (u < n_p) ? min(max(d.d_cand_n[u], 0), top_n) : 0;
      const unsigned keep = __ballot_sync(0xffffffffu, c > 0);
      int incl = c;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, s);
        if (tid >= s) incl += t;
      }
      if (c > 0) {
        const int slot = nu + __popc(keep & ((1u << tid) - 1u));
        const int off = nc + incl - c;
        peaks[slot] = d.d_upeak_x[u];
        coff[slot] = off;
        for (int j = 0; j < c; ++j) cwave[off + j] = d.d_cand_wave[(size_t)u * top_n + j];
      }
      nu += __popc(keep);
      nc += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (tid == 0) { coff[nu] = nc; s_nu = nu; s_nc = nc; s_nw = 0; }
  }
  __syncthreads();
  const int nu = s_nu, nc = s_nc;

  // 2. wavelength ids = rank among the distinct candidate wavelengths (np.unique(..., return_inverse));
  //    the candidate wavelengths are staged in shared memory for the two all-pairs passes
  double* s_w = reinterpret_cast<double*>(first + ((nc + 1) & ~1));
  for (int i = tid; i < nc; i += PREP_THREADS) s_w[i] = cwave[i];
  __syncthreads();
  for (int i = tid; i < nc; i += PREP_THREADS) {
    const double v = s_w[i];
    int f = 1;
    for (int j = 0; j < i; ++j) if (s_w[j] == v) { f = 0; break; }
    first[i] = f;
  }
  __syncthreads();
  for (int i = tid; i < nc; i += PREP_THREADS) {
    const double v = s_w[i];
    int r = 0;
    for (int j = 0; j < nc; ++j) r += (first[j] && s_w[j] < v);
    cwid[i] = r;
    if (first[i]) atomicAdd(&s_nw, 1);
  }

  // 3. Hough-density weight tables: the pp-form of the interpolating cubic spline of hist[:, 0] is a
  //    LINEAR map of that row (matrix d_spline_op, built once per xbins on the host from SciPy's
  //    not-a-knot spline on the integer grid); scaled to the gradient-centre axis here.  The operator is
  //    banded in practice but stored dense: 8 lanes share one row (coalesced 64-byte reads of the row,
  //    the strided histogram column staged in shared memory first)
  const int xb = d.xbins, yb = d.ybins, npp = d.n_pp;
  const double hx = (d.d_xedges[1] - d.d_xedges[0]) / 2.0, hy = (d.d_yedges[1] - d.d_yedges[0]) / 2.0;
  const double gc0 = d.d_xedges[1] - hx, gcn = d.d_xedges[xb] - hx, ic0 = d.d_yedges[1] - hy;
  const double step = (gcn - gc0) / (double)(xb - 1);
  double* s_h0 = s_w + nc;  // [xb] first intercept row of the histogram
  __syncthreads();
  for (int i = tid; i < xb; i += PREP_THREADS) s_h0[i] = (double)d.d_hist[(size_t)i * yb];
  __syncthreads();
  {
    const int sub = tid & 7;
    for (int r = tid >> 3; r < npp * 4; r += PREP_THREADS / 8) {
      const double* row = d.d_spline_op + (size_t)r * xb;
      double acc = 0.0;
      // the partial sums are combined in a fixed order, so the table does not depend on the launch
      for (int i = sub; i < xb; i += 8) acc = fma(row[i], s_h0[i], acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (sub == 0) {
        const int k = r & 3;
        double sc = 1.0;
        for (int q = 0; q < k; ++q) sc *= step;
        ppc[r] = acc / sc;
      }
    }
  }
  for (int j = tid; j <= npp; j += PREP_THREADS) ppb[j] = (j == npp) ? gcn : gc0 + d.d_spline_brk[j] * step;
  __syncthreads();

  // 4. scalars, sampling law and the weight's upper bound.  The law (calibrator.py:521-523) is restated bit
  //    for bit from the host's sampling_cdf32: the three running sums are sequential there, so they are
  //    sequential here (thread 0: additions only), everything around them runs over the peaks in parallel
  const bool enough = (nu > o.deg) && (nu > o.sample_size);
  __shared__ double s_edge[11];
  __shared__ int s_count[10], s_lawbad;
  __shared__ double s_total, s_last;
  double* s_q = s_h0 + xb;  // [nu] weight / running sums per peak
  if (tid < 10) s_count[tid] = 0;
  if (tid == 0) s_lawbad = 0;
  if (enough && tid < 11) {
    double lo = peaks[0], hi = peaks[nu - 1];
    if (lo == hi) { lo -= 0.5; hi += 0.5; }
    const Linspace edges(lo, hi, 11);
    s_edge[tid] = edges.at(tid);
  }
  __syncthreads();
  if (enough) {
    for (int i = tid; i < nu; i += PREP_THREADS) {  // np.histogram: searchsorted(edges, v, 'right') - 1, last edge inclusive
      int k = 0;
      while (k < 11 && s_edge[k] <= peaks[i]) ++k;
      atomicAdd(&s_count[min(k - 1, 9)], 1);
    }
    __syncthreads();
    for (int i = tid; i < nu; i += PREP_THREADS) {  // np.digitize(v, edges, right=True) - 1; index -1 wraps to the last bin
      int k = 0;
      while (k < 11 && s_edge[k] < peaks[i]) ++k;
      const int dig = (k - 1 < 0) ? 9 : min(k - 1, 9);
      // a peak exactly on an interior edge whose left bin is empty: weight 1/0, prob = NaN - the reference's
      // np.random.choice raises "probabilities contain NaN" on the first draw (calibrator.py:538)
      if (s_count[dig] == 0) s_lawbad = 1;
      s_q[i] = 1.0 / (double)s_count[dig];
    }
    __syncthreads();
    if (tid == 0) {
      double total = 0.0;
      for (int i = 0; i < nu; ++i) total = add(total, s_q[i]);
      s_total = total;
    }
    __syncthreads();
    for (int i = tid; i < nu; i += PREP_THREADS) s_q[i] = s_q[i] / s_total;
    __syncthreads();
    if (tid == 0) {  // cdf = cumsum(prob); its last value normalises once more (cdf /= cdf[-1])
      double run = 0.0;
      for (int i = 0; i < nu; ++i) { run = add(run, s_q[i]); s_q[i] = run; }
      s_last = run;
    }
    __syncthreads();
    const bool law_ok = s_lawbad == 0;
    for (int i = tid; i < nu; i += PREP_THREADS) {
      const double c = floor(mul(s_q[i] / s_last, 4294967296.0));
      unsigned v = (c >= 4294967295.0) ? 0xFFFFFFFFu : (unsigned)c;
      if (i == nu - 1) v = 0xFFFFFFFFu;
      cdf32[i] = law_ok ? v : 0u;
    }
  } else {
    for (int i = tid; i < nu; i += PREP_THREADS) cdf32[i] = 0u;
  }
  // weight upper bound (engine.weight_upper_bound): Bernstein coefficients enclose each cubic piece
  __shared__ double s_lo[PREP_THREADS / 32], s_hi[PREP_THREADS / 32];
  {
    double lo_v = CUDART_INF, hi_v = -CUDART_INF;
    for (int j = tid; j < npp; j += PREP_THREADS) {
      const double h = ppb[j + 1] - ppb[j];
      const double c0 = ppc[4 * j], c1 = ppc[4 * j + 1], c2 = ppc[4 * j + 2], c3 = ppc[4 * j + 3];
      const double b0 = c0, b1 = c0 + c1 * h / 3.0, b2 = c0 + 2.0 * c1 * h / 3.0 + c2 * h * h / 3.0,
                   b3 = c0 + c1 * h + c2 * h * h + c3 * h * h * h;
      lo_v = fmin(lo_v, fmin(fmin(b0, b1), fmin(b2, b3)));
      hi_v = fmax(hi_v, fmax(fmax(b0, b1), fmax(b2, b3)));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo_v = fmin(lo_v, __shfl_xor_sync(0xffffffffu, lo_v, off));
      hi_v = fmax(hi_v, __shfl_xor_sync(0xffffffffu, hi_v, off));
    }
    if ((tid & 31) == 0) { s_lo[tid >> 5] = lo_v; s_hi[tid >> 5] = hi_v; }
  }
  __syncthreads();
  if (tid == 0) {
    o.n_upeaks = nu; o.n_cand = nc; o.n_wids = max(s_nw, 1);
    o.gc_min = gc0; o.gc_max = gcn; o.ic_min = ic0;
    o.h_lo = (double)d.d_hist[0]; o.h_hi = (double)d.d_hist[(size_t)(xb - 1) * yb];
    if (nu > 0) {
      const double ptp = sub(peaks[nu - 1], peaks[0]);
      o.pix_min = sub(peaks[0], mul(ptp, 0.2));
      o.pix_max = add(peaks[nu - 1], mul(ptp, 0.2));
    } else {
      o.pix_min = 0.0; o.pix_max = 0.0;
    }
    const bool law_ok = s_lawbad == 0;
    if (!enough || !law_ok) o.n_hyp = 0;
    if (d.d_enough) d.d_enough[0] = !enough ? 0 : (law_ok ? 1 : -1);
    double lo_v = fmin(o.h_lo, o.h_hi), hi_v = fmax(o.h_lo, o.h_hi);
    for (int w = 0; w < PREP_THREADS / 32; ++w) { lo_v = fmin(lo_v, s_lo[w]); hi_v = fmax(hi_v, s_hi[w]); }
    const bool bounded = (lo_v >= 0.0) && (hi_v > 0.0) && (o.hough_weight > 0.0) && isfinite(hi_v);
    o.w_upper = bounded ? o.hough_weight * (double)o.n_pix * hi_v * (1.0 + 1e-9) : -1.0;
  }
}

}  // namespace rb2

extern "C" int rb2_expand_pairs(int n_problems, const rb2_pairs_desc* d_desc, const rb2_pairs_desc* h_desc,
                                void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_desc || !h_desc) return RB2_ERR_ARG;
  long long max_p = 1;
  for (int i = 0; i < n_problems; ++i) {
    const rb2_pairs_desc& h = h_desc[i];
    if (h.n_peaks < 0 || h.n_lines < 0 || (long long)h.n_peaks * h.n_lines > 0x7fffffffll) return RB2_ERR_ARG;
    if (h.n_peaks && h.n_lines && (!h.d_peaks || !h.d_lines || !h.d_pair_x || !h.d_pair_y)) return RB2_ERR_ARG;
    max_p = max(max_p, (long long)h.n_peaks * h.n_lines);
  }
  const int bx = (int)min((max_p + 255) / 256, 64ll);
  expand_pairs_kernel<<<dim3(bx, n_problems), 256, 0, (cudaStream_t)stream_>>>(d_desc);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

extern "C" int rb2_ransac_prepare(int n_problems, const rb2_prepare_desc* d_desc, const rb2_prepare_desc* h_desc,
                                  rb2_ransac_desc* d_ransac_desc, void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_desc || !h_desc || !d_ransac_desc) return RB2_ERR_ARG;
  size_t smem = 16;
  for (int i = 0; i < n_problems; ++i) {
    const rb2_prepare_desc& h = h_desc[i];
    if (h.n_peaks < 0 || h.top_n < 1 || h.top_n > RB2_MAX_TOPN || h.xbins < 4 || h.ybins < 1 || h.n_pp < 1) return RB2_ERR_ARG;
    if (!h.d_cand_wave || !h.d_cand_n || !h.d_upeak_x || !h.d_xedges || !h.d_yedges || !h.d_hist || !h.d_spline_op ||
        !h.d_spline_brk)
      return RB2_ERR_ARG;
    // first[] (int per candidate) | candidate wavelengths | first histogram row | per-peak sums
    smem = max(smem, (size_t)(h.n_peaks * h.top_n + 2) * 4 + (size_t)h.n_peaks * h.top_n * 8 + (size_t)h.xbins * 8 +
                         (size_t)h.n_peaks * 8 + 32);
  }
  if (smem > 200 * 1024) return RB2_ERR_ARG;  // > 50k candidate entries: prepare those on the host
  RB2_CHECK(cudaFuncSetAttribute(ransac_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ransac_prepare_kernel<<<n_problems, PREP_THREADS, smem, (cudaStream_t)stream_>>>(d_desc, d_ransac_desc);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}
