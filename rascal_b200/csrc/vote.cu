// K2 — candidate vote for sm_100a.
//
// Replaces Calibrator._get_candidate_points_linear (calibrator.py:219-261) and
// Calibrator._get_most_common_candidates (calibrator.py:154-217).
//
// The reference loops over ALL B = xbins*ybins ranked bin-centre lines and, for
// each, tests all P pairs (B*P evaluations; 5e9 for a 1000x1000 histogram).
// Here one CTA owns one unique peak and one thread one (peak, atlas-line) pair.
// For a fixed pair and gradient row the predicate
//       | fl(fl(g*x) + c_j) - lambda | <= tol
// is monotone in the intercept-bin index j, so the hit bins of the row are one
// short window that is located directly (float guess + exact fix-up with the
// reference's own arithmetic): P*xbins window searches instead of B*P tests.
//
// Per peak the reference then needs
//   agg[w]  = sum of Gaussian weights of every hit of wavelength w (any order),
//   U       = sorted distinct hit wavelengths,  n = min(top_n, |U|),
//   result  = W[argsort(-agg)[:n]]   -- W = the hit list ordered by (bin rank,
//             pair order): indices into U applied to W (the App. A.3 quirk).
// agg is accumulated per thread in a fixed (bin-index) order: deterministic, no
// atomics on the value path.  Only the first max(j)+1 entries of W are needed:
// a rank-bucket histogram finds the rank threshold that covers them, a second
// enumeration collects just those hits and a shared-memory bitonic sort orders
// them by (rank, pair).
#include "common.cuh"

namespace rb2 {

constexpr int VOTE_THREADS = 256;
constexpr int VOTE_NB = 4096;    // rank buckets
constexpr int VOTE_CAP = 4096;   // collected (rank, pair) keys

struct VoteCtx {
  const double* cc;   // smem [yb] intercept-bin centres fl(ye[j] + hy) (houghtransform.py:205-208)
  const double* ye;   // smem [yb+1]
  const double* gc;   // smem [xb] gradient-bin centres
  double hy, x, tol, inv_wy, c_first, c_last, inv_dg;
  int xb, yb;
};

// exp(z) for the Gaussian vote weight, z = -(lam - pred)^2 / (2 sigma^2 + 1e-9) <= 0.  With the
// reference's sigma = 1.1775 * (range_tolerance + linearity_tolerance) (hundreds of Angstrom) against a
// candidate tolerance of a few Angstrom, |z| is ~1e-4: there the degree-6 Taylor polynomial is exact
// to the last bit or two (truncation z^7/5040 < 1e-22 for |z| < 2^-9), the same 1-ulp class as exp();
// libdevice's exp() costs ~10x more and was 28 % of the kernel's instructions.
__device__ __forceinline__ double gauss_weight(double z) {
  if (z > -0.001953125) {
    double p = 1.0 / 720.0;
    p = fma(p, z, 1.0 / 120.0);
    p = fma(p, z, 1.0 / 24.0);
    p = fma(p, z, 1.0 / 6.0);
    p = fma(p, z, 0.5);
    p = fma(p, z, 1.0);
    return fma(p, z, 1.0);
  }
  return exp(z);
}

// Rows that can hold a hit of one pair: g*x + c_j within tol of lam for some bin centre c_j in
// [c_first, c_last] <=> g*x in [lam - c_last - tol, lam - c_first + tol].  The centres gc[] are equally
// spaced and increasing, so that is ONE contiguous row range, located arithmetically with a row of
// slack on each side (the exact window test in row_hits decides; skipped rows have no hits).
__device__ __forceinline__ void row_range(const VoteCtx& c, double lam, int& g_begin, int& g_end) {
  g_begin = 0; g_end = c.xb;
  if (c.xb > 2 && c.x != 0.0 && c.inv_dg > 0.0) {
    const double lo_t = lam - c.c_last - c.tol, hi_t = lam - c.c_first + c.tol;
    double ga = lo_t / c.x, gb = hi_t / c.x;
    if (ga > gb) { const double t = ga; ga = gb; gb = t; }
    const double slack = 1e-9 * (fabs(ga) + fabs(gb)) + 1e-300;
    const double ra = ((ga - slack) - c.gc[0]) * c.inv_dg, rb = ((gb + slack) - c.gc[0]) * c.inv_dg;
    if (ra == ra && rb == rb) {  // NaN / inf inputs keep the full range
      g_begin = (int)fmax(0.0, fmin((double)c.xb, floor(ra) - 1.0));
      g_end = (int)fmax(0.0, fmin((double)c.xb, ceil(rb) + 2.0));
      if (g_end < g_begin) g_end = g_begin;
    }
  }
}

// The hit bins of one pair in gradient row gi; f(gi, j, pred) for every hit, ascending j.
template <class F>
__device__ __forceinline__ void row_hits(const VoteCtx& c, double lam, int gi, F&& f) {
  const double gx = mul(c.gc[gi], c.x);
  // first j whose centre line is within reach: (gx + c_j) - lam >= -tol
  int j = (int)floor(((lam - gx) - c.tol - c.c_first) * c.inv_wy);  // the step-back loop below repairs a high guess
  j = max(0, min(c.yb, j));
  while (j > 0) {  // guess too high: step back while the previous bin is not below the window
    const double df = sub(add(gx, c.cc[j - 1]), lam);
    if (df < -c.tol) break;
    --j;
  }
  for (; j < c.yb; ++j) {
    const double pred = add(gx, c.cc[j]);
    const double df = sub(pred, lam);
    if (df > c.tol) break;
    if (fabs(df) <= c.tol) f(gi, j, pred);
  }
}

// Work units (pair, row) of one peak, flattened: pair q owns units [uoff[q], uoff[q+1]) = its rows
// ubeg[q].. in ascending order; thread t takes the contiguous slice [t*chunk, (t+1)*chunk).  Calls
// body(q, gi) for every unit of the slice in order and edge(q, first_of_pair, last_of_pair) before /
// after the run of units of pair q inside the slice.
template <class Begin, class Body, class End>
__device__ __forceinline__ void for_my_units(const int* uoff, const int* ubeg, int m, int tid, Begin&& begin, Body&& body,
                                             End&& end) {
  const int U = uoff[m];
  const int chunk = (U + VOTE_THREADS - 1) / VOTE_THREADS;
  const int u0 = tid * chunk, u1 = min(U, u0 + chunk);
  if (u0 >= u1) return;
  int lo = 0, hi = m - 1;  // pair that contains unit u0: last q with uoff[q] <= u0
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (uoff[mid] <= u0) lo = mid; else hi = mid - 1; }
  // ONE loop over the units, the change of pair handled inside it: every lane of a warp reaches body()
  // (the window search) at the same point of every iteration instead of each lane nesting its own
  // per-pair loops
  int q = lo;
  bool open = false;
  for (int uu = u0; uu < u1; ++uu) {
    if (!open || uu >= uoff[q + 1]) {
      if (open) end(q, /*starts_here=*/uoff[q] >= u0, /*ends_here=*/true);
      while (uoff[q + 1] <= uu) ++q;  // skip pairs without rows
      begin(q);
      open = true;
    }
    body(q, ubeg[q] + (uu - uoff[q]));
  }
  if (open) end(q, /*starts_here=*/uoff[q] >= u0, /*ends_here=*/uoff[q + 1] <= u1);
}

__global__ void __launch_bounds__(VOTE_THREADS)
vote_kernel(const rb2_vote_desc* __restrict__ descs, int ucap, int lam_off) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const rb2_vote_desc d = descs[blockIdx.y];
  const int u = blockIdx.x;
  if (u >= d.n_upeaks) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int xb = d.xbins, yb = d.ybins, nw = d.n_uwaves;
  const int B = xb * yb;

  double* ye = reinterpret_cast<double*>(smem_raw);
  double* gc = ye + (yb + 1);
  double* cc = gc + xb;
  double* wagg = cc + yb;
  unsigned long long* list = reinterpret_cast<unsigned long long*>(wagg + nw);
  int* wcnt = reinterpret_cast<int*>(list + VOTE_CAP);
  unsigned* rb_hist = reinterpret_cast<unsigned*>(wcnt + nw);
  int* ubeg = reinterpret_cast<int*>(rb_hist + VOTE_NB);   // [ucap] first candidate row of pair q
  int* uoff = ubeg + ucap;                                 // [ucap + 1] exclusive prefix of the row counts
  double* lam_s = reinterpret_cast<double*>(smem_raw + lam_off);  // [ucap] this peak's pair wavelengths
  __shared__ double pf_agg[VOTE_THREADS];                  // partial of a pair that began in an earlier thread
  __shared__ int pf_q[VOTE_THREADS], pf_cnt[VOTE_THREADS];
  __shared__ int sel_wid[RB2_MAX_TOPN];
  __shared__ int sel_j[RB2_MAX_TOPN];
  __shared__ unsigned scan_s[VOTE_THREADS / 32];
  __shared__ int s_L, s_tb, s_list_n, s_need;

  const double hx = (d.d_xedges[1] - d.d_xedges[0]) / 2;  // houghtransform.py:197
  const double hy = (d.d_yedges[1] - d.d_yedges[0]) / 2;  // houghtransform.py:198
  for (int i = tid; i <= yb; i += VOTE_THREADS) ye[i] = d.d_yedges[i];
  for (int i = tid; i < xb; i += VOTE_THREADS) gc[i] = add(d.d_xedges[i], hx);
  for (int i = tid; i < yb; i += VOTE_THREADS) cc[i] = add(d.d_yedges[i], hy);
  for (int i = tid; i < nw; i += VOTE_THREADS) { wagg[i] = 0.0; wcnt[i] = 0; }
  for (int i = tid; i < VOTE_NB; i += VOTE_THREADS) rb_hist[i] = 0u;
  if (tid == 0) { s_list_n = 0; s_L = 0; }
  __syncthreads();

  int shift = 0;
  while ((B >> shift) > VOTE_NB) ++shift;

  VoteCtx c;
  c.ye = ye; c.gc = gc; c.cc = cc; c.hy = hy; c.x = d.d_upeak_x[u]; c.tol = d.tol;
  c.xb = xb; c.yb = yb;
  c.c_first = ye[0] + hy;
  c.c_last = ye[yb - 1] + hy;
  c.inv_dg = (xb > 1 && gc[xb - 1] > gc[0]) ? (double)(xb - 1) / (gc[xb - 1] - gc[0]) : 0.0;
  c.inv_wy = (double)yb / (ye[yb] - ye[0]);

  const int q0 = d.d_pair_off[u];
  const int m = d.d_pair_off[u + 1] - q0;
  const int* __restrict__ rankpos = d.d_rankpos;

  // ---- phase 0: candidate row range of every pair, flattened into (pair, row) work units ----
  // (pairs whose line never crosses the histogram's intercept range have no units: without this
  //  a CTA waits at the barrier for the few threads whose pairs span all rows)
  const bool flat = m <= ucap;
  bool inv = false;  // ordered hit list from the bins in rank order (phase 3'): needs ascending pair wavelengths
  if (flat) {
    int ascending = 1;
    for (int q = tid; q < m; q += VOTE_THREADS) {
      const double lam = d.d_pair_y[q0 + q];
      int gb, ge;
      row_range(c, lam, gb, ge);
      ubeg[q] = gb; uoff[q + 1] = ge - gb;
      lam_s[q] = lam;
      if (q > 0 && !(lam >= d.d_pair_y[q0 + q - 1])) ascending = 0;
    }
    pf_q[tid] = -1;
    inv = __syncthreads_and(ascending) && d.d_rank != nullptr;
    if (warp == 0) {  // exclusive scan, 32 pairs per step
      int run = 0;
      for (int base = 0; base < m; base += 32) {
        const int q = base + lane;
        const int v = (q < m) ? uoff[q + 1] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (q < m) uoff[q + 1] = run + incl;
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) uoff[0] = 0;
    }
    __syncthreads();
  }

  // ---- phase 1: aggregate weights and counts, rank-bucket histogram ----------
  const double inv_denom = 1.0 / d.gauss_denom;
  auto vote_hit = [&](double lam, double& agg, int& cnt, int gi, int j, double pred) {
    const double t = sub(lam, pred);
    const double w = gauss_weight(-mul(t, t) * inv_denom);  // util.py:447, a = 1.0 (z to 1 ulp: w to 1e-20)
    agg = add(agg, w);
    ++cnt;
    if (!inv) atomicAdd(&rb_hist[rankpos[gi * yb + j] >> shift], 1u);
  };
  auto commit = [&](int q, double agg, int cnt) {
    if (cnt) {
      const int wid = d.d_pair_wid[q0 + q];
      atomicAdd(&wagg[wid], d.weighted ? agg : (double)cnt);
      atomicAdd(&wcnt[wid], cnt);
    }
  };
  if (flat) {
    double lam = 0.0, agg = 0.0;
    int cnt = 0;
    int own_q = -1, own_cnt = 0;  // a pair that starts in this slice and continues in the next threads
    double own_agg = 0.0;
    for_my_units(uoff, ubeg, m, tid,
                 [&](int q) { lam = d.d_pair_y[q0 + q]; agg = 0.0; cnt = 0; },
                 [&](int q, int gi) { row_hits(c, lam, gi, [&](int g, int j, double pred) { vote_hit(lam, agg, cnt, g, j, pred); }); },
                 [&](int q, bool starts_here, bool ends_here) {
                   if (starts_here && ends_here) commit(q, agg, cnt);
                   else if (starts_here) { own_q = q; own_agg = agg; own_cnt = cnt; }
                   else { pf_q[tid] = q; pf_agg[tid] = agg; pf_cnt[tid] = cnt; }  // continuation of an earlier thread's pair
                 });
    __syncthreads();
    if (own_q >= 0) {  // add the continuations in thread order: the sum does not depend on timing
      for (int t = tid + 1; t < VOTE_THREADS && pf_q[t] == own_q; ++t) { own_agg = add(own_agg, pf_agg[t]); own_cnt += pf_cnt[t]; }
      commit(own_q, own_agg, own_cnt);
    }
  } else {
    for (int q = tid; q < m; q += VOTE_THREADS) {
      const double lam = d.d_pair_y[q0 + q];
      double agg = 0.0;
      int cnt = 0;
      int gb, ge;
      row_range(c, lam, gb, ge);
      for (int gi = gb; gi < ge; ++gi) row_hits(c, lam, gi, [&](int g, int j, double pred) { vote_hit(lam, agg, cnt, g, j, pred); });
      commit(q, agg, cnt);
    }
  }
  __syncthreads();
  if (d.d_agg) {
    for (int i = tid; i < nw; i += VOTE_THREADS) {
      d.d_agg[(size_t)u * nw + i] = wagg[i];
      d.d_cnt[(size_t)u * nw + i] = wcnt[i];
    }
  }
  // L = number of distinct hit wavelengths
  {
    int l = 0;
    for (int i = tid; i < nw; i += VOTE_THREADS) l += (wcnt[i] > 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (lane == 0 && l) atomicAdd(&s_L, l);
  }
  __syncthreads();
  const int L = s_L;
  const int n = min(min(d.top_n, RB2_MAX_TOPN), L);
  if (tid == 0) d.d_cand_n[u] = n;
  if (n == 0) return;

  // ---- phase 2: top-n wavelengths by aggregated weight (stable) --------------
  // one warp does the n arg-max passes over the (few hundred) wavelengths: no block barriers
  if (warp == 0) {
    for (int k = 0; k < n; ++k) {
      double bv = -CUDART_INF;
      int bi = 0x7fffffff;
      for (int i = lane; i < nw; i += 32) {
        if (wcnt[i] > 0) {
          const double v = wagg[i];
          if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        sel_wid[k] = bi;
        wcnt[bi] = -wcnt[bi];  // mark as taken (still counts as "hit" below via != 0)
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // U-index of each selected wavelength = number of hit wavelengths with a smaller id
  for (int k = warp; k < n; k += VOTE_THREADS / 32) {
    const int wid = sel_wid[k];
    int cntl = 0;
    for (int i = lane; i < wid; i += 32) cntl += (wcnt[i] != 0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cntl += __shfl_xor_sync(0xffffffffu, cntl, o);
    if (lane == 0) sel_j[k] = cntl;
  }
  __syncthreads();
  if (tid == 0) {
    int need = 0;
    for (int k = 0; k < n; ++k) need = max(need, sel_j[k] + 1);
    s_need = need;
  }
  __syncthreads();
  const unsigned need = (unsigned)s_need;

  // ---- phase 3': the first `need` entries of W straight from the bins in rank order ----------
  // A bin (g, j) is hit by the pairs whose wavelength lies within tol of its line at this peak:
  // with ascending wavelengths that is one contiguous run of pairs, found by bisection with the
  // very test row_hits applies.  256 bins per step, a block scan places their hits; W is complete
  // after the few hundred best bins - no second enumeration of the windows, no sort.
  if (inv) {
    int* W = reinterpret_cast<int*>(list);
    const int* __restrict__ rank = d.d_rank;
    unsigned running = 0;
    for (int base = 0; base < B && running < need; base += VOTE_THREADS) {
      const int r = base + tid;
      int qa = 0, nh = 0;
      if (r < B) {
        const int b = rank[r];
        const int g = b / yb, j = b - g * yb;
        const double pred = add(mul(gc[g], c.x), cc[j]);
        int lo = 0, hi = m;  // first pair that is not below the window: pred - lam <= tol
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (sub(pred, lam_s[mid]) > c.tol) lo = mid + 1; else hi = mid;
        }
        qa = lo;
        while (qa + nh < m && fabs(sub(pred, lam_s[qa + nh])) <= c.tol) ++nh;
      }
      unsigned incl = (unsigned)nh;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) scan_s[warp] = incl;
      __syncthreads();
      unsigned wpre = 0, total = 0;
      for (int w = 0; w < VOTE_THREADS / 32; ++w) { if (w < warp) wpre += scan_s[w]; total += scan_s[w]; }
      const unsigned pos = running + wpre + incl - (unsigned)nh;
      for (int t = 0; t < nh; ++t) if (pos + t < need) W[pos + t] = qa + t;
      running += total;
      __syncthreads();
    }
    for (int k = tid; k < n; k += VOTE_THREADS) {
      const int q = W[sel_j[k]];
      d.d_cand_wave[(size_t)u * d.top_n + k] = d.d_pair_y[q0 + q];
      d.d_cand_pair[(size_t)u * d.top_n + k] = q0 + q;
    }
    return;
  }

  // ---- phase 3: rank threshold that covers the first `need` entries of W -----
  {
    constexpr int PER = VOTE_NB / VOTE_THREADS;
    unsigned loc[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { loc[j] = rb_hist[tid * PER + j]; sum += loc[j]; }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) scan_s[warp] = incl;
    __syncthreads();
    unsigned wpre = 0;
    for (int w = 0; w < warp; ++w) wpre += scan_s[w];
    unsigned run = wpre + incl - sum;  // exclusive prefix of this thread's first bucket
    if (tid == 0) s_tb = VOTE_NB - 1;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const unsigned before = run;
      run += loc[j];
      if (before < need && run >= need) s_tb = tid * PER + j;  // unique writer
    }
    __syncthreads();
  }
  const int tb = s_tb;

  // ---- phase 4: collect hits up to the threshold bucket, sort, pick W[j] ------
  auto collect = [&](int q, int gi, int j) {
    const unsigned r = (unsigned)rankpos[gi * yb + j];
    if ((int)(r >> shift) <= tb) {
      const int pos = atomicAdd(&s_list_n, 1);
      if (pos < VOTE_CAP) list[pos] = ((unsigned long long)r << 32) | (unsigned)q;
    }
  };
  if (flat) {
    double lam = 0.0;
    for_my_units(uoff, ubeg, m, tid, [&](int q) { lam = d.d_pair_y[q0 + q]; },
                 [&](int q, int gi) { row_hits(c, lam, gi, [&](int g, int j, double) { collect(q, g, j); }); },
                 [&](int, bool, bool) {});
  } else {
    for (int q = tid; q < m; q += VOTE_THREADS) {
      const double lam = d.d_pair_y[q0 + q];
      int gb, ge;
      row_range(c, lam, gb, ge);
      for (int gi = gb; gi < ge; ++gi) row_hits(c, lam, gi, [&](int g, int j, double) { collect(q, g, j); });
    }
  }
  __syncthreads();
  int ln = s_list_n;
  if (ln > VOTE_CAP) {
    if (tid == 0) atomicExch(d.d_status, (int)RB2_ERR_CAPACITY);
    ln = VOTE_CAP;
  }
  if (ln <= 1024) {
    // only the entries at the n positions sel_j[] of the sorted list are needed: position of an
    // entry = number of smaller keys (keys are distinct), found by counting - one barrier, no sort
    for (int e = tid; e < ln; e += VOTE_THREADS) {
      const unsigned long long key = list[e];
      int pos = 0;
      for (int f = 0; f < ln; ++f) pos += (list[f] < key);
      for (int k = 0; k < n; ++k) {
        if (sel_j[k] == pos) {
          const int q = (int)(key & 0xffffffffu);
          d.d_cand_wave[(size_t)u * d.top_n + k] = d.d_pair_y[q0 + q];
          d.d_cand_pair[(size_t)u * d.top_n + k] = q0 + q;
        }
      }
    }
    return;
  }
  int np2 = 1;
  while (np2 < ln) np2 <<= 1;
  for (int i = ln + tid; i < np2; i += VOTE_THREADS) list[i] = ~0ull;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += VOTE_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = list[i], b = list[ixj];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { list[i] = b; list[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int k = tid; k < n; k += VOTE_THREADS) {
    const int q = (int)(list[sel_j[k]] & 0xffffffffu);
    d.d_cand_wave[(size_t)u * d.top_n + k] = d.d_pair_y[q0 + q];
    d.d_cand_pair[(size_t)u * d.top_n + k] = q0 + q;
  }
}

}  // namespace rb2

extern "C" int rb2_candidate_vote(int n_problems, const rb2_vote_desc* d_desc, const rb2_vote_desc* h_desc,
                                  void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_desc || !h_desc) return RB2_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  int max_up = 0, max_nw = 0;
  size_t smem = 0;
  for (int i = 0; i < n_problems; ++i) {
    const rb2_vote_desc& h = h_desc[i];
    if (h.top_n < 1 || h.top_n > RB2_MAX_TOPN || h.xbins < 1 || h.ybins < 1) return RB2_ERR_ARG;
    max_up = max(max_up, h.n_upeaks);
    max_nw = max(max_nw, h.n_uwaves);
    size_t s = (size_t)(2 * h.ybins + 1) * 8 + (size_t)h.xbins * 8 + (size_t)h.n_uwaves * 8 + (size_t)VOTE_CAP * 8 +
               (size_t)h.n_uwaves * 4 + (size_t)VOTE_NB * 4;
    smem = max(smem, s);
  }
  if (max_up == 0) return RB2_OK;
  // flattened work units: room for the row tables of max_nw + 64 pairs per peak (a peak with more pairs,
  // i.e. duplicated atlas lines beyond that, takes the pair-per-thread path)
  const int ucap = max_nw + 64;
  smem += (size_t)(2 * ucap + 1) * 4 + 16;
  const int lam_off = (int)((smem + 7) & ~(size_t)7);  // this peak's pair wavelengths, after the row tables
  smem = (size_t)lam_off + (size_t)ucap * 8;
  if (smem > 200 * 1024) return RB2_ERR_ARG;
  RB2_CHECK(cudaFuncSetAttribute(vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(max_up, n_problems);
  vote_kernel<<<grid, VOTE_THREADS, smem, stream>>>(d_desc, ucap, lam_off);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

// ---- Calibrator.candidates (calibrator.py:239-261), materialised on demand --------------------------------
// For every Hough line in rank order: the pairs within candidate_tolerance of the line (mask `<=`, pairs in
// the caller's order) and their Gaussian weights - the list the reference builds before the vote and that
// its plotting code reads.  One warp per line; lanes sweep the pairs, ordered compaction by ballot.
namespace rb2 {

__global__ void __launch_bounds__(256)
candidate_points_kernel(const double* __restrict__ xedges, const double* __restrict__ yedges, int ybins,
                        const int* __restrict__ rank, int n_bins, const double* __restrict__ px,
                        const double* __restrict__ py, int n_pairs, double tol, double gauss_denom,
                        long long* __restrict__ off, double* __restrict__ out_x, double* __restrict__ out_y,
                        double* __restrict__ out_w) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const double hx = (xedges[1] - xedges[0]) / 2, hy = (yedges[1] - yedges[0]) / 2;  // houghtransform.py:197-198
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_bins; r += warps) {
    const int b = rank[r];
    const double g = add(xedges[b / ybins], hx), c = add(yedges[b % ybins], hy);
    long long run = out_x ? off[r] : 0;
    for (int p0 = 0; p0 < n_pairs; p0 += 32) {
      const int p = p0 + lane;
      bool hit = false;
      double x = 0.0, y = 0.0, pred = 0.0;
      if (p < n_pairs) {
        x = px[p]; y = py[p];
        pred = add(mul(g, x), c);                       // gradient * pairs[:, 0] + intercept (:242)
        hit = fabs(sub(pred, y)) <= tol;                // :244-245
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (out_x && hit) {
        const long long o = run + __popc(m & ((1u << lane) - 1u));
        const double dlt = sub(y, pred);
        out_x[o] = x; out_y[o] = y;
        out_w[o] = exp(-mul(dlt, dlt) / gauss_denom);  // util.gauss(actual, 1., predicted, sigma) (:250-257)
      }
      run += __popc(m);
    }
    if (!out_x && lane == 0) off[r + 1] = run;  // counts; the caller turns them into offsets
  }
}

}  // namespace rb2

extern "C" int rb2_candidate_points(const rb2_hough_desc* h_hough, double tol, double gauss_denom, int64_t* d_off,
                                    double* d_x, double* d_y, double* d_w, void* stream_) {
  using namespace rb2;
  if (!h_hough || !d_off || !h_hough->d_xedges || !h_hough->d_yedges || !h_hough->d_rank) return RB2_ERR_ARG;
  if ((d_x != nullptr) != (d_y != nullptr) || (d_x != nullptr) != (d_w != nullptr)) return RB2_ERR_ARG;
  const rb2_hough_desc& h = *h_hough;
  const long long B = (long long)h.xbins * h.ybins;
  if (B <= 0 || B > 0x7fffffffll || h.n_pairs < 0 || (h.n_pairs && (!h.d_pair_x || !h.d_pair_y))) return RB2_ERR_ARG;
  int sms = 148;
  if (rb2_sm_count(&sms) != RB2_OK) return RB2_ERR_NO_DEVICE;
  const int blocks = (int)min((long long)sms * 8, (B + 7) / 8);
  candidate_points_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(h.d_xedges, h.d_yedges, h.ybins, h.d_rank, (int)B,
                                                                     h.d_pair_x, h.d_pair_y, h.n_pairs, tol, gauss_denom,
                                                                     (long long*)d_off, d_x, d_y, d_w);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

