// K1 — Hough accumulation for sm_100a.
//
// Replaces HoughTransform.generate_hough_points (houghtransform.py:60-92) and
// HoughTransform.bin_hough_points (houghtransform.py:167-210) of the reference.
//
// The reference materialises all S*P (slope, pair) intercepts, masks them and
// histograms the survivors.  Here nothing is materialised:
//
//   pass 1  hough_interval_kernel   per pair, the surviving slopes form ONE
//           contiguous index interval [lo, hi] because fl(y - fl(s*x)) is
//           monotone in the slope index; both ends are found by binary search
//           with the reference's own arithmetic, so the decision is bit-exact.
//           The data-dependent histogram range (np.histogram2d uses min/max of
//           the survivors) falls out of the interval ends.
//   pass 2  hough_edges_kernel      np.linspace edges, bit-exact; zero hist.
//   pass 3  hough_bin_kernel        each CTA owns a contiguous slope range =
//           a few whole gradient rows -> a private shared-memory sub-histogram;
//           a thread walks one pair's surviving slopes, its intercept bin moves
//           monotonically, and runs of equal bins are flushed with ONE shared
//           atomicAdd(count) (thread-level run-length aggregation).  Work is
//           N survivors (~15 % of S*P), not S*P.
//   pass 4  hough_rank_kernel       canonical ranking of ALL bins (count
//           descending, ties by descending flat index) = stable LSD radix sort.
//
// Lazy: hough_points materialisation in the reference's exact row order.
#include "common.cuh"
#include <stdlib.h>

namespace rb2 {

struct SlopeGen {
  Linspace ls;
  __device__ SlopeGen(const rb2_hough_desc& d) : ls(d.min_slope, d.max_slope, d.n_slopes) {}
  __device__ __forceinline__ double at(int i) const { return ls.at(i); }
};

__global__ void hough_init_stats_kernel(rb2_hough_stats* stats, int n, bool points = false) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rb2_hough_stats s;
  s.icpt_min = CUDART_INF;
  s.icpt_max = -CUDART_INF;
  s.grad_min = points ? CUDART_INF : 0.0;   // explicit points: min / max are reduced over the data
  s.grad_max = points ? -CUDART_INF : 0.0;
  s.n_points = 0;
  s.slope_lo = 0x7fffffff;
  s.slope_hi = -1;
  s.max_count = 0;
  s.status = 0;
  stats[i] = s;
}

// ---- pass 1 ------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hough_interval_kernel(const rb2_hough_desc* __restrict__ descs, rb2_hough_stats* __restrict__ stats) {
  const rb2_hough_desc d = descs[blockIdx.y];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int S = d.n_slopes;
  SlopeGen sg(d);

  int lo = 1, hi = 0;  // empty
  double cmin = CUDART_INF, cmax = -CUDART_INF;
  if (p < d.n_pairs) {
    const double x = d.d_pair_x[p];
    const double y = d.d_pair_y[p];
    const double mn = d.min_intercept, mx = d.max_intercept;
    auto icpt = [&](int i) { return sub(y, mul(sg.at(i), x)); };
    // c(i) is non-increasing in i for x >= 0, non-decreasing for x < 0.
    // In-range set {i : mn <= c(i) <= mx} is an interval [lo, hi].
    const bool dec = !(x < 0.0);
    // lo = first i with "entered": dec ? c <= mx : c >= mn
    int a = 0, b = S;  // search in [0, S]
    while (a < b) {
      int m = (a + b) >> 1;
      double c = icpt(m);
      bool entered = dec ? (c <= mx) : (c >= mn);
      if (entered) b = m; else a = m + 1;
    }
    lo = a;
    // hi = last i with "not left": dec ? c >= mn : c <= mx
    a = lo - 1; b = S - 1;  // search last true in [lo, S-1]; a = lo-1 means none
    while (a < b) {
      int m = (a + b + 1) >> 1;
      double c = icpt(m);
      bool inside = dec ? (c >= mn) : (c <= mx);
      if (inside) a = m; else b = m - 1;
    }
    hi = a;
    if (lo <= hi) {
      double c0 = icpt(lo), c1 = icpt(hi);
      // NaN inputs never survive the reference's mask
      if (c0 >= mn && c0 <= mx && c1 >= mn && c1 <= mx) {
        cmax = fmax(c0, c1);
        cmin = fmin(c0, c1);
      } else { lo = 1; hi = 0; }
    }
    if (lo > hi) { lo = 1; hi = 0; }
    d.d_lo[p] = lo;
    d.d_hi[p] = hi;
  }
  // block reduction
  long long cnt = (lo <= hi) ? (long long)(hi - lo + 1) : 0;
  int slo = (lo <= hi) ? lo : 0x7fffffff;
  int shi = (lo <= hi) ? hi : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    slo = min(slo, __shfl_xor_sync(0xffffffffu, slo, o));
    shi = max(shi, __shfl_xor_sync(0xffffffffu, shi, o));
    cmin = fmin(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
    cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
  }
  if (lane_id() == 0 && cnt > 0) {
    rb2_hough_stats* s = stats + blockIdx.y;
    atomicAdd(reinterpret_cast<unsigned long long*>(&s->n_points), (unsigned long long)cnt);
    atomicMin(&s->slope_lo, slo);
    atomicMax(&s->slope_hi, shi);
    atomic_min_double(&s->icpt_min, cmin);
    atomic_max_double(&s->icpt_max, cmax);
  }
}

// ---- pass 2 ------------------------------------------------------------------
// grid (zero_blocks, n_problems): every block helps zeroing the histogram,
// block x==0 also writes the edges.
__global__ void __launch_bounds__(256)
hough_edges_kernel(const rb2_hough_desc* __restrict__ descs, rb2_hough_stats* __restrict__ stats) {
  const rb2_hough_desc d = descs[blockIdx.y];
  const int B = d.xbins * d.ybins;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) d.d_hist[i] = 0u;
  if (blockIdx.x != 0) return;
  rb2_hough_stats* s = stats + blockIdx.y;
  SlopeGen sg(d);
  double g0, g1, c0, c1;
  if (s->n_points == 0) {  // numpy: empty sample -> range (0, 1)
    g0 = 0.0; g1 = 1.0; c0 = 0.0; c1 = 1.0;
  } else {
    g0 = sg.at(s->slope_lo);
    g1 = sg.at(s->slope_hi);
    c0 = s->icpt_min;
    c1 = s->icpt_max;
    if (g0 == g1) { g0 -= 0.5; g1 += 0.5; }  // _get_outer_edges
    if (c0 == c1) { c0 -= 0.5; c1 += 0.5; }
  }
  Linspace lx(g0, g1, d.xbins + 1), ly(c0, c1, d.ybins + 1);
  for (int i = threadIdx.x; i <= d.xbins; i += blockDim.x) d.d_xedges[i] = lx.at(i);
  for (int i = threadIdx.x; i <= d.ybins; i += blockDim.x) d.d_yedges[i] = ly.at(i);
  if (threadIdx.x == 0) {
    s->grad_min = (s->n_points == 0) ? 0.0 : sg.at(s->slope_lo);
    s->grad_max = (s->n_points == 0) ? 0.0 : sg.at(s->slope_hi);
  }
}

// ---- pass 3 ------------------------------------------------------------------
// grid (n_chunks, pair_splits, n_problems), dynamic smem:
//   double ye[ybins+1]; int row_s[spc]; int row_last[xbins]; uint32 sub[cap]
// A thread walks the surviving slopes of one pair.  The inner loop touches shared memory only
// when something changes: the slope comes from np.linspace's own arithmetic in registers, the
// gradient row from a per-row "last slope" table (one load per ~S/xbins slopes), the intercept
// bin from its two edges kept in registers (reloaded when c leaves [e_lo, e_hi)); equal-bin runs
// are flushed with one shared atomicAdd(count).
__global__ void __launch_bounds__(512)
hough_bin_kernel(const rb2_hough_desc* __restrict__ descs, const rb2_hough_stats* __restrict__ stats,
                 int spc, int sub_cap_words, int crossing, int pairs_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const rb2_hough_desc d = descs[blockIdx.z];
  const rb2_hough_stats st = stats[blockIdx.z];
  if (st.n_points == 0) return;
  const int S = d.n_slopes, yb = d.ybins, xb = d.xbins;
  int s0 = blockIdx.x * spc;
  int s1 = min(S, s0 + spc);
  s0 = max(s0, st.slope_lo);
  s1 = min(s1, st.slope_hi + 1);
  if (s0 >= s1) return;
  // contiguous pair range of this CTA
  const int np = d.n_pairs;
  // pairs_per_cta > 0: fixed-size pair ranges (a batch of unequal problems: CTAs of equal work, the
  // surplus CTAs of a small problem leave at once); 0: this problem's pairs in gridDim.y equal parts
  const int per = pairs_per_cta > 0 ? pairs_per_cta : (np + gridDim.y - 1) / gridDim.y;
  const int p0 = blockIdx.y * per, p1 = min(np, p0 + per);
  if (p0 >= p1) return;

  double* ye = reinterpret_cast<double*>(smem_raw);
  int* row_s = reinterpret_cast<int*>(ye + (yb + 1));
  int* row_last = row_s + spc;
  unsigned* subh = reinterpret_cast<unsigned*>(row_last + xb);

  for (int i = threadIdx.x; i <= yb; i += blockDim.x) ye[i] = d.d_yedges[i];
  const SlopeGen sg(d);
  const bool lin_fast = (S > 1) && !sg.ls.step_zero;  // np.linspace's common case: i * step + start, last = stop
  const double* __restrict__ xe = d.d_xedges;
  const double inv_wx = (double)xb / (xe[xb] - xe[0]);
  const int ns = s1 - s0;
  for (int k = threadIdx.x; k < ns; k += blockDim.x) {
    const double g = sg.at(s0 + k);
    row_s[k] = fix_bin(xe, xb, g, (int)((g - xe[0]) * inv_wx));
  }
  __syncthreads();
  for (int k = threadIdx.x; k < ns; k += blockDim.x)
    if (k == ns - 1 || row_s[k + 1] != row_s[k]) row_last[row_s[k]] = s0 + k;  // rows are non-decreasing in the slope
  __syncthreads();
  const int row_lo = row_s[0], row_hi = row_s[ns - 1];
  const int rows_cap = max(1, sub_cap_words / yb);
  const double ye0 = ye[0];
  const double inv_wy = (double)yb / (ye[yb] - ye0);


  for (int w0 = row_lo; w0 <= row_hi; w0 += rows_cap) {
    const int w1 = min(row_hi, w0 + rows_cap - 1);
    const int words = (w1 - w0 + 1) * yb;
    for (int i = threadIdx.x; i < words; i += blockDim.x) subh[i] = 0u;
    __syncthreads();
    if (crossing && lin_fast) {
      // CROSSING MODE (many slopes per intercept bin): for a fixed pair the intercept is monotone in
      // the slope index, so inside a gradient row the slopes that fall into one intercept bin are ONE
      // run; its end is where c(i) crosses the bin edge - solved arithmetically (one division) and fixed
      // up with the reference's own arithmetic at the neighbouring indices - instead of walking the
      // run slope by slope: one step and one atomicAdd(count) per visited bin.
      const double inv_step = 1.0 / sg.ls.step;
      auto cval = [&](int i, double x, double y) {
        const double g = (i == S - 1) ? sg.ls.stop : add(mul((double)i, sg.ls.step), sg.ls.start);
        return sub(y, mul(g, x));
      };
      for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
        const int a = max(d.d_lo[p], s0), b = min(d.d_hi[p], s1 - 1);
        if (a > b) continue;
        const double x = d.d_pair_x[p], y = d.d_pair_y[p];
        const bool dec = x > 0.0;
        int i = a;
        while (i <= b) {
          const int row = row_s[i - s0];
          const int iend = min(b, row_last[row]);
          if (row < w0 || row > w1) { i = iend + 1; continue; }
          const int key_row = (row - w0) * yb;
          double c = cval(i, x, y);
          int k = fix_bin(ye, yb, c, (int)((c - ye0) * inv_wy));
          while (i <= iend) {
            int inext = iend + 1;
            if (x != 0.0 && !(dec ? k == 0 : k == yb - 1)) {
              const double e = dec ? ye[k] : ye[k + 1];  // the edge c crosses when it leaves bin k
              const double gi = (((y - e) / x) - sg.ls.start) * inv_step;
              int j = (gi < (double)iend) ? (int)fmax(gi, (double)i) + 1 : iend + 1;
              j = max(i + 1, min(iend + 1, j));
              // left(j): c(j) is past the edge (dec: c < e; inc: c >= e), monotone in j
              while (j > i + 1) {
                const double cj = cval(j - 1, x, y);
                if (!(dec ? (cj < e) : !(cj < e))) break;
                --j;
              }
              while (j <= iend) {
                const double cj = cval(j, x, y);
                if (dec ? (cj < e) : !(cj < e)) break;
                ++j;
              }
              inext = j;
            }
            atomicAdd(&subh[key_row + k], (unsigned)(inext - i));
            i = inext;
            if (i <= iend) {
              c = cval(i, x, y);
              k = fix_bin(ye, yb, c, dec ? k - 1 : k + 1);
            }
          }
        }
      }
    } else
    // A warp takes 32 consecutive pairs (one peak, neighbouring atlas lines: heavily overlapping
    // slope intervals) and walks the slope index in LOCKSTEP: slope, gradient row and row boundary
    // are warp-uniform, a lane is active while the index is inside its pair's interval.
    {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int pb = p0 + warp * 32; pb < p1; pb += n_warps * 32) {
      const int p = pb + lane;
      int a = 1, b = 0;
      double x = 0.0, y = 0.0;
      if (p < p1) {
        a = max(d.d_lo[p], s0); b = min(d.d_hi[p], s1 - 1);
        x = d.d_pair_x[p]; y = d.d_pair_y[p];
      }
      const bool any = a <= b;
      const int ia = __reduce_min_sync(0xffffffffu, any ? a : 0x7fffffff);
      const int ib = __reduce_max_sync(0xffffffffu, any ? b : -1);
      if (ia > ib) continue;
      int cur = -1;
      unsigned run = 0;
      int k = -1;
      double e_lo = 0.0, e_hi = 0.0;
      int i = ia;
      while (i <= ib) {
        const int row = row_s[i - s0];
        const int iend = min(ib, row_last[row]);
        if (row < w0 || row > w1) { i = iend + 1; continue; }
        const int key_row = (row - w0) * yb;
        for (; i <= iend; ++i) {
          const double g = lin_fast ? ((i == S - 1) ? sg.ls.stop : add(mul((double)i, sg.ls.step), sg.ls.start)) : sg.at(i);
          if (i >= a && i <= b) {
            const double c = sub(y, mul(g, x));
            if (k < 0) {
              k = fix_bin(ye, yb, c, (int)((c - ye0) * inv_wy));
              e_lo = ye[k]; e_hi = ye[k + 1];
            } else if (c < e_lo && k > 0) {  // one bin down (the common move for x > 0), more only for steep pairs
              --k; e_hi = e_lo; e_lo = ye[k];
              if (c < e_lo) { k = fix_bin(ye, yb, c, k); e_lo = ye[k]; e_hi = ye[k + 1]; }
            } else if (!(c < e_hi) && k < yb - 1) {
              ++k; e_lo = e_hi; e_hi = ye[k + 1];
              if (!(c < e_hi) && k < yb - 1) { k = fix_bin(ye, yb, c, k); e_lo = ye[k]; e_hi = ye[k + 1]; }
            }
            const int key = key_row + k;
            if (key == cur) {
              ++run;
            } else {
              if (run) atomicAdd(&subh[cur], run);
              cur = key;
              run = 1;
            }
          }
        }
      }
      if (run) atomicAdd(&subh[cur], run);
    }
    }
    __syncthreads();
    unsigned* gh = d.d_hist + (size_t)w0 * yb;
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
      unsigned v = subh[i];
      if (v) atomicAdd(&gh[i], v);
    }
    __syncthreads();
  }
}

// ---- pass 4: canonical ranking ----------------------------------------------
// One CTA (1024 threads) per problem.  Stable LSD radix sort (8-bit digits) of
// the bins visited in DESCENDING flat index, by DESCENDING count:
//   == np.argsort(hist.ravel(), kind='stable')[::-1].
// The sort is a chain of short dependent steps (per tile of 1024 bins: digit of hist[src[e]], rank inside the
// tile, scatter), i.e. latency-bound on its loads: when the histogram and both index buffers fit in shared memory
// (12 bytes per bin: every default shape, 100 x 100 ... 125 x 125) the whole sort runs there - 224 -> ~40 us for one
// 100 x 100 spectrum, which is fixed cost of every single-spectrum fit; finer binnings sort in global memory.
constexpr int RANK_SMEM_MAX_BINS = 16384;
__global__ void __launch_bounds__(1024)
hough_rank_kernel(const rb2_hough_desc* __restrict__ descs, rb2_hough_stats* __restrict__ stats, int smem_bins) {
  extern __shared__ __align__(16) unsigned char rank_smem[];
  __shared__ unsigned base[256];
  __shared__ unsigned tile_tot[256];
  __shared__ __align__(16) unsigned short wcnt[32][256];
  __shared__ unsigned red[32];
  const rb2_hough_desc d = descs[blockIdx.x];
  const int B = d.xbins * d.ybins;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool in_smem = B <= smem_bins;
  const unsigned* hist = d.d_hist;
  int* buf_a = d.d_rank;
  int* buf_b = d.d_sort_tmp;
  if (in_smem) {
    unsigned* sh = reinterpret_cast<unsigned*>(rank_smem);
    for (int i = tid; i < B; i += 1024) sh[i] = d.d_hist[i];
    hist = sh;
    buf_a = reinterpret_cast<int*>(sh + smem_bins);
    buf_b = buf_a + smem_bins;
    __syncthreads();
  }

  // max count -> number of 8-bit passes
  unsigned mx = 0;
  for (int i = tid; i < B; i += 1024) mx = max(mx, hist[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    mx = red[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) { red[0] = mx; stats[blockIdx.x].max_count = mx; }
  }
  __syncthreads();
  mx = red[0];
  int passes = 0;
  while (passes < 4 && (mx >> (8 * passes)) != 0) ++passes;

  if (passes == 0) {
    for (int e = tid; e < B; e += 1024) buf_a[e] = B - 1 - e;
  }
  const int* src = nullptr;  // pass 0 reads the implicit reversed identity
  for (int ps = 0; ps < passes; ++ps) {
    const int sh = 8 * ps;
    int* dst = (((passes - 1 - ps) & 1) == 0) ? buf_a : buf_b;  // the last pass lands in buf_a
    if (tid < 256) base[tid] = 0;
    __syncthreads();
    // digit histogram: most bins share a few digits (in the high-byte pass nearly all of them share ONE), so a
    // plain shared atomic per bin serialises on one address; lanes with equal digits are merged first
    for (int e0 = 0; e0 < B; e0 += 1024) {
      const int e = e0 + tid;
      const bool valid = e < B;
      unsigned dg = 0;
      if (valid) {
        const int idx = src ? src[e] : (B - 1 - e);
        dg = 255u - ((hist[idx] >> sh) & 255u);
      }
      const unsigned act = __ballot_sync(0xffffffffu, valid);
      if (valid) {
        const unsigned peers = __match_any_sync(act, dg);
        if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(&base[dg], (unsigned)__popc(peers));
      }
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 256 digit counters
      unsigned v[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[j] = base[lane * 8 + j]; sum += v[j]; }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      unsigned run = incl - sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) { base[lane * 8 + j] = run; run += v[j]; }
    }
    __syncthreads();
    for (int t0 = 0; t0 < B; t0 += 1024) {
      reinterpret_cast<uint4*>(&wcnt[0][0])[tid] = make_uint4(0, 0, 0, 0);  // 16 KB / 1024 threads
      __syncthreads();
      const int e = t0 + tid;
      const bool valid = e < B;
      int idx = 0;
      unsigned dg = 0;
      if (valid) {
        idx = src ? src[e] : (B - 1 - e);
        dg = 255u - ((hist[idx] >> sh) & 255u);
      }
      const unsigned act = __ballot_sync(0xffffffffu, valid);
      unsigned rank_in_warp = 0;
      if (valid) {
        const unsigned peers = __match_any_sync(act, dg);
        rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (rank_in_warp == 0) wcnt[warp][dg] = (unsigned short)__popc(peers);
      }
      __syncthreads();
      // exclusive prefix over the 32 warps, per digit; warp w owns digits 8w..8w+7
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int dgt = warp * 8 + j;
        const unsigned v = wcnt[lane][dgt];
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        wcnt[lane][dgt] = (unsigned short)(incl - v);
        if (lane == 31) tile_tot[dgt] = incl;
      }
      __syncthreads();
      if (valid) dst[base[dg] + wcnt[warp][dg] + rank_in_warp] = idx;
      __syncthreads();
      if (tid < 256) base[tid] += tile_tot[tid];
      // next iteration's zeroing of wcnt + its barrier orders this update
    }
    __syncthreads();
    src = dst;
  }
  __syncthreads();
  for (int r = tid; r < B; r += 1024) {
    const int e = buf_a[r];
    if (in_smem) d.d_rank[r] = e;
    d.d_rankpos[e] = r;
  }
}

// ---- lazy materialisation of hough_points -----------------------------------
__global__ void hough_slope_count_kernel(const rb2_hough_desc d, long long* __restrict__ off) {
  // off[i+1] accumulates (#pairs with lo<=i) - (#pairs with hi<i) via a difference array
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.n_pairs) return;
  int lo = d.d_lo[p], hi = d.d_hi[p];
  if (lo > hi) return;
  atomicAdd(reinterpret_cast<unsigned long long*>(&off[lo + 1]), 1ull);
  atomicAdd(reinterpret_cast<unsigned long long*>(&off[hi + 2]), (unsigned long long)(-1ll));
}

__global__ void __launch_bounds__(1024)
hough_slope_scan_kernel(int S, long long* __restrict__ off) {
  // single block: off[i+1] (difference) -> count[i] -> exclusive offsets in off[0..S]
  __shared__ long long carry_cnt, carry_off;
  __shared__ long long wsum[32], wsum2[32];
  if (threadIdx.x == 0) { carry_cnt = 0; carry_off = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = 0; t0 < S; t0 += 1024) {
    int i = t0 + threadIdx.x;
    long long dlt = (i < S) ? off[i + 1] : 0;
    // inclusive scan of differences -> count[i]
    long long v = dlt;
    for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
      long long w = wsum[lane], iv = w;
      for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += t; }
      wsum[lane] = iv - w;
    }
    __syncthreads();
    long long cnt = v + wsum[warp] + carry_cnt;  // count[i]
    // inclusive scan of counts -> offsets
    long long c = (i < S) ? cnt : 0;
    long long u = c;
    for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, u, o); if (lane >= o) u += t; }
    if (lane == 31) wsum2[warp] = u;
    __syncthreads();
    if (warp == 0) {
      long long w = wsum2[lane], iv = w;
      for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += t; }
      wsum2[lane] = iv - w;
    }
    __syncthreads();
    long long incl = u + wsum2[warp] + carry_off;
    __syncthreads();
    if (i < S) off[i + 1] = incl;  // exclusive offset of slope i+1
    if (threadIdx.x == 1023 || i == S - 1) { carry_cnt = cnt; carry_off = incl; }
    __syncthreads();
  }
  if (threadIdx.x == 0) off[0] = 0;
}

// one block per slope (grid-stride); rows ordered by pair index inside a slope
__global__ void __launch_bounds__(256)
hough_points_kernel(const rb2_hough_desc d, const long long* __restrict__ off, double* __restrict__ out) {
  __shared__ int wtot[8];
  __shared__ long long base;
  SlopeGen sg(d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = blockIdx.x; s < d.n_slopes; s += gridDim.x) {
    const long long o0 = off[s], o1 = off[s + 1];
    if (o0 == o1) continue;
    const double g = sg.at(s);
    if (threadIdx.x == 0) base = o0;
    __syncthreads();
    for (int p0 = 0; p0 < d.n_pairs; p0 += 256) {
      const int p = p0 + threadIdx.x;
      bool in = false;
      double c = 0.0;
      if (p < d.n_pairs) {
        in = (d.d_lo[p] <= s) && (s <= d.d_hi[p]);
        if (in) c = sub(d.d_pair_y[p], mul(g, d.d_pair_x[p]));
      }
      unsigned bal = __ballot_sync(0xffffffffu, in);
      if (lane == 0) wtot[warp] = __popc(bal);
      __syncthreads();
      int pre = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) { int t = wtot[w]; if (w < warp) pre += t; tot += t; }
      if (in) {
        long long r = base + pre + __popc(bal & ((1u << lane) - 1u));
        out[2 * r] = g;
        out[2 * r + 1] = c;
      }
      __syncthreads();
      if (threadIdx.x == 0) base += tot;
      __syncthreads();
    }
  }
}

}  // namespace rb2

// ---- host entry points -------------------------------------------------------
namespace rb2 {
// Ranking of a FEW spectra by counting: rank(e) = #{e' : hist[e'] > hist[e]} + #{e' > e : hist[e'] == hist[e]}
// is exactly the canonical order (count descending, ties by descending flat index) and needs no sort - B^2
// comparisons spread over the whole GPU (B = 10^4: 10^8, a few microseconds) instead of a chain of dependent
// radix tiles on one CTA (224 us for one 100 x 100 spectrum, fixed cost of every single-spectrum fit).
constexpr int RANKP_T = 256;
__global__ void __launch_bounds__(RANKP_T)
hough_rank_zero_kernel(const rb2_hough_desc* __restrict__ descs) {
  const rb2_hough_desc d = descs[blockIdx.y];
  const int B = d.xbins * d.ybins;
  for (int e = blockIdx.x * RANKP_T + threadIdx.x; e < B; e += gridDim.x * RANKP_T) d.d_rankpos[e] = 0;
}
__global__ void __launch_bounds__(RANKP_T)
hough_rank_pairs_kernel(const rb2_hough_desc* __restrict__ descs, int src_chunk) {
  __shared__ unsigned s_h[1024];
  const rb2_hough_desc d = descs[blockIdx.z];
  const int B = d.xbins * d.ybins;
  const int e = blockIdx.x * RANKP_T + threadIdx.x;
  if (blockIdx.x * RANKP_T >= B) return;
  const unsigned he = (e < B) ? d.d_hist[e] : 0u;
  const int s0 = blockIdx.y * src_chunk, s1 = min(B, s0 + src_chunk);
  int cnt = 0;
  for (int t0 = s0; t0 < s1; t0 += 1024) {
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += RANKP_T) s_h[i] = (t0 + i < s1) ? d.d_hist[t0 + i] : 0u;
    __syncthreads();
    const int n = min(1024, s1 - t0);
    // sources before e only count when strictly larger, sources after e also when equal
    const int split = max(0, min(n, e + 1 - t0));  // tile entries with index <= e
    for (int i = 0; i < split; ++i) cnt += (s_h[i] > he);
    for (int i = split; i < n; ++i) cnt += (s_h[i] >= he);
  }
  if (e < B && cnt) atomicAdd(&d.d_rankpos[e], cnt);
}
__global__ void __launch_bounds__(RANKP_T)
hough_rank_scatter_kernel(const rb2_hough_desc* __restrict__ descs, rb2_hough_stats* __restrict__ stats) {
  const rb2_hough_desc d = descs[blockIdx.y];
  const int B = d.xbins * d.ybins;
  unsigned mx = 0;
  for (int e = blockIdx.x * RANKP_T + threadIdx.x; e < B; e += gridDim.x * RANKP_T) {
    d.d_rank[d.d_rankpos[e]] = e;
    mx = max(mx, d.d_hist[e]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(&stats[blockIdx.y].max_count, mx);
}
}  // namespace rb2

namespace rb2 {
// Ranking of ONE fine-binned spectrum (config 5: 10^6 bins) - too many bins for the counting kernel above, and the
// single-CTA radix sort walks them tile after tile (26 ms).  The same stable LSD radix sort spread over the GPU:
// per 8-bit pass (a) every CTA histograms the digits of its tile of 1024 inputs, (b) per digit a CTA scans the tile
// counts, a last warp scans the digit totals, (c) every CTA ranks its tile's inputs inside their digit (in input
// order: warp match + prefix over the warps, as in hough_rank_kernel) and scatters them.
constexpr int RANKM_T = 1024;
__global__ void __launch_bounds__(RANKM_T)
rank_multi_hist_kernel(const unsigned* __restrict__ hist, const int* __restrict__ src, int B, int sh,
                       unsigned* __restrict__ tile_hist /* [256][tiles] */, int tiles) {
  __shared__ unsigned cnt[256];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 256) cnt[tid] = 0u;
  __syncthreads();
  const int e = blockIdx.x * RANKM_T + tid;
  const bool valid = e < B;
  unsigned dg = 0;
  if (valid) dg = 255u - ((hist[src ? src[e] : (B - 1 - e)] >> sh) & 255u);
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    const unsigned peers = __match_any_sync(act, dg);
    if ((peers & ((1u << lane) - 1u)) == 0) atomicAdd(&cnt[dg], (unsigned)__popc(peers));
  }
  __syncthreads();
  if (tid < 256) tile_hist[(size_t)tid * tiles + blockIdx.x] = cnt[tid];
}
// one CTA per digit: exclusive scan of that digit's counts over the tiles (in place), total -> digit_total
__global__ void __launch_bounds__(RANKM_T)
rank_multi_scan_kernel(unsigned* __restrict__ tile_hist, int tiles, unsigned* __restrict__ digit_total) {
  __shared__ unsigned wsum[32];
  __shared__ unsigned carry;
  unsigned* row = tile_hist + (size_t)blockIdx.x * tiles;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0u;
  __syncthreads();
  for (int t0 = 0; t0 < tiles; t0 += RANKM_T) {
    const int t = t0 + tid;
    const unsigned v = (t < tiles) ? row[t] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = wsum[lane], iw = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(0xffffffffu, iw, o); if (lane >= o) iw += x; }
      wsum[lane] = iw - w;
    }
    __syncthreads();
    const unsigned excl = carry + wsum[warp] + incl - v;
    if (t < tiles) row[t] = excl;
    __syncthreads();
    if (tid == RANKM_T - 1) carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) digit_total[blockIdx.x] = carry;
}
__global__ void rank_multi_base_kernel(unsigned* __restrict__ digit_total) {  // <<<1, 32>>>: exclusive scan of 256 totals
  const int lane = threadIdx.x;
  unsigned v[8], sum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = digit_total[lane * 8 + j]; sum += v[j]; }
  unsigned incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
  unsigned run = incl - sum;
#pragma unroll
  for (int j = 0; j < 8; ++j) { digit_total[lane * 8 + j] = run; run += v[j]; }
}
__global__ void __launch_bounds__(RANKM_T)
rank_multi_place_kernel(const unsigned* __restrict__ hist, const int* __restrict__ src, int* __restrict__ dst, int B, int sh,
                        const unsigned* __restrict__ tile_off /* [256][tiles], scanned */, int tiles,
                        const unsigned* __restrict__ digit_base) {
  __shared__ __align__(16) unsigned short wcnt[32][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  reinterpret_cast<uint4*>(&wcnt[0][0])[tid] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const int e = blockIdx.x * RANKM_T + tid;
  const bool valid = e < B;
  int idx = 0;
  unsigned dg = 0;
  if (valid) {
    idx = src ? src[e] : (B - 1 - e);
    dg = 255u - ((hist[idx] >> sh) & 255u);
  }
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  unsigned rank_in_warp = 0;
  if (valid) {
    const unsigned peers = __match_any_sync(act, dg);
    rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (rank_in_warp == 0) wcnt[warp][dg] = (unsigned short)__popc(peers);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {  // exclusive prefix over the 32 warps, per digit; warp w owns digits 8w..8w+7
    const int dgt = warp * 8 + j;
    const unsigned v = wcnt[lane][dgt];
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
    wcnt[lane][dgt] = (unsigned short)(incl - v);
  }
  __syncthreads();
  if (valid) dst[digit_base[dg] + tile_off[(size_t)dg * tiles + blockIdx.x] + wcnt[warp][dg] + rank_in_warp] = idx;
}
__global__ void rank_multi_finish_kernel(const int* __restrict__ sorted, int* __restrict__ rank, int* __restrict__ rankpos, int B) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x) {
    const int e = sorted[r];
    if (rank != sorted) rank[r] = e;
    rankpos[e] = r;
  }
}
__global__ void __launch_bounds__(256)
rank_multi_max_kernel(const unsigned* __restrict__ hist, int B, rb2_hough_stats* __restrict__ stats) {
  unsigned mx = 0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B; e += gridDim.x * blockDim.x) mx = max(mx, hist[e]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(&stats->max_count, mx);
}
}  // namespace rb2

// multi-CTA radix ranking of one spectrum; h.d_sort_tmp holds B ints, `scratch` (256 * tiles + 256 unsigned) comes
// from the caller.  Needs the maximum count on the host to know the number of passes: one small D2H inside.
static cudaError_t launch_rank_multi(const rb2_hough_desc& h, rb2_hough_stats* d_stats, unsigned* scratch, cudaStream_t stream) {
  using namespace rb2;
  const int B = h.xbins * h.ybins;
  const int tiles = (B + RANKM_T - 1) / RANKM_T;
  rank_multi_max_kernel<<<256, 256, 0, stream>>>(h.d_hist, B, d_stats);
  unsigned mx = 0;
  cudaError_t e = cudaMemcpyAsync(&mx, &d_stats->max_count, sizeof(unsigned), cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return e;
  int passes = 0;
  while (passes < 4 && (mx >> (8 * passes)) != 0) ++passes;
  unsigned* tile_hist = scratch;
  unsigned* digit_total = scratch + (size_t)256 * tiles;
  const int* src = nullptr;
  int* bufs[2] = {h.d_rank, h.d_sort_tmp};
  for (int ps = 0; ps < passes; ++ps) {
    int* dst = bufs[(passes - 1 - ps) & 1];  // the last pass lands in d_rank
    rank_multi_hist_kernel<<<tiles, RANKM_T, 0, stream>>>(h.d_hist, src, B, 8 * ps, tile_hist, tiles);
    rank_multi_scan_kernel<<<256, RANKM_T, 0, stream>>>(tile_hist, tiles, digit_total);
    rank_multi_base_kernel<<<1, 32, 0, stream>>>(digit_total);
    rank_multi_place_kernel<<<tiles, RANKM_T, 0, stream>>>(h.d_hist, src, dst, B, 8 * ps, tile_hist, tiles, digit_total);
    src = dst;
  }
  if (passes == 0) {  // an empty histogram: descending flat index
    rank_multi_hist_kernel<<<tiles, RANKM_T, 0, stream>>>(h.d_hist, nullptr, B, 0, tile_hist, tiles);
    rank_multi_scan_kernel<<<256, RANKM_T, 0, stream>>>(tile_hist, tiles, digit_total);
    rank_multi_base_kernel<<<1, 32, 0, stream>>>(digit_total);
    rank_multi_place_kernel<<<tiles, RANKM_T, 0, stream>>>(h.d_hist, nullptr, h.d_rank, B, 0, tile_hist, tiles, digit_total);
  }
  rank_multi_finish_kernel<<<256, 256, 0, stream>>>(h.d_rank, h.d_rank, h.d_rankpos, B);
  return cudaGetLastError();
}

static cudaError_t launch_rank(int n_problems, const rb2_hough_desc* d_desc, const rb2_hough_desc* h_desc,
                               rb2_hough_stats* d_stats, cudaStream_t stream) {
  using namespace rb2;
  int max_bins = 0;
  double pair_work = 0.0;
  for (int i = 0; i < n_problems; ++i) {
    const int b = h_desc[i].xbins * h_desc[i].ybins;
    max_bins = max(max_bins, b);
    pair_work += (double)b * (double)b;
  }
  if (pair_work <= 4e9 && max_bins > 0) {  // a few spectra: rank by counting on the whole GPU
    const int tiles = (max_bins + RANKP_T - 1) / RANKP_T;
    int sms = 148;
    rb2_sm_count(&sms);
    int split = max(1, min(64, (4 * sms) / max(1, tiles * n_problems)));  // source chunks per target tile: fill the GPU
    int src_chunk = (max_bins + split - 1) / split;
    src_chunk = (src_chunk + 1023) / 1024 * 1024;
    split = (max_bins + src_chunk - 1) / src_chunk;
    hough_rank_zero_kernel<<<dim3(min(tiles, 64), n_problems), RANKP_T, 0, stream>>>(d_desc);
    hough_rank_pairs_kernel<<<dim3(tiles, split, n_problems), RANKP_T, 0, stream>>>(d_desc, src_chunk);
    hough_rank_scatter_kernel<<<dim3(min(tiles, 64), n_problems), RANKP_T, 0, stream>>>(d_desc, d_stats);
    return cudaGetLastError();
  }
  if (n_problems == 1 && max_bins > 65536) {  // one fine-binned spectrum: radix passes over the whole GPU
    static unsigned* scratch[64] = {};
    static size_t scratch_n[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    const size_t need = (size_t)256 * ((max_bins + RANKM_T - 1) / RANKM_T) + 256;
    if (need > scratch_n[dev]) {
      if (scratch[dev]) cudaFree(scratch[dev]);
      e = cudaMalloc(&scratch[dev], need * sizeof(unsigned));
      if (e != cudaSuccess) { scratch[dev] = nullptr; scratch_n[dev] = 0; return e; }
      scratch_n[dev] = need;
    }
    return launch_rank_multi(h_desc[0], d_stats, scratch[dev], stream);
  }
  const int smem_bins = (max_bins <= RANK_SMEM_MAX_BINS) ? max_bins : 0;  // 0: sort in global memory
  const size_t smem = (size_t)smem_bins * 12;
  static size_t configured[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  dev = (dev >= 0 && dev < 64) ? dev : 0;
  if (smem > configured[dev]) {
    e = cudaFuncSetAttribute(hough_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev] = smem;
  }
  hough_rank_kernel<<<n_problems, 1024, smem, stream>>>(d_desc, d_stats, smem_bins);
  return cudaGetLastError();
}

extern "C" int rb2_hough_accumulate(int n_problems, const rb2_hough_desc* d_desc, const rb2_hough_desc* h_desc,
                                    rb2_hough_stats* d_stats, int slopes_per_cta, void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_desc || !h_desc || !d_stats) return RB2_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  int sms = 148;
  if (rb2_sm_count(&sms) != RB2_OK) return RB2_ERR_NO_DEVICE;

  int max_pairs = 0, max_bins = 0, max_slopes = 0, max_yb = 0;
  for (int i = 0; i < n_problems; ++i) {
    const rb2_hough_desc& h = h_desc[i];
    if (h.n_slopes < 1 || h.xbins < 1 || h.ybins < 1 || h.n_pairs < 0) return RB2_ERR_ARG;
    if (!(h.min_slope <= h.max_slope)) return RB2_ERR_ARG;
    max_pairs = max(max_pairs, h.n_pairs);
    max_bins = max(max_bins, h.xbins * h.ybins);
    max_slopes = max(max_slopes, h.n_slopes);
    max_yb = max(max_yb, h.ybins);
  }
  hough_init_stats_kernel<<<(n_problems + 255) / 256, 256, 0, stream>>>(d_stats, n_problems);
  if (max_pairs > 0) {
    dim3 g1((max_pairs + 255) / 256, n_problems);
    hough_interval_kernel<<<g1, 256, 0, stream>>>(d_desc, d_stats);
  }
  {
    int zb = max(1, min(64, (max_bins + 256 * 16 - 1) / (256 * 16)));
    dim3 g2(zb, n_problems);
    hough_edges_kernel<<<g2, 256, 0, stream>>>(d_desc, d_stats);
  }
  if (max_pairs > 0) {
    // chunking: aim at >= 4 CTAs per SM over the whole batch
    int pair_splits = 1;
    int target_chunks = (4 * sms + n_problems - 1) / n_problems;
    if (target_chunks > max_slopes) {
      pair_splits = min(8, max(1, target_chunks / max_slopes));
      pair_splits = min(pair_splits, max(1, max_pairs / 2048));
      target_chunks = max_slopes;
    }
    // batches: problems differ a lot in size (C4: 2 k ... 38 k pairs), and a grid of (slope chunks) x
    // (problems) ends on its few largest CTAs; cut every problem into pair ranges of equal size
    // instead and use only as many slope chunks as it then takes to fill the GPU
    int pairs_per_cta = 0;
    {
      static int ppc_env = -1;
      if (ppc_env < 0) { const char* e = getenv("RB2_HOUGH_PAIRS_PER_CTA"); ppc_env = e ? max(0, atoi(e)) : 4096; }
      if (n_problems > 1 && pair_splits == 1 && ppc_env > 0 && max_pairs > ppc_env) {
        pairs_per_cta = ppc_env;
        long long units = 0;
        for (int i = 0; i < n_problems; ++i) units += (h_desc[i].n_pairs + pairs_per_cta - 1) / pairs_per_cta;
        pair_splits = (max_pairs + pairs_per_cta - 1) / pairs_per_cta;
        target_chunks = (int)max(1LL, min((long long)max_slopes, (4LL * sms + units - 1) / max(1LL, units)));
      }
    }
    int spc = slopes_per_cta > 0 ? slopes_per_cta : (max_slopes + target_chunks - 1) / target_chunks;
    spc = max(1, spc);
    int n_chunks = (max_slopes + spc - 1) / spc;
    // shared memory: edges + slope tables + sub-histogram (as large as useful, <= 160 KB)
    int max_xb = 1;
    for (int i = 0; i < n_problems; ++i) max_xb = max(max_xb, h_desc[i].xbins);
    size_t fixed = (size_t)(max_yb + 1) * 8 + (size_t)spc * 4 + (size_t)max_xb * 4 + 8;
    size_t want_words = 0;
    for (int i = 0; i < n_problems; ++i) {
      const rb2_hough_desc& h = h_desc[i];
      // rows a chunk can span: spc slopes over >= 1 surviving slope per row, at most xbins
      size_t rows = (size_t)min(h.xbins, spc);
      want_words = max(want_words, rows * (size_t)h.ybins);
    }
    size_t cap_bytes = 160 * 1024;
    size_t sub_words = min(want_words, (cap_bytes - min(cap_bytes, fixed)) / 4);
    if (sub_words < (size_t)max_yb) return RB2_ERR_ARG;  // ybins too large for one row in smem
    size_t smem = fixed + sub_words * 4;
    RB2_CHECK(cudaFuncSetAttribute(hough_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 g3(n_chunks, pair_splits, n_problems);
    // crossing mode pays when a pair's ~0.2 S surviving slopes sweep the ybins intercept bins several
    // slopes at a time; finer binnings keep the slope-by-slope walk
    static const int force = getenv("RB2_HOUGH_CROSSING") ? atoi(getenv("RB2_HOUGH_CROSSING")) : -1;
    const int crossing = force >= 0 ? force : (max_slopes >= 8 * max_yb);
    hough_bin_kernel<<<g3, 512, smem, stream>>>(d_desc, d_stats, spc, (int)sub_words, crossing, pairs_per_cta);
  }
  RB2_CHECK(launch_rank(n_problems, d_desc, h_desc, d_stats, stream));
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}


// =====================================================================================
// Explicit Hough points: HoughTransform.generate_hough_points_brute_force
// (houghtransform.py:94-140) and the binning of externally supplied points
// (add_hough_points :142-165, load :291-339 followed by bin_hough_points :167-210).
// =====================================================================================
namespace rb2 {

struct BfLimits { double min_slope, max_slope, min_icpt, max_icpt; };

// gradient / intercept of the line through sorted points i < j, the reference's arithmetic (:121-122)
__device__ __forceinline__ bool bf_line(const double* __restrict__ x, const double* __restrict__ y, int i, int j,
                                        const BfLimits& L, double& g, double& c) {
  g = (y[j] - y[i]) / (x[j] - x[i]);
  c = sub(y[j], mul(g, x[j]));
  return (g > L.min_slope) && (g < L.max_slope) && (c > L.min_icpt) && (c < L.max_icpt);  // strict (:130-135)
}

// one CTA per row i: ordered compaction of the valid j > i; points == NULL -> count only
__global__ void __launch_bounds__(256)
bf_rows_kernel(const double* __restrict__ x, const double* __restrict__ y, int n, BfLimits L,
               long long* __restrict__ row_cnt, const long long* __restrict__ row_off, double* __restrict__ points) {
  __shared__ int wtot[8];
  __shared__ long long base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = blockIdx.x; i < n - 1; i += gridDim.x) {
    if (threadIdx.x == 0) base_s = points ? row_off[i] : 0;
    __syncthreads();
    for (int j0 = i + 1; j0 < n; j0 += 256) {
      const int j = j0 + threadIdx.x;
      double g = 0.0, c = 0.0;
      const bool ok = (j < n) && bf_line(x, y, i, j, L, g, c);
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (lane == 0) wtot[warp] = __popc(m);
      __syncthreads();
      int pre = 0, tot = 0;
      for (int w = 0; w < 8; ++w) { if (w < warp) pre += wtot[w]; tot += wtot[w]; }
      if (ok && points) {
        const long long slot = base_s + pre + __popc(m & ((1u << lane) - 1u));
        points[2 * slot] = g;
        points[2 * slot + 1] = c;
      }
      __syncthreads();
      if (threadIdx.x == 0) base_s += tot;
      __syncthreads();
    }
    if (threadIdx.x == 0 && !points) row_cnt[i] = base_s;
  }
}

// single block: counts[0..n) -> exclusive offsets off[0..n]
__global__ void __launch_bounds__(1024)
bf_scan_kernel(int n, const long long* __restrict__ cnt, long long* __restrict__ off) {
  __shared__ long long wsum[32];
  __shared__ long long carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n; t0 += 1024) {
    const int i = t0 + threadIdx.x;
    const long long v0 = (i < n) ? cnt[i] : 0;
    long long v = v0;
    for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
      const long long w = wsum[lane];
      long long iv = w;
      for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += t; }
      wsum[lane] = iv - w;
    }
    __syncthreads();
    const long long incl = v + wsum[warp] + carry;
    if (i < n) off[i] = incl - v0;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n] = carry;
}

// ---- np.histogram2d of explicit points: d_pair_x = gradients, d_pair_y = intercepts -----------
__global__ void __launch_bounds__(256)
points_minmax_kernel(const rb2_hough_desc* __restrict__ descs, rb2_hough_stats* __restrict__ stats) {
  const rb2_hough_desc d = descs[blockIdx.y];
  double g0 = CUDART_INF, g1 = -CUDART_INF, c0 = CUDART_INF, c1 = -CUDART_INF;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n_pairs; i += gridDim.x * blockDim.x) {
    const double g = d.d_pair_x[i], c = d.d_pair_y[i];
    g0 = fmin(g0, g); g1 = fmax(g1, g); c0 = fmin(c0, c); c1 = fmax(c1, c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    g0 = fmin(g0, __shfl_xor_sync(0xffffffffu, g0, o)); g1 = fmax(g1, __shfl_xor_sync(0xffffffffu, g1, o));
    c0 = fmin(c0, __shfl_xor_sync(0xffffffffu, c0, o)); c1 = fmax(c1, __shfl_xor_sync(0xffffffffu, c1, o));
  }
  if (lane_id() == 0 && g0 <= g1) {
    rb2_hough_stats* s = stats + blockIdx.y;
    atomic_min_double(&s->grad_min, g0); atomic_max_double(&s->grad_max, g1);
    atomic_min_double(&s->icpt_min, c0); atomic_max_double(&s->icpt_max, c1);
  }
}

__global__ void __launch_bounds__(256)
points_edges_kernel(const rb2_hough_desc* __restrict__ descs, rb2_hough_stats* __restrict__ stats) {
  const rb2_hough_desc d = descs[blockIdx.y];
  const int B = d.xbins * d.ybins;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) d.d_hist[i] = 0u;
  if (blockIdx.x != 0) return;
  rb2_hough_stats* s = stats + blockIdx.y;
  double g0 = s->grad_min, g1 = s->grad_max, c0 = s->icpt_min, c1 = s->icpt_max;
  if (d.n_pairs == 0) { g0 = 0.0; g1 = 1.0; c0 = 0.0; c1 = 1.0; }  // numpy: empty sample -> range (0, 1)
  if (g0 == g1) { g0 -= 0.5; g1 += 0.5; }  // _get_outer_edges
  if (c0 == c1) { c0 -= 0.5; c1 += 0.5; }
  Linspace lx(g0, g1, d.xbins + 1), ly(c0, c1, d.ybins + 1);
  for (int i = threadIdx.x; i <= d.xbins; i += blockDim.x) d.d_xedges[i] = lx.at(i);
  for (int i = threadIdx.x; i <= d.ybins; i += blockDim.x) d.d_yedges[i] = ly.at(i);
  if (threadIdx.x == 0) s->n_points = d.n_pairs;
}

__global__ void __launch_bounds__(256)
points_bin_kernel(const rb2_hough_desc* __restrict__ descs) {
  const rb2_hough_desc d = descs[blockIdx.y];
  const int xb = d.xbins, yb = d.ybins;
  const double* __restrict__ xe = d.d_xedges;
  const double* __restrict__ ye = d.d_yedges;
  const double inv_wx = (double)xb / (xe[xb] - xe[0]), inv_wy = (double)yb / (ye[yb] - ye[0]);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.n_pairs; i += gridDim.x * blockDim.x) {
    const double g = d.d_pair_x[i], c = d.d_pair_y[i];
    const int kx = fix_bin(xe, xb, g, (int)((g - xe[0]) * inv_wx));
    const int ky = fix_bin(ye, yb, c, (int)((c - ye[0]) * inv_wy));
    atomicAdd(&d.d_hist[(size_t)kx * yb + ky], 1u);
  }
}

}  // namespace rb2

extern "C" int rb2_hough_brute_force(const double* d_x, const double* d_y, int n, double min_slope, double max_slope,
                                     double min_intercept, double max_intercept, int64_t* d_row_cnt,
                                     int64_t* d_row_off, double* d_points, void* stream_) {
  using namespace rb2;
  if (n < 0 || !d_row_cnt || !d_row_off || (n > 0 && (!d_x || !d_y))) return RB2_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  const BfLimits L{min_slope, max_slope, min_intercept, max_intercept};
  int sms = 148;
  if (rb2_sm_count(&sms) != RB2_OK) return RB2_ERR_NO_DEVICE;
  const int rows = max(n - 1, 0);
  const int grid = max(1, min(rows, sms * 8));
  if (!d_points) {  // pass 1: per-row counts and their exclusive scan (total in d_row_off[rows])
    if (rows > 0)
      bf_rows_kernel<<<grid, 256, 0, stream>>>(d_x, d_y, n, L, (long long*)d_row_cnt, nullptr, nullptr);
    bf_scan_kernel<<<1, 1024, 0, stream>>>(rows, (const long long*)d_row_cnt, (long long*)d_row_off);
  } else if (rows > 0) {  // pass 2: the points, in the reference's order (row i, then j ascending)
    bf_rows_kernel<<<grid, 256, 0, stream>>>(d_x, d_y, n, L, (long long*)d_row_cnt, (const long long*)d_row_off, d_points);
  }
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

extern "C" int rb2_hough_bin_points(int n_problems, const rb2_hough_desc* d_desc, const rb2_hough_desc* h_desc,
                                    rb2_hough_stats* d_stats, void* stream_) {
  using namespace rb2;
  if (n_problems <= 0 || !d_desc || !h_desc || !d_stats) return RB2_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  int max_n = 0, max_bins = 0;
  for (int i = 0; i < n_problems; ++i) {
    const rb2_hough_desc& h = h_desc[i];
    if (h.xbins < 1 || h.ybins < 1 || h.n_pairs < 0) return RB2_ERR_ARG;
    if (h.n_pairs > 0 && (!h.d_pair_x || !h.d_pair_y)) return RB2_ERR_ARG;
    max_n = max(max_n, h.n_pairs);
    max_bins = max(max_bins, h.xbins * h.ybins);
  }
  int sms = 148;
  if (rb2_sm_count(&sms) != RB2_OK) return RB2_ERR_NO_DEVICE;
  hough_init_stats_kernel<<<(n_problems + 255) / 256, 256, 0, stream>>>(d_stats, n_problems, true);
  const int gx = max(1, min((max_n + 255) / 256, max(1, 8 * sms / n_problems)));
  if (max_n > 0) points_minmax_kernel<<<dim3(gx, n_problems), 256, 0, stream>>>(d_desc, d_stats);
  const int zb = max(1, min(64, (max_bins + 256 * 16 - 1) / (256 * 16)));
  points_edges_kernel<<<dim3(zb, n_problems), 256, 0, stream>>>(d_desc, d_stats);
  if (max_n > 0) points_bin_kernel<<<dim3(gx, n_problems), 256, 0, stream>>>(d_desc);
  RB2_CHECK(launch_rank(n_problems, d_desc, h_desc, d_stats, stream));
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}

extern "C" int rb2_hough_points(const rb2_hough_desc* h_desc, int64_t* d_slope_off, double* d_points,
                                int64_t n_points, void* stream_) {
  using namespace rb2;
  if (!h_desc || !d_slope_off) return RB2_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  const rb2_hough_desc d = *h_desc;
  RB2_CHECK(cudaMemsetAsync(d_slope_off, 0, sizeof(int64_t) * (size_t)(d.n_slopes + 2), stream));
  if (d.n_pairs == 0 || n_points == 0) return RB2_OK;
  if (!d_points) return RB2_ERR_ARG;
  hough_slope_count_kernel<<<(d.n_pairs + 255) / 256, 256, 0, stream>>>(d, (long long*)d_slope_off);
  hough_slope_scan_kernel<<<1, 1024, 0, stream>>>(d.n_slopes, (long long*)d_slope_off);
  int sms = 148;
  rb2_sm_count(&sms);
  hough_points_kernel<<<min(d.n_slopes, sms * 8), 256, 0, stream>>>(d, (const long long*)d_slope_off, d_points);
  RB2_CHECK(cudaGetLastError());
  return RB2_OK;
}
