/*
 * rascal_b200 — C ABI of the B200-native calibration hot path.
 *
 * The reference (rascal v0.3.9) is pure Python and has no FFI; the drop-in
 * boundary is its Python class API (rascal.calibrator.Calibrator /
 * HoughTransform).  `rascal_b200/calibrator.py` presents that API and binds the
 * entry points below with ctypes; a maintainer of the reference would add the
 * same ctypes stub.  Each entry point cites the reference code
 * it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer (the caller owns the memory:
 *     torch tensors in this repo); h_* are host pointers;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - every call returns 0 or a negative rb2_status; nothing throws;
 *   - every call is batched: a leading `n_problems` dimension, one descriptor
 *     per problem (an array of plain structs in device memory, pointers inside
 *     the structs are device pointers);
 *   - all floating point is IEEE binary64; counts are uint32/int32.
 */
#ifndef RASCAL_B200_H
#define RASCAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB2_ABI_VERSION 2
#define RB2_MAX_DEG 7       /* polynomial degree of a hypothesis            */
#define RB2_MAX_SAMPLE 16   /* sample_size + known pairs                    */
#define RB2_MAX_TOPN 32     /* top_n_candidate                              */

typedef enum {
  RB2_OK = 0,
  RB2_ERR_CUDA = -1,        /* a CUDA runtime call failed (see rb2_last_cuda_error) */
  RB2_ERR_ARG = -2,         /* bad argument                                  */
  RB2_ERR_CAPACITY = -3,    /* a device-side capacity was exceeded (flag in status word) */
  RB2_ERR_NO_DEVICE = -4
} rb2_status;

/* ---- diagnostics ------------------------------------------------------- */
int rb2_abi_version(void);
const char *rb2_status_string(int status);
const char *rb2_last_cuda_error(void);
/* Number of SMs of the current device (grid sizing is a multiple of this). */
int rb2_sm_count(int *out);
/* sizeof() of the structs below, in declaration order (0 = rb2_hough_desc ...
 * 5 = rb2_refit_result ... 11 = rb2_refine_desc); lets a binding verify its layout. */
int rb2_struct_size(int which);
/* Roofline denominator for the FP64-bound kernels (not in
 * MEASURED_PEAKS.json): launches `blocks` CTAs of 256 threads doing 8*iters
 * dependent-chain DFMAs each (2*8*iters*blocks*256 flops); the caller times it
 * with CUDA events on `stream`. */
int rb2_dfma_burn(int blocks, int iters, double *d_out, void *stream);

/* ======================================================================= *
 * K1  Hough accumulation
 *     replaces HoughTransform.generate_hough_points (houghtransform.py:60-92)
 *     and HoughTransform.bin_hough_points (houghtransform.py:167-210), as
 *     driven by Calibrator.do_hough_transform (calibrator.py:1425-1452).
 * ======================================================================= */
typedef struct {
  double min_slope, max_slope;          /* HoughTransform.set_constraints   */
  double min_intercept, max_intercept;
  int32_t n_slopes;                     /* num_slopes                       */
  int32_t xbins, ybins;                 /* gradient bins, intercept bins    */
  int32_t n_pairs;
  const double *d_pair_x;               /* pairs[:,0]  [n_pairs]            */
  const double *d_pair_y;               /* pairs[:,1]  [n_pairs]            */
  int32_t *d_lo, *d_hi;                 /* out: surviving slope interval per pair [n_pairs] */
  double *d_xedges;                     /* out: [xbins+1] == np.histogram2d xedges */
  double *d_yedges;                     /* out: [ybins+1]                   */
  uint32_t *d_hist;                     /* out: [xbins*ybins] counts, row-major (x,y) */
  int32_t *d_rank;                      /* out: [xbins*ybins] flat bin ids, canonical order */
  int32_t *d_rankpos;                   /* out: [xbins*ybins] inverse of d_rank */
  int32_t *d_sort_tmp;                  /* scratch: [xbins*ybins]           */
} rb2_hough_desc;

typedef struct {
  double icpt_min, icpt_max;            /* min/max of surviving intercepts  */
  double grad_min, grad_max;            /* min/max of surviving gradients   */
  int64_t n_points;                     /* number of surviving Hough points */
  int32_t slope_lo, slope_hi;           /* first/last slope index with survivors */
  uint32_t max_count;                   /* largest histogram count          */
  int32_t status;                       /* 0 or RB2_ERR_CAPACITY            */
} rb2_hough_stats;

/* Pass 1+2+3: intervals -> edges -> histogram -> canonical ranking.
 * d_stats[n_problems] is written by the call.  slopes_per_cta = 0 lets the
 * library choose. */
int rb2_hough_accumulate(int n_problems, const rb2_hough_desc *d_desc,
                         const rb2_hough_desc *h_desc, rb2_hough_stats *d_stats,
                         int slopes_per_cta, void *stream);

/* Lazy materialisation of HoughTransform.hough_points: [n_points,2]
 * (gradient, intercept), slope-major, exactly the reference's row order.
 * d_slope_off is scratch [n_slopes+2] int64. Needs d_lo/d_hi from
 * rb2_hough_accumulate. One problem per call. */
int rb2_hough_points(const rb2_hough_desc *h_desc, int64_t *d_slope_off,
                     double *d_points, int64_t n_points, void *stream);

/* HoughTransform.generate_hough_points_brute_force (houghtransform.py:94-140):
 * the line through every pair of (pixel, wavelength) pairs i < j of the
 * x-sorted input, kept iff gradient and intercept are strictly inside the
 * constraints.  Two calls: d_points == NULL counts (d_row_off[n-1] = total
 * after the call, d_row_cnt / d_row_off are [n] / [n+1] int64 scratch), then
 * with d_points [total*2] the rows are written in the reference's order. */
int rb2_hough_brute_force(const double *d_x_sorted, const double *d_y_sorted, int n,
                          double min_slope, double max_slope, double min_intercept,
                          double max_intercept, int64_t *d_row_cnt, int64_t *d_row_off,
                          double *d_points, void *stream);

/* HoughTransform.bin_hough_points (houghtransform.py:167-210) for EXPLICIT
 * points (brute force, add_hough_points, load): np.histogram2d with the
 * data-dependent range + canonical ranking.  Uses rb2_hough_desc with
 * d_pair_x = gradients, d_pair_y = intercepts, n_pairs = number of points;
 * slopes / limits / d_lo / d_hi are ignored. */
int rb2_hough_bin_points(int n_problems, const rb2_hough_desc *d_desc,
                         const rb2_hough_desc *h_desc, rb2_hough_stats *d_stats,
                         void *stream);

/* ======================================================================= *
 * K2  candidate vote
 *     replaces Calibrator._get_candidate_points_linear (calibrator.py:219-261)
 *     + Calibrator._get_most_common_candidates (calibrator.py:154-217),
 *     including the W[argsort(-agg)] indexing quirk.
 * ======================================================================= */
typedef struct {
  int32_t n_upeaks;                     /* unique peaks, ascending           */
  int32_t n_uwaves;                     /* unique atlas wavelengths          */
  int32_t xbins, ybins;
  int32_t top_n;                        /* top_n_candidate                   */
  int32_t weighted;                     /* candidate_weighted                */
  double tol;                           /* fit(candidate_tolerance=)         */
  double gauss_denom;                   /* 2*sigma**2 + 1e-9 (util.py:447)   */
  const double *d_upeak_x;              /* [n_upeaks]                        */
  const int32_t *d_pair_off;            /* [n_upeaks+1] CSR into the arrays below (original pair order inside a peak) */
  const double *d_pair_y;               /* [n_pairs] wavelength of each pair */
  const int32_t *d_pair_wid;            /* [n_pairs] id of that wavelength among the sorted unique wavelengths */
  const double *d_xedges, *d_yedges;    /* from K1                           */
  const int32_t *d_rankpos;             /* from K1                           */
  double *d_cand_wave;                  /* out: [n_upeaks*top_n]             */
  int32_t *d_cand_pair;                 /* out: [n_upeaks*top_n] pair index (global, CSR order) */
  int32_t *d_cand_n;                    /* out: [n_upeaks] n = min(top_n, #distinct) */
  double *d_agg;                        /* out: [n_upeaks*n_uwaves] aggregated weight (diagnostic / plotting) */
  int32_t *d_cnt;                       /* out: [n_upeaks*n_uwaves] hit counts */
  int32_t *d_status;                    /* out: [1] 0 or RB2_ERR_CAPACITY    */
  const int32_t *d_rank;                /* from K1 (bins in rank order), or NULL: with it and ascending pair
                                           wavelengths inside each peak the ordered hit list is read off the
                                           bins in rank order instead of being enumerated a second time */
} rb2_vote_desc;

int rb2_candidate_vote(int n_problems, const rb2_vote_desc *d_desc,
                       const rb2_vote_desc *h_desc, void *stream);

/* Calibrator.candidates (calibrator.py:239-261), the list the reference builds before the vote and its
 * plotting code reads (plotting.py:309-349): for every Hough line in rank order the pairs with
 * |gradient*x + intercept - wavelength| <= tol (in pair order) and their Gaussian weights.  Materialised on
 * demand only (the vote itself never builds it).  h_hough: the problem's K1 descriptor (host copy; its
 * d_xedges / d_yedges / d_rank / d_pair_x / d_pair_y are read).  Two calls: d_x == NULL writes the per-line
 * counts to d_off[1..B] (the caller turns them into offsets d_off[0..B], d_off[0] = 0), then with d_x / d_y /
 * d_w [d_off[B]] the entries of line r are written at d_off[r]. */
int rb2_candidate_points(const rb2_hough_desc *h_hough, double tol, double gauss_denom,
                         int64_t *d_off, double *d_x, double *d_y, double *d_w, void *stream);

/* ======================================================================= *
 * K3  batched RANSAC hypothesis draw / fit / test / score / reduce
 *     replaces the loop body of Calibrator._solve_candidate_ransac
 *     (calibrator.py:525-735): np.random.choice draws (:538-566), polyfit
 *     (:575), intercept test (:578-582), monotonicity test (:585-600),
 *     _match_bijective (:317-380), M-SAC cost with the Hough-density weight
 *     (:610-633) and the best-so-far selection rule (:636, App. A.4).
 * ======================================================================= */
typedef struct {
  /* candidate table (calibrator.py:488-495) */
  int32_t n_upeaks;                     /* peaks that have candidates, ascending */
  int32_t n_wids;                       /* distinct candidate wavelengths    */
  int32_t n_cand;                       /* total candidate entries == d_cand_off[n_upeaks] */
  int32_t pad0;
  const double *d_peaks;                /* [n_upeaks]                        */
  const int32_t *d_cand_off;            /* [n_upeaks+1]                      */
  const double *d_cand_wave;            /* [n_cand]                          */
  const int32_t *d_cand_wid;            /* [n_cand] id among sorted distinct candidate wavelengths */
  const uint32_t *d_cdf32;              /* [n_upeaks] floor(cdf*2^32) of the sampling law (calibrator.py:521-523) */
  /* known pairs (calibrator.py:569-571) */
  int32_t n_known;
  const double *d_known_x, *d_known_y;
  /* model */
  int32_t deg;                          /* fit_deg                           */
  int32_t sample_size;                  /* drawn peaks per hypothesis        */
  double min_intercept, max_intercept;  /* :578-582                          */
  double pix_min, pix_max;              /* monotonic grid np.arange(pix_min,pix_max,1), :585-592 */
  double tol;                           /* ransac_tolerance                  */
  /* Hough-density weight (calibrator.py:497-506, 614-627) */
  double hough_weight;
  double gc_min, gc_max;                /* first / last gradient-bin centre  */
  double ic_min;                        /* first intercept-bin centre        */
  double h_lo, h_hi;                    /* spline value at (gc_min, ic_min) / (gc_max, ic_min) */
  int32_t n_pp;                         /* pieces of the 1-D spline along the first intercept row */
  const double *d_pp_breaks;            /* [n_pp+1]                          */
  const double *d_pp_coef;              /* [n_pp*4] c0..c3 in (u - break)    */
  double w_upper;                       /* upper bound of the weight over ALL hypotheses (hough_weight *
                                           n_pix * max of the section spline), or <= 0 if the spline
                                           can go negative: lets K3 skip hypotheses that cannot win */
  int32_t n_pix;                        /* len(pixel_list)                   */
  const double *d_pix;                  /* sorted pixel_list, or NULL == arange(n_pix) */
  /* acceptance that does not need the refit (calibrator.py:640-649, 684-700) */
  int32_t min_inliers;                  /* n_inliers >= this (max(deg+1, minimum_matches)) */
  double min_inliers_frac;              /* n_inliers >= this (minimum_peak_utilisation*len(peaks)) */
  double cost_ceiling;                  /* cost <= this (1e50, or the prior's cost, :512-515) */
  double floor_cost;                    /* fall-through: only hypotheses ranked after (floor_cost, floor_id) */
  int64_t floor_id;                     /* -1 = no floor                     */
  /* work */
  int64_t hyp_begin;                    /* first hypothesis id of this call  */
  int64_t n_hyp;
  uint64_t seed;                        /* Philox4x32-10 key                 */
  const int32_t *d_replay_x;            /* [n_hyp*sample_size] peak indices, or NULL -> draw with Philox */
  const int32_t *d_replay_y;            /* [n_hyp*sample_size] candidate indices inside that peak's list */
  /* optional per-hypothesis outputs (NULL to skip) */
  uint8_t *d_out_status;                /* [n_hyp] rb2_hyp_status            */
  double *d_out_cost;                   /* [n_hyp]                           */
  double *d_out_weight;                 /* [n_hyp]                           */
  double *d_out_coeffs;                 /* [n_hyp*(deg+1)]                   */
  int32_t *d_out_counts;                /* [n_hyp*2] n_match, n_inliers      */
  int32_t *d_out_sample;                /* [n_hyp*sample_size*2] the draw itself: (peak index, candidate index
                                           inside that peak's list) per sampled point, in draw order; -1 for an
                                           aborted draw.  Lets a test check the sampler's law (App. A.8) on the
                                           device                                                                */
  const double *d_replay_coeffs;        /* with d_replay_x: [n_hyp*(deg+1)] hypothesis coefficients used INSTEAD
                                           of fitting the replayed sample (parity: the reference's own polyfit
                                           output, so that every later decision is checked on identical input)  */
  double min_fit_error;                 /* K4: the winner is accepted iff rms >= this (calibrator.py:673-679);
                                           used when rb2_refit_winners is called with minimum_fit_error = NaN   */
  /* general Hough-density weight: the clamped bicubic spline
     RectBivariateSpline(gradient centres, intercept centres, hist) itself, evaluated per pixel as the reference
     does (calibrator.py:614-627, arguments swapped and clamped to the knot range) for the hypotheses that
     leave the corner regime - nm-scale problems, where gradients and intercepts are of the same size.
     d_bs_c == NULL: such hypotheses are counted `unsupported` instead. */
  int32_t bs_nx, bs_ny;                 /* B-spline coefficients per axis (= xbins, ybins)               */
  const double *d_bs_tx, *d_bs_ty;      /* knots [bs_nx+4], [bs_ny+4] (FITPACK regrid, k = 3, s = 0)     */
  const double *d_bs_c;                 /* coefficients [bs_nx*bs_ny], row-major (x = gradient axis first) */
  const double *d_bs_ppx, *d_bs_ppy;    /* per knot interval m = 0..n-4 the four non-zero cubic B-splines
                                           N_m..N_m+3 as polynomials in (x - t[m+3]):
                                           [(n-3)][4 basis][4 coefficients, ascending powers]             */
  /* fit_type 'legendre' / 'chebyshev' (calibrator.py:1625-1631).  The hypothesis polynomial is the same
     least-squares polynomial whatever basis legfit / chebfit express it in (the coefficients everywhere in this
     ABI stay MONOMIAL); what the basis changes in the reference is (i) the intercept test, which reads the
     FIRST COEFFICIENT OF THE BASIS (:578-582), (ii) the "gradient" of the Hough-density weight -
     util._derivative, the monomial rule, applied to the basis coefficients and evaluated with the basis'
     polyval (:614-618) - and (iii) K4's residual: the basis' polyval applied to the monomial coefficients
     that models.robust_polyfit returns (:664-671).  basis != 0 needs the d_bs_* tables (the weight is then
     always evaluated per pixel). */
  int32_t basis;                        /* 0 poly, 1 legendre, 2 chebyshev                               */
  int32_t pad1;
  const double *d_basis_c0;             /* [deg+1]: first row of the monomial -> basis change matrix      */
  const double *d_basis_grad;           /* [deg][deg+1]: monomial coefficients of sum_k (k+1) c_(k+1) B_k(x),
                                           as a linear map of the hypothesis' monomial coefficients       */
} rb2_ransac_desc;

typedef enum {
  RB2_HYP_ABORT = 0,        /* impossible draw, try consumed (:557-566)      */
  RB2_HYP_INTERCEPT = 1,
  RB2_HYP_MONOTONIC = 2,
  RB2_HYP_EMPTY = 3,
  RB2_HYP_SCORED = 4,
  RB2_HYP_UNSUPPORTED = 5   /* weight outside the corner-clamped regime and no general spline tables given */
} rb2_hyp_status;

typedef struct {
  double cost;                          /* +inf when nothing eligible        */
  int64_t hyp_id;                       /* -1 when nothing eligible          */
  int32_t n_match, n_inliers;
  double coeffs[RB2_MAX_DEG + 1];       /* hypothesis coefficients, low order first */
  uint64_t counters[8];                 /* drawn, aborted, intercept, monotonic, empty, scored, eligible, unsupported */
} rb2_ransac_result;

/* Chunks of up to 2^25 hypothesis ids alternate over internal stream lanes that fork
 * from and join back into `stream` (events; capturable); everything the call
 * enqueues is ordered after prior work on `stream` and before later work on it.
 * One host thread per device drives this entry point.
 * blocks_per_problem = 0 lets the library choose (a multiple of the SM count
 * for one problem).  d_block_scratch must hold rb2_ransac_scratch_bytes().
 * All problems of one call share deg, sample_size and seed (else RB2_ERR_ARG). */
size_t rb2_ransac_scratch_bytes(int n_problems, int blocks_per_problem);
/* How rb2_ransac_batch cuts a call of n_problems x max_hyp hypothesis ids: ids per problem and chunk, the number
 * of chunks (each = 7 kernel launches, 8 with the general Hough weight) and the stream lanes they alternate over
 * (any output pointer may be NULL).  The reference has no counterpart: its loop (calibrator.py:525-735) is
 * one hypothesis at a time. */
int rb2_ransac_chunk_plan(int n_problems, long long max_hyp, long long *per_problem, int *n_chunks, int *n_lanes);
int rb2_ransac_batch(int n_problems, const rb2_ransac_desc *d_desc,
                     const rb2_ransac_desc *h_desc, rb2_ransac_result *d_result,
                     void *d_block_scratch, int blocks_per_problem, void *stream);

/* The one exchange of the multi-GPU path: merge n_sets record
 * sets laid out [set][problem] - the ncclAllGather'ed per-rank winners, or the
 * winners of successive launches - with the loop's own rule `cost <= best_cost`
 * (calibrator.py:636): lower cost, LATER hypothesis id on ties; counters add. */
int rb2_merge_winners(int n_problems, int n_sets, const rb2_ransac_result *d_sets,
                      rb2_ransac_result *d_out, void *stream);

/* ======================================================================= *
 * K4  winner refit + acceptance
 *     replaces Calibrator._match_bijective on the winner (calibrator.py:317-380),
 *     the inlier selection (:640-643), models.robust_polyfit (models.py:70-123)
 *     and the residual / rms / minimum_fit_error check (:664-679).
 *     mode 0 = the reference's solver restated step for step (SciPy TRF from
 *     ones, Huber loss, 2-point Jacobian, diff_step 1e-5); mode 1 = the exact
 *     Huber optimum that iteration approaches.
 * ======================================================================= */
typedef struct {
  double coeffs[RB2_MAX_DEG + 1];       /* refit coefficients, low order first */
  double rms;                           /* sqrt(mean(clipped residual^2))    */
  int32_t n_inliers;
  int32_t accepted;                     /* 1 if rms >= minimum_fit_error     */
  int32_t huber_iters;                  /* mode 0: SciPy nfev; mode 1: IRLS passes (0 = pure least squares) */
  int32_t pad;                          /* mode 0: SciPy termination status (1 gtol, 2 ftol, 3 xtol, 4 both, 0 max_nfev) */
} rb2_refit_result;

/* d_matched_x / d_matched_y / d_residual: [n_problems * max_upeaks] outputs
 * (inliers sorted by wavelength, first n_inliers entries valid).
 * minimum_fit_error: one threshold for every problem, or NaN = each problem's
 * own rb2_ransac_desc.min_fit_error. */
int rb2_refit_winners(int n_problems, const rb2_ransac_desc *d_desc,
                      const rb2_ransac_desc *h_desc,
                      const rb2_ransac_result *d_winner, double minimum_fit_error,
                      int mode, rb2_refit_result *d_out, double *d_matched_x,
                      double *d_matched_y, double *d_residual, int max_upeaks,
                      void *stream);

/* models.robust_polyfit (models.py:70-123) on arbitrary point sets, batched: problem p owns
 * points [off[p], off[p+1]) of d_x / d_y.  Used where the reference calls robust_polyfit
 * outside the RANSAC loop (Calibrator.match_peaks calibrator.py:1915-1920, manual_refit
 * :2113-2124).  rms = sqrt(mean((y - polyval(x))^2)) unclipped; n_inliers = point count;
 * accepted = 0 when a problem has <= deg points.  mode as in rb2_refit_winners. */
int rb2_robust_polyfit(int n_problems, const int32_t *d_off, const int32_t *h_off,
                       const double *d_x, const double *d_y, int deg, int mode,
                       rb2_refit_result *d_out, void *stream);

/* ======================================================================= *
 * K5  match_peaks: post-fit association of EVERY peak with the atlas
 *     replaces the body of Calibrator.match_peaks (calibrator.py:1824-1956,
 *     refine=False): per peak polyval -> atlas lines with |line - fit| <
 *     tolerance (:1830-1838), polyfit of the association (:1890-1903), then
 *     models.robust_polyfit on it (:1915-1920) and the signed residuals / rms
 *     (:1936-1941).  The reference's permutation search (:1841-1876) only runs
 *     to completion when every peak has at most ONE atlas line inside the
 *     tolerance (otherwise it raises); for an ambiguous peak the
 *     nearest line (first on ties) is taken and n_ambiguous counts such peaks.
 * ======================================================================= */
typedef struct {
  const double *d_peaks;                /* [n_peaks] in the caller's order   */
  const double *d_atlas;                /* [n_atlas] in atlas order          */
  int32_t n_peaks, n_atlas;
  double coeffs[RB2_MAX_DEG + 1];       /* solution to match with, low order first */
  const double *d_coeffs;               /* if non-NULL: read the n_coef coefficients from the device
                                           instead (e.g. rb2_refit_result.coeffs of K4: no host trip) */
  int32_t n_coef;                       /* its length                        */
  int32_t fit_deg;                      /* degree of polyfit / robust refit  */
  double tolerance;
  int32_t robust_refit;                 /* 0: stop after the polyfit         */
  int32_t basis;                        /* self.polyval of the fit: 0 poly, 1 legendre, 2 chebyshev - used where
                                           the reference calls it on `coeffs` (:1831) and on the refit result
                                           (:1936); polyfit_coeffs stay monomial */
} rb2_match_desc;

typedef struct {
  int32_t n_matched, n_ambiguous;
  double polyfit_coeffs[RB2_MAX_DEG + 1]; /* :1893, low order first          */
  double polyfit_err;                   /* sum |atlas - polyval| (:1896-1897) */
  double polyfit_rms;                   /* :1908-1910                        */
  rb2_refit_result refit;               /* robust refit; accepted = 0 if not run / too few points */
} rb2_match_result;

/* d_matched_x / d_matched_y / d_residual: [n_problems * max_peaks]; residual =
 * atlas - polyval(peak, refit) (signed, :1936) when robust_refit, else
 * |atlas - polyval(peak, polyfit)| (:1896). */
int rb2_match_peaks(int n_problems, const rb2_match_desc *d_desc,
                    const rb2_match_desc *h_desc, int mode,
                    rb2_match_result *d_out, double *d_matched_x,
                    double *d_matched_y, double *d_residual, int max_peaks,
                    void *stream);

/* Calibrator._adjust_polyfit (calibrator.py:754-820): the objective that match_peaks(refine=True) hands to
 * scipy.optimize.minimize (:1797-1821), for n_eval trial shifts at once.  Trial e adds d_deltas[e*n_delta + i] to
 * coefficient i of d_fit (i < n_delta); every peak is paired with its nearest atlas line (first on ties) if that
 * is closer than `tolerance`; the value is sum((line - polyval(peak))^2) / (n_matched - n_coef - 1), or +inf when
 * that divisor is < 1, fewer than min_frac * n_peaks peaks are matched, or the polynomial is not strictly
 * increasing over d_pix_sorted (np.sort(pixel_list)).  basis: self.polyval of the fit (0 poly, 1 legendre,
 * 2 chebyshev).  One warp per trial. */
int rb2_adjust_polyfit(int n_eval, const double *d_deltas, int n_delta, const double *d_fit,
                       int n_coef, int basis, const double *d_peaks, int n_peaks,
                       const double *d_atlas, int n_atlas, const double *d_pix_sorted, int n_pix,
                       double tolerance, double min_frac, double *d_out, void *stream);

/* ======================================================================= *
 * Device-resident glue of the many-spectra path: the host work the reference
 * does between its stages, on the device, so that a batch of spectra goes
 * peaks/lines -> pairs -> K1 -> K2 -> prepare -> K3 -> K4 -> K5 with no host
 * round trip.
 * ======================================================================= */
/* Calibrator._generate_pairs (calibrator.py:86-107, no polygon mask): the
 * peak-major Cartesian product peaks x lines, plus the per-peak CSR view of it
 * that K2 reads (d_pair_off / d_pair_wid may be NULL; wid = line index, i.e.
 * the lines must be strictly increasing). */
typedef struct {
  const double *d_peaks, *d_lines;
  int32_t n_peaks, n_lines;
  double *d_pair_x, *d_pair_y;          /* out: [n_peaks*n_lines]            */
  int32_t *d_pair_wid;                  /* out: [n_peaks*n_lines] or NULL    */
  int32_t *d_pair_off;                  /* out: [n_peaks+1] or NULL          */
} rb2_pairs_desc;

int rb2_expand_pairs(int n_problems, const rb2_pairs_desc *d_desc,
                     const rb2_pairs_desc *h_desc, void *stream);

/* The set-up at the top of Calibrator._solve_candidate_ransac
 * (calibrator.py:441-523, filter_close=False) from K1/K2 outputs that are
 * still on the device.  d_ransac_desc[i] must already hold every pointer and
 * every host-known scalar (deg, sample_size, tolerances, n_pix, hough_weight,
 * min_inliers..., n_hyp, and d_peaks/d_cand_off/d_cand_wave/d_cand_wid/
 * d_cdf32/d_pp_breaks/d_pp_coef sized for n_peaks, n_peaks+1,
 * n_peaks*top_n (x2), n_peaks, n_pp+1, 4*n_pp); the kernel fills those tables
 * and the data-dependent scalars (n_upeaks, n_cand, n_wids, pix_min/max,
 * gc_min/max, ic_min, h_lo/h_hi, w_upper) and sets n_hyp = 0 where the
 * reference would return before the loop (:468-469; d_enough[0] = 0) and where
 * its sampling weights are not finite (a peak on an interior edge of the 10-bin
 * histogram next to an empty bin: np.random.choice raises "probabilities contain
 * NaN", :521-540; d_enough[0] = -1). */
typedef struct {
  const double *d_upeak_x;              /* [n_peaks] ascending (K2's input)  */
  const double *d_cand_wave;            /* [n_peaks*top_n] K2 output         */
  const int32_t *d_cand_n;              /* [n_peaks]       K2 output         */
  int32_t n_peaks, top_n;
  const double *d_xedges, *d_yedges;    /* K1 outputs                        */
  const uint32_t *d_hist;               /* K1 output [xbins*ybins]           */
  int32_t xbins, ybins;
  const double *d_spline_op;            /* [4*n_pp][xbins]: pp-form (Taylor coefficients at the left breaks,
                                           integer grid) of the interpolating cubic spline as a linear map of
                                           the data; depends on xbins only                                    */
  const double *d_spline_brk;           /* [n_pp+1] breaks on the integer grid */
  int32_t n_pp, pad;
  int32_t *d_enough;                    /* out: [1] or NULL                  */
} rb2_prepare_desc;

int rb2_ransac_prepare(int n_problems, const rb2_prepare_desc *d_desc,
                       const rb2_prepare_desc *h_desc,
                       rb2_ransac_desc *d_ransac_desc, void *stream);

/* ======================================================================= *
 * K6  peak finding and sub-pixel refinement: the step BEFORE the path
 *     (SURVEY.md 8f rank 2), batched over spectra.
 * ======================================================================= */
/* rb2_find_peaks replaces scipy.signal.find_peaks(spectrum, height=h,
 * distance=d, prominence=p, threshold=None) as every reference example and
 * test calls it before refine_peaks (test/test_xe_calibration.py:69,
 * test/test_lt_sprat_manual_atlas.py:36, examples/example_*.py): strict local
 * maxima with flat tops reported at their middle sample, then - in SciPy's
 * order - the height, distance (highest first; among equal heights the later
 * sample first) and prominence (wlen=None) conditions.  A condition is
 * switched off with NaN.  One CTA per spectrum, candidates in shared memory:
 * n_pix <= RB2_FIND_MAX_PIX, else RB2_ERR_CAPACITY. */
#define RB2_FIND_MAX_PIX 16382
typedef struct {
  const double *d_spectrum;             /* [n_pix]                           */
  int32_t n_pix;
  int32_t max_peaks;                    /* capacity of the outputs           */
  double height, distance, prominence;  /* minima; NaN = not asked           */
  int32_t *d_peaks;                     /* out: [max_peaks] sample indices, ascending */
  double *d_peaks_f64;                  /* out or NULL: the same as doubles (input of rb2_refine_peaks) */
  double *d_heights;                    /* out or NULL: spectrum[peak]       */
  double *d_prominences;                /* out or NULL (needs prominence)    */
  int32_t *d_n_peaks;                   /* out: [1] number found; only the first max_peaks are stored */
} rb2_findpeaks_desc;

int rb2_find_peaks(int n_spectra, const rb2_findpeaks_desc *d_desc,
                   const rb2_findpeaks_desc *h_desc, void *stream);

/* rb2_refine_peaks replaces rascal.util.refine_peaks (util.py:450-523) with
 * util.gauss (:425-447): per peak the window [int(peak)-w, int(peak)+w) is
 * divided by its maximum IN the working copy of the spectrum (:478 - a later,
 * overlapping window sees the divided values; d_work is that copy), and a
 * Gaussian a*exp(-(x-x0)^2/(2 sigma^2+1e-9)) is fitted from p0 = [1, mean,
 * sigma] (:487-488) by MINPACK's lmdif exactly as scipy.optimize.curve_fit runs
 * it (ftol = xtol = 1.49012e-8, gtol = 0, maxfev = 800, factor 100, forward
 * differences with step sqrt(eps)*|p|); a peak is dropped when the fit fails
 * (ier not in 1..4), the height is negative or the centre leaves [0, n_pix]
 * (:491-502); the survivors peak - w + centre are filtered to (0, n_pix) and
 * de-duplicated against their predecessor with np.isclose (:509-517).
 * Four launches: windows (one CTA per spectrum), the scan that turns the peak
 * counts into one work queue, fits (persistent one-warp CTAs; an idle lane
 * takes the batch's next peak), collection (one thread per spectrum).
 * d_scratch: rb2_refine_scratch_bytes(n_spectra) bytes. */
typedef enum {
  RB2_PEAK_OK = 0,          /* fitted and appended (it may still fall to the final range / isclose filter) */
  RB2_PEAK_NO_DATA = 1,     /* no finite sample in the window (:484-485)      */
  RB2_PEAK_NOT_FINITE = 2,  /* some sample not finite: curve_fit returns NaN, removed by the final mask */
  RB2_PEAK_FIT_FAILED = 3,  /* lmdif ier not in 1..4: RuntimeError in the reference (:504) */
  RB2_PEAK_NEGATIVE = 4,    /* fitted height < 0 (:493)                       */
  RB2_PEAK_OUTSIDE = 5,     /* fitted centre outside [0, n_pix] (:496)        */
  RB2_PEAK_EMPTY_WINDOW = 6,/* window empty: the reference raises ValueError (nanmax of nothing) */
  RB2_PEAK_TOO_FEW = 7      /* fewer than 3 samples: the reference raises TypeError (curve_fit) */
} rb2_peak_status;

#define RB2_REFINE_MAX_WINDOW 32        /* largest window_width              */

typedef struct {
  const double *d_spectrum;             /* [n_pix]                           */
  double *d_work;                       /* scratch: [n_pix]                  */
  int32_t n_pix;
  int32_t n_peaks;                      /* number of peaks, or their capacity if d_n_peaks is given */
  const double *d_peaks;                /* [n_peaks] initial estimates (pixels) */
  const int32_t *d_n_peaks;             /* NULL, or [1] on the device (rb2_find_peaks output), clamped to n_peaks */
  int32_t window_width, pad;
  double *d_window;                     /* scratch: [n_peaks * 2*window_width] */
  double *d_fit;                        /* out: [n_peaks][4] height, centre, sigma, (double)nfev */
  int32_t *d_status;                    /* out: [n_peaks] rb2_peak_status; lmdif's ier in bits 8.. */
  double *d_refined;                    /* out: [n_peaks] the refined peaks, compacted */
  int32_t *d_n_refined;                 /* out: [1]                          */
} rb2_refine_desc;

size_t rb2_refine_scratch_bytes(int n_spectra);
int rb2_refine_peaks(int n_spectra, const rb2_refine_desc *d_desc,
                     const rb2_refine_desc *h_desc, void *d_scratch,
                     void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RASCAL_B200_H */
