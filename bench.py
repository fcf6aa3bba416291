#!/usr/bin/env python
"""Headline benchmark: RANSAC hypotheses/s on the synthetic 60-peak / 300-atlas-line
degree-4 spectrum (BASELINE.json configs[2]; SURVEY.md 8d "C3 concrete input").

    python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path
    python bench.py --impl reference ...                     # the reference's CPU algorithm

A step = one pass of the hot path over one batch of hypotheses: H hypothesis ids of
the Philox stream (default 1e8, the configuration the target is quoted on), split
evenly over the N ranks (strong scaling: total work fixed), each rank running K3
(draw -> fit -> intercept/monotonicity tests -> bijective match -> M-SAC cost ->
block-best), ONE all-gather of the per-rank winner records, K4 (winner refit) and the
D2H copy of the result.  The K timed steps run back to back the way a service that
calibrates one spectrum after the other runs them: the exchange, K4 (one warp) and the
D2H of step k sit on a side stream and overlap K3 of step k + 1; every result is on the
host before the closing event.  `sequential` reports the same K steps one at a time
(the latency of a resident fit).
Hough + candidate vote (K1, K2) run once per spectrum and are part of the e2e leg,
which goes through the public Calibrator API with host buffers every step.

Prints ONE JSON line (rank 0).  Beyond the contract's keys it carries three auxiliary legs
(skipped with --spectra 0): `spectra_per_sec` = a bounded sample of BASELINE configs[3] (C4
spectra through rascal_b200.batch.calibrate_batch, K1..K5 each, host arrays in / results out),
`hough_roofline` = K1 and K2 alone on 512 resident C4 spectra against the FP64-pipe roof, and
`peak_finding` = K6 (find_peaks + refine_peaks) on 512 rendered C4 arcs, resident and from the
raw arcs to wavelength solutions, with the reference's recipe timed on one host core beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ransac_hypotheses_per_sec"
UNIT = "hypotheses/s"


# ---------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------


def c3_input():
    rng = np.random.default_rng(1234)
    npix = 4096
    coeff = np.array([4000.0, 1.2, 2e-5, -3e-9, 1e-13])
    lam = lambda x: np.polynomial.polynomial.polyval(x, coeff)  # noqa: E731
    true_pix = np.sort(rng.uniform(50, npix - 50, 60))
    peaks = true_pix + rng.normal(0, 0.05, 60)
    atlas = np.sort(np.concatenate((lam(true_pix), rng.uniform(lam(0) - 400, lam(npix - 1) + 400, 240))))
    return dict(peaks=peaks, atlas=atlas, npix=npix, truth=coeff,
                hough=dict(num_slopes=2000, xbins=100, ybins=100, min_wavelength=float(lam(0)),
                           max_wavelength=float(lam(npix - 1)), range_tolerance=500, linearity_tolerance=100),
                ransac=dict(sample_size=5, top_n_candidate=5, ransac_tolerance=5),
                fit=dict(fit_deg=4, candidate_tolerance=10.0))


def c4_input(i):
    """SURVEY.md 8d "C4 concrete input": spectrum i of the 10k-spectrum batch."""
    P = np.polynomial.polynomial
    rng = np.random.default_rng(10_000 + i)
    npix = int(rng.choice([1024, 2048, 4096]))
    lam0 = rng.uniform(3500, 6500)
    dlam = rng.uniform(1500, 5000)
    while True:
        a2, a3, a4 = rng.uniform(-0.05, 0.05), rng.uniform(-0.02, 0.02), rng.uniform(-0.01, 0.01)
        cu = np.array([0.0, 1.0, a2, a3, a4]) / (1 + a2 + a3 + a4)
        coeff = np.array([lam0] + [dlam * cu[k] / npix ** k for k in range(1, 5)])
        if np.all(np.diff(P.polyval(np.arange(npix, dtype=float), coeff)) > 0):
            break
    n_peaks = int(rng.integers(30, 81))
    # sorted uniform positions conditioned on >= 3 px separation (constructive, no rejection loop)
    span = 0.98 * npix - 3.0 * (n_peaks - 1)
    pix = 0.01 * npix + np.sort(rng.uniform(0, span, n_peaks)) + 3.0 * np.arange(n_peaks)
    sigma = rng.uniform(0.02, 0.3)
    peaks = pix + rng.normal(0, sigma, n_peaks)
    spurious = rng.random(n_peaks) < rng.uniform(0, 0.3)
    true_waves = P.polyval(pix[~spurious], coeff)
    m = int(rng.integers(2, 6))
    lo, hi = P.polyval(0.0, coeff) - 400, P.polyval(npix - 1.0, coeff) + 400
    atlas = np.sort(np.concatenate((true_waves, rng.uniform(lo, hi, m * len(true_waves)))))
    return dict(peaks=peaks, lines=atlas, num_pix=npix, truth=coeff,
                hough=dict(num_slopes=2000, xbins=100, ybins=100,
                           min_wavelength=float(P.polyval(0.0, coeff) + rng.uniform(-200, 200)),
                           max_wavelength=float(P.polyval(npix - 1.0, coeff) + rng.uniform(-200, 200)),
                           range_tolerance=500, linearity_tolerance=100),
                ransac=dict(sample_size=5, top_n_candidate=5, ransac_tolerance=5),
                fit=dict(fit_deg=4, candidate_tolerance=10.0))


def c4_arc(sp, i):
    """The arc spectrum whose peaks spectrum i of the C4 batch has: Gaussian lines (sigma 1.1-1.8 px,
    200-20000 counts) at sp["peaks"] on a sloping baseline with read noise.  Input of the K6 leg."""
    rng = np.random.default_rng(20_000 + i)
    n = int(sp["num_pix"])
    s = 40.0 + 0.005 * np.arange(n) + rng.normal(0, 3.0, n)
    for c, a, w in zip(sp["peaks"], rng.uniform(200, 20000, len(sp["peaks"])), rng.uniform(1.1, 1.8, len(sp["peaks"]))):
        lo, hi = max(0, int(c) - 12), min(n, int(c) + 13)
        x = np.arange(lo, hi)
        s[lo:hi] += a * np.exp(-0.5 * ((x - c) / w) ** 2)
    return s


def flops_per_hypothesis(c, deg=4, s=5, n=60, k=5):
    """F = F_fit + f_i*F_mono + f_m*F_score (FP64 flops, algorithmic)."""
    d = deg
    f_fit = s * (d - 1) + 2 * s * (d + 1) + (2.0 / 3.0) * (d + 1) ** 3 + 2 * (d + 1) ** 2
    f_mono = 2 * (2 * d) * (2 * (d - 2) + 2) + 60
    f_score = 2 * d * n + 2 * k * n + 3 * n + 600
    drawn = max(1, c["drawn"])
    f_i = (drawn - c["intercept"]) / drawn
    f_m = (c["scored"] + c["empty"] + c["unsupported"]) / drawn
    return f_fit + f_i * f_mono + f_m * f_score, f_i, f_m


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------


class ClockSampler:
    """One long-running `nvidia-smi -lms 20` started before the timed region and stopped after
    it (B200_PROFILING.md's clocks line); samples inside [t_begin, t_end] are summarised."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._read, daemon=True)
            self._t.start()
            time.sleep(0.25)  # let the first samples arrive before the timed region starts
            self.skip = len(self.rows)
        except Exception:
            self.proc = None
            self.skip = 0
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.strip().split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        rows = self.rows[max(0, self.skip - 1):]
        ok = [r for r in rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        sm = [float(r[0]) for r in ok]
        mx = [float(r[1]) for r in ok if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in ok for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU legs: the reference itself (baseline/_ref, scripts/make_baseline_ref.py) when the snapshot carries
# it, else the oracle (a port of its algorithm; test infrastructure) - `kind` says which
# ---------------------------------------------------------------------------

REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_available():
    return os.path.isdir(os.path.join(REF_DIR, "rascal"))


def config_input(name):
    """Workload of a BASELINE config: dict(peaks, atlas, npix, hough, ransac, fit, tries).  c3 is synthetic;
    c2 / c5 take peaks and atlas lines from the recordings of the real-data cases (tests/golden/*.npz: the
    reference's own peak-finding recipe and NIST atlas, see tests/golden/make_golden_configs.py) with the
    sizes BASELINE.json states."""
    if name == "c3":
        w = c3_input()
        w["tries"] = None
        return w
    g = np.load(os.path.join(ROOT, "tests", "golden", {"c2": "c2_gmos_cuar", "c5": "c5_deimos"}[name] + ".npz"))
    cfg = json.loads(str(g["config_json"]))
    w = dict(peaks=g["peaks"], atlas=g["lines"], npix=len(g["pixel_list"]), truth=None, hough=dict(cfg["hough"]),
             ransac=dict(cfg["ransac"]), fit={k: v for k, v in cfg["fit"].items() if k != "max_tries"})
    if name == "c2":   # Gemini GMOS CuAr, degree 5, 1e5 RANSAC tries
        w["fit"]["fit_deg"] = 5
        w["tries"] = 100_000
    else:              # Keck DEIMOS 830G, fine Hough binning 2000 x (1000 x 1000), 1e6 tries
        w["hough"].update(num_slopes=2000, xbins=1000, ybins=1000)
        w["tries"] = 1_000_000
    w["fit"].setdefault("candidate_tolerance", 2.0)
    return w


WORKLOADS = {
    "c3": "C3: synthetic 60-peak/300-atlas-line arc, degree 4, sample 5, top-5 candidates (BASELINE configs[2])",
    "c2": "C2: Gemini GMOS CuAr arc (ground-truth spectrum of the reference), 67 peaks / 53 lines, Hough 5000 x (200 x 200), "
          "degree 5, 1e5 RANSAC tries (BASELINE configs[1])",
    "c5": "C5: Keck DEIMOS 830G Ne+Ar+Kr arc, 51 peaks / 100 lines, Hough 2000 x (1000 x 1000) = 1e6 lines, degree 4, "
          "1e6 RANSAC tries (BASELINE configs[4])",
}


def _reference_worker(args):
    """One process: do_hough_transform() + fit(max_tries) of the reference (or the oracle port) on `config`;
    returns (polyfit calls = hypotheses, seconds: hough, candidates, ransac loop)."""
    config, seed, max_tries, use_ref = args
    w = config_input(config)
    if not use_ref:
        from oracle import rascal_oracle as O

        trace, tm = [], {}
        O.run_path(w["peaks"], w["atlas"], num_pix=w["npix"], max_tries=max_tries, seed=seed, trace=trace,
                   timings=tm, **w["hough"], **w["ransac"], **w["fit"])
        return len(trace), tm
    import logging
    import warnings

    warnings.filterwarnings("ignore")
    logging.disable(logging.WARNING)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from rascal import calibrator as rc
    from rascal.atlas import Atlas

    n = [0]
    tm = dict(hough=0.0, candidates=0.0, ransac=0.0)
    orig_fit = np.polynomial.polynomial.polyfit

    def counted(*a, **k):
        n[0] += 1
        return orig_fit(*a, **k)

    def timed(fn):
        def wrap(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                tm["candidates"] += time.perf_counter() - t0
        return wrap

    C = rc.Calibrator
    saved = (C._get_candidate_points_linear, C._get_most_common_candidates)
    C._get_candidate_points_linear, C._get_most_common_candidates = timed(saved[0]), timed(saved[1])
    np.polynomial.polynomial.polyfit = counted
    try:
        c = C(w["peaks"], spectrum=np.zeros(w["npix"]))
        c.set_calibrator_properties(seed=seed)
        c.set_hough_properties(**w["hough"])
        c.set_ransac_properties(**w["ransac"])
        a = Atlas()
        a.add_user_atlas(elements=["X"] * len(w["atlas"]), wavelengths=w["atlas"])
        c.set_atlas(a, candidate_tolerance=10.0)
        t0 = time.perf_counter()
        c.do_hough_transform()
        tm["hough"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        try:
            c.fit(max_tries=max_tries, progress=False, **w["fit"])
        except AssertionError:  # "Couldn't fit": the timing stands
            pass
        tm["ransac"] = time.perf_counter() - t0 - tm["candidates"]
    finally:
        np.polynomial.polynomial.polyfit = orig_fit
        C._get_candidate_points_linear, C._get_most_common_candidates = saved
    return n[0], tm


def truth_rms(sp, coeff):
    """RMS of fitted minus true wavelength over the detector (A): the C4 success criterion."""
    P = np.polynomial.polynomial
    x = np.linspace(0, sp["num_pix"] - 1, 256)
    return float(np.sqrt(np.mean((P.polyval(x, coeff) - P.polyval(x, sp["truth"])) ** 2)))


def _reference_c4_worker(args):
    """Spectrum i of the C4 batch through the reference itself (or the port): (rms vs truth or None, hypotheses)."""
    i, max_tries, use_ref = args
    sp = c4_input(i)
    if not use_ref:
        from oracle import rascal_oracle as O

        trace = []
        out = O.run_path(sp["peaks"], sp["lines"], num_pix=sp["num_pix"], max_tries=max_tries, seed=i, trace=trace,
                         **sp["hough"], **sp["ransac"], **sp["fit"])
        co = out["best"]["best_p"]
        return (None if co is None else truth_rms(sp, co)), len(trace)
    import logging
    import warnings

    warnings.filterwarnings("ignore")
    logging.disable(logging.WARNING)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from rascal import calibrator as rc
    from rascal.atlas import Atlas

    n = [0]
    orig_fit = np.polynomial.polynomial.polyfit

    def counted(*a, **k):
        n[0] += 1
        return orig_fit(*a, **k)

    np.polynomial.polynomial.polyfit = counted
    try:
        c = rc.Calibrator(sp["peaks"], spectrum=np.zeros(sp["num_pix"]))
        c.set_calibrator_properties(seed=i)
        c.set_hough_properties(**sp["hough"])
        c.set_ransac_properties(**sp["ransac"])
        a = Atlas()
        a.add_user_atlas(elements=["X"] * len(sp["lines"]), wavelengths=sp["lines"])
        c.set_atlas(a, candidate_tolerance=10.0)
        c.do_hough_transform()
        try:
            co = c.fit(max_tries=max_tries, progress=False, **sp["fit"])[0]
        except AssertionError:
            co = None
    finally:
        np.polynomial.polynomial.polyfit = orig_fit
    return (None if co is None else truth_rms(sp, co)), n[0]


def c4_reference_quality(n_spectra=16, max_tries=170):
    """The reference's own success rate on the first spectra of the C4 batch at (about) the draw budget the GPU
    leg uses: fit(max_tries=170) draws ~1e4 hypotheses per spectrum on this recipe."""
    import multiprocessing as mp

    use_ref = reference_available()
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(min(os.cpu_count() or 1, n_spectra)) as pool:
        res = pool.map(_reference_c4_worker, [(i, max_tries, use_ref) for i in range(n_spectra)])
    dt = time.perf_counter() - t0
    rms = [r for r, _ in res if r is not None]
    return {"kind": "reference" if use_ref else "port", "spectra": n_spectra, "max_tries": max_tries,
            "hypotheses_per_spectrum": float(np.mean([h for _, h in res])), "fitted": len(rms),
            "within_1A": int(sum(r < 1.0 for r in rms)), "within_4A": int(sum(r < 4.0 for r in rms)),
            "spectra_per_s_all_cores": n_spectra / dt, "cores": min(os.cpu_count() or 1, n_spectra)}


REF_TRIES = {"c3": 600, "c2": 300, "c5": 150}   # bounded samples: ~10-30 s of RANSAC loop on one core


def cpu_baseline(config="c3", max_tries=None):
    """Single-thread reference (or port) on a bounded sample of the same workload."""
    use_ref = reference_available()
    max_tries = max_tries or REF_TRIES[config]
    n, tm = _reference_worker((config, 0, max_tries, use_ref))
    return {"value": n / tm["ransac"], "unit": UNIT, "cores": 1, "kind": "reference" if use_ref else "port",
            "sample": "%s spectrum, fit(max_tries=%d): %d hypotheses (polyfit calls) in %.1f s of RANSAC loop; Hough %.2f s "
                      "and candidate collection %.2f s are not in the rate" % (config.upper(), max_tries, n, tm["ransac"],
                                                                               tm["hough"], tm["candidates"]),
            "hough_s": tm["hough"], "candidates_s": tm["candidates"], "ransac_loop_s": tm["ransac"], "hypotheses": n}


def reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation (baseline/_ref; the oracle port if the
    snapshot does not carry it) on all host cores: independent worker processes with distinct seeds, each a
    bounded sample fit(max_tries=...) of the same workload."""
    import multiprocessing as mp

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_ref = reference_available()
    cores = os.cpu_count() or 1
    config = args.config
    max_tries = max(20, REF_TRIES[config] // 4)
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(args.warmup and 1):
            pool.map(_reference_worker, [(config, 1000 + i, 5, use_ref) for i in range(cores)])
        times, hyps, hough, cand = [], [], [], []
        for k in range(args.steps):
            res = pool.map(_reference_worker, [(config, k * cores + i, max_tries, use_ref) for i in range(cores)])
            # hypotheses/s of the RANSAC loop itself: every worker's draws over the slowest loop
            times.append(max(tm["ransac"] for _, tm in res))
            hyps.append(sum(n for n, _ in res))
            hough.append(float(np.median([tm["hough"] for _, tm in res])))
            cand.append(float(np.median([tm["candidates"] for _, tm in res])))
    value = sum(hyps) / sum(times)
    kind = "reference" if use_ref else "port"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic" if config == "c3" else "real arc (recorded peaks / NIST lines)",
            "config": {"workload": WORKLOADS[config], "sample": "fit(max_tries=%d) per worker" % max_tries},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "%d worker processes x fit(max_tries=%d) x %d steps of the %s spectrum (RANSAC loop time "
                                       "only; per worker the Hough transform takes %.2f s and candidate collection %.2f s on top)"
                                       % (cores, max_tries, args.steps, config.upper(), float(np.median(hough)),
                                          float(np.median(cand))),
                             "hough_s_per_spectrum": float(np.median(hough)),
                             "candidates_s_per_spectrum": float(np.median(cand))},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def k3_profile(key):
    """Figures of the K3 stage kernels from the committed `ncu --set full` capture (profiles/k3_traffic.json):
    "dram_bytes_per_chunk" = DRAM read + write bytes of the six stage kernels of one chunk ("hypotheses_per_chunk" ids,
    3.3e7 in a 1e8-hypothesis step), "issue_slots_busy_pct" = issue-slot utilisation per stage kernel; None if the
    file is absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "k3_traffic.json")) as f:
            return json.load(f)[key]
    except Exception:
        return None


def k3_traffic():
    return k3_profile("dram_bytes_per_chunk")


def k3_chunk_plan(lib, n_hyp):
    """(chunks, lanes, ids per chunk) rb2_ransac_batch uses for one problem of n_hyp hypothesis ids."""
    import ctypes as C

    pp, nc, nl = C.c_longlong(0), C.c_int(0), C.c_int(0)
    lib.rb2_ransac_chunk_plan(1, int(n_hyp), C.byref(pp), C.byref(nc), C.byref(nl))
    return nc.value, nl.value, pp.value


WORKLOAD_NAME = WORKLOADS["c3"]


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------


def fp64_peak_tflops(lib, engine, torch, sms):
    """FP64 roofline denominator: DFMA microbenchmark (rb2_dfma_burn) timed with CUDA events in this run."""
    sink = torch.zeros(1, dtype=torch.float64, device="cuda")
    blocks, iters = sms * 16, 1 << 15
    for _ in range(2):
        lib.rb2_dfma_burn(blocks, iters, sink.data_ptr(), engine.current_stream_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    lib.rb2_dfma_burn(blocks, iters, sink.data_ptr(), engine.current_stream_ptr())
    b.record()
    torch.cuda.synchronize()
    return 2.0 * 8 * iters * blocks * 256 / (a.elapsed_time(b) * 1e-3) / 1e12


def config_bench(args, torch, dist, rank, world, local, group):
    """`--config c2|c5`: the real-data BASELINE configs at their stated sizes through the UNCHANGED public
    signature - do_hough_transform() + fit(max_tries=..., fit_deg=...) - plus the stage times on resident
    inputs.  A step = one calibration.  `e2e` = hypotheses drawn / wall time of the public calls from host arrays;
    `value` = the same draws through K3 + exchange + K4 with the candidate tables resident."""
    from rascal_b200 import _native, engine, parallel
    from rascal_b200.calibrator import Calibrator, LineList

    lib = _native.load()
    name = args.config
    w = config_input(name)
    tries = int(w["tries"])
    deg, ssz = int(w["fit"].get("fit_deg", 4)), int(w["ransac"].get("sample_size", 5))

    def new_calibrator(seed):
        c = Calibrator(w["peaks"], spectrum=np.zeros(w["npix"]))
        c.set_calibrator_properties(seed=seed)
        c.set_hough_properties(**w["hough"])
        c.set_ransac_properties(**w["ransac"])
        c.set_atlas(LineList(w["atlas"]), candidate_tolerance=10.0)
        if world > 1:
            c.use_distributed(None)
        return c

    # ---- e2e: the public API, host arrays in, 7-tuple out ----
    for k in range(max(1, args.warmup)):
        c = new_calibrator(5 + k)
        c.do_hough_transform()
        res = c.fit(max_tries=tries, progress=False, **w["fit"])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    engine.TRANSFER["h2d"] = engine.TRANSFER["d2h"] = 0
    drawn = scored = 0
    t_hough = t_fit = 0.0
    with ClockSampler(local) as clocks:
        t0 = time.perf_counter()
        for k in range(args.steps):
            c = new_calibrator(100 + k)
            ta = time.perf_counter()
            c.do_hough_transform()
            c.ht.hist  # noqa: B018  (K1 runs when its results are first read)
            torch.cuda.synchronize()
            tb = time.perf_counter()
            res = c.fit(max_tries=tries, progress=False, **w["fit"])
            torch.cuda.synchronize()
            t_hough += tb - ta
            t_fit += time.perf_counter() - tb
            drawn += c.ransac_counters["drawn"] + c.ransac_counters["aborted"]
            scored += c.ransac_counters["scored"] + c.ransac_counters["aborted"]
        if world > 1:
            dist.barrier()
        t_e2e = time.perf_counter() - t0
    t_e2e_t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e_t, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e2e_t.item())
    e2e = {"value": drawn / t_e2e, "unit": UNIT, "h2d_bytes_per_step": engine.TRANSFER["h2d"] // args.steps,
           "d2h_bytes_per_step": engine.TRANSFER["d2h"] // args.steps, "ms_per_step": 1e3 * t_e2e / args.steps,
           "tries_per_s": scored / t_e2e, "completed_tries_per_step": scored // args.steps,
           "hough_ms_per_step": 1e3 * t_hough / args.steps, "fit_ms_per_step": 1e3 * t_fit / args.steps,
           "rms": float(res[3]), "n_matched": len(res[1])}

    # ---- resident legs: K1 / K2 alone, then K3 (+ exchange + K4) over the same number of draws ----
    H = drawn // args.steps
    c = new_calibrator(0)
    c.do_hough_transform()
    c.ht.hist  # noqa: B018
    hb = c.ht._batch
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    hb.run()
    torch.cuda.synchronize()
    e0.record()
    hb.run()
    e1.record()
    c._vote(w["fit"]["candidate_tolerance"])
    e2.record()
    torch.cuda.synchronize()
    k1_ms, k2_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)
    setup = engine.RansacSetup(
        c.candidate_peak, c.candidate_arc, fit_deg=deg, sample_size=max(ssz, deg + 1), ransac_tolerance=c.ransac_tolerance,
        min_intercept=c.min_intercept, max_intercept=c.max_intercept, pixel_list=c.pixel_list, hist=c.ht.hist,
        xedges=c.ht.xedges, yedges=c.ht.yedges, hough_weight=c.hough_weight, filter_close=c.filter_close,
        n_peaks_total=len(c.peaks))
    rb = engine.RansacBatch([setup])
    lo, hi = parallel.shard_range(H, rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(k):
        rb.run(k * H + lo, hi - lo, seed=2026)
        rb.exchange(group)
        rb.refit_async()
        return rb.fetch()

    for k in range(max(3, args.warmup)):
        step(k)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    counters = dict.fromkeys(_native.COUNTER_NAMES, 0)
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        kev[k][0].record()
        rb.run((3 + k) * H + lo, hi - lo, seed=2026)
        kev[k][1].record()
        rb.exchange(group)
        rb.refit_async()
        win = rb.fetch()[0]
        ev[k][1].record()
        for nm, v in engine.counters_dict(win[0]).items():
            counters[nm] += v
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device="cuda")
    kern_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in kev)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = float(total_ms.item()), float(kern_ms.item())
    peak = fp64_peak_tflops(lib, engine, torch, rb.sm_count)
    n_u, n_c = len(setup.peaks), len(setup.cand_wave)
    F, f_i, f_m = flops_per_hypothesis(counters, deg=deg, s=setup.sample_size, n=n_u, k=max(1, n_c // max(1, n_u)))
    achieved = (hi - lo) * args.steps * F / (kern_ms * 1e-3) / 1e12
    P_pairs, S, B = len(c.pairs), int(w["hough"]["num_slopes"]), int(w["hough"]["xbins"]) * int(w["hough"]["ybins"])
    line = {
        "metric": METRIC, "value": args.steps * H / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "real arc (recorded peaks / NIST lines), random draws",
        "config": {"workload": WORKLOADS[name], "max_tries": tries, "hypotheses_per_step": H,
                   "parallelism": "hypothesis-id ranges x%d" % world, "l2": "flushed (256 MiB write) between timed steps"},
        "clocks": clocks.summary(), "e2e": e2e,
        "gpu_launches": None,
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "K3 staged pipeline (deg %d, sample %d, %d peaks x ~%d candidates)"
                                                % (deg, setup.sample_size, n_u, max(1, n_c // max(1, n_u))),
                     "flops_per_hypothesis": F, "f_intercept_pass": f_i, "f_scored": f_m,
                     "peak_source": "DFMA microbenchmark in this run (rb2_dfma_burn)", "kernel_ms_per_step": kern_ms / args.steps},
        "stages_resident": {"k1_hough_ms": k1_ms, "k1_evaluations_per_s": S * P_pairs / (k1_ms * 1e-3),
                            "k2_vote_ms": k2_ms, "k2_bin_pair_tests_per_s": float(B) * P_pairs / (k2_ms * 1e-3),
                            "hough_lines": B, "pairs": P_pairs, "slopes": S},
        "counters": counters,
    }
    n_chunks = k3_chunk_plan(lib, hi - lo)[0]
    line["gpu_launches"] = (2 + 8 * n_chunks + (1 if world > 1 else 0) + 1) * args.steps
    if not args.no_cpu and rank == 0:
        line["cpu_baseline"] = cpu_baseline(name)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hyp", type=float, default=1e8, help="hypotheses per step over all ranks")
    ap.add_argument("--config", default="c3", choices=["c3", "c2", "c5"],
                    help="BASELINE config: c3 (headline, default), c2 (GMOS deg 5, 1e5 tries), c5 (DEIMOS 1000x1000 bins, 1e6 tries)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="fits in the e2e leg (0 = --steps)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--spectra", type=int, default=2048, help="C4 spectra per rank in the many-spectra leg (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    # keep fd 1 for the ONE JSON line: libraries (NCCL's version banner) write to stdout too
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = None if world > 1 else False

    import __graft_entry__

    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from rascal_b200 import _native, engine, parallel
    from rascal_b200.calibrator import Calibrator, LineList

    lib = _native.load()
    if args.config != "c3":
        line = config_bench(args, torch, dist, rank, world, local, group)
        if rank == 0:
            os.write(json_fd, (json.dumps(line) + "\n").encode())
        if world > 1:
            dist.destroy_process_group()
        return
    w = c3_input()
    H = int(args.hyp)

    def new_calibrator(seed):
        c = Calibrator(w["peaks"], spectrum=np.zeros(w["npix"]))
        c.set_calibrator_properties(seed=seed)
        c.set_hough_properties(**w["hough"])
        c.set_ransac_properties(**w["ransac"])
        c.set_atlas(LineList(w["atlas"]), candidate_tolerance=10.0)
        if world > 1:
            c.use_distributed(None)
        return c

    # ---- device-resident leg: K1 + K2 + set-up once, then K3 (+ exchange + K4) per step ----
    c = new_calibrator(0)
    c.do_hough_transform()
    c._vote(w["fit"]["candidate_tolerance"])
    setup = engine.RansacSetup(
        c.candidate_peak, c.candidate_arc, fit_deg=4, sample_size=5, ransac_tolerance=c.ransac_tolerance,
        min_intercept=c.min_intercept, max_intercept=c.max_intercept, pixel_list=c.pixel_list, hist=c.ht.hist,
        xedges=c.ht.xedges, yedges=c.ht.yedges, hough_weight=c.hough_weight, n_peaks_total=len(c.peaks))
    rb = engine.RansacBatch([setup])
    lo, hi = parallel.shard_range(H, rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step(k):
        rb.run(k * H + lo, hi - lo, seed=2026)  # K3 main + per-problem finalize
        rb.exchange(group)                      # NCCL all-gather + device merge (no-op at N=1)
        rb.refit_async()                        # K4 on the device-resident winner
        return rb.fetch()                       # the step's single D2H copy

    for k in range(args.warmup):
        step(k)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    # ---- (1) one fit after the other: the latency of a step (K3 -> exchange -> K4 -> D2H, the host waits for the
    # result before the next step starts).  Reported as `sequential`. ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (inputs are KBs, far below L2)
        ev[k][0].record()
        rb.run((args.warmup + k) * H + lo, hi - lo, seed=2026)
        rb.exchange(group)
        rb.refit_async()
        rb.fetch()
        ev[k][1].record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    seq_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(seq_ms, op=dist.ReduceOp.MAX)
    seq_ms = float(seq_ms.item())

    # ---- (2) the timed job: the same K steps back to back, as a service that calibrates one spectrum after the other
    # runs them - K3 of step k on the main stream; the exchange, K4 and the D2H of step k on a side stream, where they
    # overlap K3 of step k + 1 (K4 is ONE warp: a dependent chain that no rank count shrinks).  Two sets of result
    # buffers alternate; every step's result is on the host before the closing event.  `value` = K * H / this time. ----
    rbs = [rb, engine.RansacBatch([setup])]
    side = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    for k in range(2):  # warm the second buffer set
        rbs[1].run(k * H + lo, hi - lo, seed=2026)
        rbs[1].exchange(group)
        rbs[1].refit_async()
        rbs[1].fetch()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    counters = dict.fromkeys(_native.COUNTER_NAMES, 0)
    results = []

    def collect(r, done):
        win, fit, mx, my, rs = r.fetch_wait(done)
        results.append((win[0].hyp_id, fit[0].accepted))
        for name, v in engine.counters_dict(win[0]).items():
            counters[name] += v

    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        t_begin.record()
        pending = None
        for k in range(args.steps):
            r = rbs[k % 2]
            flush.zero_()  # L2 flush between timed iterations
            kev[k][0].record()
            r.run((args.warmup + k) * H + lo, hi - lo, seed=2026)  # K3 (main stream)
            kev[k][1].record()
            k3_done = torch.cuda.Event()
            k3_done.record()
            with torch.cuda.stream(side):
                side.wait_event(k3_done)
                r.exchange(group)          # NCCL all-gather + device merge (no-op at N=1)
                r.refit_async()            # K4
                done = r.fetch_async()     # the step's single D2H copy
            if pending is not None:
                collect(*pending)          # step k - 1: it finished while K3 of step k was queued / running
            pending = (r, done)
        collect(*pending)
        main_stream.wait_stream(side)
        t_end.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
    assert len(results) == args.steps
    ms_kernel = [a.elapsed_time(b) for a, b in kev]
    total_ms = torch.tensor([t_begin.elapsed_time(t_end)], dtype=torch.float64, device="cuda")
    kern_ms = torch.tensor([sum(ms_kernel)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern_ms, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = float(total_ms.item()), float(kern_ms.item())
    value = args.steps * H / (total_ms * 1e-3)
    sequential = {"value": args.steps * H / (seq_ms * 1e-3), "unit": UNIT, "ms_per_step": seq_ms / args.steps,
                  "what": "the same K steps one after the other (the host waits for a step's result before it starts "
                          "the next): latency of one resident fit"}

    # ---- FP64 roofline denominator (DFMA microbenchmark, same box, same run) ----
    sms = rb.sm_count
    fp64_peak = fp64_peak_tflops(lib, engine, torch, sms)

    # ---- e2e leg: the public Calibrator API with host buffers, every step ----
    e2e = None
    e2e_steps = max(1, min(args.e2e_steps or args.steps, args.steps))
    for k in range(2):  # warm the public path the timed loop takes (allocations, spline operator, NCCL)
        wc = new_calibrator(5 + k)
        wc.do_hough_transform()
        wc.fit(progress=False, n_hypotheses=max(world * 1000, H // 100), **w["fit"])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    engine.TRANSFER["h2d"] = engine.TRANSFER["d2h"] = 0
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        cal = new_calibrator(100 + k)
        cal.do_hough_transform()
        res = cal.fit(progress=False, n_hypotheses=H, **w["fit"])
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_e2e = time.perf_counter() - t0
    t_e2e_t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e_t, op=dist.ReduceOp.MAX)
    t_e2e = float(t_e2e_t.item())
    e2e = {"value": e2e_steps * H / t_e2e, "unit": UNIT,
           "h2d_bytes_per_step": engine.TRANSFER["h2d"] // e2e_steps,
           "d2h_bytes_per_step": engine.TRANSFER["d2h"] // e2e_steps,
           "ms_per_step": 1e3 * t_e2e / e2e_steps, "rms": float(res[3]), "n_matched": len(res[1])}

    # ---- the UNCHANGED signature: do_hough_transform() + fit(max_tries=500) as every example of the reference calls
    # it (its own runtime for this call: 3-8 s) ----
    api_ms, api_cal = [], None
    for k in range(4):
        api_cal = new_calibrator(200 + k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        api_cal.do_hough_transform()
        api_res = api_cal.fit(max_tries=500, progress=False, **w["fit"])
        torch.cuda.synchronize()
        api_ms.append(1e3 * (time.perf_counter() - t0))
    default_api = {"call": "do_hough_transform() + fit(max_tries=500, fit_deg=4) from host arrays",
                   "ms_per_calibration": float(np.median(api_ms[1:])), "completed_tries": 500,
                   "hypotheses_drawn": api_cal.ransac_counters["drawn"] + api_cal.ransac_counters["aborted"],
                   "rms": float(api_res[3]), "n_matched": len(api_res[1])}

    # ---- many-spectra leg (BASELINE configs[3] shape, bounded sample): spectra calibrated/s through
    # rascal_b200.batch.calibrate_batch (K1..K5 per spectrum, host buffers in, results out) ----
    spectra = None
    if args.spectra > 0:
        from rascal_b200 import batch

        specs = [c4_input(rank + world * i) for i in range(args.spectra)]  # spectrum i -> rank i mod world
        batch.calibrate_batch(specs, n_hypotheses=10_000, seed=1, match_tolerance=5.0)  # warm (untimed)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        tm = {}
        t0 = time.perf_counter()
        bres = batch.calibrate_batch(specs, n_hypotheses=10_000, seed=1, match_tolerance=5.0, timings=tm)
        # the final gather - every rank ends up with the fixed-size result records of ALL spectra
        records = parallel.gather_records(parallel.result_records(bres), world * args.spectra, group)
        torch.cuda.synchronize()
        t_b = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_b, op=dist.ReduceOp.MAX)
        all_specs = specs if world == 1 else [c4_input(i) for i in range(world * args.spectra)]
        good = sum(bool(rec[0]) and truth_rms(sp, rec[6:11]) < 1.0 for sp, rec in zip(all_specs, records))
        rate = world * args.spectra / float(t_b.item())
        spectra = {"value": rate, "unit": "spectra/s",
                   "sample": "%d C4 spectra per rank x 1e4 hypotheses each, Hough 2000x(100x100), match_peaks after the fit; "
                             "host NumPy arrays in, per-spectrum results out (device-resident pipeline), records of all "
                             "ranks gathered" % args.spectra,
                   "fitted": int(records[:, 0].sum()), "within_1A_of_truth": int(good),
                   # most C4 spectra defeat the reference's algorithm itself (below): the rate of CORRECT solutions
                   "correct_spectra_per_s": rate * good / max(1, len(records)),
                   "stage_s_rank0": {k: round(v, 4) for k, v in tm.items()}}
        if rank == 0:  # success rate against the draw budget (256 spectra), and the reference's own on the same recipe
            sub = [c4_input(i) for i in range(256)]
            sweep = {}
            for budget in (10_000, 100_000, 1_000_000):
                t0 = time.perf_counter()
                rs_ = batch.calibrate_batch(sub, n_hypotheses=budget, seed=1)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                ok = sum(r.fit_coeff is not None and truth_rms(sp, r.fit_coeff) < 1.0 for sp, r in zip(sub, rs_))
                sweep["%.0e" % budget] = {"within_1A": int(ok), "of": len(sub), "spectra_per_s": len(sub) / dt}
            spectra["quality_vs_draw_budget"] = sweep
            if not args.no_cpu:
                spectra["reference_quality"] = c4_reference_quality()

    # ---- K1 / K2 alone on resident inputs (rank 0): the Hough pass against SURVEY 8d's FP64-pipe roof ----
    hough_roof = None
    if args.spectra > 0 and rank == 0:
        from rascal_b200 import batch as _b

        ns = [_b.normalise_spec(sp) for sp in specs[:512]]
        hp, vs = [], []
        for sp in ns:
            lo_, hi_, mi_, ma_ = sp["limits"]
            h_ = sp["hough"]
            hp.append(dict(pair_x=np.repeat(sp["peaks"], len(sp["lines"])), pair_y=np.tile(sp["lines"], len(sp["peaks"])),
                           min_slope=lo_, max_slope=hi_, min_intercept=mi_, max_intercept=ma_,
                           n_slopes=int(h_["num_slopes"]), xbins=int(h_["xbins"]), ybins=int(h_["ybins"])))
            sg = (h_["range_tolerance"] + h_["linearity_tolerance"]) * 1.1775
            vs.append(dict(pairs=None, peaks=sp["peaks"], lines=sp["lines"], top_n=sp["ransac"]["top_n_candidate"],
                           weighted=sp["ransac"]["candidate_weighted"], tol=sp["fit"]["candidate_tolerance"],
                           gauss_denom=2 * sg ** 2 + 1e-9))
        hb = engine.HoughBatch(hp)
        hb.run()
        vb = engine.VoteBatch(hb, vs)
        vb.run()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        hb.run()
        e1.record()
        vb.run()
        e2.record()
        torch.cuda.synchronize()
        evals = float(sum(int(h_["n_slopes"]) * len(h_["pair_x"]) for h_ in hp))
        points = float(sum(int(st.n_points) for st in hb.stats))
        t1, t2 = e0.elapsed_time(e1) * 1e-3, e1.elapsed_time(e2) * 1e-3
        roof = fp64_peak * 1e12 / 2.0 / 4.0  # 4 FP64-pipe instructions per (slope, pair) evaluation
        hough_roof = {"kernel": "K1 rb2_hough_accumulate (interval + edges + bin + rank), 512 C4 spectra resident",
                      "evaluations_per_s": evals / t1, "surviving_points_per_s": points / t1,
                      "bound": "fp64", "peak_evaluations_per_s": roof, "frac": evals / t1 / roof, "ms": t1 * 1e3,
                      "algorithmic_bytes": int(sum(16 * len(h_["pair_x"]) + 8 * int(h_["n_slopes"]) + 8 * 202 + 8 * 10000 for h_ in hp)),
                      "k2_vote_ms": t2 * 1e3,
                      "k2_bin_pair_tests_per_s": float(sum(10000 * len(h_["pair_x"]) for h_ in hp)) / t2}

    # ---- K6 (the step before the path): find_peaks + refine_peaks for 512 rendered C4 arcs (rank 0) ----
    peak_leg = None
    if args.spectra > 0 and rank == 0:
        from rascal_b200 import util as rutil

        n_arc = min(512, args.spectra)
        arcs = [c4_arc(specs[i], i) for i in range(n_arc)]
        recipe = dict(height=150.0, distance=3, prominence=100.0)
        rutil.find_and_refine(arcs, window_width=5, max_peaks=256, **recipe)  # warm
        pbatch = rutil.PeakBatch(arcs)
        pbatch.find(max_peaks=256, **recipe).refine(5)
        torch.cuda.synchronize()
        reps = []
        for _ in range(5):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            pbatch.find(max_peaks=256, **recipe)
            e1.record()
            pbatch.refine(5)
            e2.record()
            torch.cuda.synchronize()
            reps.append((e0.elapsed_time(e2), e0.elapsed_time(e1), e1.elapsed_time(e2)))
        t_all, t_find, t_refine = sorted(reps)[2]  # median repetition
        refined = pbatch.fetch_refined()
        t0 = time.perf_counter()
        rutil.find_and_refine(arcs, window_width=5, max_peaks=256, **recipe)
        t_pk = time.perf_counter() - t0
        n_ref = sum(len(r) for r in refined)
        off_px = np.concatenate([np.min(np.abs(r[:, None] - specs[i]["peaks"][None, :]), axis=1) for i, r in enumerate(refined) if len(r)])
        peak_leg = {"kernel": "K6 rb2_find_peaks (1 launch) + rb2_refine_peaks (4 launches: windows, queue, fits, collect), %d C4 arcs resident" % n_arc,
                    "find_ms": t_find, "refine_ms": t_refine, "refined_peaks": n_ref, "timing": "median of 5 repetitions, CUDA events",
                    "peaks_per_s": n_ref / (t_all * 1e-3), "spectra_per_s": n_arc / (t_all * 1e-3),
                    "e2e_spectra_per_s": n_arc / t_pk, "median_offset_from_true_centre_px": float(np.median(off_px)),
                    "algorithmic_bytes": int(sum(8 * len(a) for a in arcs) + 8 * n_ref)}
        # the whole chain from the arcs: K6 -> K1..K5 through calibrate_batch(spectrum=...)
        arc_specs = [dict(specs[i], peaks=None, spectrum=arcs[i], find_peaks=recipe, window_width=5, max_peaks=256)
                     for i in range(n_arc)]
        batch.calibrate_batch(arc_specs, n_hypotheses=10_000, seed=1, match_tolerance=5.0)  # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ares = batch.calibrate_batch(arc_specs, n_hypotheses=10_000, seed=1, match_tolerance=5.0)
        torch.cuda.synchronize()
        t_arc = time.perf_counter() - t0
        peak_leg["arcs_to_solutions_per_s"] = n_arc / t_arc
        peak_leg["arcs_fitted"] = sum(r.fit_coeff is not None for r in ares)
        if not args.no_cpu:
            from scipy.signal import find_peaks as sp_find

            from oracle import peaks as opeaks

            t0 = time.perf_counter()
            n_cpu = 0
            for a in arcs[:16]:
                pk, _ = sp_find(a, **recipe)
                n_cpu += len(opeaks.refine_peaks(a, pk, window_width=5, fitter="scipy"))
            t_cpu = time.perf_counter() - t0
            peak_leg["cpu_baseline"] = {"spectra_per_s": 16 / t_cpu, "peaks_per_s": n_cpu / t_cpu, "cores": 1, "kind": "port",
                                        "sample": "16 of the arcs: scipy.signal.find_peaks + util.refine_peaks restated over scipy.optimize.curve_fit"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    F, f_i, f_m = flops_per_hypothesis(counters)
    per_rank_hyp = (hi - lo)
    n_chunks, n_lanes, chunk_ids = k3_chunk_plan(lib, per_rank_hyp)
    achieved = per_rank_hyp * args.steps * F / (kern_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "hypotheses_per_step": H, "parallelism": "hypothesis-id ranges x%d, %d chunks of %d ids over %d stream lanes per GPU" % (world, n_chunks, chunk_ids, n_lanes),
                   "l2": "flushed (256 MiB write) between timed steps", "philox_seed": 2026,
                   "steps": "back to back: exchange + K4 + D2H of step k on a side stream overlap K3 of step k+1; "
                            "`sequential` = one step at a time"},
        "clocks": clocks.summary(),
        "e2e": e2e,
        "sequential": sequential,
        # K3 = pipe_init + 8 kernels (reset, draw_fit, monotonic, match, cost, general weight, exact, record) per
        # chunk + pipe_join; + merge after the all-gather (N > 1) + K4 refit
        "gpu_launches": (2 + 8 * n_chunks + (1 if world > 1 else 0) + 1) * args.steps,
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp64_peak, "traffic": k3_traffic(),
                     "kernel": "K3 staged pipeline: ransac_draw_fit_kernel<4> (dominant: 71 % of K3's instructions) + "
                               "monotonic + match + cost + exact + record, timed together per step; traffic = DRAM "
                               "bytes of one chunk's six kernels (profiles/k3_traffic.json: hypotheses_per_chunk; ncu, caches flushed)",
                     "issue_slots_busy_pct": k3_profile("issue_slots_busy_pct"),
                     "flops_per_hypothesis": F, "f_intercept_pass": f_i,
                     "f_scored": f_m, "peak_source": "DFMA microbenchmark in this run (rb2_dfma_burn); FP64 is not in MEASURED_PEAKS.json",
                     "kernel_ms_per_step": kern_ms / args.steps},
        "counters": counters,
        "wall_s_timed_region": t_wall,
    }
    line["default_api"] = default_api
    if spectra is not None:
        line["spectra_per_sec"] = spectra
    if hough_roof is not None:
        line["hough_roofline"] = hough_roof
    if peak_leg is not None:
        line["peak_finding"] = peak_leg
    if not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline("c3")
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
