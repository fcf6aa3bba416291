"""CPU oracle for the step BEFORE the hot path: peak finding
and sub-pixel refinement.  TEST INFRASTRUCTURE ONLY - imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg, never by rascal_b200.

What is restated, and from where:

* ``refine_peaks``  - rascal ``util.refine_peaks`` (src/rascal/util.py:450-523) and
  ``util.gauss`` (:425-447).  The reference calls ``scipy.optimize.curve_fit``
  without bounds, i.e. MINPACK's ``lmdif`` (Levenberg-Marquardt, forward-difference
  Jacobian) through ``scipy.optimize.leastsq`` with ftol = xtol = 1.49012e-8,
  gtol = 0, maxfev = 200*(n+1), factor = 100, automatic scaling.
* ``lmdif`` / ``lmpar`` / ``qrfac`` / ``qrsolv`` / ``enorm`` - the published MINPACK-1
  algorithm (More, Garbow, Hillstrom 1980), a third-party dependency of the
  reference (SciPy >= 1.4; 1.18.1 in this image).  Pinned against
  ``scipy.optimize.leastsq`` in tests/test_oracle_peaks.py (same iterates, same
  ``nfev``, same ``ier``).
* ``find_peaks`` - the subset of ``scipy.signal.find_peaks`` every reference example
  and test uses before ``refine_peaks`` (test/test_xe_calibration.py:69,
  examples/*.py): ``height``, ``distance``, ``prominence``; ``threshold=None``.
  SciPy's algorithm (``_local_maxima_1d``, ``_select_by_peak_distance``,
  ``_peak_prominences`` in scipy/signal/_peak_finding_utils.pyx) is restated and
  pinned against ``scipy.signal.find_peaks`` itself.

Parity pinned: recordings of the reference's ``util.refine_peaks`` on synthetic and
real arcs live in tests/golden/peaks_*.npz (generator: make_golden_peaks.py).
"""
from __future__ import annotations

import math

import numpy as np

EPSMCH = np.finfo(np.float64).eps
LM_TOL = 1.49012e-8  # scipy.optimize.leastsq defaults (ftol, xtol)


# --------------------------------------------------------------------------- #
# MINPACK-1 lmdif
# --------------------------------------------------------------------------- #
_RDWARF, _RGIANT = 3.834e-20, 1.304e19


def enorm(v):
    """MINPACK enorm: Euclidean norm with separate accumulators for small,
    intermediate and large components."""
    v = np.asarray(v, dtype=np.float64)
    n = v.size
    if n == 0:
        return 0.0
    s1 = s2 = s3 = 0.0
    x1max = x3max = 0.0
    agiant = _RGIANT / n
    for x in v:
        xabs = abs(x)
        if xabs > _RDWARF and xabs < agiant:
            s2 += xabs * xabs
        elif xabs <= _RDWARF:
            if xabs > x3max:
                s3 = 1.0 + s3 * (x3max / xabs) ** 2
                x3max = xabs
            elif xabs != 0.0:
                s3 += (xabs / x3max) ** 2
        else:
            if xabs > x1max:
                s1 = 1.0 + s1 * (x1max / xabs) ** 2
                x1max = xabs
            else:
                # (a NaN component lands here with x1max still 0: IEEE arithmetic, as in the Fortran / C original,
                #  gives NaN - Python's float division would raise instead)
                s1 += (xabs / x1max) ** 2 if x1max != 0.0 else float("nan")
    if s1 != s1:  # a NaN component: every IEEE evaluation order of the line below gives NaN
        return float("nan")
    if s1 != 0.0:
        return x1max * math.sqrt(s1 + (s2 / x1max) / x1max)
    if s2 != 0.0:
        if s2 >= x3max:
            return math.sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)))
        return math.sqrt(x3max * ((s2 / x3max) + (x3max * s3)))
    return x3max * math.sqrt(s3)


def qrfac(a):
    """Householder QR with column pivoting, in place on a[m, n].
    Returns (ipvt, rdiag, acnorm)."""
    m, n = a.shape
    acnorm = np.array([enorm(a[:, j]) for j in range(n)])
    rdiag = acnorm.copy()
    wa = rdiag.copy()
    ipvt = np.arange(n)
    for j in range(min(m, n)):
        kmax = j
        for k in range(j, n):
            if rdiag[k] > rdiag[kmax]:
                kmax = k
        if kmax != j:
            a[:, [j, kmax]] = a[:, [kmax, j]]
            rdiag[kmax] = rdiag[j]
            wa[kmax] = wa[j]
            ipvt[j], ipvt[kmax] = ipvt[kmax], ipvt[j]
        ajnorm = enorm(a[j:, j])
        if ajnorm != 0.0:
            if a[j, j] < 0.0:
                ajnorm = -ajnorm
            a[j:, j] /= ajnorm
            a[j, j] += 1.0
            for k in range(j + 1, n):
                s = 0.0
                for i in range(j, m):
                    s += a[i, j] * a[i, k]
                temp = s / a[j, j]
                for i in range(j, m):
                    a[i, k] -= temp * a[i, j]
                if rdiag[k] != 0.0:
                    temp = a[j, k] / rdiag[k]
                    rdiag[k] *= math.sqrt(max(0.0, 1.0 - temp * temp))
                    if 0.05 * (rdiag[k] / wa[k]) ** 2 <= EPSMCH:
                        rdiag[k] = enorm(a[j + 1:, k])
                        wa[k] = rdiag[k]
        rdiag[j] = -ajnorm
    return ipvt, rdiag, acnorm


def qrsolv(r, ipvt, diag, qtb):
    """Solve min ||A x - b||^2 + ||D x||^2 given the pivoted QR of A.
    r[n, n]: upper triangle = R (its strict lower triangle is overwritten with
    the strict upper triangle of S^T).  Returns (x, sdiag)."""
    n = r.shape[0]
    x = np.zeros(n)
    sdiag = np.zeros(n)
    wa = np.zeros(n)
    for j in range(n):
        for i in range(j, n):
            r[i, j] = r[j, i]
        x[j] = r[j, j]
        wa[j] = qtb[j]
    for j in range(n):
        l = ipvt[j]
        if diag[l] != 0.0:
            sdiag[j:] = 0.0
            sdiag[j] = diag[l]
            qtbpj = 0.0
            for k in range(j, n):
                if sdiag[k] == 0.0:
                    continue
                if abs(r[k, k]) < abs(sdiag[k]):
                    cotan = r[k, k] / sdiag[k]
                    sin = 0.5 / math.sqrt(0.25 + 0.25 * cotan * cotan)
                    cos = sin * cotan
                else:
                    tan = sdiag[k] / r[k, k]
                    cos = 0.5 / math.sqrt(0.25 + 0.25 * tan * tan)
                    sin = cos * tan
                r[k, k] = cos * r[k, k] + sin * sdiag[k]
                temp = cos * wa[k] + sin * qtbpj
                qtbpj = -sin * wa[k] + cos * qtbpj
                wa[k] = temp
                for i in range(k + 1, n):
                    temp = cos * r[i, k] + sin * sdiag[i]
                    sdiag[i] = -sin * r[i, k] + cos * sdiag[i]
                    r[i, k] = temp
        sdiag[j] = r[j, j]
        r[j, j] = x[j]
    nsing = n
    for j in range(n):
        if sdiag[j] == 0.0 and nsing == n:
            nsing = j
        if nsing < n:
            wa[j] = 0.0
    for k in range(nsing):
        j = nsing - 1 - k
        s = 0.0
        for i in range(j + 1, nsing):
            s += r[i, j] * wa[i]
        wa[j] = (wa[j] - s) / sdiag[j]
    for j in range(n):
        x[ipvt[j]] = wa[j]
    return x, sdiag


_DWARF = np.finfo(np.float64).tiny


def lmpar(r, ipvt, diag, qtb, delta, par):
    """Levenberg-Marquardt parameter: returns (par, x, sdiag)."""
    n = r.shape[0]
    wa1 = np.zeros(n)
    nsing = n
    for j in range(n):
        wa1[j] = qtb[j]
        if r[j, j] == 0.0 and nsing == n:
            nsing = j
        if nsing < n:
            wa1[j] = 0.0
    for k in range(nsing):
        j = nsing - 1 - k
        wa1[j] /= r[j, j]
        temp = wa1[j]
        for i in range(j):
            wa1[i] -= r[i, j] * temp
    x = np.zeros(n)
    for j in range(n):
        x[ipvt[j]] = wa1[j]
    wa2 = diag * x
    dxnorm = enorm(wa2)
    fp = dxnorm - delta
    if fp <= 0.1 * delta:
        return 0.0, x, np.zeros(n)
    parl = 0.0
    if nsing >= n:
        for j in range(n):
            l = ipvt[j]
            wa1[j] = diag[l] * (wa2[l] / dxnorm)
        for j in range(n):
            s = 0.0
            for i in range(j):
                s += r[i, j] * wa1[i]
            wa1[j] = (wa1[j] - s) / r[j, j]
        temp = enorm(wa1)
        parl = ((fp / delta) / temp) / temp
    for j in range(n):
        s = 0.0
        for i in range(j + 1):
            s += r[i, j] * qtb[i]
        wa1[j] = s / diag[ipvt[j]]
    gnorm = enorm(wa1)
    paru = gnorm / delta
    if paru == 0.0:
        paru = _DWARF / min(delta, 0.1)
    par = max(par, parl)
    par = min(par, paru)
    if par == 0.0:
        par = gnorm / dxnorm
    it = 0
    sdiag = np.zeros(n)
    while True:
        it += 1
        if par == 0.0:
            par = max(_DWARF, 0.001 * paru)
        temp = math.sqrt(par)
        wa1 = temp * diag
        x, sdiag = qrsolv(r, ipvt, wa1, qtb)
        wa2 = diag * x
        dxnorm = enorm(wa2)
        temp = fp
        fp = dxnorm - delta
        if abs(fp) <= 0.1 * delta or (parl == 0.0 and fp <= temp and temp < 0.0) or it == 10:
            break
        for j in range(n):
            l = ipvt[j]
            wa1[j] = diag[l] * (wa2[l] / dxnorm)
        for j in range(n):
            wa1[j] /= sdiag[j]
            temp = wa1[j]
            for i in range(j + 1, n):
                wa1[i] -= r[i, j] * temp
        temp = enorm(wa1)
        parc = ((fp / delta) / temp) / temp
        if fp > 0.0:
            parl = max(parl, par)
        if fp < 0.0:
            paru = min(paru, par)
        par = max(parl, par + parc)
    return par, x, sdiag


def lmdif(fcn, x0, ftol=LM_TOL, xtol=LM_TOL, gtol=0.0, maxfev=None, epsfcn=EPSMCH, factor=100.0):
    """MINPACK lmdif (mode 1).  fcn(x) -> residual vector [m].  Returns (x, info, nfev)."""
    x = np.array(x0, dtype=np.float64)
    n = x.size
    if maxfev is None:
        maxfev = 200 * (n + 1)
    fvec = np.asarray(fcn(x), dtype=np.float64).copy()
    m = fvec.size
    nfev = 1
    fnorm = enorm(fvec)
    par = 0.0
    it = 1
    diag = np.ones(n)
    xnorm = delta = 0.0
    eps = math.sqrt(max(epsfcn, EPSMCH))
    info = 0
    while True:
        fjac = np.empty((m, n))
        for j in range(n):
            temp = x[j]
            h = eps * abs(temp)
            if h == 0.0:
                h = eps
            x[j] = temp + h
            wa = np.asarray(fcn(x), dtype=np.float64)
            x[j] = temp
            fjac[:, j] = (wa - fvec) / h
        nfev += n
        ipvt, wa1, wa2 = qrfac(fjac)
        if it == 1:
            diag = np.where(wa2 == 0.0, 1.0, wa2)
            xnorm = enorm(diag * x)
            delta = factor * xnorm
            if delta == 0.0:
                delta = factor
        wa4 = fvec.copy()
        qtf = np.zeros(n)
        for j in range(n):
            if fjac[j, j] != 0.0:
                s = 0.0
                for i in range(j, m):
                    s += fjac[i, j] * wa4[i]
                temp = -s / fjac[j, j]
                for i in range(j, m):
                    wa4[i] += fjac[i, j] * temp
            fjac[j, j] = wa1[j]
            qtf[j] = wa4[j]
        gnorm = 0.0
        if fnorm != 0.0:
            for j in range(n):
                l = ipvt[j]
                if wa2[l] != 0.0:
                    s = 0.0
                    for i in range(j + 1):
                        s += fjac[i, j] * (qtf[i] / fnorm)
                    gnorm = max(gnorm, abs(s / wa2[l]))
        if gnorm <= gtol:
            info = 4
            break
        diag = np.maximum(diag, wa2)
        r = fjac[:n, :n]
        while True:
            par, step, _ = lmpar(r, ipvt, diag, qtf, delta, par)
            wa1 = -step
            wa2n = x + wa1
            wa3 = diag * wa1
            pnorm = enorm(wa3)
            if it == 1:
                delta = min(delta, pnorm)
            wa4 = np.asarray(fcn(wa2n), dtype=np.float64).copy()
            nfev += 1
            fnorm1 = enorm(wa4)
            actred = -1.0
            if 0.1 * fnorm1 < fnorm:
                actred = 1.0 - (fnorm1 / fnorm) ** 2
            wa3 = np.zeros(n)
            for j in range(n):
                temp = wa1[ipvt[j]]
                for i in range(j + 1):
                    wa3[i] += r[i, j] * temp
            temp1 = enorm(wa3) / fnorm
            temp2 = (math.sqrt(par) * pnorm) / fnorm
            prered = temp1 * temp1 + temp2 * temp2 / 0.5
            dirder = -(temp1 * temp1 + temp2 * temp2)
            ratio = actred / prered if prered != 0.0 else 0.0
            if ratio <= 0.25:
                temp = 0.5 if actred >= 0.0 else 0.5 * dirder / (dirder + 0.5 * actred)
                if 0.1 * fnorm1 >= fnorm or temp < 0.1:
                    temp = 0.1
                delta = temp * min(delta, pnorm / 0.1)
                par /= temp
            elif par == 0.0 or ratio >= 0.75:
                delta = pnorm / 0.5
                par *= 0.5
            if ratio >= 1e-4:
                x = wa2n
                fvec = wa4
                xnorm = enorm(diag * x)
                fnorm = fnorm1
                it += 1
            if abs(actred) <= ftol and prered <= ftol and 0.5 * ratio <= 1.0:
                info = 1
            if delta <= xtol * xnorm:
                info = 2
            if abs(actred) <= ftol and prered <= ftol and 0.5 * ratio <= 1.0 and info == 2:
                info = 3
            if info != 0:
                break
            if nfev >= maxfev:
                info = 5
            if abs(actred) <= EPSMCH and prered <= EPSMCH and 0.5 * ratio <= 1.0:
                info = 6
            if delta <= EPSMCH * xnorm:
                info = 7
            if gnorm <= EPSMCH:
                info = 8
            if info != 0:
                break
            if ratio >= 1e-4:
                break
        if info != 0:
            break
    return x, info, nfev


# --------------------------------------------------------------------------- #
# util.refine_peaks
# --------------------------------------------------------------------------- #
def gauss(x, a, x0, sigma):
    """util.py:425-447."""
    return a * np.exp(-((x - x0) ** 2) / (2 * sigma**2 + 1e-9))


def refine_peaks(spectrum, peaks, window_width=10, fitter="lmdif", info=None, exp=np.exp):
    """util.py:450-523.  `info`: list that receives one record per fitted window (tests).
    `exp`: the exponential used inside the model (tests perturb its last bit to measure how far
    the reference's own answer moves with the rounding of libm's exp).  `fitter`: "lmdif" (the restatement above) or "scipy"
    (scipy.optimize.curve_fit, what the reference calls).  Quirks kept on purpose:
    the window is normalised IN PLACE in the working copy of the spectrum (:478),
    so a later peak whose window overlaps sees the already-divided values; `mean`
    and `sigma` of the start are sums over y, not normalised moments (:487-488);
    the result is `peak - window_width + centre` even when the window was clipped
    at pixel 0 (:499); `centre` is compared with len(spectrum) (:496)."""
    spectrum = np.array(spectrum, dtype=np.float64)
    length = len(spectrum)
    refined = []
    for peak in peaks:
        lo = max(0, int(peak) - window_width)
        hi = min(int(peak) + window_width, length)
        y = spectrum[lo:hi]
        y /= np.nanmax(y)
        x = np.arange(len(y))
        mask = np.isfinite(y)
        n = int(mask.sum())
        if n == 0:
            continue
        mean = np.sum(x * y) / n
        sigma = np.sum(y * (x - mean) ** 2) / n
        p0 = [1.0, mean, sigma]
        xm, ym = x[mask].astype(np.float64), y[mask]
        rec = dict(peak=peak, lo=lo, y=ym.copy(), p0=p0, ier=None, nfev=None, popt=None, appended=False)
        if info is not None:
            info.append(rec)
        if fitter == "scipy":
            from scipy.optimize import curve_fit
            try:
                popt, _ = curve_fit(gauss, xm, ym, p0=p0)
            except RuntimeError:
                rec["ier"] = -1
                continue
        else:
            popt, ier, nfev = lmdif(lambda p: p[0] * exp(-((xm - p[1]) ** 2) / (2 * p[2]**2 + 1e-9)) - ym, p0)
            rec.update(ier=ier, nfev=nfev)
            if ier not in (1, 2, 3, 4):
                continue
        rec["popt"] = np.array(popt)
        height, centre, _ = popt
        if height < 0:
            continue
        if centre > length or centre < 0:
            continue
        rec["appended"] = True
        refined.append(peak - window_width + centre)
    refined = np.array(refined, dtype=np.float64)
    mask = (refined > 0) & (refined < length)
    distance_mask = np.isclose(refined[:-1], refined[1:])
    distance_mask = np.insert(distance_mask, 0, False)
    return refined[mask & ~distance_mask]


# --------------------------------------------------------------------------- #
# scipy.signal.find_peaks (height / distance / prominence)
# --------------------------------------------------------------------------- #
def local_maxima_1d(x):
    """Midpoints of all strict local maxima, flat tops included (the middle
    sample, rounded down); the first and last sample are never maxima."""
    x = np.asarray(x, dtype=np.float64)
    mid = []
    i, i_max = 1, x.size - 1
    while i < i_max:
        if x[i - 1] < x[i]:
            ahead = i + 1
            while ahead < i_max and x[ahead] == x[i]:
                ahead += 1
            if x[ahead] < x[i]:
                mid.append((i + ahead - 1) // 2)
                i = ahead
        i += 1
    return np.array(mid, dtype=np.intp)


def select_by_peak_distance(peaks, priority, distance):
    """Keep-mask: peaks are visited from the highest priority down, each kept peak
    removes every not-yet-visited peak closer than `distance` samples."""
    n = peaks.size
    keep = np.ones(n, dtype=bool)
    distance_ = math.ceil(distance)
    order = np.argsort(priority, kind="stable")
    for i in range(n - 1, -1, -1):
        j = order[i]
        if not keep[j]:
            continue
        k = j - 1
        while k >= 0 and peaks[j] - peaks[k] < distance_:
            keep[k] = False
            k -= 1
        k = j + 1
        while k < n and peaks[k] - peaks[j] < distance_:
            keep[k] = False
            k += 1
    return keep


def peak_prominences(x, peaks):
    """Prominence with wlen=None: height above the higher of the two lowest
    points between the peak and the next strictly higher sample on each side."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty(peaks.size)
    for k, p in enumerate(peaks):
        i = p
        left_min = x[p]
        while i >= 0 and x[i] <= x[p]:
            if x[i] < left_min:
                left_min = x[i]
            i -= 1
        i = p
        right_min = x[p]
        while i < x.size and x[i] <= x[p]:
            if x[i] < right_min:
                right_min = x[i]
            i += 1
        out[k] = x[p] - max(left_min, right_min)
    return out


def find_peaks(x, height=None, distance=None, prominence=None):
    """scipy.signal.find_peaks(x, height=h, distance=d, prominence=p) with scalar
    minima; returns (peaks, properties) in SciPy's order of conditions: height,
    distance, prominence."""
    x = np.asarray(x, dtype=np.float64)
    peaks = local_maxima_1d(x)
    props = {}
    if height is not None:
        peaks = peaks[x[peaks] >= height]
        props["peak_heights"] = x[peaks]
    if distance is not None:
        keep = select_by_peak_distance(peaks, x[peaks], distance)
        peaks = peaks[keep]
        props = {k: v[keep] for k, v in props.items()}
    if prominence is not None:
        prom = peak_prominences(x, peaks)
        keep = prom >= prominence
        peaks = peaks[keep]
        props = {k: v[keep] for k, v in props.items()}
        props["prominences"] = prom[keep]
    return peaks, props
