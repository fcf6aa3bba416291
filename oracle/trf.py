"""Step-for-step restatement of what models.robust_polyfit makes SciPy do.
TEST INFRASTRUCTURE ONLY (see oracle/rascal_oracle.py header).

The reference's robust refit (src/rascal/models.py:70-123) is

    scipy.optimize.least_squares(poly_cost_function, x0=ones(deg+1),
        args=(x/std(x), y/std(y), deg), loss="huber", diff_step=1e-5)

i.e. SciPy's Trust Region Reflective solver without bounds
(scipy 1.18.1, optimize/_lsq/trf.py::trf_no_bounds, tr_solver='exact',
x_scale=1, ftol=xtol=gtol=1e-8, max_nfev=100*n) on a forward-difference
Jacobian (optimize/_numdiff.py::approx_derivative, '2-point',
rel_step=1e-5).  SciPy is a third-party dependency that is not part of
/root/reference; this file restates its published algorithm in scalar loops
(own one-sided Jacobi SVD instead of LAPACK gesdd) exactly as the device
kernel K4 implements it, and tests/test_oracle_trf.py pins it against the
real scipy call on the golden inlier sets and on random problems.

The result of TRF is NOT the least-squares optimum: it is wherever the
iteration stops (typically after 10-30 function evaluations, 1e-8..1e-5
relative away in the coefficients); parity with the reference at 1e-9 needs
the same iteration path, hence this restatement.
"""
import numpy as np

EPS = np.finfo(float).eps


def power_int(x, i):
    """x**i as numpy evaluates it for a Python-int exponent (np.power -> libm pow,
    correctly rounded in all but vanishingly rare cases): correctly rounded here
    via exact rational-free double-double products."""
    if i == 0:
        return np.ones_like(x)
    if i == 1:
        return x.copy()
    if i == 2:
        return x * x  # np.square
    return np.array([_pow_dd(float(v), i) for v in x])


def _two_prod(a, b):
    p = a * b
    # Veltkamp split (no FMA in numpy); exact for doubles without overflow
    c = 134217729.0
    ah = a * c
    ah = ah - (ah - a)
    al = a - ah
    bh = b * c
    bh = bh - (bh - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def _pow_dd(x, n):
    hi, lo = x, 0.0
    for _ in range(n - 1):
        p, e = _two_prod(hi, x)
        e += lo * x
        s = p + e
        lo = e - (s - p)
        hi = s
    return hi + lo


def residuals(a, x, y, deg):
    """models.py:10-54 in its own operation order: y - (a[-1] + sum_i a[deg-i] * x**i)."""
    t = np.full_like(x, a[-1])
    for i in range(1, deg + 1):
        t = t + a[deg - i] * power_int(x, i)
    return y - t


def jacobian_2point(a, f0, x, y, deg, rel_step=1e-5):
    """_numdiff.approx_derivative(method='2-point', rel_step=1e-5), no bounds."""
    n = len(a)
    J = np.empty((len(x), n))
    for i in range(n):
        sign = 1.0 if a[i] >= 0 else -1.0
        h = rel_step * sign * abs(a[i])
        if (a[i] + h) - a[i] == 0:
            h = EPS ** 0.5 * sign * max(1.0, abs(a[i]))
        a1 = a.copy()
        a1[i] = a[i] + h
        dx = (a[i] + h) - a[i]
        J[:, i] = (residuals(a1, x, y, deg) - f0) / dx
    return J


def huber_rho(f):
    """least_squares.huber on z = f**2 (f_scale = 1): rho, rho', rho''."""
    z = f * f
    r0, r1, r2 = z.copy(), np.ones_like(z), np.zeros_like(z)
    out = z > 1
    r0[out] = 2 * z[out] ** 0.5 - 1
    r1[out] = z[out] ** -0.5
    r2[out] = -0.5 * z[out] ** -1.5
    return r0, r1, r2


def robust_scale(J, f):
    """common.scale_for_robust_loss_function."""
    r0, r1, r2 = huber_rho(f)
    js = r1 + 2 * r2 * f * f
    js[js < EPS] = EPS
    js = js ** 0.5
    return J * js[:, None], f * (r1 / js), 0.5 * np.sum(r0)


def jacobi_svd(A):
    """One-sided Jacobi SVD of an m x n matrix (m >= n): returns U[m,n], s[n] (descending), V[n,n]."""
    U = A.astype(float).copy()
    m, n = U.shape
    V = np.eye(n)
    for _ in range(60):
        rotated = False
        for p in range(n - 1):
            for q in range(p + 1, n):
                alpha = U[:, p] @ U[:, p]
                beta = U[:, q] @ U[:, q]
                gamma = U[:, p] @ U[:, q]
                if abs(gamma) <= 1e-300 or abs(gamma) <= EPS * (alpha * beta) ** 0.5:
                    continue
                rotated = True
                zeta = (beta - alpha) / (2.0 * gamma)
                t = (1.0 if zeta >= 0 else -1.0) / (abs(zeta) + (1.0 + zeta * zeta) ** 0.5)
                c = 1.0 / (1.0 + t * t) ** 0.5
                s = c * t
                up, uq = U[:, p].copy(), U[:, q].copy()
                U[:, p], U[:, q] = c * up - s * uq, s * up + c * uq
                vp, vq = V[:, p].copy(), V[:, q].copy()
                V[:, p], V[:, q] = c * vp - s * vq, s * vp + c * vq
        if not rotated:
            break
    s = np.sqrt(np.sum(U * U, axis=0))
    order = np.argsort(-s, kind="stable")
    s, U, V = s[order], U[:, order], V[:, order]
    Un = np.where(s > 0, U / np.where(s > 0, s, 1.0), 0.0)
    return Un, s, V


def solve_lsq_trust_region(n, m, uf, s, V, Delta, initial_alpha, rtol=0.01, max_iter=10):
    """common.solve_lsq_trust_region."""
    def phi_and_derivative(alpha):
        denom = s * s + alpha
        p_norm = np.sqrt(np.sum((suf / denom) ** 2))
        return p_norm - Delta, -np.sum(suf ** 2 / denom ** 3) / p_norm

    suf = s * uf
    full_rank = (m >= n) and (s[-1] > EPS * m * s[0])
    if full_rank:
        p = -V @ (uf / s)
        if np.sqrt(p @ p) <= Delta:
            return p, 0.0, 0
    alpha_upper = np.sqrt(suf @ suf) / Delta
    if full_rank:
        phi, phi_prime = phi_and_derivative(0.0)
        alpha_lower = -phi / phi_prime
    else:
        alpha_lower = 0.0
    if (not full_rank) and initial_alpha == 0:
        alpha = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
    else:
        alpha = initial_alpha
    it = 0
    for it in range(max_iter):
        if alpha < alpha_lower or alpha > alpha_upper:
            alpha = max(0.001 * alpha_upper, (alpha_lower * alpha_upper) ** 0.5)
        phi, phi_prime = phi_and_derivative(alpha)
        if phi < 0:
            alpha_upper = alpha
        ratio = phi / phi_prime
        alpha_lower = max(alpha_lower, alpha - ratio)
        alpha -= (phi + Delta) * ratio / Delta
        if abs(phi) < rtol * Delta:
            break
    p = -V @ (suf / (s * s + alpha))
    p *= Delta / np.sqrt(p @ p)
    return p, alpha, it + 1


def trf_huber(x, y, deg, ftol=1e-8, xtol=1e-8, gtol=1e-8, trace=None):
    """trf.trf_no_bounds for the polynomial residual with Huber loss, from x0 = ones."""
    n, m = deg + 1, len(x)
    a = np.ones(n)
    f_true = residuals(a, x, y, deg)
    nfev = 1
    J = jacobian_2point(a, f_true, x, y, deg)
    J, f, cost = robust_scale(J, f_true)
    g = J.T @ f
    Delta = np.sqrt(a @ a)
    if Delta == 0:
        Delta = 1.0
    max_nfev = 100 * n
    alpha = 0.0
    status = None
    while True:
        g_norm = np.max(np.abs(g))
        if g_norm < gtol:
            status = 1
        if status is not None or nfev == max_nfev:
            break
        U, s, V = jacobi_svd(J)
        uf = U.T @ f
        actual_reduction = -1.0
        while actual_reduction <= 0 and nfev < max_nfev:
            step, alpha, _ = solve_lsq_trust_region(n, m, uf, s, V, Delta, alpha)
            Js = J @ step
            predicted_reduction = -(0.5 * (Js @ Js) + step @ g)
            a_new = a + step
            f_new = residuals(a_new, x, y, deg)
            nfev += 1
            step_norm = np.sqrt(step @ step)
            if not np.all(np.isfinite(f_new)):
                Delta = 0.25 * step_norm
                continue
            cost_new = 0.5 * np.sum(huber_rho(f_new)[0])
            actual_reduction = cost - cost_new
            if predicted_reduction > 0:
                ratio = actual_reduction / predicted_reduction
            elif predicted_reduction == actual_reduction == 0:
                ratio = 1.0
            else:
                ratio = 0.0
            Delta_new = Delta
            if ratio < 0.25:
                Delta_new = 0.25 * step_norm
            elif ratio > 0.75 and step_norm > 0.95 * Delta:
                Delta_new = Delta * 2.0
            ftol_ok = actual_reduction < ftol * cost and ratio > 0.25
            xtol_ok = step_norm < xtol * (xtol + np.sqrt(a @ a))
            if ftol_ok and xtol_ok:
                status = 4
            elif ftol_ok:
                status = 2
            elif xtol_ok:
                status = 3
            if status is not None:
                break
            alpha *= Delta / Delta_new
            Delta = Delta_new
        if actual_reduction > 0:
            a = a_new
            f_true = f_new
            cost = cost_new
            J = jacobian_2point(a, f_true, x, y, deg)
            J, f, _ = robust_scale(J, f_true)
            g = J.T @ f
        if trace is not None:
            trace.append((nfev, cost, status))
    if status is None:
        status = 0
    return a, status, nfev


def robust_polyfit_restated(x, y, deg, trace=None):
    """models.robust_polyfit (models.py:70-123) with the SciPy call restated."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    xs, ys = x.std(), y.std()
    p, status, nfev = trf_huber(x / xs, y / ys, deg, trace=trace)
    p = p * ys
    for i in range(deg):
        p[i] /= xs ** (deg - i)
    return p[::-1], status, nfev
