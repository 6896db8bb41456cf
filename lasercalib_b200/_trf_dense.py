"""Trust-region loop for the two small dense variants of the reference
(`bundle_adjustment_camonly`, `bundleAdjust_transform_points_3d`, lasercalib/pySBA.py:158-206):
`scipy.optimize.least_squares(fun, x0, method='trf', ftol=ftol)` with every other argument at
its default, i.e. dense Jacobian, tr_solver='exact', x_scale=1.0, xtol = gtol = 1e-8,
max_nfev = 100 n  (scipy/optimize/_lsq/trf.py:392-560, common.py:80-180).

The O(N) work -- residuals, J^T f and J^T J over all observations -- runs on the GPU
(`lcba_sq_normal`); this module only holds the n x n control arithmetic (n = 12 or 11 C).
scipy drives the exact trust-region step from the SVD J = U S V^T; the same quantities follow
from the eigen-decomposition J^T J = V S^2 V^T and V^T J^T f = S U^T f.  When the caller says
the normal matrix is block diagonal (camera-only mode: cameras do not couple while the points
are fixed) the decomposition is done block by block, which keeps the small singular values of
one camera from drowning in the large ones of another.  The trust region is a ball in the
unscaled variables (x_scale = 1), so no column scaling is applied.
"""
from __future__ import annotations

import numpy as np

EPS = np.finfo(np.float64).eps


class DenseResult(dict):
    """Field-compatible with scipy.optimize.OptimizeResult for the fields the loop owns."""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__


def _eig_blocks(H, block):
    """Eigen-decomposition of a symmetric matrix that is block diagonal with `block`-sized
    blocks (block=None: dense).  Returns (lam ascending-within-block, V dense)."""
    n = H.shape[0]
    if block is None or block >= n:
        lam, V = np.linalg.eigh(H)
        return lam, V
    lam = np.empty(n)
    V = np.zeros((n, n))
    for a in range(0, n, block):
        b = a + block
        l, v = np.linalg.eigh(H[a:b, a:b])
        lam[a:b] = l
        V[a:b, a:b] = v
    return lam, V


def solve_tr_exact(lam, V, g, m, Delta, initial_alpha, rtol=0.01, max_iter=10):
    """common.py:80-180 (`solve_lsq_trust_region`) restated on (s^2, V, V^T g)."""
    s2 = np.maximum(lam, 0.0)
    s = np.sqrt(s2)
    suf = V.T @ g                                    # = s * (U^T f)
    smax, smin = s.max(), s.min()
    full_rank = (m >= g.size) and (smin > EPS * m * smax)

    def phi_and_derivative(alpha):
        denom = s2 + alpha
        p_norm = np.linalg.norm(suf / denom)
        return p_norm - Delta, -np.sum(suf ** 2 / denom ** 3) / p_norm

    if full_rank:
        p = -V @ (suf / s2)
        if np.linalg.norm(p) <= Delta:
            return p, 0.0, 0
    alpha_upper = np.linalg.norm(suf) / Delta
    if full_rank:
        phi, phi_prime = phi_and_derivative(0.0)
        alpha_lower = -phi / phi_prime
    else:
        alpha_lower = 0.0
    if initial_alpha is None or (not full_rank and initial_alpha == 0):
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
    p = -V @ (suf / (s2 + alpha))
    p *= Delta / np.linalg.norm(p)
    return p, alpha, it + 1


def _update_tr_radius(Delta, actual, predicted, step_norm, bound_hit):
    """common.py:197-222."""
    if predicted > 0:
        ratio = actual / predicted
    elif predicted == actual == 0:
        ratio = 1.0
    else:
        ratio = 0.0
    if ratio < 0.25:
        Delta = 0.25 * step_norm
    elif ratio > 0.75 and bound_hit:
        Delta *= 2.0
    return Delta, ratio


def _check_termination(dF, F, dx_norm, x_norm, ratio, ftol, xtol):
    """common.py:689-703."""
    ftol_ok = dF < ftol * F and ratio > 0.25
    xtol_ok = dx_norm < xtol * (xtol + x_norm)
    if ftol_ok and xtol_ok:
        return 4
    if ftol_ok:
        return 2
    if xtol_ok:
        return 3
    return None


MESSAGES = {-1: "Improper input parameters status returned from `leastsq`",
            0: "The maximum number of function evaluations is exceeded.",
            1: "`gtol` termination condition is satisfied.",
            2: "`ftol` termination condition is satisfied.",
            3: "`xtol` termination condition is satisfied.",
            4: "Both `ftol` and `xtol` termination conditions are satisfied."}


def trf_dense(evaluate, x0, m, ftol, xtol=1e-8, gtol=1e-8, max_nfev=None, verbose=2, block=None):
    """`evaluate(x, derivs)` -> (cost, g, H) with g = J^T f, H = J^T J (None, None when
    derivs is False).  m = number of residuals.  Follows trf_no_bounds (trf.py:392-560)."""
    x = np.array(x0, dtype=np.float64).ravel()
    n = x.size
    if max_nfev is None:
        max_nfev = 100 * n
    cost, g, H = evaluate(x, True)
    if not np.isfinite(cost):
        raise ValueError("Residuals are not finite in the initial point.")
    nfev = njev = 1
    Delta = np.linalg.norm(x)
    if Delta == 0:
        Delta = 1.0
    alpha = 0.0
    status = None
    iteration = 0
    step_norm = actual = None
    if verbose == 2:
        print("{:^15}{:^15}{:^15}{:^15}{:^15}{:^15}".format(
            "Iteration", "Total nfev", "Cost", "Cost reduction", "Step norm", "Optimality"))
    trace = []
    while True:
        g_norm = np.linalg.norm(g, ord=np.inf)
        if g_norm < gtol:
            status = 1
        if verbose == 2:
            a = "{:^15.2e}".format(actual) if actual is not None else " " * 15
            b = "{:^15.2e}".format(step_norm) if step_norm is not None else " " * 15
            print("{:^15}{:^15}{:^15.4e}{}{}{:^15.2e}".format(iteration, nfev, cost, a, b, g_norm))
        trace.append((iteration, nfev, cost, g_norm))
        if status is not None or nfev == max_nfev:
            break
        lam, V = _eig_blocks(H, block)
        actual = -1.0
        while actual <= 0 and nfev < max_nfev:
            step, alpha, _ = solve_tr_exact(lam, V, g, m, Delta, alpha)
            predicted = -(0.5 * step @ (H @ step) + g @ step)
            x_new = x + step
            cost_new, _, _ = evaluate(x_new, False)
            nfev += 1
            step_h_norm = np.linalg.norm(step)
            if not np.isfinite(cost_new):
                Delta = 0.25 * step_h_norm
                continue
            actual = cost - cost_new
            Delta_new, ratio = _update_tr_radius(Delta, actual, predicted, step_h_norm,
                                                 step_h_norm > 0.95 * Delta)
            step_norm = step_h_norm
            status = _check_termination(actual, cost, step_norm, np.linalg.norm(x), ratio, ftol, xtol)
            if status is not None:
                break
            alpha *= Delta / Delta_new
            Delta = Delta_new
        if actual > 0:
            x = x_new
            cost, g, H = evaluate(x, True)
            njev += 1
        else:
            step_norm = 0
            actual = 0
        iteration += 1
    if status is None:
        status = 0
    if verbose >= 1:
        print(MESSAGES[status])
        print("Function evaluations {}, initial cost {:.4e}, final cost {:.4e}, "
              "first-order optimality {:.2e}.".format(nfev, trace[0][2], cost, g_norm))
    return DenseResult(x=x, cost=cost, grad=g, optimality=g_norm, active_mask=np.zeros(n),
                       nfev=nfev, njev=njev, status=status, message=MESSAGES[status],
                       success=status > 0, trace=trace)
