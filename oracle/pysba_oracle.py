"""CPU oracle for the bundle-adjustment hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product path (``lasercalib_b200/``) may import this module.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` use it, and only as the checker / the timed CPU arm.

What it restates (reference = JohnsonLabJanelia/laserCalib, file:line under
``/root/reference``):

* ``rotate``                       <- ``lasercalib/pySBA.py:61-73``  (Rodrigues, theta=0 -> identity)
* ``project``                      <- ``lasercalib/pySBA.py:76-89``  (pinhole + k1,k2 radial)
* ``fun``                          <- ``lasercalib/pySBA.py:92-101`` (weighted residual, [u0,v0,u1,v1,...])
* ``bundle_adjustment_sparsity``   <- ``lasercalib/pySBA.py:103-118`` (28 nnz per observation)
* ``bundle_adjust``                <- ``lasercalib/pySBA.py:132-147`` (scipy least_squares: TRF,
                                      x_scale='jac', jac='3-point', jac_sparsity=A)

The solver arithmetic is NOT in the reference tree: it is the third-party
``scipy.optimize.least_squares`` (un-pinned by the reference: ``setup.py:43``
``install_requires=[]``).  The oracle calls the scipy installed in this image
(1.18.1 when the golden vectors were made; the version is stored in every golden file)
exactly as the reference's call site does.  For step-by-step checks of the CUDA engine
``trf_exact`` restates scipy's ``trf_no_bounds`` (``scipy/optimize/_lsq/trf.py:415-587``)
with the inexact LSMR Gauss-Newton direction replaced by the exact solution of the same
regularised normal equations (see DESIGN.md "solver semantics").

Parity pinning: the reference has no tests or golden vectors, so
the oracle is pinned against outputs of the unmodified reference itself, generated in
the build container by ``tests/golden/make_golden.py`` and committed under
``tests/golden/*.npz`` (``tests/test_oracle_golden.py`` checks bit-for-bit / 1e-15).
"""
from __future__ import annotations

import numpy as np
import scipy
from scipy.optimize import OptimizeResult, least_squares
from scipy.sparse import csr_matrix

NCP = 11  # camera parameters: rotvec(3) t(3) f k1 k2 cx cy   (pySBA.py:31-35)


# --------------------------------------------------------------------------- model
def rotate(points, rot_vecs):
    """Rodrigues rotation of each row of ``points`` by its own rotation vector.

    Restates ``PySBA.rotate`` (pySBA.py:61-73): unit axis = r/|r| with 0/0 -> 0 so
    that |r| = 0 is the identity.
    """
    ang = np.sqrt(np.sum(rot_vecs * rot_vecs, axis=1))[:, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        axis = np.where(ang > 0, rot_vecs / ang, 0.0)
    c, s = np.cos(ang), np.sin(ang)
    along = np.sum(points * axis, axis=1)[:, None]
    return c * points + s * np.cross(axis, points) + along * (1.0 - c) * axis


def project(points, cams):
    """(M,3) points with per-row (M,11) camera vectors -> (M,2) pixels (pySBA.py:76-89)."""
    pc = rotate(points, cams[:, 0:3]) + cams[:, 3:6]
    xy = pc[:, 0:2] / pc[:, 2:3]
    n = np.sum(xy * xy, axis=1)
    dist = 1.0 + cams[:, 7] * n + cams[:, 8] * n * n
    return xy * (dist * cams[:, 6])[:, None] + cams[:, 9:11]


def fun(params, n_cameras, n_points, camera_indices, point_indices, points_2d, weights):
    """Weighted reprojection residual, interleaved (u,v) per observation (pySBA.py:92-101).

    ``weights`` has shape (N,1) like ``PySBA.pointWeights``.
    """
    cams = params[: n_cameras * NCP].reshape(n_cameras, NCP)
    pts = params[n_cameras * NCP:].reshape(n_points, 3)
    uv = project(pts[point_indices], cams[camera_indices])
    return (weights * (uv - points_2d)).ravel()


def bundle_adjustment_sparsity(n_cameras, n_points, camera_indices, point_indices):
    """0/1 Jacobian pattern (2N, 11C+3P): rows 2i and 2i+1 both hold the 11 camera
    columns and the 3 point columns of observation i (pySBA.py:103-118).

    Built directly as CSR (the reference fills a lil_matrix; same pattern, same dtype).
    """
    N = camera_indices.size
    cols = np.empty((N, 2, NCP + 3), dtype=np.int64)
    cols[:, :, :NCP] = (camera_indices[:, None] * NCP + np.arange(NCP))[:, None, :]
    cols[:, :, NCP:] = (n_cameras * NCP + point_indices[:, None] * 3 + np.arange(3))[:, None, :]
    indptr = np.arange(0, 2 * N * (NCP + 3) + 1, NCP + 3, dtype=np.int64)
    data = np.ones(cols.size, dtype=int)
    return csr_matrix((data, cols.ravel(), indptr), shape=(2 * N, n_cameras * NCP + 3 * n_points))


def default_weights(point_indices):
    """``np.full_like(point2DIndices, 1).reshape(-1,1)`` (pySBA.py:56-58): int64 ones."""
    return np.full_like(point_indices, 1).reshape(-1, 1)


def bundle_adjust(cams, pts, points_2d, camera_indices, point_indices, weights=None,
                  ftol=1e-4, verbose=0, **ls_kwargs):
    """The reference's ``bundleAdjust`` (pySBA.py:132-147) through scipy ``least_squares``.

    Extra keyword arguments go to ``least_squares`` (e.g. ``tr_options`` for the
    tight-LSMR variant of SURVEY App. C.0, explicit ``xtol``/``gtol``/``max_nfev``).
    Returns (OptimizeResult, cams_out (C,11), pts_out (P,3)).
    """
    C, P = cams.shape[0], pts.shape[0]
    if weights is None:
        weights = default_weights(point_indices)
    weights = np.asarray(weights).reshape(-1, 1)
    x0 = np.hstack((cams.ravel(), pts.ravel()))
    A = bundle_adjustment_sparsity(C, P, camera_indices, point_indices)
    res = least_squares(fun, x0, jac_sparsity=A, verbose=verbose, x_scale="jac", ftol=ftol,
                        method="trf", jac="3-point",
                        args=(C, P, camera_indices, point_indices, points_2d, weights),
                        **ls_kwargs)
    return res, res.x[: C * NCP].reshape(C, NCP), res.x[C * NCP:].reshape(P, 3)


# ------------------------------------------------------------ Jacobian oracles
def jacobian_fd(x, n_cameras, n_points, camera_indices, point_indices, points_2d, weights):
    """scipy's own sparse 3-point finite-difference Jacobian of ``fun`` — the matrix the
    reference's solver actually uses (scipy/optimize/_numdiff.py:288,770-893)."""
    from scipy.optimize._numdiff import approx_derivative
    A = bundle_adjustment_sparsity(n_cameras, n_points, camera_indices, point_indices)
    return approx_derivative(fun, x, method="3-point", sparsity=A,
                             args=(n_cameras, n_points, camera_indices, point_indices,
                                   points_2d, weights))


def jacobian_truth_longdouble(x, n_cameras, n_points, camera_indices, point_indices,
                              points_2d, weights, h_rel=1e-5):
    """Dense extended-precision Jacobian truth: Richardson-extrapolated central
    differences of ``fun`` evaluated in np.longdouble.  O(n) evaluations
    of ``fun`` -> small rigs only."""
    ld = np.longdouble
    xl = x.astype(ld)
    p2 = points_2d.astype(ld)
    w = weights.astype(ld)
    args = (n_cameras, n_points, camera_indices, point_indices, p2, w)
    m = 2 * camera_indices.size
    J = np.zeros((m, x.size), dtype=np.float64)

    def cd(j, h):
        e = np.zeros_like(xl)
        e[j] = h
        return (fun(xl + e, *args) - fun(xl - e, *args)) / (2 * h)

    for j in range(x.size):
        h = ld(h_rel) * max(ld(1.0), abs(xl[j]))
        J[:, j] = ((4 * cd(j, h / 2) - cd(j, h)) / 3).astype(np.float64)
    return J


def rotation_and_derivatives(r):
    """R(r) and dR/dr_k (k=0..2) for one rotation vector, series-safe near 0.

    R = I + a K + b K^2,  K = [r]x,  a = sin(t)/t,  b = (1-cos t)/t^2.
    dR/dr_k = c1 r_k K + a E_k + c2 r_k K^2 + b (E_k K + K E_k),  E_k = [e_k]x,
    c1 = (t cos t - sin t)/t^3,  c2 = (t sin t - 2(1-cos t))/t^4.   
    """
    r = np.asarray(r, dtype=np.float64)
    t2 = float(r @ r)
    t = np.sqrt(t2)
    if t < 0.1:
        a = 1 - t2 / 6 * (1 - t2 / 20 * (1 - t2 / 42 * (1 - t2 / 72 * (1 - t2 / 110))))
        b = 0.5 * (1 - t2 / 12 * (1 - t2 / 30 * (1 - t2 / 56 * (1 - t2 / 90 * (1 - t2 / 132)))))
        c1 = -1 / 3 + t2 / 30 - t2**2 / 840 + t2**3 / 45360 - t2**4 / 3991680 + t2**5 / 518918400
        c2 = -1 / 12 + t2 / 180 - t2**2 / 6720 + t2**3 / 453600 - t2**4 / 47900160 \
            + t2**5 / 7264857600
    else:
        s, c = np.sin(t), np.cos(t)
        a = s / t
        b = (1 - c) / t2
        c1 = (t * c - s) / (t * t2)
        c2 = (t * s - 2 * (1 - c)) / (t2 * t2)
    K = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    K2 = K @ K
    R = np.eye(3) + a * K + b * K2
    dR = np.empty((3, 3, 3))
    for k in range(3):
        E = np.zeros((3, 3))
        i, j = (k + 1) % 3, (k + 2) % 3
        E[j, i], E[i, j] = 1.0, -1.0
        dR[k] = c1 * r[k] * K + a * E + c2 * r[k] * K2 + b * (E @ K + K @ E)
    return R, dR


def jacobian_blocks(cams, pts, camera_indices, point_indices, weights=None):
    """Analytic per-observation Jacobian blocks of ``fun`` (derived):
    returns (uv (N,2) projected pixels, Jc (N,2,11), Jp (N,2,3)), weights applied to Jc/Jp.

    This is a derived quantity (the reference differentiates numerically); it is checked
    against ``jacobian_truth_longdouble`` and ``jacobian_fd`` in tests/test_oracle_golden.py.
    """
    C = cams.shape[0]
    N = camera_indices.size
    Rs = np.empty((C, 3, 3))
    dRs = np.empty((C, 3, 3, 3))
    for c in range(C):
        Rs[c], dRs[c] = rotation_and_derivatives(cams[c, :3])
    cam = cams[camera_indices]
    X = pts[point_indices]
    R = Rs[camera_indices]                       # (N,3,3)
    Xc = np.einsum("nij,nj->ni", R, X) + cam[:, 3:6]
    iz = 1.0 / Xc[:, 2]
    x, y = Xc[:, 0] * iz, Xc[:, 1] * iz
    n = x * x + y * y
    f, k1, k2 = cam[:, 6], cam[:, 7], cam[:, 8]
    d = 1 + k1 * n + k2 * n * n
    dp = k1 + 2 * k2 * n
    uv = np.stack([f * d * x + cam[:, 9], f * d * y + cam[:, 10]], axis=1)
    # d(u,v)/d(x,y)
    E = np.empty((N, 2, 2))
    E[:, 0, 0] = f * (d + 2 * x * x * dp)
    E[:, 0, 1] = E[:, 1, 0] = f * 2 * x * y * dp
    E[:, 1, 1] = f * (d + 2 * y * y * dp)
    # d(x,y)/dXc
    D = np.zeros((N, 2, 3))
    D[:, 0, 0] = iz
    D[:, 1, 1] = iz
    D[:, 0, 2] = -x * iz
    D[:, 1, 2] = -y * iz
    G = E @ D                                    # (N,2,3) = d(u,v)/dXc
    Jc = np.zeros((N, 2, NCP))
    dRX = np.einsum("nkij,nj->nik", dRs[camera_indices], X)   # (N,3,3): column k = dR_k X
    Jc[:, :, 0:3] = G @ dRX
    Jc[:, :, 3:6] = G
    Jc[:, 0, 6], Jc[:, 1, 6] = d * x, d * y
    Jc[:, 0, 7], Jc[:, 1, 7] = f * n * x, f * n * y
    Jc[:, 0, 8], Jc[:, 1, 8] = f * n * n * x, f * n * n * y
    Jc[:, 0, 9] = 1.0
    Jc[:, 1, 10] = 1.0
    Jp = G @ R
    if weights is not None:
        w = np.asarray(weights, dtype=np.float64).reshape(-1, 1, 1)
        Jc = Jc * w
        Jp = Jp * w
    return uv, Jc, Jp


def jacobian_csr(Jc, Jp, n_cameras, n_points, camera_indices, point_indices):
    """Assemble blocks into the (2N, 11C+3P) CSR with the reference's column order."""
    N = camera_indices.size
    A = bundle_adjustment_sparsity(n_cameras, n_points, camera_indices, point_indices)
    data = np.concatenate([Jc, Jp], axis=2).reshape(-1)
    return csr_matrix((data, A.indices, A.indptr), shape=A.shape)


# ------------------------------------------------- normal equations / Schur oracle
def normal_blocks(res2, Jc, Jp, n_cameras, n_points, camera_indices, point_indices):
    """J^T J / J^T f in block form: U (C,11,11), gc (C,11), V (P,3,3), gp (P,3), W (N,11,3)."""
    U = np.zeros((n_cameras, NCP, NCP))
    gc = np.zeros((n_cameras, NCP))
    V = np.zeros((n_points, 3, 3))
    gp = np.zeros((n_points, 3))
    np.add.at(U, camera_indices, np.einsum("nia,nib->nab", Jc, Jc))
    np.add.at(gc, camera_indices, np.einsum("nia,ni->na", Jc, res2))
    np.add.at(V, point_indices, np.einsum("nia,nib->nab", Jp, Jp))
    np.add.at(gp, point_indices, np.einsum("nia,ni->na", Jp, res2))
    W = np.einsum("nia,nib->nab", Jc, Jp)
    return U, gc, V, gp, W


def reduced_camera_system(U, gc, V, gp, W, camera_indices, point_indices, lam, scale_inv):
    """Schur complement of the damped normal equations
        (J^T J + lam * diag(scale_inv^2)) p = J^T f
    onto the cameras: S (11C,11C), rhs (11C), and the per-point inverse blocks.
    """
    C, P = U.shape[0], V.shape[0]
    sc = scale_inv[: C * NCP].reshape(C, NCP)
    sp = scale_inv[C * NCP:].reshape(P, 3)
    Vd = V.copy()
    Vd[:, np.arange(3), np.arange(3)] += lam * sp * sp
    Vinv = np.linalg.inv(Vd)
    S = np.zeros((C * NCP, C * NCP))
    for c in range(C):
        S[c * NCP:(c + 1) * NCP, c * NCP:(c + 1) * NCP] = U[c] + lam * np.diag(sc[c] ** 2)
    rhs = gc.reshape(-1).copy()
    # dense E_p (C*11 x 3) per point, chunked
    chunk = max(1, int(2e7 // (C * NCP * 3)))
    order = np.argsort(point_indices, kind="stable")
    pi_sorted = point_indices[order]
    bounds = np.searchsorted(pi_sorted, np.arange(0, P + chunk, chunk))
    for b in range(len(bounds) - 1):
        p0 = b * chunk
        p1 = min(P, p0 + chunk)
        sel = order[bounds[b]:bounds[b + 1]]
        if sel.size == 0:
            continue
        E = np.zeros((p1 - p0, C, NCP, 3))
        np.add.at(E, (point_indices[sel] - p0, camera_indices[sel]), W[sel])
        E = E.reshape(p1 - p0, C * NCP, 3)
        T = E @ Vinv[p0:p1]                                  # (p, 11C, 3)
        S -= np.einsum("pac,pbc->ab", T, E, optimize=True)
        rhs -= np.einsum("pac,pc->a", T, gp[p0:p1])
    return S, rhs, Vinv


def solve_damped_normal_equations(U, gc, V, gp, W, camera_indices, point_indices, lam, scale_inv):
    """Exact p = (J^T J + lam diag(scale_inv^2))^-1 J^T f via the reduced camera system."""
    from scipy.linalg import cho_factor, cho_solve
    C, P = U.shape[0], V.shape[0]
    S, rhs, Vinv = reduced_camera_system(U, gc, V, gp, W, camera_indices, point_indices,
                                         lam, scale_inv)
    # When reg_term underflows the 7 gauge modes make S numerically indefinite; the CUDA
    # engine then adds extra damping mu*diag(scale_inv_c^2) on the camera block only and
    # retries (DESIGN.md "Cholesky breakdown"); the oracle follows the same schedule.
    mu = 0.0
    dc2 = scale_inv[: C * NCP] ** 2
    while True:
        try:
            pc = cho_solve(cho_factor(S + np.diag(mu * dc2), lower=True), rhs)
            break
        except np.linalg.LinAlgError:
            mu = max(10.0 * mu, 10.0 * lam, 1e-13)
            if mu > 1e6:
                raise
    t = gp.copy()
    np.subtract.at(t, point_indices,
                   np.einsum("nab,na->nb", W, pc.reshape(C, NCP)[camera_indices]))
    pp = np.einsum("pab,pb->pa", Vinv, t)
    return np.hstack((pc, pp.ravel())), S, rhs


# -------------------------------------------------- exact-solve TRF restatement
def trf_exact(cams, pts, points_2d, camera_indices, point_indices, weights=None,
              ftol=1e-4, xtol=1e-8, gtol=1e-8, max_nfev=None, verbose=0):
    """scipy ``trf_no_bounds`` (scipy/optimize/_lsq/trf.py:415-587) restated for the BA
    problem with two substitutions: analytic Jacobian blocks instead of the 3-point finite
    difference, and the exact solution of (J_h^T J_h + reg_term I) p = J_h^T f instead of
    LSMR's inexact one (atol=btol=1e-6).  Everything else — Jacobi scaling with running
    max (common.py:598-610), reg_term from the 1-D Cauchy model (trf.py:488-492), the 2-D
    subspace trust-region problem (common.py:171-219), radius update (common.py:222-245)
    and termination tests (common.py:705-717) — uses scipy's own helper functions.

    Returns an OptimizeResult with an extra ``trace`` list (one dict per outer iteration).
    """
    from numpy.linalg import norm
    from scipy.linalg import qr
    from scipy.optimize._lsq.common import (check_termination, minimize_quadratic_1d,
                                            solve_trust_region_2d, update_tr_radius)
    C, P = cams.shape[0], pts.shape[0]
    if weights is None:
        weights = default_weights(point_indices)
    weights = np.asarray(weights).reshape(-1, 1)
    args = (C, P, camera_indices, point_indices, points_2d, weights)
    x = np.hstack((cams.ravel(), pts.ravel())).astype(np.float64)
    n = x.size

    def linearize(xv):
        cm = xv[: C * NCP].reshape(C, NCP)
        pt = xv[C * NCP:].reshape(P, 3)
        uv, Jc, Jp = jacobian_blocks(cm, pt, camera_indices, point_indices, weights)
        return Jc, Jp

    f = fun(x, *args)
    if not np.all(np.isfinite(f)):
        raise ValueError("Residuals are not finite in the initial point.")
    nfev = 1
    Jc, Jp = linearize(x)
    njev = 1
    cost = 0.5 * np.dot(f, f)
    blocks = normal_blocks(f.reshape(-1, 2), Jc, Jp, C, P, camera_indices, point_indices)
    g = np.hstack((blocks[1].ravel(), blocks[3].ravel()))

    def col_norms(bl):
        U, _, V, _, _ = bl
        dc = np.sqrt(np.einsum("caa->ca", U)).ravel()
        dp_ = np.sqrt(np.einsum("paa->pa", V)).ravel()
        return np.hstack((dc, dp_))

    scale_inv = col_norms(blocks)
    scale_inv[scale_inv == 0] = 1
    scale = 1 / scale_inv
    Delta = norm(x * scale_inv)
    if Delta == 0:
        Delta = 1.0
    if max_nfev is None:
        max_nfev = n * 100

    def J_dot(v):
        """(J v) as (N,2) for an unscaled parameter-space vector v."""
        vc = v[: C * NCP].reshape(C, NCP)[camera_indices]
        vp = v[C * NCP:].reshape(P, 3)[point_indices]
        return np.einsum("nia,na->ni", Jc, vc) + np.einsum("nia,na->ni", Jp, vp)

    termination_status = None
    iteration = 0
    step_norm = None
    actual_reduction = None
    trace = []
    if verbose == 2:
        from scipy.optimize._lsq.common import print_header_nonlinear
        print_header_nonlinear()
    while True:
        g_norm = norm(g, ord=np.inf)
        if g_norm < gtol:
            termination_status = 1
        if verbose == 2:
            from scipy.optimize._lsq.common import print_iteration_nonlinear
            print_iteration_nonlinear(iteration, nfev, cost, actual_reduction, step_norm, g_norm)
        if termination_status is not None or nfev == max_nfev:
            break
        d = scale
        g_h = d * g
        # reg_term (trf.py:488-492): a = 0.5 |J_h g_h|^2 along s = -g_h, b = -|g_h|^2
        Jg = J_dot(d * g_h)
        a = 0.5 * np.sum(Jg * Jg)
        b = -np.dot(g_h, g_h)
        to_tr = Delta / norm(g_h)
        ag_value = minimize_quadratic_1d(a, b, 0, to_tr)[1]
        reg_term = -ag_value / Delta**2
        # exact regularised Gauss-Newton direction in unscaled variables
        p, S_red, rhs_red = solve_damped_normal_equations(
            *blocks, camera_indices, point_indices, reg_term, scale_inv)
        gn_h = p * scale_inv
        Sq = np.vstack((g_h, gn_h)).T
        Sq, _ = qr(Sq, mode="economic")
        JS = np.stack([J_dot(d * Sq[:, 0]).ravel(), J_dot(d * Sq[:, 1]).ravel()], axis=1)
        B_S = JS.T @ JS
        g_S = Sq.T @ g_h
        rec = dict(iteration=iteration, cost=cost, g_norm=g_norm, Delta=Delta,
                   reg_term=reg_term, trials=[])
        actual_reduction = -1
        while actual_reduction <= 0 and nfev < max_nfev:
            p_S, _ = solve_trust_region_2d(B_S, g_S, Delta)
            step_h = Sq @ p_S
            Js = JS @ p_S
            predicted_reduction = -(0.5 * np.dot(Js, Js) + np.dot(step_h, g_h))
            step = d * step_h
            x_new = x + step
            f_new = fun(x_new, *args)
            nfev += 1
            step_h_norm = norm(step_h)
            if not np.all(np.isfinite(f_new)):
                Delta = 0.25 * step_h_norm
                continue
            cost_new = 0.5 * np.dot(f_new, f_new)
            actual_reduction = cost - cost_new
            Delta_new, ratio = update_tr_radius(Delta, actual_reduction, predicted_reduction,
                                                step_h_norm, step_h_norm > 0.95 * Delta)
            step_norm = norm(step)
            termination_status = check_termination(actual_reduction, cost, step_norm, norm(x),
                                                   ratio, ftol, xtol)
            rec["trials"].append(dict(Delta=Delta, predicted=predicted_reduction,
                                      actual=actual_reduction, ratio=ratio,
                                      step_norm=step_norm, step_h_norm=step_h_norm,
                                      cost_new=cost_new))
            if termination_status is not None:
                break
            Delta = Delta_new
        trace.append(rec)
        if actual_reduction > 0:
            x = x_new
            f = f_new
            cost = cost_new
            Jc, Jp = linearize(x)
            njev += 1
            blocks = normal_blocks(f.reshape(-1, 2), Jc, Jp, C, P, camera_indices, point_indices)
            g = np.hstack((blocks[1].ravel(), blocks[3].ravel()))
            scale_inv = np.maximum(col_norms(blocks), scale_inv)
            scale = 1 / scale_inv
        else:
            step_norm = 0
            actual_reduction = 0
        iteration += 1
    if termination_status is None:
        termination_status = 0
    res = OptimizeResult(x=x, cost=cost, fun=f, grad=g, optimality=g_norm,
                         active_mask=np.zeros_like(x), nfev=nfev, njev=njev,
                         status=termination_status, success=termination_status > 0)
    res["trace"] = trace
    res["scipy_version"] = scipy.__version__
    return res


def fun_nocam(params, camera_params, n_points, camera_indices, point_indices, points_2d, weights):
    """Residuals with the 3-D points as the only unknowns (pySBA.py:226-235)."""
    pts = params.reshape((n_points, 3))
    uv = project(pts[point_indices], camera_params[camera_indices])
    return (weights * (uv - points_2d)).ravel()


def bundle_adjust_nocam(cams, pts, points_2d, camera_indices, point_indices, weights=None,
                        ftol=1e-7, verbose=0, **ls_kwargs):
    """The reference's ``bundleAdjust_nocam`` (pySBA.py:237-250) through scipy."""
    P = pts.shape[0]
    if weights is None:
        weights = default_weights(point_indices)
    weights = np.asarray(weights).reshape(-1, 1)
    N = point_indices.size
    cols = (point_indices[:, None] * 3 + np.arange(3))[:, None, :].repeat(2, axis=1)
    A = csr_matrix((np.ones(cols.size, dtype=int), cols.ravel(),
                    np.arange(0, 6 * N + 1, 3, dtype=np.int64)), shape=(2 * N, 3 * P))
    return least_squares(fun_nocam, pts.ravel(), jac_sparsity=A, verbose=verbose, x_scale="jac",
                         ftol=ftol, method="trf", jac="3-point",
                         args=(cams, P, camera_indices, point_indices, points_2d, weights),
                         **ls_kwargs)


def trf_exact_nocam(cams, pts, points_2d, camera_indices, point_indices, weights=None,
                    ftol=1e-7, xtol=1e-8, gtol=1e-8, max_nfev=None):
    """``trf_exact`` with the cameras held fixed (x = points): the damped normal equations are
    block diagonal, p_p = (V_p + reg_term Dp^2)^-1 g_p."""
    from numpy.linalg import norm
    from scipy.linalg import qr
    from scipy.optimize._lsq.common import (check_termination, minimize_quadratic_1d,
                                            solve_trust_region_2d, update_tr_radius)
    C, P = cams.shape[0], pts.shape[0]
    if weights is None:
        weights = default_weights(point_indices)
    weights = np.asarray(weights).reshape(-1, 1)
    args = (cams, P, camera_indices, point_indices, points_2d, weights)
    x = pts.ravel().astype(np.float64)

    def lin(xv):
        fv = fun_nocam(xv, *args)
        _, _, Jp = jacobian_blocks(cams, xv.reshape(P, 3), camera_indices, point_indices, weights)
        V = np.zeros((P, 3, 3))
        gp = np.zeros((P, 3))
        np.add.at(V, point_indices, np.einsum("nia,nib->nab", Jp, Jp))
        np.add.at(gp, point_indices, np.einsum("nia,ni->na", Jp, fv.reshape(-1, 2)))
        return fv, Jp, V, gp

    f, Jp, V, gp = lin(x)
    if not np.all(np.isfinite(f)):
        raise ValueError("Residuals are not finite in the initial point.")
    nfev = njev = 1
    cost = 0.5 * f @ f
    g = gp.ravel()
    scale_inv = np.sqrt(np.einsum("paa->pa", V)).ravel()
    scale_inv[scale_inv == 0] = 1
    Delta = norm(x * scale_inv) or 1.0
    if max_nfev is None:
        max_nfev = x.size * 100
    Jdot = lambda v: np.einsum("nia,na->ni", Jp, v.reshape(P, 3)[point_indices])
    status, trace = None, []
    g_norm = norm(g, ord=np.inf)
    while True:
        g_norm = norm(g, ord=np.inf)
        if g_norm < gtol:
            status = 1
        trace.append(cost)
        if status is not None or nfev == max_nfev:
            break
        d = 1 / scale_inv
        g_h = d * g
        Jg = Jdot(d * g_h)
        a, b = 0.5 * np.sum(Jg * Jg), -g_h @ g_h
        reg = -minimize_quadratic_1d(a, b, 0, Delta / norm(g_h))[1] / Delta**2
        Vd = V.copy()
        Vd[:, np.arange(3), np.arange(3)] += reg * scale_inv.reshape(P, 3) ** 2
        p = np.linalg.solve(Vd, gp[:, :, None])[:, :, 0].ravel()
        Sq, _ = qr(np.vstack((g_h, p * scale_inv)).T, mode="economic")
        JS = np.stack([Jdot(d * Sq[:, 0]).ravel(), Jdot(d * Sq[:, 1]).ravel()], axis=1)
        B_S, g_S = JS.T @ JS, Sq.T @ g_h
        actual = -1
        while actual <= 0 and nfev < max_nfev:
            p_S, _ = solve_trust_region_2d(B_S, g_S, Delta)
            step_h = Sq @ p_S
            Js = JS @ p_S
            pred = -(0.5 * Js @ Js + step_h @ g_h)
            x_new = x + d * step_h
            f_new = fun_nocam(x_new, *args)
            nfev += 1
            shn = norm(step_h)
            if not np.all(np.isfinite(f_new)):
                Delta = 0.25 * shn
                continue
            cost_new = 0.5 * f_new @ f_new
            actual = cost - cost_new
            Delta_new, ratio = update_tr_radius(Delta, actual, pred, shn, shn > 0.95 * Delta)
            status = check_termination(actual, cost, norm(d * step_h), norm(x), ratio, ftol, xtol)
            if status is not None:
                break
            Delta = Delta_new
        if actual > 0:
            x, cost = x_new, cost_new
            f, Jp, V, gp = lin(x)
            njev += 1
            g = gp.ravel()
            scale_inv = np.maximum(np.sqrt(np.einsum("paa->pa", V)).ravel(), scale_inv)
    return OptimizeResult(x=x, cost=cost, fun=f, grad=g, optimality=g_norm, nfev=nfev, njev=njev,
                          status=status or 0, trace=trace)


# ------------------------------------------------------- shared intrinsics (pySBA.py:252-325)
def sharedcam_pack(cams, pts):
    """x0 of ``bundleAdjust_sharedcam``: [mean(f,k1,k2) | extrinsics C*6 | centroids C*2 | points]."""
    return np.hstack((np.mean(cams[:, 6:9], axis=0).ravel(), cams[:, :6].ravel(),
                      cams[:, 9:].ravel(), pts.ravel()))


def sharedcam_unpack(x, C, P):
    sh = x[:3]
    ext = x[3:3 + 6 * C].reshape(C, 6)
    cen = x[3 + 6 * C:3 + 8 * C].reshape(C, 2)
    cams = np.concatenate((ext, np.tile(sh, (C, 1)), cen), axis=1)
    return cams, x[3 + 8 * C:].reshape(P, 3)


def fun_sharedcam(params, n_cameras, n_points, camera_indices, point_indices, points_2d, weights):
    cams, pts = sharedcam_unpack(params, n_cameras, n_points)
    uv = project(pts[point_indices], cams[camera_indices])
    return (weights * (uv - points_2d)).ravel()


def sparsity_sharedcam(n_cameras, n_points, camera_indices, point_indices):
    """pySBA.py:252-276: every row sees the 3 shared columns, 6 + 2 own camera columns, 3 point
    columns."""
    N, C = camera_indices.size, n_cameras
    cols = np.empty((N, 14), dtype=np.int64)
    cols[:, 0:3] = np.arange(3)
    cols[:, 3:9] = 3 + camera_indices[:, None] * 6 + np.arange(6)
    cols[:, 9:11] = 3 + 6 * C + camera_indices[:, None] * 2 + np.arange(2)
    cols[:, 11:14] = 3 + 8 * C + point_indices[:, None] * 3 + np.arange(3)
    cols = np.repeat(cols[:, None, :], 2, axis=1)
    return csr_matrix((np.ones(cols.size, dtype=int), cols.ravel(),
                       np.arange(0, 28 * N + 1, 14, dtype=np.int64)),
                      shape=(2 * N, 3 + 8 * C + 3 * n_points))


def bundle_adjust_sharedcam(cams, pts, points_2d, camera_indices, point_indices, weights=None,
                            ftol=1e-6, verbose=0, **ls_kwargs):
    """The reference's ``bundleAdjust_sharedcam`` (pySBA.py:286-325) through scipy."""
    C, P = cams.shape[0], pts.shape[0]
    if weights is None:
        weights = default_weights(point_indices)
    weights = np.asarray(weights).reshape(-1, 1)
    A = sparsity_sharedcam(C, P, camera_indices, point_indices)
    return least_squares(fun_sharedcam, sharedcam_pack(cams, pts), jac_sparsity=A, verbose=verbose,
                         x_scale="jac", ftol=ftol, method="trf", jac="3-point",
                         args=(C, P, camera_indices, point_indices, points_2d, weights), **ls_kwargs)


def trf_exact_generic(fun_x, jac_x, x0, ftol, xtol=1e-8, gtol=1e-8, max_nfev=None):
    """``trf_no_bounds`` with an analytic sparse Jacobian ``jac_x(x)`` (CSR) and the exact
    solution of (J_h^T J_h + reg_term I) p = J_h^T f by sparse LU.  Same restatement as
    ``trf_exact`` for any parameterisation."""
    from numpy.linalg import norm
    from scipy.linalg import qr
    from scipy.optimize._lsq.common import (check_termination, minimize_quadratic_1d,
                                            solve_trust_region_2d, update_tr_radius)
    from scipy.sparse import diags, identity
    from scipy.sparse.linalg import splu
    x = np.asarray(x0, dtype=np.float64).copy()
    f = fun_x(x)
    if not np.all(np.isfinite(f)):
        raise ValueError("Residuals are not finite in the initial point.")
    J = jac_x(x)
    nfev = njev = 1
    cost = 0.5 * f @ f
    g = J.T @ f
    scale_inv = np.sqrt(np.asarray(J.power(2).sum(axis=0)).ravel())
    scale_inv[scale_inv == 0] = 1
    Delta = norm(x * scale_inv) or 1.0
    if max_nfev is None:
        max_nfev = x.size * 100
    status, trace = None, []
    while True:
        g_norm = norm(g, ord=np.inf)
        if g_norm < gtol:
            status = 1
        trace.append(cost)
        if status is not None or nfev == max_nfev:
            break
        d = 1 / scale_inv
        Jh = J @ diags(d)
        g_h = d * g
        Jg = Jh @ g_h
        reg = -minimize_quadratic_1d(0.5 * Jg @ Jg, -g_h @ g_h, 0, Delta / norm(g_h))[1] / Delta**2
        gn_h = splu((Jh.T @ Jh + reg * identity(x.size)).tocsc()).solve(Jh.T @ f)
        Sq, _ = qr(np.vstack((g_h, gn_h)).T, mode="economic")
        JS = Jh @ Sq
        B_S, g_S = JS.T @ JS, Sq.T @ g_h
        actual = -1
        while actual <= 0 and nfev < max_nfev:
            p_S, _ = solve_trust_region_2d(B_S, g_S, Delta)
            step_h = Sq @ p_S
            Js = JS @ p_S
            pred = -(0.5 * Js @ Js + step_h @ g_h)
            x_new = x + d * step_h
            f_new = fun_x(x_new)
            nfev += 1
            shn = norm(step_h)
            if not np.all(np.isfinite(f_new)):
                Delta = 0.25 * shn
                continue
            cost_new = 0.5 * f_new @ f_new
            actual = cost - cost_new
            Delta_new, ratio = update_tr_radius(Delta, actual, pred, shn, shn > 0.95 * Delta)
            status = check_termination(actual, cost, norm(d * step_h), norm(x), ratio, ftol, xtol)
            if status is not None:
                break
            Delta = Delta_new
        if actual > 0:
            x, f, cost = x_new, f_new, cost_new
            J = jac_x(x)
            njev += 1
            g = J.T @ f
            scale_inv = np.maximum(np.sqrt(np.asarray(J.power(2).sum(axis=0)).ravel()), scale_inv)
    return OptimizeResult(x=x, cost=cost, fun=f, grad=g, optimality=g_norm, nfev=nfev, njev=njev,
                          status=status or 0, trace=trace)


def trf_exact_sharedcam(cams, pts, points_2d, camera_indices, point_indices, weights=None,
                        ftol=1e-6, **kw):
    """Exact-solve TRF in the shared-intrinsics parameterisation: J_red = J_full T."""
    C, P = cams.shape[0], pts.shape[0]
    if weights is None:
        weights = default_weights(point_indices)
    weights = np.asarray(weights).reshape(-1, 1)
    args = (C, P, camera_indices, point_indices, points_2d, weights)
    A = sparsity_sharedcam(C, P, camera_indices, point_indices)

    def jac(x):
        cm, pt = sharedcam_unpack(x, C, P)
        _, Jc, Jp = jacobian_blocks(cm, pt, camera_indices, point_indices, weights)
        data = np.concatenate([Jc[:, :, 6:9], Jc[:, :, 0:6], Jc[:, :, 9:11], Jp], axis=2).reshape(-1)
        return csr_matrix((data, A.indices, A.indptr), shape=A.shape)

    return trf_exact_generic(lambda x: fun_sharedcam(x, *args), jac, sharedcam_pack(cams, pts),
                             ftol, **kw)


# ------------------------------------------- squared-residual dense variants
def fun_camonly(params, n_cameras, camera_indices, point_indices, points_2d, weights, points_3d):
    """pySBA.py:151-156: cameras only, points fixed, residual = w * (proj - obs)**2."""
    cams = params.reshape(n_cameras, NCP)
    uv = project(points_3d[point_indices], cams[camera_indices])
    return (weights * (uv - points_2d) ** 2).ravel()


def bundle_adjust_camonly(cams, pts, points_2d, camera_indices, point_indices, weights=None,
                          ftol=1e-4, verbose=0, **ls_kwargs):
    """pySBA.py:160-173: dense least_squares, every option but ftol at scipy's default
    (2-point Jacobian, tr_solver 'exact', x_scale 1).  Returns (res, cams_out)."""
    C = cams.shape[0]
    if weights is None:
        weights = default_weights(point_indices)
    weights = np.asarray(weights).reshape(-1, 1)
    res = least_squares(fun_camonly, cams.ravel(), verbose=verbose, ftol=ftol, method="trf",
                        args=(C, camera_indices, point_indices, points_2d, weights, pts),
                        **ls_kwargs)
    return res, res.x.reshape(C, NCP)


def apply_transform(theta, pts):
    """Rows of [A | t] (3x4, row-major 12-vector) applied to points (pySBA.py:177-184)."""
    M = np.asarray(theta).reshape(3, 4)
    return pts @ M[:, :3].T + M[:, 3]


def fun_transform_points_3d(theta, cams, camera_indices, point_indices, points_2d, weights,
                            points_3d):
    """pySBA.py:176-188: one 12-parameter affine map of all points, cameras fixed,
    residual = w * (proj - obs)**2."""
    X = apply_transform(theta, points_3d)
    uv = project(X[point_indices], cams[camera_indices])
    return (weights * (uv - points_2d) ** 2).ravel()


def bundle_adjust_transform_points_3d(cams, pts, points_2d, camera_indices, point_indices,
                                      weights=None, ftol=1e-3, verbose=0, **ls_kwargs):
    """pySBA.py:191-206.  Returns (res, transformed points)."""
    if weights is None:
        weights = default_weights(point_indices)
    weights = np.asarray(weights).reshape(-1, 1)
    x0 = np.hstack((np.eye(3), np.zeros((3, 1)))).ravel()
    res = least_squares(fun_transform_points_3d, x0, verbose=verbose, ftol=ftol, method="trf",
                        args=(cams, camera_indices, point_indices, points_2d, weights, pts),
                        **ls_kwargs)
    return res, apply_transform(res.x, pts)


def sq_normal_camonly(cams, pts, points_2d, camera_indices, point_indices, weights=None):
    """Analytic (cost, g (11C), H (C,11,11)) of fun_camonly: f = w e^2, J = 2 w e dproj."""
    C = cams.shape[0]
    w = np.ones(camera_indices.size) if weights is None else np.asarray(weights, float).reshape(-1)
    uv, Jc, _ = jacobian_blocks(cams, pts, camera_indices, point_indices)
    e = uv - points_2d
    f = w[:, None] * e * e
    J = 2.0 * (w[:, None] * e)[:, :, None] * Jc
    H = np.zeros((C, NCP, NCP))
    g = np.zeros((C, NCP))
    np.add.at(H, camera_indices, np.einsum("nia,nib->nab", J, J))
    np.add.at(g, camera_indices, np.einsum("nia,ni->na", J, f))
    return 0.5 * np.sum(f * f), g.ravel(), H


def sq_normal_transform(theta, cams, pts, points_2d, camera_indices, point_indices, weights=None):
    """Analytic (cost, g (12), H (12,12)) of fun_transform_points_3d."""
    w = np.ones(camera_indices.size) if weights is None else np.asarray(weights, float).reshape(-1)
    X = apply_transform(theta, pts)
    uv, _, Jp = jacobian_blocks(cams, X, camera_indices, point_indices)
    e = uv - points_2d
    f = w[:, None] * e * e
    Xh = np.column_stack([pts, np.ones(pts.shape[0])])[point_indices]      # (N,4)
    # d proj_d / d theta[k,j] = Jp[d,k] * Xh[j]
    Jt = np.einsum("ndk,nj->ndkj", Jp, Xh).reshape(-1, 2, 12)
    J = 2.0 * (w[:, None] * e)[:, :, None] * Jt
    Jf = J.reshape(-1, 12)
    return 0.5 * np.sum(f * f), Jf.T @ f.ravel(), Jf.T @ Jf


def rmse_px(res_vec):
    """sqrt(mean(|r_i|^2)) over observations, r_i the 2-vector pixel residual."""
    r = np.asarray(res_vec).reshape(-1, 2)
    return float(np.sqrt(np.mean(np.sum(r * r, axis=1))))
