"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; numpy + scipy are the only
imports of lasercalib/pySBA.py):

    python tests/golden/make_golden.py

The reference has no tests or recorded outputs, so these files are
what pins the oracle (``oracle/pysba_oracle.py``) and, through it, the CUDA engine.
Every file records the numpy / scipy versions that produced it.
"""
import contextlib
import io
import os
import sys

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)

from lasercalib.pySBA import PySBA  # noqa: E402  (the reference itself)
from lasercalib_b200.synth import make_rig  # noqa: E402

VERS = dict(numpy_version=np.__version__, scipy_version=scipy.__version__)


def inputs_of(pb):
    return dict(cams0=pb["cams0"], pts0=pb["pts0"], points_2d=pb["points_2d"],
                camera_ind=pb["camera_ind"], point_ind=pb["point_ind"],
                cams_gt=pb["cams_gt"], pts_gt=pb["pts_gt"])


def blocks_from_dense(J, C, cam_idx, pt_idx):
    """(2N, n) dense -> (N, 2, 14) in-pattern entries, plus max |out-of-pattern|."""
    N = cam_idx.size
    cols = np.concatenate([cam_idx[:, None] * 11 + np.arange(11),
                           C * 11 + pt_idx[:, None] * 3 + np.arange(3)], axis=1)
    rows = np.arange(N)
    out = np.empty((N, 2, 14))
    out[:, 0, :] = J[2 * rows[:, None], cols]
    out[:, 1, :] = J[2 * rows[:, None] + 1, cols]
    mask = np.ones(J.shape, dtype=bool)
    mask[2 * rows[:, None], cols] = False
    mask[2 * rows[:, None] + 1, cols] = False
    return out, float(np.abs(J[mask]).max())


def golden_model():
    """fun / project / pattern / Jacobians on a 4-cam x 300-pt planar rig."""
    pb = make_rig("ring4", 300, seed=0, variant="planar")
    C, P = pb["n_cams"], pb["n_points"]
    ci, pi, p2 = pb["camera_ind"], pb["point_ind"], pb["points_2d"]
    sba = PySBA(pb["cams0"].copy(), pb["pts0"].copy(), p2, ci, pi)
    x0 = np.hstack((pb["cams0"].ravel(), pb["pts0"].ravel()))
    args = (C, P, ci, pi, p2, sba.pointWeights)
    f0 = sba.fun(x0, *args)
    proj = sba.project(pb["pts0"][pi], pb["cams0"][ci])
    A = sba.bundle_adjustment_sparsity(C, P, ci, pi).tocsr()
    A.sort_indices()
    # float weights
    rng = np.random.default_rng(7)
    w = rng.uniform(0.25, 2.0, ci.size)
    sba_w = PySBA(pb["cams0"].copy(), pb["pts0"].copy(), p2, ci, pi, pointWeights=w)
    f0_w = sba_w.fun(x0, C, P, ci, pi, p2, sba_w.pointWeights)
    # scipy's own FD Jacobian of the reference fun
    from scipy.optimize._numdiff import approx_derivative
    Jfd = approx_derivative(sba.fun, x0, method="3-point", sparsity=A, args=args).toarray()
    Jfd_b, _ = blocks_from_dense(Jfd, C, ci, pi)
    # extended-precision truth from the reference fun (dtype-generic)
    ld = np.longdouble
    xl, p2l, wl = x0.astype(ld), p2.astype(ld), sba.pointWeights.astype(ld)
    argl = (C, P, ci, pi, p2l, wl)
    Jt = np.zeros((2 * ci.size, x0.size))

    def cd(j, h):
        e = np.zeros_like(xl)
        e[j] = h
        return (sba.fun(xl + e, *argl) - sba.fun(xl - e, *argl)) / (2 * h)

    for j in range(x0.size):
        h = ld(1e-5) * max(ld(1.0), abs(xl[j]))
        Jt[:, j] = ((4 * cd(j, h / 2) - cd(j, h)) / 3).astype(np.float64)
    Jt_b, off = blocks_from_dense(Jt, C, ci, pi)
    np.savez_compressed(os.path.join(HERE, "model_ring4_planar300.npz"),
                        f0=f0, proj=proj, f0_w=f0_w, weights=w,
                        A_indices=A.indices.astype(np.int32), A_indptr=A.indptr.astype(np.int32),
                        A_shape=np.array(A.shape), A_dtype=str(A.dtype),
                        J_fd_blocks=Jfd_b, J_truth_blocks=Jt_b, J_truth_offpattern_max=off,
                        **inputs_of(pb), **VERS)
    print("model_ring4_planar300: N", ci.size, "max|J|", np.abs(Jt_b).max(), "off-pattern", off)


def golden_example_cams():
    """The 17 real camera vectors of example/calib_init_2024_05_02 through the reference's
    own loader, and the reference's project()/rotate() on them."""
    import json
    from lasercalib.convert_params import initialize_from_checkerboard
    with open("/root/reference/example/config.json") as f:
        cfg = json.load(f)
    names = ["Cam" + s for s in cfg["cam_serials"]]
    have = [n for n in names
            if os.path.exists("/root/reference/example/calib_init_2024_05_02/%s.yaml" % n)]
    cams = initialize_from_checkerboard("/root/reference/example/calib_init_2024_05_02",
                                        len(have), have)
    rng = np.random.default_rng(3)
    M = 40
    pts = np.column_stack([rng.uniform(-700, 700, M * len(have)),
                           rng.uniform(-700, 700, M * len(have)),
                           rng.uniform(0, 600, M * len(have))])
    cam_rows = np.repeat(cams, M, axis=0)
    sba = PySBA(cams, pts, np.zeros((1, 2)), np.zeros(1, dtype=int), np.zeros(1, dtype=int))
    proj = sba.project(pts, cam_rows)
    rot = sba.rotate(pts, cam_rows[:, :3])
    # theta = 0 and tiny-theta rows
    rv = np.zeros((6, 3))
    rv[1] = [1e-9, 0, 0]
    rv[2] = [1e-5, -2e-5, 3e-5]
    rv[3] = [1e-3, 2e-3, -1e-3]
    rv[4] = [0.05, -0.02, 0.03]
    rv[5] = [np.pi, 0, 0]
    rot_small = sba.rotate(pts[:6], rv)
    cam_small = cam_rows[:6].copy()
    cam_small[:, :3] = rv
    cam_small[:, 3:6] = [10.0, -20.0, 1500.0]
    proj_small = sba.project(pts[:6], cam_small)
    np.savez_compressed(os.path.join(HERE, "model_example17.npz"), cams=cams, pts=pts,
                        cam_rows_per_point=M, proj=proj, rot=rot, rv_small=rv,
                        rot_small=rot_small, cam_small=cam_small, proj_small=proj_small,
                        cam_names=np.array(have), **VERS)
    print("model_example17:", len(have), "cameras; |rotvec| range",
          np.linalg.norm(cams[:, :3], axis=1).min(), np.linalg.norm(cams[:, :3], axis=1).max())


def run_ba(pb, ftol, **kw):
    """PySBA.bundleAdjust as shipped (pySBA.py:132-147); kw -> direct least_squares call with
    the same keyword set (PySBA exposes only ftol)."""
    from scipy.optimize import least_squares
    C, P = pb["n_cams"], pb["n_points"]
    sba = PySBA(pb["cams0"].copy(), pb["pts0"].copy(), pb["points_2d"], pb["camera_ind"],
                pb["point_ind"])
    buf = io.StringIO()
    xs = []
    with contextlib.redirect_stdout(buf):
        if not kw:
            res = sba.bundleAdjust(ftol)
        else:
            x0 = np.hstack((sba.cameraArray.ravel(), sba.points3D.ravel()))
            A = sba.bundle_adjustment_sparsity(C, P, sba.cameraIndices, sba.point2DIndices)
            res = least_squares(sba.fun, x0, jac_sparsity=A, verbose=2, x_scale="jac", ftol=ftol,
                                method="trf", jac="3-point",
                                callback=lambda intermediate_result: xs.append(
                                    intermediate_result.x.copy()),
                                args=(C, P, sba.cameraIndices, sba.point2DIndices, sba.points2D,
                                      sba.pointWeights), **kw)
    return res, buf.getvalue(), xs


def golden_ba(name, rig, npts, variant, p_vis=1.0):
    pb = make_rig(rig, npts, seed=0, variant=variant, p_vis=p_vis)
    res, log, _ = run_ba(pb, 1e-4)
    tight = dict(tr_options={"atol": 1e-14, "btol": 1e-14, "conlim": 1e16})
    rt, logt, xs = run_ba(pb, 1e-4, **tight)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        ref_x=res.x, ref_cost=res.cost, ref_nfev=res.nfev, ref_njev=res.njev,
        ref_status=res.status, ref_optimality=res.optimality, ref_grad=res.grad,
        ref_log=log, ref_message=res.message,
        tight_x=rt.x, tight_cost=rt.cost, tight_nfev=rt.nfev, tight_njev=rt.njev,
        tight_status=rt.status, tight_optimality=rt.optimality, tight_log=logt,
        tight_x1=xs[0], rig=rig, n_points_requested=npts, variant=variant, p_vis=p_vis,
        **inputs_of(pb), **VERS)
    print(name, "N", pb["n_obs"], "ref cost", res.cost, "nfev", res.nfev, "| tight cost", rt.cost,
          "nfev", rt.nfev)


def golden_nocam():
    """PySBA.bundleAdjust_nocam (points only, pySBA.py:237-250) on an 8-camera rig."""
    pb = make_rig("ring8", 600, seed=9, variant="volume", p_vis=0.8)
    sba = PySBA(pb["cams_gt"].copy(), pb["pts0"].copy(), pb["points_2d"], pb["camera_ind"],
                pb["point_ind"])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = sba.bundleAdjust_nocam(1e-7)
    np.savez_compressed(os.path.join(HERE, "ba_nocam_ring8_600.npz"), ref_x=res.x,
                        ref_cost=res.cost, ref_nfev=res.nfev, ref_status=res.status,
                        ref_log=buf.getvalue(), cams=pb["cams_gt"], **inputs_of(pb), **VERS)
    print("ba_nocam_ring8_600: cost", res.cost, "nfev", res.nfev, "status", res.status)


def golden_sharedcam():
    """PySBA.bundleAdjust_sharedcam (pySBA.py:286-325) on an 8-camera rig."""
    pb = make_rig("ring8", 500, seed=12, variant="volume", p_vis=0.8)
    sba = PySBA(pb["cams0"].copy(), pb["pts0"].copy(), pb["points_2d"], pb["camera_ind"],
                pb["point_ind"])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = sba.bundleAdjust_sharedcam(1e-6)
    np.savez_compressed(os.path.join(HERE, "ba_sharedcam_ring8_500.npz"), ref_x=res.x,
                        ref_cost=res.cost, ref_nfev=res.nfev, ref_status=res.status,
                        ref_cams=sba.cameraArray, ref_log=buf.getvalue(), **inputs_of(pb), **VERS)
    print("ba_sharedcam_ring8_500: cost", res.cost, "nfev", res.nfev, "status", res.status)


def golden_squared():
    """The two dense squared-residual variants (pySBA.py:151-206) through the reference:
    bundle_adjustment_camonly on an 8-camera rig (points at ground truth, cameras perturbed)
    and bundleAdjust_transform_points_3d (cameras at ground truth, points moved by a known
    affine map).  Also records fun at x0 so the restated residual functions are pinned."""
    pb = make_rig("ring8", 400, seed=21, variant="volume", p_vis=0.85)
    sba = PySBA(pb["cams0"].copy(), pb["pts_gt"].copy(), pb["points_2d"], pb["camera_ind"],
                pb["point_ind"])
    f0 = sba.fun_camonly(pb["cams0"].ravel(), 8, pb["n_points"], pb["camera_ind"], pb["point_ind"],
                         pb["points_2d"], sba.pointWeights, pb["pts_gt"])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = sba.bundle_adjustment_camonly(1e-4)
    np.savez_compressed(os.path.join(HERE, "ba_camonly_ring8_400.npz"), ref_x=res.x, ref_f0=f0,
                        ref_cost=res.cost, ref_nfev=res.nfev, ref_njev=res.njev,
                        ref_status=res.status, ref_cams=sba.cameraArray, ref_log=buf.getvalue(),
                        **inputs_of(pb), **VERS)
    print("ba_camonly_ring8_400: cost", res.cost, "nfev", res.nfev, "status", res.status)

    rng = np.random.default_rng(5)
    A = np.eye(3) + 0.02 * rng.normal(size=(3, 3))
    t = np.array([4.0, -3.0, 6.0])
    moved = (pb["pts_gt"] - t) @ np.linalg.inv(A).T          # A moved + t = pts_gt
    sba = PySBA(pb["cams_gt"].copy(), moved.copy(), pb["points_2d"], pb["camera_ind"],
                pb["point_ind"])
    x0 = np.hstack((np.eye(3), np.zeros((3, 1)))).ravel()
    f0 = sba.fun_transform_points_3d(x0, 8, pb["n_points"], pb["cams_gt"], pb["camera_ind"], pb["point_ind"],
                                     pb["points_2d"], sba.pointWeights, moved)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = sba.bundleAdjust_transform_points_3d(1e-3)
    np.savez_compressed(os.path.join(HERE, "ba_transform_ring8_400.npz"), ref_x=res.x, ref_f0=f0,
                        ref_cost=res.cost, ref_nfev=res.nfev, ref_njev=res.njev,
                        ref_status=res.status, ref_points=sba.points3D, moved=moved, A=A, t=t,
                        ref_log=buf.getvalue(), **inputs_of(pb), **VERS)
    print("ba_transform_ring8_400: cost", res.cost, "nfev", res.nfev, "status", res.status)


def golden_io():
    """Export / init formats of lasercalib/convert_params.py on the 17 example cameras."""
    import cv2
    np.NaN = np.nan            # the reference still uses np.NaN (removed in numpy 2)
    from lasercalib.convert_params import readable_to_red_format, sba_to_readable_format
    g = np.load(os.path.join(HERE, "model_example17.npz"))
    cams = g["cams"]
    readable = [sba_to_readable_format(c) for c in cams]
    red = readable_to_red_format(readable)
    raw = []
    for n in [str(x) for x in g["cam_names"]]:
        fs = cv2.FileStorage("/root/reference/example/calib_init_2024_05_02/%s.yaml" % n,
                             cv2.FILE_STORAGE_READ)
        raw.append((fs.getNode("camera_matrix").mat(), fs.getNode("distortion_coefficients").mat(),
                    fs.getNode("rc_ext").mat(), fs.getNode("tc_ext").mat()))
    # the reference's YAML export (convert_params.py:105-113) and its own loader on those files
    import tempfile
    from lasercalib.convert_params import initialize_from_checkerboard, readable_format_to_aruco_format
    names = [str(x) for x in g["cam_names"]]
    with tempfile.TemporaryDirectory() as td:
        readable_format_to_aruco_format(td + "/", len(names), readable, names)
        aruco = []
        for n in names:
            fs = cv2.FileStorage("%s/%s.yaml" % (td, n), cv2.FILE_STORAGE_READ)
            aruco.append((fs.getNode("camera_matrix").mat(), fs.getNode("distortion_coefficients").mat(),
                          fs.getNode("rc_ext").mat(), fs.getNode("tc_ext").mat()))
        aruco_text0 = open("%s/%s.yaml" % (td, names[0])).read()
        reloaded = initialize_from_checkerboard(td, len(names), names)
    init_cams = initialize_from_checkerboard("/root/reference/example/calib_init_2024_05_02", len(names), names)
    np.savez_compressed(os.path.join(HERE, "io_example17.npz"), cams=cams,
                        K=np.array([r["K"] for r in readable]), R=np.array([r["R"] for r in readable]),
                        red=red, camera_matrix=np.array([r[0] for r in raw]),
                        distortion=np.array([r[1] for r in raw]), rc_ext=np.array([r[2] for r in raw]),
                        tc_ext=np.array([r[3] for r in raw]),
                        aruco_camera_matrix=np.array([r[0] for r in aruco]),
                        aruco_distortion=np.array([r[1] for r in aruco]),
                        aruco_rc_ext=np.array([r[2] for r in aruco]), aruco_tc_ext=np.array([r[3] for r in aruco]),
                        aruco_text0=np.array(aruco_text0), aruco_reloaded_cams=reloaded, init_cams=init_cams,
                        cam_names=np.array(names), **VERS)


def golden_unproject():
    """lasercalib/rigid_body.py:205-243 Unproject on three example cameras (+ a strongly
    distorted variant with tangential terms and k3)."""
    from lasercalib.rigid_body import Unproject
    g = np.load(os.path.join(HERE, "io_example17.npz"))
    rng = np.random.default_rng(5)
    outs, pts_all, Zs = [], [], []
    for i in (0, 8, 16):
        pts = np.column_stack([rng.uniform(0, 3208, 200), rng.uniform(0, 2200, 200)])
        Z = rng.choice([0.0, 106.0], 200)
        outs.append(Unproject(pts, Z, g["camera_matrix"][i], g["distortion"][i], g["rc_ext"][i],
                              g["tc_ext"][i]))
        pts_all.append(pts)
        Zs.append(Z)
    d2 = np.array([[-0.1], [0.02], [1e-3], [-5e-4], [3e-3]])
    pts = np.column_stack([rng.uniform(0, 3208, 200), rng.uniform(0, 2200, 200)])
    out2 = Unproject(pts, [106.0], g["camera_matrix"][0], d2, g["rc_ext"][0], g["tc_ext"][0])
    np.savez_compressed(os.path.join(HERE, "unproject_example.npz"), cam_ids=np.array([0, 8, 16]),
                        pts=np.array(pts_all), Z=np.array(Zs), out=np.array(outs), d2=d2, pts2=pts,
                        out2=out2, **VERS)


if __name__ == "__main__":
    only = {"--only-unproject": golden_unproject, "--only-sharedcam": golden_sharedcam,
            "--only-io": golden_io, "--only-nocam": golden_nocam, "--only-squared": golden_squared}
    picked = [fn for flag, fn in only.items() if flag in sys.argv]
    if picked:
        for fn in picked:
            fn()
        sys.exit(0)
    golden_nocam()
    golden_sharedcam()
    golden_squared()
    golden_model()
    golden_example_cams()
    golden_ba("ba_ring4_planar2000", "ring4", 2000, "planar")
    golden_ba("ba_ring8_volume1500", "ring8", 1500, "volume")
    golden_ba("ba_example18_vis60_800", "example18", 800, "volume", p_vis=0.6)
    golden_io()
    golden_unproject()
