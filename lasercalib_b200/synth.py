"""Synthetic multi-camera laser-calibration rigs (data generator for tests and bench).

The reference ships no laser centroids (its ``.gitignore`` excludes results/), so
every parity and benchmark input is generated here from the rig geometry the
reference documents:

* camera vector layout ``[rotvec(3), t(3), f, k1, k2, cx, cy]``
  (reference ``lasercalib/pySBA.py:31-35``),
* arena footprint and the 65 MP sensor (reference ``scripts/65MP.py:54,67-70``),
* image size 3208x2200 (reference ``scripts/timeline_rerun.py:92``),
* calibration planes z in {0, 106} mm and ``min_num_cam_per_point`` = 4
  (reference ``example/config.json:8-11,23``),
* observation order: point-major, camera-ascending, int64 indices
  (reference ``scripts/get_points3d.py:74-86``), visibility filter
  ">= min cams and seen by the reference camera" (``get_points3d.py:52-56``).

This module is a *generator*: it has its own small numpy pinhole model so that it
neither imports the oracle nor the CUDA engine.
"""
from __future__ import annotations

import numpy as np

CAM_PARAMS = 11

# name -> rig description
RIGS = {
    # config 1: the reference's own CPU-runnable case
    "ring4": dict(rings=[dict(n=4, radius=1500.0, height=1700.0)],
                  f=2400.0, k1=1e-3, k2=-1e-2, cx=1604.0, cy=1100.0,
                  image=(3208, 2200), target_z=0.0),
    # config 2: example/config.json style layout: 8 upper, 8 lower, 2 overhead
    "example18": dict(rings=[dict(n=8, radius=(860.0, 1300.0), height=(1520.0, 1715.0)),
                             dict(n=8, radius=(1627.0, 1649.0), height=(468.0, 506.0)),
                             dict(n=1, radius=117.0, height=1913.0, f=1774.0),
                             dict(n=1, radius=60.0, height=1900.0, f=5200.0,
                                  image=(9344, 7000), cx=4672.0, cy=3500.0)],
                      f=(1774.0, 2777.0), k1=(-0.014, 0.010), k2=(-0.051, -0.0004),
                      cx=(1539.0, 1640.0), cy=(1059.0, 1165.0),
                      image=(3208, 2200), target_z=50.0),
    # config 3: two rings of 12
    "ring24": dict(rings=[dict(n=12, radius=1100.0, height=1600.0),
                          dict(n=12, radius=1640.0, height=500.0, phase=0.5)],
                   f=(1774.0, 2777.0), k1=(-0.014, 0.010), k2=(-0.051, -0.0004),
                   cx=(1539.0, 1640.0), cy=(1059.0, 1165.0),
                   image=(3208, 2200), target_z=50.0),
    # config 4: 8 x 65 MP wide-angle
    "wide8": dict(rings=[dict(n=8, radius=600.0, height=800.0)],
                  f=4500.0, k1=-0.10, k2=0.01, cx=4672.0, cy=3500.0,
                  image=(9344, 7000), target_z=0.0),
    # config 5: four rings of 16
    "ring64": dict(rings=[dict(n=16, radius=1000.0, height=1700.0),
                          dict(n=16, radius=1300.0, height=1300.0, phase=0.5),
                          dict(n=16, radius=1500.0, height=900.0),
                          dict(n=16, radius=1640.0, height=500.0, phase=0.5)],
                   f=(1774.0, 2777.0), k1=(-0.014, 0.010), k2=(-0.051, -0.0004),
                   cx=(1539.0, 1640.0), cy=(1059.0, 1165.0),
                   image=(3208, 2200), target_z=50.0),
    # small well-conditioned rig used by convergence tests
    "ring8": dict(rings=[dict(n=8, radius=1500.0, height=1500.0)],
                  f=2400.0, k1=1e-3, k2=-1e-2, cx=1604.0, cy=1100.0,
                  image=(3208, 2200), target_z=100.0),
}


def _draw(rng, spec):
    if isinstance(spec, tuple):
        return float(rng.uniform(spec[0], spec[1]))
    return float(spec)


def rotmat_to_rotvec(R):
    """Rotation matrix -> rotation vector (angle * axis), robust near pi."""
    R = np.asarray(R, dtype=np.float64)
    cos_t = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    theta = np.arccos(cos_t)
    if theta < 1e-12:
        return np.zeros(3)
    if np.pi - theta < 1e-6:
        # near pi: axis from the symmetric part
        A = (R + np.eye(3)) * 0.5
        i = int(np.argmax(np.diag(A)))
        v = A[:, i] / np.sqrt(A[i, i])
        w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
        if np.dot(w, v) < 0:
            v = -v
        return theta * v / np.linalg.norm(v)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return theta * w / (2.0 * np.sin(theta))


def rotvec_to_rotmat(r):
    r = np.asarray(r, dtype=np.float64)
    theta = np.linalg.norm(r)
    if theta == 0.0:
        return np.eye(3)
    v = r / theta
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + np.sin(theta) * K + (1 - np.cos(theta)) * (K @ K)


def look_at_camera(center, target):
    """world->camera rotation R (rows x,y,z of the camera frame) and t = -R C."""
    center = np.asarray(center, dtype=np.float64)
    z = np.asarray(target, dtype=np.float64) - center
    z /= np.linalg.norm(z)
    x = np.cross(np.array([0.0, 0.0, 1.0]), z)
    if np.linalg.norm(x) < 1e-6:
        x = np.array([1.0, 0.0, 0.0])
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.vstack([x, y, z])
    return R, -R @ center


def project_one_camera(points, cam):
    """Pinhole + 2-term radial model for ONE camera vector over (P,3) points.

    Same model as the reference ``PySBA.project`` (``pySBA.py:76-89``) written with
    a rotation matrix; used only to synthesise observations.
    """
    R = rotvec_to_rotmat(cam[:3])
    Pc = points @ R.T + cam[3:6]
    z = Pc[:, 2]
    x = Pc[:, 0] / z
    y = Pc[:, 1] / z
    n = x * x + y * y
    d = 1.0 + cam[7] * n + cam[8] * n * n
    uv = np.empty((points.shape[0], 2))
    uv[:, 0] = x * d * cam[6] + cam[9]
    uv[:, 1] = y * d * cam[6] + cam[10]
    return uv, z


def make_cameras(rig, rng):
    spec = RIGS[rig] if isinstance(rig, str) else rig
    cams, images = [], []
    for ring in spec["rings"]:
        n = ring["n"]
        phase = ring.get("phase", 0.0)
        for i in range(n):
            ang = 2.0 * np.pi * (i + phase) / n
            rad = _draw(rng, ring["radius"])
            h = _draw(rng, ring["height"])
            C = np.array([rad * np.cos(ang), rad * np.sin(ang), h])
            R, t = look_at_camera(C, (0.0, 0.0, spec["target_z"]))
            cam = np.empty(CAM_PARAMS)
            cam[:3] = rotmat_to_rotvec(R)
            cam[3:6] = t
            cam[6] = _draw(rng, ring.get("f", spec["f"]))
            cam[7] = _draw(rng, spec["k1"])
            cam[8] = _draw(rng, spec["k2"])
            cam[9] = _draw(rng, ring.get("cx", spec["cx"]))
            cam[10] = _draw(rng, ring.get("cy", spec["cy"]))
            cams.append(cam)
            images.append(ring.get("image", spec["image"]))
    return np.asarray(cams), np.asarray(images, dtype=np.float64)


def make_rig(rig="ring4", n_points=10_000, seed=0, variant="volume", p_vis=1.0,
             noise_px=0.3, min_cams=4, ref_cam=0, perturb=True):
    """Build one synthetic bundle-adjustment problem.

    Returns a dict with ground truth (``cams_gt``, ``pts_gt``), the perturbed initial
    guess (``cams0``, ``pts0``) and the observation arrays in the reference's wire
    format (``points_2d`` (N,2) f64, ``camera_ind`` (N,) i64, ``point_ind`` (N,) i64).
    """
    rng = np.random.default_rng(seed)
    cams_gt, images = make_cameras(rig, rng)
    C = cams_gt.shape[0]
    P = int(n_points)
    pts = np.empty((P, 3))
    pts[:, 0] = rng.uniform(-700.0, 700.0, P)
    pts[:, 1] = rng.uniform(-700.0, 700.0, P)
    if variant == "planar":
        pts[:, 2] = np.where(rng.random(P) < 0.5, 0.0, 106.0)
    elif variant == "volume":
        pts[:, 2] = rng.uniform(0.0, 600.0, P)
    else:
        raise ValueError("variant must be 'planar' or 'volume'")

    uv_all = np.empty((P, C, 2))
    vis = np.empty((P, C), dtype=bool)
    for c in range(C):
        uv, z = project_one_camera(pts, cams_gt[c])
        uv_all[:, c, :] = uv
        inside = (uv[:, 0] >= 0) & (uv[:, 0] < images[c, 0]) & \
                 (uv[:, 1] >= 0) & (uv[:, 1] < images[c, 1]) & (z > 0)
        vis[:, c] = inside
    if p_vis < 1.0:
        vis &= rng.random((P, C)) < p_vis
    keep = (vis.sum(axis=1) >= min(min_cams, C)) & vis[:, ref_cam]
    pts = pts[keep]
    vis = vis[keep]
    uv_all = uv_all[keep]
    P = pts.shape[0]

    point_ind, camera_ind = np.nonzero(vis)          # row-major: point-major, cam ascending
    points_2d = uv_all[point_ind, camera_ind, :]
    points_2d = points_2d + rng.normal(0.0, noise_px, points_2d.shape)

    cams0 = cams_gt.copy()
    pts0 = pts.copy()
    if perturb:
        cams0[:, :3] += rng.normal(0.0, 1e-3, (C, 3))
        cams0[:, 3:6] += rng.normal(0.0, 1.0, (C, 3))
        cams0[:, 6] += rng.normal(0.0, 5.0, C)
        pts0 += rng.normal(0.0, 2.0, pts0.shape)

    return dict(rig=rig if isinstance(rig, str) else "custom", seed=seed, variant=variant,
                p_vis=p_vis, n_cams=C, n_points=P, n_obs=int(point_ind.size),
                cams_gt=cams_gt, pts_gt=pts, cams0=cams0, pts0=pts0,
                points_2d=np.ascontiguousarray(points_2d),
                camera_ind=camera_ind.astype(np.int64), point_ind=point_ind.astype(np.int64),
                images=images)


def shuffle_observations(problem, seed=1):
    """Return a copy with observations in random order (tests the ingest sort)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(problem["n_obs"])
    out = dict(problem)
    out["points_2d"] = np.ascontiguousarray(problem["points_2d"][perm])
    out["camera_ind"] = problem["camera_ind"][perm].copy()
    out["point_ind"] = problem["point_ind"][perm].copy()
    return out
