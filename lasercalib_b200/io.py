"""Host-side data formats either side of the bundle-adjustment path (SURVEY.md section 8f,
rows 2 and 3).  Tiny numpy code, no GPU: these only define WHAT the engine ingests and what the
downstream tools consume.

* ingest:  ``points_dataset.pkl`` (list of per-laser-dataset dicts written by the reference's
  ``scripts/get_points3d.py:102-127``) -> the five arrays ``PySBA`` takes, concatenated exactly as
  ``scripts/calibrate_camera.py:32-44`` does;
* export:  11-vector -> {K, R, t, d} (``lasercalib/convert_params.py:18-27``) -> 25-column "red"
  CSV rows (``convert_params.py:7-16``);
* init:    (camera_matrix, distortion, rc_ext, tc_ext) -> 11-vector
  (``convert_params.py:76-82``, the numeric part of ``initialize_from_checkerboard``).
"""
from __future__ import annotations

import numpy as np

from .synth import rotmat_to_rotvec, rotvec_to_rotmat


def concat_points_dataset(points_dataset, cumulative_offsets=False):
    """Concatenate the per-laser datasets.  The reference offsets the point indices of dataset i
    by ``n_pts`` of dataset i-1 only (``calibrate_camera.py:41-44``), which is the cumulative
    offset for at most two datasets; ``cumulative_offsets=True`` gives the intended prefix sum."""
    n_cams = points_dataset[0]["n_cams"]
    points_3d = np.vstack([d["points_3d"] for d in points_dataset])
    points_2d = np.vstack([d["points_2d"] for d in points_dataset])
    camera_ind = np.hstack([d["camera_ind"] for d in points_dataset])
    offsets = [0]
    for i in range(len(points_dataset) - 1):
        n = points_dataset[i]["n_pts"]
        offsets.append(offsets[-1] + n if cumulative_offsets else n)
    point_ind = np.hstack([d["point_ind"] + offsets[i] for i, d in enumerate(points_dataset)])
    return n_cams, points_3d, points_2d, camera_ind, point_ind


def sba_to_readable_format(cam_vec):
    """11-vector -> dict(K, R, t, d).  K is stored transposed (cx, cy in the last ROW) and
    R = R(-rotvec) = R(rotvec)^T, as the reference keeps them (convert_params.py:18-27)."""
    cam_vec = np.asarray(cam_vec, dtype=np.float64)
    K = np.zeros((3, 3))
    K[0, 0] = K[1, 1] = cam_vec[6]
    K[2, 2] = 1.0
    K[2, :2] = cam_vec[9:11]
    return {"K": K, "R": rotvec_to_rotmat(-cam_vec[:3]), "t": cam_vec[3:6], "d": cam_vec[7:9]}


def readable_to_red_format(cam_list):
    """List of readable dicts -> (n, 25) rows [K^T (9), R^T (9), t (3), d (2), 0, 0]
    (convert_params.py:7-16)."""
    out = np.full((len(cam_list), 25), np.nan)
    for i, p in enumerate(cam_list):
        out[i] = np.hstack((np.transpose(p["K"]).ravel(), np.transpose(p["R"]).ravel(), p["t"],
                            p["d"], [0.0, 0.0]))
    return out


def camera_vector_from_calibration(camera_matrix, distortion, rc_ext, tc_ext):
    """OpenCV-style single-camera calibration -> the 11-vector PySBA uses
    (convert_params.py:76-82): rotvec(rc_ext), tc_ext, fx, k1, k2, cx, cy."""
    v = np.empty(11)
    v[0:3] = rotmat_to_rotvec(np.asarray(rc_ext, dtype=np.float64))
    v[3:6] = np.asarray(tc_ext, dtype=np.float64).ravel()[:3]
    d = np.asarray(distortion, dtype=np.float64).ravel()
    v[6:9] = [camera_matrix[0][0], d[0], d[1]]
    v[9:11] = [camera_matrix[0][2], camera_matrix[1][2]]
    return v


def Unproject(points, Z, intrinsic, distortion, rotation_matrix, tvec, engine=None):
    """Pixels of one camera -> world points on the plane(s) z = Z, on the GPU
    (``lasercalib/rigid_body.py:205-243``: OpenCV undistortPoints + ray/plane intersection;
    same name and argument order as the reference).  ``Z``: one number or [num_pts]."""
    from . import _cabi
    eng = engine or _cabi.Engine()
    try:
        return eng.unproject(points, Z, intrinsic, distortion, rotation_matrix, tvec)
    finally:
        if engine is None:
            eng.close()
