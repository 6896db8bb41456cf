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


# ---- OpenCV-YAML camera files (the hand-off format either side of the path) ------------------
# The reference reads its initial cameras from per-camera OpenCV FileStorage YAMLs
# (convert_params.py:63-83) and writes the adjusted ones back in the same format
# (convert_params.py:105-113) for the ArUco stage.  The files are `%YAML:1.0` documents of
# `!!opencv-matrix` nodes (rows, cols, dt: d, row-major data).  This writer/reader handles exactly
# that subset without importing OpenCV; tests check it against cv2.FileStorage where cv2 exists.

def _fmt_cv(v):
    """cv2.FileStorage's number text: integers as '1.', everything else '%.17g'."""
    v = float(v)
    if np.isnan(v):
        return ".Nan"
    if np.isinf(v):
        return ".Inf" if v > 0 else "-.Inf"
    if v == int(v) and abs(v) < 2 ** 31:
        return "%d." % int(v)
    return "%.17g" % v


def write_opencv_yaml(path, matrices):
    """Write `matrices` (ordered mapping name -> array; 1-D arrays become column vectors, as
    cv2.FileStorage.write stores them) as an OpenCV FileStorage YAML 1.0 document, with OpenCV's
    own number format and line wrapping (flow sequence, wrap margin 71, 7-space continuation)."""
    lines = ["%YAML:1.0", "---"]
    for name, m in matrices.items():
        a = np.asarray(m, dtype=np.float64)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        lines += ["%s: !!opencv-matrix" % name, "   rows: %d" % a.shape[0], "   cols: %d" % a.shape[1],
                  "   dt: d"]
        cur = "   data: ["
        for i, v in enumerate(a.ravel()):
            t = _fmt_cv(v)
            if i > 0:
                cur += ","
            if i > 0 and len(cur) + len(t) > 71:
                lines.append(cur)
                cur = "       " + t
            else:
                cur += " " + t
        lines.append(cur + " ]")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def read_opencv_yaml(path):
    """Parse the `!!opencv-matrix` nodes of an OpenCV FileStorage YAML into {name: ndarray}."""
    import re
    text = open(path).read()
    out = {}
    pat = re.compile(r"^(\w+):\s*!!opencv-matrix\s*\n\s*rows:\s*(\d+)\s*\n\s*cols:\s*(\d+)\s*\n\s*dt:\s*(\w+)\s*\n"
                     r"\s*data:\s*\[(.*?)\]", re.S | re.M)
    for name, rows, cols, dt, data in pat.findall(text):
        vals = [float(t.replace(".Nan", "nan").replace("-.Inf", "-inf").replace(".Inf", "inf"))
                for t in data.replace("\n", " ").split(",") if t.strip()]
        out[name] = np.array(vals, dtype=np.float64).reshape(int(rows), int(cols))
    return out


def initialize_from_checkerboard(filedir, nCams, cam_names):
    """Per-camera YAMLs (`camera_matrix`, `distortion_coefficients`, `rc_ext`, `tc_ext`) -> the
    (nCams, 11) cameraArray PySBA starts from; same name, arguments and result as the reference's
    loader (convert_params.py:63-83)."""
    cams = np.zeros((nCams, 11))
    for i in range(nCams):
        m = read_opencv_yaml("%s/%s.yaml" % (filedir, cam_names[i]))
        cams[i] = camera_vector_from_calibration(m["camera_matrix"], m["distortion_coefficients"],
                                                 m["rc_ext"], m["tc_ext"])
    return cams


def readable_format_to_aruco_format(save_root, nCams, camList, cam_names):
    """Readable dicts -> one OpenCV YAML per camera for the ArUco stage: camera_matrix = K^T,
    distortion [k1, k2, 0, 0, 0], rc_ext = R^T, tc_ext = t (convert_params.py:105-113; `save_root`
    is a prefix the file name is appended to, as in the reference)."""
    for i in range(nCams):
        p = camList[i]
        write_opencv_yaml(save_root + "%s.yaml" % cam_names[i], {
            "camera_matrix": np.asarray(p["K"]).T,
            "distortion_coefficients": np.asarray([p["d"][0], p["d"][1], 0.0, 0.0, 0.0]),
            "rc_ext": np.asarray(p["R"]).T,
            "tc_ext": np.asarray(p["t"])})


def save_calibration_results(sba, results_dir, cam_names):
    """Everything `scripts/calibrate_camera.py:75-106` writes after `sba.bundleAdjust`:
    calibration.pkl (list of readable dicts), calibration_red.csv (25 columns, '%f', every row
    ending in ','), sba.pkl (the PySBA object itself) and calibration_aruco/<cam>.yaml."""
    import os
    import pickle
    n_cams = np.asarray(sba.cameraArray).shape[0]
    cam_list = [sba_to_readable_format(np.asarray(sba.cameraArray)[i, :]) for i in range(n_cams)]
    os.makedirs(results_dir, exist_ok=True)
    with open(os.path.join(results_dir, "calibration.pkl"), "wb") as f:
        pickle.dump(cam_list, f)
    np.savetxt(os.path.join(results_dir, "calibration_red.csv"), readable_to_red_format(cam_list),
               delimiter=",", newline=",\n", fmt="%f")
    with open(os.path.join(results_dir, "sba.pkl"), "wb") as f:
        pickle.dump(sba, f)
    save_root = os.path.join(results_dir, "calibration_aruco") + "/"
    os.makedirs(save_root, exist_ok=True)
    readable_format_to_aruco_format(save_root, n_cams, cam_list, cam_names)
    return cam_list
