"""Host-side formats either side of the path against the reference's own
functions (golden: tests/golden/io_example17.npz, made from lasercalib/convert_params.py)."""
import numpy as np

from lasercalib_b200 import io as lio


def test_readable_and_red_formats_match_reference(golden):
    g = golden("io_example17")
    readable = [lio.sba_to_readable_format(c) for c in g["cams"]]
    for i, r in enumerate(readable):
        np.testing.assert_allclose(r["K"], g["K"][i], rtol=0, atol=0)
        np.testing.assert_allclose(r["R"], g["R"][i], rtol=0, atol=1e-14)
        np.testing.assert_array_equal(r["t"], g["cams"][i, 3:6])
        np.testing.assert_array_equal(r["d"], g["cams"][i, 7:9])
    np.testing.assert_allclose(lio.readable_to_red_format(readable), g["red"], rtol=0, atol=1e-13)


def test_camera_vector_from_calibration_matches_reference_loader(golden):
    g = golden("io_example17")
    for i in range(g["cams"].shape[0]):
        v = lio.camera_vector_from_calibration(g["camera_matrix"][i], g["distortion"][i],
                                               g["rc_ext"][i], g["tc_ext"][i])
        # rotation vectors near pi: compare the rotations, and the vector up to 1e-9
        np.testing.assert_allclose(v[3:], g["cams"][i, 3:], rtol=0, atol=0)
        np.testing.assert_allclose(v[:3], g["cams"][i, :3], rtol=0, atol=1e-9)


def test_concat_points_dataset_offsets():
    rng = np.random.default_rng(0)
    ds = []
    for n in (5, 7, 4):
        obs = n * 2
        ds.append(dict(n_cams=3, n_pts=n, points_3d=rng.normal(size=(n, 3)),
                       points_2d=rng.normal(size=(obs, 2)), camera_ind=rng.integers(0, 3, obs),
                       point_ind=np.repeat(np.arange(n), 2)))
    n_cams, p3, p2, ci, pi = lio.concat_points_dataset(ds[:2])
    assert n_cams == 3 and p3.shape == (12, 3) and p2.shape == (24, 2) and ci.shape == (24,)
    assert pi.max() == 11 and np.array_equal(pi[10:], np.repeat(np.arange(7), 2) + 5)
    # three datasets: the reference's offsets are [0, n0, n1] (calibrate_camera.py:41-44)
    _, _, _, _, pi_ref = lio.concat_points_dataset(ds)
    assert np.array_equal(pi_ref[24:], np.repeat(np.arange(4), 2) + 7)
    _, _, _, _, pi_fix = lio.concat_points_dataset(ds, cumulative_offsets=True)
    assert np.array_equal(pi_fix[24:], np.repeat(np.arange(4), 2) + 12)


def test_aruco_yaml_export_matches_reference(golden, tmp_path):
    """convert_params.py:105-113 (`readable_format_to_aruco_format`) and :63-83
    (`initialize_from_checkerboard`): files written here carry the reference's values exactly, the
    text of the first one is byte-identical to what the reference wrote through cv2.FileStorage,
    and loading them back gives the reference's 11-vectors."""
    g = golden("io_example17")
    names = [str(n) for n in g["cam_names"]]
    readable = [lio.sba_to_readable_format(c) for c in g["cams"]]
    root = str(tmp_path) + "/"
    lio.readable_format_to_aruco_format(root, len(names), readable, names)
    for i, n in enumerate(names):
        m = lio.read_opencv_yaml(root + n + ".yaml")
        np.testing.assert_allclose(m["camera_matrix"], g["aruco_camera_matrix"][i], rtol=0, atol=0)
        np.testing.assert_array_equal(m["distortion_coefficients"], g["aruco_distortion"][i])
        np.testing.assert_allclose(m["rc_ext"], g["aruco_rc_ext"][i], rtol=0, atol=1e-15)
        np.testing.assert_array_equal(m["tc_ext"], g["aruco_tc_ext"][i])
    R0 = g["R"][0]
    lio.write_opencv_yaml(root + "same.yaml", {
        "camera_matrix": g["K"][0].T, "distortion_coefficients": [g["cams"][0, 7], g["cams"][0, 8], 0, 0, 0],
        "rc_ext": R0.T, "tc_ext": g["cams"][0, 3:6]})
    assert open(root + "same.yaml").read() == str(g["aruco_text0"])
    back = lio.initialize_from_checkerboard(str(tmp_path), len(names), names)
    np.testing.assert_allclose(back, g["aruco_reloaded_cams"], rtol=0, atol=1e-9)
    cv2 = __import__("pytest").importorskip("cv2")
    fs = cv2.FileStorage(root + names[3] + ".yaml", cv2.FILE_STORAGE_READ)
    np.testing.assert_array_equal(fs.getNode("rc_ext").mat(), lio.read_opencv_yaml(root + names[3] + ".yaml")["rc_ext"])


def test_initialize_from_checkerboard_reads_opencv_files(golden, tmp_path):
    """The loader on files in the example's own (older OpenCV, %.16e) number format."""
    g = golden("io_example17")
    names = [str(n) for n in g["cam_names"]][:3]
    for i, n in enumerate(names):
        with open(tmp_path / (n + ".yaml"), "w") as f:
            f.write("%YAML:1.0\n---\n")
            for key, a in (("camera_matrix", g["camera_matrix"][i]), ("distortion_coefficients", g["distortion"][i]),
                           ("rc_ext", g["rc_ext"][i]), ("tc_ext", g["tc_ext"][i])):
                vals = ",\n       ".join("%.16e" % v for v in np.asarray(a).ravel())
                f.write("%s: !!opencv-matrix\n   rows: %d\n   cols: %d\n   dt: d\n   data: [ %s ]\n"
                        % (key, a.shape[0], a.shape[1], vals))
    cams = lio.initialize_from_checkerboard(str(tmp_path), 3, names)
    np.testing.assert_allclose(cams, g["init_cams"][:3], rtol=0, atol=1e-9)


class FakeSBA:      # stands in for PySBA on a CPU-only box: only cameraArray is read
    def __init__(self, cams):
        self.cameraArray = cams


def test_save_calibration_results_files(golden, tmp_path):
    """The result files of scripts/calibrate_camera.py:75-106 (pickles, red CSV, YAMLs)."""
    import pickle

    g = golden("io_example17")
    names = [str(n) for n in g["cam_names"]]
    out = tmp_path / "results"
    lio.save_calibration_results(FakeSBA(g["cams"]), str(out), names)
    cam_list = pickle.load(open(out / "calibration.pkl", "rb"))
    np.testing.assert_allclose(cam_list[5]["R"], g["R"][5], rtol=0, atol=1e-14)
    rows = open(out / "calibration_red.csv").read().splitlines()
    assert len(rows) == len(names) and all(r.endswith(",") and r.count(",") == 25 for r in rows)
    np.testing.assert_allclose(np.array([float(v) for v in rows[2].split(",")[:-1]]), g["red"][2], atol=5e-7)
    assert isinstance(pickle.load(open(out / "sba.pkl", "rb")), FakeSBA)
    assert sorted(p.name for p in (out / "calibration_aruco").iterdir()) == sorted(n + ".yaml" for n in names)
