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
