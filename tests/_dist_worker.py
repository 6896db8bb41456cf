"""Worker of tests/test_dist_gloo.py: world_size-N gloo job on CPU exercising the host-side
sharding / gather / reduction logic of lasercalib_b200.dist with the oracle supplying the
per-shard arithmetic (the CUDA engine does the same sums with NCCL on the GPUs)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from lasercalib_b200 import dist as D  # noqa: E402
from lasercalib_b200.synth import make_rig, shuffle_observations  # noqa: E402
from oracle import pysba_oracle as O  # noqa: E402


def main():
    rank, ws, _ = D.init_from_env(backend="gloo")
    assert ws == int(os.environ["WORLD_SIZE"]) and ws > 1
    pb = make_rig("ring8", 400, seed=5, variant="volume", p_vis=0.7)
    if os.environ.get("LCBA_TEST_SHUFFLE") == "1":
        pb = shuffle_observations(pb, seed=2)
    C, P = pb["n_cams"], pb["n_points"]
    ci, pi = pb["camera_ind"], pb["point_ind"]
    lam = 1e-5
    # ---- unsharded oracle ----
    w = O.default_weights(pi)
    x0 = np.hstack((pb["cams0"].ravel(), pb["pts0"].ravel()))
    f = O.fun(x0, C, P, ci, pi, pb["points_2d"], w)
    _, Jc, Jp = O.jacobian_blocks(pb["cams0"], pb["pts0"], ci, pi, w)
    U, gc, V, gp, W = O.normal_blocks(f.reshape(-1, 2), Jc, Jp, C, P, ci, pi)
    sc = np.hstack((np.sqrt(np.einsum("caa->ca", U)).ravel(), np.sqrt(np.einsum("paa->pa", V)).ravel()))
    S_ref, rhs_ref, _ = O.reduced_camera_system(U, gc, V, gp, W, ci, pi, lam, sc)
    # ---- this rank's shard ----
    sh = D.shard_problem(pb["pts0"], pb["points_2d"], ci, pi, None, rank, ws, collective=True)
    ref_sh = D.shard_problem(pb["pts0"], pb["points_2d"], ci, pi, None, rank, ws)          # full local check
    assert sh["lo"] == ref_sh["lo"] and sh["hi"] == ref_sh["hi"] and np.array_equal(sh["bounds"], ref_sh["bounds"])
    # the order check split over the ranks: one verdict on every rank, wherever the violation sits
    srt = np.sort(pi)
    assert D._is_point_major(srt, collective=True) and D._is_point_major(srt)
    n = srt.size
    for pos in (1, n // ws, n // ws - 1, (n * (ws - 1)) // ws, n - 1):     # piece interiors and piece borders
        bad = srt.copy()
        bad[pos - 1] = bad[pos] + 1
        assert not D._is_point_major(bad) and not D._is_point_major(bad, collective=True), pos
    b = sh["bounds"]
    sh["point_ind"] = sh["point_ind"] - sh["pt_offset"]        # local indices for the oracle
    assert b[0] == 0 and b[-1] == P and np.all(np.diff(b) >= 0)
    Pl = sh["pts"].shape[0]
    wl = O.default_weights(sh["point_ind"])
    xl = np.hstack((pb["cams0"].ravel(), sh["pts"].ravel()))
    fl = O.fun(xl, C, Pl, sh["camera_ind"], sh["point_ind"], sh["points_2d"], wl)
    np.testing.assert_allclose(fl, f.reshape(-1, 2)[sh["obs_sel"]].ravel(), atol=1e-12)
    _, Jcl, Jpl = O.jacobian_blocks(pb["cams0"], sh["pts"], sh["camera_ind"], sh["point_ind"], wl)
    Ul, gcl, Vl, gpl, Wl = O.normal_blocks(fl.reshape(-1, 2), Jcl, Jpl, C, Pl, sh["camera_ind"],
                                           sh["point_ind"])
    # camera sums + cost are all-reduced (what lcba_solve does with ncclAllReduce)
    cam = np.concatenate([Ul.ravel(), gcl.ravel(), [fl @ fl]])
    D.allreduce_sum(cam)
    Ug = cam[: C * 121].reshape(C, 11, 11)
    gcg = cam[C * 121: C * 132].reshape(C, 11)
    np.testing.assert_allclose(cam[-1], f @ f, rtol=1e-12)
    np.testing.assert_allclose(Ug, U, rtol=0, atol=1e-9 * np.abs(U).max())
    sc_c = np.sqrt(np.einsum("caa->ca", Ug)).ravel()
    sc_l = np.hstack((np.zeros(C * 11), np.sqrt(np.einsum("paa->pa", Vl)).ravel()))
    Sl, rl, _ = O.reduced_camera_system(np.zeros_like(Ul), np.zeros_like(gcl), Vl, gpl, Wl,
                                        sh["camera_ind"], sh["point_ind"], lam, sc_l)
    red = np.concatenate([Sl.ravel(), rl])
    D.allreduce_sum(red)
    n = C * 11
    S = red[: n * n].reshape(n, n)
    for c in range(C):
        S[c * 11:(c + 1) * 11, c * 11:(c + 1) * 11] += Ug[c] + lam * np.diag(sc_c[c * 11:(c + 1) * 11] ** 2)
    rhs = red[n * n:] + gcg.ravel()
    assert np.abs(S - S_ref).max() <= 1e-11 * np.abs(S_ref).max()
    assert np.abs(rhs - rhs_ref).max() <= 1e-10 * np.abs(rhs_ref).max()
    # ---- point gather ----
    full = D.allgather_rows(sh["pts"] + 1.0, b)
    np.testing.assert_array_equal(full, pb["pts0"] + 1.0)
    # every observation belongs to exactly one rank
    cnt = np.zeros(pi.size)
    cnt[sh["obs_sel"]] = 1
    D.allreduce_sum(cnt)
    assert np.all(cnt == 1)
    # balance: no rank holds more than 1.5x its fair share of observations (+ one point)
    assert sh["point_ind"].size <= 1.5 * pi.size / ws + C
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d ok" % rank)


if __name__ == "__main__":
    main()
