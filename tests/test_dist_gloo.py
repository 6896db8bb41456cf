"""N > 1 host logic on CPU: world_size-2 (and 3) gloo jobs (no GPU)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from lasercalib_b200 import dist as D

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_world(ws, extra_env=None):
    port = _free_port()
    procs = []
    for r in range(ws):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(ws), LOCAL_RANK=str(r),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1",
                   CUDA_VISIBLE_DEVICES="")
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, os.path.join(REPO, "tests", "_dist_worker.py")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, out[-3000:])
        assert "rank %d ok" % r in out


def test_two_rank_gloo_shard_reduce_gather():
    _run_world(2)


def test_three_rank_gloo_unsorted_observations():
    _run_world(3, {"LCBA_TEST_SHUFFLE": "1"})


def test_shard_bounds_properties():
    rng = np.random.default_rng(0)
    P = 1000
    counts = rng.integers(0, 25, P)
    counts[100:140] = 0                        # points nobody sees
    pi = np.repeat(np.arange(P), counts)
    for ws in (1, 2, 4, 8):
        b = D.shard_bounds(pi, P, ws)
        assert b[0] == 0 and b[-1] == P and len(b) == ws + 1 and np.all(np.diff(b) >= 0)
        per = [int(((pi >= b[r]) & (pi < b[r + 1])).sum()) for r in range(ws)]
        assert sum(per) == pi.size
        assert max(per) <= pi.size / ws + 25
    # degenerate: fewer points than ranks
    b = D.shard_bounds(np.array([0, 0, 1]), 2, 8)
    assert b[0] == 0 and b[-1] == 2 and np.all(np.diff(b) >= 0)


def test_sortedness_is_checked_on_every_call():
    """The point-major fast path must not trust a cached verdict for a reused buffer address:
    sort an array in place -> unsorted content at the same address gets the general path."""
    rng = np.random.default_rng(0)
    P, ws = 50, 2
    pi = np.repeat(np.arange(P), 3)
    ci = np.tile(np.arange(3), P)
    uv = rng.normal(size=(pi.size, 2))
    pts = rng.normal(size=(P, 3))
    a = D.shard_problem(pts, uv, ci, pi, None, 0, ws)
    assert isinstance(a["obs_sel"], slice)
    perm = rng.permutation(pi.size)
    pi[:] = pi[perm]                                   # same buffer, now unsorted
    seen = np.zeros(pi.size, dtype=int)
    for r in range(ws):
        sh = D.shard_problem(pts, uv[perm], ci[perm], pi, None, r, ws)
        assert not isinstance(sh["obs_sel"], slice)
        np.testing.assert_array_equal(pi[sh["obs_sel"]], sh["point_ind"] + sh["lo"])
        seen[sh["obs_sel"]] += 1
    assert np.all(seen == 1)


def test_empty_shard_is_refused_on_all_ranks():
    pi = np.array([0, 0, 1, 1])
    for r in range(8):
        with pytest.raises(ValueError, match="without work"):
            D.shard_problem(np.zeros((2, 3)), np.zeros((4, 2)), np.zeros(4, dtype=int), pi, None, r, 8)


def test_world_defaults_to_single_process():
    assert D.world() == (0, 1, 0)


def test_shard_problem_properties_hypothesis():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 60), st.integers(1, 9), st.integers(0, 2**31 - 1), st.booleans())
    def check(P, ws, seed, shuffle):
        rng = np.random.default_rng(seed)
        counts = rng.integers(0, 6, P)
        pi = np.repeat(np.arange(P), counts)
        if pi.size == 0:
            pi = np.array([0])
        ci = rng.integers(0, 5, pi.size)
        uv = rng.normal(size=(pi.size, 2))
        if shuffle:
            perm = rng.permutation(pi.size)
            pi, ci, uv = pi[perm], ci[perm], uv[perm]
        pts = rng.normal(size=(P, 3))
        seen = np.zeros(pi.size, dtype=int)
        covered = np.zeros(P, dtype=int)
        # a sharding that leaves some rank without work is refused on EVERY rank alike (each rank
        # derives the same bounds), so no rank can fail alone while the others wait in NCCL
        verdicts = []
        for r in range(ws):
            try:
                D.shard_problem(pts, uv, ci, pi, None, r, ws)
                verdicts.append(True)
            except ValueError:
                verdicts.append(False)
        assert len(set(verdicts)) == 1
        if not verdicts[0]:
            return
        for r in range(ws):
            sh = D.shard_problem(pts, uv, ci, pi, None, r, ws)
            assert sh["point_ind"].size > 0 and sh["pts"].shape[0] > 0
            local = sh["point_ind"] - sh["pt_offset"]
            assert sh["pts"].shape[0] == sh["hi"] - sh["lo"]
            assert local.size == 0 or (local.min() >= 0 and local.max() < sh["pts"].shape[0])
            np.testing.assert_array_equal(pi[sh["obs_sel"]], local + sh["lo"])
            np.testing.assert_array_equal(uv[sh["obs_sel"]], sh["points_2d"])
            seen[sh["obs_sel"]] += 1
            covered[sh["lo"]:sh["hi"]] += 1
        assert np.all(seen == 1) and np.all(covered == 1)

    check()
