"""CPU-side checks of the C-ABI boundary: the library builds, loads, exports every symbol
include/lcba.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "lcba.h")


@pytest.fixture(scope="module")
def lib():
    from lasercalib_b200 import build, _cabi
    build.build()
    return _cabi.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lcba_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    from lasercalib_b200 import _cabi
    decl = declared_symbols()
    assert len(decl) >= 20
    assert sorted(_cabi.SYMBOLS) == decl
    out = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lcba_[a-z0-9_]+)", out))
    for s in decl:
        assert s in exported, s
        assert getattr(lib, s) is not None


def test_version_and_default_options(lib):
    from lasercalib_b200 import _cabi
    assert lib.lcba_version() == 100
    o = _cabi.Options()
    lib.lcba_default_options(ctypes.byref(o))
    assert (o.ftol, o.xtol, o.gtol, o.max_nfev) == (1e-8, 1e-8, 1e-8, 0)


def test_struct_layouts_match_header():
    from lasercalib_b200 import _cabi
    assert ctypes.sizeof(_cabi.Options) == 3 * 8 + 8 + 8 * 4
    assert ctypes.sizeof(_cabi.TraceRow) == 64
    assert ctypes.sizeof(_cabi.Result) == 3 * 8 + 3 * 8 + 8 + 8 + 7 * 8
    assert ctypes.sizeof(_cabi.KernelStat) == 48


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lasercalib_b200 import _cabi
    with pytest.raises(_cabi.LcbaError, match="no CUDA device"):
        _cabi.Engine()
    from lasercalib_b200.pySBA import PySBA
    sba = PySBA(np.zeros((2, 11)), np.zeros((3, 3)), np.zeros((4, 2)), np.zeros(4, dtype=int),
                np.zeros(4, dtype=int))
    with pytest.raises(_cabi.LcbaError):
        sba.project(np.zeros((1, 3)), np.zeros((1, 11)))
    with pytest.raises(_cabi.LcbaError):
        sba.bundleAdjust()


def test_product_does_not_import_oracle():
    pkg = os.path.join(REPO, "lasercalib_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_tr2d_solver_matches_scipy(lib):
    """The device control kernel's 2-D trust-region solver (host-callable hook) against
    scipy's solve_trust_region_2d (scipy/optimize/_lsq/common.py:171-219)."""
    from scipy.optimize._lsq.common import solve_trust_region_2d
    from lasercalib_b200._cabi import debug_tr2d
    rng = np.random.default_rng(0)
    for i in range(3000):
        A = rng.normal(size=(3, 2))
        if i % 5 == 0:
            A[:, 1] = A[:, 0] * rng.normal() + 1e-9 * rng.normal(size=3)   # nearly rank-1
        B = A.T @ A * 10 ** rng.uniform(-3, 3)
        g = rng.normal(size=2) * 10 ** rng.uniform(-3, 3)
        D = 10 ** rng.uniform(-3, 3)
        p, newton = debug_tr2d(B[0, 0], B[0, 1], B[1, 1], g[0], g[1], D)
        ps, newton_s = solve_trust_region_2d(B, g, D)
        val = lambda q: 0.5 * q @ B @ q + g @ q
        assert np.linalg.norm(p) <= D * (1 + 1e-12)
        # same model value (the quartic route loses digits when B is nearly singular)
        assert val(p) <= val(ps) + 1e-6 * abs(val(ps)) + 1e-300
        if newton == newton_s and np.linalg.cond(B) < 1e8:
            np.testing.assert_allclose(p, ps, rtol=1e-7, atol=1e-9 * D)


def test_tensor_path_unit_plan_covers_every_tile_once():
    """Host logic of the tensor-path Schur kernel (schur_mma.cuh make_mma_plan, no GPU): for every
    camera count the units tile the lower triangle of the (11 C + 1)-row matrix exactly once,
    every kind fits 8 consumer warps, and the busiest SM sub-partition (warp w -> w % 4) stays
    within 40 % of the mean load (10 % at the 24 cameras of the benchmark)."""
    import ctypes as C
    from lasercalib_b200 import _cabi
    lib = _cabi.load()
    for cams in list(range(8, 33)) + [40, 48, 56, 64]:
        buf = np.zeros((16, 8, 5), dtype=np.int32)
        ns, ok = C.c_int32(), C.c_int32()
        nk = lib.lcba_debug_mma_plan(cams, 148, buf.ctypes.data_as(C.c_void_p), 16, C.byref(ns), C.byref(ok))
        assert nk >= 1 and ok.value == 1 and ns.value == 148 // nk
        nt = (11 * cams + 1 + 7) // 8
        cover = np.zeros((nt, nt), dtype=int)
        worst = 0.0
        for k in range(nk):
            load = np.zeros(4)
            for w in range(8):
                tr0, tc0, nr, nc, tri = buf[k, w]
                for t in range(nr):
                    for u in range(nc):
                        if tri and u > t:
                            continue
                        cover[tr0 + t, tc0 + u] += 1
                        load[w % 4] += 1
            assert load.sum() > 0
            worst = max(worst, load.max() / load.mean())
        low = np.tril(np.ones((nt, nt), dtype=int))
        np.testing.assert_array_equal(cover, low)
        assert worst <= (1.10 if cams == 24 else 1.40), (cams, worst)


def test_int8_schur_plan_covers_the_lower_triangle_once():
    """Host logic of the int8 tensor-core Schur kernel (schur_i8.cuh make_i8_plan, no GPU): for every
    camera count the tiles cover each lower-triangle entry of the (11 C + 1)-row matrix exactly once
    (what k_i8_gather keeps of a tile: block 2 is transposed, entries above the diagonal are dropped),
    block widths are multiples of 16 columns, both blocks together at most 64 (7 anti-diagonals x 64 TMEM
    columns), row tiles at most 128 rows, every tile's K ranges partition [0, nkb), the grid is exactly one
    wave (one CTA per SM) with more ranges for the tiles that cost more per K block (a second column block),
    and bench.py's executed-op count agrees."""
    import ctypes as C
    import importlib.util
    import os
    from lasercalib_b200 import _cabi
    lib = _cabi.load()
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for cams in list(range(8, 33)) + [40, 48, 56, 64]:
        for P in (1000, 125_000, 1_000_000):
            tiles = np.zeros((256, 8), dtype=np.int32)
            work = np.zeros((512, 3), dtype=np.int32)
            nwork, nrg, nkb = C.c_int32(), C.c_int32(), C.c_int64()
            nt = lib.lcba_debug_i8_plan(cams, P, 148, tiles.ctypes.data_as(C.c_void_p), 256, work.ctypes.data_as(C.c_void_p), 512,
                                        C.byref(nwork), C.byref(nrg), C.byref(nkb))
            assert 1 <= nt <= 256 and 1 <= nwork.value <= 148, (cams, nt, nwork.value)
            R = 11 * cams + 1
            assert nrg.value * 8 >= R and nrg.value % 2 == 0 and nkb.value == -(-P // 21)
            cover = np.zeros((R, R), dtype=int)
            cols = 0
            nws = []
            for m0, mn, n0, nn, n20, n2n, w0, nw in tiles[:nt]:
                assert 1 <= mn <= 16 and nn in (2, 4, 6, 8) and n2n in (0, 2, 4) and nn + n2n <= 8 and nw >= 1
                cols += 8 * (nn + n2n)
                nws.append(nw)
                rows = np.arange(8 * m0, 8 * (m0 + mn))
                for c_lo, c_n, transposed in ((n0, nn, False), (n20, n2n, True)):
                    if c_n == 0:
                        continue
                    cs = np.arange(8 * c_lo, 8 * (c_lo + c_n))
                    rr, cc = np.meshgrid(rows, cs, indexing="ij")
                    if transposed:
                        rr, cc = cc, rr
                    ok = (rr < R) & (cc < R) & (cc <= rr)
                    np.add.at(cover, (rr[ok], cc[ok]), 1)
                ranges = work[w0:w0 + nw]
                assert ranges[0, 1] == 0 and ranges[-1, 2] == nkb.value and np.all(ranges[1:, 1] == ranges[:-1, 2])
                assert np.all(ranges[:, 2] >= ranges[:, 1])
            np.testing.assert_array_equal(cover, np.tril(np.ones((R, R), dtype=int)))
            assert sum(nws) == nwork.value == min(148, nt * nkb.value), (cams, nws)
            two = [nw for (t, nw) in zip(tiles[:nt], nws) if t[5] > 0 and t[3] >= 4]
            one = [nw for (t, nw) in zip(tiles[:nt], nws) if t[5] == 0 and t[3] <= 4]
            if two and one and nkb.value > 148:
                assert min(two) >= max(one), (cams, nws)     # the two-block tiles get at least as many CTAs
            alg, exe = bench.i8_ops(cams, P)
            assert exe == 26 * 2.0 * 128 * cols * nkb.value * 64 and alg <= exe


def test_host_order_check_of_the_index_array():
    """lcba_host_is_nondecreasing_i64 (host-only, no GPU): the sharding layer's check that the caller's
    point indices are point-major (scripts/get_points3d.py:74-86 order) -- every thread count, violations
    inside a thread's piece and exactly on the piece borders, tiny and empty inputs; and dist._nondecreasing
    (which routes contiguous int64 through it and everything else through numpy) agrees on sub-ranges."""
    from lasercalib_b200 import _cabi, dist as D
    lib = _cabi.load()
    rng = np.random.default_rng(0)
    n = 5_000_000
    a = np.sort(rng.integers(0, 200_000, n))
    ptr = lambda v: v.ctypes.data
    for thr in (1, 2, 3, 4, 8, 64):
        assert lib.lcba_host_is_nondecreasing_i64(ptr(a), n, thr) == 1
        for pos in (1, n - 1, n // 2, n // 3, n // 4, n // 4 + 1, n // 4 - 1, 3 * n // 4, 2 * (n // 3)):
            b = a.copy()
            b[pos - 1] = b[pos] + 1
            assert lib.lcba_host_is_nondecreasing_i64(ptr(b), n, thr) == 0, (thr, pos)
            assert lib.lcba_host_is_nondecreasing_i64(ptr(b) + 8 * pos, n - pos, thr) == 1
    assert lib.lcba_host_is_nondecreasing_i64(ptr(a), 0, 4) == 1 and lib.lcba_host_is_nondecreasing_i64(ptr(a), 1, 4) == 1
    assert lib.lcba_host_is_nondecreasing_i64(None, 10, 4) == 1
    b = a.copy()
    b[1000] = -1
    assert not D._nondecreasing(b, 0, n) and D._nondecreasing(b, 1000, n) and not D._nondecreasing(b, 999, 1002)
    assert D._nondecreasing(a[::2], 0, n // 2) and D._nondecreasing(a.astype(np.int32), 0, n)      # numpy route
    assert not D._nondecreasing(b[::2], 0, n // 2) and not D._nondecreasing(b.astype(np.int32), 0, n)
