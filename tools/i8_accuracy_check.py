"""Accuracy and kernel time of the int8 tensor-core Schur path against the FP64 DMMA path on the resident
ring24 rig (S of lcba_linearize, final cost of bundleAdjust(1e-4)).  Usage: python tools/i8_accuracy_check.py [points ...]"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from lasercalib_b200._cabi import Engine  # noqa: E402
from lasercalib_b200.synth import make_rig  # noqa: E402

for npts in [int(a) for a in sys.argv[1:]] or [125_000, 262_144, 1_000_000]:
    pb = make_rig("ring24", npts, seed=0, variant="volume", p_vis=1.0)
    out = {}
    for name, env in (("i8_26", {"LCBA_SCHUR_I8": "1"}), ("dmma", {"LCBA_SCHUR_I8": "0"})):
        os.environ.update(env)
        eng = Engine()
        eng.set_problem(pb["cams0"], pb["pts0"], pb["points_2d"], pb["camera_ind"], pb["point_ind"])
        out[name] = eng.linearize(1e-4)["S"]
        eng.set_params(pb["cams0"], pb["pts0"])
        r, _ = eng.solve(ftol=1e-4, profile=True)
        pr = eng.profile()
        out[name + "_ms"] = pr["schur"]["total_ms"] / pr["schur"]["launches"]
        out[name + "_cost"] = r.cost
        eng.close()
        for k in env:
            del os.environ[k]
    ref = out["i8_26"]
    d = np.sqrt(np.abs(np.diag(ref)))
    for name in ("dmma",):
        e = np.abs(out[name] - ref)
        print("P=%d %s vs i8_26: max|dS|/max|S| %.2e, Jacobi-scaled %.2e; kernel %.3f ms (i8_26 %.3f ms); final cost rel diff %.1e"
              % (pb["n_points"], name, e.max() / np.abs(ref).max(), (e / np.outer(d, d)).max(), out[name + "_ms"], out["i8_26_ms"],
                 abs(out[name + "_cost"] - out["i8_26_cost"]) / out["i8_26_cost"]))
