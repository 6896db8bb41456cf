"""Per-kernel device time of one bundleAdjust on a synthetic rig (CUDA events around every
launch, lcba_options.profile=1).  Usage: python tools/profile_solve.py ring24 1000000 1.0 [iters]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lasercalib_b200._cabi import Engine  # noqa: E402
from lasercalib_b200.synth import make_rig  # noqa: E402

rig = sys.argv[1] if len(sys.argv) > 1 else "ring24"
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
pvis = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
t = time.time()
pb = make_rig(rig, npts, seed=0, variant="volume", p_vis=pvis)
k = np.bincount(pb["point_ind"], minlength=pb["n_points"])
print("rig %s: C=%d P=%d N=%d mean k=%.2f (gen %.1fs)" % (rig, pb["n_cams"], pb["n_points"], pb["n_obs"],
                                                         k.mean(), time.time() - t))
eng = Engine()
t = time.time()
eng.set_problem(pb["cams0"], pb["pts0"], pb["points_2d"], pb["camera_ind"], pb["point_ind"])
print("set_problem %.3fs" % (time.time() - t))
for prof in (False, True):
    eng.set_params(pb["cams0"], pb["pts0"])
    res, trace = eng.solve(ftol=0.0, xtol=0.0, gtol=0.0, max_iterations=iters, profile=prof)
    print("profile=%s: %d iterations, nfev %d, %.3f ms total, %.3f ms/iter, launches %d, cost %.4f -> %.4f"
          % (prof, res.iterations, res.nfev, res.solve_ms, res.solve_ms / max(1, res.iterations),
             res.gpu_launches, res.initial_cost, res.cost))
st = eng.profile()
tot = sum(v["total_ms"] for v in st.values())
flop_pairs = float(np.sum(726.0 * k * (k + 1) / 2 + 264.0 * k) + 300.0 * pb["n_obs"])
for name, v in sorted(st.items(), key=lambda kv: -kv[1]["total_ms"]):
    print("  %-16s %5d launches %10.3f ms  %5.1f%%  %.3f ms/launch" % (name, v["launches"], v["total_ms"],
                                                                    100 * v["total_ms"] / tot, v["total_ms"] / v["launches"]))
if "schur" in st:
    ms = st["schur"]["total_ms"] / st["schur"]["launches"]
    print("schur: %.3e flop/launch -> %.2f TFLOP/s" % (flop_pairs, flop_pairs / ms * 1e-9))
for what, name, bytes_per_obs in ((0, "residual", 40.0), (1, "jacobian_blocks", 264.0)):
    ms = eng.time_device(what, 5)
    print("%s: %.3f ms -> %.2f G obs/s, %.0f GB/s algorithmic" % (name, ms, pb["n_obs"] / ms * 1e-6,
                                                              (bytes_per_obs * pb["n_obs"] + 24.0 * pb["n_points"]) / ms * 1e-6))
print(json.dumps(trace[-1]))
