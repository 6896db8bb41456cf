"""Warm per-iteration time on laserCalib-sized problems (the real rig: 17 cameras, ~4000 points)."""
import sys, time
sys.path.insert(0, ".")
from lasercalib_b200._cabi import Engine
from lasercalib_b200.synth import make_rig
for rig, n in (("example18", 4000), ("ring4", 10000), ("ring24", 20000)):
    pb = make_rig(rig, n, seed=0, variant="volume", p_vis=0.8)
    eng = Engine()
    eng.set_problem(pb["cams0"], pb["pts0"], pb["points_2d"], pb["camera_ind"], pb["point_ind"])
    for rep in range(3):
        eng.set_params(pb["cams0"], pb["pts0"])
        t = time.perf_counter()
        res, _ = eng.solve(ftol=1e-4)
        wall = (time.perf_counter() - t) * 1e3
    print("%s %d pts, %d obs: %d iterations, nfev %d, device %.3f ms (%.3f ms/iter), wall %.3f ms, launches %d"
          % (rig, pb["n_points"], pb["n_obs"], res.iterations, res.nfev, res.solve_ms,
             res.solve_ms / res.iterations, wall, res.gpu_launches))
    eng.close()
