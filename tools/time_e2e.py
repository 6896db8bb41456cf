"""Wall-clock breakdown of the end-to-end path (host numpy arrays -> result) on one GPU."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from lasercalib_b200._cabi import Engine
from lasercalib_b200.pySBA import PySBA
from lasercalib_b200.synth import make_rig

pb = make_rig("ring24", int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, seed=0, variant="volume", p_vis=1.0)
def pin(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory(); return t.numpy(), t
keep = []
arrs = {}
for k in ("cams0", "pts0", "points_2d", "camera_ind", "point_ind"):
    arrs[k], t = pin(pb[k]); keep.append(t)
for label, src in (("pageable", pb), ("pinned", arrs)):
    for rep in range(3):
        eng = Engine()
        t0 = time.perf_counter()
        eng.set_problem(src["cams0"], src["pts0"], src["points_2d"], src["camera_ind"], src["point_ind"])
        t1 = time.perf_counter()
        res, _ = eng.solve(ftol=1e-4)
        t2 = time.perf_counter()
        c, p = eng.get_params()
        t3 = time.perf_counter()
        eng.close()
        t4 = time.perf_counter()
        print("%s rep%d: set_problem %.1f ms, solve %.1f ms (device %.1f, %d iters), get_params %.1f ms, close %.1f ms"
              % (label, rep, 1e3*(t1-t0), 1e3*(t2-t1), res.solve_ms, res.iterations, 1e3*(t3-t2), 1e3*(t4-t3)))
for rep in range(3):
    t0 = time.perf_counter()
    sba = PySBA(arrs["cams0"], arrs["pts0"], arrs["points_2d"], arrs["camera_ind"], arrs["point_ind"])
    r = sba.bundleAdjust(1e-4, verbose=0)
    t1 = time.perf_counter()
    print("PySBA.bundleAdjust pinned: %.1f ms total, %d iterations, device %.1f ms" % (1e3*(t1-t0), r["nit"], r["solve_ms"]))
