import sys, numpy as np
sys.path.insert(0, ".")
from lasercalib_b200._cabi import Engine
from lasercalib_b200.synth import make_rig
pb = make_rig("ring24", 200000, seed=0, variant="volume", p_vis=1.0)
eng = Engine()
eng.set_problem(pb["cams0"], pb["pts0"], pb["points_2d"], pb["camera_ind"], pb["point_ind"])
for i in range(3):
    eng.set_params(pb["cams0"], pb["pts0"])
    r, _ = eng.solve(ftol=1e-4)
    print("solve", i, "iters", r.iterations, "ms/iter %.3f" % (r.solve_ms / r.iterations), "graph replays", r.reserved[0], "launches", r.gpu_launches)
