import sys; sys.path.insert(0, ".")
from lasercalib_b200._cabi import Engine
from lasercalib_b200.synth import make_rig
pb = make_rig("ring24", 1000000, seed=0, variant="volume", p_vis=1.0)
eng = Engine()
eng.set_problem(pb["cams0"], pb["pts0"], pb["points_2d"], pb["camera_ind"], pb["point_ind"])
print(eng.time_device(1, 5), eng.time_device(0, 5))
