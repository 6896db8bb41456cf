import ctypes as C, os, sys
os.environ["LCBA_SCHUR_STATS"] = "1"
sys.path.insert(0, ".")
import numpy as np
from lasercalib_b200._cabi import Engine
from lasercalib_b200.synth import make_rig
pb = make_rig("ring24", int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, seed=0, variant="volume", p_vis=1.0)
eng = Engine()
eng.set_problem(pb["cams0"], pb["pts0"], pb["points_2d"], pb["camera_ind"], pb["point_ind"])
eng.linearize(1e-6); eng.linearize(1e-6)
buf = np.zeros((1024, 4), dtype=np.int64)
nk, ns = C.c_int(), C.c_int()
eng.lib.lcba_debug_schur_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
rc = eng.lib.lcba_debug_schur_stats(eng.h, buf.ctypes.data_as(C.c_void_p), 1024, C.byref(nk), C.byref(ns))
st = buf[: nk.value * ns.value].reshape(nk.value, ns.value, 4).astype(float)
for k in range(nk.value):
    cw, ct, pw, pt = st[k].mean(axis=0)
    print("kind %d: consumer total %.2f Mcyc (max %.2f), waiting at FULL %.1f%% | producer total %.2f Mcyc, waiting at EMPTY %.1f%%"
          % (k, ct / 1e6, st[k][:, 1].max() / 1e6, 100 * cw / ct, pt / 1e6, 100 * pw / pt))
