import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
from edgegraph3d_b200 import lib as E, synthetic as syn
import ctypes as C
from edgegraph3d_b200 import _abi as A
cfg = dict(n_views=25, width=1600, height=1200, focal=2900.0, n_curves=600, segs_per_curve=20, curve_len=0.12, seed=1234, extent=0.55, closed_frac=0.05, n_tracks=6268, track_cap=21, per_ring=25)
sc = syn.make_scene(**cfg)
cands = syn.curve_candidate_sets(sc, seed=1234)
dev = E.DeviceScene(sc)
L = E.load()
for rep in range(3):
    t0 = time.perf_counter()
    h = C.c_void_p(); tm = A.Timing()
    L.eg3d_match_refpoints(dev.h, 0, sc.n_tracks, C.byref(h), C.byref(tm))
    t1 = time.perf_counter()
    v = A.PointsView(); L.eg3d_points_get(h, C.byref(v))
    t2 = time.perf_counter()
    L.eg3d_points_free(h)
    cd = cands.desc()
    t3 = time.perf_counter()
    h2 = C.c_void_p(); tm2 = A.Timing()
    L.eg3d_match_polyline_sets(dev.h, C.byref(cd), 0, sc.n_views, C.byref(h2), C.byref(tm2))
    t4 = time.perf_counter()
    v2 = A.PointsView(); L.eg3d_points_get(h2, C.byref(v2))
    t5 = time.perf_counter()
    L.eg3d_points_free(h2)
    print("p3: call %.1f ms (device %.1f, seeds %d) get %.1f ms (%d pts, %d obs) | p12: call %.1f ms (device %.1f, seeds %d) get %.1f ms (%d pts)" % (
        1e3*(t1-t0), tm.total_ms, tm.n_seeds, 1e3*(t2-t1), v.n_points, v.n_obs, 1e3*(t4-t3), tm2.total_ms, tm2.n_seeds, 1e3*(t5-t4), v2.n_points), flush=True)
