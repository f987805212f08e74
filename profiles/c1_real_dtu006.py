"""Documentation run (not a bench line): the reference's packaged example end to end on real data — the 25 real edge
maps (tests/golden/dtu006_edges.npz) -> polyline graphs (row f1, host) -> pipeline 3 (refpoint-seeded matching) on the
device, checked against the CPU oracle.  Usage: python profiles/c1_real_dtu006.py [out.json]"""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edgegraph3d_b200 import lib as E, real_scene  # noqa: E402
from tests import oracle_lib as O  # noqa: E402

if __name__ == "__main__":
    t = time.time(); sc, plgs = real_scene.dtu006_scene(os.path.join(ROOT, "tests", "golden")); t_plg = time.time() - t
    res = {"views": sc.n_views, "tracks": sc.n_tracks, "segments_per_view": [sc.n_segments(v) for v in range(sc.n_views)], "plg_build_s": t_plg}
    prm = E.default_params(max_chain_points=256, max_follow_points=320)
    with E.DeviceScene(sc, prm) as dev:
        g3, tm = dev.match_refpoints(0, sc.n_tracks)
        g3b, tm = dev.match_refpoints(0, sc.n_tracks)
    res["gpu"] = {"points": g3.n_points, "obs": g3.n_obs, "timing": tm}
    t = time.time(); r3 = O.OracleScene(sc, prm).match_refpoints(0, sc.n_tracks, n_threads=os.cpu_count()); res["oracle_s"] = time.time() - t
    res["oracle"] = {"points": r3.n_points, "obs": r3.n_obs, "threads": os.cpu_count()}
    same = (g3.n_points == r3.n_points and np.array_equal(g3.obs_off, r3.obs_off) and np.array_equal(g3.obs_view, r3.obs_view)
            and np.array_equal(g3.obs_poly, r3.obs_poly) and np.array_equal(g3.obs_seg, r3.obs_seg) and g3.obs_xy.tobytes() == r3.obs_xy.tobytes())
    res["identical_chains_and_observations"] = bool(same)
    res["max_abs_xyz_diff"] = float(np.abs(g3.xyz - r3.xyz).max()) if same and g3.n_points else None
    print(json.dumps(res))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"))
