"""Development aid: where the end-to-end wall time of the packaged dtu006 example goes after the three matching calls
(profiles/c1_real_dtu006.py reports the total)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from edgegraph3d_b200 import lib as E, pipeline as P, real_scene
from edgegraph3d_b200.scene import PointSet
sc, _ = real_scene.dtu006_scene(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
cands1, cands2, _ = P.candidate_sets(sc)
prm = E.default_params(**P.REAL_DATA_CAPACITIES)
with E.DeviceScene(sc, prm) as dev:
    for rep in range(3):
        t = [time.perf_counter()]
        parts, tms = P.run_pipelines(dev, sc, cands1, cands2); t.append(time.perf_counter())
        allp = PointSet.concat(parts); t.append(time.perf_counter())
        keep = dev.dedup_close_points(allp); t.append(time.perf_counter())
        xyz, obs_off, obs_view, obs_xy = P.add_points_to_tracks(sc, allp, keep); t.append(time.perf_counter())
        fx, inl, tmf = dev.filter(xyz, obs_off, obs_view, obs_xy, sc.n_tracks); t.append(time.perf_counter())
        d = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
        print("rep %d: pipelines %.1f ms (device %.1f) | concat %.1f | density limiter %.1f | add_points_to_tracks %.1f | filter %.1f (gn %.2f) | total %.1f ms" % (
            rep, d[0], sum(x["total_ms"] for x in tms), d[1], d[2], d[3], d[4], tmf["gn_ms"] if tmf else -1, sum(d)), flush=True)
