"""Development aid: repeat pipeline 3 of the packaged dtu006 example and print device / wall time per call (host-side jitter inside the timed region)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgegraph3d_b200 import lib as E, pipeline as P, real_scene
sc, _ = real_scene.dtu006_scene(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
prm = E.default_params(**P.REAL_DATA_CAPACITIES)
with E.DeviceScene(sc, prm) as dev:
    for i in range(8):
        t = time.time(); pts, tm = dev.match_refpoints(0, sc.n_tracks); w = (time.time() - t) * 1e3
        print("call %d: device %.1f ms, library wall %.1f ms, python wall %.1f ms, k3a %.1f k3b %.1f, points %d" % (i, tm["total_ms"], tm["host_wall_ms"], w, tm["k3a_ms"], tm["k3b_ms"], pts.n_points), flush=True)
