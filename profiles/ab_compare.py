"""Development aid (not a bench): run the C2 batch through one build of the library and print kernel times plus a
checksum of the result, so that A/B builds (EG3D_LIB=<path>) can be compared for speed AND identical output.
  EG3D_LIB=edgegraph3d_b200/libeg3d_x.so python profiles/ab_compare.py [reps] [seeds_limit]"""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from edgegraph3d_b200 import synthetic as syn, lib as E

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
limit = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = syn.make_scene(n_views=200, width=1920, height=1080, focal=1600.0, n_curves=400, segs_per_curve=20, curve_len=0.2, seed=1234,
                    extent=0.9, closed_frac=0.05)
seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=250)
if limit:
    seeds = seeds.take(np.linspace(0, len(seeds) - 1, limit).astype(np.int64))
dev = E.DeviceScene(sc)
best = None
for rep in range(reps):
    dp, tm = dev.match_seeds(seeds, fetch=False)
    if rep == reps - 1:
        ps = dp.fetch()
    dp.free()
    if best is None or tm["total_ms"] < best["total_ms"]:
        best = tm
crc = 0
for a in (ps.seed, ps.chain_pos, ps.obs_off, ps.obs_view, ps.obs_poly, ps.obs_seg, ps.obs_xy):
    crc = zlib.crc32(np.ascontiguousarray(a).tobytes(), crc)
print(os.path.basename(E.LIB_PATH), {k: os.environ[k] for k in os.environ if k.startswith("EG3D_") and k != "EG3D_LIB"},
      "total %.1f k1 any %.1f cnt %.1f fill %.1f scan %.2f k3a %.1f k3b %.1f pack %.2f hits %d" % (
          best["total_ms"], best["k1_any_ms"], best["k1_count_ms"], best["k1_fill_ms"], best["scan_ms"], best["k3a_ms"], best["k3b_ms"], best["pack_ms"], best["n_hits"]),
      "pts", ps.n_points, "obs", ps.n_obs, "crc %08x" % crc, "xyzsum %.6f" % float(np.abs(ps.xyz.astype(np.float64)).sum()), flush=True)
