"""Development aid: run the dtu006-shaped scene (tests/test_dtu006_fixture.py) on the GPU and on the oracle, print the first
points whose observation lists differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from edgegraph3d_b200 import lib as E, synthetic as syn
from tests import oracle_lib as O
from tests.test_dtu006_fixture import dtu_scene
sc, _ = dtu_scene()
cands = syn.curve_candidate_sets(sc, seed=6)
osc = O.OracleScene(sc)
with E.DeviceScene(sc) as dev:
    g12, _ = dev.match_polyline_sets(cands, 0, 6)
    g3, _ = dev.match_refpoints(0, 1200)
r12 = osc.match_polyline_sets(cands, 0, 6, n_threads=16)
r3 = osc.match_refpoints(0, 1200, n_threads=16)
for name, g, r in (("p12", g12, r12), ("p3", g3, r3)):
    print(name, g.n_points, r.n_points, g.n_obs, r.n_obs)
    gl, rl = np.diff(g.obs_off), np.diff(r.obs_off)
    n = min(len(gl), len(rl))
    bad = np.where(gl[:n] != rl[:n])[0]
    print(" points with different obs counts:", bad[:10], len(bad))
    for i in bad[:3]:
        print("  point", i, "seed", g.seed[i], r.seed[i], "pos", g.chain_pos[i], r.chain_pos[i])
        a, b = g.obs_off[i], g.obs_off[i + 1]; c, d = r.obs_off[i], r.obs_off[i + 1]
        print("  gpu views", g.obs_view[a:b].tolist()); print("  ora views", r.obs_view[c:d].tolist())
        print("  gpu xyz", g.xyz[i], "ora xyz", r.xyz[i])
        s = g.seed[i]
        sel = np.where(g.seed == s)[0]
        print("  chain of that seed: gpu lens", gl[sel].tolist(), " ora lens", rl[np.where(r.seed == r.seed[i])[0]].tolist())
    np.savez("gpurun_out/debug_diff_%s.npz" % name, g_seed=g.seed, g_pos=g.chain_pos, g_off=g.obs_off, g_view=g.obs_view, g_poly=g.obs_poly, g_seg=g.obs_seg, g_xy=g.obs_xy, g_xyz=g.xyz)
