import sys, time, os; sys.path.insert(0,'/root/repo')
import numpy as np
from edgegraph3d_b200 import synthetic as syn, lib as E
sc = syn.make_scene(n_views=200, width=1920, height=1080, focal=1600.0, n_curves=400, segs_per_curve=20, curve_len=0.2, seed=1234, extent=0.9, closed_frac=0.05)
seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=250)
dev = E.DeviceScene(sc)
for rep in range(3):
    dp, tm = dev.match_seeds(seeds, fetch=False); dp.free()
print(os.path.basename(E.LIB_PATH), 'k3_ms', round(tm['k3_ms'],1), 'k1', round(tm['k1_count_ms']+tm['k1_fill_ms'],1), 'pts', tm['n_points'], flush=True)
