import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from edgegraph3d_b200 import synthetic as syn, lib as E
t=time.time()
sc = syn.make_scene(n_views=200, width=1920, height=1080, focal=1600.0, n_curves=400, segs_per_curve=20, curve_len=0.2, seed=1234, extent=0.9, closed_frac=0.05)
seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=250)
print('gen', time.time()-t, len(seeds), flush=True)
t=time.time(); dev = E.DeviceScene(sc); print('scene create', time.time()-t, flush=True)
rng = np.random.default_rng(0)
for n in (512, 4096, 50000):
    sub = seeds if n == len(seeds) else seeds.take(np.sort(rng.choice(len(seeds), n, replace=False)))
    for rep in range(2):
        t=time.time(); dp, tm = dev.match_seeds(sub, fetch=False); wall=time.time()-t
        t=time.time(); nb = dp.fetch_raw(); d2h=time.time()-t
        print(n, 'wall', round(wall,3), 'd2h', round(d2h,3), 'MB', nb/1e6, {k:(round(v,2) if isinstance(v,float) else v) for k,v in tm.items()}, flush=True)
        dp.free()
