#!/usr/bin/env python
"""Documentation runs of the BASELINE.json configurations that are not the bench headline (one JSON line each):

  python profiles/run_configs.py c3 [--shards 8] [--shard 0]   # 1000-view rig, the seed block one of 8 GPUs owns
  python profiles/run_configs.py c4                            # dtu006-shaped geometry, seeds every 2 px (10x density)
  python profiles/run_configs.py c5 [--n 10000000]             # GN filter microbench, 10 M hypotheses x 20 observations

Each run checks the CUDA result against the CPU oracle (tests/oracle_lib.py: test infrastructure) on a bounded sample.
Not a bench: bench.py stays the one contract line.  Needs a GPU."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from edgegraph3d_b200 import lib as E, synthetic as syn  # noqa: E402
from tests import oracle_lib as O  # noqa: E402


def same_points(a, b):
    return bool(a.n_points == b.n_points and np.array_equal(a.obs_off, b.obs_off) and np.array_equal(a.obs_view, b.obs_view)
                and np.array_equal(a.obs_poly, b.obs_poly) and np.array_equal(a.obs_seg, b.obs_seg)
                and a.obs_xy.tobytes() == b.obs_xy.tobytes() and (a.n_points == 0 or float(np.abs(a.xyz - b.xyz).max()) < 1e-4))


def run_c3(args):
    V = 1000
    t = time.perf_counter()
    sc = syn.make_scene(n_views=V, width=1920, height=1080, focal=1600.0, n_curves=400, segs_per_curve=20, curve_len=0.2, seed=1234,
                        extent=0.9, closed_frac=0.05)
    gen_s = time.perf_counter() - t
    lo, hi = (args.shard * V) // args.shards, ((args.shard + 1) * V) // args.shards
    seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=250, views=range(lo, hi))
    t = time.perf_counter()
    dev = E.DeviceScene(sc)
    up_s = time.perf_counter() - t
    dev.match_seeds(seeds.take(np.arange(0, len(seeds), 50)), fetch=False)[0].free()     # warm-up
    dp, tm = dev.match_seeds(seeds, fetch=False)
    pts = dp.fetch(); dp.free()
    sel = np.linspace(0, len(seeds) - 1, args.oracle_seeds).astype(np.int64)
    t = time.perf_counter()
    ref = O.OracleScene(sc).match_seeds(seeds.take(sel), n_threads=os.cpu_count())
    cpu_s = time.perf_counter() - t
    got, _ = dev.match_seeds(seeds.take(sel))
    print(json.dumps({"config": "c3 (BASELINE configs[2]): synthetic 1000-view rig, 1920x1080, %d segments/view, 250 seeds/view; the seed block of "
                                "shard %d of %d (starting views %d..%d), all-segment sweep" % (sc.n_segments(0), args.shard, args.shards, lo, hi - 1),
                      "seeds": len(seeds), "points": pts.n_points, "observations": pts.n_obs, "accepted_seeds": tm["n_accepted_seeds"],
                      "device_ms": tm["total_ms"], "points_per_s": pts.n_points / (tm["total_ms"] * 1e-3),
                      "kernel_ms": {k: tm[k] for k in ("k1_any_ms", "k1_count_ms", "k1_fill_ms", "scan_ms", "k3a_ms", "k3b_ms", "pack_ms")},
                      "hits_materialised": tm["n_hits"], "segment_tests_full_sweep": tm["n_segment_tests"],
                      "scene_generate_s": gen_s, "scene_upload_s": up_s,
                      "oracle_sample": {"seeds": len(sel), "points": ref.n_points, "identical": same_points(got, ref), "cpu_s": cpu_s,
                                        "cpu_threads": os.cpu_count(), "cpu_points_per_s": ref.n_points / cpu_s}}))


def run_c4(args):
    cfg = dict(n_views=25, width=1600, height=1200, focal=2900.0, n_curves=600, segs_per_curve=20, curve_len=0.12, seed=1234, extent=0.55,
               closed_frac=0.05, n_tracks=6268, track_cap=21, per_ring=25)
    sc = syn.make_scene(**cfg)
    cands = syn.curve_candidate_sets(sc, seed=cfg["seed"])
    prm = E.default_params(split_interval_distance=2.0)          # SPLIT_INTERVAL_DISTANCE 20 -> 2
    dev = E.DeviceScene(sc, prm)
    dev.match_polyline_sets(cands)                              # warm-up: the stream-ordered memory pool grows to its working size once
    t = time.perf_counter()
    pts, tm = dev.match_polyline_sets(cands)
    wall = time.perf_counter() - t
    # oracle on the first starting views only (10x seeds make the full CPU run long)
    ve = args.oracle_views
    osc = O.OracleScene(sc, prm)
    t = time.perf_counter()
    ref = osc.match_polyline_sets(cands, 0, ve, n_threads=os.cpu_count())
    cpu_s = time.perf_counter() - t
    got, _ = dev.match_polyline_sets(cands, 0, ve)
    print(json.dumps({"config": "c4 (BASELINE configs[3]): dtu006-shaped geometry (25 views 1600x1200, %d segments/view), seeds every 2 px, "
                                "candidate-set mode, pipelines 1-2" % sc.n_segments(0),
                      "seeds": tm["n_seeds"], "points": pts.n_points, "observations": pts.n_obs, "device_ms": tm["total_ms"], "e2e_ms": 1e3 * wall,
                      "points_per_s": pts.n_points / (tm["total_ms"] * 1e-3),
                      "kernel_ms": {k: tm[k] for k in ("k1_count_ms", "k1_fill_ms", "scan_ms", "k3a_ms", "k3b_ms", "pack_ms")},
                      "oracle_sample": {"starting_views": ve, "points": ref.n_points, "identical": same_points(got, ref), "cpu_s": cpu_s,
                                        "cpu_threads": os.cpu_count(), "cpu_points_per_s": ref.n_points / cpu_s}}))


def run_c5(args):
    import torch
    sc = syn.make_scene(n_views=200, width=1920, height=1080, focal=1600.0, n_curves=8, segs_per_curve=20, curve_len=0.2, seed=1234,
                        extent=0.9, closed_frac=0.05)
    dev = E.DeviceScene(sc)
    base_n, k = 1_000_000, 20
    views, xy, init, _ = syn.gn_microbench_inputs(sc, base_n, k, seed=99)
    reps = max(1, args.n // base_n)
    n = base_n * reps
    # 10 M hypotheses = the 1 M generated ones repeated `reps` times on the device (host generation of 10 M x 200 view
    # permutations does not fit the box's memory); every repetition is solved independently
    dv = torch.from_numpy(views).cuda().repeat(reps, 1).contiguous(); dxy = torch.from_numpy(xy).cuda().repeat(reps, 1, 1).contiguous()
    di = torch.from_numpy(init).cuda().repeat(reps, 1).contiguous()
    ox = torch.empty((n, 3), dtype=torch.float32, device="cuda"); om = torch.empty(n, dtype=torch.float32, device="cuda")
    ok = torch.empty(n, dtype=torch.uint8, device="cuda")
    out = {}
    for fp64 in (0, 1):
        ms = []
        for _ in range(4):
            tm = dev.gn_triangulate_device(n, k, dv.data_ptr(), dxy.data_ptr(), di.data_ptr(), fp64, ox.data_ptr(), om.data_ptr(), ok.data_ptr())
            ms.append(tm["gn_ms"])
        torch.cuda.synchronize()
        sel = np.arange(0, base_n, 4001)
        off = np.arange(len(sel) + 1, dtype=np.int64) * k
        x, m, o = O.OracleScene(sc).gn_triangulate(off, views[sel].reshape(-1), xy[sel].reshape(-1, 2), init[sel], fp64, n_threads=os.cpu_count())
        gx, go = ox[:base_n].cpu().numpy()[sel], ok[:base_n].cpu().numpy()[sel]
        good = o == 1
        best = min(ms[1:])
        bytes_alg = n * (k * 12 + 12 + 17)
        out["fp64" if fp64 else "fp32"] = {"ms": best, "hypotheses_per_s": n / (best * 1e-3), "algorithmic_GBps": bytes_alg / (best * 1e-3) / 1e9,
                                           "inlier_frac": float(ok.float().mean()), "oracle_sample": len(sel),
                                           "flags_identical": bool(np.array_equal(go, o)),
                                           "x_max_abs_diff": float(np.abs(gx[good] - x[good]).max()) if good.any() else 0.0}
    print(json.dumps({"config": "c5 (BASELINE configs[4]): Gauss-Newton filter microbench, %d hypotheses x %d observations, device-resident "
                                "(fp32 = the outlier filter's GN, fp64 = the matching path's GN)" % (n, k), **out}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c4", "c5"])
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--shard", type=int, default=0)
    ap.add_argument("--oracle-seeds", type=int, default=160)
    ap.add_argument("--oracle-views", type=int, default=3)
    ap.add_argument("--n", type=int, default=10_000_000)
    args = ap.parse_args()
    {"c3": run_c3, "c4": run_c4, "c5": run_c5}[args.config](args)


if __name__ == "__main__":
    main()
