#!/usr/bin/env python
"""bench.py — triangulated 3D edge-points/sec on the BASELINE.json workload (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|small|c1|c1real]

A "step" is one pass of the hot path (K1 epipolar intersection -> K3 triple enumeration / PLG following / view expansion
-> ordered packing [-> NCCL all-gather of the accepted points when N > 1]) over the whole seed batch of the workload.
K1 runs lazily inside the path (any-hit pass, the three selected views of every seed, every view of the accepted seeds:
the lists nobody reads are not materialised); `roofline_k1` therefore times the FULL sweep of the same batch separately,
through eg3d_epipolar_intersect_device, after the timed steps.
N = 1 runs BASELINE configs[1]: the synthetic 200-view rig, 1920x1080, 8 000 polyline segments per view, 50 000 seeds.
N > 1 is weak scaling on the same rig: every rank keeps 50 000 seeds (seed spacing 20/N px => 250*N seeds per view) and
owns the starting views r, r+N, ... (the reference's outer loop over starting views, polyline_matching.cpp:162, dealt round-robin).

`value` = accepted (pre-dedup) 3D edge-points of all ranks / max-over-ranks device time of a step (CUDA events on the
library's stream; inputs already resident).  `e2e` = the same count / wall time of the C-ABI call sequence a user makes
with HOST buffers: eg3d_match_seeds (pinned seeds H2D + kernels) + eg3d_points_get (D2H of the result into pinned memory).
--impl reference times the CPU oracle (the reference cannot be compiled in this image: DESIGN.md) on a bounded seed sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from edgegraph3d_b200 import synthetic as syn  # noqa: E402

METRIC = "triangulated 3D edge-points/sec (device-timed)"
UNIT = "points/s"


def build_workload(name, n_gpus):
    if name == "c2":
        cfg = dict(n_views=200, width=1920, height=1080, focal=1600.0, n_curves=400, segs_per_curve=20, curve_len=0.2,
                   seed=1234, extent=0.9, closed_frac=0.05)
        per_view = 250
    elif name == "c1":
        # dtu006-shaped: 25 views, 1600x1200, ~12k segments/view, 6268 tracks; candidate-set mode
        cfg = dict(n_views=25, width=1600, height=1200, focal=2900.0, n_curves=600, segs_per_curve=20, curve_len=0.12,
                   seed=1234, extent=0.55, closed_frac=0.05, n_tracks=6268, track_cap=21, per_ring=25)
        per_view = 0
    elif name == "small":
        cfg = dict(n_views=24, width=1280, height=720, focal=1000.0, n_curves=120, segs_per_curve=20, curve_len=0.3,
                   seed=1234, extent=0.8, closed_frac=0.05)
        per_view = 120
    else:
        raise SystemExit(f"unknown workload {name}")
    scene = syn.make_scene(**cfg)
    return scene, cfg, per_view


def make_seeds(scene, sampler, per_view, n_gpus, rank):
    """250*N seeds per view at spacing 20/N px, first polylines first; rank r owns the starting views r, r+N, r+2N, ...
    (multigpu.view_round_robin: better balanced than contiguous blocks; the merged result is put back into the
    reference's loop order by global seed ordinal, multigpu.all_gather_points(order_keys=...))."""
    from edgegraph3d_b200 import multigpu as mg
    return syn.sample_seeds(sampler, scene, per_view=per_view * n_gpus, spacing=20.0 / n_gpus, views=mg.view_round_robin(scene.n_views, n_gpus, rank))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch from the committed ncu --set full capture (profiles/r01_traffic.json); None if absent."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(p):
        return {}
    return json.load(open(p))


def pinned_seeds(seeds):
    """Move the seed arrays into page-locked host memory (the e2e contract copies inputs from pinned memory)."""
    import torch
    from edgegraph3d_b200.scene import SeedBatch

    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t
    keep = []
    out = []
    for a in (seeds.view, seeds.polyline.view(np.int32), seeds.segment.view(np.int32), seeds.xy):
        n, t = pin(a)
        keep.append(t)
        out.append(n)
    sb = SeedBatch(out[0], out[1].view(np.uint32), out[2].view(np.uint32), out[3], None)
    sb._pinned = keep
    return sb


def cpu_baseline(scene, seeds, target_seconds=15.0, threads=None):
    """The CPU oracle (the port of the reference path) on a bounded stratified sample of the same seed batch."""
    from tests import oracle_lib as O
    threads = threads or os.cpu_count()
    osc = O.OracleScene(scene)
    n = len(seeds)
    probe = seeds.take(np.linspace(0, n - 1, min(n, 64)).astype(np.int64))
    t = time.perf_counter()
    osc.match_seeds(probe, n_threads=threads)
    per_seed = max((time.perf_counter() - t) / len(probe), 1e-6)
    m = int(min(n, max(64, target_seconds / per_seed)))
    sample = seeds.take(np.linspace(0, n - 1, m).astype(np.int64))
    t = time.perf_counter()
    pts = osc.match_seeds(sample, n_threads=threads)
    dt = time.perf_counter() - t
    return {"value": pts.n_points / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{m} of {n} seeds (stratified: every {n / m:.1f}th seed of the batch), full per-seed path (K1 sweep + "
                      f"triples + PLG following + view expansion), {pts.n_points} points in {dt:.2f} s",
            "seconds": dt, "seeds": m, "points": pts.n_points}, osc, sample


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (reference unbuildable here)."""
    if rank != 0:
        return
    from tests import oracle_lib as O
    scene, cfg, per_view = build_workload(args.workload, 1)
    seeds = make_seeds(scene, O.sample_seeds, per_view, 1, 0)
    threads = os.cpu_count()
    osc = O.OracleScene(scene)
    n = len(seeds)
    probe = seeds.take(np.linspace(0, n - 1, 64).astype(np.int64))
    t = time.perf_counter(); osc.match_seeds(probe, n_threads=threads); per_seed = (time.perf_counter() - t) / 64
    budget = 150.0 / max(1, args.steps + args.warmup)             # whole run within a few minutes
    m = int(min(n, max(64, budget / max(per_seed, 1e-6))))
    sample = seeds.take(np.linspace(0, n - 1, m).astype(np.int64))
    for _ in range(args.warmup):
        osc.match_seeds(sample, n_threads=threads)
    t0 = time.perf_counter()
    npts = 0
    for _ in range(args.steps):
        npts += osc.match_seeds(sample, n_threads=threads).n_points
    dt = time.perf_counter() - t0
    value = npts / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, cfg, per_view, 1), "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"each step = {m} of {n} seeds (stratified), OpenMP over seeds on {threads} threads; "
                                       "the reference itself is effectively single-threaded (orphaned omp for, SURVEY finding 4)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_c1(args, scene, cfg, E):
    """dtu006-shaped input, reference semantics (candidate sets): pipelines 1-2 + pipeline 3 + density limiter + filter,
    GPU vs the CPU oracle on the full input.  Prints one JSON line (documentation; not the headline configuration)."""
    import torch
    from tests import oracle_lib as O
    cands = syn.curve_candidate_sets(scene, seed=cfg["seed"])
    dev = E.DeviceScene(scene)
    out = {}
    for _ in range(max(1, args.warmup)):
        dev.match_polyline_sets(cands); dev.match_refpoints()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(args.steps):
        p12, tm12 = dev.match_polyline_sets(cands)
        p3, tm3 = dev.match_refpoints()
    wall = (time.perf_counter() - t) / args.steps
    from edgegraph3d_b200.scene import PointSet
    allp = PointSet.concat([p12, p3])
    t = time.perf_counter(); keep = dev.dedup_close_points(allp); t_dedup = time.perf_counter() - t
    kept = np.where(keep)[0]
    xyz = np.concatenate([scene.track_xyz, allp.xyz[kept]])
    lens = allp.obs_off[kept + 1] - allp.obs_off[kept]
    obs_off = np.concatenate([scene.track_off, int(scene.track_off[-1]) + np.cumsum(lens)])
    idx = np.concatenate([np.arange(allp.obs_off[i], allp.obs_off[i + 1]) for i in kept]) if len(kept) else np.zeros(0, np.int64)
    obs_view = np.concatenate([scene.track_view, allp.obs_view[idx]]); obs_xy = np.concatenate([scene.track_xy, allp.obs_xy[idx]])
    t = time.perf_counter(); fx, inl, tmf = dev.filter(xyz, obs_off, obs_view, obs_xy, scene.n_tracks); t_filter = time.perf_counter() - t
    dev_ms = tm12["total_ms"] + tm3["total_ms"]
    threads = os.cpu_count()
    osc = O.OracleScene(scene)
    t = time.perf_counter(); o12 = osc.match_polyline_sets(cands, n_threads=threads); o3 = osc.match_refpoints(n_threads=threads); cpu_s = time.perf_counter() - t
    t = time.perf_counter(); osc.match_polyline_sets(cands, n_threads=1); cpu1_s = time.perf_counter() - t
    same = (o12.n_points == p12.n_points and o3.n_points == p3.n_points and np.array_equal(o12.obs_off, p12.obs_off)
            and np.array_equal(o12.obs_seg, p12.obs_seg) and np.array_equal(o3.obs_seg, p3.obs_seg))
    n_pts = p12.n_points + p3.n_points
    print(json.dumps({"workload": "c1 dtu006-shaped: 25 views 1600x1200, %d segments/view, %d tracks, candidate-set mode, pipelines 1-3" % (scene.n_segments(0), scene.n_tracks),
                      "points": n_pts, "points_pipelines12": p12.n_points, "points_pipeline3": p3.n_points,
                      "device_ms": dev_ms, "value_points_per_s": n_pts / (dev_ms * 1e-3), "e2e_ms": 1e3 * wall, "e2e_points_per_s": n_pts / wall,
                      "kernel_ms_p12": {k: tm12[k] for k in ("k1_count_ms", "k1_fill_ms", "k3a_ms", "k3b_ms", "pack_ms")},
                      "kernel_ms_p3": {k: tm3[k] for k in ("k1_count_ms", "k1_fill_ms", "k3a_ms", "k3b_ms", "pack_ms")},
                      "dedup_s": t_dedup, "kept_after_dedup": int(keep.sum()), "filter_s": t_filter, "filter_gn_ms": tmf["gn_ms"], "inliers": int(inl.sum()),
                      "cpu_oracle": {"cores": threads, "seconds": cpu_s, "points_per_s": (o12.n_points + o3.n_points) / cpu_s,
                                     "pipelines12_1thread_s": cpu1_s, "pipelines12_1thread_points_per_s": o12.n_points / cpu1_s},
                      "identical_to_oracle": bool(same)}))


def workload_name(name, cfg, per_view, n):
    if name == "c2":
        return (f"BASELINE configs[1]: synthetic {cfg['n_views']}-view rig, {cfg['width']}x{cfg['height']}, "
                f"{cfg['n_curves'] * cfg['segs_per_curve']} polyline segments/view, {per_view * cfg['n_views']} seeds per GPU "
                f"({per_view * n}/view at {20.0 / n:g} px), all-segment sweep, seed {cfg['seed']}")
    return f"{name}: {cfg}"


def main():
    # Only the JSON line may reach stdout: library banners (NCCL prints its version there) are sent to stderr by
    # pointing fd 1 at fd 2 for the rest of the process and keeping the real stdout for the one line.
    global print
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    _print = print

    def print(*a, **k):  # noqa: A001
        k.setdefault("file", real_stdout)
        _print(*a, **k)
        real_stdout.flush()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seeds-limit", type=int, default=0, help="profiling aid: keep only a stratified subset of the seed batch (NOT a bench configuration)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from edgegraph3d_b200 import lib as E
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world

    if args.workload == "c1real":
        # BASELINE configs[0] on the packaged files: real edge maps -> polyline graphs -> candidate sets -> pipelines 2+3 ->
        # density limiter -> filter, every stage checked against the oracle (documentation line, not the headline configuration)
        sys.path.insert(0, os.path.join(ROOT, "profiles"))
        import c1_real_dtu006
        print(json.dumps(c1_real_dtu006.run(os.path.join(ROOT, "tests", "golden"), os.cpu_count())))
        return
    scene, cfg, per_view = build_workload(args.workload, n_gpus)
    if args.workload == "c1":
        return run_c1(args, scene, cfg, E)
    seeds = make_seeds(scene, E.sample_seeds, per_view, n_gpus, rank)
    if args.seeds_limit and args.seeds_limit < len(seeds):
        seeds = seeds.take(np.linspace(0, len(seeds) - 1, args.seeds_limit).astype(np.int64))
    seeds = pinned_seeds(seeds)
    t = time.perf_counter()
    dev = E.DeviceScene(scene)
    torch.cuda.synchronize()
    scene_ms = 1e3 * (time.perf_counter() - t)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allgather_points(dp):
        """The path's one exchange step: accepted records of every rank to every rank over NCCL (NVLink)."""
        v = dp.device_view()
        n, m = int(v.n_points), int(v.n_obs)
        cnt = torch.tensor([n, m], dtype=torch.int64, device="cuda")
        cnts = torch.empty(world * 2, dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(cnts, cnt)
        cnts = cnts.view(world, 2).cpu()
        maxn, maxm = int(cnts[:, 0].max()), int(cnts[:, 1].max())

        def wrap(ptr, nbytes):
            import ctypes
            if nbytes == 0:
                return torch.empty(0, dtype=torch.uint8, device="cuda")
            addr = ctypes.cast(ptr, ctypes.c_void_p).value

            class _W:
                __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (addr, False), "version": 2}
            return torch.as_tensor(_W(), device="cuda")
        total = 0
        for ptr, per, mx, k in ((v.xyz, 12, maxn, n), (v.seed, 4, maxn, n), (v.chain_pos, 4, maxn, n), (v.obs_off, 8, maxn + 1, n + 1),
                                (v.obs_view, 4, maxm, m), (v.obs_poly, 4, maxm, m), (v.obs_seg, 4, maxm, m), (v.obs_xy, 8, maxm, m)):
            send = torch.zeros(mx * per, dtype=torch.uint8, device="cuda")
            send[:k * per] = wrap(ptr, k * per)
            recv = torch.empty(world * mx * per, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(recv, send)
            total += recv.numel()
        return int(cnts[:, 0].sum()), total

    def step(timed):
        flush.fill_(1)                                   # evict L2 between iterations
        barrier()
        t0 = time.perf_counter()
        dp, tm = dev.match_seeds(seeds, fetch=False)     # pinned seeds H2D + K1 + K3 + pack (device-resident result)
        ag_ms, total_pts = 0.0, tm["n_points"]
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            total_pts, _ = allgather_points(dp)
            e1.record(); torch.cuda.synchronize()
            ag_ms = e0.elapsed_time(e1)
        d2h = dp.fetch_raw()                             # D2H of this rank's result into pinned host memory
        t1 = time.perf_counter()
        dp.free()
        return tm, ag_ms, d2h, t1 - t0, total_pts

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, e2e_s, launches = [], [], 0
    k1_ms, k3_ms, pack_ms, ag_all, k3a_ms, k3b_ms, k1any_ms, scan_ms = [], [], [], [], [], [], [], []
    last_tm = None
    pts_total = 0
    for _ in range(args.steps):
        tm, ag_ms, d2h, wall, total_pts = step(True)
        dev_ms.append(tm["total_ms"] + ag_ms); e2e_s.append(wall); launches += tm["kernel_launches"] + (9 if world > 1 else 0)
        k1_ms.append((tm["k1_count_ms"], tm["k1_fill_ms"])); k3_ms.append(tm["k3_ms"]); pack_ms.append(tm["pack_ms"]); ag_all.append(ag_ms); k3a_ms.append(tm["k3a_ms"]); k3b_ms.append(tm["k3b_ms"])
        k1any_ms.append(tm["k1_any_ms"]); scan_ms.append(tm["scan_ms"])
        last_tm = tm; pts_total = total_pts; d2h_bytes = d2h
    clocks = sampler.stop() if rank == 0 else None
    # max over ranks of the summed device time / wall time
    tdev = torch.tensor([sum(dev_ms), sum(e2e_s)], dtype=torch.float64, device="cuda")
    npts_local = torch.tensor([last_tm["n_points"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tdev, op=dist.ReduceOp.MAX)
        dist.all_reduce(npts_local, op=dist.ReduceOp.SUM)
    dev_total_ms, e2e_total_s = float(tdev[0]), float(tdev[1])
    job_points = float(npts_local[0])                   # accepted points of all ranks per step
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # the full epipolar sweep of the same batch (BASELINE configs 2-4's kernel), timed on its own: 1 warm-up + 2 runs
    sweep = None
    if args.workload == "c2" or args.workload == "small":
        dev.epipolar_intersect_device(seeds)
        runs = [dev.epipolar_intersect_device(seeds) for _ in range(2)]
        sweep = {k: float(np.mean([r[k] for r in runs])) for k in ("k1_count_ms", "k1_fill_ms", "scan_ms")}
        sweep.update({k: int(runs[-1][k]) for k in ("n_hits", "n_segment_tests", "k1_algorithmic_bytes")})

    value = job_points * args.steps / (dev_total_ms / 1e3)
    e2e_value = job_points * args.steps / e2e_total_s
    peak, peak_src = measured_peaks()
    tr = ncu_traffic()
    t_k3b = tr.get("k3b_expand_kernel")
    t_k1 = [tr.get("k1_sweep_kernel<count>"), tr.get("k1_sweep_kernel<fill>")]
    traffic_k3b = (t_k3b["dram_read_bytes"] + t_k3b["dram_write_bytes"]) if (t_k3b and args.workload == "c2" and not args.seeds_limit) else None
    traffic_k1 = (sum(x["dram_read_bytes"] + x["dram_write_bytes"] for x in t_k1) / 2) if (all(t_k1) and args.workload == "c2" and not args.seeds_limit) else None
    icache = {}
    ip = os.path.join(ROOT, "profiles", "r01_icache.json")
    if os.path.exists(ip) and args.workload == "c2" and not args.seeds_limit:
        icache = json.load(open(ip))
    k1c = float(np.mean([a for a, _ in k1_ms])); k1f = float(np.mean([b for _, b in k1_ms])); k3 = float(np.mean(k3_ms)); k3a = float(np.mean(k3a_ms)); k3b = float(np.mean(k3b_ms))
    step_ms = dev_total_ms / args.steps
    k1_bytes = sweep["k1_algorithmic_bytes"] if sweep else last_tm["k1_algorithmic_bytes"]
    k1_avg_launch_ms = (sweep["k1_count_ms"] + sweep["k1_fill_ms"]) / 2 if sweep else (k1c + k1f) / 2
    # K3 algorithmic bytes: the hit lists it reads (16 B/hit) + the observations it writes (20 B/obs) + point headers
    # K3b algorithmic bytes: the hit lists of the accepted seeds it reads (16 B/hit) + the observations (20 B) and point headers it writes
    acc_frac = last_tm["n_accepted_seeds"] / max(1, last_tm["n_seeds"])
    full_hits = sweep["n_hits"] if sweep else last_tm["n_hits"]
    k3b_bytes = 16 * full_hits * acc_frac + 20 * last_tm["n_obs"] + 24 * last_tm["n_points"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.workload, cfg, per_view, n_gpus),
                   "l2": "flushed between iterations (256 MiB write); per-step hit lists (GBs) exceed L2 anyway",
                   "seeds_per_gpu": len(seeds), "points_per_step": job_points, "scene_upload_ms": scene_ms,
                   "parallelism": f"starting views dealt round-robin to {n_gpus} GPU(s); one NCCL all-gather of accepted records" if n_gpus > 1 else "1 GPU"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": seeds.nbytes(), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": 1e3 * e2e_total_s / args.steps,
                "note": "eg3d_match_seeds (pinned seeds H2D + kernels) + eg3d_points_get (D2H into pinned memory); the scene handle "
                        "is resident, as the reference's PLGs / plmaps are across its per-match calls"},
        "gpu_launches": launches,
        "roofline": {"kernel": "k3b_expand_kernel (view expansion of the accepted seeds: warm-started FP64 Gauss-Newton + polyline walks)",
                     "bound": "hbm", "achieved": k3b_bytes / (k3b * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": k3b_bytes / (k3b * 1e-3) / 1e9 / peak, "traffic": traffic_k3b, "peak_source": peak_src, "share_of_step": k3b / step_ms,
                     "traffic_note": "bytes per launch from the committed ncu capture (profiles/r01_traffic.json); above the algorithmic "
                                     "bytes because the per-lane stack frames and per-warp scratch arenas of the scalar walk spill past L1 "
                                     "(16 warps/SM keep them inside the L2: 21 GB; at 28 warps/SM the same kernel moved 93 GB at the same speed)",
                     "algorithmic_bytes_per_launch": k3b_bytes, "avg_launch_ms": k3b,
                     "note": "dominant kernel of the step; bound by INSTRUCTION-CACHE refills, not by bandwidth or occupancy (ncu: GPC "
                             "instruction-cache request rate at 78-80 % of peak, same kernel time with 8..32 resident warps/SM, "
                             "profiles/r01_k3b_icache.md); the HBM fraction is reported because the contract asks for the dominant "
                             "kernel; see roofline_k1 for north_star's epipolar-intersection kernel and DESIGN.md §5"},
        "roofline_k1": {"kernel": "k1_sweep_kernel (epipolar intersection, north_star's roofline kernel)", "bound": "hbm",
                        "achieved": k1_bytes / (k1_avg_launch_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": k1_bytes / (k1_avg_launch_ms * 1e-3) / 1e9 / peak, "traffic": traffic_k1, "peak_source": peak_src,
                        "share_of_step": (float(np.mean(k1any_ms)) + k1c + k1f) / step_ms, "launches_per_sweep": 2, "avg_launch_ms": k1_avg_launch_ms,
                        "algorithmic_bytes_per_launch": k1_bytes, "segment_tests_per_launch": sweep["n_segment_tests"] if sweep else last_tm["n_segment_tests"],
                        "full_sweep_ms": sweep, "hits_full_sweep": full_hits, "hits_materialised_in_step": last_tm["n_hits"],
                        "note": "FULL sweep (count + fill passes over every seed x view pair) timed on its own after the steps; inside "
                                "the step K1 is lazy (kernel_ms.k1_*; share_of_step is that lazy K1).  Algorithmic (streaming) bytes per "
                                "SURVEY 8(d); a view's segments are staged once per CTA in shared memory, so real DRAM traffic is far "
                                "lower and the fraction exceeds 1"},
        "roofline_icache": ({"kernel": "k3b_expand_kernel", "bound": "GPC instruction-cache request rate (what actually limits K3, profiles/r01_k3b_icache.md)",
                             "achieved": icache["k3b_expand_kernel"]["instruction_requests"] / (k3b * 1e-3), "peak": icache["gcc_peak_requests_per_s"],
                             "unit": "instruction-line requests/s", "frac": icache["k3b_expand_kernel"]["instruction_requests"] / (k3b * 1e-3) / icache["gcc_peak_requests_per_s"],
                             "note": "request count per launch from the committed ncu counter capture (profiles/r01_icache.json), divided by the "
                                     "kernel time measured live in this run; ncu itself reports 77 % (k3b) and 92 % (k3a) of peak"} if icache else None),
        "kernel_ms": {"k1_any": float(np.mean(k1any_ms)), "k1_count": k1c, "k1_fill": k1f, "scan_select": float(np.mean(scan_ms)), "k3a_hypothesis": k3a, "k3b_expand": k3b,
                      "pack": float(np.mean(pack_ms)), "allgather": float(np.mean(ag_all))},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and n_gpus == 1:
        cb, _, _ = cpu_baseline(scene, seeds)
        line["cpu_baseline"] = cb
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
