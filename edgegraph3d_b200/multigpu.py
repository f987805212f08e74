"""Multi-GPU plumbing of the hot path: shard planning and the one exchange step.

The path shards by independent units (seeds): ranks own contiguous blocks of STARTING VIEWS (the reference's outer loop,
polyline_matching.cpp:162) or of SfM point ids (plg_matching_from_refpoints.cpp:90) and hold the full read-only scene.
The only exchange is one all-gather of the accepted records before the order-dependent density limiter and the filter;
concatenating the shards in rank order reproduces the reference's loop order, so any GPU count gives identical results.
torch.distributed is plumbing here (NCCL over NVLink on GPUs; the same code runs on gloo/CPU tensors in the tests).
"""
import numpy as np
from .scene import PointSet


def view_block(n_views, world_size, rank):
    """Contiguous block of starting views owned by `rank`."""
    return (rank * n_views) // world_size, ((rank + 1) * n_views) // world_size


def view_round_robin(n_views, world_size, rank):
    """Starting views rank, rank + world_size, ...: per-view cost varies slowly with the view index (neighbouring cameras
    see similar geometry), so interleaving balances the ranks better than contiguous blocks (measured at N=8: the slowest
    block rank is ~20 % above the fastest).  The merged result is put back into the reference's loop order with
    `order_keys` (global seed ordinals) in all_gather_points."""
    return list(range(rank, n_views, world_size))


def balanced_view_blocks(seeds_per_view, world_size):
    """Contiguous view blocks with (nearly) equal seed counts: prefix sums over the per-view seed counts."""
    c = np.concatenate([[0], np.cumsum(np.asarray(seeds_per_view, np.int64))])
    total = int(c[-1])
    cuts = [0]
    for r in range(1, world_size):
        cuts.append(int(np.searchsorted(c, total * r / world_size, side="left")))
    cuts.append(len(seeds_per_view))
    cuts = np.maximum.accumulate(np.minimum(cuts, len(seeds_per_view)))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world_size)]


def polyline_set_seed_views(scene, cands, spacing, sampler):
    """Starting view of every seed of eg3d_match_polyline_sets(cands) in the reference's loop order (candidate set,
    starting view, candidate polyline, 20 px steps: pipelines.cpp:92-100 around polyline_matching.cpp:162-190).  A rank that
    owns the starting views [lo, hi) computes the sub-sequence with lo <= view < hi, so
    `np.where((views >= lo) & (views < hi))[0]` are that rank's `order_keys` for all_gather_points: with several candidate
    sets a plain rank-order concatenation would be view-major, not set-major, and the order-dependent density limiter
    (filtering_close_plgps.cpp:75-124) would see a different sequence.  `sampler` = lib.sample_seeds (host C++)."""
    V = scene.n_views
    counts = np.diff(cands.off).astype(np.int64)                      # per (set, view) range, already in (set, view) order
    views = np.repeat(np.tile(np.arange(V, dtype=np.int32), cands.n_sets), counts)
    sv = sampler(scene, views, cands.polyline, spacing)[0]
    return np.asarray(sv, np.int32)


def track_block(n_tracks, world_size, rank):
    """Contiguous block of SfM point ids owned by `rank` (plg_matching_from_refpoints.cpp:90): rank-order concatenation of
    the blocks is the reference's loop order."""
    return (rank * n_tracks) // world_size, ((rank + 1) * n_tracks) // world_size


_FIELDS = (("xyz", np.float32, 3), ("seed", np.int32, 1), ("chain_pos", np.int32, 1), ("obs_view", np.int32, 1),
           ("obs_poly", np.uint32, 1), ("obs_seg", np.uint32, 1), ("obs_xy", np.float32, 2))


def all_gather_points(local, dist, device=None, order_keys=None):
    """HOST-STAGED all-gather of a PointSet over torch.distributed (any backend): the CPU / gloo form of the exchange, used by
    the tests and by callers without a device communicator.  On GPUs the exchange lives in the library
    (eg3d_points_allgather through lib.DeviceScene.points_allgather: NCCL, device-side merge); pipeline.run_pipelines picks it
    whenever the scene handle has a communicator.  Returns the concatenation in rank order on every rank; without
    `order_keys` the `seed` field keeps the rank-LOCAL ordinals.

    `local` holds this rank's accepted points (host numpy arrays); tensors are staged on `device` (cuda for NCCL, cpu for
    gloo).  One count all-gather + one padded all-gather per field.  `order_keys` (int64 per local SEED: its ordinal in
    the unsharded seed list) makes the merged result independent of the shard plan: points are sorted by (global seed
    ordinal, chain position) = the reference's loop order, and `seed` is rewritten to the global ordinal.
    """
    import torch
    world = dist.get_world_size()
    dev = device if device is not None else "cpu"
    lens = np.diff(local.obs_off).astype(np.int64)
    cnt = torch.tensor([local.n_points, local.n_obs], dtype=torch.int64, device=dev)
    cnts = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    cnts = torch.stack(cnts).cpu().numpy()
    maxn, maxm = int(cnts[:, 0].max()), int(cnts[:, 1].max())

    def gather(arr, rows, width, dtype):
        pad = np.zeros((rows, width), dtype)
        a = np.asarray(arr, dtype).reshape(-1, width)
        pad[:a.shape[0]] = a
        t = torch.from_numpy(pad.view(np.uint8).reshape(-1).copy()).to(dev)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [o.cpu().numpy().view(dtype).reshape(rows, width) for o in outs]

    got = {}
    for name, dt, w in _FIELDS:
        rows = maxm if name.startswith("obs_") else maxn
        got[name] = gather(getattr(local, name), rows, w, dt)
    got_len = gather(lens, maxn, 1, np.int64)
    got_key = None
    if order_keys is not None:
        pk = np.asarray(order_keys, np.int64)[local.seed] if local.n_points else np.zeros(0, np.int64)
        got_key = gather(pk, maxn, 1, np.int64)
    parts = []
    for r in range(world):
        n, m = int(cnts[r, 0]), int(cnts[r, 1])
        off = np.concatenate([[0], np.cumsum(got_len[r][:n, 0])]).astype(np.int64)
        parts.append(PointSet(got["xyz"][r][:n], got["seed"][r][:n, 0], got["chain_pos"][r][:n, 0], off,
                              got["obs_view"][r][:m, 0], got["obs_poly"][r][:m, 0], got["obs_seg"][r][:m, 0], got["obs_xy"][r][:m]))
    merged = PointSet.concat(parts)
    if got_key is not None and merged.n_points:
        keys = np.concatenate([got_key[r][:int(cnts[r, 0]), 0] for r in range(world)])
        order = np.lexsort((merged.chain_pos, keys))
        merged = merged.take(order)
        merged.seed = keys[order].astype(np.int32 if keys.max(initial=0) < 2 ** 31 else np.int64)
    return merged, cnts
