"""The library's exchange step (eg3d_comm_create / eg3d_points_allgather).  On the one GPU the driver's `-m gpu` run
has, a communicator of world size 1 still goes through every part of it — NCCL (counts all-gather, the grouped broadcast), the
byte-packed record layout, the radix sort by (global seed ordinal, chain position) and the gather pass — so the merged result
must be the rank's own result, re-keyed.  tests/mgpu_exchange_check.py is the 2/4/8-rank form (torchrun), and
tests/test_multigpu_gloo.py covers the shard plans and the merge order on CPU ranks."""
import numpy as np
import pytest
from edgegraph3d_b200 import synthetic as syn


@pytest.mark.gpu
def test_single_rank_exchange_is_the_identity_and_rekeys_seeds():
    from edgegraph3d_b200 import lib as E
    sc = syn.make_scene(n_views=8, n_curves=20, seed=3)
    seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=30)
    with E.DeviceScene(sc) as dev:
        dev.comm_create(E.comm_unique_id(), 0, 1)
        dp, _ = dev.match_seeds(seeds, fetch=False)
        own = dp.fetch()
        assert own.n_points > 100
        # rank-major keys: seed ordinals unchanged
        merged, tm = dev.points_allgather(dp, None, fetch=True)
        assert tm["n_points"] == own.n_points and tm["total_ms"] > 0
        for f in ("xyz", "seed", "chain_pos", "obs_off", "obs_view", "obs_poly", "obs_seg", "obs_xy"):
            assert np.array_equal(getattr(merged, f), getattr(own, f)), f
        # explicit global ordinals that REVERSE the seed order: chains come back last seed first, chain order inside a seed kept
        g = (len(seeds) - 1 - np.arange(len(seeds))).astype(np.int64) * 3 + 7
        rev, _ = dev.points_allgather(dp, g, fetch=True)
        order = np.lexsort((own.chain_pos, g[own.seed]))
        assert np.array_equal(rev.seed, g[own.seed][order]) and np.array_equal(rev.chain_pos, own.chain_pos[order])
        assert np.array_equal(rev.xyz, own.xyz[order])
        lens = np.diff(own.obs_off)[order]
        assert np.array_equal(np.diff(rev.obs_off), lens)
        src = np.concatenate([np.arange(own.obs_off[i], own.obs_off[i + 1]) for i in order]) if len(order) else np.zeros(0, np.int64)
        assert np.array_equal(rev.obs_view, own.obs_view[src]) and np.array_equal(rev.obs_xy, own.obs_xy[src]) and np.array_equal(rev.obs_seg, own.obs_seg[src])
        dp.free()
        dev.comm_destroy()
