"""N>1 path on CPU: two gloo ranks shard the starting views, each computes its shard (with the oracle standing in as the
per-rank compute — this test is about the shard plan and the gather/merge, not the kernels), all-gather the accepted
records and check that every rank ends up with exactly the unsharded result in the reference's order."""
import os
import socket
import sys
import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch.distributed as dist
    from edgegraph3d_b200 import synthetic as syn, multigpu as mg
    from tests import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = syn.make_scene(n_views=6, n_curves=12, seed=21)
    cands = syn.curve_candidate_sets(sc, seed=21)
    osc = O.OracleScene(sc)
    lo, hi = mg.view_block(sc.n_views, world, rank)
    local = osc.match_polyline_sets(cands, lo, hi, n_threads=2)
    merged, cnts = mg.all_gather_points(local, dist)
    full = osc.match_polyline_sets(cands, n_threads=2)
    # the unsharded call iterates (set, view): compare as multisets of (xyz, observation identity), and per-rank order
    key = lambda p: sorted((tuple(x.tolist()), k[2]) for x, k in zip(p.xyz, p.identity_keys()))
    ok = merged.n_points == full.n_points and key(merged) == key(full)
    # rank-order concatenation: rank r's block is contiguous and equal to its local result
    start = int(cnts[:rank, 0].sum())
    ok = ok and np.array_equal(merged.xyz[start:start + local.n_points], local.xyz)
    ok = ok and np.array_equal(merged.obs_off[start:start + local.n_points + 1] - merged.obs_off[start], local.obs_off)
    # round-robin shard plan on the all-segment sweep form + keyed merge: exactly the unsharded result, in its order
    seeds_all = syn.sample_seeds(O.sample_seeds, sc, per_view=10)
    mine = np.where(np.isin(seeds_all.view, mg.view_round_robin(sc.n_views, world, rank)))[0]
    loc2 = osc.match_seeds(seeds_all.take(mine), n_threads=2)
    m2, _ = mg.all_gather_points(loc2, dist, order_keys=mine)
    full2 = osc.match_seeds(seeds_all, n_threads=2)
    ok = ok and m2.n_points == full2.n_points and full2.n_points > 0 and np.array_equal(m2.seed, full2.seed)
    ok = ok and np.array_equal(m2.chain_pos, full2.chain_pos) and np.array_equal(m2.obs_off, full2.obs_off)
    ok = ok and np.array_equal(m2.xyz, full2.xyz) and np.array_equal(m2.obs_seg, full2.obs_seg) and np.array_equal(m2.obs_xy, full2.obs_xy)
    q.put((rank, bool(ok), merged.n_points, int(cnts[:, 0].sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert res[0][2] == res[1][2] == res[0][3] > 0


def test_balanced_view_blocks():
    from edgegraph3d_b200 import multigpu as mg
    blocks = mg.balanced_view_blocks([10, 0, 30, 5, 5, 50, 0, 20], 4)
    assert blocks[0][0] == 0 and blocks[-1][1] == 8
    assert all(b[0] <= b[1] for b in blocks) and all(blocks[i][1] == blocks[i + 1][0] for i in range(3))
    assert mg.view_block(200, 8, 7) == (175, 200) and mg.view_block(6, 2, 0) == (0, 3)
    assert mg.view_round_robin(10, 4, 1) == [1, 5, 9] and sorted(sum((mg.view_round_robin(7, 3, r) for r in range(3)), [])) == list(range(7))


def _pipeline_worker(rank, world, port, q):
    """pipeline.edge_reconstruction with two ranks: view-sharded pipelines 1-2 over SEVERAL candidate sets (merged back into
    the reference's set-major order by seed-ordinal keys), track-sharded pipeline 3, then the order-dependent density
    limiter and the filter — every array must equal the single-process run."""
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch.distributed as dist
    from edgegraph3d_b200 import synthetic as syn, pipeline as P
    from edgegraph3d_b200.scene import CandidateSets
    from tests import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = syn.make_scene(n_views=6, n_curves=12, seed=21, n_tracks=60)
    c1 = syn.curve_candidate_sets(sc, seed=21)
    V = sc.n_views
    # a second family of sets: the first one regrouped two by two (still several sets, other boundaries)
    groups = [[sorted(set(c1.polyline[c1.off[i * V + v]:c1.off[i * V + v + 1]].tolist()) |
                      (set(c1.polyline[c1.off[(i + 1) * V + v]:c1.off[(i + 1) * V + v + 1]].tolist()) if i + 1 < c1.n_sets else set()))
               for v in range(V)] for i in range(0, c1.n_sets, 2)]
    c2 = CandidateSets.from_lists(groups, V)
    dev = O.OracleDevice(sc, n_threads=2)
    single = P.edge_reconstruction(dev, sc, c1, c2)
    sharded = P.edge_reconstruction(dev, sc, c1, c2, dist=dist)
    ok = c1.n_sets > 2 and c2.n_sets > 1 and single["points"].n_points > 50
    a, b = single["points"], sharded["points"]
    ok = ok and a.n_points == b.n_points and np.array_equal(a.xyz, b.xyz) and np.array_equal(a.obs_off, b.obs_off)
    ok = ok and np.array_equal(a.obs_view, b.obs_view) and np.array_equal(a.obs_poly, b.obs_poly) and np.array_equal(a.obs_seg, b.obs_seg)
    ok = ok and np.array_equal(a.obs_xy, b.obs_xy) and np.array_equal(a.chain_pos, b.chain_pos)
    ok = ok and np.array_equal(single["keep"], sharded["keep"]) and np.array_equal(single["inliers"], sharded["inliers"])
    ok = ok and np.array_equal(single["filtered_xyz"], sharded["filtered_xyz"])
    # the plain rank-order concatenation would NOT have been the reference order for the multi-set pipelines
    lo, hi = (rank * V) // world, ((rank + 1) * V) // world
    local = dev.match_polyline_sets(c1, lo, hi)[0]
    from edgegraph3d_b200 import multigpu as mg
    naive, _ = mg.all_gather_points(local, dist)
    differs = naive.n_points == single["parts"][0].n_points and not np.array_equal(naive.xyz, single["parts"][0].xyz)
    q.put((rank, bool(ok), bool(differs), a.n_points))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_pipeline_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res), res
    assert all(r[2] for r in res), res      # the keyed merge is doing real work on this input
