"""GPU parity tests proper (-m gpu): the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded
inputs.  Bars: bit-exact for index / integer / 2D-geometry outputs (hit lists, identity sets, inlier bitmaps);
3D coordinates within 1e-4 reprojection MSE of the oracle (north_star) — and in practice far tighter."""
import numpy as np
import pytest
from edgegraph3d_b200 import lib as E, synthetic as syn
from edgegraph3d_b200.scene import FlatScene
from tests import oracle_lib as O

pytestmark = pytest.mark.gpu


def _reproj_mse(scene, pts):
    """per-point mean squared reprojection error over its observations (px^2)"""
    P = scene.cameras.astype(np.float64).reshape(-1, 3, 4)
    out = np.zeros(pts.n_points)
    for i in range(pts.n_points):
        a, b = int(pts.obs_off[i]), int(pts.obs_off[i + 1])
        Pv = P[pts.obs_view[a:b]]
        h = Pv[:, :, :3] @ pts.xyz[i].astype(np.float64) + Pv[:, :, 3]
        r = pts.obs_xy[a:b] - h[:, :2] / h[:, 2:3]
        out[i] = (r ** 2).sum() / (2 * (b - a))
    return out


def assert_points_parity(scene, gpu, ref, mse_tol=1e-4):
    assert gpu.n_points == ref.n_points, (gpu.n_points, ref.n_points)
    assert np.array_equal(gpu.seed, ref.seed) and np.array_equal(gpu.chain_pos, ref.chain_pos)
    assert np.array_equal(gpu.obs_off, ref.obs_off)
    assert np.array_equal(gpu.obs_view, ref.obs_view)
    assert np.array_equal(gpu.obs_poly, ref.obs_poly) and np.array_equal(gpu.obs_seg, ref.obs_seg)
    assert gpu.obs_xy.tobytes() == ref.obs_xy.tobytes()      # 2D geometry is float arithmetic in a fixed order: bit-exact
    if gpu.n_points:
        d = np.abs(_reproj_mse(scene, gpu) - _reproj_mse(scene, ref))
        assert d.max() <= mse_tol, d.max()
        assert np.abs(gpu.xyz - ref.xyz).max() < 1e-4


@pytest.fixture(scope="module")
def small():
    sc = syn.make_scene(n_views=6, n_curves=24, seed=1, n_tracks=150)
    return sc, E.DeviceScene(sc), O.OracleScene(sc)


def assert_hits_equal(a, b):
    off_a, h_a, V_a = a[0], a[1], a[2]
    off_b, h_b, V_b = b[0], b[1], b[2]
    assert V_a == V_b
    assert np.array_equal(off_a, off_b)
    assert h_a.tobytes() == h_b.tobytes()


def test_k1_sweep_bit_exact(small):
    sc, dev, orc = small
    seeds = syn.sample_seeds(E.sample_seeds, sc)
    g = dev.epipolar_intersect(seeds)
    assert_hits_equal(g, orc.epipolar_intersect(seeds))
    assert g[1].shape[0] > 1000
    tm = g[3]
    assert tm["n_segment_tests"] == sum(sum(sc.n_segments(v) for v in range(sc.n_views) if v != s) for s in seeds.view)


def test_k1_sweep_ragged_and_multichunk():
    # a view with > K1_CHUNK segments (several TMA stages, ragged tail), views with invalid polylines, and an invalid F pair
    sc = syn.make_scene(n_views=4, width=1920, height=1080, focal=1600.0, n_curves=260, segs_per_curve=20, curve_len=0.2,
                        seed=11, extent=0.9, closed_frac=0.05)
    sc.fundamental_valid[1, 2] = 0
    assert max(sc.n_segments(v) for v in range(4)) > 4096
    seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=70)
    with E.DeviceScene(sc) as dev:
        g = dev.epipolar_intersect(seeds)
    assert_hits_equal(g, O.OracleScene(sc).epipolar_intersect(seeds))


def test_k1_empty_batch(small):
    sc, dev, orc = small
    seeds = syn.sample_seeds(E.sample_seeds, sc).slice(0, 0)
    off, h, V, _ = dev.epipolar_intersect(seeds)
    assert off.tolist() == [0] and len(h) == 0


def test_k1_candidates_bit_exact(small):
    sc, dev, orc = small
    cands = syn.curve_candidate_sets(sc)
    seeds = syn.sample_seeds(E.sample_seeds, sc, polylines_per_view=6)
    seeds.cand_set = (seeds.polyline.astype(np.int32)) % cands.n_sets
    assert_hits_equal(dev.epipolar_intersect(seeds, cands), orc.epipolar_intersect(seeds, cands))


def test_match_seeds_sweep_parity(small):
    sc, dev, orc = small
    seeds = syn.sample_seeds(E.sample_seeds, sc)
    gpu, tm = dev.match_seeds(seeds)
    ref = orc.match_seeds(seeds)
    assert ref.n_points > 300
    assert_points_parity(sc, gpu, ref)
    assert tm["n_points"] == gpu.n_points and tm["kernel_launches"] >= 5


def test_match_polyline_sets_parity(small):
    sc, dev, orc = small
    cands = syn.curve_candidate_sets(sc)
    gpu, _ = dev.match_polyline_sets(cands)
    ref = orc.match_polyline_sets(cands)
    assert ref.n_points > 1000
    assert_points_parity(sc, gpu, ref)


def test_match_polyline_sets_view_shards_concatenate(small):
    # the multi-GPU shard axis (starting views): shards concatenate to the unsharded result
    sc, dev, orc = small
    cands = syn.curve_candidate_sets(sc)
    full, _ = dev.match_polyline_sets(cands)
    a, _ = dev.match_polyline_sets(cands, 0, 3)
    b, _ = dev.match_polyline_sets(cands, 3, 6)
    assert a.n_points + b.n_points == full.n_points
    keys = sorted((tuple(x.tolist()), k[2]) for x, k in zip(full.xyz, full.identity_keys()))
    keys2 = sorted((tuple(x.tolist()), k[2]) for p in (a, b) for x, k in zip(p.xyz, p.identity_keys()))
    assert keys == keys2


@pytest.mark.parametrize("seed,views,curves", [(2, 5, 16), (7, 9, 20), (13, 12, 12)])
def test_match_more_scenes_parity(seed, views, curves):
    sc = syn.make_scene(n_views=views, n_curves=curves, seed=seed, closed_frac=0.2, drop_view_frac=0.1)
    cands = syn.curve_candidate_sets(sc, seed=seed)
    with E.DeviceScene(sc) as dev:
        gpu, _ = dev.match_polyline_sets(cands)
        seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=25)
        gpu2, _ = dev.match_seeds(seeds)
    orc = O.OracleScene(sc)
    assert_points_parity(sc, gpu, orc.match_polyline_sets(cands))
    assert_points_parity(sc, gpu2, orc.match_seeds(seeds))


def test_gn64_golden_and_oracle(golden, small):
    V = golden["cams"].shape[0]
    sc = FlatScene(640, 480, golden["cams"], np.zeros((V, V, 9)), np.zeros((V, V), np.uint8), np.arange(V + 1), np.arange(V + 1) * 2,
                   np.tile(np.array([[10, 10], [20, 20]], np.float32), (V, 1)), np.zeros(V, np.uint32), np.ones(V, np.uint32))
    with E.DeviceScene(sc) as dev:
        xyz, mse, ok, _ = dev.gn_triangulate(golden["gn64_off"], golden["gn64_view"], golden["gn64_xy"], golden["gn64_init"].astype(np.float32), True)
    assert np.array_equal(ok, golden["gn64_ok"])
    sel = ok == 1
    assert np.abs(xyz[sel] - golden["gn64_X"][sel].astype(np.float32)).max() < 2e-6
    assert np.abs(mse - golden["gn64_mse"].astype(np.float32)).max() < 1e-4 * max(1.0, np.abs(golden["gn64_mse"]).max())


def test_gn32_filter_bit_exact_vs_cv2_golden(golden):
    V = golden["cams"].shape[0]
    sc = FlatScene(640, 480, golden["cams"], np.zeros((V, V, 9)), np.zeros((V, V), np.uint8), np.arange(V + 1), np.arange(V + 1) * 2,
                   np.tile(np.array([[10, 10], [20, 20]], np.float32), (V, 1)), np.zeros(V, np.uint32), np.ones(V, np.uint32))
    with E.DeviceScene(sc) as dev:
        xyz, mse, ok, _ = dev.gn_triangulate(golden["gn32_off"], golden["gn32_view"], golden["gn32_xy"], golden["gn32_init"].astype(np.float32), False)
    assert np.array_equal(ok, golden["gn32_ok"])
    sel = ok == 1
    assert np.array_equal(xyz[sel], golden["gn32_X"][sel].astype(np.float32))
    assert np.array_equal(mse, golden["gn32_mse"].astype(np.float32))


def test_dedup_and_filter_parity(small):
    sc, dev, orc = small
    cands = syn.curve_candidate_sets(sc)
    pts, _ = dev.match_polyline_sets(cands)
    keep = dev.dedup_close_points(pts)
    assert np.array_equal(keep, orc.dedup_close_points(pts))
    assert 0 < keep.sum() < len(keep)
    # filter over SfM tracks + kept edge points (edge_matcher.cpp:132)
    kept = np.where(keep)[0]
    xyz = np.concatenate([sc.track_xyz, pts.xyz[kept]])
    offs = [sc.track_off]
    views = [sc.track_view]
    xys = [sc.track_xy]
    base = int(sc.track_off[-1])
    lens = (pts.obs_off[kept + 1] - pts.obs_off[kept])
    offs.append(base + np.cumsum(lens))
    idx = np.concatenate([np.arange(pts.obs_off[i], pts.obs_off[i + 1]) for i in kept])
    views.append(pts.obs_view[idx]); xys.append(pts.obs_xy[idx])
    obs_off = np.concatenate(offs); obs_view = np.concatenate(views); obs_xy = np.concatenate(xys)
    gx, gi, _ = dev.filter(xyz, obs_off, obs_view, obs_xy, sc.n_tracks)
    ox, oi = orc.filter(xyz, obs_off, obs_view, obs_xy, sc.n_tracks)
    assert np.array_equal(gi, oi)
    assert np.array_equal(gx, ox)      # FP32 filter arithmetic is reproduced operation by operation
    assert 0 < gi.sum() < len(gi)


def test_match_refpoints_parity(small):
    # pipeline 3 (a6): seeds from SfM tracks via the 30 px grid, hits filtered by the per-seed radius
    sc, dev, orc = small
    gpu, tm = dev.match_refpoints()
    ref = orc.match_refpoints()
    assert ref.n_points > 200
    assert_points_parity(sc, gpu, ref)
    # a shard of the refpoint ids (multi-GPU axis, plg_matching_from_refpoints.cpp:90)
    a, _ = dev.match_refpoints(0, 70)
    b, _ = dev.match_refpoints(70, sc.n_tracks)
    assert a.n_points + b.n_points == gpu.n_points
    assert np.array_equal(np.concatenate([a.xyz, b.xyz]), gpu.xyz)


def test_refpoint_seeding_capacities_grow_inside_the_call(small, monkeypatch):
    """The device seeding of pipeline 3 holds a bounded number of polyline ids per 30 px neighbourhood; the reference has no
    bound, so the bound grows inside the call (x4 per re-run).  Started from absurdly small capacities the result is unchanged."""
    sc, dev, orc = small
    full, _ = dev.match_refpoints()
    monkeypatch.setenv("EG3D_TEST_A6_CAPS", "4,4")
    tiny, tm = dev.match_refpoints()
    monkeypatch.delenv("EG3D_TEST_A6_CAPS")
    assert tm["n_capacity_retries"] >= 1
    assert tiny.n_points == full.n_points and tiny.xyz.tobytes() == full.xyz.tobytes()
    assert np.array_equal(tiny.obs_off, full.obs_off) and np.array_equal(tiny.obs_poly, full.obs_poly) and tiny.obs_xy.tobytes() == full.obs_xy.tobytes()


def test_output_capacity_retry(small, monkeypatch):
    # the internal output buffers start tiny (test knob) and the call transparently re-runs K3 with larger ones
    sc, dev, orc = small
    cands = syn.curve_candidate_sets(sc)
    ref, _ = dev.match_polyline_sets(cands)
    monkeypatch.setenv("EG3D_TEST_TINY_CAPS", "1")
    again, tm = dev.match_polyline_sets(cands)
    assert again.n_points == ref.n_points > 64 and np.array_equal(again.xyz, ref.xyz) and np.array_equal(again.obs_off, ref.obs_off)
    assert tm["kernel_launches"] > 10          # more than one K3 attempt


def test_per_seed_capacities_grow_inside_the_call(small):
    """max_chain_points / max_follow_points are STARTING sizes: a batch with longer chains is re-run inside the call with doubled
    capacities (real polyline graphs give chains of 140+ points against the default 96) — same result as with roomy capacities,
    never a truncated chain, never an error the caller has to handle."""
    sc, dev, orc = small
    cands = syn.curve_candidate_sets(sc)
    ref, tm0 = dev.match_polyline_sets(cands)
    assert tm0["n_capacity_retries"] == 0
    with E.DeviceScene(sc, E.default_params(max_chain_points=3, max_follow_points=4)) as tiny:
        got, tm = tiny.match_polyline_sets(cands)
        assert tm["n_capacity_retries"] >= 2
        for f in ("xyz", "seed", "chain_pos", "obs_off", "obs_view", "obs_poly", "obs_seg", "obs_xy"):
            assert np.array_equal(getattr(got, f), getattr(ref, f)), f
        g3, tm3 = tiny.match_refpoints(0, sc.n_tracks)
        w3, _ = dev.match_refpoints(0, sc.n_tracks)
        assert tm3["n_capacity_retries"] >= 1 and np.array_equal(g3.obs_off, w3.obs_off) and np.array_equal(g3.xyz, w3.xyz)


def test_bad_arguments_are_refused(small):
    """Out-of-range views / polyline ids / ranges come back as EG3D_ERR_INVALID_ARG instead of reading past the caller's arrays."""
    from edgegraph3d_b200 import _abi as A
    from edgegraph3d_b200.scene import SeedBatch, CandidateSets, PointSet
    sc, dev, _ = small
    seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=5)
    bad = SeedBatch(seeds.view.copy(), seeds.polyline.copy(), seeds.segment, seeds.xy, None)
    bad.view[0] = sc.n_views
    with pytest.raises(E.Eg3dError) as ei:
        dev.match_seeds(bad)
    assert ei.value.status == A.EG3D_ERR_INVALID_ARG
    bad.view[0] = 0; bad.polyline[0] = 10 ** 6
    with pytest.raises(E.Eg3dError) as ei:
        dev.match_seeds(bad)
    assert ei.value.status == A.EG3D_ERR_INVALID_ARG
    cands = syn.curve_candidate_sets(sc)
    with pytest.raises(E.Eg3dError) as ei:
        dev.match_polyline_sets(cands, 0, sc.n_views + 1)
    assert ei.value.status == A.EG3D_ERR_INVALID_ARG
    c2 = CandidateSets(cands.n_sets, cands.off.copy(), cands.polyline.copy())
    c2.polyline[0] = 10 ** 6
    with pytest.raises(E.Eg3dError) as ei:
        dev.match_polyline_sets(c2)
    assert ei.value.status == A.EG3D_ERR_INVALID_ARG
    pts, _ = dev.match_seeds(seeds)
    if pts.n_obs:
        xy = pts.obs_xy.copy(); xy[0] = (-5.0, 10.0)
        outside = PointSet(pts.xyz, pts.seed, pts.chain_pos, pts.obs_off, pts.obs_view, pts.obs_poly, pts.obs_seg, xy)
        with pytest.raises(E.Eg3dError) as ei:
            dev.dedup_close_points(outside)
        assert ei.value.status == A.EG3D_ERR_INVALID_ARG


def test_lazy_sweep_equals_full_sweep(small):
    """eg3d_match_seeds materialises only the hit lists its per-seed code reads (any-hit pass, the three selected views,
    every view of the accepted seeds); EG3D_K1_FULL=1 makes it run the reference's full sweep first.  Same result,
    bit for bit, from far fewer hit records — and both equal the oracle."""
    import os
    sc, dev, orc = small
    seeds = syn.sample_seeds(E.sample_seeds, sc)
    lazy, tm_lazy = dev.match_seeds(seeds)
    os.environ["EG3D_K1_FULL"] = "1"
    try:
        full, tm_full = dev.match_seeds(seeds)
    finally:
        del os.environ["EG3D_K1_FULL"]
    assert lazy.n_points == full.n_points > 0
    for f in ("seed", "chain_pos", "obs_off", "obs_view", "obs_poly", "obs_seg"):
        assert np.array_equal(getattr(lazy, f), getattr(full, f)), f
    assert lazy.obs_xy.tobytes() == full.obs_xy.tobytes() and lazy.xyz.tobytes() == full.xyz.tobytes()
    assert tm_lazy["k1_any_ms"] > 0 and tm_full["k1_any_ms"] == 0
    assert 0 < tm_lazy["n_hits"] < tm_full["n_hits"]
    assert_points_parity(sc, lazy, orc.match_seeds(seeds))


def test_k1_device_only_entry_point_counts_the_same_hits(small):
    sc, dev, _ = small
    seeds = syn.sample_seeds(E.sample_seeds, sc)
    off, hits, V, tm = dev.epipolar_intersect(seeds)
    tm_dev = dev.epipolar_intersect_device(seeds)
    assert tm_dev["n_hits"] == hits.shape[0] == int(off[-1])
    assert tm_dev["n_segment_tests"] == tm["n_segment_tests"] and tm_dev["k1_algorithmic_bytes"] == tm["k1_algorithmic_bytes"]


@pytest.mark.parametrize("overrides", [dict(dlt_wellposed=1), dict(filter_abs_int=1), dict(split_interval_distance=7.0, follow_first_image_distance=6.0),
                                        dict(max_proj_distsq_expand=4.0, quasiparallel_cos=0.9)])
def test_parameter_variants_parity(overrides):
    """The policy switches and a few of the reference's compile-time constants (eg3d_params) away from their defaults:
    sweep form, candidate form and the outlier filter still equal the oracle run with the same parameters."""
    sc = syn.make_scene(n_views=7, n_curves=20, seed=11, n_tracks=80)
    prm_g, prm_o = E.default_params(**overrides), O.default_params(**overrides)
    orc = O.OracleScene(sc, prm_o)
    with E.DeviceScene(sc, prm_g) as dev:
        seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=40, spacing=prm_g.split_interval_distance)
        assert_points_parity(sc, dev.match_seeds(seeds)[0], orc.match_seeds(seeds))
        cands = syn.curve_candidate_sets(sc, seed=11)
        g = dev.match_polyline_sets(cands)[0]
        r = orc.match_polyline_sets(cands)
        assert_points_parity(sc, g, r)
        if g.n_points:
            fx, inl, _ = dev.filter(g.xyz, g.obs_off, g.obs_view, g.obs_xy, 0)
            ox, oinl = orc.filter(r.xyz, r.obs_off, r.obs_view, r.obs_xy, 0)
            assert np.array_equal(inl, oinl) and np.array_equal(fx[inl == 1], ox[oinl == 1])


def test_degenerate_scenes():
    """Three views (the minimum for a hypothesis), every fundamental matrix but one pair invalid, and a view without
    polylines: no crash, results equal to the oracle (mostly empty)."""
    sc = syn.make_scene(n_views=3, n_curves=10, seed=5)
    seeds = syn.sample_seeds(E.sample_seeds, sc)
    with E.DeviceScene(sc) as dev:
        assert_points_parity(sc, dev.match_seeds(seeds)[0], O.OracleScene(sc).match_seeds(seeds))
    sc2 = syn.make_scene(n_views=5, n_curves=10, seed=6)
    sc2.fundamental_valid[:] = 0
    sc2.fundamental_valid[0, 1] = sc2.fundamental_valid[1, 0] = 1
    seeds2 = syn.sample_seeds(E.sample_seeds, sc2)
    with E.DeviceScene(sc2) as dev:
        g = dev.match_seeds(seeds2)[0]
        assert_points_parity(sc2, g, O.OracleScene(sc2).match_seeds(seeds2))
        assert g.n_points == 0          # fewer than three views can ever hold hits
    # a view whose polylines are all invalidated (empty vertex ranges keep their ids)
    sc3 = syn.make_scene(n_views=5, n_curves=10, seed=7)
    g0, g1 = int(sc3.view_poly_off[2]), int(sc3.view_poly_off[3])
    lens = np.diff(sc3.poly_vert_off)
    keep = np.ones(len(sc3.verts), bool)
    keep[int(sc3.poly_vert_off[g0]):int(sc3.poly_vert_off[g1])] = False
    lens[g0:g1] = 0
    sc3 = FlatScene(sc3.width, sc3.height, sc3.cameras, sc3.fundamental, sc3.fundamental_valid, sc3.view_poly_off,
                    np.concatenate([[0], np.cumsum(lens)]).astype(np.int64), sc3.verts[keep], sc3.poly_start, sc3.poly_end)
    seeds3 = syn.sample_seeds(E.sample_seeds, sc3)
    assert not (seeds3.view == 2).any()
    with E.DeviceScene(sc3) as dev:
        assert_points_parity(sc3, dev.match_seeds(seeds3)[0], O.OracleScene(sc3).match_seeds(seeds3))


def test_match_correspondences_is_the_consensus_step_alone(small):
    """B4 (eg3d_match_correspondences): K3 on CALLER-supplied hit lists — here the lists K1 itself produces, so the result must be
    what eg3d_match_seeds gives for the same seeds, and the oracle's; then with the lists of pipeline 3's EdgeManager side
    (eg3d_refpoint_correspondences, B3), where it must equal eg3d_match_refpoints."""
    sc, dev, orc = small
    seeds = syn.sample_seeds(E.sample_seeds, sc, per_view=40)
    off, hits, V, _ = dev.epipolar_intersect(seeds)
    g, tm = dev.match_correspondences(seeds.view, off, hits)
    assert tm["n_seeds"] == len(seeds) and g.n_points > 100
    assert_points_parity(sc, g, orc.match_seeds(seeds))
    whole, _ = dev.match_seeds(seeds)
    for f in ("xyz", "seed", "chain_pos", "obs_off", "obs_view", "obs_poly", "obs_seg", "obs_xy"):
        assert np.array_equal(getattr(g, f), getattr(whole, f)), f
    # pipeline 3: seeds + per-view hit lists of the EdgeManager side, bit-exact vs the oracle's, then the consensus step on them
    s3, track, off3, hits3 = dev.refpoint_correspondences(0, sc.n_tracks)
    o_off, o_hits, _ = orc.refpoint_hits(0, sc.n_tracks)
    assert np.array_equal(off3, o_off) and hits3.tobytes() == o_hits.tobytes() and len(s3) > 50
    assert np.all(np.diff(track) >= 0)
    g3, _ = dev.match_correspondences(s3.view, off3, hits3)
    w3, _ = dev.match_refpoints(0, sc.n_tracks)
    for f in ("xyz", "seed", "chain_pos", "obs_off", "obs_view", "obs_poly", "obs_seg", "obs_xy"):
        assert np.array_equal(getattr(g3, f), getattr(w3, f)), f
    assert_points_parity(sc, g3, orc.match_refpoints(0, sc.n_tracks))
    # empty batch
    e, _ = dev.match_correspondences(np.zeros(0, np.int32), np.zeros(1, np.int64), hits[:0])
    assert e.n_points == 0


@pytest.mark.parametrize("seed,views,curves", [(7, 9, 20), (13, 12, 12)])
def test_shared_first_iteration_kernel_variant_parity(seed, views, curves, monkeypatch):
    """k3b_expand_kernel<true> (per-slot cache of the base normal-equation sums, the variant picked for rigs with >= 400 views):
    forced on a small rig it must give the oracle's chains and observation lists, and the plain variant's points."""
    sc = syn.make_scene(n_views=views, n_curves=curves, seed=seed, closed_frac=0.15, drop_view_frac=0.1)
    dev, orc = E.DeviceScene(sc), O.OracleScene(sc)
    seeds = syn.sample_seeds(E.sample_seeds, sc)
    plain, _ = dev.match_seeds(seeds)
    monkeypatch.setenv("EG3D_GN_CACHE", "1")
    cached, _ = dev.match_seeds(seeds)
    monkeypatch.delenv("EG3D_GN_CACHE")
    assert cached.n_points > 50
    assert_points_parity(sc, cached, orc.match_seeds(seeds))
    assert_points_parity(sc, cached, plain)
