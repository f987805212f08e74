"""The oracle against the REFERENCE'S OWN CODE, executed.

oracle/_ref/libref_path.so is built from the reference's unmodified hot-path sources where they lie under /root/reference
(oracle/Makefile: polyline_matching.cpp, plg_matching.cpp, triangulation.cpp, polyline_graph_2d.cpp, plg_edge_manager.cpp,
plg_matching_from_refpoints.cpp, the consensus manager, polyLine_2d_map*.cpp, the three filtering files, geometric_utilities.cpp,
edge_graph_3d_utilities.cpp, OpenMvgParser.cpp ...) against stand-in headers for OpenCV / CGAL / Boost (oracle/ref_stubs_path/;
the cv::Mat arithmetic there is the model the cv2 goldens pinned).  So the control flow of matching, PLG following, view
expansion, pipeline-3 seeding, the density limiter and the outlier filter that these tests run IS the reference's, and the
oracle (the restatement every GPU parity test is checked against) must reproduce it bit for bit: chains, observation lists,
2D coordinates and the float bits of every 3D point.

One function of the reference has undefined behaviour on inputs the path does reach: next_pl_point_by_distance falls off its
end without a return statement when the direction is neither extreme of the polyline (zero-filled direction arrays, SURVEY
A.2.16; polyline_graph_2d.cpp:391-447).  The build interposes ONE guard there (oracle/ref_path_wrapper.cpp): that case gets the
rule of the oracle and of the product ("cannot drive"), every defined call is forwarded to the reference's own body.  The
oracle reports which seeds reach it (PointSet.seed_ub: ~3 % of the synthetic seeds, ~5 % of the real ones, and most of the
productive ones); with the guard they run through the reference's code like all the others, so every comparison below is over
ALL seeds."""
import os
import numpy as np
import pytest
from edgegraph3d_b200 import synthetic as syn
from edgegraph3d_b200.scene import SeedBatch
from tests import oracle_lib as O, ref_lib as R

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref_path.so not built (needs /root/reference: make -C oracle ref)")

SCENES = [dict(n_views=6, n_curves=16, seed=4, n_tracks=60),
          dict(n_views=9, n_curves=20, seed=7, closed_frac=0.2, drop_view_frac=0.1, n_tracks=80),
          dict(n_views=7, n_curves=14, seed=9, closed_frac=0.1, drop_view_frac=0.15, n_tracks=80)]


def identical(r, o):
    return (r.n_points == o.n_points and np.array_equal(r.obs_off, o.obs_off) and np.array_equal(r.obs_view, o.obs_view)
            and np.array_equal(r.obs_poly, o.obs_poly) and np.array_equal(r.obs_seg, o.obs_seg)
            and r.obs_xy.tobytes() == o.obs_xy.tobytes() and r.xyz.tobytes() == o.xyz.tobytes())


def set_seeds(sc, cands):
    """the seeds eg3d_match_polyline_sets / the reference's set loop sample: sets in order, views ascending, polylines ascending"""
    V = sc.n_views
    vs, pls, cs = [], [], []
    for s in range(cands.n_sets):
        for v in range(V):
            for k in range(int(cands.off[s * V + v]), int(cands.off[s * V + v + 1])):
                vs.append(v); pls.append(int(cands.polyline[k])); cs.append(s)
    view, pl, seg, xy, src = O.sample_seeds(sc, np.array(vs, np.int32), np.array(pls, np.uint32), 20.0)
    return SeedBatch(view, pl, seg, xy, np.array(cs, np.int32)[src])


@pytest.mark.parametrize("kw", SCENES)
def test_oracle_equals_the_reference_code_on_synthetic_scenes(kw):
    sc = syn.make_scene(**kw)
    rs, osc = R.RefScene(sc), O.OracleScene(sc)
    # --- a5: find_epipolar_correspondences, every seed (no undefined behaviour on this part of the path)
    seeds = syn.sample_seeds(O.sample_seeds, sc, per_view=40)
    off_r, hits_r = rs.epipolar_intersect(seeds)
    off_o, hits_o, _ = osc.epipolar_intersect(seeds)
    assert np.array_equal(off_r, off_o) and hits_r.tobytes() == hits_o.tobytes() and len(hits_r) > 5000
    # --- a8-a11, sweep form: one call of find_new_3d_points_from_compatible_polylines_starting_plgp_expandallviews per seed
    o = osc.match_seeds(seeds, n_threads=8)
    assert 0 < o.seed_ub.sum() < 0.1 * len(seeds)                       # the guarded case does occur
    r = rs.match_seeds(seeds)
    assert identical(r, o) and r.n_points > 500
    assert np.array_equal(r.seed, o.seed) and np.array_equal(r.chain_pos, o.chain_pos)
    assert identical(rs.match_seeds(seeds, n_threads=4), o)             # the per-seed entry from an OpenMP loop (bench.py's reference arm)
    # --- candidate-set form (pipelines 1-2): per seed, and the reference's own loop over sets / starting views / polylines
    #     (find_new_3d_points_from_compatible_polylines_expandallviews_parallel, its seed sampler included)
    cands = syn.curve_candidate_sets(sc, seed=kw["seed"])
    cs = set_seeds(sc, cands)
    assert identical(rs.match_seeds(cs, cands), osc.match_seeds(cs, cands, n_threads=8))
    r_sets, o_sets = rs.match_polyline_sets(cands), osc.match_polyline_sets(cands, n_threads=8)
    assert identical(r_sets, o_sets) and r_sets.n_points > 500
    # --- a6 + a12: plg_matching_from_refpoint (PLGEdgeManager seeding -> PLGPCM3ViewsPLGFollowing -> a8-a11), per SfM point
    #     the whole range goes through plg_matching_from_refpoints_parallel itself
    o3 = osc.match_refpoints(0, sc.n_tracks, n_threads=8)
    assert identical(rs.match_refpoints(0, sc.n_tracks), o3) and o3.n_points > 1000
    for t in range(0, sc.n_tracks, 7):
        assert identical(rs.match_refpoints(t, t + 1), osc.match_refpoints(t, t + 1, n_threads=1)), t
    # --- a13 + a14 on what the matching produced: filter_3d_points_close_2d_array, compute_inliers (gaussNewtonFiltering + view-count rule)
    keep_r, keep_o = rs.dedup_close_points(o), osc.dedup_close_points(o)
    assert np.array_equal(keep_r, keep_o) and 0 < keep_o.sum() < o.n_points
    xr, ir = rs.filter(o.xyz, o.obs_off, o.obs_view, o.obs_xy, 0)
    xo, io = osc.filter(o.xyz, o.obs_off, o.obs_view, o.obs_xy, 0)[:2]
    assert np.array_equal(ir, io) and xr.tobytes() == np.ascontiguousarray(xo, np.float32).tobytes() and 0 < io.sum()
    first = o.n_points // 3                                            # the view-count rule only applies from first_edgepoint on
    xr, ir = rs.filter(o.xyz, o.obs_off, o.obs_view, o.obs_xy, first)
    xo, io = osc.filter(o.xyz, o.obs_off, o.obs_view, o.obs_xy, first)[:2]
    assert np.array_equal(ir, io) and xr.tobytes() == np.ascontiguousarray(xo, np.float32).tobytes()


def test_oracle_equals_the_reference_code_on_the_packaged_dtu006_example():
    """Real edge maps -> polyline graphs (row f1), real cameras / tracks, LMedS fundamental matrices: pipeline 3 for the first 300
    SfM points and pipeline 2 for the first candidate sets, then the density limiter and the outlier filter."""
    from edgegraph3d_b200 import real_scene, pipeline as P
    real, _ = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    rs, osc = R.RefScene(real), O.OracleScene(real)
    # pipeline 2 first, as pipelines.cpp orders them: pipeline 3 records matched intervals in the shared PLGMatchesManager
    # (plg_matching_from_refpoints.cpp:75) which find_epipolar_correspondences would then filter by (polyline_matching.cpp:65);
    # candidate sets from SfM points (row f2), seeds of the first sets, those without undefined behaviour
    _, c2, _ = P.candidate_sets(real)
    cs = set_seeds(real, c2)
    cs = cs.take(np.arange(min(len(cs), 600)))
    r2, o2 = rs.match_seeds(cs, c2, n_threads=4), osc.match_seeds(cs, c2, n_threads=8)
    assert identical(r2, o2) and r2.n_points > 1000
    assert np.array_equal(rs.dedup_close_points(o2), osc.dedup_close_points(o2))
    xr, ir = rs.filter(o2.xyz, o2.obs_off, o2.obs_view, o2.obs_xy, 0)
    xo, io = osc.filter(o2.xyz, o2.obs_off, o2.obs_view, o2.obs_xy, 0)[:2]
    assert np.array_equal(ir, io) and xr.tobytes() == np.ascontiguousarray(xo, np.float32).tobytes() and 0 < io.sum() < len(io)

    # pipeline 3
    o3 = osc.match_refpoints(0, 400, n_threads=8)
    n_ub = int(o3.seed_ub.sum())
    assert 0 < n_ub < 0.1 * len(o3.seed_ub)               # ~5 % of the real seeds reach the guarded case
    r3 = rs.match_refpoints(0, 400)
    assert identical(r3, o3) and r3.n_points > 5000


@pytest.mark.skipif(not os.path.exists("/root/reference/example/dtu006/input.json"), reason="needs the reference tree (build container only)")
def test_openmvg_parser_of_the_reference_gives_the_same_cameras():
    """Row f3: the reference's own OpenMvgParser (compiled into libref_path.so) on the packaged input.json against openmvg_io."""
    from edgegraph3d_b200 import openmvg_io as io
    cams, w, h, npts = R.parse_openmvg("/root/reference/example/dtu006/input.json")
    d = io.load_sfm_data("/root/reference/example/dtu006/input.json")
    assert (w, h, npts) == (d["width"], d["height"], len(d["track_xyz"])) and cams.shape == d["cameras"].shape
    assert cams.tobytes() == np.ascontiguousarray(d["cameras"], np.float32).tobytes()
