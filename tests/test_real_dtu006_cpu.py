"""The packaged example on the CPU side: real edge maps -> polyline graphs (f1) -> candidate sets (f2) -> the ORACLE's
pipelines 2 and 3 (the GPU equals it point for point, tests/test_zz_dtu006_real.py), checked against invariants that follow
from the reference text alone, not from either implementation: every accepted point has observations in >= 3 distinct views
(triangulation.cpp:1035-1066 picks three views), every observation lies ON the polyline segment it names
(intersect_segment_line / first_plus_ratio_of_segment only produce points of the segment), and the point reprojects within
em_GaussNewton's acceptance bound (mean squared residual per coordinate < 9, triangulation.cpp:150-168)."""
import os
import numpy as np
from edgegraph3d_b200 import lib as E, pipeline as P, real_scene
from tests import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))


def check_invariants(sc, pts):
    assert pts.n_points > 1000
    lens = np.diff(pts.obs_off)
    assert lens.min() >= 3
    pid = np.repeat(np.arange(pts.n_points), lens)
    # at least three DISTINCT views per point.  (A view can appear twice in a point's list, with the identical observation: the
    # interval bookkeeping of expand_allpoints_to_other_view_using_plmap, triangulation.cpp:789-830, only skips the interval an
    # epipolar hit already covered when the scan lands exactly on its first index; 0.07 % of the points of this scene, in runs
    # along a chain.  Both implementations reproduce it; it is not treated as an error here.)
    key = pid.astype(np.int64) * sc.n_views + pts.obs_view
    distinct = np.bincount(np.unique(key) // sc.n_views, minlength=pts.n_points)
    assert distinct.min() >= 3
    assert (lens - distinct).sum() < 0.002 * len(key)
    # observation on its segment: distance to the segment (double arithmetic) below float rounding of ~1e-3 px
    g = sc.view_poly_off[pts.obs_view] + pts.obs_poly.astype(np.int64)
    nseg = sc.poly_vert_off[g + 1] - sc.poly_vert_off[g] - 1
    assert (pts.obs_seg.astype(np.int64) < nseg).all()
    a = sc.verts[sc.poly_vert_off[g] + pts.obs_seg.astype(np.int64)].astype(np.float64)
    b = sc.verts[sc.poly_vert_off[g] + pts.obs_seg.astype(np.int64) + 1].astype(np.float64)
    p = pts.obs_xy.astype(np.float64)
    ab = b - a
    t = np.clip(((p - a) * ab).sum(1) / np.maximum((ab * ab).sum(1), 1e-30), 0, 1)
    d = np.linalg.norm(p - (a + t[:, None] * ab), axis=1)
    assert d.max() < 2e-3, d.max()
    # reprojection within the acceptance bound of the solver that accepted the point
    Pm = sc.cameras.astype(np.float64).reshape(-1, 3, 4)[pts.obs_view]
    X = np.concatenate([pts.xyz.astype(np.float64)[pid], np.ones((len(pid), 1))], 1)
    h = np.einsum("nij,nj->ni", Pm, X)
    r2 = (((h[:, :2] / h[:, 2:3]) - p) ** 2).sum(1)
    mse = np.bincount(pid, r2, pts.n_points) / (2 * lens)
    assert mse.max() < 9.0 + 1e-3, mse.max()
    return float(np.median(mse))


def test_real_dtu006_oracle_results_satisfy_the_references_invariants():
    sc, _ = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    c2, ref = E.polyline_sets_from_refpoints(sc)
    dev = O.OracleDevice(sc, E.default_params(**P.REAL_DATA_CAPACITIES), n_threads=8)
    O.dlt_stats()
    p2 = dev.match_polyline_sets(c2)[0]
    calls, degenerate = O.dlt_stats()
    assert calls > 100000 and 0.05 < degenerate / calls < 0.25      # the get_min_max quirk is not a corner case
    p3 = dev.match_refpoints(0, 1500)[0]
    for pts in (p2, p3):
        med = check_invariants(sc, pts)
        assert med < 2.0           # typical accepted point: about a pixel


def test_filter_view_count_rule_second_reading():
    """compute_ray_stats + compute_inliers (src/edgegraph3d/filtering/outliers_filtering.cpp:14-64) written out again in numpy on top of
    the per-point Gauss-Newton verdicts (those are pinned bit for bit against cv2 in test_oracle_golden.py): the median
    track length is read off a histogram with the `>= count/2` integer rule, and edge points need more than
    max(3, median/2 - 1) observations.  Input: the real dtu006 tracks + the oracle's pipeline-2 points after the density limiter."""
    sc, _ = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    c2, _ = E.polyline_sets_from_refpoints(sc)
    dev = O.OracleDevice(sc, E.default_params(**P.REAL_DATA_CAPACITIES), n_threads=8)
    pts = dev.match_polyline_sets(c2)[0]
    keep = dev.dedup_close_points(pts)
    xyz, obs_off, obs_view, obs_xy = P.add_points_to_tracks(sc, pts, keep)
    first = sc.n_tracks
    for forced in (-1, 5):
        _, inl = dev.osc.filter(xyz, obs_off, obs_view, obs_xy, first, forced_min_filter=forced, n_threads=8)
        _, _, gn_ok = dev.osc.gn_triangulate(obs_off, obs_view, obs_xy, xyz, fp64=0, n_threads=8)      # GaussNewton of the filter, per point
        lens = np.diff(obs_off)
        hist = np.bincount(lens[gn_ok == 1] - 1, minlength=sc.n_views)                                   # point_rays_amount_distribution
        count = int((gn_ok == 1).sum())
        acc, median = 0, 0
        for median in range(sc.n_views):
            acc += int(hist[median])
            if acc >= count // 2:
                break
        intended = 3 if 3 >= median // 2 - 1 else median // 2 - 1
        if forced > -1:
            intended = forced
        want = gn_ok.astype(bool).copy()
        want[first:] &= lens[first:] > intended
        assert np.array_equal(inl.astype(bool), want)
        assert 0 < want[first:].sum() < (gn_ok[first:] == 1).sum() or forced == -1
