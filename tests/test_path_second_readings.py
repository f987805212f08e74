"""Second readings one level above the primitives, CPU only.

Row a6 (pipeline-3 seeding: PLGEdgeManager::detect_nearby_intersections_and_correspondences_plgp,
src/edgegraph3d/edge_managers/plg_edge_manager.cpp:191-300) and row a5 (find_epipolar_correspondences,
polyline_matching.cpp:45-73): the numpy re-derivations of the primitives (grids, point-to-polyline distance, segment/line
intersection, double-intermediate distance) composed with REAL cv2.computeCorrespondEpilines calls, against the oracle's
seeds and per-view hit lists, bit for bit.  Rows a8 / a10: rules that the reference text fixes and that are visible in the
output (the view triple a hypothesis is built from; 10 px steps on the driving view), checked on the oracle's result from
independently computed inputs.  The GPU path is compared with the same oracle in test_gpu_parity.py."""
import os
import numpy as np
import pytest
from edgegraph3d_b200 import synthetic as syn, real_scene
from tests import oracle_lib as O
from tests.test_oracle_primitives import (f32, ref_build_grid, ref_grid_query, ref_compute_distancesq, ref_intersect_segment_line,
                                          ref_squared_2d_distance)

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_refpoint_hits(sc, tb, te, start_dist=10.0, mult=3.0):
    """-> list over seeds of [V lists of (polyline, segment, x, y)] in the order plg_matching_from_refpoint consumes them
    (plg_matching_from_refpoints.cpp:64-81 -> plgpcm_3views_plg_following.cpp:40-50 scatters to a V-vector)."""
    cv2 = pytest.importorskip("cv2")
    V = sc.n_views
    cell = start_dist * mult                                              # plg_edge_manager.cpp:73-74: 30 px grid
    grids = {}
    start_dsq, corr_dsq = f32(f32(start_dist) * f32(start_dist)), f32(f32(cell) * f32(cell))
    F = sc.fundamental.reshape(V, V, 3, 3)
    out = []
    for t in range(tb, te):
        o0, o1 = int(sc.track_off[t]), int(sc.track_off[t + 1])
        cams = [int(c) for c in sc.track_view[o0:o1]]
        obs = [sc.track_xy[k] for k in range(o0, o1)]

        def coords_on(cam):                                               # get_2d_coordinates_of_point_on_image: the last match
            return obs[max(i for i, c in enumerate(cams) if c == cam)]
        pcps, snis = [], []
        for cam in cams:                                                  # :272-288
            if cam not in grids:
                grids[cam] = ref_build_grid(sc, cam, cell, sc.width, sc.height)
            grid, gw, gh = grids[cam]
            p = coords_on(cam)
            cur_pcps, cur_sni = [], []
            for pl in ref_grid_query(grid, gw, gh, cell, sc.width, sc.height, p):
                d, seg, proj = ref_compute_distancesq(sc.polyline(cam, pl), p)
                if d <= start_dsq:
                    cur_pcps.append(pl)
                    cur_sni.append((pl, seg, proj))
                elif d <= corr_dsq:
                    cur_pcps.append(pl)
            pcps.append(cur_pcps)
            snis.append(cur_sni)
        for i, start in enumerate(cams):                                  # :290-297
            init = coords_on(start)
            for (pl, seg, proj) in snis[i]:                               # :246-259
                radius = f32(np.sqrt(ref_squared_2d_distance(init, proj)) * f32(mult))
                rsq = f32(radius * radius)
                rows = [[] for _ in range(V)]
                for j, cur in enumerate(cams):                            # :208-243
                    if cur == start:
                        rows[cur] = [(pl, seg, f32(proj[0]), f32(proj[1]))]   # a later duplicate of the same camera overwrites (scatter)
                        continue
                    if not sc.fundamental_valid[start, cur]:
                        rows[cur] = []
                        continue
                    line = cv2.computeCorrespondEpilines(np.array([[proj]], np.float32), 1, F[start, cur]).reshape(3)
                    hits = []
                    for cpl in pcps[j]:                                   # :191-205
                        pc = sc.polyline(cur, cpl)
                        for k in range(1, len(pc)):
                            found, q = ref_intersect_segment_line((pc[k][0], pc[k][1], pc[k - 1][0], pc[k - 1][1]), line)
                            if found and ref_squared_2d_distance(obs[j], q) <= rsq:
                                hits.append((cpl, k - 1, q[0], q[1]))
                    rows[cur] = hits
                out.append(rows)
    return out


def assert_same(sc, tb, te):
    off, hits, V = O.OracleScene(sc).refpoint_hits(tb, te)
    want = ref_refpoint_hits(sc, tb, te)
    assert len(off) == len(want) * V + 1
    n_hits = 0
    for s, rows in enumerate(want):
        for v in range(V):
            got = hits[off[s * V + v]:off[s * V + v + 1]]
            assert len(got) == len(rows[v]), (s, v)
            for g, w in zip(got, rows[v]):
                assert (int(g["polyline"]), int(g["segment"])) == (int(w[0]), int(w[1])), (s, v)
                assert np.float32(g["x"]).tobytes() == np.float32(w[2]).tobytes() and np.float32(g["y"]).tobytes() == np.float32(w[3]).tobytes(), (s, v)
            n_hits += len(got)
    return len(want), n_hits


def test_refpoint_seeding_second_reading_synthetic():
    sc = syn.make_scene(n_views=5, n_curves=14, seed=8, closed_frac=0.15, n_tracks=80)
    n_seeds, n_hits = assert_same(sc, 0, sc.n_tracks)
    assert n_seeds > 50 and n_hits > 200


def test_refpoint_seeding_second_reading_real_dtu006():
    """Real polyline graphs, real tracks, LMedS fundamental matrices (some pairs invalid): the first 60 SfM points."""
    sc, _ = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    n_seeds, n_hits = assert_same(sc, 0, 60)
    assert n_seeds > 100 and n_hits > 300


# ------------------------------------------------------------------------------------------------------------------
# Row a5: find_epipolar_correspondences (src/edgegraph3d/matching/plg_matching/polyline_matching.cpp:45-73)
# ------------------------------------------------------------------------------------------------------------------
def ref_epipolar_correspondences(sc, seeds, cands=None):
    """Per seed, V lists of (polyline, segment, x, y): every other view with a valid F, the candidate polylines of that view
    (or all of them in the sweep form) in ascending id, polyline::intersect_line segment by segment (polyline_graph_2d.cpp:312-327)."""
    cv2 = pytest.importorskip("cv2")
    V = sc.n_views
    F = sc.fundamental.reshape(V, V, 3, 3)
    out = []
    for s in range(len(seeds)):
        start = int(seeds.view[s])
        rows = [[] for _ in range(V)]
        for cur in range(V):
            if cur == start:                                              # :54-55: the starting view holds the seed itself
                rows[cur] = [(int(seeds.polyline[s]), int(seeds.segment[s]), seeds.xy[s][0], seeds.xy[s][1])]
                continue
            if not sc.fundamental_valid[start, cur]:
                continue
            line = cv2.computeCorrespondEpilines(seeds.xy[s].reshape(1, 1, 2), 1, F[start, cur]).reshape(3)
            if cands is None:
                pls = range(sc.n_polylines(cur))
            else:
                k = int(seeds.cand_set[s]) * V + cur
                pls = cands.polyline[cands.off[k]:cands.off[k + 1]].tolist()
            for pl in pls:
                pc = sc.polyline(cur, pl)
                for k in range(1, len(pc)):
                    found, q = ref_intersect_segment_line((pc[k][0], pc[k][1], pc[k - 1][0], pc[k - 1][1]), line)
                    if found:
                        rows[cur].append((pl, k - 1, q[0], q[1]))
        out.append(rows)
    return out


@pytest.mark.parametrize("mode", ["sweep", "candidates"])
def test_epipolar_correspondences_second_reading(mode):
    sc = syn.make_scene(n_views=5, n_curves=12, seed=5, closed_frac=0.2, drop_view_frac=0.1)
    cands = syn.curve_candidate_sets(sc, seed=5) if mode == "candidates" else None
    seeds = syn.sample_seeds(O.sample_seeds, sc, per_view=6, cand_set=0 if cands is not None else None)
    off, hits, V = O.OracleScene(sc).epipolar_intersect(seeds, cands)
    want = ref_epipolar_correspondences(sc, seeds, cands)
    total = 0
    for s, rows in enumerate(want):
        for v in range(V):
            got = hits[off[s * V + v]:off[s * V + v + 1]]
            assert [(int(g["polyline"]), int(g["segment"])) for g in got] == [(int(w[0]), int(w[1])) for w in rows[v]], (s, v)
            assert np.array([[g["x"], g["y"]] for g in got], np.float32).tobytes() == np.array([[w[2], w[3]] for w in rows[v]], np.float32).tobytes()
            total += len(got)
    assert total > 50


# ------------------------------------------------------------------------------------------------------------------
# Row a8: the view triple of compute_3D_point_multiple_views_plg_following_expandallviews_vector (triangulation.cpp:1035-1066)
# ------------------------------------------------------------------------------------------------------------------
def test_view_triple_selection_is_visible_in_the_output():
    """The three views a hypothesis is built from are (first view with hits, the starting view — or, if that is the first or
    the last one, the middle non-empty view —, last view with hits).  Every chain point that came out of the 3-view stage
    lists them first, in that order (view expansion only appends); points appended later by all-view following
    (follow_direction_vector_start/end) carry whichever views could be followed.  Checked on the oracle's output from hit
    lists computed independently: every accepted seed's chain holds at least two such points, and they are the large majority."""
    sc = syn.make_scene(n_views=7, n_curves=14, seed=9, closed_frac=0.1, drop_view_frac=0.15)
    seeds = syn.sample_seeds(O.sample_seeds, sc, per_view=40)
    osc = O.OracleScene(sc)
    off, hits, V = osc.epipolar_intersect(seeds)
    pts = osc.match_seeds(seeds)
    assert pts.n_points > 100
    with_triple = {}
    for i in range(pts.n_points):
        s = int(pts.seed[i])
        nonempty = [v for v in range(V) if off[s * V + v + 1] > off[s * V + v]]     # the starting view holds the seed itself
        assert len(nonempty) >= 3
        start = int(seeds.view[s])
        want = [nonempty[0], nonempty[len(nonempty) // 2] if start in (nonempty[0], nonempty[-1]) else start, nonempty[-1]]
        views = pts.obs_view[pts.obs_off[i]:pts.obs_off[i + 1]].tolist()
        with_triple[s] = with_triple.get(s, 0) + (views[:3] == want)
    # compatible_new_plg_point needs >= 2 points in a direction (plg_matching.cpp:1276-1287), all of them 3-view points
    assert all(n >= 2 for n in with_triple.values()), with_triple
    assert sum(with_triple.values()) > 0.8 * pts.n_points        # points appended by later all-view following are the exception


def test_following_steps_ten_pixels_on_the_driving_view():
    """PLG following advances 10 px (Euclidean, FOLLOW_FIRST_IMAGE_DISTANCE, plg_matching.hpp:39) along the polyline of the
    first selected view and finds the other views by epipolar intersection (plg_matching.cpp:51-132): consecutive chain points
    of the 3-view stage are therefore 10 px apart on that view (linear interpolation on the crossing segment), on one polyline."""
    sc = syn.make_scene(n_views=7, n_curves=14, seed=9, closed_frac=0.1, drop_view_frac=0.15)
    seeds = syn.sample_seeds(O.sample_seeds, sc, per_view=40)
    pts = O.OracleScene(sc).match_seeds(seeds)
    d, same_pl = [], []
    for i in range(pts.n_points - 1):
        a, b = int(pts.obs_off[i]), int(pts.obs_off[i + 1])
        if pts.seed[i] != pts.seed[i + 1] or pts.obs_view[a:a + 3].tolist() != pts.obs_view[b:b + 3].tolist():
            continue
        d.append(float(np.linalg.norm(pts.obs_xy[a].astype(np.float64) - pts.obs_xy[b].astype(np.float64))))
        same_pl.append(pts.obs_poly[a] == pts.obs_poly[b])
    d = np.array(d)
    assert len(d) > 500
    ok = np.abs(d - 10.0) < 0.05
    assert ok.mean() > 0.97, ok.mean()               # the rest: joints between the 3-view stage and later all-view following
    assert np.mean(same_pl) > 0.99


def test_view_expansion_appends_views_in_ascending_order():
    """expand_point_to_other_views_expandallviews_vector (triangulation.cpp:960-973) visits the views that are not in the
    triple in ascending order and update_new_3dpoint_plgp_matches appends: after the first three entries the view list of a
    point of the 3-view stage never decreases, and no point repeats one of its first three views later.  Synthetic scene and the real dtu006 graphs."""
    import os
    from edgegraph3d_b200 import lib as E, pipeline as P
    sc = syn.make_scene(n_views=7, n_curves=14, seed=9, closed_frac=0.1, drop_view_frac=0.15)
    sets = [O.OracleScene(sc).match_seeds(syn.sample_seeds(O.sample_seeds, sc, per_view=40))]
    real, _ = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    sets.append(O.OracleDevice(real, E.default_params(**P.REAL_DATA_CAPACITIES), n_threads=8).match_refpoints(0, 800)[0])
    for pts in sets:
        assert pts.n_points > 500
        lens = np.diff(pts.obs_off)
        pos = np.arange(pts.n_obs) - np.repeat(pts.obs_off[:-1], lens)                   # index of each observation in its point's list
        v = pts.obs_view.astype(np.int64)
        later = pos >= 4
        dec = v[later] < v[np.where(later)[0] - 1]                                        # a decrease after entry 3
        # points PREPENDED / APPENDED to the chain by later all-view following start with more than three views in following
        # order (driving view first); they are rare (22 of 20 143 on the real sample in the default, OpenCV-faithful DLT mode; 5 of
        # 20 311 with dlt_wellposed = 1) and the only exception
        assert dec.sum() <= 0.002 * pts.n_points, int(dec.sum())
        first3 = np.stack([v[pts.obs_off[:-1] + k] for k in range(3)], 1)                 # [n_points, 3]
        rest = pos >= 3
        owner = np.repeat(np.arange(pts.n_points), lens)[rest]
        # (repeats exist only as the duplicated observations described in tests/test_real_dtu006_cpu.py)
        assert (v[rest][:, None] == first3[owner]).any(1).sum() <= 0.002 * pts.n_obs


def test_three_view_points_are_epipolar_consistent():
    """`compatible` (plg_matching.cpp:51-132) builds every followed point from a 10 px step on the driving view a and the
    intersections of a's epipolar lines (F[a][b], F[a][c]) with the polylines of b and c; the central hypothesis is built
    from the hits of the SEED's epipolar lines (polyline_matching.cpp:45-73).  So every 3-view-stage point of the output lies,
    in views b and c, on the real cv2 epipolar line of its own observation in a — or, for the central point, of the seed."""
    cv2 = pytest.importorskip("cv2")
    sc = syn.make_scene(n_views=7, n_curves=14, seed=9, closed_frac=0.1, drop_view_frac=0.15)
    seeds = syn.sample_seeds(O.sample_seeds, sc, per_view=40)
    osc = O.OracleScene(sc)
    off, hits, V = osc.epipolar_intersect(seeds)
    pts = osc.match_seeds(seeds)
    F = sc.fundamental.reshape(V, V, 3, 3)

    def dist(x, src, a, b):            # distance of x (view b) to the epipolar line of src (view a); the line is normalised
        l = cv2.computeCorrespondEpilines(src.reshape(1, 1, 2), 1, F[a, b]).reshape(3).astype(np.float64)
        return abs(l[0] * float(x[0]) + l[1] * float(x[1]) + l[2])

    followed = central = 0
    for i in range(pts.n_points):
        s, o = int(pts.seed[i]), int(pts.obs_off[i])
        nonempty = [v for v in range(V) if off[s * V + v + 1] > off[s * V + v]]
        start = int(seeds.view[s])
        a, b, c = nonempty[0], nonempty[len(nonempty) // 2] if start in (nonempty[0], nonempty[-1]) else start, nonempty[-1]
        if pts.obs_view[o:o + 3].tolist() != [a, b, c]:
            continue                                           # appended later by all-view following
        xa, xb, xc = pts.obs_xy[o], pts.obs_xy[o + 1], pts.obs_xy[o + 2]
        if max(dist(xb, xa, a, b), dist(xc, xa, a, c)) < 0.01:
            followed += 1
            continue
        sx = seeds.xy[s]
        d = max(np.abs(x - sx).max() if vw == start else dist(x, sx, start, vw) for vw, x in ((a, xa), (b, xb), (c, xc)))
        assert d < 0.01, (i, s, d)
        central += 1
    assert followed > 500 and central > 10


def test_all_view_followed_points_are_epipolar_consistent_with_their_driving_view():
    """The all-view `compatible` (plg_matching.cpp:633-759) steps 10 px on ONE of the point's views (the new list starts with
    it) and finds every other view by intersecting that observation's epipolar line, so a point appended to a chain by
    all-view following has its next observations on the real cv2 epipolar lines of its FIRST observation (views that view
    expansion appends afterwards need not be)."""
    cv2 = pytest.importorskip("cv2")
    sc = syn.make_scene(n_views=7, n_curves=14, seed=9, closed_frac=0.1, drop_view_frac=0.15)
    seeds = syn.sample_seeds(O.sample_seeds, sc, per_view=40)
    osc = O.OracleScene(sc)
    off, hits, V = osc.epipolar_intersect(seeds)
    pts = osc.match_seeds(seeds)
    F = sc.fundamental.reshape(V, V, 3, 3)
    n = full = 0
    for i in range(pts.n_points):
        s, o, e = int(pts.seed[i]), int(pts.obs_off[i]), int(pts.obs_off[i + 1])
        nonempty = [v for v in range(V) if off[s * V + v + 1] > off[s * V + v]]
        start = int(seeds.view[s])
        triple = [nonempty[0], nonempty[len(nonempty) // 2] if start in (nonempty[0], nonempty[-1]) else start, nonempty[-1]]
        views = pts.obs_view[o:e].tolist()
        if views[:3] == triple:
            continue                                                  # 3-view stage: covered above
        ok = []
        for k in range(1, len(views)):
            l = cv2.computeCorrespondEpilines(pts.obs_xy[o].reshape(1, 1, 2), 1, F[views[0], views[k]]).reshape(3).astype(np.float64)
            ok.append(abs(l[0] * float(pts.obs_xy[o + k][0]) + l[1] * float(pts.obs_xy[o + k][1]) + l[2]) < 0.01)
        assert ok[0] and ok[1], (i, views, ok)                        # a followed point needs >= 3 views: the first two others came from following
        n += 1
        full += all(ok)
    assert n > 100 and full > 0.8 * n
