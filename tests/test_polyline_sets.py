"""Row f2 (producer of pipeline 2's candidate sets, host C++ behind the C-ABI) against a second reading composed of the
numpy re-derivations of the reference's primitives (tests/test_oracle_primitives.py: 10 px polyline grid build and 3x3
lookup, point-to-polyline distance) plus the set / graph logic of polyline_matching_closeness_to_refpoints
(src/edgegraph3d/matching/polyline_matching/polyline_matcher.cpp:75-168) written out again here.  CPU only."""
import os
import numpy as np
import pytest
from edgegraph3d_b200 import lib as E, synthetic as syn, real_scene
from edgegraph3d_b200.scene import FlatScene
from tests.test_oracle_primitives import ref_build_grid, ref_grid_query, ref_compute_distancesq, f32

HERE = os.path.dirname(os.path.abspath(__file__))
FLT_MAX, FLT_MIN = np.finfo(np.float32).max, np.finfo(np.float32).tiny


def ref_polyline_sets(sc, find_within_dist=10.0, mult=3.0):
    V = sc.n_views
    grids = [ref_build_grid(sc, v, find_within_dist, sc.width, sc.height) for v in range(V)]
    dsq_max = f32(f32(find_within_dist) * f32(find_within_dist))
    node_of, nodes, adj, refpoints = {}, [], [], []
    for t in range(sc.n_tracks):
        o0, o1 = int(sc.track_off[t]), int(sc.track_off[t + 1])
        cams = [int(c) for c in sc.track_view[o0:o1]]
        res = []
        for cam in cams:
            last = max(k for k in range(o0, o1) if sc.track_view[k] == cam)      # get_2d_coordinates_of_point_on_image keeps the last match
            p = sc.track_xy[last]
            grid, gw, gh = grids[cam]
            near = []
            for pl in ref_grid_query(grid, gw, gh, find_within_dist, sc.width, sc.height, p):
                d, seg, proj = ref_compute_distancesq(sc.polyline(cam, pl), p)
                if d <= dsq_max:
                    near.append((pl, np.sqrt(f32(d))))
            res.append(near)
        if max((len(r) for r in res), default=0) != 1:
            continue
        pairs, dmin, dmax = set(), FLT_MAX, FLT_MIN
        for cam, r in zip(cams, res):
            if not r:
                continue
            pl, dist = r[0]
            dmin = dmin if dmin <= dist else dist
            dmax = dmax if dmax >= dist else dist
            pairs.add((cam, pl))
        if len(pairs) < len(cams) * 0.7 or dmin < f32(dmax / f32(mult)) or dmax > f32(dmin * f32(mult)) or len(pairs) < 2:
            continue
        ids = []
        for cp in sorted(pairs):
            if cp not in node_of:
                node_of[cp] = len(nodes)
                nodes.append(cp)
                adj.append(set())
            ids.append(node_of[cp])
        for i in range(len(ids)):
            for j in range(i + 1, len(ids)):
                adj[ids[i]].add(ids[j])
                adj[ids[j]].add(ids[i])
        refpoints.append(t)
    sets, seen = [], [False] * len(nodes)
    for s in range(len(nodes)):
        if seen[s]:
            continue
        comp, stack = [set() for _ in range(V)], [s]
        seen[s] = True
        while stack:
            c = stack.pop()
            comp[nodes[c][0]].add(nodes[c][1])
            for nb in sorted(adj[c]):
                if not seen[nb]:
                    seen[nb] = True
                    stack.append(nb)
        sets.append([sorted(x) for x in comp])
    return sets, refpoints


def assert_same(sc):
    cs, ref = E.polyline_sets_from_refpoints(sc)
    want_sets, want_ref = ref_polyline_sets(sc)
    assert ref.tolist() == want_ref
    assert cs.n_sets == len(want_sets)
    V = sc.n_views
    for i, s in enumerate(want_sets):
        for v in range(V):
            assert cs.polyline[cs.off[i * V + v]:cs.off[i * V + v + 1]].tolist() == s[v], (i, v)
    return cs, ref


@pytest.mark.parametrize("seed", [3, 8])
def test_product_equals_second_reading_on_synthetic_scenes(seed):
    sc = syn.make_scene(n_views=5, n_curves=14, seed=seed, closed_frac=0.15, n_tracks=120)
    cs, ref = assert_same(sc)
    assert cs.n_sets > 0 and len(ref) > 0
    # every set is usable as an eg3d_match_polyline_sets input: >= 2 (view, polyline) pairs, ids ascending and in range
    V = sc.n_views
    for i in range(cs.n_sets):
        assert cs.off[(i + 1) * V] - cs.off[i * V] >= 2
        for v in range(V):
            ids = cs.polyline[cs.off[i * V + v]:cs.off[i * V + v + 1]]
            assert (np.diff(ids.astype(np.int64)) > 0).all() and (ids < sc.n_polylines(v)).all()


def test_product_equals_second_reading_on_real_dtu006_views():
    """Real polyline graphs (row f1 on the packaged edge maps) of 3 views with the real SfM tracks restricted to them."""
    full, plgs = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    keep_v = [0, 1, 2]
    V = len(keep_v)
    from edgegraph3d_b200 import plg_build as PB
    arr = PB.scene_polyline_arrays([plgs[v] for v in keep_v])
    ko = np.isin(full.track_view, keep_v)
    cnt = np.add.reduceat(ko.astype(np.int64), full.track_off[:-1])
    cnt[np.diff(full.track_off) == 0] = 0
    kt = cnt >= 2
    off = np.concatenate([[0], np.cumsum(cnt[kt])])
    obs_keep = ko & np.repeat(kt, np.diff(full.track_off))
    sc = FlatScene(width=full.width, height=full.height, cameras=full.cameras[keep_v],
                   fundamental=full.fundamental.reshape(25, 25, 9)[np.ix_(keep_v, keep_v)], fundamental_valid=full.fundamental_valid[np.ix_(keep_v, keep_v)],
                   track_xyz=full.track_xyz[kt], track_off=off, track_view=full.track_view[obs_keep], track_xy=full.track_xy[obs_keep], **arr)
    assert sc.n_tracks > 500
    cs, ref = assert_same(sc)
    assert cs.n_sets > 5


def test_full_dtu006_counts_and_errors():
    import ctypes as C
    from edgegraph3d_b200 import _abi as A
    full, _ = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    cs, ref = E.polyline_sets_from_refpoints(full)
    assert cs.n_sets > 100 and len(ref) >= cs.n_sets and (np.diff(ref) > 0).all()
    # a (view, polyline) pair belongs to exactly one set (sets are connected components)
    V = full.n_views
    seen = set()
    for i in range(cs.n_sets):
        for v in range(V):
            for pl in cs.polyline[cs.off[i * V + v]:cs.off[i * V + v + 1]]:
                assert (v, int(pl)) not in seen
                seen.add((v, int(pl)))
    L = E.load()
    h = C.c_void_p()
    assert L.eg3d_polyline_sets_from_refpoints(None, 10.0, 3.0, C.byref(h)) == A.EG3D_ERR_INVALID_ARG
    no_tracks = FlatScene(width=full.width, height=full.height, cameras=full.cameras, fundamental=full.fundamental, fundamental_valid=full.fundamental_valid,
                          view_poly_off=full.view_poly_off, poly_vert_off=full.poly_vert_off, verts=full.verts, poly_start=full.poly_start, poly_end=full.poly_end)
    d = no_tracks.desc()
    assert L.eg3d_polyline_sets_from_refpoints(C.byref(d), 10.0, 3.0, C.byref(h)) == A.EG3D_ERR_INVALID_ARG
