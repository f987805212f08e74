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


def ref_polyline_sets(sc, find_within_dist=10.0, mult=3.0, return_graph=False):
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
    if return_graph:
        return sets, refpoints, nodes, adj
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


# ------------------------------------------------------------------------------------------------------------------
# Pipeline 1's producer: the weighted compatibility graph of polyline_matching_similarity_graph (polyline_matcher.cpp:222-336)
# ------------------------------------------------------------------------------------------------------------------
def ref_similarity_graph(sc, find_within_dist=10.0):
    """Second reading: -> (nodes [(view, polyline)], edges [(a, b, w float32)] with a < b in the reference's order, file text)."""
    V, NT = sc.n_views, sc.n_tracks
    grids = [ref_build_grid(sc, v, find_within_dist, sc.width, sc.height) for v in range(V)]
    dsq_max = f32(f32(find_within_dist) * f32(find_within_dist))
    node_of, nodes, adj = {}, [], []
    close_refpoints = {}
    close_polylines = []
    visible = [set() for _ in range(V)]
    for t in range(NT):
        o0, o1 = int(sc.track_off[t]), int(sc.track_off[t + 1])
        cams = [int(c) for c in sc.track_view[o0:o1]]
        pairs = set()
        for cam in cams:
            visible[cam].add(t)
            last = max(k for k in range(o0, o1) if sc.track_view[k] == cam)
            p = sc.track_xy[last]
            grid, gw, gh = grids[cam]
            for pl in ref_grid_query(grid, gw, gh, find_within_dist, sc.width, sc.height, p):
                d, _, _ = ref_compute_distancesq(sc.polyline(cam, pl), p)
                if d <= dsq_max:
                    pairs.add((cam, pl))
        per_view = [set() for _ in range(V)]
        ids = []
        for cp in sorted(pairs):
            per_view[cp[0]].add(cp[1])
            close_refpoints.setdefault(cp, []).append(t)
            if cp not in node_of:
                node_of[cp] = len(nodes)
                nodes.append(cp)
                adj.append(set())
            ids.append(node_of[cp])
        for i in range(len(ids)):
            for j in range(i + 1, len(ids)):
                adj[ids[i]].add(ids[j])
                adj[ids[j]].add(ids[i])
        close_polylines.append(per_view)
    weights = []
    for per_view in close_polylines:                       # compute_refpoint_weight
        non_empty = sum(1 for s in per_view if s)
        total = sum(len(s) for s in per_view)
        weights.append(f32(0) if non_empty == 0 else f32(f32(non_empty) / f32(total)))
    edges, wadj = [], [dict() for _ in nodes]
    for n1 in range(len(nodes)):
        for n2 in sorted(adj[n1]):
            if not n1 < n2:
                continue
            (c1, p1), (c2, p2) = nodes[n1], nodes[n2]
            la = [t for t in close_refpoints[(c1, p1)] if t in visible[c2]]
            lb = [t for t in close_refpoints[(c2, p2)] if t in visible[c1]]
            inter = f32(0)
            for t in sorted(set(la) & set(lb)):
                inter = f32(inter + weights[t])
            if inter == 0:
                continue
            uni = f32(0)
            for t in sorted(set(la) | set(lb)):
                uni = f32(uni + weights[t])
            w = f32(inter / uni)
            if w > 0:
                edges.append((n1, n2, w))
                wadj[n1][n2] = w
                wadj[n2][n1] = w
    text = "p sp %d %d\n" % (len(nodes), sum(len(a) for a in wadj))
    for n1 in range(len(nodes)):
        for n2 in sorted(wadj[n1]):
            text += "a %d %d %s\n" % (n1 + 1, n2 + 1, "%g" % float(wadj[n1][n2]))     # ostream << float: 6 significant digits
    return nodes, edges, text


def modularity(g, com):
    w = g.edge_weight.astype(np.float64)
    a, b = com[g.edge_a], com[g.edge_b]
    ok = (a >= 0) & (b >= 0)
    mm, n = 2 * w.sum(), int(com.max()) + 1
    tot = np.bincount(a[ok], w[ok], n) + np.bincount(b[ok], w[ok], n)
    inn = np.bincount(a[ok & (a == b)], 2 * w[ok & (a == b)], n)
    return float((inn / mm - (tot / mm) ** 2).sum())


@pytest.mark.parametrize("seed", [3, 8])
def test_similarity_graph_equals_second_reading(seed):
    sc = syn.make_scene(n_views=5, n_curves=14, seed=seed, closed_frac=0.15, n_tracks=120)
    g = E.SimilarityGraph(sc)
    nodes, edges, text = ref_similarity_graph(sc)
    assert list(zip(g.node_view.tolist(), g.node_polyline.tolist())) == nodes and len(nodes) > 10
    assert list(zip(g.edge_a.tolist(), g.edge_b.tolist())) == [(a, b) for a, b, _ in edges] and len(edges) > 10
    assert g.edge_weight.tobytes() == np.array([w for _, _, w in edges], np.float32).tobytes()
    assert g.dimacs == text                                                       # the file the reference hands to Grappolo, byte for byte
    com, q = g.communities()
    com2, q2 = g.communities()
    assert np.array_equal(com, com2) and q == q2                                  # deterministic
    assert abs(q - modularity(g, com)) < 1e-9 and q > 0.3
    deg = np.bincount(np.concatenate([g.edge_a, g.edge_b]), minlength=len(nodes))
    assert ((com < 0) == (deg == 0)).all()                                        # nodes without an edge belong to no community
    assert (com[g.edge_a] >= 0).all()
    cs = g.candidate_sets(com)
    assert cs.n_sets == com.max() + 1
    V = sc.n_views
    for c in range(cs.n_sets):                                                    # compute_polyline_matches_from_nodes_component_ids
        for v in range(V):
            want = sorted(int(p) for (vv, p), cc in zip(nodes, com) if cc == c and vv == v)
            assert cs.polyline[cs.off[c * V + v]:cs.off[c * V + v + 1]].tolist() == want


def test_communities_against_the_references_own_grappolo(tmp_path):
    """oracle/_ref/libgrappolo_ref.so is the reference's vendored Louvain compiled from its own sources (oracle/Makefile,
    `make ref`).  It cannot serve as a bit-level oracle — it writes nothing on one thread and its partition changes with
    the thread count and from run to run — so the product's deterministic Louvain is held to the quantity both optimise:
    its modularity on the real dtu006 compatibility graph must not be below the reference's."""
    import ctypes as C
    so = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libgrappolo_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libgrappolo_ref.so not built (needs /root/reference: make -C oracle ref)")
    L = C.CDLL(so)
    L.eg3d_ref_grappolo_communities.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    full, _ = real_scene.dtu006_scene(os.path.join(HERE, "golden"))
    g = E.SimilarityGraph(full)
    assert len(g.node_view) > 20000 and len(g.edge_a) > 300000
    gf = tmp_path / "compatibility_graph.txt"
    gf.write_text(g.dimacs)
    one = tmp_path / "one_thread.txt"
    L.eg3d_ref_grappolo_communities(str(gf).encode(), str(one).encode(), 1)
    assert not one.exists()                                   # driverForGraphClustering_edited.cpp:52-58: `if (nT <= 1) return 0;`
    out = tmp_path / "communities.txt"
    L.eg3d_ref_grappolo_communities(str(gf).encode(), str(out).encode(), 4)
    ref = np.loadtxt(out, dtype=np.int64)
    assert len(ref) == len(g.node_view)                      # the reference's parser accepts the file as written
    com, q = g.communities()
    assert ((ref < 0) == (com < 0)).all()                    # the same nodes (those without an edge) are left out by both
    q_ref = modularity(g, ref)
    assert q >= q_ref - 1e-3 and q > 0.9, (q, q_ref)


def test_component_order_equals_the_references_own_get_components():
    """GraphAdjacencySetUndirectedNoType::get_components (graph_adjacency_set_undirected_no_type.cpp:44-69) decides the ORDER of
    pipeline 2's candidate sets; here the reference's compiled class (oracle/_ref/libref_graph.so) gets the polyline-match
    graph of a synthetic scene and must number the components as the product does."""
    import ctypes as C
    so = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_graph.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_graph.so not built (needs /root/reference: make -C oracle ref)")
    L = C.CDLL(so)
    L.eg3d_ref_graph_new.restype = C.c_void_p
    L.eg3d_ref_graph_new.argtypes = [C.c_ulong]
    L.eg3d_ref_graph_free.argtypes = [C.c_void_p]
    L.eg3d_ref_graph_add_edge.argtypes = [C.c_void_p, C.c_ulong, C.c_ulong]
    L.eg3d_ref_graph_components.argtypes = [C.c_void_p, C.POINTER(C.c_ulong)]
    sc = syn.make_scene(n_views=5, n_curves=14, seed=8, closed_frac=0.15, n_tracks=120)
    cs, _ = E.polyline_sets_from_refpoints(sc)
    V = sc.n_views
    _, _, order, adj = ref_polyline_sets(sc, return_graph=True)          # nodes in first-seen order + adjacency of the second reading
    edges = [(a, b) for a in range(len(order)) for b in sorted(adj[a]) if a < b]
    g = L.eg3d_ref_graph_new(len(order))
    try:
        for a, b in edges:
            L.eg3d_ref_graph_add_edge(g, a, b)
        comp = (C.c_ulong * len(order))()
        L.eg3d_ref_graph_components(g, comp)
    finally:
        L.eg3d_ref_graph_free(g)
    # every node's component number (the reference's own numbering) = index of the product's candidate set that holds it
    covered = 0
    for k, cp in enumerate(order):
        i = int(comp[k])
        ids_v = cs.polyline[cs.off[i * V + cp[0]]:cs.off[i * V + cp[0] + 1]].tolist()
        assert cp[1] in ids_v, (k, cp, i)
        covered += 1
    assert covered >= 6 and len(set(comp)) == cs.n_sets
