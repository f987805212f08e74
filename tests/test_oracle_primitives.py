"""Second reading of the reference's geometric primitives: numpy float32 re-derivations written directly from the
reference text (file:line in every docstring), independent of oracle/*.cpp, compared BIT FOR BIT with the oracle's
exported helpers on random and adversarial inputs.  The reference ships no tests; this guards the oracle's
control-flow primitives against transcription mistakes the cv2 goldens cannot see.  CPU only."""
import ctypes as C
import numpy as np
import pytest
from edgegraph3d_b200 import synthetic as syn, _abi as A
from tests import oracle_lib as O

f32 = np.float32


def ref_squared_2d_distance(a, b):
    """geometric_utilities.cpp:555-557: pow(float - float, 2) is a double power, the sum is double, the return float."""
    dx, dy = f32(a[0] - b[0]), f32(a[1] - b[1])
    return f32(np.float64(dx) ** 2 + np.float64(dy) ** 2)


def ref_distance_point_line(p, l):
    """geometric_utilities.cpp:997-1009: den = a x + b y + c (float), den *= den, den / (a a + b b), sqrt."""
    den = f32(f32(f32(l[0] * p[0]) + f32(l[1] * p[1])) + l[2])
    den = f32(den * den)
    return f32(np.sqrt(f32(den / f32(f32(l[0] * l[0]) + f32(l[1] * l[1])))))


def ref_anglecos(segm, l):
    """geometric_utilities.cpp:579-581, 604-618: signed cosine between (x2-x1, y2-y1) and (1, -a/b) (or (0,1) if b == 0)."""
    a = (f32(segm[2] - segm[0]), f32(segm[3] - segm[1]))
    b = (f32(0.0), f32(1.0)) if l[1] == 0 else (f32(1.0), f32(-l[0] / l[1]))
    dot = lambda u, v: f32(f32(u[0] * v[0]) + f32(u[1] * v[1]))
    with np.errstate(all="ignore"):
        return f32(dot(a, b) / f32(np.sqrt(f32(dot(a, a) * dot(b, b)))))


def ref_intersect_segment_line(segm, l):
    """geometric_utilities.cpp:272-312 -> (found, intersection)."""
    d = (f32(segm[2] - segm[0]), f32(segm[3] - segm[1]))
    num = f32(f32(f32(l[0] * segm[0]) + f32(l[1] * segm[1])) + l[2])
    den = f32(f32(l[0] * d[0]) + f32(l[1] * d[1]))
    if den != 0:
        t = f32(-num / den)
        if t >= 0 and t <= 1:
            return True, (f32(segm[0] + f32(t * d[0])), f32(segm[1] + f32(t * d[1])))
    return False, (f32(0), f32(0))


def ref_intersect_nqp(segm, l, max_cos, max_dist):
    """geometric_utilities.cpp:365-430 -> (intersection_found, quasiparallel_within_distance, intersection)."""
    d = (f32(segm[2] - segm[0]), f32(segm[3] - segm[1]))
    num = f32(f32(f32(l[0] * segm[0]) + f32(l[1] * segm[1])) + l[2])
    den = f32(f32(l[0] * d[0]) + f32(l[1] * d[1]))
    found, qp, inter = False, False, (f32(0), f32(0))
    if den != 0:
        t = f32(-num / den)
        if t >= 0 and t <= 1:
            inter = (f32(segm[0] + f32(t * d[0])), f32(segm[1] + f32(t * d[1])))
            found = True
        if ref_anglecos(segm, l) > max_cos:
            if t < 0:
                dist = ref_distance_point_line((segm[0], segm[1]), l)
            elif t > 1:
                dist = ref_distance_point_line((segm[2], segm[3]), l)
            else:
                dist = f32(0)
            if dist <= max_dist:
                qp = True
    else:
        if ref_distance_point_line((segm[0], segm[1]), l) <= max_dist:
            qp = True
    return found, qp, inter


def ref_lerp(a, b, r):
    """first_plus_ratio_of_segment, geometric_utilities.cpp:1370-1372."""
    return (f32(a[0] + f32(r * f32(b[0] - a[0]))), f32(a[1] + f32(r * f32(b[1] - a[1]))))


def ref_dist(a, b):
    return f32(np.sqrt(ref_squared_2d_distance(a, b)))          # compute_2d_distance, :571-573


def ref_next_by_distance(pc, seg, c, towards_end, distance):
    """polyline::next_pl_point_by_distance, polyline_graph_2d.cpp:391-447 -> (segment_index, coords, reached_extreme)."""
    n = len(pc)
    distance = f32(distance)
    if not towards_end:
        cur = ref_dist(pc[seg], c)
        if cur >= distance:
            return seg, ref_lerp(c, pc[seg], f32(distance / cur)), False
        i = seg
        prev = cur
        while i > 0:
            prev = cur
            cur = ref_dist(pc[i - 1], c)
            if cur >= distance:
                break
            i -= 1
        if i == 0:
            return 0, (pc[0][0], pc[0][1]), True
        return i - 1, ref_lerp(pc[i], pc[i - 1], f32(f32(distance - prev) / f32(cur - prev))), False
    if seg >= n - 1:
        return n - 2, (pc[n - 1][0], pc[n - 1][1]), True
    cur = ref_dist(pc[seg + 1], c)
    if cur >= distance:
        return seg, ref_lerp(c, pc[seg + 1], f32(distance / cur)), False
    i = seg + 1
    prev = cur
    while i < n - 1:
        prev = cur
        cur = ref_dist(pc[i + 1], c)
        if cur >= distance:
            break
        i += 1
    if i == n - 1:
        return n - 2, (pc[n - 1][0], pc[n - 1][1]), True
    return i, ref_lerp(pc[i], pc[i + 1], f32(f32(distance - prev) / f32(cur - prev))), False


def ref_next_by_line(pc, seg, c, towards_end, l, max_cos, max_dist):
    """polyline::next_pl_point_by_line_intersection, polyline_graph_2d.cpp:579-664 -> (found, segment_index, coords)."""
    n = len(pc)
    first_to = pc[seg + 1] if towards_end else pc[seg]
    f, q, p = ref_intersect_nqp((c[0], c[1], first_to[0], first_to[1]), l, max_cos, max_dist)
    if q:
        return False, seg, c
    if f:
        return True, seg, p
    rng = range(seg + 1, n - 1) if towards_end else range(seg, 0, -1)
    for i in rng:
        a, b = pc[i], (pc[i + 1] if towards_end else pc[i - 1])
        f, q, p = ref_intersect_nqp((a[0], a[1], b[0], b[1]), l, max_cos, max_dist)
        if q:
            return False, seg, c
        if f:
            return True, (i if towards_end else i - 1), p
    return False, seg, c


def _rand_lines(rng, n):
    ang = rng.uniform(0, 2 * np.pi, n)
    a, b = np.cos(ang).astype(f32), np.sin(ang).astype(f32)
    c = (-(a * rng.uniform(0, 640, n) + b * rng.uniform(0, 480, n))).astype(f32)
    return np.stack([a, b, c], 1)


def test_segment_line_primitives_bit_exact():
    L = O.lib()
    L.eg3d_oracle_intersect_segment_line.restype = C.c_int
    L.eg3d_oracle_intersect_segment_line_nqp.restype = C.c_int
    rng = np.random.default_rng(5)
    n = 4000
    segs = rng.uniform(0, 640, (n, 4)).astype(f32)
    segs[::7, 2:] = segs[::7, :2] + rng.uniform(-3, 3, (len(segs[::7]), 2)).astype(f32)      # short segments
    lines = _rand_lines(rng, n)
    lines[::11, 1] = 0                                                                       # vertical lines (b == 0 branch)
    # quasi-parallel cases: line almost along the segment, passing nearby
    for k in range(0, n, 5):
        d = segs[k, 2:] - segs[k, :2]
        nn = np.array([-d[1], d[0]], f32) / max(float(np.hypot(*d)), 1e-6)
        nn = (nn + rng.normal(0, 0.05, 2)).astype(f32)
        off = f32(rng.uniform(-8, 8))
        lines[k] = [nn[0], nn[1], -(nn[0] * segs[k, 0] + nn[1] * segs[k, 1]) + off]
    n_found = n_qp = 0
    out = np.zeros(2, f32)
    for k in range(n):
        s, l = segs[k].copy(), lines[k].copy()
        f = L.eg3d_oracle_intersect_segment_line(A.ptr(s, A.c_f32p), A.ptr(l, A.c_f32p), A.ptr(out, A.c_f32p))
        rf, rp = ref_intersect_segment_line(s, l)
        assert bool(f) == rf and (not rf or (out[0] == rp[0] and out[1] == rp[1])), k
        r = L.eg3d_oracle_intersect_segment_line_nqp(A.ptr(s, A.c_f32p), A.ptr(l, A.c_f32p), f32(0.965), f32(5.0), A.ptr(out, A.c_f32p))
        rf, rq, rp = ref_intersect_nqp(s, l, f32(0.965), f32(5.0))
        assert r == (1 if rf else 0) | (2 if rq else 0), (k, r, rf, rq)
        assert not rf or (out[0] == rp[0] and out[1] == rp[1]), k
        n_found += rf; n_qp += rq
        a, b = s[:2], s[2:]
        assert L.eg3d_oracle_squared_2d_distance(a[0], a[1], b[0], b[1]) == ref_squared_2d_distance(a, b)
    assert n_found > 300 and n_qp > 100          # both branches are exercised


def test_polyline_walkers_bit_exact():
    L = O.lib()
    L.eg3d_oracle_walk.restype = C.c_int
    sc = syn.make_scene(n_views=3, n_curves=14, seed=9, closed_frac=0.2)
    osc = O.OracleScene(sc)
    rng = np.random.default_rng(2)
    out_seg = C.c_uint32(); out_xy = np.zeros(2, f32)
    n_reached = n_found = checked = 0
    for view in range(3):
        for pl in range(sc.n_polylines(view)):
            pc = sc.polyline(view, pl)
            if len(pc) < 2:
                continue
            g = int(sc.view_poly_off[view]) + pl
            closed = sc.poly_start[g] == sc.poly_end[g]     # a loop: `direction == start` is tested first (:401, :589), so both extremes walk towards index 0
            for _ in range(6):
                seg = int(rng.integers(0, len(pc) - 1))
                t = f32(rng.uniform(0, 1))
                c = ref_lerp(pc[seg], pc[seg + 1], t)
                for towards_end in (0, 1):
                    dist = f32(rng.choice([2.0, 10.0, 20.0, 75.0]))
                    fl = L.eg3d_oracle_walk(osc.h, view, pl, 0, seg, c[0], c[1], towards_end, None, dist, C.byref(out_seg), A.ptr(out_xy, A.c_f32p))
                    rs, rc, rr = ref_next_by_distance(pc, seg, c, bool(towards_end) and not closed, dist)
                    assert (bool(fl), out_seg.value, out_xy[0], out_xy[1]) == (rr, rs, rc[0], rc[1]), (view, pl, seg, towards_end, dist)
                    n_reached += rr
                    l = _rand_lines(rng, 1)[0]
                    # aim the line at a point further along the polyline so that intersections are common
                    tgt = pc[min(len(pc) - 1, seg + 2)] if (towards_end and not closed) else pc[max(0, seg - 1)]
                    l[2] = f32(-(l[0] * tgt[0] + l[1] * tgt[1]))
                    fl = L.eg3d_oracle_walk(osc.h, view, pl, 1, seg, c[0], c[1], towards_end, A.ptr(l, A.c_f32p), f32(0), C.byref(out_seg), A.ptr(out_xy, A.c_f32p))
                    rf, rs, rc = ref_next_by_line(pc, seg, c, bool(towards_end) and not closed, l, f32(0.965), f32(5.0))
                    assert bool(fl) == rf, (view, pl, seg, towards_end)
                    if rf:
                        assert (out_seg.value, out_xy[0], out_xy[1]) == (rs, rc[0], rc[1])
                    n_found += rf
                    checked += 1
    assert checked > 400 and n_reached > 20 and n_found > 100


# ---------------------------------------------------------------------------------------------------------------------
# uniform polyline grid: build (polyLine_2d_map.cpp:40-58, polyline_graph_2d.cpp:555-577, 816-835) and lookup
# (edge_graph_3d_utilities.cpp:600-629, polyLine_2d_map_search.cpp:46-88), re-derived in numpy float32

def ref_floor_or_upper_if_close(v):
    """edge_graph_3d_utilities.cpp:600-605: ceil(v) - v (float) < 0.001 (double)."""
    c = f32(np.ceil(v))
    return c if np.float64(f32(c - v)) < 0.001 else f32(np.floor(v))


def ref_is_multiple(m, n):
    """:607-614 (abs is std::abs(float): the file has `using namespace std` and <cmath>)."""
    mul = f32(ref_floor_or_upper_if_close(f32(m / n)) * n)
    return np.float64(f32(abs(f32(m - mul)))) < 0.001


def ref_cell(cell, p):
    return int(ref_floor_or_upper_if_close(f32(p[0] / cell))), int(ref_floor_or_upper_if_close(f32(p[1] / cell)))


def ref_build_grid(sc, view, cell, w, h):
    cell = f32(cell)
    gw, gh = int(np.ceil(f32(w) / cell)), int(np.ceil(f32(h) / cell))
    grid = [[[] for _ in range(gw)] for _ in range(gh)]
    step = f32(np.float64(cell) / (1.414 + 0.1))               # cell_dim / PL_CELL_SPLIT_RATIO, narrowed at the call
    for pl in range(sc.n_polylines(view)):
        pc = sc.polyline(view, pl)
        if len(pc) < 2:
            continue                                           # invalidated polyline
        g = int(sc.view_poly_off[view]) + pl
        closed = sc.poly_start[g] == sc.poly_end[g]            # get_other_end(start) == start: the walk stops at once
        plps = [(0, (pc[0][0], pc[0][1]))]
        seg, c, reached = 0, (pc[0][0], pc[0][1]), False
        while not reached:
            seg, c, reached = ref_next_by_distance(pc, seg, c, not closed, step)
            plps.append((seg, c))
        cells = set()
        for _, c in plps:
            on_boundary = ref_is_multiple(c[0], cell) or ref_is_multiple(c[1], cell)
            if not on_boundary:
                cells.add(ref_cell(cell, c))
        for cx, cy in sorted(cells):
            grid[cy][cx].append(pl)                            # pls_id_maps[cell.second][cell.first]; no bounds check upstream either
    return grid, gw, gh


def ref_grid_query(grid, gw, gh, cell, w, h, p):
    cell = f32(cell)
    if p[0] <= 0 or p[0] >= w or p[1] <= 0 or p[1] >= h:
        return []
    on_row, on_col = ref_is_multiple(p[0], cell), ref_is_multiple(p[1], cell)
    cx, cy = ref_cell(cell, p)
    cx, cy = min(cx, gw - 1), min(cy, gh - 1)
    res = set()
    for i in range(-1 if cy > 0 else 0, (0 if on_row else (1 if cy < gh - 1 else 0)) + 1):
        for j in range(-1 if cx > 0 else 0, (0 if on_col else (1 if cx < gw - 1 else 0)) + 1):
            res.update(grid[cy + i][cx + j])
    return sorted(res)


@pytest.mark.parametrize("which,cell", [(0, 4.0), (1, 30.0)])
def test_polyline_grid_build_and_lookup(which, cell):
    sc = syn.make_scene(n_views=3, n_curves=18, seed=13, closed_frac=0.2, n_tracks=10)     # tracks => the 30 px grid exists too
    assert sc.n_tracks > 0
    osc = O.OracleScene(sc)
    rng = np.random.default_rng(4)
    for view in range(2):
        grid, gw, gh = ref_build_grid(sc, view, cell, sc.width, sc.height)
        pts = [rng.uniform([0, 0], [sc.width, sc.height]).astype(f32) for _ in range(250)]
        for pl in range(sc.n_polylines(view)):                                              # points on and next to polylines
            pc = sc.polyline(view, pl)
            if len(pc) >= 2:
                k = int(rng.integers(0, len(pc)))
                pts.append((pc[k] + rng.normal(0, 1.5, 2)).astype(f32))
                pts.append(pc[k].copy())
        pts += [np.array([f32(cell * k), f32(cell * m + 1.3)], f32) for k, m in ((3, 5), (7, 2))]   # x on a cell boundary
        pts += [np.array([f32(cell * 4 + 0.6), f32(cell * 6)], f32), np.array([f32(cell * 5), f32(cell * 5)], f32)]
        pts += [np.array([f32(0), f32(10)], f32), np.array([f32(sc.width), f32(10)], f32), np.array([f32(sc.width - 0.01), f32(sc.height - 0.01)], f32)]
        nonempty = 0
        for p in pts:
            got = osc.grid_query(view, which, float(p[0]), float(p[1]))
            want = ref_grid_query(grid, gw, gh, cell, sc.width, sc.height, p)
            assert got == want, (view, p, got, want)
            nonempty += bool(want)
        assert nonempty > 30


def ref_minimum_distancesq(p, v, w):
    """geometric_utilities.cpp:940-954 -> (distance squared, projection)."""
    l2 = ref_squared_2d_distance(v, w)
    if l2 == 0.0:
        return ref_squared_2d_distance(p, v), (v[0], v[1])
    pv, wv = (f32(p[0] - v[0]), f32(p[1] - v[1])), (f32(w[0] - v[0]), f32(w[1] - v[1]))
    q = f32(f32(f32(pv[0] * wv[0]) + f32(pv[1] * wv[1])) / l2)
    m = q if q < f32(1) else f32(1)                  # std::min<float>(1, q)
    t = m if f32(0) < m else f32(0)                  # std::max<float>(0, m)
    proj = (f32(v[0] + f32(t * wv[0])), f32(v[1] + f32(t * wv[1])))
    return ref_squared_2d_distance(p, proj), proj


def ref_compute_distancesq(pc, p):
    """polyline::compute_distancesq, polyline_graph_2d.cpp:845-862: the FIRST closest segment wins (strict <)."""
    md, proj = ref_minimum_distancesq(p, pc[0], pc[1])
    seg = 0
    for i in range(2, len(pc)):
        d, cp = ref_minimum_distancesq(p, pc[i - 1], pc[i])
        if d < md:
            md, proj, seg = d, cp, i - 1
    return md, seg, proj


def test_point_to_polyline_distance_bit_exact():
    L = O.lib()
    L.eg3d_oracle_polyline_distancesq.restype = C.c_float
    L.eg3d_oracle_polyline_distancesq.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_float, C.c_float, C.POINTER(C.c_uint32), A.c_f32p]
    sc = syn.make_scene(n_views=2, n_curves=14, seed=17, closed_frac=0.2)
    osc = O.OracleScene(sc)
    rng = np.random.default_rng(8)
    seg = C.c_uint32(); proj = np.zeros(2, f32)
    n = 0
    for view in range(2):
        for pl in range(sc.n_polylines(view)):
            pc = sc.polyline(view, pl)
            if len(pc) < 2:
                continue
            for _ in range(12):
                k = int(rng.integers(0, len(pc)))
                p = (pc[k] + rng.normal(0, rng.choice([0.0, 0.5, 6.0, 60.0]), 2)).astype(f32)
                d = L.eg3d_oracle_polyline_distancesq(osc.h, view, pl, p[0], p[1], C.byref(seg), A.ptr(proj, A.c_f32p))
                rd, rs, rp = ref_compute_distancesq(pc, p)
                assert (f32(d), seg.value, proj[0], proj[1]) == (rd, rs, rp[0], rp[1]), (view, pl, p)
                n += 1
    assert n > 200


def test_density_limiter_second_reading():
    """filter_3d_points_close_2d_array, filtering_close_plgps.cpp:75-124: a point is kept iff at least one of its
    observations falls in a still-empty 3 px cell of its view; a kept point marks all its cells."""
    from edgegraph3d_b200.scene import PointSet
    sc = syn.make_scene(n_views=4, n_curves=6, seed=3)
    rng = np.random.default_rng(12)
    npts = 600
    lens = rng.integers(2, 5, npts)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    views = np.concatenate([rng.choice(4, k, replace=False) for k in lens]).astype(np.int32)
    base = rng.uniform(20, 200, (npts, 2))
    xy = np.concatenate([base[i] + rng.normal(0, 2.0, (lens[i], 2)) for i in range(npts)]).astype(f32)
    ps = PointSet(np.zeros((npts, 3), f32), np.arange(npts, dtype=np.int32), np.zeros(npts, np.int32), off, views,
                  np.zeros(len(views), np.uint32), np.zeros(len(views), np.uint32), xy)
    got = O.OracleScene(sc).dedup_close_points(ps)
    gw, gh = int(np.ceil(f32(sc.width) / 3)), int(np.ceil(f32(sc.height) / 3))
    bm = np.zeros((4, gh, gw), bool)
    want = np.zeros(npts, np.uint8)
    for i in range(npts):
        cells = [(int(views[o]), int(f32(xy[o, 1] / f32(3))), int(f32(xy[o, 0] / f32(3)))) for o in range(off[i], off[i + 1])]
        if any(not bm[c] for c in cells):
            want[i] = 1
            for c in cells:
                bm[c] = True
    assert np.array_equal(got, want) and 0 < want.sum() < npts
