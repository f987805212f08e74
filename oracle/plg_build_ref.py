"""CPU ORACLE for SURVEY §8 row f1 (test infrastructure, NOT product code): edge image -> optimized polyline graph.

A second, independent reading of the reference text in pure Python + numpy float32 scalars (slow: use it on small
images only).  It follows the reference's own containers (std::set -> sorted(set), std::stack -> list, the
unordered_map point map -> dict keyed by the coordinate pair) rather than the flat structures of the product
(edgegraph3d_b200/csrc/eg3d_plg_build.cpp), so that agreement between the two is agreement of two readings.
Only tests/ may import this module.

PARITY UNPINNED by the reference itself: it ships no tests / expected outputs for this stage and cannot be compiled in
the build container (OpenCV, CGAL, Boost headers absent).

Reference (paths relative to the reference root):
  convertEdgeImagePolyLineGraph_optimized      src/edgegraph3d/io/input/convert_edge_images_pixel_to_segment.cpp:880-883
  convertEdgeImagePixelToGraph_NoCycles        same file :347-426 (+ :294-343)
  convert_EdgeGraph_to_PolyLineGraph           same file :583-626 (+ :428-581)
  GraphAdjacencySetNoType::is_connected        src/edgegraph3d/plgs/graph_adjacency_set_no_type.cpp:95-137
  PolyLineGraph2DHMapImpl (get_node_id, add_polyline, optimize and its steps)
                                               src/edgegraph3d/plgs/polyline_graph_2d_hmap_impl.cpp:50-266
  PolyLineGraph2D (simplify, merge, components, smooth-length filter, closest pairs)
                                               src/edgegraph3d/plgs/polyline_graph_2d.cpp:76-99, 295-310, 901-1155, 1315-1350,
                                               1423-1455, 1926-2066
  geometry                                     src/edgegraph3d/utils/geometry/geometric_utilities.cpp:272-312, 432-442, 555-588,
                                               997-1001, 1337-1372
"""
import numpy as np

F = np.float32
INVALID = F(-1)
LOOP_CHECK_DIST = 8                      # convert_edge_images_pixel_to_segment.cpp:345
MAXIMUM_LINEARIZABILITY_DISTANCE = F(1)  # polyline_graph_2d.hpp:68
DIRECT_CONNECTION_EXTREMES_MAXDIST = F(6)  # polyline_graph_2d.hpp:56
TOP_FILTER_BY_POLYLINESMOOTHLENGTH = 0.82  # polyline_graph_2d.hpp:66
STAGE_FULL, STAGE_PIXEL_GRAPH, STAGE_RAW, STAGE_MERGED, STAGE_SIMPLIFIED, STAGE_CONNECTED = range(6)


# ----------------------------------------------------------------------------------------------- geometry (float32)
def squared_2d_distance(a, b):
    """:555-557 — float differences, pow(float,2) and the sum in double, float result."""
    dx, dy = F(a[0] - b[0]), F(a[1] - b[1])
    return F(float(dx) * float(dx) + float(dy) * float(dy))


def compute_2d_distance(a, b):
    return np.sqrt(squared_2d_distance(a, b))  # :571-573, float sqrt


def compute_2dline(a, b):
    """:1341-1354"""
    if a[0] == b[0]:
        return (F(1), F(0), F(-a[0]))
    m = F(F(b[1] - a[1]) / F(b[0] - a[0]))
    q = F(a[1] - F(m * a[0]))
    return (m, F(-1), q)


def distance_point_line_sq(p, line):
    """:997-1001"""
    den = F(F(F(line[0] * p[0]) + F(line[1] * p[1])) + line[2])
    den = F(den * den)
    return F(den / F(F(line[0] * line[0]) + F(line[1] * line[1])))


def intersect_segment_line(segm, line):
    """:272-312 -> (found, point)"""
    dx, dy = F(segm[2] - segm[0]), F(segm[3] - segm[1])
    num = F(F(F(line[0] * segm[0]) + F(line[1] * segm[1])) + line[2])
    den = F(F(line[0] * dx) + F(line[1] * dy))
    if den != 0:
        t = F(-num / den)
        if t >= 0 and t <= 1:
            return True, (F(segm[0] + F(t * dx)), F(segm[1] + F(t * dy)))
    return False, None


def intersect_segment_segment(s1, s2):
    """:437-442 with point_in_segment_bounding_box :432-435"""
    found, p = intersect_segment_line(s2, compute_2dline((s1[0], s1[1]), (s1[2], s1[3])))
    if not found:
        return False
    inx = (s1[0] <= p[0] <= s1[2]) or (s1[2] <= p[0] <= s1[0])
    iny = (s1[1] <= p[1] <= s1[3]) or (s1[3] <= p[1] <= s1[1])
    return inx and iny


def anglecos(a1, a2, b1, b2):
    """:579-588"""
    a = (F(a2[0] - a1[0]), F(a2[1] - a1[1]))
    b = (F(b2[0] - b1[0]), F(b2[1] - b1[1]))
    dot = F(F(a[0] * b[0]) + F(a[1] * b[1]))
    na = F(F(a[0] * a[0]) + F(a[1] * a[1]))
    nb = F(F(b[0] * b[0]) + F(b[1] * b[1]))
    with np.errstate(all="ignore"):
        return F(dot / np.sqrt(F(na * nb)))


# ------------------------------------------------------------------------------------------------------ pixel graph
class AdjacencySetGraph:
    """GraphAdjacencySetUndirectedNoType; `visited` is the member that survives between is_connected calls."""

    def __init__(self, n):
        self.adj = [set() for _ in range(n)]
        self.visited = [False] * n

    def add_edge(self, a, b):
        self.adj[a].add(b)
        self.adj[b].add(a)

    def is_connected(self, start, end, max_dist):
        visited_vec = [bool(start)]            # vector<bool>: the node id is narrowed to true/false
        self.visited[start] = True
        found = False
        cur = [start]                          # std::stack
        dist = 0
        while dist <= max_dist and cur:
            nxt = []
            while not found and cur:
                node = cur.pop()
                for c in sorted(self.adj[node]):
                    if c == end:
                        found = True
                        break
                    if not self.visited[c]:
                        nxt.append(c)
                        self.visited[c] = True
                        visited_vec.append(bool(c))
            for v in visited_vec:              # visited[v] with v in {false, true} -> indices 0 / 1
                if int(v) < len(self.visited):
                    self.visited[int(v)] = False
            cur = nxt
            dist += 1
        return found


def pixel_graph(mask):
    """:294-426.  mask: bool [rows, cols]; returns (graph, node coords)."""
    rows, cols = mask.shape
    flat = mask.astype(bool).ravel().copy()    # Mat(c_img) shares the pixels: clearing is visible to everything after

    def is_edge(i, j):                         # Mat::at without bound checks on a continuous buffer
        k = i * cols + j
        return 0 <= k < flat.size and bool(flat[k])

    node_id = {}
    coords = []
    for i in range(rows):
        for j in range(cols):
            if not is_edge(i, j):
                continue
            useless = ((i > 1 and j > 1 and is_edge(i - 1, j) and is_edge(i, j - 1) and not is_edge(i + 1, j + 1)) or
                       (i > 1 and j < cols - 1 and is_edge(i - 1, j) and is_edge(i, j + 1) and not is_edge(i + 1, j - 1)) or
                       (i < rows - 1 and j < cols - 1 and is_edge(i + 1, j) and is_edge(i, j + 1) and not is_edge(i - 1, j - 1)) or
                       (i < rows - 1 and j > 1 and is_edge(i + 1, j) and is_edge(i, j - 1) and not is_edge(i - 1, j + 1)))
            if useless:
                flat[i * cols + j] = False
            else:
                node_id[(i, j)] = len(coords)
                coords.append((F(j + 0.5), F(i + 0.5)))
    g = AdjacencySetGraph(len(coords))
    for i in range(rows - 1):
        for j in range(cols - 1):
            if not flat[i * cols + j]:
                continue
            p = node_id[(i, j)]
            neighbours = [(i, j + 1), (i + 1, j), (i + 1, j + 1)]
            if j > 1:
                neighbours.append((i + 1, j - 1))
            for (ci, cj) in neighbours:
                if flat[ci * cols + cj]:
                    c = node_id[(ci, cj)]
                    if p != c and not g.is_connected(p, c, LOOP_CHECK_DIST):
                        g.add_edge(p, c)
    return g, coords


# ---------------------------------------------------------------------------------------- pixel graph -> polylines
def find_polylineend_no_come_back(start, adj, no_come_back, res):
    prev, cur = no_come_back, start
    res.append(cur)
    while cur != no_come_back and len(adj[cur]) == 2:
        lo, hi = sorted(adj[cur])
        nxt = lo if lo != prev else hi
        prev, cur = cur, nxt
        res.append(cur)


def find_polylines(start, adj):
    n = len(adj[start])
    if n == 2:
        prev, nxt = sorted(adj[start])
        res = []
        find_polylineend_no_come_back(prev, adj, start, res)
        res.reverse()
        res.append(start)
        if res[0] == res[-1]:
            return [res]
        find_polylineend_no_come_back(nxt, adj, start, res)
        return [res]
    if n == 1:
        res = [start]
        find_polylineend_no_come_back(next(iter(adj[start])), adj, start, res)
        return [res]
    if n > 2:
        out = []
        for nb in sorted(adj[start]):
            res = [start]
            find_polylineend_no_come_back(nb, adj, start, res)
            out.append(res)
        return out
    return []


# --------------------------------------------------------------------------------------------------- polyline graph
class Polyline:
    def __init__(self, start, end, coords):
        self.start, self.end, self.coords = start, end, list(coords)
        self.update_length()

    def update_length(self):
        length = F(0)
        for i in range(1, len(self.coords)):
            length = F(length + compute_2d_distance(self.coords[i], self.coords[i - 1]))
        self.length = length

    def same_as(self, o):
        return ((self.start == o.start and self.end == o.end and self.coords == o.coords) or
                (self.start == o.end and self.end == o.start and self.coords == o.coords[::-1]))

    def other_end(self, n):
        return self.end if n == self.start else self.start


def linearizable(c, s, e, max_dsq):
    line = compute_2dline(c[s], c[e])
    return all(not (distance_point_line_sq(c[i], line) > max_dsq) for i in range(s + 1, e))


def find_max_se(c, s, max_se, max_dsq):
    if max_se <= s:
        return s
    k = max_se
    while k > s + 1:
        if linearizable(c, s, k, max_dsq):
            return k
        k -= 1
    return s + 1


def find_min_eb(c, e, min_eb, max_dsq):
    if min_eb >= e:
        return e
    k = min_eb
    while k < e - 1:
        if linearizable(c, k, e, max_dsq):
            return k
        k += 1
    return e - 1


def find_compatible_se_eb(c, s, e, max_dsq):
    if s >= e:
        return s, s
    max_se, min_eb = e, s
    while True:
        se = find_max_se(c, s, max_se, max_dsq)
        if se == e:
            return se, 0
        eb = find_min_eb(c, e, min_eb, max_dsq)
        max_se -= 1
        min_eb += 1
        if not (eb < se):
            return se, eb


def simplify_polyline(c, max_d):
    max_dsq = F(max_d * max_d)
    s, e = 0, len(c) - 1
    head, tail = [c[s]], [c[e]]
    while e > s + 1:
        se, eb = find_compatible_se_eb(c, s, e, max_dsq)
        if se == e:
            break
        head.append(c[se])
        if se != eb:
            tail.append(c[eb])
        s, e = se, eb
    return head + tail[::-1]


class PolylineGraph:
    def __init__(self):
        self.polylines, self.connections, self.nodes, self.point_map = [], [], [], {}

    # --- validity
    def valid_node(self, n):
        return self.nodes[n][0] != INVALID and self.nodes[n][1] != INVALID

    def valid_polyline(self, i):
        p = self.polylines[i]
        return (self.valid_node(p.start) and self.valid_node(p.end) and len(p.coords) > 1 and
                self.nodes[p.start] == p.coords[0] and self.nodes[p.end] == p.coords[-1])

    # --- removal
    def base_invalidate_node(self, n):
        self.nodes[n] = (INVALID, INVALID)
        for pid in list(self.connections[n]):
            self.remove_polyline(pid)
        self.connections[n] = []

    def hmap_invalidate_node(self, n):
        self.point_map.pop(self.nodes[n], None)
        self.base_invalidate_node(n)

    def remove_connection(self, n, pid):
        self.connections[n] = [x for x in self.connections[n] if x != pid]
        if not self.connections[n]:
            self.base_invalidate_node(n)      # the base-class method: the point map keeps its (now stale) entry

    def remove_polyline(self, pid):
        p = self.polylines[pid]
        self.remove_connection(p.start, pid)
        self.remove_connection(p.end, pid)
        p.coords = []
        p.length = INVALID

    # --- construction
    def get_node_id(self, xy):
        n = self.point_map.get(xy)
        if n is not None and not self.valid_node(n):
            self.hmap_invalidate_node(n)
            n = None
        if n is None:
            n = len(self.nodes)
            self.point_map[xy] = n
            self.connections.append([])
            self.nodes.append(xy)
        return n

    def is_duplicate(self, pl):
        s, e = self.connections[pl.start], self.connections[pl.end]
        smallest = s if len(s) < len(e) else e
        return any(self.polylines[i].same_as(pl) for i in smallest)

    def internal_add_polyline(self, pl):
        if self.is_duplicate(pl):
            return
        pid = len(self.polylines)
        self.polylines.append(pl)
        self.connections[pl.start].append(pid)
        if pl.start != pl.end:
            self.connections[pl.end].append(pid)

    def add_polyline(self, coords):
        if coords[0] == coords[-1] and len(coords) == 4 and squared_2d_distance(coords[1], coords[2]) <= 4:
            mid = (F(F(coords[1][0] + coords[2][0]) / 2), F(F(coords[1][1] + coords[2][1]) / 2))
            coords = [coords[0], mid]
        s = self.get_node_id(coords[0])
        e = self.get_node_id(coords[-1])
        self.internal_add_polyline(Polyline(s, e, coords))

    # --- predicates
    def is_loop(self, pid):
        return self.polylines[pid].start == self.polylines[pid].end

    def is_extreme(self, n):
        return len(self.connections[n]) == 1 and not self.is_loop(self.connections[n][0])

    # --- optimize() steps
    def remove_invalid_polylines(self):
        for i in range(len(self.polylines)):
            if not self.valid_polyline(i):
                self.remove_polyline(i)

    def remove_degenerate_loops(self):
        for i in range(len(self.polylines)):
            if self.valid_polyline(i):
                p = self.polylines[i]
                if (p.start == p.end or p.coords[0] == p.coords[-1]) and len(p.coords) < 5:
                    self.remove_polyline(i)

    @staticmethod
    def merge_polylines(p1, p2):
        if p1.start == p2.start:
            return Polyline(p1.end, p2.end, p1.coords[::-1] + p2.coords[1:])
        if p1.start == p2.end:
            return Polyline(p2.start, p1.end, p2.coords + p1.coords[1:])
        if p1.end == p2.start:
            return Polyline(p1.start, p2.end, p1.coords + p2.coords[1:])
        if p1.end == p2.end:
            return Polyline(p1.start, p2.start, p1.coords + p2.coords[::-1][1:])
        raise ValueError("cannot merge disconnected polylines")

    def remove_2connection_nodes(self):
        for n in range(len(self.connections)):
            if len(self.connections[n]) != 2:
                continue
            id1, id2 = self.connections[n]
            p1, p2 = self.polylines[id1], self.polylines[id2]
            o1, o2 = p1.other_end(n), p2.other_end(n)
            if p1.coords == p2.coords or p1.coords == p2.coords[::-1]:
                self.remove_polyline(id2)
                continue
            if o1 != n and o2 != n:
                p3 = self.merge_polylines(p1, p2)
                self.internal_add_polyline(p3)
                self.remove_polyline(id1)
                self.remove_polyline(id2)
                self.hmap_invalidate_node(n)

    def simplify(self):
        for i in range(len(self.polylines)):
            if self.valid_polyline(i):
                p = self.polylines[i]
                p.coords = simplify_polyline(p.coords, MAXIMUM_LINEARIZABILITY_DISTANCE)
                p.update_length()

    def compute_components(self):
        n_nodes = len(self.nodes)
        comp_of = [0] * n_nodes
        comps = []
        explored = [False] * n_nodes
        in_to_explore = [False] * n_nodes
        stack = []
        for s in range(n_nodes):
            if explored[s]:
                continue
            explored[s] = True
            cur = {s}
            cid = len(comps)
            comp_of[s] = cid
            for pid in self.connections[s]:
                o = self.polylines[pid].other_end(s)
                in_to_explore[o] = True
                stack.append(o)
            while stack:
                c = stack.pop()
                cur.add(c)
                comp_of[c] = cid
                in_to_explore[c] = False
                explored[c] = True
                for pid in self.connections[c]:
                    o = self.polylines[pid].other_end(c)
                    if not explored[o] or o == c:
                        if not in_to_explore[o] and o != c:
                            stack.append(o)
            comps.append(cur)
        return comp_of, comps

    def intersect_polylines_nonempty(self, a, b):
        q = (a[0], a[1], b[0], b[1])
        for i in range(len(self.polylines)):
            if self.valid_polyline(i):
                c = self.polylines[i].coords
                for k in range(1, len(c)):
                    if intersect_segment_segment((c[k][0], c[k][1], c[k - 1][0], c[k - 1][1]), q):
                        return True
        return False

    def connect_close_extremes(self):
        ids = [n for n in range(len(self.nodes)) if self.valid_node(n) and self.is_extreme(n)]
        pts = [self.nodes[n] for n in ids]
        max_dsq = F(DIRECT_CONNECTION_EXTREMES_MAXDIST * DIRECT_CONNECTION_EXTREMES_MAXDIST)
        closest = [None] * len(pts)
        pairs = []
        for i in range(len(pts)):
            best, bi = np.finfo(np.float32).max, None
            for j in range(len(pts)):
                if j != i:
                    d = squared_2d_distance(pts[i], pts[j])
                    if d < best:
                        best, bi = d, j
            closest[i] = bi
            if bi is not None and bi < i and closest[bi] == i and squared_2d_distance(pts[i], pts[bi]) <= max_dsq:
                pairs.append((ids[i], ids[bi]))
        comp_of, comps = self.compute_components()
        for a, b in pairs:
            if comp_of[a] == comp_of[b]:
                continue
            if self.intersect_polylines_nonempty(self.nodes[a], self.nodes[b]):
                continue
            self.internal_add_polyline(Polyline(a, b, [self.nodes[a], self.nodes[b]]))
            if len(comps[comp_of[a]]) < len(comps[comp_of[b]]):
                new_id, change = comp_of[b], comp_of[a]
            else:
                new_id, change = comp_of[a], comp_of[b]
            for n in comps[change]:
                comp_of[n] = new_id

    @staticmethod
    def compute_max_smooth_length(c):
        maxl = F(0)
        i = 1
        while i < len(c):
            cur = compute_2d_distance(c[i], c[i - 1])
            i += 1
            while i < len(c):
                cosv = anglecos(c[i - 1], c[i], c[i - 2], c[i - 1])
                if cosv != 0:                     # `if(float)`: only an exact 0 ends the section (NaN is truthy)
                    cur = F(cur + compute_2d_distance(c[i], c[i - 1]))
                    i += 1
                else:
                    break
            if maxl < cur:
                maxl = cur
        return maxl

    def filter_components_by_polylinesmoothlength(self):
        npl = len(self.polylines)
        if npl == 0:
            return
        comp_of, comps = self.compute_components()
        comp_polys = [set() for _ in comps]
        for n in range(len(self.nodes)):
            if self.valid_node(n):
                for pid in self.connections[n]:
                    comp_polys[comp_of[n]].add(pid)
        sl = [F(0)] * npl
        for i in range(npl):
            if self.valid_polyline(i):
                sl[i] = self.compute_max_smooth_length(self.polylines[i].coords)
        k = int(npl * TOP_FILTER_BY_POLYLINESMOOTHLENGTH)
        thr = sorted(sl)[k]                       # nth_element: the k-th smallest value
        for cid, polys in enumerate(comp_polys):
            if not any(sl[pid] >= thr for pid in sorted(polys)):
                for pid in sorted(polys):
                    self.remove_polyline(pid)


def polyline_graph_from_mask(mask, stop_after=STAGE_FULL):
    """convertEdgeImagePolyLineGraph_optimized on a boolean edge mask.  Returns (PolylineGraph, AdjacencySetGraph, coords)."""
    g, coords = pixel_graph(np.asarray(mask, bool))
    plg = PolylineGraph()
    if stop_after == STAGE_PIXEL_GRAPH:
        return plg, g, coords
    processed = [False] * len(coords)
    for i in range(len(coords)):
        if processed[i]:
            continue
        for ids in find_polylines(i, g.adj):
            s, e = ids[0], ids[-1]
            plg.get_node_id(coords[s])
            plg.get_node_id(coords[e])
            if not len(g.adj[s]) > 2:
                processed[s] = True
            if not len(g.adj[e]) > 2:
                processed[e] = True
            for k in ids[1:-1]:
                processed[k] = True
            plg.add_polyline([coords[k] for k in ids])
        processed[i] = True
    if stop_after == STAGE_RAW:
        return plg, g, coords
    plg.remove_invalid_polylines()
    plg.remove_degenerate_loops()
    plg.remove_2connection_nodes()
    if stop_after == STAGE_MERGED:
        return plg, g, coords
    plg.simplify()
    if stop_after == STAGE_SIMPLIFIED:
        return plg, g, coords
    plg.connect_close_extremes()
    plg.simplify()
    # split_loops(): is_loop() means start == end, so next_pl_point_by_length takes its `direction == start` branch from
    # segment 0 with zero length walked and reports reached_polyline_extreme; split_loop then does nothing (hmap_impl.cpp:238-254).
    if stop_after == STAGE_CONNECTED:
        return plg, g, coords
    plg.filter_components_by_polylinesmoothlength()
    return plg, g, coords
