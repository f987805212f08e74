// eg3d_plg_build.cpp — SURVEY §8 row f1 (host side, upstream of the hot path): edge image -> optimized 2D polyline graph.
//
// Replaces convertEdgeImagePolyLineGraph_optimized (src/edgegraph3d/io/input/convert_edge_images_pixel_to_segment.cpp:880-883),
// i.e. convertEdgeImagePixelToGraph_NoCycles (:347-426) -> convert_EdgeGraph_to_PolyLineGraph (:583-626) ->
// PolyLineGraph2DHMapImpl::optimize (src/edgegraph3d/plgs/polyline_graph_2d_hmap_impl.cpp:255-266).
//
// The whole stage is order-dependent graph surgery on one image (raster-scan pixel clearing, a breadth-first loop check
// whose visited flags persist between queries, id-ordered merges), so it stays on the host; views are independent and
// the Python side runs them on a thread pool.  Polyline ids, node ids, vertex order and every float are meant to equal
// the reference's: the downstream matching path is sensitive to all of them.  Expressions are written in the
// reference's operation order (float unless noted; x86-64 host code has no FMA contraction without -march flags).
//
// Quirks reproduced on purpose (paths relative to the reference root):
//  * GraphAdjacencySetNoType::is_connected(a,b,max_dist) (src/edgegraph3d/plgs/graph_adjacency_set_no_type.cpp:95-137)
//    keeps its `visited` member between calls: the per-call undo list is a vector<bool>, so node ids are narrowed to
//    0/1 and only visited[0] / visited[1] are ever reset.  The loop check therefore only explores nodes no earlier
//    query has touched.
//  * convertEdgeImagesPixelToNodesNoSquaresNoTriangles_remove_useless_hubs (:294-343) works on `Mat img = Mat(c_img)`,
//    a header sharing the caller's pixels, so cleared pixels are seen by the edge pass; its guards are `i>1`, `j>1`
//    (not >0) and it reads (i+1, j±1) / (i-1, j±1) without a matching bound check.  On a continuous Mat those reads
//    wrap to the neighbouring row; reads outside the buffer are treated as "not an edge" here.
//  * polyline::compute_max_smooth_length (src/edgegraph3d/plgs/polyline_graph_2d.cpp:82-99) uses the angle cosine as a
//    boolean: a section only ends at a cosine of exactly 0.
//  * PolyLineGraph2DHMapImpl::split_loops (:247-254) never splits: for a loop start == end, so
//    next_pl_point_by_length takes the `direction == start` branch from segment 0 with zero accumulated length and
//    reports reached_polyline_extreme (polyline_graph_2d.cpp:455-472), and split_loop only splits when it does not.
//  * PolyLineGraph2D::remove_connection (polyline_graph_2d.cpp:1040-1045) calls the BASE invalidate_node, which does not
//    erase the point-map entry; get_node_id (hmap_impl.cpp:50-72) repairs that lazily.
#include "../../include/eg3d.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

eg3d_status eg3d_internal_fail(eg3d_status s, const char* msg);  // eg3d_capi.cu

namespace {
typedef uint64_t u64;
struct V2 { float x, y; };
inline bool eq(const V2& a, const V2& b) { return a.x == b.x && a.y == b.y; }

// geometric_utilities.cpp:555-557 (pow(float,2) is a double product), :571-573
inline float sqdist(const V2& a, const V2& b) {
  float dx = a.x - b.x, dy = a.y - b.y;
  return (float)((double)dx * (double)dx + (double)dy * (double)dy);
}
inline float dist(const V2& a, const V2& b) { return std::sqrt(sqdist(a, b)); }

// geometric_utilities.cpp:1341-1354
inline void line_through(const V2& a, const V2& b, float l[3]) {
  if (a.x == b.x) { l[0] = 1; l[1] = 0; l[2] = -a.x; }
  else { float m = (b.y - a.y) / (b.x - a.x); float q = a.y - m * a.x; l[0] = m; l[1] = -1.0f; l[2] = q; }
}
// geometric_utilities.cpp:997-1001
inline float dist_point_line_sq(const V2& p, const float l[3]) {
  float den = l[0] * p.x + l[1] * p.y + l[2];
  den *= den;
  return den / (l[0] * l[0] + l[1] * l[1]);
}
// geometric_utilities.cpp:272-312 (only `intersection_found` and the point are consumed here)
inline bool segment_line(const float s[4], const float l[3], V2& out) {
  float dx = s[2] - s[0], dy = s[3] - s[1];
  float num = l[0] * s[0] + l[1] * s[1] + l[2];
  float den = l[0] * dx + l[1] * dy;
  if (den != 0) {
    float t = -num / den;
    if (t >= 0 && t <= 1) { out.x = s[0] + t * dx; out.y = s[1] + t * dy; return true; }
  }
  return false;
}
// geometric_utilities.cpp:432-442
inline bool segment_segment(const float s1[4], const float s2[4], V2& out) {
  float l1[3]; line_through(V2{s1[0], s1[1]}, V2{s1[2], s1[3]}, l1);
  if (!segment_line(s2, l1, out)) return false;
  return ((s1[0] <= out.x && out.x <= s1[2]) || (s1[2] <= out.x && out.x <= s1[0])) &&
         ((s1[1] <= out.y && out.y <= s1[3]) || (s1[3] <= out.y && out.y <= s1[1]));
}

// ---------------------------------------------------------------------------------------------------------------
// Stage 1+2: pixel graph (convert_edge_images_pixel_to_segment.cpp:294-426)
// ---------------------------------------------------------------------------------------------------------------
struct PixelGraph {
  std::vector<std::vector<u64>> adj;   // std::set<ulong> per node: kept sorted, unique
  std::vector<V2> xy;
  std::vector<uint8_t> visited;        // GraphAdjacencySetNoType::visited — persists between is_connected calls
  void insert(u64 a, u64 b) {
    auto& v = adj[a];
    auto it = std::lower_bound(v.begin(), v.end(), b);
    if (it == v.end() || *it != b) v.insert(it, b);
  }
  void add_edge(u64 a, u64 b) { insert(a, b); insert(b, a); }  // graph_adjacency_set_undirected_no_type.cpp:39-42
  // graph_adjacency_set_no_type.cpp:95-137
  bool is_connected(u64 start, u64 end, u64 max_dist) {
    bool undo0 = (start == 0), undo1 = (start != 0);  // vector<bool> visited_vec: ids narrowed to bool
    visited[start] = 1;
    bool found = false;
    std::vector<u64> cur{start}, next;                 // std::stack: back() is the top
    u64 d = 0;
    while (d <= max_dist && !cur.empty()) {
      next.clear();
      while (!found && !cur.empty()) {
        const u64 c = cur.back(); cur.pop_back();
        for (u64 nb : adj[c]) {
          if (nb == end) { found = true; break; }
          if (!visited[nb]) { next.push_back(nb); visited[nb] = 1; (nb == 0 ? undo0 : undo1) = true; }
        }
      }
      if (undo0 && !visited.empty()) visited[0] = 0;
      if (undo1 && visited.size() > 1) visited[1] = 0;
      cur.swap(next);                                  // cur_to_visit = next_to_visit (leftovers of cur are dropped)
      d++;
    }
    return found;
  }
};

void build_pixel_graph(std::vector<uint8_t>& m, int rows, int cols, PixelGraph& g) {
  const int64_t total = (int64_t)rows * cols;
  auto E = [&](int i, int j) -> bool {  // img.at<Vec3b>(i,j) == edge_color on a continuous Mat, no bound check
    int64_t k = (int64_t)i * cols + j;
    return k >= 0 && k < total && m[k];
  };
  std::vector<u64> id((size_t)total, 0);
  // :294-343, raster order; a cleared pixel changes what later pixels see
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++)
      if (E(i, j)) {
        if ((i > 1 && j > 1 && E(i - 1, j) && E(i, j - 1) && !E(i + 1, j + 1)) ||
            (i > 1 && j < cols - 1 && E(i - 1, j) && E(i, j + 1) && !E(i + 1, j - 1)) ||
            (i < rows - 1 && j < cols - 1 && E(i + 1, j) && E(i, j + 1) && !E(i - 1, j - 1)) ||
            (i < rows - 1 && j > 1 && E(i + 1, j) && E(i, j - 1) && !E(i - 1, j + 1))) {
          m[(size_t)i * cols + j] = 0;
        } else {
          id[(size_t)i * cols + j] = g.xy.size();
          g.xy.push_back(V2{(float)(j + 0.5), (float)(i + 0.5)});
        }
      }
  g.adj.assign(g.xy.size(), {});
  g.visited.assign(g.xy.size(), 0);
  // :347-426  P -> right, below, below-right, and (for j > 1) below-left, unless already linked within 8 hops
  auto link = [&](u64 p, int ci, int cj) {
    if (!m[(size_t)ci * cols + cj]) return;
    const u64 c = id[(size_t)ci * cols + cj];
    if (p != c && !g.is_connected(p, c, 8)) g.add_edge(p, c);
  };
  for (int i = 0; i < rows - 1; i++)
    for (int j = 0; j < cols - 1; j++)
      if (m[(size_t)i * cols + j]) {
        const u64 p = id[(size_t)i * cols + j];
        link(p, i, j + 1);
        link(p, i + 1, j);
        link(p, i + 1, j + 1);
        if (j > 1) link(p, i + 1, j - 1);
      }
}

// ---------------------------------------------------------------------------------------------------------------
// Stage 3: pixel graph -> polylines (convert_edge_images_pixel_to_segment.cpp:428-581)
// ---------------------------------------------------------------------------------------------------------------
// :487-501
void walk_to_polyline_end(u64 start, const std::vector<std::vector<u64>>& adj, u64 no_come_back, std::vector<u64>& res) {
  u64 prev = no_come_back, cur = start;
  res.push_back(cur);
  while (cur != no_come_back && adj[cur].size() == 2) {
    const u64 a = adj[cur][0], b = adj[cur][1];
    const u64 nx = (a != prev) ? a : b;  // :476-482
    prev = cur; cur = nx;
    res.push_back(cur);
  }
}
// :557-574
std::vector<std::vector<u64>> find_polylines(u64 s, const std::vector<std::vector<u64>>& adj) {
  std::vector<std::vector<u64>> res;
  const size_t n = adj[s].size();
  if (n == 2) {                       // :514-539
    std::vector<u64> r;
    walk_to_polyline_end(adj[s][0], adj, s, r);
    std::reverse(r.begin(), r.end());
    r.push_back(s);
    if (r.front() != r.back()) walk_to_polyline_end(adj[s][1], adj, s, r);
    res.push_back(std::move(r));
  } else if (n == 1) {                // :541-555
    std::vector<u64> r{s};
    walk_to_polyline_end(adj[s][0], adj, s, r);
    res.push_back(std::move(r));
  } else if (n > 2) {                 // :503-512
    for (u64 nb : adj[s]) {
      std::vector<u64> r{s};
      walk_to_polyline_end(nb, adj, s, r);
      res.push_back(std::move(r));
    }
  }
  return res;
}

// ---------------------------------------------------------------------------------------------------------------
// PolyLineGraph2DHMapImpl: only what construction + optimize() touch
// ---------------------------------------------------------------------------------------------------------------
struct Poly {
  u64 start = 0, end = 0;
  std::vector<V2> pc;
  float length = 0;
  void update_length() {  // polyline_graph_2d.cpp:76-80
    length = 0.0f;
    for (size_t i = 1; i < pc.size(); i++) length += dist(pc[i], pc[i - 1]);
  }
};
inline bool coords_equal(const std::vector<V2>& a, const std::vector<V2>& b) {      // edge_graph_3d_utilities.cpp:470-480
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); i++) if (!eq(a[i], b[i])) return false;
  return true;
}
inline bool coords_equal_inv(const std::vector<V2>& a, const std::vector<V2>& b) {  // :497-507
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); i++) if (!eq(a[i], b[a.size() - i - 1])) return false;
  return true;
}
inline bool same_polyline(const Poly& a, const Poly& b) {  // polyline_graph_2d.cpp:1026-1029
  return (a.start == b.start && a.end == b.end && coords_equal(a.pc, b.pc)) ||
         (a.start == b.end && a.end == b.start && coords_equal_inv(a.pc, b.pc));
}

// polyline_graph_2d.cpp:910-1013
bool linearizable(const std::vector<V2>& c, u64 s, u64 e, float max_dsq) {
  float l[3]; line_through(c[s], c[e], l);
  for (u64 i = s + 1; i < e; i++) if (dist_point_line_sq(c[i], l) > max_dsq) return false;
  return true;
}
u64 find_max_se(const std::vector<V2>& c, u64 s, u64 max_se, float max_dsq) {
  if (max_se <= s) return s;
  for (u64 k = max_se; k > s + 1; k--) if (linearizable(c, s, k, max_dsq)) return k;
  return s + 1;
}
u64 find_min_eb(const std::vector<V2>& c, u64 e, u64 min_eb, float max_dsq) {
  if (min_eb >= e) return e;
  for (u64 k = min_eb; k < e - 1; k++) if (linearizable(c, k, e, max_dsq)) return k;
  return e - 1;
}
void find_compatible_se_eb(const std::vector<V2>& c, u64 s, u64 e, float max_dsq, u64& se, u64& eb) {
  if (s >= e) { se = eb = s; return; }
  u64 max_se = e, min_eb = s;
  do {
    se = find_max_se(c, s, max_se, max_dsq);
    if (se == e) { eb = 0; break; }
    eb = find_min_eb(c, e, min_eb, max_dsq);
    max_se--; min_eb++;
  } while (eb < se);
}
std::vector<V2> simplify_coords(const std::vector<V2>& c, float max_d) {
  const float max_dsq = max_d * max_d;
  u64 s = 0, e = c.size() - 1;
  std::vector<V2> head{c[s]}, tail{c[e]};
  while (e > s + 1) {
    u64 se, eb;
    find_compatible_se_eb(c, s, e, max_dsq, se, eb);
    if (se == e) break;
    head.push_back(c[se]);
    if (se != eb) tail.push_back(c[eb]);
    s = se; e = eb;
  }
  for (auto it = tail.rbegin(); it != tail.rend(); ++it) head.push_back(*it);
  return head;
}

struct Plg {
  std::vector<Poly> polylines;
  std::vector<std::vector<u64>> connections;   // node -> polyline ids, insertion order
  std::vector<V2> nodes;
  std::unordered_map<u64, u64> point_map;      // key = the two float bit patterns (coordinates are positive finite)

  static u64 key(const V2& p) { uint32_t a, b; std::memcpy(&a, &p.x, 4); std::memcpy(&b, &p.y, 4); return ((u64)a << 32) | b; }
  bool valid_node(u64 n) const { return nodes[n].x != -1.0f && nodes[n].y != -1.0f; }  // polyline_graph_2d.cpp:1137-1139
  bool valid_polyline(u64 i) const {                                                       // :1141-1147
    const Poly& p = polylines[i];
    return valid_node(p.start) && valid_node(p.end) && p.pc.size() > 1 && eq(nodes[p.start], p.pc.front()) && eq(nodes[p.end], p.pc.back());
  }
  void invalidate_node_base(u64 n) {  // :1129-1135 (the reference iterates the list it is erasing from; in every flow here it is already empty)
    nodes[n] = V2{-1.0f, -1.0f};
    const std::vector<u64> c = connections[n];
    for (u64 pid : c) remove_polyline(pid);
    connections[n].clear();
  }
  void invalidate_node_hmap(u64 n) {  // polyline_graph_2d_hmap_impl.cpp:162-165
    point_map.erase(key(nodes[n]));
    invalidate_node_base(n);
  }
  void remove_connection(u64 n, u64 pid) {  // polyline_graph_2d.cpp:1040-1045
    auto& v = connections[n];
    v.erase(std::remove(v.begin(), v.end(), pid), v.end());
    if (v.empty()) invalidate_node_base(n);
  }
  void remove_polyline(u64 pid) {  // :1047-1052, :1035-1038
    const u64 s = polylines[pid].start, e = polylines[pid].end;
    remove_connection(s, pid);
    remove_connection(e, pid);
    polylines[pid].pc.clear();
    polylines[pid].length = -1.0f;
  }
  u64 get_node_id(const V2& p) {  // polyline_graph_2d_hmap_impl.cpp:50-72
    auto it = point_map.find(key(p));
    if (it != point_map.end() && !valid_node(it->second)) { invalidate_node_hmap(it->second); it = point_map.end(); }
    if (it == point_map.end()) {
      const u64 n = nodes.size();
      point_map[key(p)] = n;
      connections.emplace_back();
      nodes.push_back(p);
      return n;
    }
    return it->second;
  }
  bool is_duplicate(const Poly& pl) const {  // :112-123
    const auto& s = connections[pl.start]; const auto& e = connections[pl.end];
    const auto& c = s.size() < e.size() ? s : e;
    for (u64 id : c) if (same_polyline(polylines[id], pl)) return true;
    return false;
  }
  void internal_add_polyline(const Poly& pl) {  // :74-82
    if (is_duplicate(pl)) return;
    const u64 id = polylines.size();
    polylines.push_back(pl);
    connections[pl.start].push_back(id);
    if (pl.start != pl.end) connections[pl.end].push_back(id);
  }
  void add_polyline(const std::vector<V2>& in) {  // :84-104
    std::vector<V2> c;
    if (eq(in.front(), in.back()) && in.size() == 4 && sqdist(in[1], in[2]) <= 4) {
      c.push_back(in[0]);
      c.push_back(V2{(in[1].x + in[2].x) / 2, (in[1].y + in[2].y) / 2});
    } else c = in;
    Poly pl; pl.start = get_node_id(c.front()); pl.end = get_node_id(c.back()); pl.pc = std::move(c); pl.update_length();
    internal_add_polyline(pl);
  }
  bool is_loop(u64 pid) const { return polylines[pid].start == polylines[pid].end; }
  bool is_extreme(u64 n) const { return connections[n].size() == 1 && !is_loop(connections[n][0]); }  // polyline_graph_2d.cpp:1444-1447
  u64 other_end(const Poly& p, u64 n) const { return n == p.start ? p.end : p.start; }                 // :901-908

  // --- optimize() steps, polyline_graph_2d_hmap_impl.cpp ---
  void remove_invalid_polylines() {  // :215-221
    const u64 n = polylines.size();
    for (u64 i = 0; i < n; i++) if (!valid_polyline(i)) remove_polyline(i);
  }
  void remove_degenerate_loops() {  // :203-213
    const u64 n = polylines.size();
    for (u64 i = 0; i < n; i++)
      if (valid_polyline(i)) {
        const Poly& p = polylines[i];
        if ((p.start == p.end || eq(p.pc.front(), p.pc.back())) && p.pc.size() < 5) remove_polyline(i);
      }
  }
  static Poly merge(const Poly& a, const Poly& b) {  // polyline_graph_2d.cpp:1060-1115
    Poly r;
    if (a.start == b.start) {
      r.start = a.end; r.end = b.end;
      r.pc.assign(a.pc.rbegin(), a.pc.rend());
      r.pc.insert(r.pc.end(), b.pc.begin() + 1, b.pc.end());
    } else if (a.start == b.end) {
      r.start = b.start; r.end = a.end;
      r.pc = b.pc;
      r.pc.insert(r.pc.end(), a.pc.begin() + 1, a.pc.end());
    } else if (a.end == b.start) {
      r.start = a.start; r.end = b.end;
      r.pc = a.pc;
      r.pc.insert(r.pc.end(), b.pc.begin() + 1, b.pc.end());
    } else {  // a.end == b.end
      r.start = a.start; r.end = b.start;
      r.pc = a.pc;
      r.pc.insert(r.pc.end(), b.pc.rbegin() + 1, b.pc.rend());
    }
    r.update_length();
    return r;
  }
  void remove_2connection_nodes() {  // :167-201
    for (u64 n = 0; n < connections.size(); n++)
      if (connections[n].size() == 2) {
        const u64 id1 = connections[n][0], id2 = connections[n][1];
        const u64 o1 = other_end(polylines[id1], n), o2 = other_end(polylines[id2], n);
        if (coords_equal(polylines[id1].pc, polylines[id2].pc) || coords_equal_inv(polylines[id1].pc, polylines[id2].pc)) {
          remove_polyline(id2);
          continue;
        }
        if (o1 != n && o2 != n) {
          const Poly p3 = merge(polylines[id1], polylines[id2]);
          internal_add_polyline(p3);
          remove_polyline(id1);
          remove_polyline(id2);
          invalidate_node_hmap(n);
        }
      }
  }
  void simplify_all() {  // polyline_graph_2d.cpp:1149-1155, :1019-1024; MAXIMUM_LINEARIZABILITY_DISTANCE 1.0
    const u64 n = polylines.size();
    for (u64 i = 0; i < n; i++)
      if (valid_polyline(i)) { polylines[i].pc = simplify_coords(polylines[i].pc, 1.0f); polylines[i].update_length(); }
  }
  // polyline_graph_2d.cpp:1926-1986
  void compute_components(std::vector<u64>& comp_of, std::vector<std::vector<u64>>& comp_nodes) const {
    const u64 N = nodes.size();
    comp_of.assign(N, 0); comp_nodes.clear();
    std::vector<uint8_t> explored(N, 0), queued(N, 0);
    std::vector<u64> st;
    for (u64 s = 0; s < N; s++) {
      if (explored[s]) continue;
      explored[s] = 1;
      std::vector<u64> cur{s};
      const u64 cid = comp_nodes.size();
      comp_of[s] = cid;
      for (u64 pid : connections[s]) { const u64 o = other_end(polylines[pid], s); queued[o] = 1; st.push_back(o); }
      while (!st.empty()) {
        const u64 c = st.back(); st.pop_back();
        cur.push_back(c);
        comp_of[c] = cid;
        queued[c] = 0; explored[c] = 1;
        for (u64 pid : connections[c]) {
          const u64 o = other_end(polylines[pid], c);
          if ((!explored[o] || o == c) && !queued[o] && o != c) st.push_back(o);
        }
      }
      std::sort(cur.begin(), cur.end());
      cur.erase(std::unique(cur.begin(), cur.end()), cur.end());  // set<ulong>
      comp_nodes.push_back(std::move(cur));
    }
  }
  bool segment_hits_any_polyline(const V2& a, const V2& b) const {  // intersect_polylines(...).size() != 0, polyline_graph_2d.cpp:2054-2066, :295-310
    const float q[4] = {a.x, a.y, b.x, b.y};
    V2 tmp;
    for (u64 i = 0; i < polylines.size(); i++)
      if (valid_polyline(i)) {
        const auto& c = polylines[i].pc;
        for (size_t k = 1; k < c.size(); k++) {
          const float s1[4] = {c[k].x, c[k].y, c[k - 1].x, c[k - 1].y};
          if (segment_segment(s1, q, tmp)) return true;
        }
      }
    return false;
  }
  void connect_close_extremes() {  // polyline_graph_2d_hmap_impl.cpp:133-160; DIRECT_CONNECTION_EXTREMES_MAXDIST 6
    std::vector<u64> ids; std::vector<V2> pts;
    for (u64 n = 0; n < nodes.size(); n++) if (valid_node(n) && is_extreme(n)) { ids.push_back(n); pts.push_back(nodes[n]); }
    // find_closest_pairs_with_max_dist, polyline_graph_2d.cpp:1315-1350 (reciprocal nearest neighbours)
    const float max_dsq = 6.0f * 6.0f;
    const u64 NONE = ~(u64)0;
    std::vector<u64> closest(pts.size());
    std::vector<std::pair<u64, u64>> pairs;
    for (u64 i = 0; i < pts.size(); i++) {
      float best = std::numeric_limits<float>::max(); u64 bi = NONE;
      for (u64 j = 0; j < pts.size(); j++) if (j != i) { const float d = sqdist(pts[i], pts[j]); if (d < best) { best = d; bi = j; } }
      closest[i] = bi;
      if (bi < i && i == closest[bi] && sqdist(pts[i], pts[bi]) <= max_dsq) pairs.emplace_back(ids[i], ids[bi]);
    }
    std::vector<u64> comp_of; std::vector<std::vector<u64>> comp_nodes;
    compute_components(comp_of, comp_nodes);
    for (const auto& pp : pairs) {
      const u64 a = pp.first, b = pp.second;
      if (comp_of[a] == comp_of[b]) continue;
      if (segment_hits_any_polyline(nodes[a], nodes[b])) continue;
      Poly pl; pl.start = a; pl.end = b; pl.pc = {nodes[a], nodes[b]}; pl.update_length();  // add_direct_connection :125-131
      internal_add_polyline(pl);
      u64 new_id, change;
      if (comp_nodes[comp_of[a]].size() < comp_nodes[comp_of[b]].size()) { new_id = comp_of[b]; change = comp_of[a]; }
      else { new_id = comp_of[a]; change = comp_of[b]; }
      for (u64 n : comp_nodes[change]) comp_of[n] = new_id;
    }
  }
  static float max_smooth_length(const std::vector<V2>& c) {  // polyline_graph_2d.cpp:82-99
    float maxl = 0.0f;
    size_t i = 1;
    while (i < c.size()) {
      float cur = dist(c[i], c[i - 1]);
      for (i++; i < c.size(); i++) {
        const float ax = c[i].x - c[i - 1].x, ay = c[i].y - c[i - 1].y, bx = c[i - 1].x - c[i - 2].x, by = c[i - 1].y - c[i - 2].y;
        const float cosv = (ax * bx + ay * by) / std::sqrt((ax * ax + ay * ay) * (bx * bx + by * by));  // geometric_utilities.cpp:579-588
        if (cosv != 0) cur += dist(c[i], c[i - 1]); else break;   // `if(float)`: NaN counts as true, as here
      }
      maxl = maxl < cur ? cur : maxl;
    }
    return maxl;
  }
  void filter_components_by_smooth_length() {  // polyline_graph_2d.cpp:2011-2052; TOP_FILTER_BY_POLYLINESMOOTHLENGTH 0.82
    const u64 NPL = polylines.size();
    if (NPL == 0) return;
    std::vector<u64> comp_of; std::vector<std::vector<u64>> comp_nodes;
    compute_components(comp_of, comp_nodes);
    std::vector<std::vector<u64>> comp_polys(comp_nodes.size());  // :1994-2009
    for (u64 n = 0; n < nodes.size(); n++)
      if (valid_node(n)) for (u64 pid : connections[n]) comp_polys[comp_of[n]].push_back(pid);
    for (auto& v : comp_polys) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    std::vector<float> sl(NPL, 0.0f);
    for (u64 i = 0; i < NPL; i++) if (valid_polyline(i)) sl[i] = max_smooth_length(polylines[i].pc);
    std::vector<float> sorted = sl;
    const u64 k = (u64)((double)NPL * 0.82);
    std::nth_element(sorted.begin(), sorted.begin() + k, sorted.end());
    const float thr = sorted[k];
    std::vector<uint8_t> drop(comp_polys.size(), 1);
    for (u64 c = 0; c < comp_polys.size(); c++)
      for (u64 pid : comp_polys[c]) if (sl[pid] >= thr) { drop[c] = 0; break; }
    for (u64 c = 0; c < comp_polys.size(); c++)
      if (drop[c]) for (u64 pid : comp_polys[c]) remove_polyline(pid);
  }
};

}  // namespace

struct eg3d_plg {
  std::vector<int64_t> poly_vert_off;
  std::vector<float> verts, poly_length, node_xy, pix_xy;
  std::vector<uint32_t> poly_start, poly_end;
  std::vector<int64_t> pix_adj_off;
  std::vector<uint32_t> pix_adj;
};

extern "C" {

eg3d_status eg3d_plg_from_edge_image(const uint8_t* img, int32_t rows, int32_t cols, int32_t channels, const uint8_t* edge_color,
                                     int32_t stop_after, eg3d_plg** out) {
  if (!img || !out || !edge_color) return eg3d_internal_fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (rows <= 0 || cols <= 0 || channels < 1 || channels > 4) return eg3d_internal_fail(EG3D_ERR_INVALID_ARG, "bad image shape");
  try {
    std::vector<uint8_t> mask((size_t)rows * cols);
    for (size_t k = 0; k < mask.size(); k++) {
      bool e = true;
      for (int c = 0; c < channels; c++) e = e && img[k * channels + c] == edge_color[c];
      mask[k] = e;
    }
    PixelGraph g;
    build_pixel_graph(mask, rows, cols, g);
    Plg plg;
    if (stop_after != EG3D_PLG_STAGE_PIXEL_GRAPH) {
      // convert_EdgeGraph_to_PolyLineGraph, convert_edge_images_pixel_to_segment.cpp:583-626
      std::vector<uint8_t> processed(g.xy.size(), 0);
      for (u64 i = 0; i < g.xy.size(); i++)
        if (!processed[i]) {
          for (const auto& ids : find_polylines(i, g.adj)) {
            const u64 s = ids.front(), e = ids.back();
            plg.get_node_id(g.xy[s]);
            plg.get_node_id(g.xy[e]);
            if (g.adj[s].size() <= 2) processed[s] = 1;
            if (g.adj[e].size() <= 2) processed[e] = 1;
            for (size_t k = 1; k + 1 < ids.size(); k++) processed[ids[k]] = 1;
            std::vector<V2> c; c.reserve(ids.size());
            for (u64 n : ids) c.push_back(g.xy[n]);
            plg.add_polyline(c);
          }
          processed[i] = 1;
        }
      // PolyLineGraph2DHMapImpl::optimize, polyline_graph_2d_hmap_impl.cpp:255-266
      if (stop_after != EG3D_PLG_STAGE_RAW) {
        plg.remove_invalid_polylines();
        plg.remove_degenerate_loops();
        plg.remove_2connection_nodes();
        if (stop_after != EG3D_PLG_STAGE_MERGED) {
          plg.simplify_all();
          if (stop_after != EG3D_PLG_STAGE_SIMPLIFIED) {
            plg.connect_close_extremes();
            plg.simplify_all();
            /* split_loops(): never splits (see the header comment) */
            if (stop_after != EG3D_PLG_STAGE_CONNECTED) plg.filter_components_by_smooth_length();
          }
        }
      }
    }
    eg3d_plg* r = new eg3d_plg();
    r->poly_vert_off.push_back(0);
    for (const Poly& p : plg.polylines) {
      for (const V2& v : p.pc) { r->verts.push_back(v.x); r->verts.push_back(v.y); }
      r->poly_vert_off.push_back((int64_t)r->verts.size() / 2);
      r->poly_start.push_back((uint32_t)p.start); r->poly_end.push_back((uint32_t)p.end); r->poly_length.push_back(p.length);
    }
    for (const V2& v : plg.nodes) { r->node_xy.push_back(v.x); r->node_xy.push_back(v.y); }
    r->pix_adj_off.push_back(0);
    for (size_t n = 0; n < g.xy.size(); n++) {
      r->pix_xy.push_back(g.xy[n].x); r->pix_xy.push_back(g.xy[n].y);
      for (u64 nb : g.adj[n]) r->pix_adj.push_back((uint32_t)nb);
      r->pix_adj_off.push_back((int64_t)r->pix_adj.size());
    }
    *out = r;
    return EG3D_OK;
  } catch (const std::bad_alloc&) {
    return eg3d_internal_fail(EG3D_ERR_OOM, "out of host memory while building the polyline graph");
  }
}

eg3d_status eg3d_plg_get(const eg3d_plg* p, eg3d_plg_view* v) {
  if (!p || !v) return eg3d_internal_fail(EG3D_ERR_INVALID_ARG, "null argument");
  v->n_polylines = (int64_t)p->poly_start.size();
  v->poly_vert_off = p->poly_vert_off.data(); v->verts = p->verts.data();
  v->poly_start = p->poly_start.data(); v->poly_end = p->poly_end.data(); v->poly_length = p->poly_length.data();
  v->n_nodes = (int64_t)p->node_xy.size() / 2; v->node_xy = p->node_xy.data();
  v->n_pixel_nodes = (int64_t)p->pix_xy.size() / 2; v->pixel_node_xy = p->pix_xy.data();
  v->pixel_adj_off = p->pix_adj_off.data(); v->pixel_adj = p->pix_adj.data();
  return EG3D_OK;
}

void eg3d_plg_free(eg3d_plg* p) { delete p; }

}  // extern "C"
