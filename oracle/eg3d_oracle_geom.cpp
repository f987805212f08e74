/*
 * eg3d_oracle_geom.cpp — CPU ORACLE (test infrastructure, NOT product code): geometry, polyline primitives,
 * uniform grid.  See eg3d_oracle.hpp for the parity-pinning statement.  References are to the EdgeGraph3D tree.
 */
#include "eg3d_oracle.hpp"
#include <cmath>
#include <algorithm>
#include <set>

namespace eg3d_oracle {

/* geometric_utilities.cpp:555-557 — float subtraction, double square/sum (pow(float,2) promotes), float result */
float squared_2d_distance(const V2& a, const V2& b) {
  float dx = a.x - b.x, dy = a.y - b.y;
  return (float)((double)dx * (double)dx + (double)dy * (double)dy);
}
/* geometric_utilities.cpp:571-573 */
float compute_2d_distance(const V2& a, const V2& b) { return std::sqrt(squared_2d_distance(a, b)); }

/* geometric_utilities.cpp:272-312.  segm = (x1,y1,x2,y2); t measured from (x1,y1); both ends inclusive. */
void intersect_segment_line(const float segm[4], const V3& line, bool& found, V2& inter) {
  float dx = segm[2] - segm[0], dy = segm[3] - segm[1];
  found = false;
  float num = line.x * segm[0] + line.y * segm[1] + line.z;
  float den = line.x * dx + line.y * dy;
  if (den != 0) {
    float t = -num / den;
    if (t >= 0 && t <= 1) {
      inter.x = segm[0] + t * dx;
      inter.y = segm[1] + t * dy;
      found = true;
    }
  }
}

/* geometric_utilities.cpp:608-614 */
static V2 get_2d_direction_line(const V3& line) {
  if (line.y == 0) return V2{0.0f, 1.0f};
  return V2{1.0f, -line.x / line.y};
}
/* geometric_utilities.cpp:579-581 (signed cosine) */
static float compute_anglecos_vec2_vec2(const V2& a, const V2& b) {
  float ab = a.x * b.x + a.y * b.y;
  float aa = a.x * a.x + a.y * a.y;
  float bb = b.x * b.x + b.y * b.y;
  return ab / std::sqrt(aa * bb);
}
/* geometric_utilities.cpp:997-1001, 1007-1009 */
static float distance_point_line(const V2& p, const V3& line) {
  float den = line.x * p.x + line.y * p.y + line.z;
  den *= den;
  return std::sqrt(den / (line.x * line.x + line.y * line.y));
}

/* geometric_utilities.cpp:365-430 */
void intersect_segment_line_no_quasiparallel(const float segm[4], const V3& line, float max_cos, float max_dist,
                                             bool& found, bool& quasiparallel_within_distance, V2& inter) {
  float dx = segm[2] - segm[0], dy = segm[3] - segm[1];
  found = false;
  quasiparallel_within_distance = false;
  float distance;
  float num = line.x * segm[0] + line.y * segm[1] + line.z;
  float den = line.x * dx + line.y * dy;
  if (den != 0) {
    float t = -num / den;
    if (t >= 0 && t <= 1) {
      inter.x = segm[0] + t * dx;
      inter.y = segm[1] + t * dy;
      distance = 0;
      found = true;
    }
    if (compute_anglecos_vec2_vec2(V2{dx, dy}, get_2d_direction_line(line)) > max_cos) {
      if (t < 0) distance = distance_point_line(V2{segm[0], segm[1]}, line);
      else if (t > 1) distance = distance_point_line(V2{segm[2], segm[3]}, line);
      else distance = 0;
      if (distance <= max_dist) quasiparallel_within_distance = true;
    }
  } else {
    distance = distance_point_line(V2{segm[0], segm[1]}, line);
    if (distance <= max_dist) quasiparallel_within_distance = true;
  }
}

/* geometric_utilities.cpp:940-954 */
float minimum_distancesq(const V2& p, const V2& v, const V2& w, V2& projection) {
  const float l2 = squared_2d_distance(v, w);
  if (l2 == 0.0) { projection = v; return squared_2d_distance(p, v); }
  float pvx = p.x - v.x, pvy = p.y - v.y, wvx = w.x - v.x, wvy = w.y - v.y;
  float q = (pvx * wvx + pvy * wvy) / l2;
  const float t = std::max<float>(0, std::min<float>(1, q));
  projection.x = v.x + t * wvx;
  projection.y = v.y + t * wvy;
  return squared_2d_distance(p, projection);
}

/* geometric_utilities.cpp:824-843 -> cv::computeCorrespondEpilines(whichImage=1) (pinned vs cv2) */
bool computeCorrespondEpilineSinglePoint(const V2& p, const double* F, bool Fvalid, V3& line) {
  if (!Fvalid) return false;
  double x = p.x, y = p.y;
  double a = F[0] * x + F[1] * y + F[2];
  double b = F[3] * x + F[4] * y + F[5];
  double c = F[6] * x + F[7] * y + F[8];
  double nu = a * a + b * b;
  nu = nu ? 1. / std::sqrt(nu) : 1.;
  line.x = (float)(a * nu); line.y = (float)(b * nu); line.z = (float)(c * nu);
  return true;
}

/* geometric_utilities.cpp:973-983: vec4(X,1) * cameraMatrix (glm row-vector product, type_mat4x4.inl:640-651) */
V2 compute_projection(const float P[12], const V3& X) {
  float h0 = P[0] * X.x + P[1] * X.y + P[2] * X.z + P[3] * 1.0f;
  float h1 = P[4] * X.x + P[5] * X.y + P[6] * X.z + P[7] * 1.0f;
  float h2 = P[8] * X.x + P[9] * X.y + P[10] * X.z + P[11] * 1.0f;
  return V2{h0 / h2, h1 / h2};
}

/* geometric_utilities.cpp:1370-1372 */
static V2 first_plus_ratio_of_segment(const V2& a, const V2& b, float ratio) {
  return V2{a.x + ratio * (b.x - a.x), a.y + ratio * (b.y - a.y)};
}

thread_local long long g_foreign_direction_calls = 0;
/* ------------------------------------------------------------------ polyline ---- */
PlPoint Polyline::get_start_plp() const { return PlPoint{0, pc[0]}; } /* polyline_graph_2d.cpp:134-136 */

/* polyline_graph_2d.cpp:391-447.  A direction that is neither start nor end is undefined behaviour in the
 * reference (falls off the function); oracle rule: "cannot drive" = reached extreme. */
PlPoint Polyline::next_pl_point_by_distance(const PlPoint& init, ulong_t direction, float distance, bool& reached) const {
  float prevdist = 0, curdist, ratio;
  reached = false;
  ulong_t i;
  const ulong_t n = pc.size();
  if (direction == start) {
    curdist = compute_2d_distance(pc[init.seg], init.c);
    if (curdist >= distance) {
      ratio = distance / curdist;
      return PlPoint{init.seg, first_plus_ratio_of_segment(init.c, pc[init.seg], ratio)};
    }
    for (i = init.seg; i > 0; i--) {
      prevdist = curdist;
      curdist = compute_2d_distance(pc[i - 1], init.c);
      if (curdist >= distance) break;
    }
    if (i == 0) { reached = true; return PlPoint{0, pc[0]}; }
    ratio = (distance - prevdist) / (curdist - prevdist);
    return PlPoint{i - 1, first_plus_ratio_of_segment(pc[i], pc[i - 1], ratio)};
  } else if (direction == end) {
    if (init.seg >= n - 1) { reached = true; return PlPoint{n - 2, pc[n - 1]}; }
    curdist = compute_2d_distance(pc[init.seg + 1], init.c);
    if (curdist >= distance) {
      ratio = distance / curdist;
      return PlPoint{init.seg, first_plus_ratio_of_segment(init.c, pc[init.seg + 1], ratio)};
    }
    for (i = init.seg + 1; i < n - 1; i++) {
      prevdist = curdist;
      curdist = compute_2d_distance(pc[i + 1], init.c);
      if (curdist >= distance) break;
    }
    if (i == n - 1) { reached = true; return PlPoint{n - 2, pc[n - 1]}; }
    ratio = (distance - prevdist) / (curdist - prevdist);
    return PlPoint{i, first_plus_ratio_of_segment(pc[i], pc[i + 1], ratio)};
  }
  g_foreign_direction_calls++;   /* the reference's behaviour is undefined here: tests that execute the reference's own code skip such seeds */
  reached = true;
  return init;
}

/* polyline_graph_2d.cpp:579-664.  Quasi-parallel test comes BEFORE accepting an intersection (:593-602). */
void Polyline::next_pl_point_by_line_intersection(const PlPoint& init, ulong_t direction, const V3& line, float qcos, float qdist,
                                                  PlPoint& next, bool& found) const {
  found = false;
  bool inter_found, qp;
  V2 inter{0, 0};
  float segm[4];
  const ulong_t n = pc.size();
  if (direction == start) {
    segm[0] = init.c.x; segm[1] = init.c.y; segm[2] = pc[init.seg].x; segm[3] = pc[init.seg].y;
    intersect_segment_line_no_quasiparallel(segm, line, qcos, qdist, inter_found, qp, inter);
    if (qp) return;
    if (inter_found) { next = PlPoint{init.seg, inter}; found = true; return; }
    for (ulong_t i = init.seg; i > 0; i--) {
      segm[0] = pc[i].x; segm[1] = pc[i].y; segm[2] = pc[i - 1].x; segm[3] = pc[i - 1].y;
      intersect_segment_line_no_quasiparallel(segm, line, qcos, qdist, inter_found, qp, inter);
      if (qp) return;
      if (inter_found) { next = PlPoint{i - 1, inter}; found = true; return; }
    }
    return;
  } else if (direction == end) {
    segm[0] = init.c.x; segm[1] = init.c.y; segm[2] = pc[init.seg + 1].x; segm[3] = pc[init.seg + 1].y;
    intersect_segment_line_no_quasiparallel(segm, line, qcos, qdist, inter_found, qp, inter);
    if (qp) return;
    if (inter_found) { next = PlPoint{init.seg, inter}; found = true; return; }
    for (ulong_t i = init.seg + 1; i < n - 1; i++) {
      segm[0] = pc[i].x; segm[1] = pc[i].y; segm[2] = pc[i + 1].x; segm[3] = pc[i + 1].y;
      intersect_segment_line_no_quasiparallel(segm, line, qcos, qdist, inter_found, qp, inter);
      if (qp) return;
      if (inter_found) { next = PlPoint{i, inter}; found = true; return; }
    }
    return;
  }
  /* neither start nor end: flags stay cleared (polyline_graph_2d.cpp:586-588, 662-663) */
}

/* polyline_graph_2d.cpp:666-780 */
void Polyline::next_pl_point_by_line_intersection_bounded_distance(const PlPoint& init, ulong_t direction, const V3& line,
                                                                   float qcos, float qdist, float min_dist, float max_dist,
                                                                   PlPoint& next, bool& found) const {
  found = false;
  bool inter_found, qp;
  V2 inter{0, 0};
  float segm[4];
  const ulong_t n = pc.size();
  auto bounded = [&](const V2& p) {
    float dsq = squared_2d_distance(p, init.c);
    return !(dsq < (min_dist * min_dist) || dsq > (max_dist * max_dist));
  };
  if (direction == start) {
    segm[0] = init.c.x; segm[1] = init.c.y; segm[2] = pc[init.seg].x; segm[3] = pc[init.seg].y;
    intersect_segment_line_no_quasiparallel(segm, line, qcos, qdist, inter_found, qp, inter);
    if (qp) return;
    if (inter_found) { next = PlPoint{init.seg, inter}; found = bounded(inter); return; }
    for (ulong_t i = init.seg; i > 0; i--) {
      segm[0] = pc[i].x; segm[1] = pc[i].y; segm[2] = pc[i - 1].x; segm[3] = pc[i - 1].y;
      intersect_segment_line_no_quasiparallel(segm, line, qcos, qdist, inter_found, qp, inter);
      if (qp) return;
      if (inter_found) { next = PlPoint{i - 1, inter}; found = bounded(inter); return; }
    }
    return;
  } else if (direction == end) {
    segm[0] = init.c.x; segm[1] = init.c.y; segm[2] = pc[init.seg + 1].x; segm[3] = pc[init.seg + 1].y;
    intersect_segment_line_no_quasiparallel(segm, line, qcos, qdist, inter_found, qp, inter);
    if (qp) return;
    if (inter_found) { next = PlPoint{init.seg, inter}; found = bounded(inter); return; }
    for (ulong_t i = init.seg + 1; i < n - 1; i++) {
      segm[0] = pc[i].x; segm[1] = pc[i].y; segm[2] = pc[i + 1].x; segm[3] = pc[i + 1].y;
      intersect_segment_line_no_quasiparallel(segm, line, qcos, qdist, inter_found, qp, inter);
      if (qp) return;
      if (inter_found) { next = PlPoint{i, inter}; found = bounded(inter); return; }
    }
    return;
  }
}

/* polyline_graph_2d.cpp:312-327: segments are passed reversed (P[i], P[i-1]); segment_index = i-1 */
std::vector<PlPoint> Polyline::intersect_line(const V3& line) const {
  std::vector<PlPoint> res;
  bool found; V2 inter;
  float segm[4];
  for (ulong_t i = 1; i < pc.size(); i++) {
    segm[0] = pc[i].x; segm[1] = pc[i].y; segm[2] = pc[i - 1].x; segm[3] = pc[i - 1].y;
    intersect_segment_line(segm, line, found, inter);
    if (found) res.push_back(PlPoint{i - 1, inter});
  }
  return res;
}

/* polyline_graph_2d.cpp:845-862 */
float Polyline::compute_distancesq(const V2& p, ulong_t& closest_segm, V2& projection) const {
  float min_dist = minimum_distancesq(p, pc[0], pc[1], projection);
  closest_segm = 0;
  V2 cur_projection;
  for (ulong_t i = 2; i < pc.size(); i++) {
    float cur = minimum_distancesq(p, pc[i - 1], pc[i], cur_projection);
    if (cur < min_dist) { min_dist = cur; projection = cur_projection; closest_segm = i - 1; }
  }
  return min_dist;
}

/* edge_graph_3d_utilities.cpp:600-629 */
static float floor_or_upper_if_close(float v) {
  if (std::ceil(v) - v < 0.001) return std::ceil(v);
  return std::floor(v);
}
static bool is_m_multiple_of_n_float(float m, float n) {
  float div = m / n;
  float mul = floor_or_upper_if_close(div) * n;
  return std::abs(m - mul) < 0.001;
}
static std::pair<ulong_t, ulong_t> cell_from_coords(float cell_dim, const V2& c) {
  return std::make_pair((ulong_t)(int)floor_or_upper_if_close(c.x / cell_dim), (ulong_t)(int)floor_or_upper_if_close(c.y / cell_dim));
}

/* polyline_graph_2d.cpp:798-835 (+ :568-577, :551-562): sample every cell/(1.414+0.1) by Euclidean stepping from the
 * start extreme, drop samples lying on a cell boundary, collect the set of cells. */
std::vector<std::pair<ulong_t, ulong_t>> Polyline::get_intersectedcells_2dmap_set(float cell_dim) const {
  std::set<std::pair<ulong_t, ulong_t>> cells;
  const float step = (float)(cell_dim / (1.414 + 0.1));
  /* split_equal_size_intervals(start, step): direction = get_other_end(start) */
  ulong_t direction = end; /* get_other_end(start): extreme == start -> end (polyline_graph_2d.cpp:901-908) */
  std::vector<PlPoint> plps;
  plps.push_back(get_start_plp());
  PlPoint next = plps[0];
  bool reached = false;
  while (!reached) {
    next = next_pl_point_by_distance(next, direction, step, reached);
    plps.push_back(next);
  }
  for (const auto& plp : plps) {
    bool on_boundary = is_m_multiple_of_n_float(plp.c.x, cell_dim) || is_m_multiple_of_n_float(plp.c.y, cell_dim);
    if (!on_boundary) cells.insert(cell_from_coords(cell_dim, plp.c));
  }
  return std::vector<std::pair<ulong_t, ulong_t>>(cells.begin(), cells.end());
}

/* ------------------------------------------------------------------ grid ---- */
/* polyLine_2d_map.cpp:40-58.  Cells outside the map are an out-of-bounds write in the reference; the oracle
 * drops them (scenes under test keep polylines strictly inside the image). */
void Grid::build(const std::vector<Polyline>& pls, int iw, int ih, float cell) {
  cell_dim = cell; img_w = iw; img_h = ih;
  w = (int)std::ceil(iw / cell); h = (int)std::ceil(ih / cell);
  cells.assign((size_t)w * h, {});
  for (ulong_t id = 0; id < pls.size(); id++) {
    if (!pls[id].valid()) continue;
    for (const auto& c : pls[id].get_intersectedcells_2dmap_set(cell))
      if ((long long)c.first < w && (long long)c.second < h) cells[(size_t)c.second * w + c.first].push_back(id);
  }
}

/* polyLine_2d_map_search.cpp:46-77.  NB the "row" flag (x on a boundary) clips the ROW (y) loop. */
std::vector<ulong_t> Grid::find_polylines_potentially_within_search_dist(const V2& c) const {
  std::set<ulong_t> res;
  if (c.x <= 0 || c.x >= img_w || c.y <= 0 || c.y >= img_h) return {};
  bool on_row = is_m_multiple_of_n_float(c.x, cell_dim);
  bool on_col = is_m_multiple_of_n_float(c.y, cell_dim);
  auto cc = cell_from_coords(cell_dim, c);
  long long cx = (long long)cc.first >= w ? w - 1 : (long long)cc.first;
  long long cy = (long long)cc.second >= h ? h - 1 : (long long)cc.second;
  int i0 = cy > 0 ? -1 : 0, i1 = on_row ? 0 : (cy < h - 1 ? 1 : 0);
  int j0 = cx > 0 ? -1 : 0, j1 = on_col ? 0 : (cx < w - 1 ? 1 : 0);
  for (int i = i0; i <= i1; i++)
    for (int j = j0; j <= j1; j++)
      for (ulong_t id : cells[(size_t)(cy + i) * w + (cx + j)]) res.insert(id);
  return std::vector<ulong_t>(res.begin(), res.end());
}

/* polyLine_2d_map_search.cpp:81-88 */
void Grid::find_unique_polyline_potentially_within_search_dist(const V2& c, ulong_t& pl_id, bool& valid) const {
  valid = false;
  auto r = find_polylines_potentially_within_search_dist(c);
  if (r.size() == 1) { valid = true; pl_id = r[0]; }
}

}  // namespace eg3d_oracle
