/*
 * eg3d_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain C++17 restatement of the reference algorithm for EdgeGraph3D's epipolar polyline matching +
 * Gauss-Newton triangulation path.  Every function cites the reference
 * file:line it follows (paths relative to the reference root).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the product (edgegraph3d_b200/) never does.
 *
 * PARITY PINNING: the reference ships no tests or golden vectors and cannot be compiled in the build
 * container (OpenCV/CGAL/Boost/Eigen headers absent), so the oracle is pinned at the only third-party
 * boundary the path has — OpenCV — against fixtures produced by real `cv2` calls
 * (tests/golden/make_golden_cv2.py: computeCorrespondEpilines, triangulatePoints, and the two Gauss-Newton
 * loops rebuilt from cv2.gemm / cv2.invert / cv2.determinant).  Everything else (control flow of the
 * matching path) is "parity unpinned" by the reference itself and rests on this restatement.
 *
 * Built with -O3 -ffp-contract=off so float/double expressions round exactly as written (the reference's
 * x86-64 build has no FMA contraction either).
 */
#pragma once
#include <cstdint>
#include <vector>
#include <array>
#include "../include/eg3d.h"

namespace eg3d_oracle {

typedef uint64_t ulong_t;

struct V2 { float x, y; };
struct V3 { float x, y, z; };
inline bool operator==(const V2& a, const V2& b) { return a.x == b.x && a.y == b.y; }

/* PolyLineGraph2D::polyline::pl_point (polyline_graph_2d.hpp:87-99) */
struct PlPoint { ulong_t seg; V2 c; };
/* PolyLineGraph2D::plg_point (polyline_graph_2d.hpp:278-294) */
struct PlgPoint { ulong_t pl; PlPoint plp; };
/* new_3dpoint_plgp_matches (polyline_graph_2d.hpp:451) */
struct Match { V3 X; std::vector<PlgPoint> obs; std::vector<int> views; };

/* PolyLineGraph2D::polyline (polyline_graph_2d.hpp:85-119): only what the hot path reads. */
struct Polyline {
  ulong_t start = 0, end = 0;
  std::vector<V2> pc;
  bool valid() const { return pc.size() > 1; } /* polyline_graph_2d.cpp:1141-1147 */
  PlPoint get_start_plp() const;
  PlPoint next_pl_point_by_distance(const PlPoint& init, ulong_t direction, float distance, bool& reached_extreme) const;
  void next_pl_point_by_line_intersection(const PlPoint& init, ulong_t direction, const V3& line, float qcos, float qdist,
                                          PlPoint& next, bool& found) const;
  void next_pl_point_by_line_intersection_bounded_distance(const PlPoint& init, ulong_t direction, const V3& line,
                                                            float qcos, float qdist, float min_dist, float max_dist,
                                                            PlPoint& next, bool& found) const;
  std::vector<PlPoint> intersect_line(const V3& line) const;
  float compute_distancesq(const V2& p, ulong_t& closest_segm, V2& projection) const;
  std::vector<std::pair<ulong_t, ulong_t>> get_intersectedcells_2dmap_set(float cell_dim) const; /* sorted unique (col,row) */
};

/* PolyLine2DMapSearch (polyline_2d_map_search.hpp:47-63, polyLine_2d_map.cpp:40-58) */
struct Grid {
  float cell_dim = 0; int w = 0, h = 0; int img_w = 0, img_h = 0;
  std::vector<std::vector<ulong_t>> cells; /* [row*w + col] */
  void build(const std::vector<Polyline>& pls, int img_w, int img_h, float cell);
  std::vector<ulong_t> find_polylines_potentially_within_search_dist(const V2& c) const; /* sorted unique */
  void find_unique_polyline_potentially_within_search_dist(const V2& c, ulong_t& pl_id, bool& valid) const;
};

struct Scene {
  int V = 0, width = 0, height = 0;
  eg3d_params prm;
  std::vector<std::array<float, 12>> P;        /* rows 0..2 of cameraMatrix */
  std::vector<double> F; std::vector<uint8_t> Fvalid;
  std::vector<std::vector<Polyline>> plgs;     /* [view][polyline] */
  std::vector<Grid> plmaps;                    /* 4 px, edge_matcher.cpp:101-103 */
  std::vector<Grid> corr_maps;                 /* 30 px, plg_edge_manager.cpp:73-74 */
  std::vector<V3> points; std::vector<std::vector<int>> track_views; std::vector<std::vector<V2>> track_xy;
  const double* Fm(int a, int b) const { return &F[((size_t)a * V + b) * 9]; }
  bool Fok(int a, int b) const { return Fvalid[(size_t)a * V + b] != 0; }
};

/* --- geometry (geometric_utilities.cpp) --- */
float squared_2d_distance(const V2& a, const V2& b);
float compute_2d_distance(const V2& a, const V2& b);
void intersect_segment_line(const float segm[4], const V3& line, bool& found, V2& inter);
void intersect_segment_line_no_quasiparallel(const float segm[4], const V3& line, float max_cos, float max_dist,
                                             bool& found, bool& quasiparallel_within_distance, V2& inter);
float minimum_distancesq(const V2& p, const V2& v, const V2& w, V2& projection);
bool computeCorrespondEpilineSinglePoint(const V2& p, const double* F9, bool Fvalid, V3& line);
V2 compute_projection(const float P12[12], const V3& X);

/* --- triangulation (triangulation.cpp) --- */
extern thread_local int g_trace;
extern thread_local long long g_foreign_direction_calls;   /* calls of next_pl_point_by_distance with a direction that is neither extreme (UB in the reference) */
void triangulate_dlt(const float P1[12], const float P2[12], const V2& x1, const V2& x2, float out4[4]);
void triangulate_dlt_opencv(const float P1[12], const float P2[12], const V2& x1, const V2& x2, float out4[4]);
int em_GaussNewton(const Scene& s, const std::vector<int>& views, const std::vector<V2>& pts, const double init[3],
                   double out[3], double* last_mse_out);
void em_estimate3Dpositions(const Scene& s, const std::vector<V2>& coords, const std::vector<int>& ids, V3& X, bool& valid);
void em_add_new_observation_to_3Dpositions(const Scene& s, const V3& X0, const std::vector<V2>& coords, const std::vector<int>& ids,
                                           const V2& new_coords, int new_view, V3& X, bool& valid);

} // namespace eg3d_oracle
#include <atomic>
namespace eg3d_oracle {
extern std::atomic<long long> g_dlt_calls, g_dlt_degenerate; /* test statistics, eg3d_oracle_match.cpp */

/* --- the per-seed path --- */
std::vector<std::vector<PlgPoint>> find_epipolar_correspondences(const Scene& s, const std::vector<std::vector<ulong_t>>* cand,
                                                                 int starting_plg_id, const PlgPoint& starting_plgp);
std::vector<Match> compute_3D_point_multiple_views_plg_following_expandallviews_vector(
    const Scene& s, int starting_plg_id, const std::vector<std::vector<PlgPoint>>& epipolar_correspondences);

/* --- filtering --- */
int GaussNewton_f32(const Scene& s, const std::vector<int>& views, const std::vector<V2>& pts, const float init[3],
                    float out[3], float gn_max_mse, float* last_mse_out);

Scene* scene_from_desc(const eg3d_scene_desc* d, const eg3d_params* p);

}  // namespace eg3d_oracle
