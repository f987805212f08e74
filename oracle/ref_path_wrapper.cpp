/*
 * ref_path_wrapper.cpp — TEST INFRASTRUCTURE: a flat C interface over the REFERENCE'S OWN hot-path code, compiled unmodified
 * from /root/reference (oracle/Makefile, target _ref/libref_path.so; never copied into this repository):
 *   plgs/polyline_graph_2d.cpp, polyline_graph_2d_hmap_impl.cpp, polyline_graph_3d*.cpp, graph*.cpp
 *   matching/plg_matching/{polyline_matching,plg_matching,plg_matching_from_refpoints,plg_matches_manager,polyLine_2d_map,
 *   polyLine_2d_map_search}.cpp, matching/consensus_manager/*.cpp, edge_managers/plg_edge_manager*.cpp,
 *   utils/geometry/{triangulation,geometric_utilities}.cpp, utils/{edge_graph_3d_utilities,datatypes}.cpp,
 *   filtering/{gauss_newton,outliers_filtering,filtering_close_plgps}.cpp, external/manifoldReconstructor/src/OpenMvgParser.cpp
 * against the stand-in headers of oracle/ref_stubs_path/ (OpenCV, CGAL, Boost are not in this image).  The control flow of
 * matching, PLG following, view expansion, the density limiter and the outlier filter that runs here IS the reference's; the
 * OpenCV arithmetic underneath is the model the cv2 goldens pinned (see eg3d_cv_stub.hpp).  tests/test_ref_path.py compares the
 * oracle (oracle/libeg3d_oracle.so) with this library; the product never loads either.
 *
 * Only the reference's compile-time constants are available here (SPLIT_INTERVAL_DISTANCE = 20, ... ), i.e. eg3d_params defaults.
 */
#include <omp.h>
#include <iostream>
#include <set>
#include <tuple>
#include <vector>

#include "SfMData.h"
#include "glm.hpp"
#include "polyline_graph_2d_hmap_impl.hpp"
#include "polyline_graph_3d_hmap_impl.hpp"
#include "polyline_2d_map_search.hpp"
#include "plg_matches_manager.hpp"
#include "polyline_matching.hpp"
#include "plg_matching_from_refpoints.hpp"
#include "plg_edge_manager.hpp"
#include "plgpcm_3views_plg_following.hpp"
#include "triangulation.hpp"
#include "filtering_close_plgps.hpp"
#include "outliers_filtering.hpp"
#include "gauss_newton.hpp"
#include "geometric_utilities.hpp"
#include "edge_graph_3d_utilities.hpp"
#include "global_defines.hpp"
#include "OpenMvgParser.h"

#include "../include/eg3d.h"

/* the oracle's bit-exact restatements of the two OpenCV calls of the path (pinned against real cv2, tests/test_oracle_golden.py) */
namespace eg3d_oracle {
struct V2 { float x, y; };
struct V3 { float x, y, z; };
void triangulate_dlt_opencv(const float P1[12], const float P2[12], const V2& x1, const V2& x2, float out4[4]);
bool computeCorrespondEpilineSinglePoint(const V2& p, const double* F9, bool Fvalid, V3& line);
}

namespace cv {
/* cv::triangulatePoints(P1, P2, x1, x2, out): 3x4 CV_32F cameras, one point each (triangulation.cpp:216,290) */
void triangulatePoints(const Mat& P1, const Mat& P2, const std::vector<Point2f>& x1, const std::vector<Point2f>& x2, Vec4f& out) {
  float a[12], b[12];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) { a[4 * r + c] = P1.at<float>(r, c); b[4 * r + c] = P2.at<float>(r, c); }
  float o[4];
  eg3d_oracle::triangulate_dlt_opencv(a, b, eg3d_oracle::V2{x1[0].x, x1[0].y}, eg3d_oracle::V2{x2[0].x, x2[0].y}, o);
  for (int k = 0; k < 4; k++) out[k] = o[k];
}
/* cv::computeCorrespondEpilines(points, 1, F, lines) (geometric_utilities.cpp:832); F is 3x3 CV_64F */
void computeCorrespondEpilines(const Mat& points, int whichImage, const Mat& F, std::vector<Vec3f>& lines) {
  if (whichImage != 1 || F.depth() != CV_64F) eg3d_stub_unreachable("computeCorrespondEpilines(whichImage != 1 or non-double F)");
  double f[9]; for (int i = 0; i < 9; i++) f[i] = F.at<double>(i / 3, i % 3);
  lines.clear();
  for (int i = 0; i < points.rows; i++) {
    eg3d_oracle::V3 l;
    eg3d_oracle::computeCorrespondEpilineSinglePoint(eg3d_oracle::V2{points.at<float>(2 * i), points.at<float>(2 * i + 1)}, f, true, l);
    lines.push_back(Vec3f(l.x, l.y, l.z));
  }
}
}  // namespace cv

/* ONE GUARD around the reference's code.  PolyLineGraph2D::polyline::next_pl_point_by_distance (polyline_graph_2d.cpp:391-447) falls
 * off its end without a return statement when `direction` is neither extreme of the polyline, and the path does call it that way
 * (zero-filled per-view direction arrays; ~5-10 % of the productive seeds): undefined behaviour — garbage in the
 * author's build, a crash in this one.  oracle/Makefile compiles polyline_graph_2d.cpp with
 * -Dnext_pl_point_by_distance=next_pl_point_by_distance_refimpl, so the reference's own two overloads keep their bodies under
 * that name, and the symbols every OTHER translation unit calls are defined here: the undefined case gets the rule the oracle
 * and the product use ("cannot drive": reached_polyline_extreme = true), every defined case is forwarded to the reference's
 * code untouched. */
PolyLineGraph2D::polyline::pl_point eg3d_refimpl_next_a(PolyLineGraph2D::polyline* self, const PolyLineGraph2D::polyline::pl_point init_plp, const ulong direction,
                                                        const float distance, bool& reached)
    asm("_ZN15PolyLineGraph2D8polyline33next_pl_point_by_distance_refimplENS0_8pl_pointEmfRb");
PolyLineGraph2D::polyline::pl_point eg3d_refimpl_next_b(PolyLineGraph2D::polyline* self, const ulong starting_extreme, const glm::vec2 coords, const float distance, bool& reached)
    asm("_ZN15PolyLineGraph2D8polyline33next_pl_point_by_distance_refimplEmN3glm5tvec2IfLNS1_9precisionE0EEEfRb");
PolyLineGraph2D::polyline::pl_point PolyLineGraph2D::polyline::next_pl_point_by_distance(const PolyLineGraph2D::polyline::pl_point init_plp, const ulong direction,
                                                                                        const float distance, bool& reached_polyline_extreme) {
  if (direction != start && direction != end) { reached_polyline_extreme = true; return init_plp; }
  return eg3d_refimpl_next_a(this, init_plp, direction, distance, reached_polyline_extreme);
}
PolyLineGraph2D::polyline::pl_point PolyLineGraph2D::polyline::next_pl_point_by_distance(const ulong starting_extreme, const glm::vec2 coords, const float distance,
                                                                                        bool& reached_polyline_extreme) {
  return eg3d_refimpl_next_b(this, starting_extreme, coords, distance, reached_polyline_extreme);
}

/* debug drawing (src/edgegraph3d/utils/drawing_utilities.cpp, OpenCV imgproc) is behind the reference's -i switch and never
 * reached here: the symbols its callers name are defined as traps */
#define EG3D_TRAP(name) { fprintf(stderr, "ref_path_wrapper: " name " (debug drawing) is not part of this build\n"); abort(); }
typedef std::vector<std::pair<PolyLineGraph2D::plg_point, std::vector<std::vector<PolyLineGraph2D::plg_point>>>> eg3d_iac_t;
void draw_single_point_process_no_epilines(std::vector<cv::Mat>&, const SfMData&, cv::Mat**, int, int, int, const eg3d_iac_t&) EG3D_TRAP("draw_single_point_process_no_epilines")
void draw_single_point_process(std::vector<cv::Mat>&, const SfMData&, cv::Mat**, int, int, int, const eg3d_iac_t&) EG3D_TRAP("draw_single_point_process")
void draw_overlay_MultiColorComponents_PolyLineGraph_simplified(cv::Mat&, const PolyLineGraph2D&) EG3D_TRAP("draw_overlay_MultiColorComponents_PolyLineGraph_simplified")
cv::Mat draw_MultiColorComponents_PolyLineGraph_simplified(const cv::Mat&, const PolyLineGraph2D&) EG3D_TRAP("draw_MultiColorComponents_PolyLineGraph_simplified")
void draw_3dpoints_on_imgs(std::vector<cv::Mat>&, const std::vector<std::tuple<glm::vec3, std::vector<glm::vec2>, std::vector<int>>>&) EG3D_TRAP("draw_3dpoints_on_imgs")

/* functions of the reference that its headers do not declare (defined in the translation units named) */
vector<bool> compute_inliers(SfMData& sfmd, const int first_edgepoint, const float gn_max_mse, const int forced_min_filter);  /* outliers_filtering.cpp:37 */
vector<vector<PolyLineGraph2D::plg_point>> find_epipolar_correspondences(const SfMData& sfmd, const vector<PolyLineGraph2DHMapImpl>& plgs, const Mat** all_fundamental_matrices,
    const vector<set<ulong>>& potentially_compatible_polylines, const int starting_plg_id, const PolyLineGraph2D::plg_point& starting_plgp, PLGMatchesManager& plgmm);   /* polyline_matching.cpp:45 */
vector<std::tuple<glm::vec3, vector<PolyLineGraph2D::plg_point>, vector<int>>> find_new_3d_points_from_compatible_polylines_starting_plgp_expandallviews(
    const SfMData& sfmd, const vector<PolyLineGraph2DHMapImpl>& plgs, const Mat** all_fundamental_matrices, const vector<set<ulong>>& potentially_compatible_polylines,
    const int starting_plg_id, const PolyLineGraph2D::plg_point& starting_plgp, PLGMatchesManager& plgmm, const vector<PolyLine2DMapSearch>& plmaps);                  /* polyline_matching.cpp:134 */
#define EG3D_INVALID_FORCED_MIN_FILTER -1   /* INVALID_FORCED_MIN_FILTER, outliers_filtering.cpp:12 (file-local there) */

typedef std::tuple<glm::vec3, vector<PolyLineGraph2D::plg_point>, vector<int>> p3d_t;

namespace {
struct RefScene {
  int V = 0;
  SfMData sfmd;
  vector<PolyLineGraph2DHMapImpl> plgs;
  cv::Mat** F = nullptr;
  vector<PolyLine2DMapSearch> plmaps;              /* 4 px, edge_matcher.cpp:101-103 */
  vector<cv::Mat> imgs;                            /* headers only: the path reads their size */
  PolyLineGraph3DHMapImpl plg3d;
  PLGMatchesManager* plgmm = nullptr;
  PLGEdgeManager* em = nullptr;
  PLGPCM3ViewsPLGFollowing* cm = nullptr;
  std::string error;
};
struct RefPoints {
  std::vector<float> xyz; std::vector<int32_t> seed, chain_pos; std::vector<int64_t> obs_off{0};
  std::vector<int32_t> obs_view; std::vector<uint32_t> obs_poly, obs_seg; std::vector<float> obs_xy;
  void append(const vector<p3d_t>& chain, int32_t seed_ord) {
    for (size_t k = 0; k < chain.size(); k++) {
      const glm::vec3& X = std::get<0>(chain[k]);
      xyz.push_back(X[0]); xyz.push_back(X[1]); xyz.push_back(X[2]);
      seed.push_back(seed_ord); chain_pos.push_back((int32_t)k);
      const auto& obs = std::get<1>(chain[k]); const auto& views = std::get<2>(chain[k]);
      for (size_t j = 0; j < obs.size(); j++) {
        obs_view.push_back(views[j]); obs_poly.push_back((uint32_t)obs[j].polyline_id); obs_seg.push_back((uint32_t)obs[j].plp.segment_index);
        obs_xy.push_back(obs[j].plp.coords[0]); obs_xy.push_back(obs[j].plp.coords[1]);
      }
      obs_off.push_back((int64_t)obs_view.size());
    }
  }
};
struct Quiet {   /* the reference prints progress lines from its loops */
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(nullptr)) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
vector<set<ulong>> cand_of(const RefScene& s, const eg3d_candidates* c, int set_id) {
  vector<set<ulong>> r(s.V);
  for (int v = 0; v < s.V; v++)
    for (int64_t k = c->off[(size_t)set_id * s.V + v]; k < c->off[(size_t)set_id * s.V + v + 1]; k++) r[v].insert((ulong)c->polyline[k]);
  return r;
}
}  // namespace

extern "C" {

const char* eg3d_ref_last_error(void* sc) { return ((RefScene*)sc)->error.c_str(); }

/* SfMData + PolyLineGraph2DHMapImpl + Mat** F + plmaps from the flat scene (the inverse of what include/eg3d_ref_api.hpp does) */
void* eg3d_ref_scene_create(const eg3d_scene_desc* d) {
  Quiet q;
  RefScene* s = new RefScene();
  const int V = s->V = d->n_views;
  SfMData& m = s->sfmd;
  m.numCameras_ = V; m.imageWidth_ = d->width; m.imageHeight_ = d->height;
  m.camerasList_.resize(V); m.camerasPaths_.resize(V); m.pointsVisibleFromCamN_.resize(V);
  for (int v = 0; v < V; v++) {
    glm::mat4 cm(0.0f);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) cm[r][c] = d->cameras[(size_t)v * 12 + r * 4 + c];   /* read back as [row][col], SURVEY A.1 */
    cm[3][3] = 1.0f;
    m.camerasList_[v].cameraMatrix = cm; m.camerasList_[v].imageWidth = d->width; m.camerasList_[v].imageHeight = d->height;
  }
  m.numPoints_ = (int)d->n_tracks;
  for (int64_t t = 0; t < d->n_tracks; t++) {
    m.points_.push_back(glm::vec3(d->track_xyz[3 * t], d->track_xyz[3 * t + 1], d->track_xyz[3 * t + 2]));
    vector<int> vv; vector<glm::vec2> xy;
    for (int64_t k = d->track_off[t]; k < d->track_off[t + 1]; k++) {
      vv.push_back(d->track_view[k]); xy.push_back(glm::vec2(d->track_xy[2 * k], d->track_xy[2 * k + 1]));
      m.pointsVisibleFromCamN_[d->track_view[k]].push_back((int)t);
    }
    m.camViewingPointN_.push_back(vv); m.point2DoncamViewingPoint_.push_back(xy);
  }
  /* polyline graphs: ids, node ids and coordinates as given; node coordinates are those of the polylines' extremes, which is
   * what PolyLineGraph2D::is_valid_polyline (polyline_graph_2d.cpp:1141-1147) checks */
  s->plgs.resize(V);
  for (int v = 0; v < V; v++) {
    PolyLineGraph2DHMapImpl& g = s->plgs[v];
    const int64_t p0 = d->view_poly_off[v], p1 = d->view_poly_off[v + 1];
    ulong max_node = 0;
    for (int64_t p = p0; p < p1; p++) { max_node = std::max<ulong>(max_node, d->poly_start[p]); max_node = std::max<ulong>(max_node, d->poly_end[p]); }
    g.nodes_coords.assign(max_node + 1, glm::vec2(INVALID_POINT_COORDS, INVALID_POINT_COORDS));
    g.connections.assign(max_node + 1, vector<ulong>());
    g.visited_nodes.assign(max_node + 1, false);
    g.nodes_amount = g.real_nodes_amount = g.next_node_id = max_node + 1;
    for (int64_t p = p0; p < p1; p++) {
      vector<glm::vec2> pcs;
      for (int64_t k = d->poly_vert_off[p]; k < d->poly_vert_off[p + 1]; k++) pcs.push_back(glm::vec2(d->verts[2 * k], d->verts[2 * k + 1]));
      const ulong a = d->poly_start[p], b = d->poly_end[p];
      if (pcs.size() > 1) {
        const glm::vec2 inv(INVALID_POINT_COORDS, INVALID_POINT_COORDS);
        if ((g.nodes_coords[a] != inv && g.nodes_coords[a] != pcs[0]) || (a == b && pcs[0] != pcs[pcs.size() - 1]) ||
            (g.nodes_coords[b] != inv && a != b && g.nodes_coords[b] != pcs[pcs.size() - 1])) {
          s->error = "node coordinates are not consistent with the polylines' extremes: the reference's is_valid_polyline would reject them";
        }
        g.nodes_coords[a] = pcs[0]; g.nodes_coords[b] = pcs[pcs.size() - 1];
        g.connections[a].push_back((ulong)(p - p0)); if (b != a) g.connections[b].push_back((ulong)(p - p0));
        g.polylines.push_back(PolyLineGraph2D::polyline(a, b, pcs));
      } else {
        g.polylines.push_back(PolyLineGraph2D::polyline(a, b, vector<glm::vec2>(), -1.0f));   /* an invalidated polyline keeps its id and node ids, coordinates cleared (polyline_graph_2d.cpp:1035-1038, 1047-1058) */
      }
    }
  }
  s->F = new cv::Mat*[V];
  for (int a = 0; a < V; a++) {
    s->F[a] = new cv::Mat[V];
    for (int b = 0; b < V; b++) {
      if (a != b && d->fundamental_valid[(size_t)a * V + b]) {
        cv::Mat f(3, 3, CV_64F);
        for (int i = 0; i < 9; i++) f.at<double>(i / 3, i % 3) = d->fundamental[((size_t)a * V + b) * 9 + i];
        s->F[a][b] = f;
      } else if (a != b) s->F[a][b] = cv::Mat(1, 1, CV_32F);      /* geometric_utilities.cpp:780 */
    }
  }
  const cv::Size img_sz(d->width, d->height);
  for (int v = 0; v < V; v++) s->plmaps.push_back(PolyLine2DMapSearch(s->plgs[v], img_sz, 4.0f));   /* edge_matcher.cpp:101-103 */
  for (int v = 0; v < V; v++) s->imgs.push_back(cv::Mat::header_only(d->height, d->width, CV_8UC3));
  s->plgmm = new PLGMatchesManager(s->plgs, s->plg3d);
  if (d->n_tracks > 0) {
    s->em = new PLGEdgeManager(s->imgs, s->sfmd, (const cv::Mat**)s->F, s->plgs, DETECTION_STARTING_RADIUS, DETECTION_CORRESPONDENCES_MULTIPLICATION_FACTOR);   /* edge_matcher.cpp:107 */
    s->cm = new PLGPCM3ViewsPLGFollowing(s->imgs, s->sfmd, img_sz, (const cv::Mat**)s->F, s->plgs, s->plmaps);                                              /* edge_matcher.cpp:115 */
  }
  return s;
}
void eg3d_ref_scene_destroy(void* sc) { /* the reference leaks these objects as well; test processes are short-lived */ (void)sc; }

/* pipelines 1-2: find_new_3d_points_from_compatible_polylines_expandallviews_parallel once per candidate set, as
 * pipelines.cpp:92-100 / :140-146 do.  The reference returns one flat list: seed = set index, chain_pos = running index. */
void* eg3d_ref_match_polyline_sets(void* sc, const eg3d_candidates* c) {
  Quiet q;
  RefScene* s = (RefScene*)sc;
  RefPoints* out = new RefPoints();
  for (int k = 0; k < c->n_sets; k++) {
    const vector<set<ulong>> cs = cand_of(*s, c, k);
    vector<p3d_t> r = find_new_3d_points_from_compatible_polylines_expandallviews_parallel(s->sfmd, s->plgs, (const cv::Mat**)s->F, cs, *s->plgmm, s->plmaps);
    out->append(r, k);
  }
  return out;
}

/* one call of find_new_3d_points_from_compatible_polylines_starting_plgp_expandallviews (polyline_matching.cpp:134-144) per seed;
 * cands == NULL: every polyline of every view is a candidate (the all-segment sweep of BASELINE configs 2-4) */
void* eg3d_ref_match_seeds(void* sc, const eg3d_seeds* seeds, const eg3d_candidates* c) {
  Quiet q;
  RefScene* s = (RefScene*)sc;
  RefPoints* out = new RefPoints();
  vector<vector<set<ulong>>> csets;
  if (c) for (int k = 0; k < c->n_sets; k++) csets.push_back(cand_of(*s, c, k));
  vector<set<ulong>> all(s->V);
  if (!c) for (int v = 0; v < s->V; v++) for (ulong p = 0; p < s->plgs[v].polylines.size(); p++) if (s->plgs[v].is_valid_polyline(p)) all[v].insert(p);
  for (int64_t i = 0; i < seeds->n; i++) {
    const vector<set<ulong>>& cs = (c && seeds->cand_set && seeds->cand_set[i] >= 0) ? csets[seeds->cand_set[i]] : all;
    PolyLineGraph2D::plg_point p((ulong)seeds->polyline[i], (ulong)seeds->segment[i], glm::vec2(seeds->xy[2 * i], seeds->xy[2 * i + 1]));
    vector<p3d_t> r = find_new_3d_points_from_compatible_polylines_starting_plgp_expandallviews(s->sfmd, s->plgs, (const cv::Mat**)s->F, cs, seeds->view[i], p, *s->plgmm, s->plmaps);
    out->append(r, (int32_t)i);
  }
  return out;
}

/* The same per-seed calls issued from a real `omp parallel for` over the seeds (the reference's own `#pragma omp for` is orphaned
 * and runs on one thread, SURVEY finding 4; its per-seed function only reads shared state): bench.py's reference arm. */
void* eg3d_ref_match_seeds_mt(void* sc, const eg3d_seeds* seeds, const eg3d_candidates* c, int n_threads) {
  Quiet q;
  RefScene* s = (RefScene*)sc;
  vector<vector<set<ulong>>> csets;
  if (c) for (int k = 0; k < c->n_sets; k++) csets.push_back(cand_of(*s, c, k));
  vector<set<ulong>> all(s->V);
  if (!c) for (int v = 0; v < s->V; v++) for (ulong p = 0; p < s->plgs[v].polylines.size(); p++) if (s->plgs[v].is_valid_polyline(p)) all[v].insert(p);
  vector<vector<p3d_t>> res((size_t)seeds->n);
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
  for (int64_t i = 0; i < seeds->n; i++) {
    const vector<set<ulong>>& cs = (c && seeds->cand_set && seeds->cand_set[i] >= 0) ? csets[seeds->cand_set[i]] : all;
    PolyLineGraph2D::plg_point p((ulong)seeds->polyline[i], (ulong)seeds->segment[i], glm::vec2(seeds->xy[2 * i], seeds->xy[2 * i + 1]));
    res[i] = find_new_3d_points_from_compatible_polylines_starting_plgp_expandallviews(s->sfmd, s->plgs, (const cv::Mat**)s->F, cs, seeds->view[i], p, *s->plgmm, s->plmaps);
  }
  RefPoints* out = new RefPoints();
  for (int64_t i = 0; i < seeds->n; i++) out->append(res[i], (int32_t)i);
  return out;
}

/* K1 alone: find_epipolar_correspondences (polyline_matching.cpp:45-73): CSR over (seed, view) */
void* eg3d_ref_epipolar_intersect(void* sc, const eg3d_seeds* seeds, const eg3d_candidates* c, int64_t* n_hits) {
  Quiet q;
  RefScene* s = (RefScene*)sc;
  struct Hits { std::vector<int64_t> off{0}; std::vector<eg3d_hit> hits; };
  Hits* h = new Hits();
  vector<vector<set<ulong>>> csets;
  if (c) for (int k = 0; k < c->n_sets; k++) csets.push_back(cand_of(*s, c, k));
  vector<set<ulong>> all(s->V);
  if (!c) for (int v = 0; v < s->V; v++) for (ulong p = 0; p < s->plgs[v].polylines.size(); p++) if (s->plgs[v].is_valid_polyline(p)) all[v].insert(p);
  for (int64_t i = 0; i < seeds->n; i++) {
    const vector<set<ulong>>& cs = (c && seeds->cand_set && seeds->cand_set[i] >= 0) ? csets[seeds->cand_set[i]] : all;
    PolyLineGraph2D::plg_point p((ulong)seeds->polyline[i], (ulong)seeds->segment[i], glm::vec2(seeds->xy[2 * i], seeds->xy[2 * i + 1]));
    vector<vector<PolyLineGraph2D::plg_point>> e = find_epipolar_correspondences(s->sfmd, s->plgs, (const cv::Mat**)s->F, cs, seeds->view[i], p, *s->plgmm);
    for (int v = 0; v < s->V; v++) {
      for (const auto& g : e[v]) h->hits.push_back(eg3d_hit{(uint32_t)g.polyline_id, (uint32_t)g.plp.segment_index, g.plp.coords[0], g.plp.coords[1]});
      h->off.push_back((int64_t)h->hits.size());
    }
  }
  *n_hits = (int64_t)h->hits.size();
  return h;
}
void eg3d_ref_hits_get(void* hp, const int64_t** off, const eg3d_hit** hits) {
  struct Hits { std::vector<int64_t> off; std::vector<eg3d_hit> hits; };
  Hits* h = (Hits*)hp; *off = h->off.data(); *hits = h->hits.data();
}

/* pipeline 3: plg_matching_from_refpoint (plg_matching_from_refpoints.cpp:64-82, the body of the _parallel loop :89-96) for the
 * SfM points [tb, te); the whole range goes through plg_matching_from_refpoints_parallel itself */
void* eg3d_ref_match_refpoints(void* sc, int64_t tb, int64_t te) {
  Quiet q;
  RefScene* s = (RefScene*)sc;
  RefPoints* out = new RefPoints();
  if (tb == 0 && te == s->sfmd.numPoints_) {
    out->append(plg_matching_from_refpoints_parallel(s->sfmd, s->em, s->cm, *s->plgmm), 0);
    return out;
  }
  for (int64_t rp = tb; rp < te; rp++) out->append(plg_matching_from_refpoint(s->sfmd, s->em, s->cm, (ulong)rp, *s->plgmm), (int32_t)(rp - tb));
  return out;
}

void eg3d_ref_points_get(void* h, eg3d_points_view* v) {
  RefPoints* p = (RefPoints*)h;
  v->n_points = (int64_t)p->seed.size(); v->n_obs = (int64_t)p->obs_view.size();
  v->xyz = p->xyz.data(); v->seed = p->seed.data(); v->chain_pos = p->chain_pos.data(); v->obs_off = p->obs_off.data();
  v->obs_view = p->obs_view.data(); v->obs_poly = p->obs_poly.data(); v->obs_seg = p->obs_seg.data(); v->obs_xy = p->obs_xy.data();
}
void eg3d_ref_points_free(void* h) { delete (RefPoints*)h; }

/* a13: filter_3d_points_close_2d_array (filtering_close_plgps.cpp:97-124).  The function returns the kept records; keep[i] is
 * recovered by walking both lists in order. */
void eg3d_ref_dedup_close_points(void* sc, const eg3d_points_view* pts, uint8_t* keep) {
  RefScene* s = (RefScene*)sc;
  vector<p3d_t> in;
  for (int64_t i = 0; i < pts->n_points; i++) {
    vector<PolyLineGraph2D::plg_point> obs; vector<int> views;
    for (int64_t o = pts->obs_off[i]; o < pts->obs_off[i + 1]; o++) {
      obs.push_back(PolyLineGraph2D::plg_point((ulong)pts->obs_poly[o], (ulong)pts->obs_seg[o], glm::vec2(pts->obs_xy[2 * o], pts->obs_xy[2 * o + 1])));
      views.push_back(pts->obs_view[o]);
    }
    in.push_back(p3d_t(glm::vec3((float)i, 0.f, 0.f), obs, views));   /* X carries the input index */
  }
  vector<p3d_t> kept = filter_3d_points_close_2d_array(s->plgs, cv::Size(s->sfmd.imageWidth_, s->sfmd.imageHeight_), in);
  for (int64_t i = 0; i < pts->n_points; i++) keep[i] = 0;
  for (const auto& k : kept) keep[(int64_t)std::get<0>(k)[0]] = 1;
}

/* a14: compute_inliers (outliers_filtering.cpp:37-64) = gaussNewtonFiltering + the view-count rule; xyz is overwritten for
 * Gauss-Newton inliers as gauss_newton.cpp:168-173 does.  forced_min_filter < 0 = INVALID_FORCED_MIN_FILTER. */
void eg3d_ref_filter(void* sc, int64_t n, float* xyz, const int64_t* obs_off, const int32_t* obs_view, const float* obs_xy,
                     int64_t first_edgepoint, float gn_max_mse, int32_t forced_min_filter, uint8_t* inliers) {
  Quiet q;
  RefScene* s = (RefScene*)sc;
  SfMData m;
  m.numCameras_ = s->sfmd.numCameras_; m.camerasList_ = s->sfmd.camerasList_; m.camerasPaths_ = s->sfmd.camerasPaths_;
  m.imageWidth_ = s->sfmd.imageWidth_; m.imageHeight_ = s->sfmd.imageHeight_;
  m.pointsVisibleFromCamN_.resize(m.numCameras_);
  m.numPoints_ = (int)n;
  for (int64_t i = 0; i < n; i++) {
    m.points_.push_back(glm::vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    vector<int> vv; vector<glm::vec2> xy;
    for (int64_t o = obs_off[i]; o < obs_off[i + 1]; o++) { vv.push_back(obs_view[o]); xy.push_back(glm::vec2(obs_xy[2 * o], obs_xy[2 * o + 1])); m.pointsVisibleFromCamN_[obs_view[o]].push_back((int)i); }
    m.camViewingPointN_.push_back(vv); m.point2DoncamViewingPoint_.push_back(xy);
  }
  vector<bool> inl = compute_inliers(m, (int)first_edgepoint, gn_max_mse, forced_min_filter < 0 ? EG3D_INVALID_FORCED_MIN_FILTER : forced_min_filter);
  for (int64_t i = 0; i < n; i++) { inliers[i] = inl[i] ? 1 : 0; xyz[3 * i] = m.points_[i][0]; xyz[3 * i + 1] = m.points_[i][1]; xyz[3 * i + 2] = m.points_[i][2]; }
}

/* f3 (reference side): OpenMvgParser (external/manifoldReconstructor/src/OpenMvgParser.cpp) on a JSON file; cameras [V][12]
 * as the path reads them (rows 0..2 of cameraMatrix read as [row][col]).  Returns the number of views, or -1. */
int eg3d_ref_parse_openmvg(const char* path, int max_views, float* cameras12, int* width, int* height, int64_t* n_points) {
  Quiet q;
  OpenMvgParser op(path);
  op.parse();
  SfMData m = op.getSfmData();
  if (m.numCameras_ > max_views) return -1;
  for (int v = 0; v < m.numCameras_; v++)
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) cameras12[(size_t)v * 12 + r * 4 + c] = m.camerasList_[v].cameraMatrix[r][c];
  *width = m.numCameras_ > 0 ? m.camerasList_[0].imageWidth : 0; *height = m.numCameras_ > 0 ? m.camerasList_[0].imageHeight : 0;   /* the parser fills the per-camera fields */
  *n_points = (int64_t)m.points_.size();
  return m.numCameras_;
}

}  // extern "C"
