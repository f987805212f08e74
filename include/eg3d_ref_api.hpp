// eg3d_ref_api.hpp — header-only C++ shim that keeps EdgeGraph3D's own names and signatures for the hot path on top of
// the C-ABI of libeg3d.so (include/eg3d.h).  A maintainer of the reference includes this file instead of
// polyline_matching.hpp / outliers_filtering.hpp / filtering_close_plgps.hpp in pipelines.cpp and edge_matcher.cpp
// (INTEGRATION.md shows the patch).  Inside the reference tree define EG3D_REF_USE_REFERENCE_TYPES before including it
// and the shim binds to the real SfMData / PolyLineGraph2DHMapImpl / cv::Mat types; stand-alone (this repository's
// tests) it declares layout-compatible minimal versions of those types.
#pragma once
#include <cstdint>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <cstring>
#include <vector>
#include "eg3d.h"

#ifndef EG3D_REF_USE_REFERENCE_TYPES
namespace eg3d_ref {
typedef unsigned long ulong;
struct vec2 { float x, y; vec2() : x(0), y(0) {} vec2(float a, float b) : x(a), y(b) {} float& operator[](int i) { return i ? y : x; } float operator[](int i) const { return i ? y : x; } };
struct vec3 { float x, y, z; vec3() : x(0), y(0), z(0) {} vec3(float a, float b, float c) : x(a), y(b), z(c) {} float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); } float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); } };
struct mat4 { float m[4][4]; const float* operator[](int i) const { return m[i]; } float* operator[](int i) { return m[i]; } };
// external/manifoldReconstructor/include/types_reconstructor.hpp:68-82 (fields the path reads)
struct CameraType { mat4 cameraMatrix; int imageWidth = 0, imageHeight = 0; };
// external/manifoldReconstructor/include/SfMData.h:16-30
struct SfMData {
  int numPoints_ = 0, numCameras_ = 0;
  std::vector<vec3> points_;
  std::vector<CameraType> camerasList_;
  std::vector<std::vector<int>> camViewingPointN_;
  std::vector<std::vector<int>> pointsVisibleFromCamN_;
  std::vector<std::vector<vec2>> point2DoncamViewingPoint_;
  int imageWidth_ = 0, imageHeight_ = 0;
};
// include/edgegraph3d/plgs/polyline_graph_2d.hpp:85-119, 278-294
struct PolyLineGraph2D {
  struct polyline {
    struct pl_point { ulong segment_index; vec2 coords; pl_point() : segment_index(0) {} pl_point(ulong s, const vec2& c) : segment_index(s), coords(c) {} };
    ulong start = 0, end = 0;
    std::vector<vec2> polyline_coords;
  };
  struct plg_point {
    ulong polyline_id; polyline::pl_point plp;
    plg_point() : polyline_id(0) {}
    plg_point(ulong id, ulong seg, const vec2& c) : polyline_id(id), plp(seg, c) {}
  };
  std::vector<polyline> polylines;
};
typedef PolyLineGraph2D PolyLineGraph2DHMapImpl;
}  // namespace eg3d_ref
#define EG3D_REF_NS eg3d_ref
#else
#define EG3D_REF_NS
#endif

namespace eg3d_shim {
using namespace EG3D_REF_NS;
typedef unsigned long ulong_t;

// polyline_graph_2d.hpp:451
typedef std::tuple<vec3, std::vector<PolyLineGraph2D::plg_point>, std::vector<int>> new_3dpoint_plgp_matches;

inline void check(eg3d_status st) {
  if (st != EG3D_OK) throw std::runtime_error(std::string("libeg3d: ") + eg3d_last_error());
}

// `const Mat** all_fundamental_matrices` (edge_graph_3d_utilities.cpp:581-589) flattened: F[a][b] row-major 3x3 doubles,
// valid[a][b] = 0 for the 1x1 dummy Mat of pairs with < 10 common tracks (geometric_utilities.cpp:780).
struct FundamentalSet {
  std::vector<double> F; std::vector<uint8_t> valid; int V = 0;
  explicit FundamentalSet(int n = 0) : F((size_t)n * n * 9, 0.0), valid((size_t)n * n, 0), V(n) {}
  void set(int a, int b, const double f9[9]) { for (int i = 0; i < 9; i++) F[((size_t)a * V + b) * 9 + i] = f9[i]; valid[(size_t)a * V + b] = 1; }
#ifdef EG3D_REF_USE_REFERENCE_TYPES
  FundamentalSet(const cv::Mat** all_fundamental_matrices, int n) : FundamentalSet(n) {
    for (int a = 0; a < n; a++) for (int b = 0; b < n; b++) {
      const cv::Mat& m = all_fundamental_matrices[a][b];
      if (a != b && m.rows == 3 && m.cols == 3) { double f[9]; for (int i = 0; i < 9; i++) f[i] = m.at<double>(i / 3, i % 3); set(a, b, f); }
    }
  }
#endif
};

// OpenMvgParser::parse (OpenMvgParser.cpp:75-153, 241-301) without rapidjson: the C++ reader of libeg3d.so fills the SfMData fields
// the path reads (cameras bit-identical to the reference's glm arithmetic; tracks in file order).
inline SfMData load_sfm_data(const std::string& path) {
  eg3d_sfm* h = nullptr;
  check(eg3d_sfm_load(path.c_str(), &h));
  eg3d_sfm_view v;
  check(eg3d_sfm_get(h, &v));
  SfMData s;
  s.numCameras_ = v.n_views; s.imageWidth_ = v.width; s.imageHeight_ = v.height; s.numPoints_ = (int)v.n_tracks;
  s.camerasList_.resize((size_t)v.n_views);
  for (int i = 0; i < v.n_views; i++) {
    CameraType& c = s.camerasList_[(size_t)i];
    for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) c.cameraMatrix[r][k] = r < 3 ? v.cameras[(size_t)i * 12 + r * 4 + k] : (k == 3 ? 1.f : 0.f);
    c.imageWidth = v.width; c.imageHeight = v.height;
  }
  for (int64_t p = 0; p < v.n_tracks; p++) {
    s.points_.push_back(vec3(v.track_xyz[3 * p], v.track_xyz[3 * p + 1], v.track_xyz[3 * p + 2]));
    s.camViewingPointN_.push_back(std::vector<int>()); s.point2DoncamViewingPoint_.push_back(std::vector<vec2>());
    for (int64_t o = v.track_off[p]; o < v.track_off[p + 1]; o++) {
      s.camViewingPointN_.back().push_back(v.track_view[o]);
      s.point2DoncamViewingPoint_.back().push_back(vec2(v.track_xy[2 * o], v.track_xy[2 * o + 1]));
    }
  }
  eg3d_sfm_free(h);
  return s;
}
// output_sfm_data (output_sfm_data.cpp:186-229): the original file's views / intrinsics / extrinsics + the points of `sfmd` as `structure`
// (optionally only those flagged in `inliers`); returns the number of points written.
inline int64_t output_sfm_data(const std::string& original_json, const SfMData& sfmd, const std::string& out_path, const std::vector<bool>* inliers = nullptr) {
  std::vector<float> xyz, xy; std::vector<int64_t> off(1, 0); std::vector<int32_t> view; std::vector<uint8_t> keep;
  for (size_t p = 0; p < sfmd.points_.size(); p++) {
    xyz.push_back(sfmd.points_[p][0]); xyz.push_back(sfmd.points_[p][1]); xyz.push_back(sfmd.points_[p][2]);
    for (size_t k = 0; k < sfmd.camViewingPointN_[p].size(); k++) {
      view.push_back((int32_t)sfmd.camViewingPointN_[p][k]);
      xy.push_back(sfmd.point2DoncamViewingPoint_[p][k][0]); xy.push_back(sfmd.point2DoncamViewingPoint_[p][k][1]);
    }
    off.push_back((int64_t)view.size());
    if (inliers) keep.push_back((*inliers)[p] ? 1 : 0);
  }
  int64_t n = 0;
  check(eg3d_sfm_save(out_path.c_str(), original_json.c_str(), (int64_t)sfmd.points_.size(), xyz.data(), off.data(), view.data(), xy.data(), inliers ? keep.data() : nullptr, &n));
  return n;
}

// generate_all_fundamental_matrices (geometric_utilities.cpp:816-820 -> _from_Points :800-812): LMedS matrices from the SfM tracks,
// host code of libeg3d.so (eg3d_fundamental_from_tracks: same estimator family as cv::findFundamentalMat(FM_LMEDS), not bit-identical;
// a maintainer who keeps OpenCV's matrices uses the `cv::Mat**` constructor of FundamentalSet above instead).
inline FundamentalSet generate_all_fundamental_matrices(const SfMData& sfmd) {
  const int V = (int)sfmd.camerasList_.size();
  std::vector<int64_t> off(1, 0); std::vector<int32_t> view; std::vector<float> xy;
  for (size_t p = 0; p < sfmd.points_.size(); p++) {
    for (size_t k = 0; k < sfmd.camViewingPointN_[p].size(); k++) {
      view.push_back((int32_t)sfmd.camViewingPointN_[p][k]);
      xy.push_back(sfmd.point2DoncamViewingPoint_[p][k][0]); xy.push_back(sfmd.point2DoncamViewingPoint_[p][k][1]);
    }
    off.push_back((int64_t)view.size());
  }
  eg3d_scene_desc d; std::memset(&d, 0, sizeof d);
  d.n_views = V; d.n_tracks = (int64_t)sfmd.points_.size(); d.track_off = off.data(); d.track_view = view.data(); d.track_xy = xy.data();
  FundamentalSet F(V);
  check(eg3d_fundamental_from_tracks(&d, 10 /* MIN_CORRESPONDENCES_AMOUNT */, F.F.data(), F.valid.data()));
  return F;
}

// Owns the device-resident scene: the role `plgs`, `plmaps`, `all_fundamental_matrices`, `em` play as long-lived
// arguments of the reference's entry points (edge_matcher.cpp:83-115 builds them once).
class Eg3dScene {
 public:
  Eg3dScene(const SfMData& sfmd, const std::vector<PolyLineGraph2DHMapImpl>& plgs, const FundamentalSet& F, const eg3d_params* params = nullptr) {
    const int V = (int)plgs.size();
    cams_.resize((size_t)V * 12);
    for (int v = 0; v < V; v++)   // convert_glm_mat4_to_cv_Mat34: glm_mat[row][col] (edge_graph_3d_utilities.cpp:190-206)
      for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) cams_[(size_t)v * 12 + r * 4 + c] = sfmd.camerasList_[v].cameraMatrix[r][c];
    view_poly_off_.assign(1, 0); poly_vert_off_.assign(1, 0);
    for (int v = 0; v < V; v++) {
      for (const auto& pl : plgs[v].polylines) {
        for (const auto& c : pl.polyline_coords) { verts_.push_back(c[0]); verts_.push_back(c[1]); }
        poly_vert_off_.push_back((int64_t)verts_.size() / 2);
        start_.push_back((uint32_t)pl.start); end_.push_back((uint32_t)pl.end);
      }
      view_poly_off_.push_back((int64_t)start_.size());
    }
    track_off_.assign(1, 0);
    for (size_t t = 0; t < sfmd.points_.size(); t++) {
      track_xyz_.push_back(sfmd.points_[t][0]); track_xyz_.push_back(sfmd.points_[t][1]); track_xyz_.push_back(sfmd.points_[t][2]);
      for (size_t k = 0; k < sfmd.camViewingPointN_[t].size(); k++) {
        track_view_.push_back(sfmd.camViewingPointN_[t][k]);
        track_xy_.push_back(sfmd.point2DoncamViewingPoint_[t][k][0]); track_xy_.push_back(sfmd.point2DoncamViewingPoint_[t][k][1]);
      }
      track_off_.push_back((int64_t)track_view_.size());
    }
    desc_ = eg3d_scene_desc();
    desc_.n_views = V; desc_.width = sfmd.imageWidth_; desc_.height = sfmd.imageHeight_;
    desc_.cameras = cams_.data(); desc_.fundamental = F.F.data(); desc_.fundamental_valid = F.valid.data();
    desc_.view_poly_off = view_poly_off_.data(); desc_.poly_vert_off = poly_vert_off_.data(); desc_.verts = verts_.data();
    desc_.poly_start = start_.data(); desc_.poly_end = end_.data();
    desc_.n_tracks = (int64_t)sfmd.points_.size(); desc_.track_xyz = track_xyz_.data(); desc_.track_off = track_off_.data();
    desc_.track_view = track_view_.data(); desc_.track_xy = track_xy_.data();
    check(eg3d_scene_create(&desc_, params, &scene_));
  }
  ~Eg3dScene() { if (scene_) eg3d_scene_destroy(scene_); }
  Eg3dScene(const Eg3dScene&) = delete;
  Eg3dScene& operator=(const Eg3dScene&) = delete;
  eg3d_scene* handle() const { return scene_; }
  const eg3d_scene_desc& desc() const { return desc_; }
  int n_views() const { return desc_.n_views; }

 private:
  std::vector<float> cams_, verts_, track_xyz_, track_xy_;
  std::vector<int64_t> view_poly_off_, poly_vert_off_, track_off_;
  std::vector<uint32_t> start_, end_;
  std::vector<int32_t> track_view_;
  eg3d_scene_desc desc_;
  eg3d_scene* scene_ = nullptr;
};

inline std::vector<new_3dpoint_plgp_matches> points_to_reference(eg3d_points* pts) {
  eg3d_points_view v; check(eg3d_points_get(pts, &v));
  std::vector<new_3dpoint_plgp_matches> res((size_t)v.n_points);
  for (int64_t i = 0; i < v.n_points; i++) {
    std::get<0>(res[i]) = vec3(v.xyz[3 * i], v.xyz[3 * i + 1], v.xyz[3 * i + 2]);
    for (int64_t o = v.obs_off[i]; o < v.obs_off[i + 1]; o++) {
      std::get<1>(res[i]).push_back(PolyLineGraph2D::plg_point(v.obs_poly[o], v.obs_seg[o], vec2(v.obs_xy[2 * o], v.obs_xy[2 * o + 1])));
      std::get<2>(res[i]).push_back(v.obs_view[o]);
    }
  }
  eg3d_points_free(pts);
  return res;
}

// B1 (polyline_matching.hpp:55-56), batched over polyline matches exactly as the loop at pipelines.cpp:92-100 /
// :138-147 calls it: one `vector<set<ulong>>` per match.  The PLGMatchesManager argument of the reference is not needed:
// with SWITCH_RUNPARALLEL it is never written on this path (SURVEY finding 5).
inline std::vector<new_3dpoint_plgp_matches> find_new_3d_points_from_compatible_polylines_expandallviews_parallel(
    Eg3dScene& scene, const std::vector<std::vector<std::set<ulong_t>>>& potentially_compatible_polylines_per_match,
    int starting_view_begin = 0, int starting_view_end = -1) {
  const int V = scene.n_views();
  std::vector<int64_t> off(1, 0); std::vector<uint32_t> ids;
  for (const auto& match : potentially_compatible_polylines_per_match)
    for (int v = 0; v < V; v++) { for (ulong_t id : match[v]) ids.push_back((uint32_t)id); off.push_back((int64_t)ids.size()); }
  eg3d_candidates c; c.n_sets = (int32_t)potentially_compatible_polylines_per_match.size(); c.off = off.data(); c.polyline = ids.data();
  eg3d_points* pts = nullptr;
  check(eg3d_match_polyline_sets(scene.handle(), &c, starting_view_begin, starting_view_end < 0 ? V : starting_view_end, &pts, nullptr));
  return points_to_reference(pts);
}

// B2 (plg_matching_from_refpoints.hpp:53-55)
inline std::vector<new_3dpoint_plgp_matches> plg_matching_from_refpoints_parallel(Eg3dScene& scene, int64_t refpoint_begin = 0, int64_t refpoint_end = -1) {
  eg3d_points* pts = nullptr;
  check(eg3d_match_refpoints(scene.handle(), refpoint_begin, refpoint_end < 0 ? scene.desc().n_tracks : refpoint_end, &pts, nullptr));
  return points_to_reference(pts);
}

// ---------------------------------------------------------------------------------------------------------------------
// B4 + B3: the triangulation entry and the two abstract classes the reference carries in its data_bundle
// (utils/data_bundle.hpp:45-54).  A maintainer may keep an EdgeManager of their own and swap only the consensus step, or the
// other way round: both sides of the seam are here with the reference's signatures.
// ---------------------------------------------------------------------------------------------------------------------
typedef std::pair<std::vector<PolyLineGraph2D::plg_point>, std::vector<std::vector<std::vector<PolyLineGraph2D::plg_point>>>> intersections_and_correspondences_t;

// B4, batched: compute_3D_point_multiple_views_plg_following_expandallviews_vector (triangulation.hpp:98) for several seeds at
// once; epipolar_correspondences[i] is the V-vector of hit lists of seed i.  One result vector per seed.
inline std::vector<std::vector<new_3dpoint_plgp_matches>> compute_3D_point_multiple_views_plg_following_expandallviews_vector_batch(
    Eg3dScene& scene, const std::vector<int>& starting_plg_ids, const std::vector<std::vector<std::vector<PolyLineGraph2D::plg_point>>>& epipolar_correspondences) {
  const int V = scene.n_views();
  const int64_t n = (int64_t)starting_plg_ids.size();
  std::vector<int32_t> sv(starting_plg_ids.begin(), starting_plg_ids.end());
  std::vector<int64_t> off(1, 0); std::vector<eg3d_hit> hits;
  for (int64_t i = 0; i < n; i++)
    for (int v = 0; v < V; v++) {
      if (v < (int)epipolar_correspondences[(size_t)i].size())
        for (const auto& q : epipolar_correspondences[(size_t)i][(size_t)v]) { eg3d_hit h; h.polyline = (uint32_t)q.polyline_id; h.segment = (uint32_t)q.plp.segment_index; h.x = q.plp.coords[0]; h.y = q.plp.coords[1]; hits.push_back(h); }
      off.push_back((int64_t)hits.size());
    }
  eg3d_points* pts = nullptr;
  check(eg3d_match_correspondences(scene.handle(), n, sv.data(), off.data(), hits.data(), &pts, nullptr));
  eg3d_points_view v; check(eg3d_points_get(pts, &v));
  std::vector<std::vector<new_3dpoint_plgp_matches>> res((size_t)n);
  for (int64_t i = 0; i < v.n_points; i++) {
    new_3dpoint_plgp_matches p;
    std::get<0>(p) = vec3(v.xyz[3 * i], v.xyz[3 * i + 1], v.xyz[3 * i + 2]);
    for (int64_t o = v.obs_off[i]; o < v.obs_off[i + 1]; o++) {
      std::get<1>(p).push_back(PolyLineGraph2D::plg_point(v.obs_poly[o], v.obs_seg[o], vec2(v.obs_xy[2 * o], v.obs_xy[2 * o + 1])));
      std::get<2>(p).push_back(v.obs_view[o]);
    }
    res[(size_t)v.seed[i]].push_back(p);
  }
  eg3d_points_free(pts);
  return res;
}
// B4 with the reference's own shape: one seed (the sfmd / plgs / F / plmaps arguments live in the scene handle)
inline std::vector<new_3dpoint_plgp_matches> compute_3D_point_multiple_views_plg_following_expandallviews_vector(
    Eg3dScene& scene, const int starting_plg_id, const std::vector<std::vector<PolyLineGraph2D::plg_point>>& epipolar_correspondences) {
  return compute_3D_point_multiple_views_plg_following_expandallviews_vector_batch(scene, std::vector<int>(1, starting_plg_id),
                                                                                  std::vector<std::vector<std::vector<PolyLineGraph2D::plg_point>>>(1, epipolar_correspondences))[0];
}

// plgp_consensus_manager.hpp:56-72
class PLGPConsensusManager {
 public:
  virtual std::vector<new_3dpoint_plgp_matches> consensus_strategy_single_point(const int starting_img_id, const int refpoint, const intersections_and_correspondences_t& intersections_and_correspondences_pair) = 0;
  virtual std::vector<std::vector<new_3dpoint_plgp_matches>> consensus_strategy_single_point_vector(const int starting_img_id, const int refpoint, const intersections_and_correspondences_t& intersections_and_correspondences_pair) = 0;
  virtual ~PLGPConsensusManager() {}
 protected:
  const SfMData& sfmd;
  Eg3dScene& scene;          // stands for the reference's (imgs, img_size, all_fundamental_matrices, plgs) members
  PLGPConsensusManager(const SfMData& input_sfmd, Eg3dScene& input_scene) : sfmd(input_sfmd), scene(input_scene) {}
};

// plgpcm_3views_plg_following.hpp:50-58, plgpcm_3views_plg_following.cpp:40-69: scatter the per-observing-view hit lists of
// every intersection into a V-vector and run the triangulation entry — all intersections of the pair in ONE device call.
class PLGPCM3ViewsPLGFollowing : public PLGPConsensusManager {
 public:
  PLGPCM3ViewsPLGFollowing(const SfMData& input_sfmd, Eg3dScene& input_scene) : PLGPConsensusManager(input_sfmd, input_scene) {}
  std::vector<std::vector<new_3dpoint_plgp_matches>> consensus_strategy_single_point_vector(const int starting_img_id, const int refpoint, const intersections_and_correspondences_t& iac) override {
    const size_t n = iac.first.size();
    std::vector<std::vector<std::vector<PolyLineGraph2D::plg_point>>> all(n, std::vector<std::vector<PolyLineGraph2D::plg_point>>((size_t)scene.n_views()));
    for (size_t k = 0; k < n; k++)      // consensus_strategy_single_point_single_intersection, :40-44
      for (size_t i = 0; i < sfmd.camViewingPointN_[(size_t)refpoint].size(); i++) all[k][(size_t)sfmd.camViewingPointN_[(size_t)refpoint][i]] = iac.second[k][i];
    return compute_3D_point_multiple_views_plg_following_expandallviews_vector_batch(scene, std::vector<int>(n, starting_img_id), all);
  }
  std::vector<new_3dpoint_plgp_matches> consensus_strategy_single_point(const int starting_img_id, const int refpoint, const intersections_and_correspondences_t& iac) override {
    std::vector<new_3dpoint_plgp_matches> res;
    for (auto& cur : consensus_strategy_single_point_vector(starting_img_id, refpoint, iac)) res.insert(res.end(), cur.begin(), cur.end());
    return res;
  }
};

// edge_manager.hpp:54-73 narrowed to what pipeline 3 calls through its `(PLGEdgeManager*) em` downcast
// (plg_matching_from_refpoints.cpp:67): one (intersections, correspondences) pair per observing view of the SfM point.
class EdgeManager {
 public:
  virtual std::vector<intersections_and_correspondences_t> detect_nearby_intersections_and_correspondences_plgp(const int starting_point_id) = 0;
  virtual ~EdgeManager() {}
};

// plg_edge_manager.hpp:58-105 (seeding part, plg_edge_manager.cpp:191-300) on the device: 10 px seeds, 30 px candidate
// neighbourhoods, epipolar intersections within 3 x |observation - seed|.
class PLGEdgeManager : public EdgeManager {
 public:
  PLGEdgeManager(const SfMData& input_sfmd, Eg3dScene& input_scene) : sfmd(input_sfmd), scene(input_scene) {}
  std::vector<intersections_and_correspondences_t> detect_nearby_intersections_and_correspondences_plgp(const int starting_point_id) override {
    eg3d_corr* c = nullptr;
    check(eg3d_refpoint_correspondences(scene.handle(), starting_point_id, starting_point_id + 1, &c));
    int64_t n = 0; int32_t V = 0; const int32_t* view; const uint32_t *pl, *seg; const float* xy; const int64_t* off; const eg3d_hit* hits;
    check(eg3d_corr_get(c, &n, &V, &view, &pl, &seg, &xy, nullptr, &off, &hits));
    const std::vector<int>& cams = sfmd.camViewingPointN_[(size_t)starting_point_id];
    std::vector<intersections_and_correspondences_t> res(cams.size());
    int64_t s = 0;
    for (size_t i = 0; i < cams.size(); i++)            // seeds come in the reference's order: observing view by observing view
      for (; s < n && view[s] == cams[i]; s++) {
        res[i].first.push_back(PolyLineGraph2D::plg_point(pl[s], seg[s], vec2(xy[2 * s], xy[2 * s + 1])));
        std::vector<std::vector<PolyLineGraph2D::plg_point>> per_cam(cams.size());
        for (size_t j = 0; j < cams.size(); j++)
          for (int64_t k = off[s * V + cams[j]]; k < off[s * V + cams[j] + 1]; k++)
            per_cam[j].push_back(PolyLineGraph2D::plg_point(hits[k].polyline, hits[k].segment, vec2(hits[k].x, hits[k].y)));
        res[i].second.push_back(per_cam);
      }
    eg3d_corr_free(c);
    return res;
  }
 private:
  const SfMData& sfmd;
  Eg3dScene& scene;
};

// B2 with the reference's own shape (plg_matching_from_refpoints.hpp:53-55; plg_matching_from_refpoints.cpp:64-104): ANY EdgeManager
// with ANY PLGPConsensusManager — the fused eg3d_match_refpoints (plg_matching_from_refpoints_parallel above) is the fast path
// when both are the library's own.
inline std::vector<new_3dpoint_plgp_matches> plg_matching_from_refpoints(const SfMData& sfm_data, EdgeManager* em, PLGPConsensusManager* cm) {
  std::vector<new_3dpoint_plgp_matches> res;
  for (size_t refpoint_id = 0; refpoint_id < sfm_data.points_.size(); refpoint_id++) {
    const std::vector<intersections_and_correspondences_t> all = em->detect_nearby_intersections_and_correspondences_plgp((int)refpoint_id);
    for (size_t i = 0; i < sfm_data.camViewingPointN_[refpoint_id].size(); i++)
      for (const auto& cur : cm->consensus_strategy_single_point_vector(sfm_data.camViewingPointN_[refpoint_id][i], (int)refpoint_id, all[i])) res.insert(res.end(), cur.begin(), cur.end());
  }
  return res;
}

// a13 (filtering_close_plgps.hpp): first-come-first-kept density limiter over the gathered points
inline std::vector<new_3dpoint_plgp_matches> filter_3d_points_close_2d_array(Eg3dScene& scene, const std::vector<new_3dpoint_plgp_matches>& p3ds) {
  std::vector<float> xyz, xy; std::vector<int32_t> seed(p3ds.size(), 0), pos(p3ds.size(), 0), view; std::vector<uint32_t> pl, seg; std::vector<int64_t> off(1, 0);
  for (const auto& p : p3ds) {
    xyz.push_back(std::get<0>(p)[0]); xyz.push_back(std::get<0>(p)[1]); xyz.push_back(std::get<0>(p)[2]);
    for (size_t k = 0; k < std::get<1>(p).size(); k++) {
      const auto& q = std::get<1>(p)[k];
      view.push_back(std::get<2>(p)[k]); pl.push_back((uint32_t)q.polyline_id); seg.push_back((uint32_t)q.plp.segment_index);
      xy.push_back(q.plp.coords[0]); xy.push_back(q.plp.coords[1]);
    }
    off.push_back((int64_t)view.size());
  }
  eg3d_points_view v; v.n_points = (int64_t)p3ds.size(); v.n_obs = (int64_t)view.size(); v.xyz = xyz.data(); v.seed = seed.data(); v.chain_pos = pos.data();
  v.obs_off = off.data(); v.obs_view = view.data(); v.obs_poly = pl.data(); v.obs_seg = seg.data(); v.obs_xy = xy.data();
  std::vector<uint8_t> keep(p3ds.size());
  check(eg3d_dedup_close_points(scene.handle(), &v, keep.data()));
  std::vector<new_3dpoint_plgp_matches> res;
  for (size_t i = 0; i < p3ds.size(); i++) if (keep[i]) res.push_back(p3ds[i]);
  return res;
}

// B6 (outliers_filtering.hpp:18-21): GN refinement + inlier test + view-count rule + removeOutliers compaction
inline void filter(Eg3dScene& scene, SfMData& sfmd, int first_edgepoint, float gn_max_mse = 2.25f, int forced_min_filter = -1) {
  const int64_t n = (int64_t)sfmd.points_.size();
  std::vector<float> xyz, xy; std::vector<int32_t> view; std::vector<int64_t> off(1, 0);
  for (int64_t i = 0; i < n; i++) {
    xyz.push_back(sfmd.points_[i][0]); xyz.push_back(sfmd.points_[i][1]); xyz.push_back(sfmd.points_[i][2]);
    for (size_t k = 0; k < sfmd.camViewingPointN_[i].size(); k++) {
      view.push_back(sfmd.camViewingPointN_[i][k]);
      xy.push_back(sfmd.point2DoncamViewingPoint_[i][k][0]); xy.push_back(sfmd.point2DoncamViewingPoint_[i][k][1]);
    }
    off.push_back((int64_t)view.size());
  }
  std::vector<uint8_t> inl((size_t)n);
  check(eg3d_filter(scene.handle(), n, xyz.data(), off.data(), view.data(), xy.data(), first_edgepoint, gn_max_mse, forced_min_filter, inl.data(), nullptr));
  SfMData res;   // removeOutliers, outliers_filtering.cpp:66-92
  res.camerasList_ = sfmd.camerasList_; res.numCameras_ = sfmd.numCameras_; res.imageWidth_ = sfmd.imageWidth_; res.imageHeight_ = sfmd.imageHeight_;
  res.pointsVisibleFromCamN_.resize(res.numCameras_);
  int cur = 0;
  for (int64_t i = 0; i < n; i++) if (inl[i]) {
    res.points_.push_back(vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));   // GN inliers carry the refined position (gauss_newton.cpp:168-173)
    res.camViewingPointN_.push_back(sfmd.camViewingPointN_[i]); res.point2DoncamViewingPoint_.push_back(sfmd.point2DoncamViewingPoint_[i]);
    for (int cam : sfmd.camViewingPointN_[i]) res.pointsVisibleFromCamN_[cam].push_back(cur);
    cur++;
  }
  res.numPoints_ = (int)res.points_.size();
  sfmd = res;
}

// f1: convertEdgeImagePolyLineGraph_optimized / convert_edge_images_to_optimized_polyline_graphs
// (convert_edge_images_pixel_to_segment.hpp:98-101).  img = the pixels of a continuous CV_8UC3 Mat (img.data, img.rows,
// img.cols); with EG3D_REF_USE_REFERENCE_TYPES the cv::Mat overloads below forward to it.  Only the fields the matching
// path reads are filled (polyline start / end / coords); removed polylines keep their id with empty coords, as in the
// reference (polyline_graph_2d.cpp:1035-1038).
inline PolyLineGraph2DHMapImpl convertEdgeImagePolyLineGraph_optimized(const uint8_t* bgr, int rows, int cols, const uint8_t edge_color[3]) {
  eg3d_plg* h = nullptr;
  check(eg3d_plg_from_edge_image(bgr, rows, cols, 3, edge_color, EG3D_PLG_STAGE_FULL, &h));
  eg3d_plg_view v; check(eg3d_plg_get(h, &v));
  PolyLineGraph2DHMapImpl plg;
  plg.polylines.resize((size_t)v.n_polylines);
  for (int64_t i = 0; i < v.n_polylines; i++) {
    auto& pl = plg.polylines[(size_t)i];
    pl.start = v.poly_start[i]; pl.end = v.poly_end[i];
    for (int64_t k = v.poly_vert_off[i]; k < v.poly_vert_off[i + 1]; k++) pl.polyline_coords.push_back(vec2(v.verts[2 * k], v.verts[2 * k + 1]));
  }
  eg3d_plg_free(h);
  return plg;
}
#ifdef EG3D_REF_USE_REFERENCE_TYPES
inline PolyLineGraph2DHMapImpl convertEdgeImagePolyLineGraph_optimized(const cv::Mat& img, const cv::Vec3b& edge_color) {
  const cv::Mat c = img.isContinuous() ? img : img.clone();
  const uint8_t col[3] = {edge_color[0], edge_color[1], edge_color[2]};
  return convertEdgeImagePolyLineGraph_optimized(c.data, c.rows, c.cols, col);
}
inline std::vector<PolyLineGraph2DHMapImpl> convert_edge_images_to_optimized_polyline_graphs(const std::vector<cv::Mat>& imgs, const cv::Vec3b& edge_color) {
  std::vector<PolyLineGraph2DHMapImpl> res;
  for (const auto& img : imgs) res.push_back(convertEdgeImagePolyLineGraph_optimized(img, edge_color));
  return res;
}
#endif

// f2: polyline_matching_closeness_to_refpoints (polyline_matcher.hpp:64): pair(refpoint ids, one vector<set<ulong>> per match)
inline std::pair<std::vector<ulong_t>, std::vector<std::vector<std::set<ulong_t>>>> polyline_matching_closeness_to_refpoints(const Eg3dScene& scene) {
  eg3d_polyline_sets* h = nullptr;
  check(eg3d_polyline_sets_from_refpoints(&scene.desc(), 10.0f /* FIND_WITHIN_DIST */, 3.0f /* DETECTION_CORRESPONDENCES_MULTIPLICATION_FACTOR */, &h));
  eg3d_candidates c; int64_t n_ref = 0; const int64_t* ref = nullptr;
  check(eg3d_polyline_sets_get(h, &c, &n_ref, &ref));
  std::pair<std::vector<ulong_t>, std::vector<std::vector<std::set<ulong_t>>>> res;
  res.first.assign(ref, ref + n_ref);
  const int V = scene.n_views();
  res.second.assign((size_t)c.n_sets, std::vector<std::set<ulong_t>>((size_t)V));
  for (int i = 0; i < c.n_sets; i++)
    for (int v = 0; v < V; v++)
      for (int64_t k = c.off[(size_t)i * V + v]; k < c.off[(size_t)i * V + v + 1]; k++) res.second[(size_t)i][(size_t)v].insert(c.polyline[k]);
  eg3d_polyline_sets_free(h);
  return res;
}

// f2, pipeline 1: polyline_matching_similarity_graph (polyline_matcher.hpp:47).  The reference returns
// tuple(close_polylines, close_refpoints, matches); only the third element is consumed (pipelines.cpp:77), so that is what
// this returns: one vector<set<ulong>> per community.  compatibility_graph_file, when given, receives the text the
// reference writes for Grappolo (byte for byte); the communities themselves come from the library's deterministic Louvain.
inline std::vector<std::vector<std::set<ulong_t>>> polyline_matching_similarity_graph(const Eg3dScene& scene, std::string* compatibility_graph_file = nullptr) {
  eg3d_similarity_graph* g = nullptr;
  check(eg3d_polyline_similarity_graph(&scene.desc(), 10.0f /* FIND_WITHIN_DIST */, &g));
  eg3d_similarity_graph_view gv; check(eg3d_similarity_graph_get(g, &gv));
  if (compatibility_graph_file) compatibility_graph_file->assign(gv.dimacs, (size_t)gv.dimacs_len);
  std::vector<int64_t> com((size_t)gv.n_nodes);
  check(eg3d_similarity_graph_communities(g, com.data(), nullptr));
  eg3d_polyline_sets* h = nullptr;
  check(eg3d_polyline_sets_from_communities(g, com.data(), &h));
  eg3d_candidates c; check(eg3d_polyline_sets_get(h, &c, nullptr, nullptr));
  const int V = scene.n_views();
  std::vector<std::vector<std::set<ulong_t>>> res((size_t)c.n_sets, std::vector<std::set<ulong_t>>((size_t)V));
  for (int i = 0; i < c.n_sets; i++)
    for (int v = 0; v < V; v++)
      for (int64_t k = c.off[(size_t)i * V + v]; k < c.off[(size_t)i * V + v + 1]; k++) res[(size_t)i][(size_t)v].insert(c.polyline[k]);
  eg3d_polyline_sets_free(h);
  eg3d_similarity_graph_free(g);
  return res;
}

// add_3dpoints_to_sfmd (src/edgegraph3d/io/output/output_utilities.cpp:96-111)
inline void add_3dpoints_to_sfmd(SfMData& sfmd, const std::vector<new_3dpoint_plgp_matches>& p3ds) {
  if ((int)sfmd.pointsVisibleFromCamN_.size() < sfmd.numCameras_) sfmd.pointsVisibleFromCamN_.resize((size_t)sfmd.numCameras_);
  for (const auto& p : p3ds) {
    const int id = (int)sfmd.points_.size();
    sfmd.points_.push_back(std::get<0>(p));
    for (int cam : std::get<2>(p)) sfmd.pointsVisibleFromCamN_[(size_t)cam].push_back(id);
    std::vector<vec2> coords;
    for (const auto& q : std::get<1>(p)) coords.push_back(q.plp.coords);
    sfmd.point2DoncamViewingPoint_.push_back(coords);
    sfmd.camViewingPointN_.push_back(std::get<2>(p));
  }
  sfmd.numPoints_ = (int)sfmd.points_.size();
}

// edge_reconstruction_pipeline (src/edgegraph3d/matching/plg_matching/pipelines.cpp:201-246) without its image / timing /
// serialisation arguments: pipelines 1-3 in the reference's order, the density limiter, add_3dpoints_to_sfmd.
// `scene` must have been built from the same sfmd and plgs.  Returns first_edgepoint (the number of SfM points before the
// edge points were appended), the argument filter() takes next (edge_matcher.cpp:118-132).
inline int edge_reconstruction_pipeline(Eg3dScene& scene, SfMData& sfmd) {
  const int first_edgepoint = (int)sfmd.points_.size();
  std::vector<new_3dpoint_plgp_matches> p3ds =
      find_new_3d_points_from_compatible_polylines_expandallviews_parallel(scene, polyline_matching_similarity_graph(scene));
  const auto pmctr = polyline_matching_closeness_to_refpoints(scene);
  const auto p2 = find_new_3d_points_from_compatible_polylines_expandallviews_parallel(scene, pmctr.second);
  p3ds.insert(p3ds.end(), p2.begin(), p2.end());
  const auto p3 = plg_matching_from_refpoints_parallel(scene);
  p3ds.insert(p3ds.end(), p3.begin(), p3.end());
  add_3dpoints_to_sfmd(sfmd, filter_3d_points_close_2d_array(scene, p3ds));
  return first_edgepoint;
}

}  // namespace eg3d_shim
