/*
 * eg3d_oracle_api.cpp — CPU ORACLE (test infrastructure, NOT product code): pipeline drivers, density limiter,
 * outlier filter and the flat C interface used by tests/ and bench.py's cpu_baseline leg (via ctypes).
 * See eg3d_oracle.hpp.  References are to the EdgeGraph3D tree.
 */
#include "eg3d_oracle.hpp"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <omp.h>

namespace eg3d_oracle {

Scene* scene_from_desc(const eg3d_scene_desc* d, const eg3d_params* p) {
  Scene* s = new Scene();
  s->V = d->n_views; s->width = d->width; s->height = d->height;
  s->prm = *p;
  s->P.resize(s->V);
  for (int v = 0; v < s->V; v++) std::memcpy(s->P[v].data(), d->cameras + 12 * (size_t)v, 12 * sizeof(float));
  s->F.assign(d->fundamental, d->fundamental + (size_t)s->V * s->V * 9);
  s->Fvalid.assign(d->fundamental_valid, d->fundamental_valid + (size_t)s->V * s->V);
  s->plgs.resize(s->V);
  for (int v = 0; v < s->V; v++) {
    int64_t p0 = d->view_poly_off[v], p1 = d->view_poly_off[v + 1];
    s->plgs[v].resize(p1 - p0);
    for (int64_t g = p0; g < p1; g++) {
      Polyline& pl = s->plgs[v][g - p0];
      pl.start = d->poly_start[g]; pl.end = d->poly_end[g];
      for (int64_t k = d->poly_vert_off[g]; k < d->poly_vert_off[g + 1]; k++) pl.pc.push_back(V2{d->verts[2 * k], d->verts[2 * k + 1]});
    }
  }
  s->plmaps.resize(s->V); s->corr_maps.resize(s->V);
  const float corr = p->detection_starting_radius * p->detection_mult; /* plg_edge_manager.cpp:46 init list */
#pragma omp parallel for schedule(dynamic)
  for (int v = 0; v < s->V; v++) {
    s->plmaps[v].build(s->plgs[v], s->width, s->height, p->expand_grid_cell);
    if (d->n_tracks > 0) s->corr_maps[v].build(s->plgs[v], s->width, s->height, corr);
  }
  for (int64_t t = 0; t < d->n_tracks; t++) {
    s->points.push_back(V3{d->track_xyz[3 * t], d->track_xyz[3 * t + 1], d->track_xyz[3 * t + 2]});
    std::vector<int> vv; std::vector<V2> xy;
    for (int64_t k = d->track_off[t]; k < d->track_off[t + 1]; k++) { vv.push_back(d->track_view[k]); xy.push_back(V2{d->track_xy[2 * k], d->track_xy[2 * k + 1]}); }
    s->track_views.push_back(vv); s->track_xy.push_back(xy);
  }
  return s;
}

/* ---------------------------------------------------------------- result container ---- */
struct Points {
  std::vector<float> xyz; std::vector<int32_t> seed, chain_pos; std::vector<int64_t> obs_off{0};
  std::vector<int32_t> obs_view; std::vector<uint32_t> obs_poly, obs_seg; std::vector<float> obs_xy;
  std::vector<int64_t> seed_track; std::vector<int32_t> seed_set;   /* per seed: SfM point (pipeline 3) / candidate set (pipelines 1-2) */
  std::vector<uint8_t> seed_ub;   /* per seed: the reference's behaviour is undefined on it */
  void append(const std::vector<Match>& chain, int32_t seed_ord) {
    for (size_t k = 0; k < chain.size(); k++) {
      const Match& m = chain[k];
      xyz.push_back(m.X.x); xyz.push_back(m.X.y); xyz.push_back(m.X.z);
      seed.push_back(seed_ord); chain_pos.push_back((int32_t)k);
      for (size_t j = 0; j < m.obs.size(); j++) {
        obs_view.push_back(m.views[j]); obs_poly.push_back((uint32_t)m.obs[j].pl); obs_seg.push_back((uint32_t)m.obs[j].plp.seg);
        obs_xy.push_back(m.obs[j].plp.c.x); obs_xy.push_back(m.obs[j].plp.c.y);
      }
      obs_off.push_back((int64_t)obs_view.size());
    }
  }
};

struct SeedRec { int view; PlgPoint p; int cand_set; int64_t track = -1; };

static std::vector<std::vector<ulong_t>> cand_of(const Scene& s, const eg3d_candidates* c, int set) {
  std::vector<std::vector<ulong_t>> r(s.V);
  for (int v = 0; v < s.V; v++)
    for (int64_t k = c->off[(size_t)set * s.V + v]; k < c->off[(size_t)set * s.V + v + 1]; k++) r[v].push_back(c->polyline[k]);
  return r;
}

/* polyline_matching.cpp:134-144 for a batch of seeds */
static Points* run_seeds(const Scene& s, const std::vector<SeedRec>& seeds, const eg3d_candidates* c, int n_threads) {
  std::vector<std::vector<Match>> chains(seeds.size());
  std::vector<std::vector<std::vector<ulong_t>>> csets;
  if (c) for (int k = 0; k < c->n_sets; k++) csets.push_back(cand_of(s, c, k));
  if (n_threads < 1) n_threads = 1;
  std::vector<uint8_t> ub(seeds.size(), 0);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
  for (int64_t i = 0; i < (int64_t)seeds.size(); i++) {
    const SeedRec& sd = seeds[i];
    const std::vector<std::vector<ulong_t>>* cs = (c && sd.cand_set >= 0) ? &csets[sd.cand_set] : nullptr;
    auto epc = find_epipolar_correspondences(s, cs, sd.view, sd.p);
    { const char* t = getenv("EG3D_ORACLE_TRACE_SEED"); g_trace = (t && atoll(t) == i) ? 1 : 0; }
    const long long ub0 = g_foreign_direction_calls;
    chains[i] = compute_3D_point_multiple_views_plg_following_expandallviews_vector(s, sd.view, epc);
    ub[i] = g_foreign_direction_calls != ub0;
    g_trace = 0;
  }
  Points* out = new Points();
  for (size_t i = 0; i < seeds.size(); i++) out->append(chains[i], (int32_t)i);
  out->seed_ub = ub;
  for (size_t i = 0; i < seeds.size(); i++) { out->seed_set.push_back(seeds[i].cand_set); out->seed_track.push_back(-1); }
  return out;
}

/* seed sampler, polyline_matching.cpp:168-190 */
static void sample_polyline(const Polyline& pl, float spacing, std::vector<PlPoint>& out) {
  if (!pl.valid()) return;
  PlPoint plp = pl.get_start_plp();
  bool reached;
  plp = pl.next_pl_point_by_distance(plp, pl.end, spacing, reached);
  while (!reached) {
    out.push_back(plp);
    plp = pl.next_pl_point_by_distance(plp, pl.end, spacing, reached);
  }
}

/* ---------------------------------------------------------------- pipeline 3 (refpoints) ---- */
/* plg_edge_manager.cpp:261-300 + :246-259 + :208-243 + :191-205, then plgpcm_3views_plg_following.cpp:40-69 */
static void refpoint_seeds(const Scene& s, int64_t rp, std::vector<SeedRec>& seeds, std::vector<std::vector<std::vector<PlgPoint>>>& epcs) {
  const auto& views = s.track_views[rp];
  const auto& xy = s.track_xy[rp];
  const float start_dsq = s.prm.detection_starting_radius * s.prm.detection_starting_radius;
  const float corr_d = s.prm.detection_starting_radius * s.prm.detection_mult;
  const float corr_dsq = corr_d * corr_d;
  std::vector<std::vector<ulong_t>> pcp(views.size());
  std::vector<std::vector<PlgPoint>> sni(views.size());
  auto coords_on = [&](int img, V2& pt) { /* get_2d_coordinates_of_point_on_image, edge_graph_3d_utilities.cpp:395-403 (last match wins) */
    for (size_t i = 0; i < views.size(); i++) if (views[i] == img) pt = xy[i];
  };
  for (size_t i = 0; i < views.size(); i++) {
    V2 sp{0, 0}; coords_on(views[i], sp);
    for (ulong_t pl_id : s.corr_maps[views[i]].find_polylines_potentially_within_search_dist(sp)) {
      ulong_t cs; V2 proj;
      float dsq = s.plgs[views[i]][pl_id].compute_distancesq(sp, cs, proj);
      if (dsq <= start_dsq) { pcp[i].push_back(pl_id); sni[i].push_back(PlgPoint{pl_id, PlPoint{cs, proj}}); }
      else if (dsq <= corr_dsq) pcp[i].push_back(pl_id);
    }
  }
  for (size_t i = 0; i < views.size(); i++) {
    const int simg = views[i];
    V2 init{0, 0}; coords_on(simg, init);
    for (const auto& seed : sni[i]) {
      const float radius = compute_2d_distance(init, seed.plp.c) * s.prm.detection_mult;
      const float rsq = radius * radius;
      std::vector<std::vector<PlgPoint>> all(s.V);
      for (size_t j = 0; j < views.size(); j++) {
        const int img = views[j];
        std::vector<PlgPoint> cur;
        if (img != simg) {
          V3 epi;
          if (computeCorrespondEpilineSinglePoint(seed.plp.c, s.Fm(simg, img), s.Fok(simg, img), epi)) {
            for (ulong_t pl_id : pcp[j])
              for (const auto& plp : s.plgs[img][pl_id].intersect_line(epi))
                if (squared_2d_distance(xy[j], plp.c) <= rsq) cur.push_back(PlgPoint{pl_id, plp});
          }
        } else cur.push_back(seed);
        all[img] = cur; /* scatter to the V-vector, plgpcm_3views_plg_following.cpp:41-43 */
      }
      seeds.push_back(SeedRec{simg, seed, -1, rp});
      epcs.push_back(all);
    }
  }
}

/* ---------------------------------------------------------------- filter GN (FP32) ---- */
/* filtering/gauss_newton.cpp:83-134 + :26-75.  OpenCV CV_32F arithmetic as probed against cv2 4.13:
 *  4x4 * 4x1 : float products, float sequential accumulate;  J^T J and (H^-1 J^T) r : double sequential accumulate,
 *  cast to float;  determinant / inverse of the float 3x3 : double cofactor formula. */
int GaussNewton_f32(const Scene& s, const std::vector<int>& views, const std::vector<V2>& pts, const float init[3],
                    float out[3], float gn_max_mse, float* last_mse_out) {
  const int n = (int)pts.size();
  std::vector<float> r(2 * n), J(6 * n);
  float X[3] = {init[0], init[1], init[2]};
  float last_mse = 0;
  for (int it = 0; it < s.prm.gn_max_iters; it++) {
    float mse = 0;
    for (int m = 0; m < n; m++) {
      const float* P = s.P[views[m]].data();
      float h0 = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3] * 1.0f;
      float h1 = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7] * 1.0f;
      float h2 = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] * 1.0f;
      r[2 * m] = pts[m].x - h0 / h2;
      mse += r[2 * m] * r[2 * m];
      r[2 * m + 1] = pts[m].y - h1 / h2;
      mse += r[2 * m + 1] * r[2 * m + 1];
    }
    float diff = mse / (n * 2) - last_mse;
    bool stop = s.prm.filter_abs_int ? ((double)std::abs((int)diff) < (double)s.prm.filter_gn_stop)
                                     : ((double)std::abs(diff) < (double)s.prm.filter_gn_stop);
    if (stop) break;
    last_mse = mse / (n * 2);
    for (int m = 0; m < n; m++) {
      const float* P = s.P[views[m]].data();
      float xH = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3] * 1.0f;
      float yH = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7] * 1.0f;
      float zH = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] * 1.0f;
      float* j0 = &J[6 * m]; float* j1 = j0 + 3;
      j0[0] = (P[0] * zH - P[8] * xH) / (zH * zH);  j1[0] = (P[4] * zH - P[8] * yH) / (zH * zH);
      j0[1] = (P[1] * zH - P[9] * xH) / (zH * zH);  j1[1] = (P[5] * zH - P[9] * yH) / (zH * zH);
      j0[2] = (P[2] * zH - P[10] * xH) / (zH * zH); j1[2] = (P[6] * zH - P[10] * yH) / (zH * zH);
    }
    float H[9];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        double acc = 0;
        for (int k = 0; k < 2 * n; k++) acc += (double)J[3 * k + a] * (double)J[3 * k + b];
        H[3 * a + b] = (float)acc;
      }
    double m9[9]; for (int i = 0; i < 9; i++) m9[i] = H[i];
    double d = m9[0] * (m9[4] * m9[8] - m9[5] * m9[7]) - m9[1] * (m9[3] * m9[8] - m9[5] * m9[6]) + m9[2] * (m9[3] * m9[7] - m9[4] * m9[6]);
    float df = (float)d; /* `float d = cv::determinant(hessian)` */
    if ((double)df < (double)s.prm.filter_gn_det_min) { if (last_mse_out) *last_mse_out = last_mse; return -1; }
    float Hi[9];
    {
      double dd = d != 0 ? 1. / d : 0;
      Hi[0] = (float)((m9[4] * m9[8] - m9[5] * m9[7]) * dd); Hi[1] = (float)((m9[2] * m9[7] - m9[1] * m9[8]) * dd); Hi[2] = (float)((m9[1] * m9[5] - m9[2] * m9[4]) * dd);
      Hi[3] = (float)((m9[5] * m9[6] - m9[3] * m9[8]) * dd); Hi[4] = (float)((m9[0] * m9[8] - m9[2] * m9[6]) * dd); Hi[5] = (float)((m9[2] * m9[3] - m9[0] * m9[5]) * dd);
      Hi[6] = (float)((m9[3] * m9[7] - m9[4] * m9[6]) * dd); Hi[7] = (float)((m9[1] * m9[6] - m9[0] * m9[7]) * dd); Hi[8] = (float)((m9[0] * m9[4] - m9[1] * m9[3]) * dd);
    }
    for (int a = 0; a < 3; a++) {
      double acc = 0;
      for (int k = 0; k < 2 * n; k++) {
        float mk = (float)((double)Hi[3 * a + 0] * (double)J[3 * k + 0] + (double)Hi[3 * a + 1] * (double)J[3 * k + 1] + (double)Hi[3 * a + 2] * (double)J[3 * k + 2]);
        acc += (double)mk * (double)r[k];
      }
      X[a] += (float)acc;
    }
  }
  if (last_mse_out) *last_mse_out = last_mse;
  if (last_mse < gn_max_mse) { out[0] = X[0]; out[1] = X[1]; out[2] = X[2]; return 1; }
  return -1;
}

}  // namespace eg3d_oracle

/* ======================================================================= flat C interface ============== */
using namespace eg3d_oracle;

extern "C" {

void eg3d_oracle_params_default(eg3d_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->split_interval_distance = 20.0f; p->follow_first_image_distance = 10.0f;
  p->follow_corr_min = 5.0f; p->follow_corr_max = 20.0f;
  p->quasiparallel_cos = 0.965f; p->quasiparallel_dist = 5.0f;
  p->max_proj_distsq_expand = 16.0f; p->expand_grid_cell = 4.0f;
  p->detection_starting_radius = 10.0f; p->detection_mult = 3.0f;
  p->gn_max_iters = 30; p->gn_stop = 0.0000005; p->gn_det_min = 0.00001; p->gn_accept_mse = 9;
  p->filter_gn_stop = 0.0000000005; p->filter_gn_det_min = 0.0000000001; p->filter_gn_max_mse = 2.25f;
  p->filter_3views_amount = 3; p->dedup_cell = 3.0f; p->dlt_wellposed = 2; p->filter_abs_int = 0;
  p->max_chain_points = 96; p->max_follow_points = 160;
}

void* eg3d_oracle_scene_create(const eg3d_scene_desc* d, const eg3d_params* p) { return scene_from_desc(d, p); }
void eg3d_oracle_scene_destroy(void* s) { delete (Scene*)s; }

/* --- primitives exposed for unit tests / golden checks --- */
int eg3d_oracle_epiline(const double* F9, float x, float y, float* out3) {
  V3 l; bool ok = computeCorrespondEpilineSinglePoint(V2{x, y}, F9, true, l);
  out3[0] = l.x; out3[1] = l.y; out3[2] = l.z; return ok;
}
/* (calls, degenerate calls) of the 2-view DLT since the last reset; reset = 1 clears the counters after reading */
void eg3d_oracle_dlt_stats(long long* calls, long long* degenerate, int reset) {
  *calls = g_dlt_calls.load(); *degenerate = g_dlt_degenerate.load();
  if (reset) { g_dlt_calls = 0; g_dlt_degenerate = 0; }
}
void eg3d_oracle_triangulate_dlt(const float* P1, const float* P2, const float* x1, const float* x2, float* out4) {
  triangulate_dlt(P1, P2, V2{x1[0], x1[1]}, V2{x2[0], x2[1]}, out4);
}
void eg3d_oracle_compute_projection(const float* cam12, const float* x3, float* out2) {
  V2 q = compute_projection(cam12, V3{x3[0], x3[1], x3[2]}); out2[0] = q.x; out2[1] = q.y;
}
void eg3d_oracle_triangulate_dlt_opencv(const float* P1, const float* P2, const float* x1, const float* x2, float* out4) {
  triangulate_dlt_opencv(P1, P2, V2{x1[0], x1[1]}, V2{x2[0], x2[1]}, out4);
}
int eg3d_oracle_intersect_segment_line(const float* segm4, const float* line3, float* inter2) {
  bool f; V2 p{0, 0}; intersect_segment_line(segm4, V3{line3[0], line3[1], line3[2]}, f, p);
  inter2[0] = p.x; inter2[1] = p.y; return f;
}
/* returns bit0 = intersection_found, bit1 = quasiparallel_within_distance */
int eg3d_oracle_intersect_segment_line_nqp(const float* segm4, const float* line3, float qcos, float qdist, float* inter2) {
  bool f, q; V2 p{0, 0}; intersect_segment_line_no_quasiparallel(segm4, V3{line3[0], line3[1], line3[2]}, qcos, qdist, f, q, p);
  inter2[0] = p.x; inter2[1] = p.y; return (f ? 1 : 0) | (q ? 2 : 0);
}
float eg3d_oracle_squared_2d_distance(float ax, float ay, float bx, float by) { return squared_2d_distance(V2{ax, ay}, V2{bx, by}); }
/* grid query on the 4 px (which=0) or 30 px (which=1) grid; returns count, ids into out (cap) */
int eg3d_oracle_grid_query(void* sc, int view, int which, float x, float y, uint32_t* out, int cap) {
  Scene* s = (Scene*)sc;
  auto r = (which ? s->corr_maps : s->plmaps)[view].find_polylines_potentially_within_search_dist(V2{x, y});
  for (int i = 0; i < (int)r.size() && i < cap; i++) out[i] = (uint32_t)r[i];
  return (int)r.size();
}
/* polyline walkers: kind 0 = by distance, 1 = line intersection, 2 = bounded. returns flags bit0 found / reached */
int eg3d_oracle_walk(void* sc, int view, uint32_t pl_id, int kind, uint32_t seg, float x, float y, int towards_end,
                     const float* line3, float dist, uint32_t* out_seg, float* out_xy) {
  Scene* s = (Scene*)sc;
  const Polyline& pl = s->plgs[view][pl_id];
  PlPoint init{seg, V2{x, y}}, next{seg, V2{x, y}};
  ulong_t dir = towards_end ? pl.end : pl.start;
  int flag = 0;
  if (kind == 0) { bool reached; next = pl.next_pl_point_by_distance(init, dir, dist, reached); flag = reached; }
  else if (kind == 1) { bool f; pl.next_pl_point_by_line_intersection(init, dir, V3{line3[0], line3[1], line3[2]}, s->prm.quasiparallel_cos, s->prm.quasiparallel_dist, next, f); flag = f; }
  else { bool f; pl.next_pl_point_by_line_intersection_bounded_distance(init, dir, V3{line3[0], line3[1], line3[2]}, s->prm.quasiparallel_cos, s->prm.quasiparallel_dist, s->prm.follow_corr_min, s->prm.follow_corr_max, next, f); flag = f; }
  *out_seg = (uint32_t)next.seg; out_xy[0] = next.c.x; out_xy[1] = next.c.y;
  return flag;
}
float eg3d_oracle_polyline_distancesq(void* sc, int view, uint32_t pl_id, float x, float y, uint32_t* seg, float* proj) {
  Scene* s = (Scene*)sc; ulong_t cs; V2 p;
  float d = s->plgs[view][pl_id].compute_distancesq(V2{x, y}, cs, p);
  *seg = (uint32_t)cs; proj[0] = p.x; proj[1] = p.y; return d;
}

/* --- seed sampler (same contract as eg3d_sample_seeds) --- */
int eg3d_oracle_sample_seeds(const eg3d_scene_desc* d, const int32_t* views, const uint32_t* polylines, int64_t n_pl, float spacing,
                             int64_t capacity, int32_t* o_view, uint32_t* o_pl, uint32_t* o_seg, float* o_xy, int32_t* o_src, int64_t* n_out) {
  int64_t n = 0;
  for (int64_t k = 0; k < n_pl; k++) {
    int64_t g = d->view_poly_off[views[k]] + polylines[k];
    Polyline pl; pl.start = d->poly_start[g]; pl.end = d->poly_end[g];
    for (int64_t j = d->poly_vert_off[g]; j < d->poly_vert_off[g + 1]; j++) pl.pc.push_back(V2{d->verts[2 * j], d->verts[2 * j + 1]});
    std::vector<PlPoint> pts; sample_polyline(pl, spacing, pts);
    for (const auto& p : pts) {
      if (n < capacity) { o_view[n] = views[k]; o_pl[n] = polylines[k]; o_seg[n] = (uint32_t)p.seg; o_xy[2 * n] = p.c.x; o_xy[2 * n + 1] = p.c.y; if (o_src) o_src[n] = (int32_t)k; }
      n++;
    }
  }
  *n_out = n;
  return n <= capacity ? 0 : 4;
}

/* --- K1 --- */
struct OHits { std::vector<int64_t> off; std::vector<eg3d_hit> hits; int64_t n_seeds; int V; };
void* eg3d_oracle_epipolar_intersect(void* sc, const eg3d_seeds* seeds, const eg3d_candidates* c, int n_threads) {
  Scene* s = (Scene*)sc;
  std::vector<std::vector<std::vector<ulong_t>>> csets;
  if (c) for (int k = 0; k < c->n_sets; k++) csets.push_back(cand_of(*s, c, k));
  std::vector<std::vector<std::vector<PlgPoint>>> all(seeds->n);
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
  for (int64_t i = 0; i < seeds->n; i++) {
    int cs = (c && seeds->cand_set) ? seeds->cand_set[i] : -1;
    PlgPoint p{seeds->polyline[i], PlPoint{seeds->segment[i], V2{seeds->xy[2 * i], seeds->xy[2 * i + 1]}}};
    all[i] = find_epipolar_correspondences(*s, cs >= 0 ? &csets[cs] : nullptr, seeds->view[i], p);
  }
  OHits* h = new OHits(); h->n_seeds = seeds->n; h->V = s->V; h->off.push_back(0);
  for (int64_t i = 0; i < seeds->n; i++)
    for (int v = 0; v < s->V; v++) {
      for (const auto& q : all[i][v]) h->hits.push_back(eg3d_hit{(uint32_t)q.pl, (uint32_t)q.plp.seg, q.plp.c.x, q.plp.c.y});
      h->off.push_back((int64_t)h->hits.size());
    }
  return h;
}
void eg3d_oracle_hits_get(void* hh, int64_t* n_seeds, int32_t* n_views, const int64_t** off, const eg3d_hit** hits) {
  OHits* h = (OHits*)hh; *n_seeds = h->n_seeds; *n_views = h->V; *off = h->off.data(); *hits = h->hits.data();
}
void eg3d_oracle_hits_free(void* h) { delete (OHits*)h; }

/* --- B1/B4 per seed batch --- */
void* eg3d_oracle_match_seeds(void* sc, const eg3d_seeds* seeds, const eg3d_candidates* c, int n_threads) {
  Scene* s = (Scene*)sc;
  std::vector<SeedRec> v(seeds->n);
  for (int64_t i = 0; i < seeds->n; i++)
    v[i] = SeedRec{seeds->view[i], PlgPoint{seeds->polyline[i], PlPoint{seeds->segment[i], V2{seeds->xy[2 * i], seeds->xy[2 * i + 1]}}},
                   (c && seeds->cand_set) ? seeds->cand_set[i] : -1};
  return run_seeds(*s, v, c, n_threads);
}

/* --- B1 batched over candidate sets: polyline_matching.cpp:153-201 called per match as pipelines.cpp:92-100 --- */
void* eg3d_oracle_match_polyline_sets(void* sc, const eg3d_candidates* c, int32_t view_begin, int32_t view_end, int n_threads) {
  Scene* s = (Scene*)sc;
  std::vector<SeedRec> seeds;
  for (int set = 0; set < c->n_sets; set++)
    for (int v = view_begin; v < view_end; v++)
      for (int64_t k = c->off[(size_t)set * s->V + v]; k < c->off[(size_t)set * s->V + v + 1]; k++) {
        const ulong_t pl_id = c->polyline[k];
        std::vector<PlPoint> pts; sample_polyline(s->plgs[v][pl_id], s->prm.split_interval_distance, pts);
        for (const auto& p : pts) seeds.push_back(SeedRec{v, PlgPoint{pl_id, p}, set});
      }
  return run_seeds(*s, seeds, c, n_threads);
}

/* --- B2: plg_matching_from_refpoints.cpp:64-104 --- */
void* eg3d_oracle_match_refpoints(void* sc, int64_t tb, int64_t te, int n_threads) {
  Scene* s = (Scene*)sc;
  std::vector<SeedRec> seeds; std::vector<std::vector<std::vector<PlgPoint>>> epcs;
  for (int64_t rp = tb; rp < te; rp++) refpoint_seeds(*s, rp, seeds, epcs);
  std::vector<std::vector<Match>> chains(seeds.size());
  if (n_threads < 1) n_threads = 1;
  std::vector<uint8_t> ub(seeds.size(), 0);
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
  for (int64_t i = 0; i < (int64_t)seeds.size(); i++) {
    { const char* t = getenv("EG3D_ORACLE_TRACE_SEED"); g_trace = (t && atoll(t) == i) ? 1 : 0; }
    const long long ub0 = g_foreign_direction_calls;
    chains[i] = compute_3D_point_multiple_views_plg_following_expandallviews_vector(*s, seeds[i].view, epcs[i]);
    ub[i] = g_foreign_direction_calls != ub0;
    g_trace = 0;
  }
  Points* out = new Points();
  for (size_t i = 0; i < seeds.size(); i++) out->append(chains[i], (int32_t)i);
  out->seed_ub = ub;
  for (size_t i = 0; i < seeds.size(); i++) { out->seed_track.push_back(seeds[i].track); out->seed_set.push_back(-1); }
  return out;
}
/* the seeds + hit lists pipeline 3 feeds to the consensus manager (for K1-level parity of the refpoint variant) */
void* eg3d_oracle_refpoint_hits(void* sc, int64_t tb, int64_t te, int64_t* n_seeds_out) {
  Scene* s = (Scene*)sc;
  std::vector<SeedRec> seeds; std::vector<std::vector<std::vector<PlgPoint>>> epcs;
  for (int64_t rp = tb; rp < te; rp++) refpoint_seeds(*s, rp, seeds, epcs);
  OHits* h = new OHits(); h->n_seeds = (int64_t)seeds.size(); h->V = s->V; h->off.push_back(0);
  for (size_t i = 0; i < seeds.size(); i++)
    for (int v = 0; v < s->V; v++) {
      for (const auto& q : epcs[i][v]) h->hits.push_back(eg3d_hit{(uint32_t)q.pl, (uint32_t)q.plp.seg, q.plp.c.x, q.plp.c.y});
      h->off.push_back((int64_t)h->hits.size());
    }
  *n_seeds_out = h->n_seeds;
  return h;
}

int eg3d_oracle_points_get(void* pp, eg3d_points_view* v) {
  Points* p = (Points*)pp;
  v->n_points = (int64_t)p->seed.size(); v->n_obs = (int64_t)p->obs_view.size();
  v->xyz = p->xyz.data(); v->seed = p->seed.data(); v->chain_pos = p->chain_pos.data(); v->obs_off = p->obs_off.data();
  v->obs_view = p->obs_view.data(); v->obs_poly = p->obs_poly.data(); v->obs_seg = p->obs_seg.data(); v->obs_xy = p->obs_xy.data();
  return 0;
}
/* per seed of the call: ub = the reference's behaviour is undefined on it, the SfM point (pipeline 3) or the
 * candidate set (pipelines 1-2) it belongs to */
int64_t eg3d_oracle_points_seed_info(void* pp, const uint8_t** ub, const int64_t** track, const int32_t** cand_set) {
  Points* p = (Points*)pp;
  *ub = p->seed_ub.data(); *track = p->seed_track.data(); *cand_set = p->seed_set.data();
  return (int64_t)p->seed_ub.size();
}
void eg3d_oracle_points_free(void* p) { delete (Points*)p; }

/* --- K2 alone --- */
int eg3d_oracle_gn_triangulate(void* sc, int64_t n, const int64_t* obs_off, const int32_t* obs_view, const float* obs_xy,
                               const float* init_xyz, int fp64, float* out_xyz, float* out_mse, uint8_t* out_ok, int n_threads) {
  Scene* s = (Scene*)sc;
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
  for (int64_t i = 0; i < n; i++) {
    std::vector<int> views; std::vector<V2> pts;
    for (int64_t k = obs_off[i]; k < obs_off[i + 1]; k++) { views.push_back(obs_view[k]); pts.push_back(V2{obs_xy[2 * k], obs_xy[2 * k + 1]}); }
    int ok;
    if (fp64) {
      double init[3] = {init_xyz[3 * i], init_xyz[3 * i + 1], init_xyz[3 * i + 2]}, o[3] = {init[0], init[1], init[2]}, lm = 0;
      ok = em_GaussNewton(*s, views, pts, init, o, &lm);
      out_xyz[3 * i] = (float)o[0]; out_xyz[3 * i + 1] = (float)o[1]; out_xyz[3 * i + 2] = (float)o[2]; out_mse[i] = (float)lm;
    } else {
      float o[3] = {init_xyz[3 * i], init_xyz[3 * i + 1], init_xyz[3 * i + 2]}, lm = 0;
      ok = GaussNewton_f32(*s, views, pts, &init_xyz[3 * i], o, s->prm.filter_gn_max_mse, &lm);
      out_xyz[3 * i] = o[0]; out_xyz[3 * i + 1] = o[1]; out_xyz[3 * i + 2] = o[2]; out_mse[i] = lm;
    }
    out_ok[i] = ok == 1;
  }
  return 0;
}

/* --- a13: filter_3d_points_close_2d_array, filtering_close_plgps.cpp:75-124 --- */
int eg3d_oracle_dedup_close_points(void* sc, const eg3d_points_view* pts, uint8_t* keep) {
  Scene* s = (Scene*)sc;
  const float cs = s->prm.dedup_cell;
  int w = (int)std::ceil((float)s->width / cs), h = (int)std::ceil((float)s->height / cs);
  std::vector<std::vector<uint8_t>> bm(s->V, std::vector<uint8_t>((size_t)w * h, 0));
  auto cell = [&](int64_t o) -> size_t { return (size_t)(int)(pts->obs_xy[2 * o + 1] / cs) * w + (size_t)(int)(pts->obs_xy[2 * o] / cs); };
  for (int64_t i = 0; i < pts->n_points; i++) {
    bool is_new = false;
    for (int64_t o = pts->obs_off[i]; o < pts->obs_off[i + 1]; o++)
      if (!bm[pts->obs_view[o]][cell(o)]) { is_new = true; break; }
    keep[i] = is_new;
    if (is_new) for (int64_t o = pts->obs_off[i]; o < pts->obs_off[i + 1]; o++) bm[pts->obs_view[o]][cell(o)] = 1;
  }
  return 0;
}

/* --- a14: filter(), outliers_filtering.cpp:14-64 + gaussNewtonFiltering gauss_newton.cpp:136-178 --- */
int eg3d_oracle_filter(void* sc, int64_t n, float* xyz, const int64_t* obs_off, const int32_t* obs_view, const float* obs_xy,
                       int64_t first_edgepoint, float gn_max_mse, int32_t forced_min_filter, uint8_t* inliers, int n_threads) {
  Scene* s = (Scene*)sc;
  if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
  for (int64_t i = 0; i < n; i++) {
    std::vector<int> views; std::vector<V2> p;
    for (int64_t k = obs_off[i]; k < obs_off[i + 1]; k++) { views.push_back(obs_view[k]); p.push_back(V2{obs_xy[2 * k], obs_xy[2 * k + 1]}); }
    float o[3];
    if (GaussNewton_f32(*s, views, p, &xyz[3 * i], o, gn_max_mse, nullptr) != -1) { xyz[3 * i] = o[0]; xyz[3 * i + 1] = o[1]; xyz[3 * i + 2] = o[2]; inliers[i] = 1; }
    else inliers[i] = 0;
  }
  /* compute_ray_stats, outliers_filtering.cpp:14-35 */
  std::vector<int64_t> dist(s->V, 0);
  int64_t count = 0;
  for (int64_t i = 0; i < n; i++) if (inliers[i]) { count++; int64_t len = obs_off[i + 1] - obs_off[i]; if (len >= 1 && len <= s->V) dist[len - 1]++; }
  int median = 0; int64_t m = 0;
  for (median = 0; median < s->V; median++) { m += dist[median]; if (m >= count / 2) break; }
  int intended = (s->prm.filter_3views_amount >= median / 2 - 1) ? s->prm.filter_3views_amount : (median / 2 - 1);
  if (forced_min_filter > -1) intended = forced_min_filter;
  for (int64_t i = first_edgepoint; i < n; i++) inliers[i] = inliers[i] && ((obs_off[i + 1] - obs_off[i]) > intended);
  return 0;
}

}  // extern "C"
