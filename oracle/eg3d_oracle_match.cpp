/*
 * eg3d_oracle_match.cpp — CPU ORACLE (test infrastructure, NOT product code): Gauss-Newton triangulation,
 * PLG following, view expansion and the per-seed entry.  See eg3d_oracle.hpp.  References are to the
 * EdgeGraph3D tree (SURVEY.md Appendix B is the cite-checked pseudo-code of this file).
 */
#include "eg3d_oracle.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <tuple>

#include <atomic>
namespace eg3d_oracle {

/* ============================================================ 2-view DLT initialiser ====================== */
/* cv::triangulatePoints (triangulation.cpp:216,290): rows x*P[2]-P[0], y*P[2]-P[1] for both cameras in double,
 * null vector = last right-singular vector, cast to float32 (pinned against cv2 in tests/golden).
 * SVD here = one-sided Jacobi (Hestenes) on the 4x4 matrix. */
void triangulate_dlt(const float P1[12], const float P2[12], const V2& x1, const V2& x2, float out4[4]) {
  double A[4][4];
  const float* Ps[2] = {P1, P2};
  const V2 xs[2] = {x1, x2};
  for (int j = 0; j < 2; j++) {
    double x = xs[j].x, y = xs[j].y;
    for (int k = 0; k < 4; k++) {
      A[j * 2 + 0][k] = x * (double)Ps[j][8 + k] - (double)Ps[j][k];
      A[j * 2 + 1][k] = y * (double)Ps[j][8 + k] - (double)Ps[j][4 + k];
    }
  }
  double Vm[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < 3; p++)
      for (int q = p + 1; q < 4; q++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 4; i++) { alpha += A[i][p] * A[i][p]; beta += A[i][q] * A[i][q]; gamma += A[i][p] * A[i][q]; }
        if (gamma == 0) continue;
        off = std::max(off, std::fabs(gamma) / std::sqrt(alpha * beta + 1e-300));
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 4; i++) {
          double ap = A[i][p], aq = A[i][q];
          A[i][p] = c * ap - s * aq; A[i][q] = s * ap + c * aq;
          double vp = Vm[i][p], vq = Vm[i][q];
          Vm[i][p] = c * vp - s * vq; Vm[i][q] = s * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
  int best = 0; double bestn = 1e300;
  for (int j = 0; j < 4; j++) {
    double nn = 0;
    for (int i = 0; i < 4; i++) nn += A[i][j] * A[i][j];
    if (nn < bestn) { bestn = nn; best = j; }
  }
  for (int i = 0; i < 4; i++) out4[i] = (float)Vm[i][best];
}

/* The same initialiser with OpenCV's own SVD restated: cv::triangulatePoints (OpenCV 4.x, modules/calib3d/src/triangulate.cpp:
 * the 4x4 system above in double) -> cv::SVD::compute -> JacobiSVDImpl_<double> (modules/core/src/lapack.cpp) run on the
 * TRANSPOSED matrix: rows = columns of A, squared norms cached in W, pairs (i<j) skipped when |p| <= eps*sqrt(a*b) with
 * eps = DBL_EPSILON*10, rotation from hypot(2p, a-b), at most max(m,30) sweeps, singular values sorted descending by
 * selection sort with the rows of Vt swapped along; result = last row of Vt, narrowed to float.  OpenCV is not vendored in
 * the reference; this is the published algorithm, pinned against REAL cv2 4.13 calls: bit-identical (sign included) on the
 * 400 golden cases and on > 99.9 % of live inputs incl. degenerate ones (same camera twice), within a few float ulps on the
 * rest (hypot, below); there the null space is a whole ray and every other SVD returns a different point of it
 * (tests/test_oracle_golden.py).  eg3d_params.dlt_wellposed == 2 selects it. */
/* hypot from IEEE operations only, correctly rounded (sqrt of the double-double sum of squares + one exact-residual
 * correction), so that this oracle and the device code agree bit for bit.  OpenCV calls the C library's hypot, which in
 * glibc is accurate to ~0.8 ulp, not correctly rounded, and differs between its FMA / non-FMA builds: in 4 of 6 000 live
 * cases (all fully degenerate) that moves the float result by a few ulps (<= 3e-7 relative) — the noise floor of "identical". */
static double hypot_cr(double x, double y) {
  x = std::fabs(x); y = std::fabs(y);
  if (x < y) std::swap(x, y);
  if (y == 0) return x;
  int ex;
  std::frexp(x, &ex);
  x = std::ldexp(x, -ex); y = std::ldexp(y, -ex);
  const double x2 = x * x, xx = std::fma(x, x, -x2), y2 = y * y, yy = std::fma(y, y, -y2);
  const double s = x2 + y2, e = y2 - (s - x2);
  const double lo = (xx + yy) + e;
  const double h = std::sqrt(s);
  return std::ldexp(h + (std::fma(-h, h, s) + lo) / (2 * h), ex);
}
void triangulate_dlt_opencv(const float P1[12], const float P2[12], const V2& x1, const V2& x2, float out4[4]) {
  double At[4][4], Vt[4][4], W[4];   /* At[i][k] = A[k][i] */
  const float* Ps[2] = {P1, P2};
  const V2 xs[2] = {x1, x2};
  for (int j = 0; j < 2; j++) {
    const double x = xs[j].x, y = xs[j].y;
    for (int k = 0; k < 4; k++) {
      At[k][j * 2 + 0] = x * (double)Ps[j][8 + k] - (double)Ps[j][k];
      At[k][j * 2 + 1] = y * (double)Ps[j][8 + k] - (double)Ps[j][4 + k];
    }
  }
  const int n = 4, m = 4;
  const double eps = 2.220446049250313e-16 * 10;
  for (int i = 0; i < n; i++) {
    double sd = 0;
    for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
    W[i] = sd;
    for (int k = 0; k < n; k++) Vt[i][k] = 0;
    Vt[i][i] = 1;
  }
  for (int iter = 0; iter < 30; iter++) {
    bool changed = false;
    for (int i = 0; i < n - 1; i++)
      for (int j = i + 1; j < n; j++) {
        double a = W[i], p = 0, b = W[j];
        for (int k = 0; k < m; k++) p += At[i][k] * At[j][k];
        if (std::fabs(p) <= eps * std::sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = hypot_cr(p, beta);
        double c, sn;
        if (beta < 0) { const double delta = (gamma - beta) * 0.5; sn = std::sqrt(delta / gamma); c = p / (gamma * sn * 2); }
        else { c = std::sqrt((gamma + beta) / (gamma * 2)); sn = p / (gamma * c * 2); }
        a = b = 0;
        for (int k = 0; k < m; k++) {
          const double t0 = c * At[i][k] + sn * At[j][k], t1 = -sn * At[i][k] + c * At[j][k];
          At[i][k] = t0; At[j][k] = t1;
          a += t0 * t0; b += t1 * t1;
        }
        W[i] = a; W[j] = b;
        changed = true;
        for (int k = 0; k < n; k++) {
          const double t0 = c * Vt[i][k] + sn * Vt[j][k], t1 = -sn * Vt[i][k] + c * Vt[j][k];
          Vt[i][k] = t0; Vt[j][k] = t1;
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; i++) {
    double sd = 0;
    for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
    W[i] = std::sqrt(sd);
  }
  for (int i = 0; i < n - 1; i++) {
    int j = i;
    for (int k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
    if (i != j) {
      std::swap(W[i], W[j]);
      for (int k = 0; k < m; k++) std::swap(At[i][k], At[j][k]);
      for (int k = 0; k < n; k++) std::swap(Vt[i][k], Vt[j][k]);
    }
  }
  for (int k = 0; k < 4; k++) out4[k] = (float)Vt[3][k];
}

/* ============================================================ em_GaussNewton (FP64) ======================= */
static inline void cam4(const Scene& s, int view, double P[12]) {
  for (int i = 0; i < 12; i++) P[i] = (double)s.P[view][i]; /* float -> CV_64F, triangulation.cpp:301-306 */
}
static inline double det3(const double m[9]) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
/* cv::invert 3x3 closed form (cofactors * 1/det), as probed against cv2.invert */
static inline void inv3(const double m[9], double d, double t[9]) {
  d = 1. / d;
  t[0] = (m[4] * m[8] - m[5] * m[7]) * d; t[1] = (m[2] * m[7] - m[1] * m[8]) * d; t[2] = (m[1] * m[5] - m[2] * m[4]) * d;
  t[3] = (m[5] * m[6] - m[3] * m[8]) * d; t[4] = (m[0] * m[8] - m[2] * m[6]) * d; t[5] = (m[2] * m[3] - m[0] * m[5]) * d;
  t[6] = (m[3] * m[7] - m[4] * m[6]) * d; t[7] = (m[1] * m[6] - m[0] * m[7]) * d; t[8] = (m[0] * m[4] - m[1] * m[3]) * d;
}

/* triangulation.cpp:105-176 (+ em_point2D3DJacobian :53-103).  Returns 1 / -1. */
thread_local int g_trace = 0;   /* EG3D_ORACLE_TRACE_SEED=<ordinal>: print the margins of every GN call of that seed (debug aid) */
int em_GaussNewton(const Scene& s, const std::vector<int>& views, const std::vector<V2>& pts, const double init[3],
                   double out[3], double* last_mse_out) {
  const int n = (int)pts.size();
  double tr_min_stop = 1e300, tr_min_det = 1e300; int tr_it = 0;
  std::vector<double> r(2 * n), J(6 * n), cams(12 * n);
  for (int m = 0; m < n; m++) cam4(s, views[m], &cams[12 * m]);
  double X[3] = {init[0], init[1], init[2]};
  double last_mse = 0;
  for (int it = 0; it < s.prm.gn_max_iters; it++) {
    double mse = 0;
    for (int m = 0; m < n; m++) {
      const double* P = &cams[12 * m];
      double h0 = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3] * 1.0;
      double h1 = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7] * 1.0;
      double h2 = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] * 1.0;
      r[2 * m] = (double)pts[m].x - h0 / h2;
      mse += r[2 * m] * r[2 * m];
      r[2 * m + 1] = (double)pts[m].y - h1 / h2;
      mse += r[2 * m + 1] * r[2 * m + 1];
    }
    if (g_trace) { tr_it = it; tr_min_stop = std::min(tr_min_stop, std::abs(std::abs(mse / (n * 2) - last_mse) - s.prm.gn_stop)); }
    if (std::abs(mse / (n * 2) - last_mse) < s.prm.gn_stop) break;
    last_mse = mse / (n * 2);
    for (int m = 0; m < n; m++) {
      const double* P = &cams[12 * m];
      double xH = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3] * 1.0;
      double yH = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7] * 1.0;
      double zH = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] * 1.0;
      double* j0 = &J[6 * m]; double* j1 = j0 + 3;
      j0[0] = (P[0] * zH - P[8] * xH) / (zH * zH);  j1[0] = (P[4] * zH - P[8] * yH) / (zH * zH);
      j0[1] = (P[1] * zH - P[9] * xH) / (zH * zH);  j1[1] = (P[5] * zH - P[9] * yH) / (zH * zH);
      j0[2] = (P[2] * zH - P[10] * xH) / (zH * zH); j1[2] = (P[6] * zH - P[10] * yH) / (zH * zH);
    }
    double H[9];
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        double acc = 0;
        for (int k = 0; k < 2 * n; k++) acc += J[3 * k + a] * J[3 * k + b];
        H[3 * a + b] = acc;
      }
    double d = det3(H);
    if (g_trace) tr_min_det = std::min(tr_min_det, d);
    if (d < s.prm.gn_det_min) { if (g_trace) fprintf(stderr, "[gn] n=%d it=%d DET FAIL det=%.3e last_mse=%.6g views0=%d init=(%.4f %.4f %.4f)\n", n, it, d, last_mse, views[0], init[0], init[1], init[2]); if (last_mse_out) *last_mse_out = last_mse; return -1; }
    double Hi[9]; inv3(H, d, Hi);
    /* curEstimate += H.inv() * J.t() * r  evaluates (H^-1 J^T) first, then times r (MatExpr order) */
    for (int a = 0; a < 3; a++) {
      double acc = 0;
      for (int k = 0; k < 2 * n; k++) {
        double mk = Hi[3 * a + 0] * J[3 * k + 0] + Hi[3 * a + 1] * J[3 * k + 1] + Hi[3 * a + 2] * J[3 * k + 2];
        acc += mk * r[k];
      }
      X[a] += acc;
    }
  }
  if (g_trace) {
    fprintf(stderr, "[gn] n=%d its=%d last_mse=%.9g min|stop margin|=%.3e min det=%.3e lastview=%d init=(%.4f %.4f %.4f) X=(%.5f %.5f %.5f) %s\n", n, tr_it, last_mse, tr_min_stop, tr_min_det, views[n - 1],
            init[0], init[1], init[2], X[0], X[1], X[2], last_mse < s.prm.gn_accept_mse ? "ok" : "REJECT");
  }
  if (last_mse_out) *last_mse_out = last_mse;
  if (last_mse < s.prm.gn_accept_mse) { out[0] = X[0]; out[1] = X[1]; out[2] = X[2]; return 1; }
  return -1;
}

/* triangulation.cpp:178-250 / 252-323.  get_min_max (edge_graph_3d_utilities.hpp:69-92): min = first arg-min,
 * "max" = ALWAYS the last index (missing braces). */
/* statistics for the tests: how often the quirk hands the DLT the same camera twice (eg3d_oracle_dlt_stats) */
std::atomic<long long> g_dlt_calls{0}, g_dlt_degenerate{0};

void em_estimate3Dpositions(const Scene& s, const std::vector<V2>& coords, const std::vector<int>& ids, V3& Xo, bool& valid) {
  int min_index = 0;
  for (int i = 0; i < (int)ids.size(); i++) if (ids[i] < ids[min_index]) min_index = i;
  int max_index = (int)ids.size() - 1;
  g_dlt_calls.fetch_add(1, std::memory_order_relaxed);
  if (ids[max_index] == ids[min_index]) g_dlt_degenerate.fetch_add(1, std::memory_order_relaxed);
  if (s.prm.dlt_wellposed == 1 && ids[max_index] == ids[min_index]) {
    for (int j = (int)ids.size() - 1; j >= 0; j--) if (ids[j] != ids[min_index]) { max_index = j; break; }
  }
  float t4[4];
  /* always OpenCV's own SVD (cv::triangulatePoints restated): with dlt_wellposed != 1 this is the reference linked against
   * OpenCV 4.x, quirk and all; with 1 the same arithmetic on a well-posed camera pair */
  triangulate_dlt_opencv(s.P[ids[min_index]].data(), s.P[ids[max_index]].data(), coords[min_index], coords[max_index], t4);
  double init[3] = {(double)(t4[0] / t4[3]), (double)(t4[1] / t4[3]), (double)(t4[2] / t4[3])}; /* Vec4f / float */
  double out[3];
  if (em_GaussNewton(s, ids, coords, init, out, nullptr) != -1) {
    Xo = V3{(float)out[0], (float)out[1], (float)out[2]};
    valid = true;
  } else valid = false;
}

/* triangulation.cpp:347-405 / 408-466: GN warm-started at the current X with one more observation */
void em_add_new_observation_to_3Dpositions(const Scene& s, const V3& X0, const std::vector<V2>& coords, const std::vector<int>& ids,
                                           const V2& new_coords, int new_view, V3& Xo, bool& valid) {
  std::vector<V2> c = coords; std::vector<int> v = ids;
  c.push_back(new_coords); v.push_back(new_view);
  double init[3] = {X0.x, X0.y, X0.z}, out[3];
  if (em_GaussNewton(s, v, c, init, out, nullptr) != -1) {
    Xo = V3{(float)out[0], (float)out[1], (float)out[2]};
    valid = true;
  } else valid = false;
}

static std::vector<V2> coords_of(const std::vector<PlgPoint>& v) {
  std::vector<V2> r; r.reserve(v.size());
  for (const auto& p : v) r.push_back(p.plp.c);
  return r;
}

/* triangulation.cpp:468-481 */
static void compute_3d_point_coords(const Scene& s, const std::vector<V2>& coords, const std::vector<int>& ids, V3& X, bool& valid) {
  valid = false;
  if (coords.size() >= 2) em_estimate3Dpositions(s, coords, ids, X, valid);
}

/* triangulation.cpp:1105-1158 */
static void compute_3d_point_coords_combinations(const Scene& s, const std::vector<V2>& all_coords, const std::vector<int>& all_ids,
                                                 int min_combinations, std::vector<V2>& sel_coords, std::vector<int>& sel_ids,
                                                 std::vector<bool>& selected, V3& X, bool& valid) {
  valid = false;
  selected.resize(all_ids.size());
  std::fill(selected.begin() + min_combinations, selected.end(), false);
  std::fill(selected.begin(), selected.begin() + min_combinations, true);
  do {
    sel_coords.clear(); sel_ids.clear();
    for (size_t i = 0; i < all_ids.size(); ++i)
      if (selected[i]) { sel_coords.push_back(all_coords[i]); sel_ids.push_back(all_ids[i]); }
    compute_3d_point_coords(s, sel_coords, sel_ids, X, valid);
  } while (!valid && std::prev_permutation(selected.begin(), selected.end()));
  if (!valid) return;
  V3 nX; bool need_reorder = false;
  for (size_t i = 0; i < all_ids.size(); ++i) {
    if (!selected[i]) {
      em_add_new_observation_to_3Dpositions(s, X, sel_coords, sel_ids, all_coords[i], all_ids[i], nX, valid);
      if (valid) {
        selected[i] = true;
        X = nX;
        sel_coords.push_back(all_coords[i]);
        if (!need_reorder && all_ids[i] > sel_ids[sel_ids.size() - 1]) need_reorder = true;
        sel_ids.push_back(all_ids[i]);
      }
    }
  }
  valid = true;
  if (need_reorder) { /* reorder_pair_of_vector (edge_graph_3d_utilities.hpp:257-273): sort by view id */
    std::vector<std::pair<int, V2>> z;
    for (size_t i = 0; i < sel_ids.size(); i++) z.push_back({sel_ids[i], sel_coords[i]});
    std::sort(z.begin(), z.end(), [](const std::pair<int, V2>& a, const std::pair<int, V2>& b) { return a.first < b.first; });
    for (size_t i = 0; i < z.size(); i++) { sel_ids[i] = z[i].first; sel_coords[i] = z[i].second; }
  }
}

/* ============================================================ K1: find_epipolar_correspondences =========== */
/* polyline_matching.cpp:45-73.  cand == nullptr: sweep every valid polyline of every other view (the superset
 * formulation of BASELINE configs 2-4, SURVEY finding 6).  plgmm.is_matched is always false on this path
 * (SURVEY finding 5). */
std::vector<std::vector<PlgPoint>> find_epipolar_correspondences(const Scene& s, const std::vector<std::vector<ulong_t>>* cand,
                                                                 int starting_plg_id, const PlgPoint& starting_plgp) {
  std::vector<std::vector<PlgPoint>> res;
  V3 epipolar;
  for (int other = 0; other < s.V; other++) {
    std::vector<PlgPoint> filtered;
    if (other == starting_plg_id) {
      filtered.push_back(starting_plgp);
    } else if (computeCorrespondEpilineSinglePoint(starting_plgp.plp.c, s.Fm(starting_plg_id, other), s.Fok(starting_plg_id, other), epipolar)) {
      auto do_pl = [&](ulong_t pl_id) {
        const Polyline& pl = s.plgs[other][pl_id];
        if (!pl.valid()) return;
        for (const auto& plp : pl.intersect_line(epipolar)) filtered.push_back(PlgPoint{pl_id, plp});
      };
      if (cand) for (ulong_t id : (*cand)[other]) do_pl(id);
      else for (ulong_t id = 0; id < s.plgs[other].size(); id++) do_pl(id);
    }
    res.push_back(filtered);
  }
  return res;
}

/* ============================================================ PLG following (plg_matching.cpp) =========== */
struct Plgp3 { PlgPoint a, b, c; };
typedef std::array<ulong_t, 3> Dir3;

/* 3-view compatible, plg_matching.cpp:51-132 */
static bool compatible3(const Scene& s, const std::vector<int>& ids, const Plgp3& cur, const Dir3& dir, Plgp3& next, V3& X) {
  const eg3d_params& p = s.prm;
  const Polyline& pl_a = s.plgs[ids[0]][cur.a.pl];
  const Polyline& pl_b = s.plgs[ids[1]][cur.b.pl];
  const Polyline& pl_c = s.plgs[ids[2]][cur.c.pl];
  bool reached;
  const PlPoint na = pl_a.next_pl_point_by_distance(cur.a.plp, dir[0], p.follow_first_image_distance, reached);
  if (reached) return false;
  V3 epi; bool found;
  if (!computeCorrespondEpilineSinglePoint(na.c, s.Fm(ids[0], ids[1]), s.Fok(ids[0], ids[1]), epi)) return false;
  PlPoint nb{};
  pl_b.next_pl_point_by_line_intersection(cur.b.plp, dir[1], epi, p.quasiparallel_cos, p.quasiparallel_dist, nb, found);
  if (!found) return false;
  if (!computeCorrespondEpilineSinglePoint(na.c, s.Fm(ids[0], ids[2]), s.Fok(ids[0], ids[2]), epi)) return false;
  PlPoint nc{};
  pl_c.next_pl_point_by_line_intersection(cur.c.plp, dir[2], epi, p.quasiparallel_cos, p.quasiparallel_dist, nc, found);
  if (!found) return false;
  bool valid;
  std::vector<V2> coords = {na.c, nb.c, nc.c};
  compute_3d_point_coords(s, coords, ids, X, valid);
  if (!valid) return false;
  next.a = PlgPoint{cur.a.pl, na}; next.b = PlgPoint{cur.b.pl, nb}; next.c = PlgPoint{cur.c.pl, nc};
  return true;
}

/* plg_matching.cpp:142-203 */
static bool find_direction_given_first_extreme(const Scene& s, const std::vector<int>& ids, const Plgp3& cur, ulong_t first_direction,
                                               Dir3& valid_direction, std::vector<Match>& valid_points) {
  const Polyline& pl_b = s.plgs[ids[1]][cur.b.pl];
  const Polyline& pl_c = s.plgs[ids[2]][cur.c.pl];
  Dir3 pd[4] = {{first_direction, pl_b.start, pl_c.start}, {first_direction, pl_b.start, pl_c.end},
                {first_direction, pl_b.end, pl_c.start}, {first_direction, pl_b.end, pl_c.end}};
  bool vd[4] = {true, true, true, true};
  Plgp3 last[4] = {cur, cur, cur, cur};
  std::vector<Match> tri[4];
  int amount_of_valid = 4;
  while (amount_of_valid > 1) {
    for (int i = 0; i < 4; i++) {
      if (!vd[i]) continue;
      Plgp3 next; V3 X;
      if (compatible3(s, ids, last[i], pd[i], next, X)) {
        last[i] = next;
        tri[i].push_back(Match{X, {next.a, next.b, next.c}, ids});
      } else { vd[i] = false; amount_of_valid--; }
    }
  }
  if (amount_of_valid == 0) return false;
  for (int i = 0; i < 4; i++) if (vd[i]) { valid_direction = pd[i]; valid_points = tri[i]; }
  return true;
}

/* plg_matching.cpp:205-265 + :305-370 (vector-of-directions wrapper, which only clears/overwrites the output
 * vectors of a direction that is valid for THIS hypothesis — SURVEY A.2.15) */
static void find_directions_3view_firstlast(const Scene& s, const Match& m, std::vector<ulong_t>& dir1, bool& d1ok, std::vector<Match>& pts1,
                                            std::vector<ulong_t>& dir2, bool& d2ok, std::vector<Match>& pts2) {
  const int n = (int)m.views.size();
  const int sel[3] = {0, n / 2, n - 1};
  std::vector<int> ids = {m.views[sel[0]], m.views[sel[1]], m.views[sel[2]]};
  Plgp3 cur{m.obs[sel[0]], m.obs[sel[1]], m.obs[sel[2]]};
  const Polyline& pl_a = s.plgs[ids[0]][cur.a.pl];
  const Polyline& pl_b = s.plgs[ids[1]][cur.b.pl];
  const Polyline& pl_c = s.plgs[ids[2]][cur.c.pl];
  d1ok = false; d2ok = false;
  Dir3 direction1{}, direction2{};
  std::vector<Match> p1t, p2t;
  auto opposite = [&](const Dir3& d) {
    return Dir3{pl_a.start == d[0] ? pl_a.end : pl_a.start, pl_b.start == d[1] ? pl_b.end : pl_b.start,
                pl_c.start == d[2] ? pl_c.end : pl_c.start};
  };
  std::vector<Match> towards;
  if (find_direction_given_first_extreme(s, ids, cur, pl_a.start, direction1, towards)) {
    d1ok = true; p1t = towards;
    direction2 = opposite(direction1);
    Plgp3 next; V3 X;
    if (compatible3(s, ids, cur, direction2, next, X)) {
      d2ok = true;
      p2t.clear(); p2t.push_back(Match{X, {next.a, next.b, next.c}, ids});
    }
  } else if (find_direction_given_first_extreme(s, ids, cur, pl_a.end, direction1, towards)) {
    d1ok = true; p1t = towards;
    direction2 = opposite(direction1);
  }
  if (d1ok) {
    dir1.assign(s.V, 0);
    dir1[ids[0]] = direction1[0]; dir1[ids[1]] = direction1[1]; dir1[ids[2]] = direction1[2];
    pts1 = p1t;
    dir2.assign(s.V, 0);
    dir2[ids[0]] = direction2[0]; dir2[ids[1]] = direction2[1]; dir2[ids[2]] = direction2[2];
    if (d2ok) pts2 = p2t;
  }
}

/* all-view compatible, plg_matching.cpp:633-759 */
static bool compatible_all(const Scene& s, const std::vector<ulong_t>& directions, const Match& cur, Match& out) {
  const eg3d_params& p = s.prm;
  for (int si = 0; si < (int)cur.views.size(); si++) {
    const int sv = cur.views[si];
    const PlgPoint& sp = cur.obs[si];
    const Polyline& pl_s = s.plgs[sv][sp.pl];
    bool reached;
    const PlPoint ns = pl_s.next_pl_point_by_distance(sp.plp, directions[sv], p.follow_first_image_distance, reached);
    if (reached) continue;
    std::vector<PlgPoint> sel_plgps; std::vector<int> sel_ids; std::vector<V2> sel_coords;
    sel_plgps.push_back(PlgPoint{sp.pl, ns}); sel_ids.push_back(sv); sel_coords.push_back(ns.c);
    V3 epi; bool found;
    for (int i = 0; i < (int)cur.views.size(); i++) {
      if (i == si) continue;
      const int v = cur.views[i];
      const PlgPoint& cp = cur.obs[i];
      const Polyline& pl = s.plgs[v][cp.pl];
      if (!computeCorrespondEpilineSinglePoint(ns.c, s.Fm(sv, v), s.Fok(sv, v), epi)) continue;
      PlPoint np{};
      pl.next_pl_point_by_line_intersection_bounded_distance(cp.plp, directions[v], epi, p.quasiparallel_cos, p.quasiparallel_dist,
                                                             p.follow_corr_min, p.follow_corr_max, np, found);
      if (found) { sel_plgps.push_back(PlgPoint{cp.pl, np}); sel_ids.push_back(v); sel_coords.push_back(np.c); }
    }
    if (sel_ids.size() < 3) continue; /* PLG_MATCHING_TRIANGULATION_MINIMUM_AMOUNT_OF_POINTS */
    bool valid; V3 X;
    compute_3d_point_coords(s, sel_coords, sel_ids, X, valid);
    if (!valid) {
      std::vector<bool> selected; std::vector<V2> ac; std::vector<int> ai;
      compute_3d_point_coords_combinations(s, sel_coords, sel_ids, 3, ac, ai, selected, X, valid);
      if (valid) {
        /* plg_matching.cpp:720-732: coords come (possibly re-ordered) from the combinations routine, but plgps and
         * ids are re-derived from the selection mask in original order */
        std::vector<PlgPoint> ap; std::vector<int> ids2;
        for (size_t i = 0; i < sel_plgps.size(); i++) if (selected[i]) { ap.push_back(sel_plgps[i]); ids2.push_back(sel_ids[i]); }
        sel_plgps = ap; sel_ids = ids2;
      }
    }
    if (valid) { out.X = X; out.obs = sel_plgps; out.views = sel_ids; return true; }
  }
  return false;
}

/* plg_matching.cpp:765-769 (and follow_direction_vector_end :791-795) */
static void follow_direction(const Scene& s, const std::vector<ulong_t>& directions, std::vector<Match>& pts) {
  Match np;
  while (compatible_all(s, directions, pts[pts.size() - 1], np)) pts.push_back(np);
}
/* plg_matching.cpp:771-789 */
static void follow_direction_vector_start(const Scene& s, const std::vector<ulong_t>& directions, std::vector<Match>& pts) {
  Match np; std::vector<Match> nv;
  if (compatible_all(s, directions, pts[0], np)) {
    nv.push_back(np);
    while (compatible_all(s, directions, nv[nv.size() - 1], np)) nv.push_back(np);
    std::vector<Match> res;
    for (int i = (int)nv.size() - 1; i >= 0; i--) res.push_back(nv[i]);
    for (size_t i = 0; i < pts.size(); i++) res.push_back(pts[i]);
    pts = res;
  }
}

/* plg_matching.cpp:1276-1287 <- follow_plgs_from_match4 :1249-1270 <- find_directions_all_views :1060-1076
 * (its two per-view loops are empty for a 3-view hypothesis) */
static bool compatible_new_plg_point(const Scene& s, const Match& m, std::vector<ulong_t>& dir1, bool& d1ok, std::vector<Match>& pts1,
                                     std::vector<ulong_t>& dir2, bool& d2ok, std::vector<Match>& pts2) {
  d1ok = false; d2ok = false;
  if (m.views.size() >= 3) {
    find_directions_3view_firstlast(s, m, dir1, d1ok, pts1, dir2, d2ok, pts2);
    if (d1ok) follow_direction(s, dir1, pts1);
    if (d2ok) follow_direction(s, dir2, pts2);
  }
  if (d1ok && pts1.size() >= 2) return true;
  if (d2ok && pts2.size() >= 2) return true;
  return false;
}

/* ============================================================ view expansion ============================== */
/* plg_matching.cpp:797-819 */
static void get_plgp_by_epipolar_intersection_from_known_point(const Scene& s, int cur_view, const PlgPoint& cur_plgp, ulong_t direction,
                                                               const Match& known, PlPoint& next, bool& valid) {
  const int sv = known.views[0];
  V3 epi;
  if (!computeCorrespondEpilineSinglePoint(known.obs[0].plp.c, s.Fm(sv, cur_view), s.Fok(sv, cur_view), epi)) { valid = false; return; }
  s.plgs[cur_view][cur_plgp.pl].next_pl_point_by_line_intersection(cur_plgp.plp, direction, epi, s.prm.quasiparallel_cos,
                                                                  s.prm.quasiparallel_dist, next, valid);
}

/* plg_matching.cpp:866-914 */
static bool compatible_direction_noupdate_vector(const Scene& s, int cur_view, const PlgPoint& cur_plgp, ulong_t direction,
                                                 const std::vector<Match>& pts, std::vector<std::pair<V3, PlgPoint>>& to_add,
                                                 int start_interval, int cur_index, int end_interval, bool towards_start) {
  to_add.clear();
  if (pts.empty()) return false;
  PlgPoint actual = cur_plgp;
  int i = towards_start ? cur_index - 1 : cur_index + 1;
  while ((towards_start && i >= start_interval) || (!towards_start && i < end_interval)) {
    const Match& cp = pts[i];
    PlPoint np{}; bool valid;
    get_plgp_by_epipolar_intersection_from_known_point(s, cur_view, actual, direction, cp, np, valid);
    if (!valid) break;
    V3 X;
    em_add_new_observation_to_3Dpositions(s, cp.X, coords_of(cp.obs), cp.views, np.c, cur_view, X, valid);
    if (!valid) break;
    to_add.push_back({X, PlgPoint{cur_plgp.pl, np}});
    actual = PlgPoint{actual.pl, np};
    if (towards_start) i--; else i++;
  }
  return to_add.size() > 0;
}

/* plg_matching.cpp:1011-1058 */
static void find_directions_on_plg_known_3D_point_no_update_vector(const Scene& s, int cur_view, const PlgPoint& cur_plgp,
                                                                   const std::vector<Match>& pts, std::vector<std::pair<V3, PlgPoint>>& n1,
                                                                   std::vector<std::pair<V3, PlgPoint>>& n2, ulong_t& nd1, ulong_t& nd2,
                                                                   int start_interval, int cur_index, int end_interval) {
  const Polyline& pl = s.plgs[cur_view][cur_plgp.pl];
  const ulong_t start = pl.start, end = pl.end;
  if (cur_index > start_interval) {
    if (compatible_direction_noupdate_vector(s, cur_view, cur_plgp, start, pts, n1, start_interval, cur_index, end_interval, true)) {
      nd1 = start; nd2 = end;
      if (cur_index < end_interval)
        compatible_direction_noupdate_vector(s, cur_view, cur_plgp, end, pts, n2, start_interval, cur_index, end_interval, false);
    } else if (compatible_direction_noupdate_vector(s, cur_view, cur_plgp, end, pts, n1, start_interval, cur_index, end_interval, true)) {
      nd1 = end; nd2 = start;
      if (cur_index < end_interval)
        compatible_direction_noupdate_vector(s, cur_view, cur_plgp, start, pts, n2, start_interval, cur_index, end_interval, false);
    } else {
      if (cur_index < end_interval) {
        if (compatible_direction_noupdate_vector(s, cur_view, cur_plgp, end, pts, n2, start_interval, cur_index, end_interval, false)) {
          nd2 = end; nd1 = start;
        } else if (compatible_direction_noupdate_vector(s, cur_view, cur_plgp, start, pts, n2, start_interval, cur_index, end_interval, false)) {
          nd2 = start; nd1 = end;
        }
      }
    }
  }
}

static void update_match(Match& m, int view, const PlgPoint& p, const V3& X) { /* polyline_graph_2d.cpp:1604-1608 */
  m.X = X; m.obs.push_back(p); m.views.push_back(view);
}

/* plg_matching.cpp:1345-1412 */
static std::pair<int, int> add_view_to_3dpoint_and_sides_plgp_matches_vector(const Scene& s, std::vector<Match>& pts, std::vector<ulong_t>& start_dirs,
                                                                            std::vector<ulong_t>& end_dirs, int cur_view, const PlgPoint& cur_plgp,
                                                                            int start_interval, int cur_index, int end_interval, bool& success) {
  success = false;
  V3 nX; bool ok;
  em_add_new_observation_to_3Dpositions(s, pts[cur_index].X, coords_of(pts[cur_index].obs), pts[cur_index].views, cur_plgp.plp.c, cur_view, nX, ok);
  if (!ok) return {0, 0};
  ulong_t nd1 = 0, nd2 = 0;
  std::vector<std::pair<V3, PlgPoint>> n1, n2;
  find_directions_on_plg_known_3D_point_no_update_vector(s, cur_view, cur_plgp, pts, n1, n2, nd1, nd2, start_interval, cur_index, end_interval);
  if (cur_index > 0 && n1.size() == 0) return {0, 0};
  if (cur_index < (int)pts.size() - 1 && n2.size() == 0) return {0, 0};
  success = true;
  int towards_start = (int)n1.size(), towards_end = (int)n2.size();
  update_match(pts[cur_index], cur_view, cur_plgp, nX);
  for (size_t i = 0; i < n1.size(); i++) update_match(pts[cur_index - 1 - i], cur_view, n1[i].second, n1[i].first);
  for (size_t i = 0; i < n2.size(); i++) update_match(pts[cur_index + 1 + i], cur_view, n2[i].second, n2[i].first);
  if (n1.size() > 0 && (int)n1.size() == cur_index) {
    int sz0 = (int)pts.size();
    start_dirs[cur_view] = nd1;
    follow_direction_vector_start(s, start_dirs, pts);
    int added = (int)pts.size() - sz0;
    towards_start += added;
    cur_index += added;
  }
  if (n2.size() > 0 && (int)n2.size() == ((int)pts.size() - cur_index - 1)) {
    int sz0 = (int)pts.size();
    end_dirs[cur_view] = nd2;
    follow_direction(s, end_dirs, pts);
    towards_end += (int)pts.size() - sz0;
  }
  return {towards_start, towards_end};
}

/* triangulation.cpp:742-830 (active branches: SWITCH_PLG_MATCHING_ADDPOINT_BOTHDIR_ONE, SWITCH_DISABLE_INTERVAL) */
static void expand_allpoints_to_other_view_using_plmap(const Scene& s, int other, const std::vector<PlgPoint>& epcs, const Grid& plmap,
                                                       std::vector<Match>& pts, std::vector<ulong_t>& start_dirs, std::vector<ulong_t>& end_dirs,
                                                       int& central_index) {
  bool success;
  std::pair<int, int> m, mi{0, 0};
  bool matched = false;
  for (const auto& epc : epcs) {
    m = add_view_to_3dpoint_and_sides_plgp_matches_vector(s, pts, start_dirs, end_dirs, other, epc, 0, central_index, (int)pts.size(), matched);
    if (g_trace) fprintf(stderr, "[epc] view=%d hit pl=%lu seg=%lu matched=%d ns=%d ne=%d central=%d len=%d\n", other, (unsigned long)epc.pl, (unsigned long)epc.plp.seg, (int)matched, m.first, m.second, central_index, (int)pts.size());
    if (matched) {
      if (m.first > central_index) {
        central_index = m.first;
        mi.first = 0; mi.second = m.first + m.second;
      } else {
        mi.first = central_index - m.first; mi.second = central_index + m.second;
      }
      break;
    }
  }
  int last_matched = -1;
  for (int cur = 0; cur < (int)pts.size(); cur++) {
    if (matched && cur == mi.first) { cur = mi.second; last_matched = mi.second; continue; }
    V2 q = compute_projection(s.P[other].data(), pts[cur].X);
    ulong_t pl_id; bool valid;
    plmap.find_unique_polyline_potentially_within_search_dist(q, pl_id, valid);
    if (g_trace) fprintf(stderr, "[main] view=%d cur=%d/%d X=(%.9g %.9g %.9g) q=(%.6f %.6f) unique=%d pl=%lu\n", other, cur, (int)pts.size(), pts[cur].X.x, pts[cur].X.y, pts[cur].X.z, q.x, q.y, (int)valid, (unsigned long)pl_id);
    if (valid) {
      int central = cur;
      const Polyline& pl = s.plgs[other][pl_id];
      PlPoint ip{};
      const float dsq_ = pl.compute_distancesq(q, ip.seg, ip.c);
      if (g_trace) fprintf(stderr, "[main]   distsq=%.9g %s\n", dsq_, dsq_ > s.prm.max_proj_distsq_expand ? "ABANDON VIEW" : "");
      if (dsq_ > s.prm.max_proj_distsq_expand) return; /* abandons the view */
      PlgPoint init{pl_id, ip};
      int interval_end = matched ? (central <= mi.first ? mi.first : (int)pts.size()) : (int)pts.size();
      auto added = add_view_to_3dpoint_and_sides_plgp_matches_vector(s, pts, start_dirs, end_dirs, other, init, last_matched + 1, central, interval_end, success);
      if (success) {
        if (added.first > central) { central_index = added.first; cur = added.first + added.second; }
        else cur = central + added.second;
        last_matched = cur;
      }
    }
  }
}

/* triangulation.cpp:1027-1088 (+ :550-601 inlined as the triple loop, :960-973 as the view loops) */
std::vector<Match> compute_3D_point_multiple_views_plg_following_expandallviews_vector(
    const Scene& s, int starting_plg_id, const std::vector<std::vector<PlgPoint>>& epc) {
  std::vector<Match> res;
  int non_empty = 0, min_index = -1, max_index = -1;
  for (int i = 0; i < (int)epc.size(); i++)
    if (epc[i].size() > 0) { non_empty++; min_index = min_index != -1 ? min_index : i; max_index = i; }
  if (non_empty < 3) return res;
  int rel = 0, rel_mid = non_empty / 2, mid_index = 0;
  for (int i = 0; i < (int)epc.size(); i++)
    if (epc[i].size() > 0) { if (rel == rel_mid) { mid_index = i; break; } else rel++; }
  const int sel[3] = {min_index, (starting_plg_id == min_index || starting_plg_id == max_index) ? mid_index : starting_plg_id, max_index};

  /* compute_unique_potential_3d_points_3views_plg_following_newpoint_compatibility, triangulation.cpp:550-601 */
  bool found = false;
  std::vector<Match> pts1, pts2; std::vector<ulong_t> dir1, dir2; /* persist across triples */
  std::vector<Match> f_pts1, f_pts2; std::vector<ulong_t> f_dir1, f_dir2; Match f_central;
  const std::vector<int> views = {sel[0], sel[1], sel[2]};
  for (const auto& p0 : epc[sel[0]])
    for (const auto& p1 : epc[sel[1]])
      for (const auto& p2 : epc[sel[2]]) {
        V3 X; bool tvalid;
        std::vector<V2> coords = {p0.plp.c, p1.plp.c, p2.plp.c};
        em_estimate3Dpositions(s, coords, views, X, tvalid);
        if (!tvalid) continue;
        Match cand{X, {p0, p1, p2}, views};
        bool d1ok, d2ok;
        if (compatible_new_plg_point(s, cand, dir1, d1ok, pts1, dir2, d2ok, pts2)) {
          if (found) return res; /* second compatible triple: ambiguous, reject the seed (:587-590) */
          found = true;
          f_pts1 = pts1; f_dir1 = dir1; f_central = cand; f_pts2 = pts2; f_dir2 = dir2;
        }
      }
  if (!found) return res;

  /* new_3dpoint_and_sides_plgp_matches_to_vector, polyline_graph_2d.cpp:1298-1306 */
  for (int i = (int)f_pts1.size() - 1; i >= 0; i--) res.push_back(f_pts1[i]);
  res.push_back(f_central);
  for (size_t i = 0; i < f_pts2.size(); i++) res.push_back(f_pts2[i]);
  int central_point = (int)f_pts1.size();

  /* expand_point_to_other_views_expandallviews_vector, triangulation.cpp:960-973 */
  if ((int)f_dir1.size() != s.V) f_dir1.assign(s.V, 0);
  if ((int)f_dir2.size() != s.V) f_dir2.assign(s.V, 0);
  for (int i = 0; i < s.V; i++) {
    if (i == sel[0] || i == sel[1] || i == sel[2]) continue;
    expand_allpoints_to_other_view_using_plmap(s, i, epc[i], s.plmaps[i], res, f_dir1, f_dir2, central_point);
  }
  return res;
}

}  // namespace eg3d_oracle
