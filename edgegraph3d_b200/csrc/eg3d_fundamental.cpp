// eg3d_fundamental.cpp — row f4 of SURVEY §8 without OpenCV: the track-based fundamental matrices the reference builds once per
// run (generate_all_fundamental_matrices_from_Points, geometric_utilities.cpp:754-820): for every ordered view pair (i, j) with at
// least MIN_CORRESPONDENCES_AMOUNT = 10 SfM points seen by both, F[i][j] = findFundamentalMat(points_i, points_j, FM_LMEDS);
// other pairs keep the reference's 1x1 dummy Mat (fundamental_valid = 0).
//
// Host code, no device.  cv::findFundamentalMat(FM_LMEDS) is a randomised estimator driven by OpenCV's own RNG, 7-point solver and
// SVD; this is NOT a restatement of it and is not bit-identical to it (the path takes F as an INPUT, exactly as the reference's
// entry points take `Mat** all_fundamental_matrices`, so a caller who wants OpenCV's matrices passes OpenCV's matrices).  It is the
// same estimator family with the same constants where they are observable: least median of squares over minimal samples
// (normalised 8-point, Hartley), error of a correspondence = the larger of its two squared point-to-epipolar-line distances,
// confidence 0.99 at an assumed outlier ratio of 0.45, inliers within 2.5 * 1.4826 * (1 + 5 / (n - 8)) * sqrt(median), final
// 8-point fit on the inliers, rank 2 enforced, scaled to F[2][2] = 1.  Deterministic: the sampler is a xorshift generator seeded
// by the view pair.  tests/test_fundamental.py holds it against the analytic matrices on exact data and against the cv2 matrices
// of the packaged example on real tracks (median epipolar distance of the common tracks).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/eg3d.h"

namespace {

struct P2 { double x, y; };

// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 9): A = V diag(w) V^T, columns of V
template <int N>
void jacobi_eigen(double A[N][N], double V[N][N], double w[N]) {
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < N; p++) for (int q = p + 1; q < N; q++) off += A[p][q] * A[p][q];
    if (off < 1e-300) break;
    for (int p = 0; p < N; p++)
      for (int q = p + 1; q < N; q++) {
        if (std::fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; k++) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < N; k++) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < N; k++) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  for (int i = 0; i < N; i++) w[i] = A[i][i];
}

// normalised 8-point algorithm on the correspondences idx[0..m): x2^T F x1 = 0.  Returns false on a degenerate configuration.
bool eight_point(const std::vector<P2>& a, const std::vector<P2>& b, const int* idx, int m, double F[9]) {
  double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
  for (int k = 0; k < m; k++) { c1x += a[idx[k]].x; c1y += a[idx[k]].y; c2x += b[idx[k]].x; c2y += b[idx[k]].y; }
  c1x /= m; c1y /= m; c2x /= m; c2y /= m;
  double d1 = 0, d2 = 0;
  for (int k = 0; k < m; k++) {
    d1 += std::hypot(a[idx[k]].x - c1x, a[idx[k]].y - c1y);
    d2 += std::hypot(b[idx[k]].x - c2x, b[idx[k]].y - c2y);
  }
  if (d1 < 1e-12 || d2 < 1e-12) return false;
  const double s1 = std::sqrt(2.0) * m / d1, s2 = std::sqrt(2.0) * m / d2;
  double M[9][9]; std::memset(M, 0, sizeof M);
  for (int k = 0; k < m; k++) {
    const double x1 = (a[idx[k]].x - c1x) * s1, y1 = (a[idx[k]].y - c1y) * s1, x2 = (b[idx[k]].x - c2x) * s2, y2 = (b[idx[k]].y - c2y) * s2;
    const double r[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
    for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) M[i][j] += r[i] * r[j];
  }
  double V[9][9], w[9];
  jacobi_eigen<9>(M, V, w);
  int best = 0, second = -1;
  for (int i = 1; i < 9; i++) if (w[i] < w[best]) best = i;
  for (int i = 0; i < 9; i++) if (i != best && (second < 0 || w[i] < w[second])) second = i;
  double wmax = 0; for (int i = 0; i < 9; i++) wmax = std::max(wmax, w[i]);
  if (m == 8 && w[second] < 1e-10 * wmax) return false;      // (numerically) two-dimensional null space: degenerate minimal sample
  double F0[3][3];
  for (int i = 0; i < 9; i++) F0[i / 3][i % 3] = V[i][best];
  // rank 2: drop the smallest singular value (SVD through the eigen-decomposition of F0^T F0)
  double G[3][3], Vg[3][3], wg[3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { G[i][j] = 0; for (int k = 0; k < 3; k++) G[i][j] += F0[k][i] * F0[k][j]; }
  jacobi_eigen<3>(G, Vg, wg);
  int lo = 0;
  for (int i = 1; i < 3; i++) if (wg[i] < wg[lo]) lo = i;
  double F2[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double fv = 0; for (int k = 0; k < 3; k++) fv += F0[i][k] * Vg[k][lo];           // (F0 v_lo)_i
    F2[i][j] = F0[i][j] - fv * Vg[j][lo];                                            // F0 (I - v v^T)
  }
  // denormalise: F = T2^T F2 T1, T = [s 0 -s c; 0 s -s c; 0 0 1]
  const double T1[3][3] = {{s1, 0, -s1 * c1x}, {0, s1, -s1 * c1y}, {0, 0, 1}}, T2[3][3] = {{s2, 0, -s2 * c2x}, {0, s2, -s2 * c2y}, {0, 0, 1}};
  double A1[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { A1[i][j] = 0; for (int k = 0; k < 3; k++) A1[i][j] += F2[i][k] * T1[k][j]; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double v = 0; for (int k = 0; k < 3; k++) v += T2[k][i] * A1[k][j]; F[3 * i + j] = v; }
  double nrm = 0; for (int i = 0; i < 9; i++) nrm += F[i] * F[i];
  if (!(nrm > 0) || !std::isfinite(nrm)) return false;
  if (std::fabs(F[8]) > 1.1920929e-07 * std::sqrt(nrm)) { const double s = 1.0 / F[8]; for (int i = 0; i < 9; i++) F[i] *= s; }
  else { const double s = 1.0 / std::sqrt(nrm); for (int i = 0; i < 9; i++) F[i] *= s; }
  return true;
}

// larger of the two squared point-to-epipolar-line distances of a correspondence (the error OpenCV's estimator ranks by)
inline double fm_error(const double F[9], const P2& p, const P2& q) {
  const double a2 = F[0] * p.x + F[1] * p.y + F[2], b2 = F[3] * p.x + F[4] * p.y + F[5], c2 = F[6] * p.x + F[7] * p.y + F[8];
  const double s2 = q.x * a2 + q.y * b2 + c2;
  const double e2 = s2 * s2 / std::max(a2 * a2 + b2 * b2, 1e-300);
  const double a1 = F[0] * q.x + F[3] * q.y + F[6], b1 = F[1] * q.x + F[4] * q.y + F[7];
  const double e1 = s2 * s2 / std::max(a1 * a1 + b1 * b1, 1e-300);
  return std::max(e1, e2);
}

struct Rng { uint64_t s; uint32_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 32); } };

bool lmeds_pair(const std::vector<P2>& a, const std::vector<P2>& b, uint64_t seed, double F[9]) {
  const int n = (int)a.size();
  if (n < 8) return false;
  std::vector<int> all(n); for (int i = 0; i < n; i++) all[i] = i;
  if (n == 8) return eight_point(a, b, all.data(), n, F);
  // iterations for confidence 0.99 at outlier ratio 0.45 with 8-point samples
  const int iters = std::min(2000, (int)std::ceil(std::log(1.0 - 0.99) / std::log(1.0 - std::pow(1.0 - 0.45, 8.0))));
  Rng rng{seed * 0x9E3779B97F4A7C15ull + 0x2545F4914F6CDD1Dull};
  double best_med = 1e300, bestF[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<double> err(n);
  bool have = false;
  for (int it = 0; it < iters; it++) {
    int idx[8];
    for (int k = 0; k < 8; k++) {
      for (;;) {
        idx[k] = (int)(rng.next() % (uint32_t)n);
        bool dup = false; for (int q = 0; q < k; q++) dup |= idx[q] == idx[k];
        if (!dup) break;
      }
    }
    double Fc[9];
    if (!eight_point(a, b, idx, 8, Fc)) continue;
    for (int k = 0; k < n; k++) err[k] = fm_error(Fc, a[k], b[k]);
    std::nth_element(err.begin(), err.begin() + n / 2, err.end());
    const double med = err[n / 2];
    if (med < best_med) { best_med = med; std::memcpy(bestF, Fc, sizeof bestF); have = true; }
  }
  if (!have) return eight_point(a, b, all.data(), n, F);
  // inliers of the best model, then the least-squares fit on them
  const double sigma = 2.5 * 1.4826 * (1.0 + 5.0 / (n - 8)) * std::sqrt(best_med);
  const double thr = std::max(sigma * sigma, 1e-12);
  std::vector<int> inl;
  for (int k = 0; k < n; k++) if (fm_error(bestF, a[k], b[k]) <= thr) inl.push_back(k);
  if ((int)inl.size() >= 8 && eight_point(a, b, inl.data(), (int)inl.size(), F)) return true;
  std::memcpy(F, bestF, sizeof bestF);
  return true;
}

}  // namespace

extern "C" eg3d_status eg3d_fundamental_from_tracks(const eg3d_scene_desc* desc, int32_t min_common, double* out_F, uint8_t* out_valid) {
  if (!desc || !out_F || !out_valid || desc->n_views <= 0) return EG3D_ERR_INVALID_ARG;
  if (desc->n_tracks > 0 && (!desc->track_off || !desc->track_view || !desc->track_xy)) return EG3D_ERR_INVALID_ARG;
  const int V = desc->n_views;
  if (min_common < 8) min_common = 8;
  std::memset(out_F, 0, sizeof(double) * 9 * (size_t)V * V);
  std::memset(out_valid, 0, (size_t)V * V);
  // per view: (track id, observation index), ascending track id; the LAST observation of a view in a track wins, as
  // get_2d_coordinates_of_point_on_image does (edge_graph_3d_utilities.cpp)
  std::vector<std::vector<std::pair<int64_t, int64_t>>> seen((size_t)V);
  for (int64_t t = 0; t < desc->n_tracks; t++)
    for (int64_t o = desc->track_off[t]; o < desc->track_off[t + 1]; o++) {
      const int v = desc->track_view[o];
      if (v < 0 || v >= V) return EG3D_ERR_INVALID_ARG;
      auto& s = seen[(size_t)v];
      if (!s.empty() && s.back().first == t) s.back().second = o; else s.emplace_back(t, o);
    }
  // view pairs are independent (each has its own sampler seed): a plain thread pool over the ordered pairs
  std::atomic<int64_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const int64_t pair = next.fetch_add(1);
      if (pair >= (int64_t)V * V) return;
      const int i = (int)(pair / V), j = (int)(pair % V);
      if (i == j) continue;
      std::vector<P2> a, b;
      const auto &si = seen[(size_t)i], &sj = seen[(size_t)j];
      size_t p = 0, q = 0;
      while (p < si.size() && q < sj.size()) {
        if (si[p].first < sj[q].first) p++;
        else if (si[p].first > sj[q].first) q++;
        else {
          a.push_back({(double)desc->track_xy[2 * si[p].second], (double)desc->track_xy[2 * si[p].second + 1]});
          b.push_back({(double)desc->track_xy[2 * sj[q].second], (double)desc->track_xy[2 * sj[q].second + 1]});
          p++; q++;
        }
      }
      if ((int)a.size() < min_common) continue;
      double F[9];
      if (!lmeds_pair(a, b, (uint64_t)i * (uint64_t)V + (uint64_t)j + 1, F)) continue;
      std::memcpy(out_F + 9 * ((size_t)i * V + j), F, sizeof F);
      out_valid[(size_t)i * V + j] = 1;
    }
  };
  const unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nt; t++) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  return EG3D_OK;
}
