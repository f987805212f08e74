// eg3d_gn.cuh — K2 stand-alone Gauss-Newton kernels (B5/B6 primitives, BASELINE config 5) and the a14 outlier filter.
//   fp64: em_GaussNewton semantics (triangulation.cpp:105-176), one hypothesis per thread, observations in order.
//   fp32: GaussNewton of the outlier filter (filtering/gauss_newton.cpp:83-134) with OpenCV's CV_32F arithmetic
//         reproduced operation by operation (float products and sequential float accumulation for the 4x4*4x1
//         projections; double sequential accumulation for J^T J and (H^-1 J^T) r; double cofactor determinant and
//         inverse) so that the inlier bitmap matches the oracle bit for bit.
// Camera matrices are staged in shared memory once per CTA (48 B per view).
#pragma once
#include "eg3d_dev.cuh"

namespace eg3d {

constexpr int GN_THREADS = 128;

struct GnProblem {       // CSR (obs_off != null) or fixed stride (obs_off == null, k per hypothesis)
  int64_t n;
  const int64_t* obs_off; int k;
  const int* obs_view; const float2* obs_xy;
  const float* init;     // [n][3]
  float* out_xyz; float* out_mse; uint8_t* out_ok;
  float gn_max_mse;      // fp32 variant
  int write_back_only_ok; // filter: leave xyz untouched unless accepted
};

// filtering/gauss_newton.cpp:83-134 for one point, bit-faithful to the oracle restatement.
EG3D_D bool gn_f32_exact(const float* __restrict__ Ps, const eg3d_params& prm, int n, const int* __restrict__ views,
                         const float2* __restrict__ pts, float X[3], float gn_max_mse, float& last_mse_out) {
  float last_mse = 0;
  const int n2 = n * 2;
  for (int it = 0; it < prm.gn_max_iters; it++) {
    float mse = 0;
    double h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0;
    for (int m = 0; m < n; m++) {
      const float* P = Ps + 12 * views[m];
      float xH = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3] * 1.0f;
      float yH = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7] * 1.0f;
      float zH = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] * 1.0f;
      float2 pt = pts[m];
      float rx = pt.x - xH / zH;
      mse += rx * rx;
      float ry = pt.y - yH / zH;
      mse += ry * ry;
      float zz = zH * zH;
      float jx0 = (P[0] * zH - P[8] * xH) / zz, jx1 = (P[1] * zH - P[9] * xH) / zz, jx2 = (P[2] * zH - P[10] * xH) / zz;
      float jy0 = (P[4] * zH - P[8] * yH) / zz, jy1 = (P[5] * zH - P[9] * yH) / zz, jy2 = (P[6] * zH - P[10] * yH) / zz;
      h00 += (double)jx0 * (double)jx0; h00 += (double)jy0 * (double)jy0;
      h01 += (double)jx0 * (double)jx1; h01 += (double)jy0 * (double)jy1;
      h02 += (double)jx0 * (double)jx2; h02 += (double)jy0 * (double)jy2;
      h11 += (double)jx1 * (double)jx1; h11 += (double)jy1 * (double)jy1;
      h12 += (double)jx1 * (double)jx2; h12 += (double)jy1 * (double)jy2;
      h22 += (double)jx2 * (double)jx2; h22 += (double)jy2 * (double)jy2;
    }
    float diff = mse / n2 - last_mse;
    bool stop = prm.filter_abs_int ? ((double)abs((int)diff) < prm.filter_gn_stop) : ((double)fabsf(diff) < prm.filter_gn_stop);
    if (stop) break;
    last_mse = mse / n2;
    float Hf[9] = {(float)h00, (float)h01, (float)h02, (float)h01, (float)h11, (float)h12, (float)h02, (float)h12, (float)h22};
    double m9[9];
#pragma unroll
    for (int i = 0; i < 9; i++) m9[i] = Hf[i];
    double d = det3d(m9);
    float df = (float)d;
    if ((double)df < prm.filter_gn_det_min) { last_mse_out = last_mse; return false; }
    float Hi[9];
    {
      double t[9];
      if (d != 0) inv3d(m9, d, t);
      else {
#pragma unroll
        for (int i = 0; i < 9; i++) t[i] = 0;
      }
#pragma unroll
      for (int i = 0; i < 9; i++) Hi[i] = (float)t[i];
    }
    // second pass: delta = (H^-1 J^T) r with M = H^-1 J^T rounded to float element-wise, then a double dot with r
    double a0 = 0, a1 = 0, a2 = 0;
    for (int m = 0; m < n; m++) {
      const float* P = Ps + 12 * views[m];
      float xH = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3] * 1.0f;
      float yH = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7] * 1.0f;
      float zH = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] * 1.0f;
      float2 pt = pts[m];
      float rx = pt.x - xH / zH, ry = pt.y - yH / zH;
      float zz = zH * zH;
      float jx0 = (P[0] * zH - P[8] * xH) / zz, jx1 = (P[1] * zH - P[9] * xH) / zz, jx2 = (P[2] * zH - P[10] * xH) / zz;
      float jy0 = (P[4] * zH - P[8] * yH) / zz, jy1 = (P[5] * zH - P[9] * yH) / zz, jy2 = (P[6] * zH - P[10] * yH) / zz;
#define EG3D_MROW(a, j0, j1, j2) ((float)((double)Hi[3 * a + 0] * (double)(j0) + (double)Hi[3 * a + 1] * (double)(j1) + (double)Hi[3 * a + 2] * (double)(j2)))
      a0 += (double)EG3D_MROW(0, jx0, jx1, jx2) * (double)rx; a0 += (double)EG3D_MROW(0, jy0, jy1, jy2) * (double)ry;
      a1 += (double)EG3D_MROW(1, jx0, jx1, jx2) * (double)rx; a1 += (double)EG3D_MROW(1, jy0, jy1, jy2) * (double)ry;
      a2 += (double)EG3D_MROW(2, jx0, jx1, jx2) * (double)rx; a2 += (double)EG3D_MROW(2, jy0, jy1, jy2) * (double)ry;
#undef EG3D_MROW
    }
    X[0] += (float)a0; X[1] += (float)a1; X[2] += (float)a2;
  }
  last_mse_out = last_mse;
  return last_mse < gn_max_mse;
}

template <bool FP64>
__global__ void __launch_bounds__(GN_THREADS) gn_kernel(const __grid_constant__ DevScene S, const __grid_constant__ GnProblem pr) {
  extern __shared__ float sP[];
  for (int i = threadIdx.x; i < S.V * 12; i += blockDim.x) sP[i] = S.P[i];
  __syncthreads();
  int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= pr.n) return;
  int64_t o0 = pr.obs_off ? pr.obs_off[h] : h * pr.k;
  int n = pr.obs_off ? (int)(pr.obs_off[h + 1] - o0) : pr.k;
  const int* views = pr.obs_view + o0;
  const float2* pts = pr.obs_xy + o0;
  bool ok; float mse_out;
  float xo[3];
  if (FP64) {
    double X[3] = {pr.init[3 * h], pr.init[3 * h + 1], pr.init[3 * h + 2]};
    double last_mse = 0;
    ok = true;
    for (int it = 0; it < S.prm.gn_max_iters; it++) {
      GnAcc a = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int i = 0; i < n; i++) { float2 p = pts[i]; gn_accumulate(sP + 12 * views[i], p.x, p.y, X, a); }
      int r = gn_update(a, n, S.prm, last_mse, X);
      if (r == 1) break;
      if (r == -1) { ok = false; break; }
    }
    if (ok) ok = last_mse < S.prm.gn_accept_mse;
    mse_out = (float)last_mse;
    xo[0] = (float)X[0]; xo[1] = (float)X[1]; xo[2] = (float)X[2];
  } else {
    float X[3] = {pr.init[3 * h], pr.init[3 * h + 1], pr.init[3 * h + 2]};
    ok = gn_f32_exact(sP, S.prm, n, views, pts, X, pr.gn_max_mse, mse_out);
    xo[0] = X[0]; xo[1] = X[1]; xo[2] = X[2];
  }
  if (pr.out_ok) pr.out_ok[h] = ok ? 1 : 0;
  if (pr.out_mse) pr.out_mse[h] = mse_out;
  if (!pr.write_back_only_ok || ok) { pr.out_xyz[3 * h] = xo[0]; pr.out_xyz[3 * h + 1] = xo[1]; pr.out_xyz[3 * h + 2] = xo[2]; }
}

}  // namespace eg3d
