// eg3d_gn.cuh — K2 stand-alone Gauss-Newton kernels (B5/B6 primitives, BASELINE config 5) and the a14 outlier filter.
//   fp64: em_GaussNewton semantics (triangulation.cpp:105-176), warp-cooperative: G lanes per hypothesis, shuffle-reduced sums.
//   fp32: GaussNewton of the outlier filter (filtering/gauss_newton.cpp:83-134) with OpenCV's CV_32F arithmetic
//         reproduced operation by operation (float products and sequential float accumulation for the 4x4*4x1
//         projections; double sequential accumulation for J^T J and (H^-1 J^T) r; double cofactor determinant and
//         inverse) so that the inlier bitmap matches the oracle bit for bit; one hypothesis per thread, Jacobian rows parked in
//         shared memory between the two passes of an iteration, exact early exit once the float iterates cycle.
// Camera matrices are staged in shared memory once per CTA.
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

// The float loop of the filter stops only when the float MSE repeats within 5e-10 (gauss_newton.cpp:114), i.e. exactly; on
// BASELINE configs[4] two thirds of the hypotheses never get there: X ends up on a short cycle of float values (period 2-4,
// entered after ~8 iterations) and the loop runs its 30 iterations.  The iteration is a deterministic map X -> X', so once
// X_t equals some X_(t-p) bit for bit every later iterate is known: if the stop test of iteration t does not fire, no later
// one will (the same pairs recur), the determinant test repeats its earlier verdicts, and the state after 30 iterations is
// read from the history — X_30 and last_mse = mse(X_29).  Exact by construction (same bits as running all 30 iterations).
constexpr int GN_HIST = 12;

// filtering/gauss_newton.cpp:83-134, bit-faithful to the oracle restatement, one hypothesis per LANE at a time (its sums must be
// accumulated observation by observation to stay bit-exact).  The loop is per lane: a lane that finishes its hypothesis takes
// its next one while its neighbours keep iterating, so the warp never idles behind its slowest member (two thirds of BASELINE
// configs[4] would otherwise hold every warp for the full 30 iterations).  An iteration is cut into phases with a warp barrier
// between them — residual / Jacobian pass, stop test + 3x3 solve, update pass — so that lanes which are at different iterations
// of different hypotheses still execute each phase together.
struct Gn32State { float X[3]; float last_mse; int it; float hx[GN_HIST][3]; float hm[GN_HIST]; };

// pass 1 of an iteration: residuals (float) and J^T J (double accumulation of float products), observations in order
EG3D_D void gn_f32_pass1(const float* __restrict__ Ps, int n, const int* __restrict__ views, const float2* __restrict__ pts, int os,
                         float X0, float X1, float X2, float& mse_out, double h[6]) {
  float mse = 0;
  double h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0;
#pragma unroll 2
  for (int m = 0; m < n; m++) {
    const float* P = Ps + 12 * views[m * os];
    float xH = P[0] * X0 + P[1] * X1 + P[2] * X2 + P[3] * 1.0f;
    float yH = P[4] * X0 + P[5] * X1 + P[6] * X2 + P[7] * 1.0f;
    float zH = P[8] * X0 + P[9] * X1 + P[10] * X2 + P[11] * 1.0f;
    float2 pt = pts[m * os];
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
  mse_out = mse;
  h[0] = h00; h[1] = h01; h[2] = h02; h[3] = h11; h[4] = h12; h[5] = h22;
}

// stop test, cycle shortcut, determinant test and inverse.  Returns 0 = update with Hi, 1 = finished and accepted, 2 = finished
// and rejected.
EG3D_D int gn_f32_decide(const eg3d_params& prm, int n, Gn32State& st, float mse, const double h[6], float gn_max_mse, float Hi[9]) {
  const int it = st.it, max_it = prm.gn_max_iters;
  const float X0 = st.X[0], X1 = st.X[1], X2 = st.X[2];
  const float cur = mse / (n * 2);
  const float diff = cur - st.last_mse;
  const bool stop = prm.filter_abs_int ? ((double)abs((int)diff) < prm.filter_gn_stop) : ((double)fabsf(diff) < prm.filter_gn_stop);
  if (stop) return st.last_mse < gn_max_mse ? 1 : 2;
  // cycle shortcut (see GN_HIST): X_it == X_(it-p)?  The stop test of this iteration has just been passed.
  int p = 0;
  const int back = it < GN_HIST ? it : GN_HIST;
  for (int q = 1; q <= back; q++) {
    const int j = (it - q) % GN_HIST;
    if (st.hx[j][0] == X0 && st.hx[j][1] == X1 && st.hx[j][2] == X2) { p = q; break; }
  }
  if (p > 0) {
    // iterates from s0 = it - p on have period p: X_k = X_(s0 + (k - s0) mod p).  After the loop: last_mse = mse(X_(max_it-1)), X = X_max_it.
    const int s0 = it - p;
    const int k29 = s0 + (max_it - 1 - s0) % p, k30 = s0 + (max_it - s0) % p;
    st.last_mse = st.hm[k29 % GN_HIST];
    st.X[0] = st.hx[k30 % GN_HIST][0]; st.X[1] = st.hx[k30 % GN_HIST][1]; st.X[2] = st.hx[k30 % GN_HIST][2];
    return st.last_mse < gn_max_mse ? 1 : 2;
  }
  st.hx[it % GN_HIST][0] = X0; st.hx[it % GN_HIST][1] = X1; st.hx[it % GN_HIST][2] = X2; st.hm[it % GN_HIST] = cur;
  st.last_mse = cur;
  float Hf[9] = {(float)h[0], (float)h[1], (float)h[2], (float)h[1], (float)h[3], (float)h[4], (float)h[2], (float)h[4], (float)h[5]};
  double m9[9];
#pragma unroll
  for (int i = 0; i < 9; i++) m9[i] = Hf[i];
  double d = det3d(m9);
  float df = (float)d;
  if ((double)df < prm.filter_gn_det_min) return 2;
  double t[9];
  if (d != 0) inv3d(m9, d, t);
  else {
#pragma unroll
    for (int i = 0; i < 9; i++) t[i] = 0;
  }
#pragma unroll
  for (int i = 0; i < 9; i++) Hi[i] = (float)t[i];
  return 0;
}

// pass 2: delta = (H^-1 J^T) r with M = H^-1 J^T rounded to float element-wise, then a double dot with r
EG3D_D void gn_f32_pass2(const float* __restrict__ Ps, int n, const int* __restrict__ views, const float2* __restrict__ pts, int os,
                         const float Hi[9], float X[3]) {
  const float X0 = X[0], X1 = X[1], X2 = X[2];
  double a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 2
  for (int m = 0; m < n; m++) {
    const float* P = Ps + 12 * views[m * os];
    float xH = P[0] * X0 + P[1] * X1 + P[2] * X2 + P[3] * 1.0f;
    float yH = P[4] * X0 + P[5] * X1 + P[6] * X2 + P[7] * 1.0f;
    float zH = P[8] * X0 + P[9] * X1 + P[10] * X2 + P[11] * 1.0f;
    float2 pt = pts[m * os];
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
  X[0] = X0 + (float)a0; X[1] = X1 + (float)a1; X[2] = X2 + (float)a2;
}

// fp32 = the outlier filter's Gauss-Newton, persistent grid: thread t solves hypotheses t, t + T, t + 2T, ...  Dynamic shared
// memory: the cameras (V x 12 floats) and, with STAGE, the observations of the lane's current hypothesis ([obs][thread] view
// ids and coordinates, copied once per hypothesis: the 20-60 passes of a solve then never leave the SM).
template <bool STAGE>
__global__ void __launch_bounds__(GN_THREADS) gn32_kernel(const __grid_constant__ DevScene S, const __grid_constant__ GnProblem pr, int stage_obs) {
  extern __shared__ float sP[];
  for (int i = threadIdx.x; i < S.V * 12; i += blockDim.x) sP[i] = S.P[i];
  __syncthreads();
  const int cam = (S.V * 12 + 3) & ~3;
  float2* sxy = reinterpret_cast<float2*>(sP + cam) + threadIdx.x;                                 // [stage_obs][GN_THREADS]
  int* sv = reinterpret_cast<int*>(sP + cam + 2 * (size_t)stage_obs * GN_THREADS) + threadIdx.x;   // [stage_obs][GN_THREADS]
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  Gn32State st;
  int64_t o0 = 0; int n = 0;
  bool fresh = true;
  bool have = h < pr.n;
  const int max_it = S.prm.gn_max_iters;
  while (__any_sync(0xffffffffu, have)) {
    // phase 0: lanes that start a hypothesis
    if (have && fresh) {
      o0 = pr.obs_off ? pr.obs_off[h] : h * pr.k;
      n = pr.obs_off ? (int)(pr.obs_off[h + 1] - o0) : pr.k;
      st.X[0] = pr.init[3 * h]; st.X[1] = pr.init[3 * h + 1]; st.X[2] = pr.init[3 * h + 2];
      st.last_mse = 0; st.it = 0;
      if (STAGE) for (int m = 0; m < n; m++) { sxy[m * GN_THREADS] = pr.obs_xy[o0 + m]; sv[m * GN_THREADS] = pr.obs_view[o0 + m]; }
      fresh = false;
    }
    __syncwarp();
    // phase 1: residuals + normal matrix of every lane that still iterates (the loop has run max_it times: finished, :97,:128)
    int r = -1;                       // -1 idle, 0 update, 1 accepted, 2 rejected
    float mse = 0; double hs[6];
    const int* vv = STAGE ? sv : pr.obs_view + o0;
    const float2* pp = STAGE ? sxy : pr.obs_xy + o0;
    const int os = STAGE ? GN_THREADS : 1;
    if (have) {
      if (st.it >= max_it) r = st.last_mse < pr.gn_max_mse ? 1 : 2;
      else gn_f32_pass1(sP, n, vv, pp, os, st.X[0], st.X[1], st.X[2], mse, hs);
    }
    __syncwarp();
    // phase 2: stop test / cycle shortcut / 3x3 inverse
    float Hi[9];
    if (have && r < 0) r = gn_f32_decide(S.prm, n, st, mse, hs, pr.gn_max_mse, Hi);
    __syncwarp();
    // phase 3: update pass
    if (have && r == 0) { gn_f32_pass2(sP, n, vv, pp, os, Hi, st.X); st.it++; }
    __syncwarp();
    // phase 4: results of the lanes that finished; they start their next hypothesis on the next trip
    if (have && r > 0) {
      const bool ok = r == 1;
      if (pr.out_ok) pr.out_ok[h] = ok ? 1 : 0;
      if (pr.out_mse) pr.out_mse[h] = st.last_mse;
      if (!pr.write_back_only_ok || ok) { pr.out_xyz[3 * h] = st.X[0]; pr.out_xyz[3 * h + 1] = st.X[1]; pr.out_xyz[3 * h + 2] = st.X[2]; }
      h += T; fresh = true; have = h < pr.n;
    }
  }
}

// fp64 = em_GaussNewton (triangulation.cpp:105-176), warp-cooperative: G lanes per hypothesis (G = 1, 2, 4, ..., 32), the lanes
// stride over the observations and the ten normal-equation sums are butterfly-reduced with warp shuffles inside the group, so
// every lane of a group holds the same iterate; closed-form 3x3 solve per iteration.  Arithmetic as in gn_group of the matching
// path (one reciprocal of the depth per observation, fused multiply-adds, cameras widened to double once in shared memory):
// rounding-level differences from the reference's division form, thresholds unaffected on every tested input.
template <int G>
__global__ void __launch_bounds__(GN_THREADS) gn64_kernel(const __grid_constant__ DevScene S, const __grid_constant__ GnProblem pr) {
  extern __shared__ double sP64[];
  for (int i = threadIdx.x; i < S.V * 12; i += blockDim.x) sP64[i] = S.P64[i];
  __syncthreads();
  // persistent grid: the camera table is staged once per resident CTA, which then walks the hypotheses with the grid's stride
  // (one CTA per 32 hypotheses re-staged 19 KB of cameras for 61 KB of camera reads)
  const int64_t total = (((int64_t)pr.n * G + GN_THREADS - 1) / GN_THREADS) * GN_THREADS;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t h = t / G;
    const int sub = (int)(t % G);
    const bool active = h < pr.n;                       // whole groups are active or not (GN_THREADS is a multiple of G)
    const int64_t hh = active ? h : 0;
    const int64_t o0 = pr.obs_off ? pr.obs_off[hh] : hh * pr.k;
    const int n = pr.obs_off ? (int)(pr.obs_off[hh + 1] - o0) : pr.k;
    const int* __restrict__ views = pr.obs_view + o0;
    const float2* __restrict__ pts = pr.obs_xy + o0;
    double X0 = pr.init[3 * hh], X1 = pr.init[3 * hh + 1], X2 = pr.init[3 * hh + 2];
    double last_mse = 0;
    bool running = active, failed = false;
    for (int it = 0; it < S.prm.gn_max_iters; it++) {
      if (!__any_sync(0xffffffffu, running)) break;
      GnAcc a = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      if (running) {
#pragma unroll 2
        for (int i = sub; i < n; i += G) { const float2 p = pts[i]; gn_accumulate_fast(sP64 + 12 * views[i], p.x, p.y, X0, X1, X2, a); }
      }
#pragma unroll
      for (int off = G >> 1; off > 0; off >>= 1) {
        a.mse += __shfl_xor_sync(0xffffffffu, a.mse, off);
        a.h00 += __shfl_xor_sync(0xffffffffu, a.h00, off); a.h01 += __shfl_xor_sync(0xffffffffu, a.h01, off);
        a.h02 += __shfl_xor_sync(0xffffffffu, a.h02, off); a.h11 += __shfl_xor_sync(0xffffffffu, a.h11, off);
        a.h12 += __shfl_xor_sync(0xffffffffu, a.h12, off); a.h22 += __shfl_xor_sync(0xffffffffu, a.h22, off);
        a.g0 += __shfl_xor_sync(0xffffffffu, a.g0, off); a.g1 += __shfl_xor_sync(0xffffffffu, a.g1, off);
        a.g2 += __shfl_xor_sync(0xffffffffu, a.g2, off);
      }
      if (running) {
        const double cur = a.mse / (n * 2);
        if (fabs(cur - last_mse) < S.prm.gn_stop) running = false;
        else {
          last_mse = cur;
          const double H[9] = {a.h00, a.h01, a.h02, a.h01, a.h11, a.h12, a.h02, a.h12, a.h22};
          const double d = det3d(H);
          if (d < S.prm.gn_det_min) { running = false; failed = true; }
          else {
            double Hi[9]; inv3d(H, d, Hi);
            X0 += Hi[0] * a.g0 + Hi[1] * a.g1 + Hi[2] * a.g2;
            X1 += Hi[3] * a.g0 + Hi[4] * a.g1 + Hi[5] * a.g2;
            X2 += Hi[6] * a.g0 + Hi[7] * a.g1 + Hi[8] * a.g2;
          }
        }
      }
    }
    if (active && sub == 0) {
      const bool ok = !failed && last_mse < S.prm.gn_accept_mse;
      if (pr.out_ok) pr.out_ok[h] = ok ? 1 : 0;
      if (pr.out_mse) pr.out_mse[h] = (float)last_mse;
      if (!pr.write_back_only_ok || ok) { pr.out_xyz[3 * h] = (float)X0; pr.out_xyz[3 * h + 1] = (float)X1; pr.out_xyz[3 * h + 2] = (float)X2; }
    }
  }
}

}  // namespace eg3d
