// eg3d_dev.cuh — device-resident scene view and the geometric / polyline / Gauss-Newton primitives shared by the
// kernels of libeg3d.so.  Host+device where the host needs the same arithmetic (grid build, seed sampler).
//
// Arithmetic contract: the file is compiled with -fmad=false, IEEE division and square root, so every float/double
// expression rounds exactly as written (the reference's x86-64 build has no FMA contraction).  Citations are to the
// EdgeGraph3D reference tree.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/eg3d.h"

#define EG3D_HD __host__ __device__ __forceinline__
#define EG3D_D __device__ __forceinline__
#define EG3D_HD_NI static __host__ __device__ __noinline__
#ifndef EG3D_GN_UNROLL
#define EG3D_GN_UNROLL 1
#endif

namespace eg3d {
constexpr int kGnUnroll = EG3D_GN_UNROLL;

struct DevGrid {            // uniform polyline grid, PolyLine2DMap (polyLine_2d_map.cpp:40-58), CSR per (view, cell)
  float cell; int w, h;
  const int* cell_off;      // [V*w*h + 1]
  const uint32_t* ids;      // ascending polyline ids per cell
};

struct DevScene {
  int V, width, height;
  const float* P;               // [V][12]
  const double* P64;            // [V][12] the same matrices widened once (triangulation.cpp:301-306 converts per call)
  const double* F;              // [V*V][9]
  const double* Fp;             // [V*V][9] fundamental matrices derived from the cameras P themselves (exact two-view geometry of
                                //          the GN problems; the input F may come from LMedS and is NOT used for pruning)
  const double* Fph;            // [V*V] Frobenius norm of Fp's upper-left 2x2 block (the `h` of pair_cannot_fit)
  const uint8_t* Fvalid;        // [V*V]
  const int* view_poly_off;     // [V+1]
  const int* poly_vert_off;     // [NP+1]
  const float2* verts;
  const uint32_t* poly_start;   // [NP]
  const uint32_t* poly_end;     // [NP]
  // K1 staging arrays: every segment of every valid polyline, ascending (polyline, segment), in the reference's
  // intersect_line orientation (P[i] first, polyline_graph_2d.cpp:318-323) stored as (x1, y1, dx, dy).
  const int* view_seg_off;      // [V+1]
  const float4* seg;            // [NSEG]
  const uint2* seg_id;          // [NSEG] (polyline id in view, segment_index = i-1)
  const int* poly_seg_off;      // [NP+1] first staged segment of each polyline
  // K1 cull structures: groups of <= 32 consecutive staged segments of ONE polyline, with an inflated bounding box, and a
  // per-view table of chunks (<= 2048 segments, <= 128 groups) that the sweep kernel streams through shared memory
  const int* view_chunk_off;    // [V+1]
  const int4* chunks;           // (first segment, #segments, first group, #groups padded to a multiple of 4)
  const float4* grp_box;        // (cx, cy, ex, ey): centre and half extents inflated by 0.05 px; ex < 0 for padding groups
  const uint32_t* grp_desc;     // (offset of the group's first segment in its chunk << 6) | #segments
  DevGrid g_expand;             // 4 px
  DevGrid g_corr;               // 30 px (only with tracks)
  // tracks
  int64_t n_tracks;
  const float* track_xyz; const int64_t* track_off; const int32_t* track_view; const float2* track_xy;
  eg3d_params prm;
};

struct PlP { uint32_t seg; float2 c; };                    // pl_point
struct Plg { uint32_t pl; uint32_t seg; float2 c; };       // plg_point
struct Pl { const float2* pc; int n; uint32_t start, end; };  // polyline view

EG3D_HD Pl get_pl(const DevScene& S, int view, uint32_t pl_id) {
  int g = S.view_poly_off[view] + (int)pl_id;
  int o = S.poly_vert_off[g];
  Pl p; p.pc = S.verts + o; p.n = S.poly_vert_off[g + 1] - o; p.start = S.poly_start[g]; p.end = S.poly_end[g];
  return p;
}

// geometric_utilities.cpp:555-557: float subtraction, double square + sum, float result.  The squares are exact in
// double, so contraction could not change the result.
EG3D_HD float sqdist2(float2 a, float2 b) {
  float dx = a.x - b.x, dy = a.y - b.y;
  return (float)((double)dx * (double)dx + (double)dy * (double)dy);
}
EG3D_HD_NI float dist2(float2 a, float2 b) { return sqrtf(sqdist2(a, b)); }  // :571-573

// geometric_utilities.cpp:824-843 -> cv::computeCorrespondEpilines(whichImage = 1)
EG3D_HD_NI bool epiline(const DevScene& S, int a, int b, float2 p, float3& l) {
  size_t idx = (size_t)a * S.V + b;
  if (!S.Fvalid[idx]) return false;
  const double* F = S.F + idx * 9;
  double x = p.x, y = p.y;
  double la = F[0] * x + F[1] * y + F[2];
  double lb = F[3] * x + F[4] * y + F[5];
  double lc = F[6] * x + F[7] * y + F[8];
  double nu = la * la + lb * lb;
  nu = nu ? 1. / sqrt(nu) : 1.;
  l.x = (float)(la * nu); l.y = (float)(lb * nu); l.z = (float)(lc * nu);
  return true;
}

// geometric_utilities.cpp:272-312 on a segment given as (x1, y1, dx, dy)
EG3D_HD bool isect_seg_line(float x1, float y1, float dx, float dy, float3 l, float2& inter) {
  float num = l.x * x1 + l.y * y1 + l.z;
  float den = l.x * dx + l.y * dy;
  if (den != 0) {
    float t = -num / den;
    if (t >= 0 && t <= 1) { inter.x = x1 + t * dx; inter.y = y1 + t * dy; return true; }
  }
  return false;
}

EG3D_HD float dist_point_line(float px, float py, float3 l) {  // geometric_utilities.cpp:997-1009
  float den = l.x * px + l.y * py + l.z;
  den *= den;
  return sqrtf(den / (l.x * l.x + l.y * l.y));
}

// geometric_utilities.cpp:365-430.  Returns bit0 = intersection_found, bit1 = quasiparallel_within_distance.
EG3D_HD_NI int isect_seg_line_nqp(float x1, float y1, float x2, float y2, float3 l, float max_cos, float max_dist, float2& inter) {
  float dx = x2 - x1, dy = y2 - y1;
  int res = 0;
  float num = l.x * x1 + l.y * y1 + l.z;
  float den = l.x * dx + l.y * dy;
  if (den != 0) {
    float t = -num / den;
    if (t >= 0 && t <= 1) { inter.x = x1 + t * dx; inter.y = y1 + t * dy; res |= 1; }
    // compute_anglecos(segm, line): signed cosine between the segment and (1, -a/b) or (0,1) (:579-581, :608-618)
    float lx, ly;
    if (l.y == 0) { lx = 0.0f; ly = 1.0f; } else { lx = 1.0f; ly = -l.x / l.y; }
    float ab = dx * lx + dy * ly, aa = dx * dx + dy * dy, bb = lx * lx + ly * ly;
    float cosv = ab / sqrtf(aa * bb);
    if (cosv > max_cos) {
      float distance;
      if (t < 0) distance = dist_point_line(x1, y1, l);
      else if (t > 1) distance = dist_point_line(x2, y2, l);
      else distance = 0;
      if (distance <= max_dist) res |= 2;
    }
  } else {
    if (dist_point_line(x1, y1, l) <= max_dist) res |= 2;
  }
  return res;
}

EG3D_HD float2 lerp2(float2 a, float2 b, float r) {  // first_plus_ratio_of_segment, geometric_utilities.cpp:1370-1372
  return make_float2(a.x + r * (b.x - a.x), a.y + r * (b.y - a.y));
}

// polyline::next_pl_point_by_distance, polyline_graph_2d.cpp:391-447.  A direction that is neither extreme is UB in
// the reference; rule shared with the oracle: cannot drive = reached extreme.
EG3D_HD_NI PlP step_by_distance(const Pl& pl, PlP init, uint32_t dir, float distance, bool& reached) {
  float prevdist = 0, curdist, ratio;
  reached = false;
  PlP r;
  const int n = pl.n;
  if (dir == pl.start) {
    curdist = dist2(pl.pc[init.seg], init.c);
    if (curdist >= distance) { ratio = distance / curdist; r.seg = init.seg; r.c = lerp2(init.c, pl.pc[init.seg], ratio); return r; }
    int i;
    for (i = (int)init.seg; i > 0; i--) {
      prevdist = curdist;
      curdist = dist2(pl.pc[i - 1], init.c);
      if (curdist >= distance) break;
    }
    if (i == 0) { reached = true; r.seg = 0; r.c = pl.pc[0]; return r; }
    ratio = (distance - prevdist) / (curdist - prevdist);
    r.seg = (uint32_t)(i - 1); r.c = lerp2(pl.pc[i], pl.pc[i - 1], ratio); return r;
  } else if (dir == pl.end) {
    if ((int)init.seg >= n - 1) { reached = true; r.seg = (uint32_t)(n - 2); r.c = pl.pc[n - 1]; return r; }
    curdist = dist2(pl.pc[init.seg + 1], init.c);
    if (curdist >= distance) { ratio = distance / curdist; r.seg = init.seg; r.c = lerp2(init.c, pl.pc[init.seg + 1], ratio); return r; }
    int i;
    for (i = (int)init.seg + 1; i < n - 1; i++) {
      prevdist = curdist;
      curdist = dist2(pl.pc[i + 1], init.c);
      if (curdist >= distance) break;
    }
    if (i == n - 1) { reached = true; r.seg = (uint32_t)(n - 2); r.c = pl.pc[n - 1]; return r; }
    ratio = (distance - prevdist) / (curdist - prevdist);
    r.seg = (uint32_t)i; r.c = lerp2(pl.pc[i], pl.pc[i + 1], ratio); return r;
  }
  reached = true;
  return init;
}

// polyline::next_pl_point_by_line_intersection[_bounded_distance], polyline_graph_2d.cpp:579-780.
// bounded = true applies the [min_dist, max_dist] window to the first intersection found (:689-699).
EG3D_HD_NI bool walk_line(const Pl& pl, PlP init, uint32_t dir, float3 line, const eg3d_params& prm, bool bounded, PlP& next) {
  const float qc = prm.quasiparallel_cos, qd = prm.quasiparallel_dist;
  float2 inter = make_float2(0.f, 0.f);
  int r;
  const int n = pl.n;
  bool found = false;
  if (dir == pl.start) {
    float2 b = pl.pc[init.seg];
    r = isect_seg_line_nqp(init.c.x, init.c.y, b.x, b.y, line, qc, qd, inter);
    if (r & 2) return false;
    if (r & 1) { next.seg = init.seg; next.c = inter; found = true; }
    else {
      for (int i = (int)init.seg; i > 0; i--) {
        float2 a = pl.pc[i], c = pl.pc[i - 1];
        r = isect_seg_line_nqp(a.x, a.y, c.x, c.y, line, qc, qd, inter);
        if (r & 2) return false;
        if (r & 1) { next.seg = (uint32_t)(i - 1); next.c = inter; found = true; break; }
      }
    }
  } else if (dir == pl.end) {
    float2 b = pl.pc[init.seg + 1];
    r = isect_seg_line_nqp(init.c.x, init.c.y, b.x, b.y, line, qc, qd, inter);
    if (r & 2) return false;
    if (r & 1) { next.seg = init.seg; next.c = inter; found = true; }
    else {
      for (int i = (int)init.seg + 1; i < n - 1; i++) {
        float2 a = pl.pc[i], c = pl.pc[i + 1];
        r = isect_seg_line_nqp(a.x, a.y, c.x, c.y, line, qc, qd, inter);
        if (r & 2) return false;
        if (r & 1) { next.seg = (uint32_t)i; next.c = inter; found = true; break; }
      }
    }
  }
  if (found && bounded) {
    float dsq = sqdist2(next.c, init.c);
    if (dsq < (prm.follow_corr_min * prm.follow_corr_min) || dsq > (prm.follow_corr_max * prm.follow_corr_max)) found = false;
  }
  return found;
}

// minimum_distancesq, geometric_utilities.cpp:940-954
EG3D_HD_NI float min_distsq_seg(float2 p, float2 v, float2 w, float2& proj) {
  const float l2 = sqdist2(v, w);
  if (l2 == 0.0) { proj = v; return sqdist2(p, v); }
  float pvx = p.x - v.x, pvy = p.y - v.y, wvx = w.x - v.x, wvy = w.y - v.y;
  float q = (pvx * wvx + pvy * wvy) / l2;
  float m = (q < 1.0f) ? q : 1.0f;      // std::min<float>(1, q)
  float t = (0.0f < m) ? m : 0.0f;      // std::max<float>(0, m)
  proj.x = v.x + t * wvx; proj.y = v.y + t * wvy;
  return sqdist2(p, proj);
}

// polyline::compute_distancesq, polyline_graph_2d.cpp:845-862
EG3D_HD_NI float pl_distancesq(const Pl& pl, float2 p, uint32_t& seg, float2& proj) {
  float md = min_distsq_seg(p, pl.pc[0], pl.pc[1], proj);
  seg = 0;
  float2 cp;
  for (int i = 2; i < pl.n; i++) {
    float cur = min_distsq_seg(p, pl.pc[i - 1], pl.pc[i], cp);
    if (cur < md) { md = cur; proj = cp; seg = (uint32_t)(i - 1); }
  }
  return md;
}

#ifdef __CUDACC__
// Warp-cooperative walk_line for warp-uniform arguments (same result in every lane, identical to walk_line): the walk
// visits the partial segment from the current point to the next vertex and then whole segments, and stops at the first
// one that is intersected (or aborts on the first quasi-parallel one, which is tested first).  Lane m takes walk
// position base + m, so the first lane with a non-zero verdict is the segment the sequential walk would stop at.
static __device__ __noinline__ bool walk_line_warp(const Pl& pl, PlP init, uint32_t dir, float3 line, const eg3d_params& prm, bool bounded, PlP& next, int lane) {
  const int n = pl.n;
  const bool to_start = dir == pl.start;
  if (!to_start && dir != pl.end) return false;
  // positions: 0 = partial segment; to_start: m >= 1 <-> i = seg-(m-1) >= 1; to_end: m >= 1 <-> i = seg+m <= n-2
  const int total = to_start ? (int)init.seg + 1 : n - 1 - (int)init.seg;
  bool found = false;
  for (int base = 0; base < total; base += 32) {
    const int m = base + lane;
    int r = 0; float2 inter = make_float2(0.f, 0.f); uint32_t sg = 0;
    if (m < total) {
      float2 a, b;
      if (m == 0) { a = init.c; b = to_start ? pl.pc[init.seg] : pl.pc[init.seg + 1]; sg = init.seg; }
      else if (to_start) { const int i = (int)init.seg - (m - 1); a = pl.pc[i]; b = pl.pc[i - 1]; sg = (uint32_t)(i - 1); }
      else { const int i = (int)init.seg + m; a = pl.pc[i]; b = pl.pc[i + 1]; sg = (uint32_t)i; }
      r = isect_seg_line_nqp(a.x, a.y, b.x, b.y, line, prm.quasiparallel_cos, prm.quasiparallel_dist, inter);
    }
    const unsigned any = __ballot_sync(0xffffffffu, r != 0);
    if (any) {
      const int w = __ffs(any) - 1;
      if (__shfl_sync(0xffffffffu, r, w) & 2) return false;
      next.seg = __shfl_sync(0xffffffffu, sg, w);
      next.c.x = __shfl_sync(0xffffffffu, inter.x, w); next.c.y = __shfl_sync(0xffffffffu, inter.y, w);
      found = true;
      break;
    }
  }
  if (found && bounded) {
    float dsq = sqdist2(next.c, init.c);
    if (dsq < (prm.follow_corr_min * prm.follow_corr_min) || dsq > (prm.follow_corr_max * prm.follow_corr_max)) found = false;
  }
  return found;
}

// Warp-cooperative pl_distancesq for warp-uniform arguments: the sequential scan keeps the FIRST segment that attains
// the minimum (strict <), i.e. the arg-min with ties to the lower index; the lanes take segments lane, lane+32, ... and
// the 32 candidates are reduced with the same ordering.
static __device__ __noinline__ float pl_distancesq_warp(const Pl& pl, float2 p, uint32_t& seg, float2& proj, int lane) {
  float md = 0.f; int ms = 0x7fffffff; float2 mp = make_float2(0.f, 0.f);
  for (int sgi = lane; sgi < pl.n - 1; sgi += 32) {
    float2 cp;
    const float cur = min_distsq_seg(p, pl.pc[sgi], pl.pc[sgi + 1], cp);
    if (ms == 0x7fffffff || cur < md) { md = cur; ms = sgi; mp = cp; }
  }
#pragma unroll 1
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, md, o);
    const int os = __shfl_xor_sync(0xffffffffu, ms, o);
    const float ox = __shfl_xor_sync(0xffffffffu, mp.x, o), oy = __shfl_xor_sync(0xffffffffu, mp.y, o);
    if (os != 0x7fffffff && (ms == 0x7fffffff || od < md || (od == md && os < ms))) { md = od; ms = os; mp.x = ox; mp.y = oy; }
  }
  seg = (uint32_t)ms; proj = mp;
  return md;
}
#endif

// edge_graph_3d_utilities.cpp:600-629
EG3D_HD float floor_or_upper_if_close(float v) {
  float c = ceilf(v);
  if ((double)(c - v) < 0.001) return c;
  return floorf(v);
}
EG3D_HD bool is_multiple_of(float m, float n) {
  float div = m / n;
  float mul = floor_or_upper_if_close(div) * n;
  return (double)fabsf(m - mul) < 0.001;
}

// compute_projection, geometric_utilities.cpp:973-983 (glm row-vector product)
EG3D_HD_NI float2 project(const float* P, float X, float Y, float Z) {
  float h0 = P[0] * X + P[1] * Y + P[2] * Z + P[3] * 1.0f;
  float h1 = P[4] * X + P[5] * Y + P[6] * Z + P[7] * 1.0f;
  float h2 = P[8] * X + P[9] * Y + P[10] * Z + P[11] * 1.0f;
  return make_float2(h0 / h2, h1 / h2);
}

// PolyLine2DMapSearch::find_unique_polyline_potentially_within_search_dist, polyLine_2d_map_search.cpp:46-88:
// valid iff the union of the (clipped) 3x3 cell neighbourhood holds exactly one polyline id.
// The "row" flag (x on a cell boundary) clips the ROW loop.
template <typename Visit>
EG3D_HD void grid_visit(const DevGrid& g, int view, int img_w, int img_h, float2 c, Visit visit) {
  if (c.x <= 0 || c.x >= img_w || c.y <= 0 || c.y >= img_h) return;
  bool on_row = is_multiple_of(c.x, g.cell);
  bool on_col = is_multiple_of(c.y, g.cell);
  int cx = (int)floor_or_upper_if_close(c.x / g.cell), cy = (int)floor_or_upper_if_close(c.y / g.cell);
  if (cx >= g.w) cx = g.w - 1;
  if (cy >= g.h) cy = g.h - 1;
  int i0 = cy > 0 ? -1 : 0, i1 = on_row ? 0 : (cy < g.h - 1 ? 1 : 0);
  int j0 = cx > 0 ? -1 : 0, j1 = on_col ? 0 : (cx < g.w - 1 ? 1 : 0);
  const int* off = g.cell_off + (size_t)view * g.w * g.h;
  for (int i = i0; i <= i1; i++)
    for (int j = j0; j <= j1; j++) {
      int cidx = (cy + i) * g.w + (cx + j);
      for (int k = off[cidx]; k < off[cidx + 1]; k++) visit(g.ids[k]);
    }
}
EG3D_HD_NI bool grid_unique(const DevGrid& g, int view, int img_w, int img_h, float2 c, uint32_t& pl_id) {
  int cnt = 0; uint32_t first = 0; bool multi = false;
  grid_visit(g, view, img_w, img_h, c, [&](uint32_t id) {
    if (cnt == 0) { first = id; cnt = 1; } else if (id != first) multi = true;
  });
  pl_id = first;
  return cnt == 1 && !multi;
}

#ifdef __CUDACC__
// Warp-cooperative grid_unique for warp-uniform arguments: the (up to) nine cells of the clipped 3x3 neighbourhood are
// read by nine lanes at once; the union holds exactly one polyline id iff every non-empty cell holds only that id.
static __device__ __noinline__ bool grid_unique_warp(const DevGrid& g, int view, int img_w, int img_h, float2 c, uint32_t& pl_id, int lane) {
  if (c.x <= 0 || c.x >= img_w || c.y <= 0 || c.y >= img_h) return false;
  const bool on_row = is_multiple_of(c.x, g.cell);
  const bool on_col = is_multiple_of(c.y, g.cell);
  int cx = (int)floor_or_upper_if_close(c.x / g.cell), cy = (int)floor_or_upper_if_close(c.y / g.cell);
  if (cx >= g.w) cx = g.w - 1;
  if (cy >= g.h) cy = g.h - 1;
  const int i0 = cy > 0 ? -1 : 0, i1 = on_row ? 0 : (cy < g.h - 1 ? 1 : 0);
  const int j0 = cx > 0 ? -1 : 0, j1 = on_col ? 0 : (cx < g.w - 1 ? 1 : 0);
  const int i = lane / 3 - 1, j = lane % 3 - 1;            // lanes 0..8 <-> (i, j) in {-1,0,1}^2
  int cnt = 0; uint32_t first = 0; bool multi = false;
  if (lane < 9 && i >= i0 && i <= i1 && j >= j0 && j <= j1) {
    const int* off = g.cell_off + (size_t)view * g.w * g.h;
    const int cidx = (cy + i) * g.w + (cx + j);
    const int k0 = off[cidx], k1 = off[cidx + 1];
    cnt = k1 - k0;
    if (cnt > 0) first = g.ids[k0];
    for (int k = k0 + 1; k < k1; k++) if (g.ids[k] != first) multi = true;
  }
  const unsigned have = __ballot_sync(0xffffffffu, cnt > 0);
  if (__any_sync(0xffffffffu, multi) || have == 0) { pl_id = 0; return false; }
  const uint32_t f0 = __shfl_sync(0xffffffffu, first, __ffs(have) - 1);
  const bool same = __all_sync(0xffffffffu, cnt == 0 || first == f0);
  pl_id = f0;
  return same;
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// 2-view DLT initialiser = cv::triangulatePoints (triangulation.cpp:216,290): null vector of the 4x4 DLT matrix by
// one-sided Jacobi SVD in double, cast to float.  Same operation sequence as the oracle's restatement.
EG3D_HD_NI void dlt_null(const float* P1, const float* P2, float2 x1, float2 x2, float out4[4]) {
  double A[4][4], Vm[4][4];   // runtime-indexed on purpose (compact code); same operation order as the oracle
  for (int k = 0; k < 4; k++) {
    A[0][k] = (double)x1.x * (double)P1[8 + k] - (double)P1[k];
    A[1][k] = (double)x1.y * (double)P1[8 + k] - (double)P1[4 + k];
    A[2][k] = (double)x2.x * (double)P2[8 + k] - (double)P2[k];
    A[3][k] = (double)x2.y * (double)P2[8 + k] - (double)P2[4 + k];
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) Vm[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
#pragma unroll 1
    for (int p = 0; p < 3; p++)
#pragma unroll 1
      for (int q = p + 1; q < 4; q++) {
        double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) { alpha += A[i][p] * A[i][p]; beta += A[i][q] * A[i][q]; gamma += A[i][p] * A[i][q]; }
        if (gamma != 0) {
          double o = fabs(gamma) / sqrt(alpha * beta + 1e-300);
          off = off > o ? off : o;
          double zeta = (beta - alpha) / (2.0 * gamma);
          double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            double ap = A[i][p], aq = A[i][q];
            A[i][p] = c * ap - s * aq; A[i][q] = s * ap + c * aq;
            double vp = Vm[i][p], vq = Vm[i][q];
            Vm[i][p] = c * vp - s * vq; Vm[i][q] = s * vp + c * vq;
          }
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

// ---------------------------------------------------------------------------------------------------------------
// The same initialiser with OpenCV's OWN SVD restated (prepared this round on the host, to be selected by
// eg3d_params.dlt_wellposed == 2 in the kernels next round): cv::triangulatePoints (OpenCV 4.x triangulate.cpp: the 4x4
// system above) -> cv::SVD::compute -> JacobiSVDImpl_<double> (core/src/lapack.cpp) on the transposed matrix, squared norms
// cached in W, a pair is skipped when |p| <= 10 eps sqrt(a b), rotation from hypot(2p, a-b), <= 30 sweeps, singular values
// sorted descending by selection sort with the rows of Vt swapped along, result = last row of Vt as float.  Why it matters:
// the reference's get_min_max quirk hands the DLT the same camera twice in 1-12 % of the fresh triangulations of the real
// example; the null space is then a whole ray and only this algorithm in this operation order returns the point of it that
// a reference linked against OpenCV returns (bit-identical to cv2 4.13, sign included: tests/test_cabi_host.py), while the
// accepted point set depends on that point.  hypot_cr: a correctly rounded hypot from IEEE operations only
// (sqrt of the double-double sum of squares + one exact-residual correction), so host and device agree bit for bit.
EG3D_HD_NI double hypot_cr(double x, double y) {
  x = fabs(x); y = fabs(y);
  if (x < y) { const double t = x; x = y; y = t; }
  if (y == 0) return x;
  int ex;
  frexp(x, &ex);                                      // scale the larger argument into [0.5, 1): exact, no under/overflow below
  x = ldexp(x, -ex); y = ldexp(y, -ex);
  const double x2 = x * x, xx = fma(x, x, -x2), y2 = y * y, yy = fma(y, y, -y2);
  const double s = x2 + y2, e = y2 - (s - x2);        // Fast2Sum (x2 >= y2)
  const double lo = (xx + yy) + e;
  const double h = sqrt(s);
  return ldexp(h + (fma(-h, h, s) + lo) / (2 * h), ex);
}
EG3D_HD_NI void dlt_null_opencv(const float* P1, const float* P2, float2 x1, float2 x2, float out4[4]) {
  double At[4][4], Vt[4][4], W[4];   // At[i][k] = A[k][i]
  for (int k = 0; k < 4; k++) {
    At[k][0] = (double)x1.x * (double)P1[8 + k] - (double)P1[k];
    At[k][1] = (double)x1.y * (double)P1[8 + k] - (double)P1[4 + k];
    At[k][2] = (double)x2.x * (double)P2[8 + k] - (double)P2[k];
    At[k][3] = (double)x2.y * (double)P2[8 + k] - (double)P2[4 + k];
  }
  const double eps = 2.220446049250313e-16 * 10;
  for (int i = 0; i < 4; i++) {
    double sd = 0;
    for (int k = 0; k < 4; k++) { sd += At[i][k] * At[i][k]; Vt[i][k] = (i == k) ? 1.0 : 0.0; }
    W[i] = sd;
  }
#pragma unroll 1
  for (int iter = 0; iter < 30; iter++) {
    bool changed = false;
#pragma unroll 1
    for (int i = 0; i < 3; i++)
#pragma unroll 1
      for (int j = i + 1; j < 4; j++) {
        double a = W[i], p = 0, b = W[j];
        for (int k = 0; k < 4; k++) p += At[i][k] * At[j][k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = hypot_cr(p, beta);
        double c, sn;
        if (beta < 0) { const double delta = (gamma - beta) * 0.5; sn = sqrt(delta / gamma); c = p / (gamma * sn * 2); }
        else { c = sqrt((gamma + beta) / (gamma * 2)); sn = p / (gamma * c * 2); }
        a = b = 0;
        for (int k = 0; k < 4; k++) {
          const double t0 = c * At[i][k] + sn * At[j][k], t1 = -sn * At[i][k] + c * At[j][k];
          At[i][k] = t0; At[j][k] = t1;
          a += t0 * t0; b += t1 * t1;
        }
        W[i] = a; W[j] = b;
        changed = true;
        for (int k = 0; k < 4; k++) {
          const double t0 = c * Vt[i][k] + sn * Vt[j][k], t1 = -sn * Vt[i][k] + c * Vt[j][k];
          Vt[i][k] = t0; Vt[j][k] = t1;
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < 4; i++) {
    double sd = 0;
    for (int k = 0; k < 4; k++) sd += At[i][k] * At[i][k];
    W[i] = sqrt(sd);
  }
  int row[4] = {0, 1, 2, 3};           // the selection sort only has to track which row of Vt ends up last
  for (int i = 0; i < 3; i++) {
    int j = i;
    for (int k = i + 1; k < 4; k++) if (W[j] < W[k]) j = k;
    if (i != j) { const double tw = W[i]; W[i] = W[j]; W[j] = tw; const int tr = row[i]; row[i] = row[j]; row[j] = tr; }
  }
  for (int k = 0; k < 4; k++) out4[k] = (float)Vt[row[3]][k];
}

EG3D_HD double det3d(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
EG3D_HD void inv3d(const double* m, double d, double* t) {  // cv::invert closed form: cofactors * (1/det)
  d = 1. / d;
  t[0] = (m[4] * m[8] - m[5] * m[7]) * d; t[1] = (m[2] * m[7] - m[1] * m[8]) * d; t[2] = (m[1] * m[5] - m[2] * m[4]) * d;
  t[3] = (m[5] * m[6] - m[3] * m[8]) * d; t[4] = (m[0] * m[8] - m[2] * m[6]) * d; t[5] = (m[2] * m[3] - m[0] * m[5]) * d;
  t[6] = (m[3] * m[7] - m[4] * m[6]) * d; t[7] = (m[1] * m[6] - m[0] * m[7]) * d; t[8] = (m[0] * m[4] - m[1] * m[3]) * d;
}

// ---------------------------------------------------------------------------------------------------------------
// em_GaussNewton for exactly three observations (triangulation.cpp:105-176), thread-local, SAME operation order as the
// oracle (J and r kept in registers, H = J^T J accumulated row by row, update = (H^-1 J^T) r) => bit-identical results.
// Returns true when accepted (last_mse < accept); X holds the result.
EG3D_HD bool gn3_exact(const DevScene& S, const int v[3], const float2 pt[3], double X[3]) {
  // Deliberately rolled loops with J / r in (L1-resident) local arrays: the routine is called from many divergent
  // sites and an unrolled body (24 double divisions) would not fit the instruction caches.
  double last_mse = 0;
  const eg3d_params& prm = S.prm;
  double r[6], J[18];
  for (int it = 0; it < prm.gn_max_iters; it++) {
    double mse = 0;
#pragma unroll 1
    for (int m = 0; m < 3; m++) {
      const float* Pf = S.P + 12 * v[m];
      double P[12];
#pragma unroll
      for (int i = 0; i < 12; i++) P[i] = (double)Pf[i];
      double h0 = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3] * 1.0;
      double h1 = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7] * 1.0;
      double h2 = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] * 1.0;
      double rx = (double)pt[m].x - h0 / h2;
      mse += rx * rx;
      double ry = (double)pt[m].y - h1 / h2;
      mse += ry * ry;
      r[2 * m] = rx; r[2 * m + 1] = ry;
      double zz = h2 * h2;
      J[6 * m + 0] = (P[0] * h2 - P[8] * h0) / zz;  J[6 * m + 3] = (P[4] * h2 - P[8] * h1) / zz;
      J[6 * m + 1] = (P[1] * h2 - P[9] * h0) / zz;  J[6 * m + 4] = (P[5] * h2 - P[9] * h1) / zz;
      J[6 * m + 2] = (P[2] * h2 - P[10] * h0) / zz; J[6 * m + 5] = (P[6] * h2 - P[10] * h1) / zz;
    }
    if (fabs(mse / 6 - last_mse) < prm.gn_stop) break;
    last_mse = mse / 6;
    double H[9];
#pragma unroll
    for (int a = 0; a < 9; a++) H[a] = 0;
#pragma unroll 1
    for (int k = 0; k < 6; k++) {
      double j0 = J[3 * k], j1 = J[3 * k + 1], j2 = J[3 * k + 2];
      H[0] += j0 * j0; H[1] += j0 * j1; H[2] += j0 * j2;
      H[3] += j1 * j0; H[4] += j1 * j1; H[5] += j1 * j2;
      H[6] += j2 * j0; H[7] += j2 * j1; H[8] += j2 * j2;
    }
    double d = det3d(H);
    if (d < prm.gn_det_min) return false;
    double Hi[9]; inv3d(H, d, Hi);
    double a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 1
    for (int k = 0; k < 6; k++) {
      double j0 = J[3 * k], j1 = J[3 * k + 1], j2 = J[3 * k + 2], rk = r[k];
      a0 += (Hi[0] * j0 + Hi[1] * j1 + Hi[2] * j2) * rk;
      a1 += (Hi[3] * j0 + Hi[4] * j1 + Hi[5] * j2) * rk;
      a2 += (Hi[6] * j0 + Hi[7] * j1 + Hi[8] * j2) * rk;
    }
    X[0] += a0; X[1] += a1; X[2] += a2;
  }
  return last_mse < prm.gn_accept_mse;
}

// em_estimate3Dpositions for three observations (triangulation.cpp:252-323) incl. the get_min_max quirk
// (edge_graph_3d_utilities.hpp:69-92: min = first arg-min, "max" = last index) and the well-posed substitution.
EG3D_HD_NI bool est3(const DevScene& S, const int v[3], const float2 pt[3], float Xo[3]) {
  int mi = 0;
  if (v[1] < v[mi]) mi = 1;
  if (v[2] < v[mi]) mi = 2;
  int ma = 2;
  if (S.prm.dlt_wellposed == 1 && v[ma] == v[mi]) {
    if (v[2] != v[mi]) ma = 2; else if (v[1] != v[mi]) ma = 1; else if (v[0] != v[mi]) ma = 0;
  }
  float2 pmi = pt[0], pma = pt[0];
  int vmi = v[0], vma = v[0];
  if (mi == 1) { pmi = pt[1]; vmi = v[1]; } else if (mi == 2) { pmi = pt[2]; vmi = v[2]; }
  if (ma == 1) { pma = pt[1]; vma = v[1]; } else if (ma == 2) { pma = pt[2]; vma = v[2]; }
  float t4[4];
  dlt_null_opencv(S.P + 12 * vmi, S.P + 12 * vma, pmi, pma, t4);
  double X[3] = {(double)(t4[0] / t4[3]), (double)(t4[1] / t4[3]), (double)(t4[2] / t4[3])};
  if (!gn3_exact(S, v, pt, X)) return false;
  Xo[0] = (float)X[0]; Xo[1] = (float)X[1]; Xo[2] = (float)X[2];
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Gauss-Newton over n observations, one problem per THREAD, observations visited in order (same summation order as
// the reference; the normal-equation right-hand side is accumulated as J^T r instead of (H^-1 J^T) r, a difference of
// rounding only).  obs(i, view, x, y) yields observation i.
struct GnAcc { double mse, h00, h01, h02, h11, h12, h22, g0, g1, g2; };

EG3D_D void gn_accumulate(const float* __restrict__ P, float px, float py, const double X[3], GnAcc& a) {
  double p0 = P[0], p1 = P[1], p2 = P[2], p3 = P[3], p4 = P[4], p5 = P[5], p6 = P[6], p7 = P[7], p8 = P[8], p9 = P[9], p10 = P[10], p11 = P[11];
  double h0 = p0 * X[0] + p1 * X[1] + p2 * X[2] + p3 * 1.0;
  double h1 = p4 * X[0] + p5 * X[1] + p6 * X[2] + p7 * 1.0;
  double h2 = p8 * X[0] + p9 * X[1] + p10 * X[2] + p11 * 1.0;
  double rx = (double)px - h0 / h2;
  a.mse += rx * rx;
  double ry = (double)py - h1 / h2;
  a.mse += ry * ry;
  double zz = h2 * h2;
  double jx0 = (p0 * h2 - p8 * h0) / zz, jx1 = (p1 * h2 - p9 * h0) / zz, jx2 = (p2 * h2 - p10 * h0) / zz;
  double jy0 = (p4 * h2 - p8 * h1) / zz, jy1 = (p5 * h2 - p9 * h1) / zz, jy2 = (p6 * h2 - p10 * h1) / zz;
  a.h00 += jx0 * jx0; a.h00 += jy0 * jy0;
  a.h01 += jx0 * jx1; a.h01 += jy0 * jy1;
  a.h02 += jx0 * jx2; a.h02 += jy0 * jy2;
  a.h11 += jx1 * jx1; a.h11 += jy1 * jy1;
  a.h12 += jx1 * jx2; a.h12 += jy1 * jy2;
  a.h22 += jx2 * jx2; a.h22 += jy2 * jy2;
  a.g0 += jx0 * rx; a.g0 += jy0 * ry;
  a.g1 += jx1 * rx; a.g1 += jy1 * ry;
  a.g2 += jx2 * rx; a.g2 += jy2 * ry;
}

// one GN update from fully reduced sums; returns 0 = continue, 1 = converged (break), -1 = det failure
EG3D_D int gn_update(const GnAcc& a, int n, const eg3d_params& prm, double& last_mse, double X[3]) {
  double cur = a.mse / (n * 2);
  if (fabs(cur - last_mse) < prm.gn_stop) return 1;
  last_mse = cur;
  double H[9] = {a.h00, a.h01, a.h02, a.h01, a.h11, a.h12, a.h02, a.h12, a.h22};
  double d = det3d(H);
  if (d < prm.gn_det_min) return -1;
  double Hi[9]; inv3d(H, d, Hi);
  X[0] += Hi[0] * a.g0 + Hi[1] * a.g1 + Hi[2] * a.g2;
  X[1] += Hi[3] * a.g0 + Hi[4] * a.g1 + Hi[5] * a.g2;
  X[2] += Hi[6] * a.g0 + Hi[7] * a.g1 + Hi[8] * a.g2;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Exact-safe pruning of hopeless Gauss-Newton problems.
// em_GaussNewton accepts iff last_mse < 9 (triangulation.cpp:168) where last_mse is the mean squared residual at one of
// its iterates, hence >= the global minimum over X, hence >= (minimal cost of ANY two of the observations) / (2n).
// For two observations pa (view a), pb (view b) every X reprojects to (qa, qb) with qb on the epipolar line F qa of the
// cameras' own geometry (Fp), so the 2-view cost is >= min_t [ t^2 + dist(pb, line(F(pa + d)))^2 ], |d| = t.  With
// s0 = pb~^T F pa~, g = |(F^T pb~)_xy|, n0 = |(F pa~)_xy|, h = |F[0:2,0:2]|_F:
//   dist >= (|s0| - g t) / (n0 + h t)   (numerator falls, denominator grows with t)
// so cost >= min(T^2, ((|s0| - g T)/(n0 + h T))^2) for any T.  If that is >= budget = accept_mse * 2n (with a safety
// margin for rounding) the solve cannot be accepted and is skipped; a rejected solve has no side effects in the
// reference, so results are unchanged.  (A first-iteration break needs mse < 3e-6 and cannot occur here.)
// One out-of-line copy (the kernels are bound by instruction-cache refills): h = |F[0:2,0:2]|_F comes precomputed, and
// the two conditions are compared in squared form, which needs one square root instead of three:
//   |s0| > g T                       <=>  s0^2 > g^2 T^2
//   |s0| - g T >= T (n0 + h T)       <=>  B := |s0| - h T^2 > 0  and  B^2 >= T^2 (g^2 + n0^2 + 2 sqrt(g^2 n0^2))
static __device__ __noinline__ bool pair_cannot_fit(const double* __restrict__ F, double h, float2 pa, float2 pb, double T) {
  const double ax = pa.x, ay = pa.y, bx = pb.x, by = pb.y;
  const double l0 = fma(F[0], ax, fma(F[1], ay, F[2])), l1 = fma(F[3], ax, fma(F[4], ay, F[5])), l2 = fma(F[6], ax, fma(F[7], ay, F[8]));
  const double s0 = fabs(fma(bx, l0, fma(by, l1, l2)));
  const double n2 = fma(l0, l0, l1 * l1);
  const double m0 = fma(F[0], bx, fma(F[3], by, F[6])), m1 = fma(F[1], bx, fma(F[4], by, F[7]));
  const double g2 = fma(m0, m0, m1 * m1);
  const double T2 = T * T;
  if (!(s0 * s0 > g2 * T2)) return false;
  const double B = fma(-h, T2, s0);
  if (!(B > 0)) return false;
  return B * B >= T2 * (g2 + n2 + 2.0 * sqrt(g2 * n2));
}
EG3D_D double prune_radius(const eg3d_params& prm, int n_obs) {  // T with T^2 = accept_mse * 2n * 1.02
  return sqrt(prm.gn_accept_mse * (double)(2 * n_obs) * 1.02);
}

// Observation source shared by every multi-observation GN call of K3: the n observations of the arrays (v, x, y)
// followed by one optional extra observation.
struct ObsSrc {
  const int* v; const float* x; const float* y; int n;
  int has_extra; int ev; float ex, ey;
  // optional (gn_group only, null = off): the ten normal-equation sums of the n base observations at the slot's stored estimate,
  // and the observation count they were taken at (valid iff *tag == n); left there by the first solve that needs them
  double* cache; int* tag;
  EG3D_D int count() const { return n + has_extra; }
};

// Residual / Jacobian / normal-equation accumulation of one observation for the multi-observation solves of K3.
// Algebraically identical to em_point2D3DJacobian + the residual loop (triangulation.cpp:53-103,127-148) but arranged
// for FP64 throughput: one reciprocal of the projective depth instead of eight divisions ((p0*z - p8*x)/z^2 ==
// (p0 - p8*(x/z))/z), explicit fused multiply-adds, cameras pre-widened to double.  Differs from the division form by
// rounding only (~1e-16 relative); the thresholded 3-view solves use gn3_exact instead.
EG3D_D void gn_accumulate_fast(const double* __restrict__ P, float px, float py, double X0, double X1, double X2, GnAcc& a) {
  // Staged for register pressure (the kernels run at 64 registers/thread): the depth row first, then the x row, then
  // the y row, each camera row dead before the next one is loaded.  Every accumulator still receives its x term and
  // then its y term, so the sums round exactly as in the one-shot form.
  const double2* P2 = reinterpret_cast<const double2*>(P);
  const double2 q4 = P2[4], q5 = P2[5];
  const double inv = 1.0 / fma(q4.x, X0, fma(q4.y, X1, fma(q5.x, X2, q5.y)));
  {
    const double2 q0 = P2[0], q1 = P2[1];
    const double u = fma(q0.x, X0, fma(q0.y, X1, fma(q1.x, X2, q1.y))) * inv;
    const double rx = (double)px - u;
    const double jx0 = fma(-q4.x, u, q0.x) * inv, jx1 = fma(-q4.y, u, q0.y) * inv, jx2 = fma(-q5.x, u, q1.x) * inv;
    a.mse = fma(rx, rx, a.mse);
    a.h00 = fma(jx0, jx0, a.h00); a.h01 = fma(jx0, jx1, a.h01); a.h02 = fma(jx0, jx2, a.h02);
    a.h11 = fma(jx1, jx1, a.h11); a.h12 = fma(jx1, jx2, a.h12); a.h22 = fma(jx2, jx2, a.h22);
    a.g0 = fma(jx0, rx, a.g0); a.g1 = fma(jx1, rx, a.g1); a.g2 = fma(jx2, rx, a.g2);
  }
  {
    const double2 q2 = P2[2], q3 = P2[3];
    const double v = fma(q2.x, X0, fma(q2.y, X1, fma(q3.x, X2, q3.y))) * inv;
    const double ry = (double)py - v;
    const double jy0 = fma(-q4.x, v, q2.x) * inv, jy1 = fma(-q4.y, v, q2.y) * inv, jy2 = fma(-q5.x, v, q3.x) * inv;
    a.mse = fma(ry, ry, a.mse);
    a.h00 = fma(jy0, jy0, a.h00); a.h01 = fma(jy0, jy1, a.h01); a.h02 = fma(jy0, jy2, a.h02);
    a.h11 = fma(jy1, jy1, a.h11); a.h12 = fma(jy1, jy2, a.h12); a.h22 = fma(jy2, jy2, a.h22);
    a.g0 = fma(jy0, ry, a.g0); a.g1 = fma(jy1, ry, a.g1); a.g2 = fma(jy2, ry, a.g2);
  }
}

// Gauss-Newton (em_GaussNewton, triangulation.cpp:105-176) for up to 32/G independent problems per warp: the warp is
// split into groups of G lanes (G a power of two, 1..32); group g = lane / G solves the problem described by `o`
// (identical in all lanes of a group), its lanes stride over the observations (the optional extra observation is index
// n) and the ten sums are butterfly-reduced inside the group, so every lane of a group holds the same iterate.  Must be
// called by all 32 lanes; groups without a problem pass active = false.  Returns the accept decision (last_mse < 9) of
// the caller's group.  Loop-invariant scalars are re-read from the (constant-bank) scene instead of being kept in
// registers: the observation loop has to stay spill-free at 64 registers.
#ifdef EG3D_K3_PROFILE
__device__ unsigned long long g_gnprof[8];   // [0] calls, [1] iterations, [2] observation-loop passes (max over lanes), [3] active lanes x iterations
#endif
// Shared first iteration (GC = true: the kernel variant used for rigs with many views, where the observation loop dominates; with
// GC = false the code below folds to the plain loop — on BASELINE configs[1] merely carrying the extra code cost 6 %): every
// problem "observations of slot s + one new observation", started from the slot's stored estimate, begins with the same sums
// over the slot's n observations.  The first iteration takes them from the slot's cache (valid while the slot is unchanged:
// *tag == n) or computes and leaves them there, then adds the new observation's terms — the same value up to the order of
// summation.
template <bool GC>
static __device__ __noinline__ bool gn_group(const DevScene& S, const ObsSrc& o, bool active, int G, int lane, double X[3]) {
  const int* ov = o.v;
  const float* ox = o.x;
  const float* oy = o.y;
  const int n = o.n, ntot = o.n + o.has_extra;
  const int sub = lane & (G - 1);
  double X0 = X[0], X1 = X[1], X2 = X[2];
  double last_mse = 0;
  bool running = active, failed = false;
  const bool use_cache = GC && active && o.has_extra;
  const bool cached = use_cache && *o.tag == n;
#ifdef EG3D_K3_PROFILE
  if (lane == 0) atomicAdd(&g_gnprof[0], 1ull);
#endif
  for (int it = 0; it < S.prm.gn_max_iters; it++) {
    if (!__any_sync(0xffffffffu, running)) break;
#ifdef EG3D_K3_PROFILE
    {
      const int passes = running ? (ntot - sub + G - 1) / G : 0;
      int mp = passes;
      for (int o2 = 16; o2 > 0; o2 >>= 1) mp = max(mp, __shfl_xor_sync(0xffffffffu, mp, o2));
      const unsigned rm = __ballot_sync(0xffffffffu, running);
      if (lane == 0) { atomicAdd(&g_gnprof[1], 1ull); atomicAdd(&g_gnprof[2], (unsigned long long)mp); atomicAdd(&g_gnprof[3], (unsigned long long)__popc(rm)); }
    }
#endif
    GnAcc a = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const bool first_shared = it == 0 && use_cache;
    if (running) {
      const double* __restrict__ P64 = S.P64;
      if (first_shared && cached) {
        if (sub == 0) { const double* q = o.cache; a.mse = q[0]; a.h00 = q[1]; a.h01 = q[2]; a.h02 = q[3]; a.h11 = q[4]; a.h12 = q[5]; a.h22 = q[6]; a.g0 = q[7]; a.g1 = q[8]; a.g2 = q[9]; }
      } else {
        const int nloop = first_shared ? n : ntot;
        // the next observation is fetched while the current one is accumulated: with ~1000 views a chain's observation lists no
        // longer fit the L2 and the loop was bound by the latency of these loads (long-scoreboard stalls, profiles/r02_c3_k3b.txt)
        int i = sub;
        int v = 0; float px = 0.f, py = 0.f;
        if (i < nloop) { if (i < n) { v = ov[i]; px = ox[i]; py = oy[i]; } else { v = o.ev; px = o.ex; py = o.ey; } }
#pragma unroll 1
        while (i < nloop) {
          const int j = i + G;
          int nv = 0; float npx = 0.f, npy = 0.f;
          if (j < nloop) { if (j < n) { nv = ov[j]; npx = ox[j]; npy = oy[j]; } else { nv = o.ev; npx = o.ex; npy = o.ey; } }
          gn_accumulate_fast(P64 + 12 * v, px, py, X0, X1, X2, a);
          v = nv; px = npx; py = npy; i = j;
        }
      }
    }
#pragma unroll 1
    for (int off = G >> 1; off > 0; off >>= 1) {
      a.mse += __shfl_xor_sync(0xffffffffu, a.mse, off);
      a.h00 += __shfl_xor_sync(0xffffffffu, a.h00, off); a.h01 += __shfl_xor_sync(0xffffffffu, a.h01, off);
      a.h02 += __shfl_xor_sync(0xffffffffu, a.h02, off); a.h11 += __shfl_xor_sync(0xffffffffu, a.h11, off);
      a.h12 += __shfl_xor_sync(0xffffffffu, a.h12, off); a.h22 += __shfl_xor_sync(0xffffffffu, a.h22, off);
      a.g0 += __shfl_xor_sync(0xffffffffu, a.g0, off); a.g1 += __shfl_xor_sync(0xffffffffu, a.g1, off);
      a.g2 += __shfl_xor_sync(0xffffffffu, a.g2, off);
    }
    if (first_shared && running) {
      if (!cached && sub == 0) {
        double* q = o.cache; q[0] = a.mse; q[1] = a.h00; q[2] = a.h01; q[3] = a.h02; q[4] = a.h11; q[5] = a.h12; q[6] = a.h22; q[7] = a.g0; q[8] = a.g1; q[9] = a.g2;
        *o.tag = n;
      }
      gn_accumulate_fast(S.P64 + 12 * o.ev, o.ex, o.ey, X0, X1, X2, a);     // the new observation (every lane of the group: identical values)
    }
    if (running) {
      const double cur = a.mse / (ntot * 2);
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
  X[0] = X0; X[1] = X1; X[2] = X2;
  if (GC) __syncwarp();                           // cache rows written above are read by other lanes in later calls
  return active && !failed && last_mse < S.prm.gn_accept_mse;
}

// em_GaussNewton (triangulation.cpp:105-176) over n observations in the reference's own operation order — true divisions, no
// FMA, every sum accumulated observation by observation — so the iterates are bit-identical to the oracle's (as gn3_exact is
// for three observations).  Used where Gauss-Newton starts from the 2-view DLT point instead of a converged estimate
// (est_slot, combos_slot): from such a start, and above all from the arbitrary point of the ray a degenerate DLT returns, the
// iteration can wander for tens of steps and amplifies rounding-level differences into different accept decisions, which the
// lane-parallel gn_group (different summation order) would not reproduce.
// Warp-uniform arguments.  The residual / Jacobian rows of 32 observations are computed one per lane (the expensive part: eight
// IEEE divisions each); the sums are then accumulated in observation order by every lane from the broadcast rows, which keeps
// the reference's summation order exactly.  Rare: ~1 call per 50 gn_group calls on BASELINE configs[1].
struct GnRows { double rx, ry, jx0, jx1, jx2, jy0, jy1, jy2; };
EG3D_D GnRows gn_rows_exact(const DevScene& S, const ObsSrc& o, int m, const double X[3]) {
  int v; float px, py;
  if (m < o.n) { v = o.v[m]; px = o.x[m]; py = o.y[m]; } else { v = o.ev; px = o.ex; py = o.ey; }
  const double* __restrict__ P = S.P64 + 12 * v;
  const double h0 = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3] * 1.0;
  const double h1 = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7] * 1.0;
  const double h2 = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] * 1.0;
  GnRows r;
  r.rx = (double)px - h0 / h2; r.ry = (double)py - h1 / h2;
  const double zz = h2 * h2;
  r.jx0 = (P[0] * h2 - P[8] * h0) / zz; r.jx1 = (P[1] * h2 - P[9] * h0) / zz; r.jx2 = (P[2] * h2 - P[10] * h0) / zz;
  r.jy0 = (P[4] * h2 - P[8] * h1) / zz; r.jy1 = (P[5] * h2 - P[9] * h1) / zz; r.jy2 = (P[6] * h2 - P[10] * h1) / zz;
  return r;
}
EG3D_D GnRows gn_rows_bcast(const GnRows& r, int k) {
  GnRows b;
  b.rx = __shfl_sync(0xffffffffu, r.rx, k); b.ry = __shfl_sync(0xffffffffu, r.ry, k);
  b.jx0 = __shfl_sync(0xffffffffu, r.jx0, k); b.jx1 = __shfl_sync(0xffffffffu, r.jx1, k); b.jx2 = __shfl_sync(0xffffffffu, r.jx2, k);
  b.jy0 = __shfl_sync(0xffffffffu, r.jy0, k); b.jy1 = __shfl_sync(0xffffffffu, r.jy1, k); b.jy2 = __shfl_sync(0xffffffffu, r.jy2, k);
  return b;
}
static __device__ __noinline__ bool gn_seq_exact(const DevScene& S, const ObsSrc& o, int lane, double X[3]) {
  const eg3d_params& prm = S.prm;
  const int ntot = o.n + o.has_extra;
  double last_mse = 0;
  for (int it = 0; it < prm.gn_max_iters; it++) {
    double mse = 0;
    double h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0;      // H is symmetric term by term (j_a * j_b == j_b * j_a): six sums
#pragma unroll 1
    for (int base = 0; base < ntot; base += 32) {
      GnRows r = {0, 0, 0, 0, 0, 0, 0, 0};
      if (base + lane < ntot) r = gn_rows_exact(S, o, base + lane, X);
      const int cnt = min(32, ntot - base);
#pragma unroll 1
      for (int k = 0; k < cnt; k++) {
        const GnRows b = gn_rows_bcast(r, k);
        mse += b.rx * b.rx; mse += b.ry * b.ry;
        h00 += b.jx0 * b.jx0; h01 += b.jx0 * b.jx1; h02 += b.jx0 * b.jx2; h11 += b.jx1 * b.jx1; h12 += b.jx1 * b.jx2; h22 += b.jx2 * b.jx2;
        h00 += b.jy0 * b.jy0; h01 += b.jy0 * b.jy1; h02 += b.jy0 * b.jy2; h11 += b.jy1 * b.jy1; h12 += b.jy1 * b.jy2; h22 += b.jy2 * b.jy2;
      }
    }
    if (fabs(mse / (ntot * 2) - last_mse) < prm.gn_stop) break;
    last_mse = mse / (ntot * 2);
    const double H[9] = {h00, h01, h02, h01, h11, h12, h02, h12, h22};
    const double d = det3d(H);
    if (d < prm.gn_det_min) return false;
    double Hi[9]; inv3d(H, d, Hi);
    double a0 = 0, a1 = 0, a2 = 0;      // curEstimate += (H^-1 J^T) r, row by row
#pragma unroll 1
    for (int base = 0; base < ntot; base += 32) {
      GnRows r = {0, 0, 0, 0, 0, 0, 0, 0};
      if (base + lane < ntot) r = gn_rows_exact(S, o, base + lane, X);
      const int cnt = min(32, ntot - base);
#pragma unroll 1
      for (int k = 0; k < cnt; k++) {
        const GnRows b = gn_rows_bcast(r, k);
        a0 += (Hi[0] * b.jx0 + Hi[1] * b.jx1 + Hi[2] * b.jx2) * b.rx; a1 += (Hi[3] * b.jx0 + Hi[4] * b.jx1 + Hi[5] * b.jx2) * b.rx; a2 += (Hi[6] * b.jx0 + Hi[7] * b.jx1 + Hi[8] * b.jx2) * b.rx;
        a0 += (Hi[0] * b.jy0 + Hi[1] * b.jy1 + Hi[2] * b.jy2) * b.ry; a1 += (Hi[3] * b.jy0 + Hi[4] * b.jy1 + Hi[5] * b.jy2) * b.ry; a2 += (Hi[6] * b.jy0 + Hi[7] * b.jy1 + Hi[8] * b.jy2) * b.ry;
      }
    }
    X[0] += a0; X[1] += a1; X[2] += a2;
  }
  return last_mse < prm.gn_accept_mse;
}

// lanes per problem for `p` (1..32) simultaneous problems
EG3D_D int gn_group_width(int p) { return p > 16 ? 1 : p > 8 ? 2 : p > 4 ? 4 : p > 2 ? 8 : p > 1 ? 16 : 32; }

}  // namespace eg3d
