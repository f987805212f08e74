// eg3d_a6.cuh — seeding of pipeline 3 on the device:
// PLGEdgeManager::detect_nearby_intersections_and_correspondences_plgp (plg_edge_manager.cpp:261-288 with the helpers
// :191-259) as called by plg_matching_from_refpoints (plg_matching_from_refpoints.cpp:64-104).  For every observation
// of every SfM track: the polylines of the 3x3 neighbourhood of the 30 px grid (PolyLine2DMapSearch,
// polyLine_2d_map_search.cpp:46-88) in ascending id order, their distance to the observation
// (polyline::compute_distancesq, polyline_graph_2d.cpp:845-862); polylines within 30 px are the candidate set of that
// (track, view), those within 10 px also yield a seed = the projection onto the polyline, with the per-seed radius
// 3 x |observation - seed| that later filters the epipolar hits.  One warp per observation, two kernels (classify ->
// prefix sums -> fill); the arithmetic is the host/oracle functions of eg3d_dev.cuh compiled for the device.
#pragma once
#include "eg3d_dev.cuh"

namespace eg3d {

constexpr int A6_THREADS = 128;
// Per-observation capacities are run-time values (A6Args::raw_cap / ids_cap; dynamic shared memory): the host starts at
// A6_RAW / A6_IDS and re-runs the seeding with four times as much when an observation's neighbourhood holds more (the
// reference has no such bound), up to A6_RAW_MAX ids, where one warp per CTA still fits an SM's shared memory.
constexpr int A6_RAW = 512;      // ids gathered from the nine cells, duplicates included
constexpr int A6_IDS = 128;      // distinct polylines per observation
constexpr int A6_RAW_MAX = 16384;
inline __host__ __device__ size_t a6_smem_per_warp(int raw_cap, int ids_cap) { return (size_t)raw_cap * 9 + (size_t)ids_cap * 4; }   // raw + sorted (u32) + dup (u8) + uniq (u32)

struct A6Rec { uint32_t id, seg; float px, py; int cls; };   // cls: 2 = seed + candidate (<= 10 px), 1 = candidate (<= 30 px), 0 = neither

struct A6Args {
  int64_t o_begin, o_end;          // observation range = track_off[tb] .. track_off[te]
  int64_t tb;
  const int* obs_track;            // [NO] track of every observation
  int raw_cap, ids_cap;            // capacities per observation (multiples of 4)
  A6Rec* recs;                     // [n_obs][ids_cap]
  int* n_ids; int* n_cand; int* n_seed; unsigned char* is_last;
  int* overflow;
  // fill
  const int64_t* seed_off;         // [n_obs + 1]
  const int64_t* coff;             // [nt * V + 1] candidate CSR rows (track - tb, view)
  int64_t* row_cnt;                // [nt * V + 1] written by the classify kernel for the last observation of a (track, view)
  int* s_view; uint32_t* s_pl; uint32_t* s_seg; float2* s_xy; int* s_set; float* s_r2;
  uint32_t* cpl; float2* center;
};

__global__ void __launch_bounds__(A6_THREADS) a6_classify_kernel(const __grid_constant__ DevScene S, const __grid_constant__ A6Args A) {
  extern __shared__ __align__(16) unsigned char a6_smem[];
  const int A6_RAW = A.raw_cap, A6_IDS = A.ids_cap;        // (shadow the defaults: this launch's capacities)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t ol = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // local observation index
  const int64_t o = A.o_begin + ol;
  if (o >= A.o_end) return;
  unsigned char* mine = a6_smem + (size_t)wib * a6_smem_per_warp(A6_RAW, A6_IDS);
  uint32_t* raw = reinterpret_cast<uint32_t*>(mine); uint32_t* sorted = raw + A6_RAW; uint32_t* uniq = sorted + A6_RAW;
  unsigned char* dup = reinterpret_cast<unsigned char*>(uniq + A6_IDS);
  const int rp = A.obs_track[o];
  const int64_t t0 = S.track_off[rp], t1 = S.track_off[rp + 1];
  const int view = S.track_view[o];
  // get_2d_coordinates_of_point_on_image: the LAST observation of the track in this view wins
  int64_t jl = -1;
  for (int64_t j0 = t0; j0 < t1; j0 += 32) {
    const int64_t j = j0 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, j < t1 && S.track_view[j] == view);
    if (m) jl = j0 + 31 - __clz(m);
  }
  const float2 sp = S.track_xy[jl];
  const bool last = jl == o;
  // the nine cells (grid_visit, eg3d_dev.cuh)
  const DevGrid& g = S.g_corr;
  int nraw = 0;
  if (!(sp.x <= 0 || sp.x >= S.width || sp.y <= 0 || sp.y >= S.height)) {
    const bool on_row = is_multiple_of(sp.x, g.cell);
    const bool on_col = is_multiple_of(sp.y, g.cell);
    int cx = (int)floor_or_upper_if_close(sp.x / g.cell), cy = (int)floor_or_upper_if_close(sp.y / g.cell);
    if (cx >= g.w) cx = g.w - 1;
    if (cy >= g.h) cy = g.h - 1;
    const int i0 = cy > 0 ? -1 : 0, i1 = on_row ? 0 : (cy < g.h - 1 ? 1 : 0);
    const int j0 = cx > 0 ? -1 : 0, j1 = on_col ? 0 : (cx < g.w - 1 ? 1 : 0);
    const int i = lane / 3 - 1, j = lane % 3 - 1;
    int k0 = 0, k1 = 0;
    if (lane < 9 && i >= i0 && i <= i1 && j >= j0 && j <= j1) {
      const int* off = g.cell_off + (size_t)view * g.w * g.h;
      const int cidx = (cy + i) * g.w + (cx + j);
      k0 = off[cidx]; k1 = off[cidx + 1];
    }
    int cnt = k1 - k0, pre = cnt;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pre, d); if (lane >= d) pre += t; }
    nraw = __shfl_sync(0xffffffffu, pre, 8);
    pre -= cnt;
    if (nraw <= A6_RAW) for (int k = 0; k < cnt; k++) raw[pre + k] = g.ids[k0 + k];
  }
  if (nraw > A6_RAW) { if (lane == 0) atomicAdd(A.overflow, 1); nraw = 0; }
  __syncwarp();
  // ascending, duplicates adjacent (std::set order): rank by (id, position)
  for (int e = lane; e < nraw; e += 32) {
    const uint32_t id = raw[e];
    int pos = 0; bool d = false;
    for (int x = 0; x < nraw; x++) {
      const uint32_t ix = raw[x];
      pos += (ix < id) || (ix == id && x < e);
      d = d || (ix == id && x < e);
    }
    sorted[pos] = id; dup[pos] = d ? 1 : 0;
  }
  __syncwarp();
  int nu = 0;
  for (int base = 0; base < nraw; base += 32) {
    const int e = base + lane;
    const bool first = e < nraw && !dup[e];
    const unsigned m = __ballot_sync(0xffffffffu, first);
    const int pos = nu + __popc(m & ((1u << lane) - 1u));
    if (first && pos < A6_IDS) uniq[pos] = sorted[e];
    nu += __popc(m);
  }
  if (nu > A6_IDS) { if (lane == 0) atomicAdd(A.overflow, 1); nu = A6_IDS; }
  __syncwarp();
  // distance of the observation to every nearby polyline
  const float start_dsq = S.prm.detection_starting_radius * S.prm.detection_starting_radius;
  const float corr_d = S.prm.detection_starting_radius * S.prm.detection_mult;
  const float corr_dsq = corr_d * corr_d;
  int ncand = 0, nseed = 0;
  A6Rec* out = A.recs + (size_t)ol * A6_IDS;
  for (int base = 0; base < nu; base += 32) {
    const int u = base + lane;
    int cls = 0;
    if (u < nu) {
      const uint32_t id = uniq[u];
      Pl pl = get_pl(S, view, id);
      uint32_t cs; float2 proj;
      const float dsq = pl_distancesq(pl, sp, cs, proj);
      cls = dsq <= start_dsq ? 2 : (dsq <= corr_dsq ? 1 : 0);
      A6Rec r; r.id = id; r.seg = cs; r.px = proj.x; r.py = proj.y; r.cls = cls;
      out[u] = r;
    }
    ncand += __popc(__ballot_sync(0xffffffffu, cls >= 1));
    nseed += __popc(__ballot_sync(0xffffffffu, cls == 2));
  }
  if (lane == 0) {
    A.n_ids[ol] = nu; A.n_cand[ol] = ncand; A.n_seed[ol] = nseed; A.is_last[ol] = last ? 1 : 0;
    if (last) A.row_cnt[(size_t)(rp - A.tb) * S.V + view] = ncand;   // the row of (track, view) is the LAST observation's list
  }
}

__global__ void __launch_bounds__(A6_THREADS) a6_fill_kernel(const __grid_constant__ DevScene S, const __grid_constant__ A6Args A) {
  const int lane = threadIdx.x & 31;
  const int64_t ol = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t o = A.o_begin + ol;
  if (o >= A.o_end) return;
  const int rp = A.obs_track[o];
  const int view = S.track_view[o];
  const int nu = A.n_ids[ol];
  const bool last = A.is_last[ol] != 0;
  const A6Rec* in = A.recs + (size_t)ol * A.ids_cap;
  // the seed's reference point: the observation the lists were built around (last one of the view)
  const int64_t t0 = S.track_off[rp], t1 = S.track_off[rp + 1];
  int64_t jl = -1;
  for (int64_t j0 = t0; j0 < t1; j0 += 32) {
    const int64_t j = j0 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, j < t1 && S.track_view[j] == view);
    if (m) jl = j0 + 31 - __clz(m);
  }
  const float2 sp = S.track_xy[jl];
  const size_t row = (size_t)(rp - A.tb) * S.V + view;
  int64_t sbase = A.seed_off[ol];
  int64_t cbase = last ? A.coff[row] : 0;
  if (last && lane == 0) A.center[row] = S.track_xy[o];
  for (int base = 0; base < nu; base += 32) {
    const int u = base + lane;
    A6Rec r; r.cls = 0;
    if (u < nu) r = in[u];
    const unsigned ms = __ballot_sync(0xffffffffu, r.cls == 2), mc = __ballot_sync(0xffffffffu, r.cls >= 1);
    const unsigned below = (1u << lane) - 1u;
    if (r.cls == 2) {
      const int64_t k = sbase + __popc(ms & below);
      const float2 c = make_float2(r.px, r.py);
      const float radius = dist2(sp, c) * S.prm.detection_mult;          // plg_edge_manager.cpp:254
      A.s_view[k] = view; A.s_set[k] = (int)(rp - A.tb); A.s_pl[k] = r.id; A.s_seg[k] = r.seg; A.s_xy[k] = c; A.s_r2[k] = radius * radius;
    }
    if (last && r.cls >= 1) A.cpl[cbase + __popc(mc & below)] = r.id;
    sbase += __popc(ms); cbase += __popc(mc);
  }
}

}  // namespace eg3d
