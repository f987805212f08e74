// eg3d_capi.cu — host side of libeg3d.so: scene upload + derived structures, kernel orchestration, ordered packing of
// the accepted points, and the extern "C" boundary declared in include/eg3d.h.  No CPU fallback: every compute entry
// point requires a CUDA device and fails with EG3D_ERR_NO_DEVICE otherwise.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <set>
#include <algorithm>
#include <memory>
#include <limits>
#include <sstream>
#include <cub/cub.cuh>
#include <dlfcn.h>
#include <chrono>
#include <nccl.h>
#include "eg3d_dev.cuh"
#include "eg3d_k1.cuh"
#include "eg3d_k3.cuh"
#include "eg3d_gn.cuh"
#include "eg3d_a6.cuh"

using namespace eg3d;

static thread_local std::string g_err;
static eg3d_status fail(eg3d_status s, const std::string& m) { g_err = m; return s; }
eg3d_status eg3d_internal_fail(eg3d_status s, const char* msg) { return fail(s, msg); }  // for the host-only translation units
#define CK(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) {                                                                                  \
      char b_[512]; snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return fail(e_ == cudaErrorMemoryAllocation ? EG3D_ERR_OOM : EG3D_ERR_CUDA, b_);                         \
    }                                                                                                         \
  } while (0)

// Device buffers come from the stream-ordered memory pool (cudaMallocAsync) of the current device; the pool's release
// threshold is raised at scene creation so that the GB-sized per-call temporaries are recycled instead of being
// returned to the driver after every call.
static thread_local cudaStream_t g_alloc_stream = nullptr;
template <typename T>
struct DBuf {  // owning device buffer
  T* p = nullptr; size_t n = 0; cudaStream_t s = nullptr;
  DBuf() {}
  DBuf(const DBuf&) = delete; DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }
  void release() { if (p) { cudaFreeAsync(p, s); p = nullptr; } }
  cudaError_t alloc(size_t count) {
    release(); n = count; s = g_alloc_stream;
    return cudaMallocAsync((void**)&p, std::max<size_t>(count, 1) * sizeof(T), s);
  }
  cudaError_t upload(const T* h, size_t count, cudaStream_t st) {
    cudaError_t e = alloc(count); if (e != cudaSuccess) return e;
    if (count) e = cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, st);
    return e;
  }
  cudaError_t upload(const std::vector<T>& h, cudaStream_t st) { return upload(h.data(), h.size(), st); }
};

// grow-only reuse of a scene-owned buffer: (re)allocates with 25 % head-room only when the request exceeds what is there
template <typename T>
static cudaError_t ensure(DBuf<T>& b, size_t count) {
  if (b.p && b.n >= count) return cudaSuccess;
  return b.alloc(count + count / 4);
}

struct HostGrid { float cell; int w, h; std::vector<int> off; std::vector<uint32_t> ids; };

// The library's stream is shared by a scene and by every result produced from it (their device buffers are freed on it):
// it is destroyed when the last of them goes, so an eg3d_points may outlive its scene.  Declared FIRST in both structs, i.e.
// destroyed after their buffers.
struct StreamHolder {
  int device = 0; cudaStream_t s = nullptr;
  ~StreamHolder() { if (s) { cudaSetDevice(device); cudaStreamSynchronize(s); cudaStreamDestroy(s); } }
};

struct eg3d_scene {
  std::shared_ptr<StreamHolder> sh;
  int device = 0; cudaStream_t stream = nullptr;
  int V = 0, width = 0, height = 0, num_sms = 0;
  eg3d_params prm;
  DBuf<float> P; DBuf<double> P64; DBuf<double> F; DBuf<double> Fp; DBuf<double> Fph; DBuf<uint8_t> Fvalid;
  DBuf<int> view_poly_off, poly_vert_off, view_seg_off, poly_seg_off;
  DBuf<float2> verts; DBuf<uint32_t> poly_start, poly_end;
  DBuf<float4> seg; DBuf<uint2> seg_id; DBuf<float4> grp_box; DBuf<uint32_t> grp_desc; DBuf<int4> chunks; DBuf<int> view_chunk_off;
  DBuf<int> g4_off, g30_off; DBuf<uint32_t> g4_ids, g30_ids;
  DBuf<float> track_xyz; DBuf<int64_t> track_off; DBuf<int32_t> track_view; DBuf<float2> track_xy;
  DBuf<int> obs_track;          // [NO] track id of every observation (pipeline-3 seeding, one warp per observation)
  DevScene dev;
  // host copies needed by host-side steps (seed sampler, refpoint seeding)
  std::vector<int> h_view_poly_off, h_poly_vert_off; std::vector<float2> h_verts; std::vector<uint32_t> h_start, h_end;
  std::vector<int> h_view_seg_off;
  HostGrid hg30;
  std::vector<int64_t> h_track_off; std::vector<int32_t> h_track_view; std::vector<float2> h_track_xy;
  int max_view_segs = 0;
  // K3's large per-call temporaries (per-warp scratch arenas: GBs; unordered chain output) stay with the scene, grow-only.  Going
  // through the stream-ordered pool per call made the pool re-grow by GBs whenever calls of different shapes alternate (the
  // three pipelines of the reference's driver: 100+ ms of host stalls per call inside the timed region).
  DBuf<unsigned char> k3_scratch; DBuf<float> k3_uX, k3_ux, k3_uy; DBuf<int> k3_unobs, k3_uv; DBuf<int64_t> k3_uobase; DBuf<uint32_t> k3_upl, k3_useg;
  // multi-GPU exchange (eg3d_comm_create): an NCCL communicator owned by the scene handle
  ncclComm_t comm = nullptr; int comm_rank = 0, comm_world = 1;
};

// Page-locked host staging for results: one block per result, recycled through a process-wide free list so that the
// (slow) cudaMallocHost happens once per size class instead of once per call.
#include <mutex>
#include <map>
struct PinnedPool {
  std::mutex mu; std::multimap<size_t, void*> free_blocks; std::map<void*, size_t> sizes;
  void* acquire(size_t bytes, cudaError_t& err) {
    err = cudaSuccess;
    bytes = std::max<size_t>(bytes, 256);
    {
      std::lock_guard<std::mutex> g(mu);
      auto it = free_blocks.lower_bound(bytes);
      if (it != free_blocks.end() && it->first <= bytes * 2 + (1 << 20)) { void* p = it->second; free_blocks.erase(it); return p; }
    }
    void* p = nullptr;
    size_t cap = bytes + bytes / 8;
    err = cudaMallocHost(&p, cap);
    if (err != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> g(mu);
    sizes[p] = cap;
    return p;
  }
  void release(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(mu);
    free_blocks.insert({sizes[p], p});
  }
};
static PinnedPool g_pinned;

struct eg3d_points {
  std::shared_ptr<StreamHolder> sh;
  int device = 0; cudaStream_t stream = nullptr;
  int64_t n_points = 0, n_obs = 0, n_seeds = 0;
  // device-resident, ordered by (seed, chain position)
  DBuf<float> d_xyz; DBuf<int> d_seed, d_pos; DBuf<int64_t> d_obs_off; DBuf<int> d_ov; DBuf<uint32_t> d_opl, d_oseg; DBuf<float> d_oxy;
  // host copies (filled on first eg3d_points_get), carved out of one pinned block
  bool on_host = false; void* hblock = nullptr;
  float* xyz = nullptr; int32_t *seed = nullptr, *chain_pos = nullptr; int64_t* obs_off = nullptr; int32_t* obs_view = nullptr;
  uint32_t *obs_poly = nullptr, *obs_seg = nullptr; float* obs_xy = nullptr;
  ~eg3d_points() { g_pinned.release(hblock); }
};
struct eg3d_hits { int64_t n_seeds; int V; std::vector<int64_t> off; std::vector<eg3d_hit> hits; };

// ------------------------------------------------------------------------------------------------ host helpers ----
// polyline::get_intersectedcells_2dmap_set (polyline_graph_2d.cpp:798-835) + PolyLine2DMap ctor (polyLine_2d_map.cpp:40-58)
static void build_grid(const eg3d_scene& sc, float cell, HostGrid& g) {
  g.cell = cell; g.w = (int)ceilf(sc.width / cell); g.h = (int)ceilf(sc.height / cell);
  const int V = sc.V; const size_t ncell = (size_t)g.w * g.h;
  std::vector<std::vector<uint32_t>> cells(ncell);
  g.off.assign((size_t)V * ncell + 1, 0);
  g.ids.clear();
  DevScene hs; memset(&hs, 0, sizeof hs);
  hs.V = V; hs.view_poly_off = sc.h_view_poly_off.data(); hs.poly_vert_off = sc.h_poly_vert_off.data();
  hs.verts = sc.h_verts.data(); hs.poly_start = sc.h_start.data(); hs.poly_end = sc.h_end.data();
  const float step = (float)(cell / (1.414 + 0.1));
  for (int v = 0; v < V; v++) {
    for (auto& c : cells) c.clear();
    const int npl = sc.h_view_poly_off[v + 1] - sc.h_view_poly_off[v];
    for (int id = 0; id < npl; id++) {
      Pl pl = get_pl(hs, v, (uint32_t)id);
      if (pl.n < 2) continue;
      std::set<std::pair<int, int>> cs;
      PlP cur; cur.seg = 0; cur.c = pl.pc[0];
      bool reached = false; bool first = true;
      while (true) {
        if (!first) { cur = step_by_distance(pl, cur, pl.end, step, reached); }
        first = false;
        bool on_boundary = is_multiple_of(cur.c.x, cell) || is_multiple_of(cur.c.y, cell);
        if (!on_boundary) cs.insert({(int)floor_or_upper_if_close(cur.c.x / cell), (int)floor_or_upper_if_close(cur.c.y / cell)});
        if (reached) break;
      }
      for (auto& c : cs) if (c.first >= 0 && c.first < g.w && c.second >= 0 && c.second < g.h) cells[(size_t)c.second * g.w + c.first].push_back((uint32_t)id);
    }
    for (size_t c = 0; c < ncell; c++) {
      g.ids.insert(g.ids.end(), cells[c].begin(), cells[c].end());
      g.off[(size_t)v * ncell + c + 1] = (int)g.ids.size();
    }
  }
}

// Fundamental matrices of the cameras themselves: F_ji = (-1)^(i+j) det[ rows of P_a without i ; rows of P_b without j ]
// (Hartley & Zisserman, Multiple View Geometry, eq. 17.3), so that x_b^T F x_a = 0 for the projections of any X.
static double det4(const double m[4][4]) {
  double d = 0;
  for (int c = 0; c < 4; c++) {
    double sub[3][3];
    for (int i = 1; i < 4; i++) { int cc = 0; for (int j = 0; j < 4; j++) { if (j == c) continue; sub[i - 1][cc++] = m[i][j]; } }
    double d3 = sub[0][0] * (sub[1][1] * sub[2][2] - sub[1][2] * sub[2][1]) - sub[0][1] * (sub[1][0] * sub[2][2] - sub[1][2] * sub[2][0]) +
                sub[0][2] * (sub[1][0] * sub[2][1] - sub[1][1] * sub[2][0]);
    d += ((c & 1) ? -1.0 : 1.0) * m[0][c] * d3;
  }
  return d;
}
static void camera_fundamentals(const float* cams, int V, std::vector<double>& Fp) {
  Fp.assign((size_t)V * V * 9, 0.0);
  for (int a = 0; a < V; a++)
    for (int b = 0; b < V; b++) {
      if (a == b) continue;
      double* F = &Fp[((size_t)a * V + b) * 9];
      double nrm = 0;
      for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) {
          double m[4][4]; int r = 0;
          for (int k = 0; k < 3; k++) if (k != i) { for (int c = 0; c < 4; c++) m[r][c] = (double)cams[(size_t)a * 12 + k * 4 + c]; r++; }
          for (int k = 0; k < 3; k++) if (k != j) { for (int c = 0; c < 4; c++) m[r][c] = (double)cams[(size_t)b * 12 + k * 4 + c]; r++; }
          double v = (((i + j) & 1) ? -1.0 : 1.0) * det4(m);
          F[j * 3 + i] = v; nrm += v * v;
        }
      nrm = nrm > 0 ? 1.0 / sqrt(nrm) : 1.0;
      for (int k = 0; k < 9; k++) F[k] *= nrm;
    }
}

static eg3d_status require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) return fail(EG3D_ERR_NO_DEVICE, "no usable CUDA device (libeg3d has no CPU fallback)");
  return EG3D_OK;
}

struct WallClock {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  float ms() const { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr; cudaStream_t s;
  explicit Timer(cudaStream_t st) : s(st) { cudaEventCreate(&a); cudaEventCreate(&b); }
  ~Timer() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  void start() { cudaEventRecord(a, s); }
  void stop() { cudaEventRecord(b, s); }
  float ms() { float m = 0; cudaEventSynchronize(b); cudaEventElapsedTime(&m, a, b); return m; }
};

// ------------------------------------------------------------------------------------------------ pack kernels ----
__global__ void pack_points_kernel(int n_seeds, const int* __restrict__ seed_npts, const int64_t* __restrict__ seed_pbase,
                                   const int64_t* __restrict__ pt_off, const float* __restrict__ uX, const int* __restrict__ unobs,
                                   float* __restrict__ xyz, int* __restrict__ seed_out, int* __restrict__ pos_out,
                                   int64_t* __restrict__ nobs_out, int64_t* __restrict__ src_out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seeds) return;
  int n = seed_npts[s];
  int64_t src = seed_pbase[s], dst = pt_off[s];
  for (int j = 0; j < n; j++) {
    xyz[3 * (dst + j)] = uX[3 * (src + j)]; xyz[3 * (dst + j) + 1] = uX[3 * (src + j) + 1]; xyz[3 * (dst + j) + 2] = uX[3 * (src + j) + 2];
    seed_out[dst + j] = s; pos_out[dst + j] = j; nobs_out[dst + j] = unobs[src + j]; src_out[dst + j] = src + j;
  }
}
__global__ void pack_obs_kernel(int64_t n_points, const int64_t* __restrict__ obs_off, const int64_t* __restrict__ src,
                                const int64_t* __restrict__ uobase, const int* __restrict__ uv, const uint32_t* __restrict__ upl,
                                const uint32_t* __restrict__ useg, const float* __restrict__ ux, const float* __restrict__ uy,
                                int* __restrict__ ov, uint32_t* __restrict__ opl, uint32_t* __restrict__ oseg, float* __restrict__ oxy) {
  int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (p >= n_points) return;
  int64_t d0 = obs_off[p], n = obs_off[p + 1] - d0, s0 = uobase[src[p]];
  for (int64_t k = lane; k < n; k += 32) {
    ov[d0 + k] = uv[s0 + k]; opl[d0 + k] = upl[s0 + k]; oseg[d0 + k] = useg[s0 + k];
    oxy[2 * (d0 + k)] = ux[s0 + k]; oxy[2 * (d0 + k) + 1] = uy[s0 + k];
  }
}

// ------------------------------------------------------------------------------------------------ seeds on device ----
struct DevSeeds {
  DBuf<int> view; DBuf<uint32_t> pl, seg; DBuf<float2> xy; DBuf<int> cand_set; int n = 0;
  K1Seeds k1() const { K1Seeds s; s.n = n; s.view = view.p; s.pl = pl.p; s.seg = seg.p; s.xy = xy.p; s.cand_set = cand_set.p; return s; }
};
struct DevCand { DBuf<int64_t> off; DBuf<uint32_t> pl; DBuf<float2> center; DBuf<float> seed_r2; bool filtered = false; };

static eg3d_status upload_seeds(eg3d_scene* sc, const eg3d_seeds* s, bool need_cand, DevSeeds& d) {
  d.n = (int)s->n;
  CK(d.view.upload(s->view, s->n, sc->stream));
  CK(d.pl.upload(s->polyline, s->n, sc->stream));
  CK(d.seg.upload(s->segment, s->n, sc->stream));
  CK(d.xy.upload((const float2*)s->xy, s->n, sc->stream));
  if (need_cand) CK(d.cand_set.upload(s->cand_set, s->n, sc->stream));
  return EG3D_OK;
}

// ---------------------------------------------------------------------------------------------------- K1 driver ----
static eg3d_status exclusive_scan_i64(eg3d_scene* sc, int64_t* p, size_t n) {
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, p, p, n, sc->stream);
  DBuf<unsigned char> tmp; CK(tmp.alloc(tb));
  cub::DeviceScan::ExclusiveSum(tmp.p, tb, p, p, n, sc->stream);
  return EG3D_OK;
}

template <int MODE>
static void launch_sweep(eg3d_scene* sc, const K1Seeds& ks, const K1Work& w, int64_t* counts, unsigned char* flags, const int64_t* off, eg3d_hit* hits) {
  if (w.n <= 0) return;
  dim3 grid((w.n + K1_THREADS - 1) / K1_THREADS, sc->V);
  if (!w.list && !w.sel) k1_sweep_kernel<MODE, true><<<grid, K1_THREADS, K1_SMEM_BYTES2, sc->stream>>>(sc->dev, ks, w, counts, flags, off, hits);
  else k1_sweep_kernel<MODE, false><<<grid, K1_THREADS, K1_SMEM_BYTES2, sc->stream>>>(sc->dev, ks, w, counts, flags, off, hits);
}

// Sweep form of K1 for the pairs described by `w` (n_rows result rows): count -> scan -> fill.
static eg3d_status sweep_lists(eg3d_scene* sc, const K1Seeds& ks, const K1Work& w, size_t n_rows, bool rows_sparse, DBuf<int64_t>& off, DBuf<eg3d_hit>& hits,
                               int64_t& n_hits, eg3d_timing* tm) {
  CK(off.alloc(n_rows + 1));
  if (rows_sparse) CK(cudaMemsetAsync(off.p, 0, (n_rows + 1) * sizeof(int64_t), sc->stream));   // rows no launch writes stay empty
  else CK(cudaMemsetAsync(off.p + n_rows, 0, sizeof(int64_t), sc->stream));
  Timer t1(sc->stream), t2(sc->stream), t3(sc->stream);
  t1.start(); launch_sweep<K1_COUNT>(sc, ks, w, off.p, nullptr, nullptr, nullptr); t1.stop();
  CK(cudaGetLastError());
  t2.start();
  eg3d_status st = exclusive_scan_i64(sc, off.p, n_rows + 1); if (st != EG3D_OK) return st;
  t2.stop();
  CK(cudaMemcpyAsync(&n_hits, off.p + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, sc->stream));
  CK(cudaStreamSynchronize(sc->stream));
  CK(hits.alloc((size_t)n_hits));
  t3.start(); launch_sweep<K1_FILL>(sc, ks, w, nullptr, nullptr, off.p, hits.p); t3.stop();
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(sc->stream));
  if (tm) { tm->k1_count_ms += t1.ms(); tm->scan_ms += t2.ms(); tm->k1_fill_ms += t3.ms(); tm->n_hits += n_hits; tm->kernel_launches += 4; }
  return EG3D_OK;
}

// K1 for every (seed, view) pair: count -> scan -> fill.  Leaves off (n*V+1) and hits on the device.
static eg3d_status run_k1(eg3d_scene* sc, const DevSeeds& ds, const DevCand* dc, DBuf<int64_t>& off, DBuf<eg3d_hit>& hits,
                          int64_t& n_hits, eg3d_timing* tm) {
  const int V = sc->V; const int64_t nsv = (int64_t)ds.n * V;
  K1Seeds ks = ds.k1();
  if (!dc) {
    K1Work w; w.list = nullptr; w.n = ds.n; w.sel = nullptr;
    return sweep_lists(sc, ks, w, (size_t)nsv, false, off, hits, n_hits, tm);
  }
  CK(off.alloc(nsv + 1));
  CK(cudaMemsetAsync(off.p + nsv, 0, sizeof(int64_t), sc->stream));
  Timer t1(sc->stream), t2(sc->stream), t3(sc->stream);
  K1Cand kc; kc.off = dc->off.p; kc.pl = dc->pl.p;
  kc.center = dc->filtered ? dc->center.p : nullptr; kc.seed_r2 = dc->filtered ? dc->seed_r2.p : nullptr;
  const int cblocks = (int)((nsv + 255) / 256);
  t1.start();
  if (ds.n > 0) k1_cand_kernel<false><<<cblocks, 256, 0, sc->stream>>>(sc->dev, ks, kc, off.p, nullptr, nullptr);
  t1.stop();
  CK(cudaGetLastError());
  t2.start();
  eg3d_status st = exclusive_scan_i64(sc, off.p, (size_t)nsv + 1); if (st != EG3D_OK) return st;
  t2.stop();
  CK(cudaMemcpyAsync(&n_hits, off.p + nsv, sizeof(int64_t), cudaMemcpyDeviceToHost, sc->stream));
  CK(cudaStreamSynchronize(sc->stream));
  CK(hits.alloc((size_t)n_hits));
  t3.start();
  if (ds.n > 0) k1_cand_kernel<true><<<cblocks, 256, 0, sc->stream>>>(sc->dev, ks, kc, nullptr, off.p, hits.p);
  t3.stop();
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(sc->stream));
  if (tm) {
    tm->k1_count_ms += t1.ms(); tm->scan_ms += t2.ms(); tm->k1_fill_ms += t3.ms();
    tm->n_hits += n_hits; tm->kernel_launches += 4;
  }
  return EG3D_OK;
}

// The hit lists K3 reads.  Full form (candidate sets / refpoints): one CSR over every (seed, view) pair.  Lazy form
// (all-segment sweep): the reference materialises every list (polyline_matching.cpp:45-73) but its per-seed code only
// ever READS the three selected views unless the seed is accepted, so the sweep is run on demand — an any-hit pass for
// the view selection, the three selected views of every seed for phase A, and every view of the accepted seeds for
// phase B.  Same lists, same order, same results; ~5x fewer segment tests and hit bytes on BASELINE configs[1].
struct HitLists {
  bool lazy = false;
  DBuf<unsigned char> flags;                       // lazy: [n*V] list is non-empty
  DBuf<int> sel;                                   // [n][3]
  DBuf<int64_t> off_a; DBuf<eg3d_hit> hits_a;      // full: [n*V+1]; lazy: [3n+1]
  DBuf<int64_t> off_b; DBuf<eg3d_hit> hits_b;      // lazy only: [n_acc*V+1], rebuilt after phase A
  DBuf<int> acc_seed;                              // [n] accepted seed by phase-A rank
};

static eg3d_status select_views(eg3d_scene* sc, const DevSeeds& ds, HitLists& H, eg3d_timing* tm) {
  CK(H.sel.alloc(3 * (size_t)std::max(ds.n, 1)));
  CK(H.acc_seed.alloc((size_t)std::max(ds.n, 1)));
  if (ds.n == 0) return EG3D_OK;
  Timer t(sc->stream);
  t.start();
  k3_select_views_kernel<<<(unsigned)(((size_t)ds.n * 32 + 127) / 128), 128, 0, sc->stream>>>(ds.n, sc->V, ds.view.p, H.lazy ? H.flags.p : nullptr,
                                                                                          H.lazy ? nullptr : H.off_a.p, H.sel.p);
  t.stop();
  CK(cudaGetLastError());
  if (tm) { tm->scan_ms += t.ms(); tm->kernel_launches += 1; }
  return EG3D_OK;
}

// Everything phase A needs.  dc != null: candidate form (full CSR).
static eg3d_status prepare_hits_a(eg3d_scene* sc, const DevSeeds& ds, const DevCand* dc, HitLists& H, eg3d_timing* tm) {
  int64_t nh = 0;
  H.lazy = (dc == nullptr) && !getenv("EG3D_K1_FULL");
  if (!H.lazy) {
    eg3d_status st = run_k1(sc, ds, dc, H.off_a, H.hits_a, nh, tm); if (st != EG3D_OK) return st;
    return select_views(sc, ds, H, tm);
  }
  const int V = sc->V;
  K1Seeds ks = ds.k1();
  CK(H.flags.alloc((size_t)std::max<int64_t>((int64_t)ds.n * V, 1)));
  Timer t(sc->stream);
  K1Work w; w.list = nullptr; w.n = ds.n; w.sel = nullptr;
  t.start(); launch_sweep<K1_ANY>(sc, ks, w, nullptr, H.flags.p, nullptr, nullptr); t.stop();
  CK(cudaGetLastError());
  if (tm) { tm->k1_any_ms += t.ms(); tm->kernel_launches += 1; }
  eg3d_status st = select_views(sc, ds, H, tm); if (st != EG3D_OK) return st;
  w.sel = H.sel.p;
  return sweep_lists(sc, ks, w, 3 * (size_t)ds.n, true, H.off_a, H.hits_a, nh, tm);
}

// lazy form: every view of the accepted seeds (rank order of phase A)
static eg3d_status prepare_hits_b(eg3d_scene* sc, const DevSeeds& ds, HitLists& H, int64_t n_acc, eg3d_timing* tm) {
  K1Seeds ks = ds.k1();
  K1Work w; w.list = H.acc_seed.p; w.n = (int)n_acc; w.sel = nullptr;
  int64_t nh = 0;
  return sweep_lists(sc, ks, w, (size_t)n_acc * sc->V, false, H.off_b, H.hits_b, nh, tm);
}

// K3 + ordered packing + D2H into an eg3d_points
static eg3d_status run_k3_once(eg3d_scene* sc, const DevSeeds& ds, HitLists& H, eg3d_points* out, eg3d_timing* tm,
                               int64_t cap_scale, int seed_cap_mult, bool& out_overflow, bool& seed_cap_overflow) {
  out_overflow = false; seed_cap_overflow = false;
  const int V = sc->V; const int n = ds.n;
  out->sh = sc->sh; out->device = sc->device; out->stream = sc->stream; out->n_seeds = n;
  if (n == 0) { CK(out->d_obs_off.alloc(1)); CK(cudaMemsetAsync(out->d_obs_off.p, 0, sizeof(int64_t), sc->stream)); CK(cudaStreamSynchronize(sc->stream)); return EG3D_OK; }
  const int capf = sc->prm.max_follow_points * seed_cap_mult, capc = sc->prm.max_chain_points * seed_cap_mult, oc = V + 16;
  const size_t spw = k3_scratch_bytes(V, capf, capc, oc);
  int blocks_per_sm = 16;  // upper bound; residency is set by the kernel's register budget
  int nblocks = sc->num_sms * blocks_per_sm;
  const int warps_per_block = K3_THREADS / 32;
  int max_useful = (n + warps_per_block - 1) / warps_per_block;
  if (nblocks > max_useful) nblocks = max_useful;
  // phase B: EG3D_K3B_MIN_BLOCKS CTAs of K3B_THREADS per SM (one CTA per SM in the lock-step form)
  const int warps_per_block_b = K3B_THREADS / 32;
  int blocks_per_sm_b = EG3D_K3B_MIN_BLOCKS;
  if (const char* e = getenv("EG3D_K3B_BLOCKS_PER_SM")) blocks_per_sm_b = std::max(1, std::min(EG3D_K3B_MIN_BLOCKS, atoi(e)));   // A/B runs
  int nblocks_b = sc->num_sms * blocks_per_sm_b;
  nblocks_b = std::max(1, std::min(nblocks_b, (n + warps_per_block_b - 1) / warps_per_block_b));
  const int batch_b = EG3D_K3B_BATCH > 0 ? EG3D_K3B_BATCH : 1;     // arenas per phase-B warp
  if (EG3D_K3B_BATCH > 0) nblocks_b = std::max(1, std::min(sc->num_sms, (n + warps_per_block_b * batch_b - 1) / (warps_per_block_b * batch_b)));
  const size_t nwarps = std::max((size_t)nblocks * warps_per_block, (size_t)nblocks_b * warps_per_block_b * batch_b);
  CK(ensure(sc->k3_scratch, nwarps * spw));
  DBuf<unsigned char>& scratch = sc->k3_scratch;
  DBuf<int> counter; CK(counter.alloc(1)); CK(cudaMemsetAsync(counter.p, 0, sizeof(int), sc->stream));
  DBuf<unsigned long long> oc4; CK(oc4.alloc(4)); CK(cudaMemsetAsync(oc4.p, 0, 4 * sizeof(unsigned long long), sc->stream));
  // unordered output capacity: generous typical-case bound; exceeding it is reported, never silently truncated
  const bool tiny = getenv("EG3D_TEST_TINY_CAPS") != nullptr;   // test knob: start from absurdly small output buffers to exercise the retry
  int64_t pt_cap = std::min<int64_t>((int64_t)n * capc, (tiny ? (int64_t)64 : std::max<int64_t>((int64_t)n * 24, 1 << 16)) * cap_scale);
  int64_t ob_cap = pt_cap * std::min<int64_t>(oc, 64 + V / 2);
  DBuf<float>& uX = sc->k3_uX; DBuf<int>& unobs = sc->k3_unobs; DBuf<int64_t>& uobase = sc->k3_uobase; DBuf<int>& uv = sc->k3_uv;
  DBuf<uint32_t>&upl = sc->k3_upl, &useg = sc->k3_useg; DBuf<float>&ux = sc->k3_ux, &uy = sc->k3_uy;
  DBuf<int> snp; DBuf<int64_t> spb, sno;
  CK(ensure(uX, 3 * pt_cap)); CK(ensure(unobs, pt_cap)); CK(ensure(uobase, pt_cap));
  CK(ensure(uv, ob_cap)); CK(ensure(upl, ob_cap)); CK(ensure(useg, ob_cap)); CK(ensure(ux, ob_cap)); CK(ensure(uy, ob_cap));
  CK(snp.alloc(n)); CK(spb.alloc(n)); CK(sno.alloc(n));
  K3Args a; memset(&a, 0, sizeof a);
  a.n_seeds = n; a.seed_view = ds.view.p; a.seed_pl = ds.pl.p; a.seed_seg = ds.seg.p; a.seed_xy = ds.xy.p;
  a.hit_off = H.off_a.p; a.hits = H.hits_a.p; a.a_compact = H.lazy ? 1 : 0;
  a.hit_off_b = H.off_a.p; a.hits_b = H.hits_a.p; a.b_compact = 0;     // lazy form: replaced after phase A
  a.sel = H.sel.p; a.acc_seed = H.acc_seed.p;
  a.capf = capf; a.capc = capc; a.oc = oc;
  // shared first Gauss-Newton iteration (k3b_expand_kernel<true>): on BASELINE configs[1] (200 views) the extra code costs more
  // instruction fetches than the saved observation passes (k3b 177 -> 193 ms); with 1000 views (configs[2]) the observation loop
  // dominates and it pays (1401 -> 1261 ms per 31 250-seed shard).  EG3D_GN_CACHE=0|1 overrides.
  a.gn_cache = V >= 400 ? 1 : 0;
  if (const char* e = getenv("EG3D_GN_CACHE")) a.gn_cache = atoi(e) ? 1 : 0;
  a.scratch = scratch.p; a.scratch_per_warp = spw; a.work_counter = counter.p;
  a.pt_cap = pt_cap; a.ob_cap = ob_cap; a.out_counters = oc4.p;
  a.o_X = uX.p; a.o_nobs = unobs.p; a.o_obase = uobase.p;
  a.ob_view = uv.p; a.ob_pl = upl.p; a.ob_seg = useg.p; a.ob_x = ux.p; a.ob_y = uy.p;
  a.seed_npts = snp.p; a.seed_pbase = spb.p; a.seed_nobs = sno.p;
  DBuf<unsigned long long> prof;
  const bool do_prof = getenv("EG3D_K3_PROF") != nullptr;
  DBuf<unsigned long long> prof_seed;
  if (do_prof) {
    CK(prof.alloc(48)); CK(cudaMemsetAsync(prof.p, 0, 48 * sizeof(unsigned long long), sc->stream)); a.prof = prof.p;
    CK(prof_seed.alloc(3 * (size_t)n)); CK(cudaMemsetAsync(prof_seed.p, 0, 3 * (size_t)n * sizeof(unsigned long long), sc->stream)); a.prof_seed = prof_seed.p;
  }
  // phase A -> phase B hand-over buffers
  DBuf<PaRec> pa_recs; DBuf<Pt3> pa_pool; DBuf<unsigned long long> pa_cnt; DBuf<int> counter_b;
  const long long pool_cap = std::min<long long>((long long)n * 2 * capf, (tiny ? 32ll : std::max<long long>((long long)n * 12, 1 << 14)) * cap_scale);
  CK(pa_recs.alloc(n)); CK(pa_pool.alloc(pool_cap)); CK(pa_cnt.alloc(2)); CK(counter_b.alloc(1));
  CK(cudaMemsetAsync(pa_cnt.p, 0, 2 * sizeof(unsigned long long), sc->stream));
  CK(cudaMemsetAsync(counter_b.p, 0, sizeof(int), sc->stream));
  CK(cudaMemsetAsync(snp.p, 0, (size_t)n * sizeof(int), sc->stream));
  a.pa_recs = pa_recs.p; a.pa_pool = pa_pool.p; a.pa_pool_cap = pool_cap; a.pa_counters = pa_cnt.p;
  Timer t3(sc->stream), t3b(sc->stream), tp(sc->stream);
  t3.start();
  k3a_hypothesis_kernel<<<nblocks, K3_THREADS, 0, sc->stream>>>(sc->dev, a);
  t3.stop();
  CK(cudaGetLastError());
  unsigned long long pac[2] = {0, 0};
  CK(cudaMemcpyAsync(pac, pa_cnt.p, sizeof pac, cudaMemcpyDeviceToHost, sc->stream));
  CK(cudaStreamSynchronize(sc->stream));
  K3Args b = a; b.work_counter = counter_b.p;
  if (H.lazy) {
    eg3d_status st = prepare_hits_b(sc, ds, H, (int64_t)pac[0], tm); if (st != EG3D_OK) return st;
    b.hit_off_b = H.off_b.p; b.hits_b = H.hits_b.p; b.b_compact = 1;
  }
  // phase-B work order: longest chains first
  DBuf<unsigned> okeys, okeys2; DBuf<int> ovals, oorder; DBuf<unsigned char> otmp;
  if (!getenv("EG3D_K3_NO_ORDER")) {
    CK(okeys.alloc(n)); CK(okeys2.alloc(n)); CK(ovals.alloc(n)); CK(oorder.alloc(n));
    k3_order_keys_kernel<<<(n + 255) / 256, 256, 0, sc->stream>>>(n, pa_cnt.p, pa_recs.p, okeys.p, ovals.p);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, okeys.p, okeys2.p, ovals.p, oorder.p, n, 0, 32, sc->stream);
    CK(otmp.alloc(tb));
    cub::DeviceRadixSort::SortPairs(otmp.p, tb, okeys.p, okeys2.p, ovals.p, oorder.p, n, 0, 32, sc->stream);
    b.pa_order = oorder.p;
    if (tm) tm->kernel_launches += 1;
  }
  t3b.start();
  if (b.gn_cache) k3b_expand_kernel<true><<<nblocks_b, K3B_THREADS, 0, sc->stream>>>(sc->dev, b);
  else k3b_expand_kernel<false><<<nblocks_b, K3B_THREADS, 0, sc->stream>>>(sc->dev, b);
  t3b.stop();
  unsigned long long cnt[4];
  CK(cudaMemcpyAsync(cnt, oc4.p, sizeof cnt, cudaMemcpyDeviceToHost, sc->stream));
  CK(cudaStreamSynchronize(sc->stream));
  if (tm) { tm->k3a_ms += t3.ms(); tm->k3b_ms += t3b.ms(); tm->k3_ms += t3.ms() + t3b.ms(); tm->n_accepted_seeds += (int64_t)pac[0]; tm->kernel_launches += 2; tm->n_capacity_overflows += (int)(cnt[2] + cnt[3]); }
  if (do_prof) {
    unsigned long long pr[48]; CK(cudaMemcpy(pr, prof.p, sizeof pr, cudaMemcpyDeviceToHost));
    const char* nm[32] = {"A.scan+prune", "A.est3", "A.plg_compatible", "B.epc_prune", "B.epc_gn", "B.add_view_finish(epc)", "B.main_loop", "seed_total", "#est3_lanes", "#est3_rounds", "#plg_compat", "#epc_solved",
                          "#epc_gn_calls", "#main_first_gn", "#avf_calls(main)", "#avf_ok(main)", "#walk_solve_calls", "#walk_solve_problems", "#main_points_visited", "#main_grid_unique_ok", "", "#follow_big", "#est_slot", "#combos_slot",
                          "#views_run", "#epc_matched", "#sum_len_at_view", "#walk_geo_calls", "#walk_geo_nwalk", "#epc_gn_accepted", "#epc_avf_attempts", "#epc_attempt_pos_sum"};
    for (int k = 0; k < 32; k++) if (nm[k][0]) fprintf(stderr, "[k3prof] %-26s %14llu%s\n", nm[k], pr[k], k < 8 ? " warp-cycles" : "");
    if (pr[45]) fprintf(stderr, "[k3prof] phase B warps: %llu, mean busy %.2f ms, last one done after %.2f ms (tail = %.1f %% of the kernel)\n", pr[45], pr[43] / (double)pr[45] * 1e-6, pr[44] * 1e-6,
                        100.0 * (1.0 - (pr[43] / (double)pr[45]) / (double)pr[44]));
    fprintf(stderr, "[k3prof] max seed cycles %llu (%.1f ms at 1.9 GHz); seeds > 50M cycles: %llu; > 200M cycles: %llu\n", pr[40], pr[40] / 1.9e6, pr[41], pr[42]);
#ifdef EG3D_K3_PROFILE
    {
      unsigned long long g[8]; CK(cudaMemcpyFromSymbol(g, g_gnprof, sizeof g));
      fprintf(stderr, "[k3prof] gn_group: calls %llu, iterations %llu, obs-loop passes %llu, running-lane-iterations %llu (cumulative over K3a+K3b of this process)\n", g[0], g[1], g[2], g[3]);
    }
#endif
    const char* dump = getenv("EG3D_K3_PROF");
    if (dump && (dump[0] == '/' || strchr(dump, '.'))) {   // EG3D_K3_PROF=<file>: per-seed (cycles, initial length, final length)
      std::vector<unsigned long long> ps(3 * (size_t)n); CK(cudaMemcpy(ps.data(), prof_seed.p, ps.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      if (FILE* f = fopen(dump, "w")) {
        for (int i = 0; i < n; i++) if (ps[3 * i]) fprintf(f, "%d %llu %llu %llu %llu\n", i, ps[3 * i], ps[3 * i + 1] & 0xffffffffull, ps[3 * i + 1] >> 32, ps[3 * i + 2]);
        fclose(f);
      }
    }
  }
  if (cnt[3]) { out_overflow = true; return EG3D_OK; }   // internal output bound exceeded: the caller retries with larger buffers
  if (cnt[2]) { seed_cap_overflow = true; return EG3D_OK; }   // a per-seed capacity was too small for some seed: the caller re-runs with doubled capacities
  // ordered packing
  tp.start();
  DBuf<int64_t> pt_off; CK(pt_off.alloc(n + 1));
  {
    // int -> int64 exclusive scan of seed_npts
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, snp.p, pt_off.p, n, sc->stream);
    DBuf<unsigned char> tmp; CK(tmp.alloc(tb));
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, snp.p, pt_off.p, n, sc->stream);
    CK(cudaStreamSynchronize(sc->stream));
  }
  const int64_t npts = (int64_t)cnt[0], nobs = (int64_t)cnt[1];
  out->device = sc->device; out->stream = sc->stream; out->n_points = npts; out->n_obs = nobs;
  DBuf<int64_t> src;
  CK(out->d_xyz.alloc(3 * npts)); CK(out->d_seed.alloc(npts)); CK(out->d_pos.alloc(npts)); CK(out->d_obs_off.alloc(npts + 1)); CK(src.alloc(npts));
  CK(cudaMemsetAsync(out->d_obs_off.p + npts, 0, sizeof(int64_t), sc->stream));
  pack_points_kernel<<<(n + 127) / 128, 128, 0, sc->stream>>>(n, snp.p, spb.p, pt_off.p, uX.p, unobs.p, out->d_xyz.p, out->d_seed.p, out->d_pos.p,
                                                             out->d_obs_off.p, src.p);
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, out->d_obs_off.p, out->d_obs_off.p, npts + 1, sc->stream);
    DBuf<unsigned char> tmp; CK(tmp.alloc(tb));
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, out->d_obs_off.p, out->d_obs_off.p, npts + 1, sc->stream);
    CK(cudaStreamSynchronize(sc->stream));
  }
  CK(out->d_ov.alloc(nobs)); CK(out->d_opl.alloc(nobs)); CK(out->d_oseg.alloc(nobs)); CK(out->d_oxy.alloc(2 * nobs));
  if (npts > 0) {
    int64_t threads = npts * 32;
    pack_obs_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, sc->stream>>>(npts, out->d_obs_off.p, src.p, uobase.p, uv.p, upl.p, useg.p, ux.p, uy.p,
                                                                            out->d_ov.p, out->d_opl.p, out->d_oseg.p, out->d_oxy.p);
  }
  tp.stop();
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(sc->stream));
  if (tm) { tm->pack_ms += tp.ms(); tm->kernel_launches += 4; tm->n_points += npts; tm->n_obs += nobs; }
  return EG3D_OK;
}

// Output buffers are sized for the typical case; if a batch produces more, K3 is simply run again with larger buffers
// (nothing is truncated and no result of the short run is used).
static eg3d_status run_k3(eg3d_scene* sc, const DevSeeds& ds, HitLists& H, eg3d_points* out, eg3d_timing* tm) {
  int64_t scale = 1; int mult = 1;
  for (int attempt = 0; attempt < 12; attempt++) {
    eg3d_timing t0; if (tm) t0 = *tm;
    bool ovf = false, seed_ovf = false;
    eg3d_status st = run_k3_once(sc, ds, H, out, tm, scale, mult, ovf, seed_ovf);
    if (st != EG3D_OK || (!ovf && !seed_ovf)) return st;
    // nothing of the short run is used; the time it took stays in the timing (it was spent)
    if (tm) { const eg3d_timing t1 = *tm; *tm = t0; tm->k3_ms = t1.k3_ms; tm->k3a_ms = t1.k3a_ms; tm->k3b_ms = t1.k3b_ms; tm->kernel_launches = t1.kernel_launches;
              tm->k1_count_ms = t1.k1_count_ms; tm->k1_fill_ms = t1.k1_fill_ms; tm->scan_ms = t1.scan_ms; tm->n_capacity_retries = t1.n_capacity_retries + (seed_ovf ? 1 : 0); }
    if (seed_ovf) { if (mult >= 64) break; mult *= 2; }     // per-seed chain / follow capacities: real polyline graphs give chains well above the synthetic rigs' (dtu006: 140 points)
    else { if (scale >= 4096) break; scale *= 4; }          // unordered output buffers
  }
  return fail(EG3D_ERR_CAPACITY, "accepted-point output exceeds 4096x the typical bound, or a seed needs more than 64x the per-seed capacities; split the seed batch");
}

// lazily bring a result to (pinned) host memory
static eg3d_status points_to_host(eg3d_points* p) {
  if (p->on_host) return EG3D_OK;
  CK(cudaSetDevice(p->device));
  const int64_t npts = p->n_points, nobs = p->n_obs;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_xyz = al(3 * npts * 4), b_seed = al(npts * 4), b_pos = al(npts * 4), b_off = al((npts + 1) * 8), b_ov = al(nobs * 4),
               b_opl = al(nobs * 4), b_oseg = al(nobs * 4), b_oxy = al(2 * nobs * 4);
  cudaError_t e;
  char* base = (char*)g_pinned.acquire(b_xyz + b_seed + b_pos + b_off + b_ov + b_opl + b_oseg + b_oxy, e);
  if (!base) return fail(EG3D_ERR_OOM, std::string("cudaMallocHost failed: ") + cudaGetErrorString(e));
  p->hblock = base;
  p->xyz = (float*)base; base += b_xyz; p->seed = (int32_t*)base; base += b_seed; p->chain_pos = (int32_t*)base; base += b_pos;
  p->obs_off = (int64_t*)base; base += b_off; p->obs_view = (int32_t*)base; base += b_ov; p->obs_poly = (uint32_t*)base; base += b_opl;
  p->obs_seg = (uint32_t*)base; base += b_oseg; p->obs_xy = (float*)base;
  p->obs_off[0] = 0;
  if (npts > 0) {
    cudaStream_t s = p->stream;
    CK(cudaMemcpyAsync(p->xyz, p->d_xyz.p, 3 * npts * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(p->seed, p->d_seed.p, npts * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(p->chain_pos, p->d_pos.p, npts * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(p->obs_off, p->d_obs_off.p, (npts + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(p->obs_view, p->d_ov.p, nobs * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(p->obs_poly, p->d_opl.p, nobs * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(p->obs_seg, p->d_oseg.p, nobs * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(p->obs_xy, p->d_oxy.p, 2 * nobs * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  p->on_host = true;
  return EG3D_OK;
}

static void k1_accounting(const eg3d_scene* sc, const eg3d_seeds* seeds, const eg3d_candidates* cands, int64_t n_hits, eg3d_timing* tm) {
  // per (seed, target view) 16 B x segments swept + 72 B (F) + 8 B (seed) + 16 B x hits written
  int64_t tests = 0;
  const int V = sc->V;
  if (!cands) {
    int64_t total = sc->h_view_seg_off[V];
    for (int64_t i = 0; i < seeds->n; i++) tests += total - (sc->h_view_seg_off[seeds->view[i] + 1] - sc->h_view_seg_off[seeds->view[i]]);
    tm->k1_algorithmic_bytes += 16 * tests + (72 + 8) * seeds->n * (int64_t)(V - 1) + 16 * n_hits;
  } else {
    // per (set, view): candidate polylines and their segments, once; then per seed the sum over the other views
    const size_t nsv = (size_t)cands->n_sets * V;
    std::vector<int64_t> segs(nsv, 0), ncs(nsv, 0), set_segs((size_t)cands->n_sets, 0), set_nc((size_t)cands->n_sets, 0);
    for (size_t r = 0; r < nsv; r++) {
      const int v = (int)(r % V);
      for (int64_t k = cands->off[r]; k < cands->off[r + 1]; k++) {
        const int g = sc->h_view_poly_off[v] + (int)cands->polyline[k];
        const int nv = sc->h_poly_vert_off[g + 1] - sc->h_poly_vert_off[g];
        segs[r] += nv > 1 ? nv - 1 : 0; ncs[r]++;
      }
      set_segs[r / V] += segs[r]; set_nc[r / V] += ncs[r];
    }
    int64_t ncand = 0;
    for (int64_t i = 0; i < seeds->n; i++) {
      const size_t set = (size_t)seeds->cand_set[i], own = set * V + seeds->view[i];
      tests += set_segs[set] - segs[own]; ncand += set_nc[set] - ncs[own];
    }
    tm->k1_algorithmic_bytes += 16 * tests + 4 * ncand + (72 + 8) * seeds->n * (int64_t)(V - 1) + 16 * n_hits;
  }
  tm->n_segment_tests += tests;
}

// ------------------------------------------------------------------------------------------- multi-GPU exchange ----
// NCCL is resolved at run time (dlopen by soname): a host process that already carries an NCCL — torch's bundled one under
// torch.distributed, for instance — shares it, a plain C++ host gets the system library, and libeg3d.so has no link-time
// dependency on either.
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
static NcclApi& nccl_api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) { a.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (a.h) break; }
    if (!a.h) return;
    auto sym = [&](const char* n) { return dlsym(a.h, n); };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId"); a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy"); a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.Broadcast = (decltype(a.Broadcast))sym("ncclBroadcast"); a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd"); a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.Broadcast && a.GroupStart && a.GroupEnd && a.GetErrorString;
  });
  return a;
}
#define NCK(call)                                                                                              \
  do {                                                                                                         \
    ncclResult_t r_ = (call);                                                                                  \
    if (r_ != ncclSuccess) {                                                                                   \
      char b_[512]; snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r_), __FILE__, __LINE__); \
      return fail(EG3D_ERR_CUDA, b_);                                                                          \
    }                                                                                                          \
  } while (0)

// One rank's records as ONE byte-packed buffer (no padding to a common size, no per-field collectives):
//   [xyz f32 x3 | n][key i64 | n][obs_off i64 | n+1][obs_view i32 | m][obs_poly u32 | m][obs_seg u32 | m][obs_xy f32 x2 | m]
// key = (global seed ordinal << 20) | chain position: the canonical order of the merged result.
struct PackLayout { int64_t n, m; size_t o_xyz, o_key, o_off, o_ov, o_opl, o_oseg, o_oxy, bytes; };
static PackLayout pack_layout(int64_t n, int64_t m) {
  PackLayout L; L.n = n; L.m = m;
  auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
  size_t o = 0;
  L.o_xyz = o; o = al(o + 12 * (size_t)n); L.o_key = o; o = al(o + 8 * (size_t)n); L.o_off = o; o = al(o + 8 * (size_t)(n + 1));
  L.o_ov = o; o = al(o + 4 * (size_t)m); L.o_opl = o; o = al(o + 4 * (size_t)m); L.o_oseg = o; o = al(o + 4 * (size_t)m); L.o_oxy = o; o = al(o + 8 * (size_t)m);
  L.bytes = o;
  return L;
}
__global__ void xchg_pack_points_kernel(int64_t n, const float* __restrict__ xyz, const int* __restrict__ seed, const int* __restrict__ pos,
                                        const int64_t* __restrict__ obs_off, const int64_t* __restrict__ seed_global, int64_t seed_base,
                                        float* __restrict__ o_xyz, int64_t* __restrict__ o_key, int64_t* __restrict__ o_off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  o_off[i] = obs_off[i];
  if (i == n) return;
  o_xyz[3 * i] = xyz[3 * i]; o_xyz[3 * i + 1] = xyz[3 * i + 1]; o_xyz[3 * i + 2] = xyz[3 * i + 2];
  const int64_t g = seed_global ? seed_global[seed[i]] : seed_base + seed[i];
  o_key[i] = (g << 20) | (int64_t)pos[i];
}
struct XchgRanks {   // per-rank views of the received buffers (<= 64 ranks)
  int world; int64_t pbase[65], obase[65];
  const float* xyz[64]; const int64_t* key[64]; const int64_t* off[64]; const int* ov[64]; const uint32_t* opl[64]; const uint32_t* oseg[64]; const float* oxy[64];
};
__device__ __forceinline__ int xchg_rank_of(const XchgRanks& R, int64_t g) { int r = 0; while (r + 1 < R.world && g >= R.pbase[r + 1]) r++; return r; }
__global__ void xchg_keys_kernel(const __grid_constant__ XchgRanks R, int64_t N, unsigned long long* __restrict__ keys, int64_t* __restrict__ idx) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  const int r = xchg_rank_of(R, g);
  keys[g] = (unsigned long long)R.key[r][g - R.pbase[r]]; idx[g] = g;
}
__global__ void xchg_counts_kernel(const __grid_constant__ XchgRanks R, int64_t N, const int64_t* __restrict__ idx, int64_t* __restrict__ nobs) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > N) return;
  if (j == N) { nobs[j] = 0; return; }
  const int64_t g = idx[j]; const int r = xchg_rank_of(R, g); const int64_t l = g - R.pbase[r];
  nobs[j] = R.off[r][l + 1] - R.off[r][l];
}
__global__ void xchg_copy_kernel(const __grid_constant__ XchgRanks R, int64_t N, const int64_t* __restrict__ idx, const unsigned long long* __restrict__ keys,
                                 const int64_t* __restrict__ obs_off, float* __restrict__ xyz, int* __restrict__ seed, int* __restrict__ pos,
                                 int* __restrict__ ov, uint32_t* __restrict__ opl, uint32_t* __restrict__ oseg, float* __restrict__ oxy) {
  const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (j >= N) return;
  const int64_t g = idx[j]; const int r = xchg_rank_of(R, g); const int64_t l = g - R.pbase[r];
  if (lane == 0) {
    xyz[3 * j] = R.xyz[r][3 * l]; xyz[3 * j + 1] = R.xyz[r][3 * l + 1]; xyz[3 * j + 2] = R.xyz[r][3 * l + 2];
    seed[j] = (int)(keys[j] >> 20); pos[j] = (int)(keys[j] & 0xfffffull);
  }
  const int64_t s0 = R.off[r][l], n = R.off[r][l + 1] - s0, d0 = obs_off[j];
  for (int64_t k = lane; k < n; k += 32) {
    ov[d0 + k] = R.ov[r][s0 + k]; opl[d0 + k] = R.opl[r][s0 + k]; oseg[d0 + k] = R.oseg[r][s0 + k];
    oxy[2 * (d0 + k)] = R.oxy[r][2 * (s0 + k)]; oxy[2 * (d0 + k) + 1] = R.oxy[r][2 * (s0 + k) + 1];
  }
}

// ================================================================================================ C ABI ============
extern "C" {

const char* eg3d_last_error(void) { return g_err.c_str(); }
void eg3d_camera_fundamentals(const float* cameras, int32_t n_views, double* out) {
  std::vector<double> fp; camera_fundamentals(cameras, n_views, fp);
  memcpy(out, fp.data(), fp.size() * sizeof(double));
}
int eg3d_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

void eg3d_params_default(eg3d_params* p) {
  memset(p, 0, sizeof(*p));
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

eg3d_status eg3d_scene_create(const eg3d_scene_desc* d, const eg3d_params* params, eg3d_scene** out) {
  if (!d || !out || d->n_views <= 0) return fail(EG3D_ERR_INVALID_ARG, "bad scene descriptor");
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  eg3d_scene* sc = new eg3d_scene();
  std::unique_ptr<eg3d_scene> guard(sc);
  CK(cudaGetDevice(&sc->device));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, sc->device));
  sc->num_sms = prop.multiProcessorCount;
  sc->sh = std::make_shared<StreamHolder>(); sc->sh->device = sc->device;
  CK(cudaStreamCreateWithFlags(&sc->sh->s, cudaStreamNonBlocking));
  sc->stream = sc->sh->s;
  g_alloc_stream = sc->stream;
  {
    cudaMemPool_t pool; CK(cudaDeviceGetDefaultMemPool(&pool, sc->device));
    uint64_t thr = ~0ull; CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  }
  if (params) sc->prm = *params; else eg3d_params_default(&sc->prm);
  const int V = sc->V = d->n_views; sc->width = d->width; sc->height = d->height;
  const int64_t NP = d->view_poly_off[V], NV = d->poly_vert_off[NP];
  if (NV > 0x7fffffff) return fail(EG3D_ERR_INVALID_ARG, "too many vertices");
  sc->h_view_poly_off.resize(V + 1); for (int i = 0; i <= V; i++) sc->h_view_poly_off[i] = (int)d->view_poly_off[i];
  sc->h_poly_vert_off.resize(NP + 1); for (int64_t i = 0; i <= NP; i++) sc->h_poly_vert_off[i] = (int)d->poly_vert_off[i];
  sc->h_verts.assign((const float2*)d->verts, (const float2*)d->verts + NV);
  sc->h_start.assign(d->poly_start, d->poly_start + NP); sc->h_end.assign(d->poly_end, d->poly_end + NP);
  // staged segments: every valid polyline, (P[i], P[i-1]) orientation, as (x1, y1, dx, dy)
  std::vector<float4> seg; std::vector<uint2> seg_id; std::vector<int> poly_seg_off(NP + 1, 0);
  sc->h_view_seg_off.assign(V + 1, 0);
  for (int v = 0; v < V; v++) {
    for (int g = sc->h_view_poly_off[v]; g < sc->h_view_poly_off[v + 1]; g++) {
      poly_seg_off[g] = (int)seg.size();
      const int o = sc->h_poly_vert_off[g], nv = sc->h_poly_vert_off[g + 1] - o;
      for (int i = 1; i < nv; i++) {
        float2 a = sc->h_verts[o + i], b = sc->h_verts[o + i - 1];
        seg.push_back(make_float4(a.x, a.y, b.x - a.x, b.y - a.y));
        seg_id.push_back(make_uint2((uint32_t)(g - sc->h_view_poly_off[v]), (uint32_t)(i - 1)));
      }
    }
    sc->h_view_seg_off[v + 1] = (int)seg.size();
    sc->max_view_segs = std::max(sc->max_view_segs, sc->h_view_seg_off[v + 1] - sc->h_view_seg_off[v]);
  }
  poly_seg_off[NP] = (int)seg.size();
  // K1 cull structures: groups (<= 32 segments of one polyline) with inflated boxes; chunks of <= K1_CHUNK segments and
  // <= K1_GMAX groups aligned to group boundaries (group ranges padded to a multiple of 4 for 16-byte bulk copies)
  std::vector<float4> gbox; std::vector<uint32_t> gdesc; std::vector<int4> chunks; std::vector<int> view_chunk_off(V + 1, 0);
  for (int v = 0; v < V; v++) {
    int cseg0 = sc->h_view_seg_off[v], cnseg = 0, cg0 = (int)gbox.size(), cng = 0;
    auto flush = [&]() {
      if (cng == 0) return;
      while (gbox.size() % 4) { gbox.push_back(make_float4(0.f, 0.f, -1.f, -1.f)); gdesc.push_back(0u); }
      chunks.push_back(make_int4(cseg0, cnseg, cg0, (int)gbox.size() - cg0));
      cseg0 += cnseg; cnseg = 0; cg0 = (int)gbox.size(); cng = 0;
    };
    for (int g = sc->h_view_poly_off[v]; g < sc->h_view_poly_off[v + 1]; g++) {
      const int ps0 = poly_seg_off[g], ps1 = (g + 1 <= (int)NP) ? poly_seg_off[g + 1] : (int)seg.size();
      const int o = sc->h_poly_vert_off[g];
      for (int k0 = ps0; k0 < ps1; k0 += K1_GROUP) {
        const int k1 = std::min(ps1, k0 + K1_GROUP), n = k1 - k0;
        if (cnseg + n > K1_CHUNK || cng + 1 > K1_GMAX - 3) flush();
        float xmin = 1e30f, xmax = -1e30f, ymin = 1e30f, ymax = -1e30f;
        for (int k = k0; k <= k1; k++) {   // vertices (k - ps0) .. (k1 - ps0) of the polyline
          const float2 q = sc->h_verts[o + (k - ps0)];
          xmin = std::min(xmin, q.x); xmax = std::max(xmax, q.x); ymin = std::min(ymin, q.y); ymax = std::max(ymax, q.y);
        }
        gbox.push_back(make_float4(0.5f * (xmin + xmax), 0.5f * (ymin + ymax), 0.5f * (xmax - xmin) + 0.05f, 0.5f * (ymax - ymin) + 0.05f));
        gdesc.push_back(((uint32_t)cnseg << 6) | (uint32_t)n);
        cnseg += n; cng++;
      }
    }
    flush();
    view_chunk_off[v + 1] = (int)chunks.size();
  }
  HostGrid g4; build_grid(*sc, sc->prm.expand_grid_cell, g4);
  const bool tracks = d->n_tracks > 0;
  if (tracks) build_grid(*sc, sc->prm.detection_starting_radius * sc->prm.detection_mult, sc->hg30);
  cudaStream_t s = sc->stream;
  CK(sc->P.upload(d->cameras, (size_t)V * 12, s));
  { std::vector<double> p64((size_t)V * 12); for (size_t i = 0; i < p64.size(); i++) p64[i] = (double)d->cameras[i]; CK(sc->P64.upload(p64, s)); CK(cudaStreamSynchronize(s)); }
  CK(sc->F.upload(d->fundamental, (size_t)V * V * 9, s));
  {
    std::vector<double> fp; camera_fundamentals(d->cameras, V, fp);
    std::vector<double> fph((size_t)V * V);
    for (size_t i = 0; i < fph.size(); i++) { const double* f = &fp[i * 9]; fph[i] = sqrt(f[0] * f[0] + f[1] * f[1] + f[3] * f[3] + f[4] * f[4]); }
    CK(sc->Fp.upload(fp, s)); CK(sc->Fph.upload(fph, s)); CK(cudaStreamSynchronize(s));
  }
  CK(sc->Fvalid.upload(d->fundamental_valid, (size_t)V * V, s));
  CK(sc->view_poly_off.upload(sc->h_view_poly_off, s)); CK(sc->poly_vert_off.upload(sc->h_poly_vert_off, s));
  CK(sc->verts.upload(sc->h_verts, s)); CK(sc->poly_start.upload(sc->h_start, s)); CK(sc->poly_end.upload(sc->h_end, s));
  CK(sc->view_seg_off.upload(sc->h_view_seg_off, s)); CK(sc->poly_seg_off.upload(poly_seg_off, s));
  CK(sc->seg.upload(seg, s)); CK(sc->seg_id.upload(seg_id, s));
  CK(sc->grp_box.upload(gbox, s)); CK(sc->grp_desc.upload(gdesc, s)); CK(sc->chunks.upload(chunks, s)); CK(sc->view_chunk_off.upload(view_chunk_off, s));
  CK(cudaStreamSynchronize(s));
  CK(sc->g4_off.upload(g4.off, s)); CK(sc->g4_ids.upload(g4.ids, s));
  if (tracks) {
    CK(sc->g30_off.upload(sc->hg30.off, s)); CK(sc->g30_ids.upload(sc->hg30.ids, s));
    const int64_t NO = d->track_off[d->n_tracks];
    CK(sc->track_xyz.upload(d->track_xyz, (size_t)d->n_tracks * 3, s)); CK(sc->track_off.upload(d->track_off, (size_t)d->n_tracks + 1, s));
    CK(sc->track_view.upload(d->track_view, (size_t)NO, s)); CK(sc->track_xy.upload((const float2*)d->track_xy, (size_t)NO, s));
    sc->h_track_off.assign(d->track_off, d->track_off + d->n_tracks + 1);
    {
      std::vector<int> ot((size_t)NO);
      for (int64_t t = 0; t < d->n_tracks; t++) for (int64_t o = d->track_off[t]; o < d->track_off[t + 1]; o++) ot[(size_t)o] = (int)t;
      CK(sc->obs_track.upload(ot, s));
    }
    sc->h_track_view.assign(d->track_view, d->track_view + NO);
    sc->h_track_xy.assign((const float2*)d->track_xy, (const float2*)d->track_xy + NO);
  }
  CK(cudaStreamSynchronize(s));
  DevScene& D = sc->dev; memset(&D, 0, sizeof D);
  D.V = V; D.width = sc->width; D.height = sc->height;
  D.P = sc->P.p; D.P64 = sc->P64.p; D.F = sc->F.p; D.Fp = sc->Fp.p; D.Fph = sc->Fph.p; D.Fvalid = sc->Fvalid.p;
  D.view_poly_off = sc->view_poly_off.p; D.poly_vert_off = sc->poly_vert_off.p; D.verts = sc->verts.p;
  D.poly_start = sc->poly_start.p; D.poly_end = sc->poly_end.p;
  D.view_seg_off = sc->view_seg_off.p; D.seg = sc->seg.p; D.seg_id = sc->seg_id.p; D.poly_seg_off = sc->poly_seg_off.p;
  D.grp_box = sc->grp_box.p; D.grp_desc = sc->grp_desc.p; D.chunks = sc->chunks.p; D.view_chunk_off = sc->view_chunk_off.p;
  D.g_expand.cell = g4.cell; D.g_expand.w = g4.w; D.g_expand.h = g4.h; D.g_expand.cell_off = sc->g4_off.p; D.g_expand.ids = sc->g4_ids.p;
  if (tracks) { D.g_corr.cell = sc->hg30.cell; D.g_corr.w = sc->hg30.w; D.g_corr.h = sc->hg30.h; D.g_corr.cell_off = sc->g30_off.p; D.g_corr.ids = sc->g30_ids.p; }
  D.n_tracks = d->n_tracks; D.track_xyz = sc->track_xyz.p; D.track_off = sc->track_off.p; D.track_view = sc->track_view.p; D.track_xy = sc->track_xy.p;
  D.prm = sc->prm;
  CK(cudaFuncSetAttribute(k1_sweep_kernel<K1_COUNT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES2));
  CK(cudaFuncSetAttribute(k1_sweep_kernel<K1_FILL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES2));
  CK(cudaFuncSetAttribute(k1_sweep_kernel<K1_ANY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES2));
  CK(cudaFuncSetAttribute(k1_sweep_kernel<K1_COUNT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES2));
  CK(cudaFuncSetAttribute(k1_sweep_kernel<K1_FILL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES2));
  CK(cudaFuncSetAttribute(k1_sweep_kernel<K1_ANY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES2));
  *out = guard.release();
  return EG3D_OK;
}

void eg3d_scene_destroy(eg3d_scene* sc) {
  if (!sc) return;
  cudaSetDevice(sc->device);
  if (sc->stream) cudaStreamSynchronize(sc->stream);
  if (sc->comm && nccl_api().ok) { nccl_api().CommDestroy(sc->comm); sc->comm = nullptr; }
  delete sc;   // buffers are returned to the pool on the stream, which lives until the last result of this scene is freed
}

eg3d_status eg3d_sample_seeds(const eg3d_scene_desc* d, const int32_t* views, const uint32_t* polylines, int64_t n_pl, float spacing,
                              int64_t capacity, int32_t* o_view, uint32_t* o_pl, uint32_t* o_seg, float* o_xy, int32_t* o_src, int64_t* n_out) {
  if (!d || !n_out) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  int64_t n = 0;
  for (int64_t k = 0; k < n_pl; k++) {
    const int64_t g = d->view_poly_off[views[k]] + polylines[k];
    Pl pl; pl.pc = (const float2*)d->verts + d->poly_vert_off[g]; pl.n = (int)(d->poly_vert_off[g + 1] - d->poly_vert_off[g]);
    pl.start = d->poly_start[g]; pl.end = d->poly_end[g];
    if (pl.n < 2) continue;
    PlP plp; plp.seg = 0; plp.c = pl.pc[0];                                  // get_start_plp
    bool reached;
    plp = step_by_distance(pl, plp, pl.end, spacing, reached);              // polyline_matching.cpp:172
    while (!reached) {
      if (n < capacity) { o_view[n] = views[k]; o_pl[n] = polylines[k]; o_seg[n] = plp.seg; o_xy[2 * n] = plp.c.x; o_xy[2 * n + 1] = plp.c.y; if (o_src) o_src[n] = (int32_t)k; }
      n++;
      plp = step_by_distance(pl, plp, pl.end, spacing, reached);            // :185
    }
  }
  *n_out = n;
  return n <= capacity ? EG3D_OK : fail(EG3D_ERR_CAPACITY, "seed capacity too small");
}

// ---------------------------------------------------------------------------------------------------------------
// Row f2 (host, upstream of pipeline 2): polyline_matching_closeness_to_refpoints,
// src/edgegraph3d/matching/polyline_matching/polyline_matcher.cpp:75-168.  Works on the caller's host arrays only.
// ---------------------------------------------------------------------------------------------------------------
struct eg3d_polyline_sets { int32_t n_sets = 0, V = 0; std::vector<int64_t> off; std::vector<uint32_t> ids; std::vector<int64_t> refpoints; };

struct CollectIds {   // grid_visit visitor usable from the host (count with cap = 0, then fill)
  uint32_t* buf; int* n; int cap;
  EG3D_HD void operator()(uint32_t id) const { if (*n < cap) buf[*n] = id; (*n)++; }
};
static void fill_host_plg(const eg3d_scene_desc* d, eg3d_scene& sc) {
  const int V = sc.V = d->n_views; sc.width = d->width; sc.height = d->height;
  const int64_t NP = d->view_poly_off[V], NV = d->poly_vert_off[NP];
  sc.h_view_poly_off.resize(V + 1); for (int i = 0; i <= V; i++) sc.h_view_poly_off[i] = (int)d->view_poly_off[i];
  sc.h_poly_vert_off.resize(NP + 1); for (int64_t i = 0; i <= NP; i++) sc.h_poly_vert_off[i] = (int)d->poly_vert_off[i];
  sc.h_verts.assign((const float2*)d->verts, (const float2*)d->verts + NV);
  sc.h_start.assign(d->poly_start, d->poly_start + NP); sc.h_end.assign(d->poly_end, d->poly_end + NP);
}

eg3d_status eg3d_polyline_sets_from_refpoints(const eg3d_scene_desc* d, float find_within_dist, float mult, eg3d_polyline_sets** out) {
  if (!d || !out) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (d->n_tracks <= 0 || !d->track_off || !d->track_view || !d->track_xy) return fail(EG3D_ERR_INVALID_ARG, "the scene has no SfM tracks");
  if (!(find_within_dist > 0)) return fail(EG3D_ERR_INVALID_ARG, "find_within_dist must be positive");
  const int64_t NVtot = d->poly_vert_off[d->view_poly_off[d->n_views]];
  if (NVtot > 0x7fffffff) return fail(EG3D_ERR_INVALID_ARG, "too many vertices");
  eg3d_scene tmp;                       // host copies only: no device buffer is ever allocated in it
  fill_host_plg(d, tmp);
  const int V = tmp.V;
  HostGrid hg; build_grid(tmp, find_within_dist, hg);   // PolyLine2DMapSearch(plg, img_sz, FIND_WITHIN_DIST), :80-81
  DevScene hs; memset(&hs, 0, sizeof hs);
  hs.V = V; hs.view_poly_off = tmp.h_view_poly_off.data(); hs.poly_vert_off = tmp.h_poly_vert_off.data();
  hs.verts = tmp.h_verts.data(); hs.poly_start = tmp.h_start.data(); hs.poly_end = tmp.h_end.data();
  DevGrid g; g.cell = hg.cell; g.w = hg.w; g.h = hg.h; g.cell_off = hg.off.data(); g.ids = hg.ids.data();
  const float search_dist_sq = find_within_dist * find_within_dist;
  // polyline-match graph: nodes = (view, polyline) in first-seen order, undirected set adjacency (:62-73, graph_adjacency_set_*.cpp)
  std::map<std::pair<int, uint32_t>, int64_t> node_of;
  std::vector<std::pair<int, uint32_t>> nodes;
  std::vector<std::set<int64_t>> adj;
  std::unique_ptr<eg3d_polyline_sets> r(new eg3d_polyline_sets());
  r->V = V;
  std::vector<uint32_t> near_ids;
  for (int64_t t = 0; t < d->n_tracks; t++) {
    const int64_t o0 = d->track_off[t], o1 = d->track_off[t + 1], n_obs = o1 - o0;
    size_t maxpl = 0;
    // per observing view: the polylines within find_within_dist of the observation (find_polylines_within_search_dist_with_reprojections,
    // polyLine_2d_map_search.cpp:119-135); only lists of length 1 are ever read, so (count, id, distance) is enough
    std::vector<size_t> cnt((size_t)n_obs, 0); std::vector<uint32_t> first_id((size_t)n_obs, 0); std::vector<float> first_d((size_t)n_obs, 0.f);
    for (int64_t k = 0; k < n_obs; k++) {
      const int cam = d->track_view[o0 + k];
      if (cam < 0 || cam >= V) return fail(EG3D_ERR_INVALID_ARG, "track_view out of range");
      // get_2d_coordinates_of_point_on_image returns the LAST observation of that camera (edge_graph_3d_utilities.cpp:382-393)
      float2 c = make_float2(0.f, 0.f);
      for (int64_t m = 0; m < n_obs; m++) if (d->track_view[o0 + m] == cam) c = make_float2(d->track_xy[2 * (o0 + m)], d->track_xy[2 * (o0 + m) + 1]);
      int n_near = 0;
      grid_visit(g, cam, tmp.width, tmp.height, c, CollectIds{nullptr, &n_near, 0});
      near_ids.resize((size_t)n_near);
      n_near = 0;
      grid_visit(g, cam, tmp.width, tmp.height, c, CollectIds{near_ids.data(), &n_near, (int)near_ids.size()});
      std::sort(near_ids.begin(), near_ids.end());
      near_ids.erase(std::unique(near_ids.begin(), near_ids.end()), near_ids.end());
      for (uint32_t id : near_ids) {
        Pl pl = get_pl(hs, cam, id);
        uint32_t seg; float2 proj;
        const float dsq = pl_distancesq(pl, c, seg, proj);
        if (dsq <= search_dist_sq) { if (cnt[k] == 0) { first_id[k] = id; first_d[k] = sqrtf(dsq); } cnt[k]++; }
      }
      maxpl = std::max(maxpl, cnt[k]);
    }
    if (maxpl != 1) continue;
    std::set<std::pair<int, uint32_t>> cams_pls;
    float min_dist = std::numeric_limits<float>::max(), max_dist = std::numeric_limits<float>::min();   // sic: min() is the smallest positive float
    for (int64_t k = 0; k < n_obs; k++) {
      if (cnt[k] == 0) continue;
      min_dist = min_dist <= first_d[k] ? min_dist : first_d[k];
      max_dist = max_dist >= first_d[k] ? max_dist : first_d[k];
      cams_pls.insert({d->track_view[o0 + k], first_id[k]});
    }
    if ((double)cams_pls.size() < (double)n_obs * 0.7) continue;
    if (min_dist < (max_dist / mult)) continue;
    if (max_dist > (min_dist * mult)) continue;
    if (cams_pls.size() < 2) continue;
    std::vector<int64_t> ids;
    for (const auto& cp : cams_pls) {
      auto it = node_of.find(cp);
      if (it == node_of.end()) { it = node_of.insert({cp, (int64_t)nodes.size()}).first; nodes.push_back(cp); adj.emplace_back(); }
      ids.push_back(it->second);
    }
    for (size_t i = 0; i < ids.size(); i++)
      for (size_t j = i + 1; j < ids.size(); j++) { adj[ids[i]].insert(ids[j]); adj[ids[j]].insert(ids[i]); }
    r->refpoints.push_back(t);
  }
  // GraphAdjacencySetUndirectedNoType::get_components (graph_adjacency_set_undirected_no_type.cpp:44-69): one candidate set
  // per connected component, in order of the component's lowest node id; per view an ascending set of polyline ids
  std::vector<uint8_t> visited(nodes.size(), 0);
  r->off.push_back(0);
  for (size_t s = 0; s < nodes.size(); s++) {
    if (visited[s]) continue;
    std::vector<std::set<uint32_t>> per_view(V);
    std::vector<int64_t> st{(int64_t)s};
    visited[s] = 1;
    while (!st.empty()) {
      const int64_t c = st.back(); st.pop_back();
      per_view[nodes[c].first].insert(nodes[c].second);
      for (int64_t nb : adj[c]) if (!visited[nb]) { visited[nb] = 1; st.push_back(nb); }
    }
    for (int v = 0; v < V; v++) { r->ids.insert(r->ids.end(), per_view[v].begin(), per_view[v].end()); r->off.push_back((int64_t)r->ids.size()); }
    r->n_sets++;
  }
  *out = r.release();
  return EG3D_OK;
}
eg3d_status eg3d_polyline_sets_get(const eg3d_polyline_sets* p, eg3d_candidates* view, int64_t* n_refpoints, const int64_t** refpoints) {
  if (!p || !view) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  view->n_sets = p->n_sets; view->off = p->off.data(); view->polyline = p->ids.data();
  if (n_refpoints) *n_refpoints = (int64_t)p->refpoints.size();
  if (refpoints) *refpoints = p->refpoints.data();
  return EG3D_OK;
}
void eg3d_polyline_sets_free(eg3d_polyline_sets* p) { delete p; }

// ---------------------------------------------------------------------------------------------------------------
// Row f2, pipeline 1's producer: polyline_matching_similarity_graph, polyline_matcher.cpp:222-336.  The weighted
// (view, polyline) compatibility graph is the reference's, number for number; its communities are found here by a
// deterministic sequential Louvain (the reference hands the graph to the vendored multi-threaded Grappolo, whose result
// depends on the thread count and interleaving and which returns nothing on one thread).
// ---------------------------------------------------------------------------------------------------------------
struct eg3d_similarity_graph {
  int32_t V = 0;
  std::vector<int32_t> node_view; std::vector<uint32_t> node_pl;
  std::vector<int64_t> ea, eb; std::vector<float> ew;      // node a < node b, in the reference's insertion order
  std::string dimacs;                                      // GraphAdjacencySetUndirectedNoTypeWeighted::write_to_file
};

// the polylines within find_within_dist of point c in view cam, ascending id (polyLine_2d_map_search.cpp:119-135)
static void near_polylines(const DevGrid& g, const DevScene& hs, int cam, int width, int height, float2 c, float dsq_max,
                           std::vector<uint32_t>& scratch, std::vector<std::pair<uint32_t, float>>& out) {
  out.clear();
  int n_near = 0;
  grid_visit(g, cam, width, height, c, CollectIds{nullptr, &n_near, 0});
  scratch.resize((size_t)n_near);
  n_near = 0;
  grid_visit(g, cam, width, height, c, CollectIds{scratch.data(), &n_near, (int)scratch.size()});
  std::sort(scratch.begin(), scratch.end());
  scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
  for (uint32_t id : scratch) {
    Pl pl = get_pl(hs, cam, id);
    uint32_t seg; float2 proj;
    const float dsq = pl_distancesq(pl, c, seg, proj);
    if (dsq <= dsq_max) out.push_back({id, sqrtf(dsq)});
  }
}

eg3d_status eg3d_polyline_similarity_graph(const eg3d_scene_desc* d, float find_within_dist, eg3d_similarity_graph** out) {
  if (!d || !out) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (d->n_tracks <= 0 || !d->track_off || !d->track_view || !d->track_xy) return fail(EG3D_ERR_INVALID_ARG, "the scene has no SfM tracks");
  if (!(find_within_dist > 0)) return fail(EG3D_ERR_INVALID_ARG, "find_within_dist must be positive");
  if (d->poly_vert_off[d->view_poly_off[d->n_views]] > 0x7fffffff) return fail(EG3D_ERR_INVALID_ARG, "too many vertices");
  eg3d_scene tmp; fill_host_plg(d, tmp);
  const int V = tmp.V; const int64_t NT = d->n_tracks;
  HostGrid hg; build_grid(tmp, find_within_dist, hg);
  DevScene hs; memset(&hs, 0, sizeof hs);
  hs.V = V; hs.view_poly_off = tmp.h_view_poly_off.data(); hs.poly_vert_off = tmp.h_poly_vert_off.data();
  hs.verts = tmp.h_verts.data(); hs.poly_start = tmp.h_start.data(); hs.poly_end = tmp.h_end.data();
  DevGrid g; g.cell = hg.cell; g.w = hg.w; g.h = hg.h; g.cell_off = hg.off.data(); g.ids = hg.ids.data();
  const float dsq_max = find_within_dist * find_within_dist;
  std::unique_ptr<eg3d_similarity_graph> r(new eg3d_similarity_graph()); r->V = V;
  std::map<std::pair<int, uint32_t>, int64_t> node_of;
  std::vector<std::set<int64_t>> adj;
  // close_refpoints[(view, polyline)] (ascending point ids), refpoint weights, visibility (:236-300)
  std::map<std::pair<int, uint32_t>, std::vector<int64_t>> close_refpoints;
  std::vector<float> weight((size_t)NT, 0.f);
  std::vector<std::vector<uint8_t>> visible((size_t)V, std::vector<uint8_t>((size_t)NT, 0));   // pointsVisibleFromCamN_ as a bitmap
  std::vector<uint32_t> scratch; std::vector<std::pair<uint32_t, float>> near;
  for (int64_t t = 0; t < NT; t++) {
    const int64_t o0 = d->track_off[t], o1 = d->track_off[t + 1];
    std::set<std::pair<int, uint32_t>> cams_pls;
    for (int64_t k = o0; k < o1; k++) {
      const int cam = d->track_view[k];
      if (cam < 0 || cam >= V) return fail(EG3D_ERR_INVALID_ARG, "track_view out of range");
      visible[cam][t] = 1;
      float2 c = make_float2(0.f, 0.f);   // the LAST observation of that camera (edge_graph_3d_utilities.cpp:382-393)
      for (int64_t m = o0; m < o1; m++) if (d->track_view[m] == cam) c = make_float2(d->track_xy[2 * m], d->track_xy[2 * m + 1]);
      near_polylines(g, hs, cam, tmp.width, tmp.height, c, dsq_max, scratch, near);
      for (const auto& q : near) cams_pls.insert({cam, q.first});
    }
    // compute_refpoint_weight (:196-205): views with close polylines / close polylines
    int non_empty = 0, sum_pls = 0, last_cam = -1;
    for (const auto& cp : cams_pls) { if (cp.first != last_cam) { non_empty++; last_cam = cp.first; } sum_pls++; }
    weight[t] = non_empty == 0 ? 0.0f : non_empty / ((float)sum_pls);
    std::vector<int64_t> ids;
    for (const auto& cp : cams_pls) {
      close_refpoints[cp].push_back(t);
      auto it = node_of.find(cp);
      if (it == node_of.end()) { it = node_of.insert({cp, (int64_t)r->node_view.size()}).first; r->node_view.push_back(cp.first); r->node_pl.push_back(cp.second); adj.emplace_back(); }
      ids.push_back(it->second);
    }
    for (size_t i = 0; i < ids.size(); i++)
      for (size_t j = i + 1; j < ids.size(); j++) { adj[ids[i]].insert(ids[j]); adj[ids[j]].insert(ids[i]); }
  }
  // compute_compatibility (:171-194): weighted Jaccard index of the two polylines' close SfM points, each list restricted
  // to the points that are also visible in the OTHER polyline's view
  const int64_t N = (int64_t)r->node_view.size();
  std::vector<std::map<int64_t, float>> wadj((size_t)N);
  std::vector<int64_t> la, lb, tmpv;
  for (int64_t n1 = 0; n1 < N; n1++)
    for (int64_t n2 : adj[n1]) {
      if (!(n1 < n2)) continue;
      const std::pair<int, uint32_t> m1{r->node_view[n1], r->node_pl[n1]}, m2{r->node_view[n2], r->node_pl[n2]};
      la.clear(); lb.clear();
      for (int64_t t : close_refpoints[m1]) if (visible[m2.first][t]) la.push_back(t);
      for (int64_t t : close_refpoints[m2]) if (visible[m1.first][t]) lb.push_back(t);
      tmpv.resize(la.size() + lb.size());
      auto e = std::set_intersection(la.begin(), la.end(), lb.begin(), lb.end(), tmpv.begin());
      float inter = 0.0f; for (auto it = tmpv.begin(); it != e; ++it) inter += weight[*it];
      float w = 0.0f;
      if (inter != 0.0f) {
        e = std::set_union(la.begin(), la.end(), lb.begin(), lb.end(), tmpv.begin());
        float uni = 0.0f; for (auto it = tmpv.begin(); it != e; ++it) uni += weight[*it];
        w = inter / uni;
      }
      if (w > 0.0f) { r->ea.push_back(n1); r->eb.push_back(n2); r->ew.push_back(w); wadj[n1][n2] = w; wadj[n2][n1] = w; }
    }
  // write_to_file (graph_adjacency_set_undirected_no_type_weighted.cpp:54-73): every edge appears in both directions and
  // the header counts both; weights in the stream's default float format
  {
    std::ostringstream f;
    int64_t m2 = 0; for (const auto& a : wadj) m2 += (int64_t)a.size();
    f << "p sp " << (unsigned long)N << " " << (unsigned long)m2 << "\n";
    for (int64_t n1 = 0; n1 < N; n1++) for (const auto& kv : wadj[n1]) f << "a " << (unsigned long)(n1 + 1) << " " << (unsigned long)(kv.first + 1) << " " << kv.second << "\n";
    r->dimacs = f.str();
  }
  *out = r.release();
  return EG3D_OK;
}

eg3d_status eg3d_similarity_graph_get(const eg3d_similarity_graph* p, eg3d_similarity_graph_view* v) {
  if (!p || !v) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  v->n_views = p->V; v->n_nodes = (int64_t)p->node_view.size(); v->node_view = p->node_view.data(); v->node_polyline = p->node_pl.data();
  v->n_edges = (int64_t)p->ew.size(); v->edge_a = p->ea.data(); v->edge_b = p->eb.data(); v->edge_weight = p->ew.data();
  v->dimacs = p->dimacs.c_str(); v->dimacs_len = (int64_t)p->dimacs.size();
  return EG3D_OK;
}
void eg3d_similarity_graph_free(eg3d_similarity_graph* p) { delete p; }

// Deterministic sequential Louvain (Blondel et al. 2008) on the weighted graph: nodes are visited in ascending id, a node
// moves to the neighbouring community with the largest modularity gain (ties: the lowest community id), levels are
// aggregated until a level makes no move.  Nodes without an edge get -1 (no community), the others ids 0..n-1 in order
// of first appearance.  *modularity (may be NULL) receives Q of the returned partition.
eg3d_status eg3d_similarity_graph_communities(const eg3d_similarity_graph* p, int64_t* community, double* modularity) {
  if (!p || !community) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  const int64_t N0 = (int64_t)p->node_view.size();
  std::vector<std::map<int64_t, double>> adj((size_t)N0);        // current level graph (self loops allowed after aggregation)
  for (size_t e = 0; e < p->ew.size(); e++) { adj[p->ea[e]][p->eb[e]] += (double)p->ew[e]; adj[p->eb[e]][p->ea[e]] += (double)p->ew[e]; }
  double m2 = 0; for (const auto& a : adj) for (const auto& kv : a) m2 += kv.second;   // 2m (a self loop {i: w} counts w, stored as 2*w below)
  std::vector<int64_t> owner((size_t)N0); for (int64_t i = 0; i < N0; i++) owner[i] = i;   // original node -> node of the current level
  int64_t N = N0;
  while (true) {
    std::vector<double> k((size_t)N, 0.0), tot((size_t)N, 0.0);
    std::vector<int64_t> com((size_t)N);
    for (int64_t i = 0; i < N; i++) { for (const auto& kv : adj[i]) k[i] += kv.second; tot[i] = k[i]; com[i] = i; }
    bool moved_any = false, moved = true;
    int sweeps = 0;
    while (moved && m2 > 0 && sweeps++ < 100) {
      moved = false;
      for (int64_t i = 0; i < N; i++) {
        std::map<int64_t, double> to;                                  // weight from i to each neighbouring community
        for (const auto& kv : adj[i]) if (kv.first != i) to[com[kv.first]] += kv.second;
        const int64_t old = com[i];
        tot[old] -= k[i];
        int64_t best = old; double best_gain = (to.count(old) ? to[old] : 0.0) - tot[old] * k[i] / m2;
        for (const auto& kv : to) {
          const double gain = kv.second - tot[kv.first] * k[i] / m2;
          if (gain > best_gain || (gain == best_gain && kv.first < best)) { best = kv.first; best_gain = gain; }
        }
        tot[best] += k[i];
        if (best != old) { com[i] = best; moved = true; moved_any = true; }
      }
    }
    // renumber the communities of this level in order of first appearance and aggregate
    std::vector<int64_t> renum((size_t)N, -1); int64_t nc = 0;
    for (int64_t i = 0; i < N; i++) if (renum[com[i]] < 0) renum[com[i]] = nc++;
    for (int64_t i = 0; i < N0; i++) owner[i] = renum[com[owner[i]]];
    if (!moved_any) break;
    std::vector<std::map<int64_t, double>> nadj((size_t)nc);
    for (int64_t i = 0; i < N; i++) for (const auto& kv : adj[i]) nadj[renum[com[i]]][renum[com[kv.first]]] += kv.second;
    adj.swap(nadj); N = nc;
  }
  // nodes without any edge belong to no community (-1), as in Grappolo's output; the others are renumbered 0..n-1 in
  // order of first appearance
  {
    std::vector<uint8_t> has_edge((size_t)N0, 0);
    for (size_t e = 0; e < p->ew.size(); e++) { has_edge[p->ea[e]] = 1; has_edge[p->eb[e]] = 1; }
    std::map<int64_t, int64_t> ren;
    for (int64_t i = 0; i < N0; i++) {
      if (!has_edge[i]) { owner[i] = -1; continue; }
      auto it = ren.find(owner[i]);
      if (it == ren.end()) it = ren.insert({owner[i], (int64_t)ren.size()}).first;
      owner[i] = it->second;
    }
  }
  for (int64_t i = 0; i < N0; i++) community[i] = owner[i];
  if (modularity) {
    // Q = sum_c [ in_c / 2m - (tot_c / 2m)^2 ] on the original graph
    std::map<int64_t, double> in, tot; double mm = 0;
    for (size_t e = 0; e < p->ew.size(); e++) {
      const double w = (double)p->ew[e]; mm += 2 * w;
      tot[owner[p->ea[e]]] += w; tot[owner[p->eb[e]]] += w;
      if (owner[p->ea[e]] == owner[p->eb[e]]) in[owner[p->ea[e]]] += 2 * w;
    }
    double q = 0; if (mm > 0) for (const auto& kv : tot) q += (in.count(kv.first) ? in[kv.first] : 0.0) / mm - (kv.second / mm) * (kv.second / mm);
    *modularity = q;
  }
  return EG3D_OK;
}

// compute_polyline_matches_from_nodes_component_ids (polyline_matcher.cpp:207-219): one candidate set per community id in
// [0, max id]; nodes with a negative id belong to none
eg3d_status eg3d_polyline_sets_from_communities(const eg3d_similarity_graph* p, const int64_t* community, eg3d_polyline_sets** out) {
  if (!p || !community || !out) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  const int64_t N = (int64_t)p->node_view.size(); const int V = p->V;
  int64_t maxc = -1; for (int64_t i = 0; i < N; i++) maxc = std::max(maxc, community[i]);
  std::vector<std::vector<std::set<uint32_t>>> sets((size_t)(maxc + 1), std::vector<std::set<uint32_t>>((size_t)V));
  for (int64_t i = 0; i < N; i++) if (community[i] >= 0) sets[(size_t)community[i]][(size_t)p->node_view[i]].insert(p->node_pl[i]);
  std::unique_ptr<eg3d_polyline_sets> r(new eg3d_polyline_sets()); r->V = V; r->n_sets = (int32_t)(maxc + 1);
  r->off.push_back(0);
  for (const auto& s : sets) for (int v = 0; v < V; v++) { r->ids.insert(r->ids.end(), s[v].begin(), s[v].end()); r->off.push_back((int64_t)r->ids.size()); }
  *out = r.release();
  return EG3D_OK;
}

// Host evaluation of compute_projection as the kernels compute it (eg3d::project; tests)
void eg3d_project_host(const float* cam12, const float* x3, float* out2) {
  const float2 q = project(cam12, x3[0], x3[1], x3[2]); out2[0] = q.x; out2[1] = q.y;
}
#define EG3D_STR2(x) #x
#define EG3D_STR(x) EG3D_STR2(x)
// compile-time switches of this build (A/B variants are told apart by it; tests gate on it)
const char* eg3d_build_info(void) {
  return "EG3D_DLT_OPENCV=1 EG3D_K3A_MIN_BLOCKS=" EG3D_STR(EG3D_K3A_MIN_BLOCKS) " EG3D_K3B_MIN_BLOCKS=" EG3D_STR(EG3D_K3B_MIN_BLOCKS)
         " EG3D_K3B_THREADS=" EG3D_STR(EG3D_K3B_THREADS) " EG3D_K3B_SYNC=" EG3D_STR(EG3D_K3B_SYNC) " EG3D_K3B_BATCH=" EG3D_STR(EG3D_K3B_BATCH)
         " EG3D_GN_UNROLL=" EG3D_STR(EG3D_GN_UNROLL) " EG3D_K1_GROUP=" EG3D_STR(EG3D_K1_GROUP);
}

// Host evaluation of the 2-view initialiser in its two forms (tests; no device needed): opencv_svd = 0 -> dlt_null (the
// kernels' current SVD), 1 -> dlt_null_opencv (OpenCV's own Jacobi SVD restated; see eg3d_dev.cuh).
void eg3d_triangulate_dlt_host(const float* P1, const float* P2, const float* x1, const float* x2, int32_t opencv_svd, float* out4) {
  if (opencv_svd) dlt_null_opencv(P1, P2, make_float2(x1[0], x1[1]), make_float2(x2[0], x2[1]), out4);
  else dlt_null(P1, P2, make_float2(x1[0], x1[1]), make_float2(x2[0], x2[1]), out4);
}

static eg3d_status check_seeds(const eg3d_scene* sc, const eg3d_seeds* seeds, const eg3d_candidates* cands) {
  if (!sc || !seeds) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (seeds->n < 0 || seeds->n > 0x7fffffff / std::max(1, sc->V)) return fail(EG3D_ERR_INVALID_ARG, "too many seeds in one call (n_seeds * n_views must fit 31 bits); split the batch");
  for (int64_t i = 0; i < seeds->n; i++) {
    const int v = seeds->view[i];
    if (v < 0 || v >= sc->V) return fail(EG3D_ERR_INVALID_ARG, "seed view out of range");
    if ((int64_t)seeds->polyline[i] >= (int64_t)(sc->h_view_poly_off[v + 1] - sc->h_view_poly_off[v])) return fail(EG3D_ERR_INVALID_ARG, "seed polyline id out of range");
  }
  if (cands) {
    if (!seeds->cand_set) return fail(EG3D_ERR_INVALID_ARG, "candidates given but seeds->cand_set is null");
    for (int64_t i = 0; i < seeds->n; i++)
      if (seeds->cand_set[i] < 0 || seeds->cand_set[i] >= cands->n_sets) return fail(EG3D_ERR_INVALID_ARG, "cand_set out of range (sweep and candidate seeds cannot be mixed in one call)");
  }
  return EG3D_OK;
}
static eg3d_status upload_cands(eg3d_scene* sc, const eg3d_candidates* c, DevCand& dc) {
  const size_t no = (size_t)c->n_sets * sc->V + 1;
  CK(dc.off.upload(c->off, no, sc->stream));
  CK(dc.pl.upload(c->polyline, (size_t)c->off[no - 1], sc->stream));
  return EG3D_OK;
}

eg3d_status eg3d_epipolar_intersect(eg3d_scene* sc, const eg3d_seeds* seeds, const eg3d_candidates* cands, eg3d_hits** out, eg3d_timing* tm) {
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  st = check_seeds(sc, seeds, cands); if (st != EG3D_OK) return st;
  CK(cudaSetDevice(sc->device));
  g_alloc_stream = sc->stream;
  eg3d_timing local; memset(&local, 0, sizeof local);
  DevSeeds ds; st = upload_seeds(sc, seeds, cands != nullptr, ds); if (st != EG3D_OK) return st;
  DevCand dc; if (cands) { st = upload_cands(sc, cands, dc); if (st != EG3D_OK) return st; }
  DBuf<int64_t> off; DBuf<eg3d_hit> hits; int64_t nh = 0;
  st = run_k1(sc, ds, cands ? &dc : nullptr, off, hits, nh, &local); if (st != EG3D_OK) return st;
  eg3d_hits* h = new eg3d_hits(); h->n_seeds = seeds->n; h->V = sc->V;
  h->off.resize((size_t)seeds->n * sc->V + 1); h->hits.resize((size_t)nh);
  cudaMemcpy(h->off.data(), off.p, h->off.size() * sizeof(int64_t), cudaMemcpyDeviceToHost);
  if (nh) cudaMemcpy(h->hits.data(), hits.p, (size_t)nh * sizeof(eg3d_hit), cudaMemcpyDeviceToHost);
  local.n_seeds = seeds->n; local.total_ms = local.k1_count_ms + local.scan_ms + local.k1_fill_ms;
  k1_accounting(sc, seeds, cands, nh, &local);
  if (tm) *tm = local;
  *out = h;
  return EG3D_OK;
}
eg3d_status eg3d_epipolar_intersect_device(eg3d_scene* sc, const eg3d_seeds* seeds, const eg3d_candidates* cands, eg3d_timing* tm) {
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  st = check_seeds(sc, seeds, cands); if (st != EG3D_OK) return st;
  CK(cudaSetDevice(sc->device));
  g_alloc_stream = sc->stream;
  eg3d_timing local; memset(&local, 0, sizeof local);
  DevSeeds ds; st = upload_seeds(sc, seeds, cands != nullptr, ds); if (st != EG3D_OK) return st;
  DevCand dc; if (cands) { st = upload_cands(sc, cands, dc); if (st != EG3D_OK) return st; }
  DBuf<int64_t> off; DBuf<eg3d_hit> hits; int64_t nh = 0;
  st = run_k1(sc, ds, cands ? &dc : nullptr, off, hits, nh, &local); if (st != EG3D_OK) return st;
  local.n_seeds = seeds->n; local.total_ms = local.k1_count_ms + local.scan_ms + local.k1_fill_ms;
  k1_accounting(sc, seeds, cands, nh, &local);
  if (tm) *tm = local;
  return EG3D_OK;
}
eg3d_status eg3d_hits_get(const eg3d_hits* h, int64_t* n_seeds, int32_t* n_views, const int64_t** off, const eg3d_hit** hits) {
  if (!h) return fail(EG3D_ERR_INVALID_ARG, "null hits");
  *n_seeds = h->n_seeds; *n_views = h->V; *off = h->off.data(); *hits = h->hits.data();
  return EG3D_OK;
}
void eg3d_hits_free(eg3d_hits* h) { delete h; }

eg3d_status eg3d_match_seeds(eg3d_scene* sc, const eg3d_seeds* seeds, const eg3d_candidates* cands, eg3d_points** out, eg3d_timing* tm) {
  WallClock wall;
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  st = check_seeds(sc, seeds, cands); if (st != EG3D_OK) return st;
  CK(cudaSetDevice(sc->device));
  g_alloc_stream = sc->stream;
  eg3d_timing local; memset(&local, 0, sizeof local);
  DevSeeds ds; st = upload_seeds(sc, seeds, cands != nullptr, ds); if (st != EG3D_OK) return st;
  DevCand dc; if (cands) { st = upload_cands(sc, cands, dc); if (st != EG3D_OK) return st; }
  Timer tall(sc->stream);
  tall.start();
  HitLists H;
  st = prepare_hits_a(sc, ds, cands ? &dc : nullptr, H, &local); if (st != EG3D_OK) return st;
  std::unique_ptr<eg3d_points> pts(new eg3d_points());
  st = run_k3(sc, ds, H, pts.get(), &local);
  tall.stop();
  local.n_seeds = seeds->n;
  local.total_ms = tall.ms();   // first kernel launch .. last kernel end, everything in between included
  k1_accounting(sc, seeds, cands, local.n_hits, &local);
  local.host_wall_ms = wall.ms();
  if (tm) *tm = local;
  if (st != EG3D_OK) return st;
  *out = pts.release();
  return EG3D_OK;
}

eg3d_status eg3d_match_polyline_sets(eg3d_scene* sc, const eg3d_candidates* c, int32_t view_begin, int32_t view_end, eg3d_points** out, eg3d_timing* tm) {
  WallClock wall;
  if (!sc || !c) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (view_begin < 0 || view_end > sc->V || view_begin > view_end) return fail(EG3D_ERR_INVALID_ARG, "starting-view range out of bounds");
  if (c->n_sets < 0 || (c->n_sets > 0 && (!c->off || c->off[0] != 0))) return fail(EG3D_ERR_INVALID_ARG, "bad candidate sets");
  for (size_t r = 0; r < (size_t)c->n_sets * sc->V; r++) {
    if (c->off[r + 1] < c->off[r]) return fail(EG3D_ERR_INVALID_ARG, "candidate offsets must be non-decreasing");
    const int v = (int)(r % sc->V), npl = sc->h_view_poly_off[v + 1] - sc->h_view_poly_off[v];
    for (int64_t k = c->off[r]; k < c->off[r + 1]; k++) if ((int64_t)c->polyline[k] >= npl) return fail(EG3D_ERR_INVALID_ARG, "candidate polyline id out of range");
  }
  // seed sampler (polyline_matching.cpp:162-190): sets in order, starting views ascending, candidate polylines ascending
  std::vector<int32_t> sv, scs; std::vector<uint32_t> spl, sseg; std::vector<float> sxy;
  DevScene hs; memset(&hs, 0, sizeof hs);
  hs.V = sc->V; hs.view_poly_off = sc->h_view_poly_off.data(); hs.poly_vert_off = sc->h_poly_vert_off.data();
  hs.verts = sc->h_verts.data(); hs.poly_start = sc->h_start.data(); hs.poly_end = sc->h_end.data();
  for (int set = 0; set < c->n_sets; set++)
    for (int v = view_begin; v < view_end; v++)
      for (int64_t k = c->off[(size_t)set * sc->V + v]; k < c->off[(size_t)set * sc->V + v + 1]; k++) {
        Pl pl = get_pl(hs, v, c->polyline[k]);
        if (pl.n < 2) continue;
        PlP plp; plp.seg = 0; plp.c = pl.pc[0];
        bool reached;
        plp = step_by_distance(pl, plp, pl.end, sc->prm.split_interval_distance, reached);
        while (!reached) {
          sv.push_back(v); scs.push_back(set); spl.push_back(c->polyline[k]); sseg.push_back(plp.seg); sxy.push_back(plp.c.x); sxy.push_back(plp.c.y);
          plp = step_by_distance(pl, plp, pl.end, sc->prm.split_interval_distance, reached);
        }
      }
  eg3d_seeds s; s.n = (int64_t)sv.size(); s.view = sv.data(); s.polyline = spl.data(); s.segment = sseg.data(); s.xy = sxy.data(); s.cand_set = scs.data();
  const float sample_ms = wall.ms();
  const eg3d_status st = eg3d_match_seeds(sc, &s, c, out, tm);
  if (tm) tm->host_wall_ms += sample_ms;
  return st;
}

eg3d_status eg3d_points_get(const eg3d_points* pc, eg3d_points_view* v) {
  if (!pc || !v) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  eg3d_points* p = const_cast<eg3d_points*>(pc);
  eg3d_status st = points_to_host(p); if (st != EG3D_OK) return st;
  v->n_points = p->n_points; v->n_obs = p->n_obs;
  v->xyz = p->xyz; v->seed = p->seed; v->chain_pos = p->chain_pos; v->obs_off = p->obs_off;
  v->obs_view = p->obs_view; v->obs_poly = p->obs_poly; v->obs_seg = p->obs_seg; v->obs_xy = p->obs_xy;
  return EG3D_OK;
}
eg3d_status eg3d_points_device_get(const eg3d_points* p, eg3d_points_view* v) {
  if (!p || !v) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  v->n_points = p->n_points; v->n_obs = p->n_obs;
  v->xyz = p->d_xyz.p; v->seed = p->d_seed.p; v->chain_pos = p->d_pos.p; v->obs_off = p->d_obs_off.p;
  v->obs_view = p->d_ov.p; v->obs_poly = p->d_opl.p; v->obs_seg = p->d_oseg.p; v->obs_xy = p->d_oxy.p;
  return EG3D_OK;
}
void eg3d_points_free(eg3d_points* p) { if (p) { cudaSetDevice(p->device); delete p; } }

}  // extern "C" (templates below)
// lanes per hypothesis of the warp-cooperative fp64 kernel: EG3D_GN_LANES=1|2|4|8|16|32 overrides (A/B runs)
static int gn64_lanes(int typical_obs) {
  if (const char* e = getenv("EG3D_GN_LANES")) { const int g = atoi(e); if (g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) return g; }
  return typical_obs >= 64 ? 8 : (typical_obs >= 32 ? 4 : 2);     // BASELINE configs[4] (20 observations): 9.7 / 7.5 / 8.1 / 10.7 ms for 1 / 2 / 4 / 8 lanes
}
template <int G>
static cudaError_t launch_gn64(eg3d_scene* sc, const GnProblem& pr, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(gn64_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int64_t want = ((int64_t)pr.n * G + GN_THREADS - 1) / GN_THREADS;
  int per_sm = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn64_kernel<G>, GN_THREADS, smem);
  if (e != cudaSuccess) return e;
  const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sc->num_sms * std::max(per_sm, 1)));   // persistent grid
  gn64_kernel<G><<<blocks, GN_THREADS, smem, sc->stream>>>(sc->dev, pr);
  return cudaGetLastError();
}
static eg3d_status launch_gn(eg3d_scene* sc, const GnProblem& pr, int fp64, int max_obs, eg3d_timing* tm) {
  Timer t(sc->stream);
  t.start();
  if (pr.n > 0) {
    if (fp64) {
      const size_t smem = (size_t)sc->V * 12 * sizeof(double);
      if (smem > 200 * 1024) return fail(EG3D_ERR_INVALID_ARG, "too many views for the stand-alone GN kernel's shared-memory camera table");
      cudaError_t e;
      switch (gn64_lanes(max_obs)) {
        case 1: e = launch_gn64<1>(sc, pr, smem); break;
        case 2: e = launch_gn64<2>(sc, pr, smem); break;
        case 8: e = launch_gn64<8>(sc, pr, smem); break;
        case 16: e = launch_gn64<16>(sc, pr, smem); break;
        case 32: e = launch_gn64<32>(sc, pr, smem); break;
        default: e = launch_gn64<4>(sc, pr, smem); break;
      }
      CK(e);
    } else {
      const size_t cam = (((size_t)sc->V * 12 + 3) & ~(size_t)3) * sizeof(float);
      const size_t stage = (size_t)max_obs * 12 * GN_THREADS;                       // float2 + int per observation and thread
      const bool use_stage = max_obs > 0 && cam + stage <= 56 * 1024 && !getenv("EG3D_GN_NO_STAGE");   // >= 4 CTAs per SM
      const size_t smem = cam + (use_stage ? stage : 0);
      if (smem > 200 * 1024) return fail(EG3D_ERR_INVALID_ARG, "too many views for the stand-alone GN kernel's shared-memory camera table");
      auto kern = use_stage ? gn32_kernel<true> : gn32_kernel<false>;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 1;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GN_THREADS, smem));
      // persistent grid: every resident thread walks the hypotheses with stride T (lane-level scheduling, see gn32_kernel)
      const int64_t want = (pr.n + GN_THREADS - 1) / GN_THREADS;
      unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sc->num_sms * std::max(per_sm, 1)));
      if (const char* e = getenv("EG3D_GN32_BLOCKS_PER_SM")) blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sc->num_sms * atoi(e)));   // A/B runs
      kern<<<blocks, GN_THREADS, smem, sc->stream>>>(sc->dev, pr, use_stage ? max_obs : 0);
#ifdef EG3D_GN_DEBUG
      { cudaStreamSynchronize(sc->stream); unsigned long long d[4]; cudaMemcpyFromSymbol(d, g_gn32_dbg, sizeof d); fprintf(stderr, "[gn32] blocks %u warp-trips %llu lane-steps %llu fresh %llu (cumulative)\n", blocks, d[0], d[1], d[2]); }
#endif
      CK(cudaGetLastError());
    }
  }
  t.stop();
  CK(cudaStreamSynchronize(sc->stream));
  if (tm) { tm->gn_ms += t.ms(); tm->total_ms += t.ms(); tm->kernel_launches += 1; }
  return EG3D_OK;
}

extern "C" {
eg3d_status eg3d_gn_triangulate(eg3d_scene* sc, int64_t n, const int64_t* obs_off, const int32_t* obs_view, const float* obs_xy,
                                const float* init_xyz, int fp64, float* out_xyz, float* out_mse, uint8_t* out_ok, eg3d_timing* tm) {
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  if (!sc || !obs_off) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  CK(cudaSetDevice(sc->device));
  g_alloc_stream = sc->stream;
  if (tm) memset(tm, 0, sizeof *tm);
  const int64_t no = obs_off[n];
  DBuf<int64_t> d_off; DBuf<int> d_v; DBuf<float2> d_xy; DBuf<float> d_init, d_x, d_m; DBuf<uint8_t> d_ok;
  CK(d_off.upload(obs_off, n + 1, sc->stream)); CK(d_v.upload(obs_view, no, sc->stream)); CK(d_xy.upload((const float2*)obs_xy, no, sc->stream));
  CK(d_init.upload(init_xyz, 3 * n, sc->stream)); CK(d_x.alloc(3 * n)); CK(d_m.alloc(n)); CK(d_ok.alloc(n));
  CK(cudaMemcpyAsync(d_x.p, d_init.p, 3 * n * sizeof(float), cudaMemcpyDeviceToDevice, sc->stream));
  GnProblem pr; memset(&pr, 0, sizeof pr);
  pr.n = n; pr.obs_off = d_off.p; pr.obs_view = d_v.p; pr.obs_xy = d_xy.p; pr.init = d_init.p;
  pr.out_xyz = d_x.p; pr.out_mse = d_m.p; pr.out_ok = d_ok.p; pr.gn_max_mse = sc->prm.filter_gn_max_mse; pr.write_back_only_ok = 0;
  int max_obs = 0; for (int64_t i = 0; i < n; i++) max_obs = std::max<int>(max_obs, (int)(obs_off[i + 1] - obs_off[i]));
  st = launch_gn(sc, pr, fp64, max_obs, tm); if (st != EG3D_OK) return st;
  CK(cudaMemcpy(out_xyz, d_x.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out_mse, d_m.p, n * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out_ok, d_ok.p, n, cudaMemcpyDeviceToHost));
  return EG3D_OK;
}

eg3d_status eg3d_gn_triangulate_device(eg3d_scene* sc, int64_t n, int32_t k, const int32_t* d_obs_view, const float* d_obs_xy,
                                       const float* d_init, int fp64, float* d_out_xyz, float* d_out_mse, uint8_t* d_out_ok, eg3d_timing* tm) {
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  if (!sc) return fail(EG3D_ERR_INVALID_ARG, "null scene");
  CK(cudaSetDevice(sc->device));
  g_alloc_stream = sc->stream;
  if (tm) memset(tm, 0, sizeof *tm);
  GnProblem pr; memset(&pr, 0, sizeof pr);
  pr.n = n; pr.obs_off = nullptr; pr.k = k; pr.obs_view = d_obs_view; pr.obs_xy = (const float2*)d_obs_xy; pr.init = d_init;
  pr.out_xyz = d_out_xyz; pr.out_mse = d_out_mse; pr.out_ok = d_out_ok; pr.gn_max_mse = sc->prm.filter_gn_max_mse; pr.write_back_only_ok = 0;
  return launch_gn(sc, pr, fp64, k, tm);
}

// a13: order-dependent first-come-first-kept density limiter (filtering_close_plgps.cpp:75-124).  The dependency chain
// is inherently sequential in the visiting order; it runs on the host over the gathered records.
eg3d_status eg3d_dedup_close_points(eg3d_scene* sc, const eg3d_points_view* pts, uint8_t* keep) {
  if (!sc || !pts || !keep) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  const float cs = sc->prm.dedup_cell;
  const int w = (int)ceilf((float)sc->width / cs), h = (int)ceilf((float)sc->height / cs);
  std::vector<std::vector<uint8_t>> bm(sc->V);
  auto cell = [&](int64_t o) -> size_t { return (size_t)(int)(pts->obs_xy[2 * o + 1] / cs) * w + (size_t)(int)(pts->obs_xy[2 * o] / cs); };
  // the reference indexes its bitmaps unchecked (filtering_close_plgps.cpp:77-85: undefined behaviour for an observation outside
  // the image); a public C ABI refuses such input instead
  for (int64_t o = 0; o < pts->n_obs; o++) {
    const float x = pts->obs_xy[2 * o], y = pts->obs_xy[2 * o + 1];
    if (pts->obs_view[o] < 0 || pts->obs_view[o] >= sc->V || !(x >= 0) || !(y >= 0) || (int)(x / cs) >= w || (int)(y / cs) >= h)
      return fail(EG3D_ERR_INVALID_ARG, "an observation lies outside its image (or names a view that does not exist)");
  }
  for (int64_t i = 0; i < pts->n_points; i++) {
    bool is_new = false;
    for (int64_t o = pts->obs_off[i]; o < pts->obs_off[i + 1]; o++) {
      auto& b = bm[pts->obs_view[o]];
      if (b.empty()) b.assign((size_t)w * h, 0);
      if (!b[cell(o)]) { is_new = true; break; }
    }
    keep[i] = is_new;
    if (is_new)
      for (int64_t o = pts->obs_off[i]; o < pts->obs_off[i + 1]; o++) {
        auto& b = bm[pts->obs_view[o]];
        if (b.empty()) b.assign((size_t)w * h, 0);
        b[cell(o)] = 1;
      }
  }
  return EG3D_OK;
}

eg3d_status eg3d_filter(eg3d_scene* sc, int64_t n, float* xyz, const int64_t* obs_off, const int32_t* obs_view, const float* obs_xy,
                        int64_t first_edgepoint, float gn_max_mse, int32_t forced_min_filter, uint8_t* inliers, eg3d_timing* tm) {
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  if (!sc || !obs_off) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  CK(cudaSetDevice(sc->device));
  g_alloc_stream = sc->stream;
  if (tm) memset(tm, 0, sizeof *tm);
  const int64_t no = obs_off[n];
  DBuf<int64_t> d_off; DBuf<int> d_v; DBuf<float2> d_xy; DBuf<float> d_x; DBuf<uint8_t> d_ok;
  CK(d_off.upload(obs_off, n + 1, sc->stream)); CK(d_v.upload(obs_view, no, sc->stream)); CK(d_xy.upload((const float2*)obs_xy, no, sc->stream));
  CK(d_x.upload(xyz, 3 * n, sc->stream)); CK(d_ok.alloc(n));
  GnProblem pr; memset(&pr, 0, sizeof pr);
  pr.n = n; pr.obs_off = d_off.p; pr.obs_view = d_v.p; pr.obs_xy = d_xy.p; pr.init = d_x.p;
  pr.out_xyz = d_x.p; pr.out_mse = nullptr; pr.out_ok = d_ok.p; pr.gn_max_mse = gn_max_mse; pr.write_back_only_ok = 1;
  int max_obs = 0; for (int64_t i = 0; i < n; i++) max_obs = std::max<int>(max_obs, (int)(obs_off[i + 1] - obs_off[i]));
  st = launch_gn(sc, pr, 0, max_obs, tm); if (st != EG3D_OK) return st;   // gaussNewtonFiltering, gauss_newton.cpp:136-178
  CK(cudaMemcpy(xyz, d_x.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(inliers, d_ok.p, n, cudaMemcpyDeviceToHost));
  // compute_ray_stats + view-count rule (outliers_filtering.cpp:14-64): O(n) bookkeeping on the returned bitmap
  std::vector<int64_t> dist(sc->V, 0);
  int64_t count = 0;
  for (int64_t i = 0; i < n; i++) if (inliers[i]) { count++; int64_t len = obs_off[i + 1] - obs_off[i]; if (len >= 1 && len <= sc->V) dist[len - 1]++; }
  int median = 0; int64_t m = 0;
  for (median = 0; median < sc->V; median++) { m += dist[median]; if (m >= count / 2) break; }
  int intended = (sc->prm.filter_3views_amount >= median / 2 - 1) ? sc->prm.filter_3views_amount : (median / 2 - 1);
  if (forced_min_filter > -1) intended = forced_min_filter;
  for (int64_t i = first_edgepoint; i < n; i++) inliers[i] = inliers[i] && ((obs_off[i + 1] - obs_off[i]) > intended);
  return EG3D_OK;
}

// B2 / a6: plg_matching_from_refpoints (plg_matching_from_refpoints.cpp:64-104).  Seeding — the polylines within 30 px of
// every observation (30 px grid), the seeds within 10 px (projection onto the polyline) and the per-seed radius
// (plg_edge_manager.cpp:261-288) — runs on the device, one warp per observation (eg3d_a6.cuh); the epipolar
// intersections with the radius filter (:191-259) run in k1_cand_kernel, then K3 exactly as for pipelines 1-2
// (plgpcm_3views_plg_following.cpp:40-50 scatters the per-observing-view lists into a V-vector = the CSR rows).
}  // extern "C"
// a6 seeding on the device for the tracks [tb, te): seeds (view, polyline, segment, xy, row = track * V), per-seed radius and the
// candidate rows, ready for k1_cand_kernel.  Shared by eg3d_match_refpoints and eg3d_refpoint_correspondences.
static eg3d_status refpoint_seeding(eg3d_scene* sc, int64_t tb, int64_t te, DevSeeds& ds, DevCand& dc, int64_t& n_seeds, eg3d_timing& local) {
  const int V = sc->V;
  const int64_t nt = te - tb, o_begin = sc->h_track_off[tb], o_end = sc->h_track_off[te], n_o = o_end - o_begin;
  const size_t n_rows = (size_t)nt * V;
  DBuf<A6Rec> recs; DBuf<int> n_ids, n_cand, n_seed, ovf; DBuf<unsigned char> is_last; DBuf<int64_t> seed_off, coff;
  CK(n_ids.alloc(n_o)); CK(n_cand.alloc(n_o)); CK(n_seed.alloc(n_o + 1)); CK(is_last.alloc(n_o)); CK(ovf.alloc(1));
  CK(seed_off.alloc(n_o + 1)); CK(coff.alloc(n_rows + 1));
  A6Args a; memset(&a, 0, sizeof a);
  a.o_begin = o_begin; a.o_end = o_end; a.tb = tb; a.obs_track = sc->obs_track.p;
  a.n_ids = n_ids.p; a.n_cand = n_cand.p; a.n_seed = n_seed.p; a.is_last = is_last.p; a.overflow = ovf.p; a.row_cnt = coff.p;
  Timer ts(sc->stream);
  ts.start();
  // Capacities per observation grow on demand: the reference has no bound on the polylines in a 30 px neighbourhood, so a dense
  // edge map must not fail the call (EG3D_TEST_A6_CAPS="raw,ids" starts from tiny values to exercise the retry).
  int raw_cap = A6_RAW, ids_cap = A6_IDS;
  if (const char* e = getenv("EG3D_TEST_A6_CAPS")) { int r0 = 0, i0 = 0; if (sscanf(e, "%d,%d", &r0, &i0) == 2 && r0 >= 4 && i0 >= 4) { raw_cap = r0 & ~3; ids_cap = i0 & ~3; } }
  unsigned blocks = 0; int threads = A6_THREADS;
  int64_t n_cands = 0; int h_ovf = 0;
  for (;;) {
    threads = a6_smem_per_warp(raw_cap, ids_cap) * (A6_THREADS / 32) <= 200 * 1024 ? A6_THREADS : 32;       // one warp per CTA for the largest tiers
    const size_t smem = a6_smem_per_warp(raw_cap, ids_cap) * (threads / 32);
    static thread_local size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) { CK(cudaFuncSetAttribute(a6_classify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); smem_set = smem; }
    CK(recs.alloc((size_t)n_o * ids_cap));
    CK(cudaMemsetAsync(ovf.p, 0, sizeof(int), sc->stream));
    CK(cudaMemsetAsync(n_seed.p, 0, (size_t)(n_o + 1) * sizeof(int), sc->stream));
    CK(cudaMemsetAsync(coff.p, 0, (n_rows + 1) * sizeof(int64_t), sc->stream));
    a.raw_cap = raw_cap; a.ids_cap = ids_cap; a.recs = recs.p;
    blocks = (unsigned)(((size_t)n_o * 32 + threads - 1) / threads);
    if (n_o > 0) a6_classify_kernel<<<blocks, threads, smem, sc->stream>>>(sc->dev, a);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&h_ovf, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, sc->stream));
    CK(cudaStreamSynchronize(sc->stream));
    local.kernel_launches += 1;
    if (!h_ovf) break;
    if (raw_cap >= A6_RAW_MAX) return fail(EG3D_ERR_CAPACITY, "an observation has more than 16384 polyline entries in its 30 px neighbourhood (A6_RAW_MAX)");
    raw_cap = std::min(A6_RAW_MAX, raw_cap * 4); ids_cap = std::min(A6_RAW_MAX / 4, ids_cap * 4);
    local.n_capacity_retries += 1;
  }
  {
    size_t tb1 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb1, n_seed.p, seed_off.p, n_o + 1, sc->stream);
    DBuf<unsigned char> tmp; CK(tmp.alloc(tb1));
    cub::DeviceScan::ExclusiveSum(tmp.p, tb1, n_seed.p, seed_off.p, n_o + 1, sc->stream);
  }
  eg3d_status st = exclusive_scan_i64(sc, coff.p, n_rows + 1); if (st != EG3D_OK) return st;
  n_seeds = 0;
  CK(cudaMemcpyAsync(&n_seeds, seed_off.p + n_o, sizeof(int64_t), cudaMemcpyDeviceToHost, sc->stream));
  CK(cudaMemcpyAsync(&n_cands, coff.p + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, sc->stream));
  CK(cudaStreamSynchronize(sc->stream));
  if (n_seeds > 0x7fffffff) return fail(EG3D_ERR_CAPACITY, "too many seeds in one call; split the track range");
  ds.n = (int)n_seeds;
  CK(ds.view.alloc(n_seeds)); CK(ds.pl.alloc(n_seeds)); CK(ds.seg.alloc(n_seeds)); CK(ds.xy.alloc(n_seeds)); CK(ds.cand_set.alloc(n_seeds));
  dc.filtered = true;
  CK(dc.pl.alloc(n_cands)); CK(dc.center.alloc(n_rows)); CK(dc.seed_r2.alloc(n_seeds));
  CK(cudaMemsetAsync(dc.center.p, 0, std::max<size_t>(n_rows, 1) * sizeof(float2), sc->stream));
  a.seed_off = seed_off.p; a.coff = coff.p;
  a.s_view = ds.view.p; a.s_pl = ds.pl.p; a.s_seg = ds.seg.p; a.s_xy = ds.xy.p; a.s_set = ds.cand_set.p; a.s_r2 = dc.seed_r2.p;
  a.cpl = dc.pl.p; a.center = dc.center.p;
  if (n_o > 0) a6_fill_kernel<<<(unsigned)(((size_t)n_o * 32 + A6_THREADS - 1) / A6_THREADS), A6_THREADS, 0, sc->stream>>>(sc->dev, a);
  ts.stop();
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(sc->stream));
  local.scan_ms += ts.ms(); local.kernel_launches += 4;
  // hand the CSR offsets to the candidate descriptor (the buffer changes owner)
  dc.off.p = coff.p; dc.off.n = coff.n; dc.off.s = coff.s; coff.p = nullptr;
  return EG3D_OK;
}
struct eg3d_corr {
  int V = 0; int64_t n_seeds = 0;
  std::vector<int32_t> view; std::vector<uint32_t> pl, seg; std::vector<float> xy; std::vector<int64_t> track; std::vector<int64_t> off; std::vector<eg3d_hit> hits;
};
extern "C" {

eg3d_status eg3d_match_refpoints(eg3d_scene* sc, int64_t tb, int64_t te, eg3d_points** out, eg3d_timing* tm) {
  WallClock wall;
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  if (!sc || !out) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (sc->dev.n_tracks <= 0) return fail(EG3D_ERR_INVALID_ARG, "the scene was created without SfM tracks");
  if (tb < 0 || te > sc->dev.n_tracks || tb > te) return fail(EG3D_ERR_INVALID_ARG, "track range out of bounds");
  CK(cudaSetDevice(sc->device));
  g_alloc_stream = sc->stream;
  eg3d_timing local; memset(&local, 0, sizeof local);
  Timer tall(sc->stream);
  tall.start();
  DevSeeds ds; DevCand dc; int64_t n_seeds = 0;
  st = refpoint_seeding(sc, tb, te, ds, dc, n_seeds, local); if (st != EG3D_OK) return st;
  HitLists H;
  st = prepare_hits_a(sc, ds, &dc, H, &local); if (st != EG3D_OK) return st;
  std::unique_ptr<eg3d_points> pts(new eg3d_points());
  st = run_k3(sc, ds, H, pts.get(), &local);
  tall.stop();
  local.n_seeds = n_seeds;
  local.total_ms = tall.ms();
  local.host_wall_ms = wall.ms();
  if (tm) *tm = local;
  if (st != EG3D_OK) return st;
  *out = pts.release();
  return EG3D_OK;
}

// B3 (EdgeManager side): PLGEdgeManager::detect_nearby_intersections_and_correspondences_plgp(refpoint) for the tracks [tb, te)
// (plg_edge_manager.cpp:261-300): the seeds (projections onto the polylines within 10 px of an observation) and, per seed, the
// epipolar hits on the other observing views within the 30 px grid neighbourhood and the radius 3 x |observation - seed|.
// Exactly what eg3d_match_refpoints feeds its own consensus step; here it is handed back so that a caller can run its own.
eg3d_status eg3d_refpoint_correspondences(eg3d_scene* sc, int64_t tb, int64_t te, eg3d_corr** out) {
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  if (!sc || !out) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (sc->dev.n_tracks <= 0) return fail(EG3D_ERR_INVALID_ARG, "the scene was created without SfM tracks");
  if (tb < 0 || te > sc->dev.n_tracks || tb > te) return fail(EG3D_ERR_INVALID_ARG, "track range out of bounds");
  CK(cudaSetDevice(sc->device));
  g_alloc_stream = sc->stream;
  eg3d_timing local; memset(&local, 0, sizeof local);
  DevSeeds ds; DevCand dc; int64_t n_seeds = 0;
  st = refpoint_seeding(sc, tb, te, ds, dc, n_seeds, local); if (st != EG3D_OK) return st;
  DBuf<int64_t> off; DBuf<eg3d_hit> hits; int64_t nh = 0;
  st = run_k1(sc, ds, &dc, off, hits, nh, &local); if (st != EG3D_OK) return st;
  std::unique_ptr<eg3d_corr> c(new eg3d_corr());
  c->V = sc->V; c->n_seeds = n_seeds;
  c->view.resize(n_seeds); c->pl.resize(n_seeds); c->seg.resize(n_seeds); c->xy.resize(2 * n_seeds); c->track.resize(n_seeds);
  c->off.resize((size_t)n_seeds * sc->V + 1); c->hits.resize((size_t)nh);
  std::vector<int> set((size_t)n_seeds);
  if (n_seeds > 0) {
    CK(cudaMemcpy(c->view.data(), ds.view.p, n_seeds * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c->pl.data(), ds.pl.p, n_seeds * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c->seg.data(), ds.seg.p, n_seeds * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c->xy.data(), ds.xy.p, n_seeds * sizeof(float2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(set.data(), ds.cand_set.p, n_seeds * sizeof(int), cudaMemcpyDeviceToHost));
  }
  for (int64_t i = 0; i < n_seeds; i++) c->track[i] = tb + set[i];                    // candidate row = track - tb
  CK(cudaMemcpy(c->off.data(), off.p, c->off.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
  if (nh) CK(cudaMemcpy(c->hits.data(), hits.p, (size_t)nh * sizeof(eg3d_hit), cudaMemcpyDeviceToHost));
  *out = c.release();
  return EG3D_OK;
}
eg3d_status eg3d_corr_get(const eg3d_corr* c, int64_t* n_seeds, int32_t* n_views, const int32_t** view, const uint32_t** polyline, const uint32_t** segment,
                          const float** xy, const int64_t** track, const int64_t** hit_off, const eg3d_hit** hits) {
  if (!c) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (n_seeds) *n_seeds = c->n_seeds; if (n_views) *n_views = c->V;
  if (view) *view = c->view.data(); if (polyline) *polyline = c->pl.data(); if (segment) *segment = c->seg.data(); if (xy) *xy = c->xy.data();
  if (track) *track = c->track.data(); if (hit_off) *hit_off = c->off.data(); if (hits) *hits = c->hits.data();
  return EG3D_OK;
}
void eg3d_corr_free(eg3d_corr* c) { delete c; }

// B4: compute_3D_point_multiple_views_plg_following_expandallviews_vector (triangulation.hpp:98, triangulation.cpp:1027-1088) for a
// batch of seeds whose V hit lists the CALLER supplies (any EdgeManager / correspondence search): view-triple selection,
// triple enumeration with the uniqueness test, PLG following, view expansion — K3 alone, no K1.
eg3d_status eg3d_match_correspondences(eg3d_scene* sc, int64_t n_seeds, const int32_t* start_view, const int64_t* hit_off, const eg3d_hit* hits,
                                       eg3d_points** out, eg3d_timing* tm) {
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  if (!sc || !out || n_seeds < 0 || (n_seeds > 0 && (!start_view || !hit_off))) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (n_seeds > 0x7fffffff / std::max(1, sc->V)) return fail(EG3D_ERR_INVALID_ARG, "too many seeds in one call");
  const int V = sc->V;
  const size_t n_rows = (size_t)n_seeds * V;
  for (int64_t i = 0; i < n_seeds; i++) if (start_view[i] < 0 || start_view[i] >= V) return fail(EG3D_ERR_INVALID_ARG, "start_view out of range");
  for (size_t r = 0; r < n_rows; r++) if (hit_off[r + 1] < hit_off[r]) return fail(EG3D_ERR_INVALID_ARG, "hit_off must be non-decreasing");
  const int64_t nh = n_seeds > 0 ? hit_off[n_rows] : 0;
  if (nh > 0 && !hits) return fail(EG3D_ERR_INVALID_ARG, "null hits");
  CK(cudaSetDevice(sc->device));
  g_alloc_stream = sc->stream;
  eg3d_timing local; memset(&local, 0, sizeof local);
  DevSeeds ds; ds.n = (int)n_seeds;
  CK(ds.view.upload(start_view, (size_t)n_seeds, sc->stream));
  CK(ds.pl.alloc((size_t)n_seeds)); CK(ds.seg.alloc((size_t)n_seeds)); CK(ds.xy.alloc((size_t)n_seeds));
  Timer tall(sc->stream);
  tall.start();
  HitLists H; H.lazy = false;
  if (n_seeds > 0) { CK(H.off_a.upload(hit_off, n_rows + 1, sc->stream)); CK(H.hits_a.upload(hits, (size_t)nh, sc->stream)); }
  else { CK(H.off_a.alloc(1)); CK(H.hits_a.alloc(0)); }
  st = select_views(sc, ds, H, &local); if (st != EG3D_OK) return st;
  std::unique_ptr<eg3d_points> pts(new eg3d_points());
  st = run_k3(sc, ds, H, pts.get(), &local);
  tall.stop();
  local.n_seeds = n_seeds; local.n_hits = nh;
  local.total_ms = tall.ms();
  if (tm) *tm = local;
  if (st != EG3D_OK) return st;
  *out = pts.release();
  return EG3D_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// the path's one exchange step.  Every rank contributes the accepted points of its shard (device resident);
// every rank receives the records of all ranks, merged on the device into the reference's loop order = ascending
// (global seed ordinal, chain position): one ncclAllGather of the counts, ONE grouped broadcast of the byte-packed
// records (rank r's buffer to everybody, all ranks in one NCCL group: no padding, no staging copies), one radix sort.
eg3d_status eg3d_comm_unique_id(uint8_t id[EG3D_COMM_ID_BYTES]) {
  if (!id) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (!nccl_api().ok) return fail(EG3D_ERR_CUDA, "NCCL (libnccl.so.2) could not be loaded");
  static_assert(sizeof(ncclUniqueId) == EG3D_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId u; NCK(nccl_api().GetUniqueId(&u));
  memcpy(id, &u, sizeof u);
  return EG3D_OK;
}
eg3d_status eg3d_comm_create(eg3d_scene* sc, const uint8_t id[EG3D_COMM_ID_BYTES], int32_t rank, int32_t world) {
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  if (!sc || !id || world < 1 || world > 64 || rank < 0 || rank >= world) return fail(EG3D_ERR_INVALID_ARG, "bad communicator arguments (1 <= world <= 64)");
  if (!nccl_api().ok) return fail(EG3D_ERR_CUDA, "NCCL (libnccl.so.2) could not be loaded");
  if (sc->comm) return fail(EG3D_ERR_INVALID_ARG, "the scene already has a communicator");
  CK(cudaSetDevice(sc->device));
  ncclUniqueId u; memcpy(&u, id, sizeof u);
  NCK(nccl_api().CommInitRank(&sc->comm, world, u, rank));
  sc->comm_rank = rank; sc->comm_world = world;
  return EG3D_OK;
}
eg3d_status eg3d_comm_destroy(eg3d_scene* sc) {
  if (!sc) return fail(EG3D_ERR_INVALID_ARG, "null scene");
  if (sc->comm) { cudaSetDevice(sc->device); cudaStreamSynchronize(sc->stream); nccl_api().CommDestroy(sc->comm); sc->comm = nullptr; sc->comm_world = 1; sc->comm_rank = 0; }
  return EG3D_OK;
}

eg3d_status eg3d_points_allgather(eg3d_scene* sc, const eg3d_points* mine, const int64_t* seed_global, eg3d_points** out, eg3d_timing* tm) {
  eg3d_status st = require_device(); if (st != EG3D_OK) return st;
  if (!sc || !mine || !out) return fail(EG3D_ERR_INVALID_ARG, "null argument");
  if (!sc->comm) return fail(EG3D_ERR_INVALID_ARG, "no communicator: call eg3d_comm_create first");
  CK(cudaSetDevice(sc->device));
  g_alloc_stream = sc->stream;
  auto& N = nccl_api();
  const int W = sc->comm_world, me = sc->comm_rank;
  cudaStream_t s = sc->stream;
  Timer tall(s); tall.start();
  // 1) counts of every rank: (points, observations, seeds of the call)
  DBuf<int64_t> d_cnt, d_cnts; CK(d_cnt.alloc(3)); CK(d_cnts.alloc(3 * (size_t)W));
  const int64_t h_cnt[3] = {mine->n_points, mine->n_obs, mine->n_seeds};
  CK(cudaMemcpyAsync(d_cnt.p, h_cnt, sizeof h_cnt, cudaMemcpyHostToDevice, s));
  NCK(N.AllGather(d_cnt.p, d_cnts.p, 3, ncclInt64, sc->comm, s));
  std::vector<int64_t> cnts(3 * (size_t)W);
  CK(cudaMemcpyAsync(cnts.data(), d_cnts.p, cnts.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));                       // the one host sync of the exchange: buffer sizes
  std::vector<PackLayout> L(W);
  int64_t NP = 0, NO = 0, seed_base = 0;
  XchgRanks R; memset(&R, 0, sizeof R); R.world = W;
  for (int r = 0; r < W; r++) {
    L[r] = pack_layout(cnts[3 * r], cnts[3 * r + 1]);
    R.pbase[r] = NP; R.obase[r] = NO; NP += cnts[3 * r]; NO += cnts[3 * r + 1];
    if (r < me) seed_base += cnts[3 * r + 2];
  }
  R.pbase[W] = NP; R.obase[W] = NO;
  // 2) pack this rank's records straight into its slot of the receive area, then one grouped broadcast per rank
  size_t total = 0; std::vector<size_t> slot(W);
  for (int r = 0; r < W; r++) { slot[r] = total; total += L[r].bytes; }
  DBuf<unsigned char> recv; CK(recv.alloc(total));
  DBuf<int64_t> d_sg;
  if (seed_global && mine->n_seeds > 0) CK(d_sg.upload(seed_global, (size_t)mine->n_seeds, s));
  {
    unsigned char* b = recv.p + slot[me]; const PackLayout& l = L[me];
    const int64_t n = mine->n_points, m = mine->n_obs;
    xchg_pack_points_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, s>>>(n, mine->d_xyz.p, mine->d_seed.p, mine->d_pos.p, mine->d_obs_off.p,
        (seed_global && mine->n_seeds > 0) ? d_sg.p : nullptr, seed_base, (float*)(b + l.o_xyz), (int64_t*)(b + l.o_key), (int64_t*)(b + l.o_off));
    CK(cudaGetLastError());
    if (m > 0) {
      CK(cudaMemcpyAsync(b + l.o_ov, mine->d_ov.p, 4 * (size_t)m, cudaMemcpyDeviceToDevice, s));
      CK(cudaMemcpyAsync(b + l.o_opl, mine->d_opl.p, 4 * (size_t)m, cudaMemcpyDeviceToDevice, s));
      CK(cudaMemcpyAsync(b + l.o_oseg, mine->d_oseg.p, 4 * (size_t)m, cudaMemcpyDeviceToDevice, s));
      CK(cudaMemcpyAsync(b + l.o_oxy, mine->d_oxy.p, 8 * (size_t)m, cudaMemcpyDeviceToDevice, s));
    }
  }
  Timer tx(s); tx.start();
  NCK(N.GroupStart());
  for (int r = 0; r < W; r++) if (L[r].bytes) NCK(N.Broadcast(recv.p + slot[r], recv.p + slot[r], L[r].bytes, ncclUint8, r, sc->comm, s));
  NCK(N.GroupEnd());
  tx.stop();
  for (int r = 0; r < W; r++) {
    const unsigned char* b = recv.p + slot[r]; const PackLayout& l = L[r];
    R.xyz[r] = (const float*)(b + l.o_xyz); R.key[r] = (const int64_t*)(b + l.o_key); R.off[r] = (const int64_t*)(b + l.o_off);
    R.ov[r] = (const int*)(b + l.o_ov); R.opl[r] = (const uint32_t*)(b + l.o_opl); R.oseg[r] = (const uint32_t*)(b + l.o_oseg); R.oxy[r] = (const float*)(b + l.o_oxy);
  }
  // 3) merge on the device: radix sort of the keys, observation offsets by scan, one gather pass
  std::unique_ptr<eg3d_points> p(new eg3d_points());
  p->sh = sc->sh; p->device = sc->device; p->stream = s; p->n_points = NP; p->n_obs = NO;
  for (int r = 0; r < W; r++) p->n_seeds += cnts[3 * r + 2];
  CK(p->d_xyz.alloc(3 * NP)); CK(p->d_seed.alloc(NP)); CK(p->d_pos.alloc(NP)); CK(p->d_obs_off.alloc(NP + 1));
  CK(p->d_ov.alloc(NO)); CK(p->d_opl.alloc(NO)); CK(p->d_oseg.alloc(NO)); CK(p->d_oxy.alloc(2 * NO));
  DBuf<unsigned long long> k0, k1; DBuf<int64_t> i0, i1; DBuf<unsigned char> tmp;
  CK(k0.alloc(NP)); CK(k1.alloc(NP)); CK(i0.alloc(NP)); CK(i1.alloc(NP));
  if (NP > 0) {
    xchg_keys_kernel<<<(unsigned)((NP + 255) / 256), 256, 0, s>>>(R, NP, k0.p, i0.p);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k0.p, k1.p, i0.p, i1.p, (int)NP, 0, 64, s);
    CK(tmp.alloc(tb));
    cub::DeviceRadixSort::SortPairs(tmp.p, tb, k0.p, k1.p, i0.p, i1.p, (int)NP, 0, 64, s);
  }
  xchg_counts_kernel<<<(unsigned)((NP + 1 + 255) / 256), 256, 0, s>>>(R, NP, i1.p, p->d_obs_off.p);
  {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, p->d_obs_off.p, p->d_obs_off.p, NP + 1, s);
    DBuf<unsigned char> t2; CK(t2.alloc(tb));
    cub::DeviceScan::ExclusiveSum(t2.p, tb, p->d_obs_off.p, p->d_obs_off.p, NP + 1, s);
    CK(cudaStreamSynchronize(s));
  }
  if (NP > 0) xchg_copy_kernel<<<(unsigned)((NP * 32 + 255) / 256), 256, 0, s>>>(R, NP, i1.p, k1.p, p->d_obs_off.p, p->d_xyz.p, p->d_seed.p, p->d_pos.p,
                                                                              p->d_ov.p, p->d_opl.p, p->d_oseg.p, p->d_oxy.p);
  CK(cudaGetLastError());
  tall.stop();
  CK(cudaStreamSynchronize(s));
  if (tm) { memset(tm, 0, sizeof *tm); tm->total_ms = tall.ms(); tm->scan_ms = tx.ms() /* the grouped broadcast alone */; tm->pack_ms = tall.ms() - tx.ms();
            tm->n_points = NP; tm->n_obs = NO; tm->n_seeds = p->n_seeds; tm->kernel_launches = 4; }
  *out = p.release();
  return EG3D_OK;
}

}  // extern "C"
