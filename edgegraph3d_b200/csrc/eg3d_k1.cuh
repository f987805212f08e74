// eg3d_k1.cuh — K1: find_epipolar_correspondences (polyline_matching.cpp:45-73 -> polyline::intersect_line,
// polyline_graph_2d.cpp:312-327 -> intersect_segment_line, geometric_utilities.cpp:272-312).
//
// Sweep form (BASELINE configs 2-4): every seed against every segment of every other view.  One CTA owns a tile of
// seeds (one seed per thread) and one target view; the view's staged segments (x1,y1,dx,dy) stream through a
// double-buffered shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier complete_tx), and every
// thread reads the same float4 per step (shared-memory broadcast), so the only per-test traffic is register math.
// Two passes (count, exclusive scan, fill) give the reference's exact output order without atomics:
// hits of a (seed, view) pair are in ascending (polyline id, segment index).
//
// Candidate form (reference semantics, configs 1/4): one thread per (seed, view) walks the few candidate polylines.
#pragma once
#include "eg3d_dev.cuh"

namespace eg3d {

constexpr int K1_THREADS = 256;
constexpr int K1_CHUNK = 2048;   // segments per stage: 32 KB
constexpr int K1_STAGES = 2;
constexpr int K1_SMEM_BYTES = K1_STAGES * K1_CHUNK * 16 + K1_STAGES * 8;

EG3D_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
EG3D_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
EG3D_D void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
EG3D_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
EG3D_D void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
EG3D_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
EG3D_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

struct K1Seeds {
  int n;
  const int* view; const uint32_t* pl; const uint32_t* seg; const float2* xy;
  const int* cand_set;  // may be null
};

// The exact reference predicate is `den != 0 && 0 <= fl(-num/den) <= 1`.  A division per test would dominate the loop,
// so lanes first apply a conservative filter that can only over-accept (|num| <= |den|(1+eps), opposite signs or a
// product that underflows), and the rare survivors evaluate the reference expression verbatim.
EG3D_D bool k1_prefilter(float num, float den) {
  return (fabsf(num) <= fabsf(den) * 1.000001f) && (num * den <= 1e-30f);
}

template <bool FILL>
__global__ void __launch_bounds__(K1_THREADS) k1_sweep_kernel(const __grid_constant__ DevScene S, const __grid_constant__ K1Seeds seeds, int view_lo,
                                                              int64_t* __restrict__ counts, const int64_t* __restrict__ off,
                                                              eg3d_hit* __restrict__ hits) {
  extern __shared__ __align__(128) unsigned char k1_smem[];
  float4 (*sbuf)[K1_CHUNK] = reinterpret_cast<float4 (*)[K1_CHUNK]>(k1_smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(k1_smem + sizeof(float4) * K1_STAGES * K1_CHUNK);
  const int tid = threadIdx.x;
  const int t = view_lo + blockIdx.y;                 // target view
  const int sidx = blockIdx.x * K1_THREADS + tid;     // my seed
  const int seg0 = S.view_seg_off[t];
  const int nseg = S.view_seg_off[t + 1] - seg0;
  const int nchunks = (nseg + K1_CHUNK - 1) / K1_CHUNK;

  if (tid == 0) {
    for (int s = 0; s < K1_STAGES; s++) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  bool active = sidx < seeds.n;
  int sv = -1; float2 p = make_float2(0.f, 0.f); float3 l = make_float3(0.f, 0.f, 0.f);
  if (active) {
    sv = seeds.view[sidx]; p = seeds.xy[sidx];
    if (sv == t) active = false;
    else active = epiline(S, sv, t, p, l);
  }
  int cnt = 0;
  int64_t obase = 0;
  if (FILL && sidx < seeds.n) obase = off[(size_t)sidx * S.V + t];
  const uint2* sid = S.seg_id + seg0;

  auto issue = [&](int c) {
    int n = min(K1_CHUNK, nseg - c * K1_CHUNK);
    uint32_t bytes = (uint32_t)n * 16u;
    mbar_expect_tx(&bars[c % K1_STAGES], bytes);
    tma_bulk_g2s(&sbuf[c % K1_STAGES][0], S.seg + seg0 + (size_t)c * K1_CHUNK, bytes, &bars[c % K1_STAGES]);
  };
  if (tid == 0 && nchunks > 0) issue(0);
  for (int c = 0; c < nchunks; c++) {
    if (tid == 0 && c + 1 < nchunks) issue(c + 1);
    mbar_wait(&bars[c % K1_STAGES], (uint32_t)((c / K1_STAGES) & 1));
    if (active) {
      const float4* sb = sbuf[c % K1_STAGES];
      const int n = min(K1_CHUNK, nseg - c * K1_CHUNK);
#pragma unroll 4
      for (int j = 0; j < n; j++) {
        float4 s = sb[j];
        float num = l.x * s.x + l.y * s.y + l.z;
        float den = l.x * s.z + l.y * s.w;
        if (k1_prefilter(num, den)) {
          if (den != 0) {
            float tt = -num / den;
            if (tt >= 0 && tt <= 1) {
              if (FILL) {
                uint2 id = sid[c * K1_CHUNK + j];
                eg3d_hit h; h.polyline = id.x; h.segment = id.y; h.x = s.x + tt * s.z; h.y = s.y + tt * s.w;
                hits[obase + cnt] = h;
              }
              cnt++;
            }
          }
        }
      }
    }
    __syncthreads();
  }
  if (sidx < seeds.n) {
    if (sv == t) {  // the starting view holds the seed itself (polyline_matching.cpp:54-55)
      if (FILL) { eg3d_hit h; h.polyline = seeds.pl[sidx]; h.segment = seeds.seg[sidx]; h.x = p.x; h.y = p.y; hits[obase] = h; }
      cnt = 1;
    }
    if (!FILL) counts[(size_t)sidx * S.V + t] = cnt;
  }
}

// center/seed_r2 (optional): the refpoint variant keeps a hit only within radius 3*|obs - seed| of that view's own
// observation of the SfM point (plg_edge_manager.cpp:191-205, :246-259)
struct K1Cand { const int64_t* off; const uint32_t* pl; const float2* center; const float* seed_r2; };

template <bool FILL>
__global__ void __launch_bounds__(256) k1_cand_kernel(const __grid_constant__ DevScene S, const __grid_constant__ K1Seeds seeds, const __grid_constant__ K1Cand cand, int64_t* __restrict__ counts,
                                                      const int64_t* __restrict__ off, eg3d_hit* __restrict__ hits) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)seeds.n * S.V) return;
  int sidx = (int)(idx / S.V), t = (int)(idx % S.V);
  int sv = seeds.view[sidx];
  float2 p = seeds.xy[sidx];
  int cnt = 0;
  int64_t obase = FILL ? off[idx] : 0;
  if (sv == t) {
    if (FILL) { eg3d_hit h; h.polyline = seeds.pl[sidx]; h.segment = seeds.seg[sidx]; h.x = p.x; h.y = p.y; hits[obase] = h; }
    cnt = 1;
  } else {
    float3 l;
    if (epiline(S, sv, t, p, l)) {
      int set = seeds.cand_set[sidx];
      size_t ci = (size_t)set * S.V + t;
      for (int64_t k = cand.off[ci]; k < cand.off[ci + 1]; k++) {
        uint32_t pl = cand.pl[k];
        int g = S.view_poly_off[t] + (int)pl;
        for (int s = S.poly_seg_off[g]; s < S.poly_seg_off[g + 1]; s++) {
          float4 sg = S.seg[s];
          float2 inter;
          if (isect_seg_line(sg.x, sg.y, sg.z, sg.w, l, inter)) {
            if (cand.center && !(sqdist2(cand.center[ci], inter) <= cand.seed_r2[sidx])) continue;
            if (FILL) { uint2 id = S.seg_id[s]; eg3d_hit h; h.polyline = id.x; h.segment = id.y; h.x = inter.x; h.y = inter.y; hits[obase + cnt] = h; }
            cnt++;
          }
        }
      }
    }
  }
  if (!FILL) counts[idx] = cnt;
}

}  // namespace eg3d
