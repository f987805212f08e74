// eg3d_k1.cuh — K1: find_epipolar_correspondences (polyline_matching.cpp:45-73 -> polyline::intersect_line,
// polyline_graph_2d.cpp:312-327 -> intersect_segment_line, geometric_utilities.cpp:272-312).
//
// Sweep form (BASELINE configs 2-4): every seed against every segment of every other view.  One CTA owns one target
// view and a tile of 256 seeds; the view's staged segments (x1,y1,dx,dy), grouped by polyline into groups of <= 16 with
// an inflated bounding box each, stream through a double-buffered shared-memory ring filled by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx).  Each warp visits its 32 seeds one after the other: the lanes cull 32 group
// boxes per step against the epipolar line, then test the surviving groups' segments (two groups per step) with the
// reference predicate and ballot-compact the hits.  Two passes (count, exclusive scan, fill) give the reference's exact
// output order without atomics: hits of a (seed, view) pair are in ascending (polyline id, segment index).
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

// Sweep kernel.  One CTA = one target view x a tile of 256 (seed) entries; the view's staged segments and the bounding boxes of
// their groups of 32 stream through a double-buffered shared-memory ring (TMA bulk copies, mbarrier complete_tx).
// Each warp owns 32 seeds of the tile and visits them one after the other; for a seed the 32 lanes
//   1. test 32 group boxes at a time against the epipolar line (|a cx + b cy + c| <= |a| ex + |b| ey + slack) — a
//      conservative cull: a float hit needs the line within ~1e-3 px of the segment, the boxes are inflated by 0.05 px;
//   2. for every surviving group test its 32 segments at once with the reference predicate verbatim
//      (den != 0, t = -num/den, 0 <= t <= 1) and ballot-compact the hits, which keeps the reference's order
//      (polyline id ascending, segment ascending) because groups and lanes are visited in staging order.
#ifndef EG3D_K1_GROUP
#define EG3D_K1_GROUP 16
#endif
constexpr int K1_GROUP = EG3D_K1_GROUP;        // max segments per group (one group never spans two polylines)
constexpr int K1_GPR = 32 / K1_GROUP;          // groups handled per warp step (K1_GROUP lanes each)
constexpr int K1_GMAX = 4096 / K1_GROUP;       // max groups per chunk
constexpr int K1_SMEM_BYTES2 = K1_STAGES * (K1_CHUNK * 16 + K1_GMAX * 16 + K1_GMAX * 4) + K1_STAGES * 8;

// Which (seed, view) pairs a launch handles and where their results go.  The matching path never reads most of the
// hit lists the reference materialises (a seed that is not accepted only ever looks at its three selected views), so
// besides the full sweep there are two targeted forms:
//   list == null, sel == null : every seed x every view; result row of (seed, t) = seed * V + t        (K1 alone)
//   sel != null               : only the views sel[seed][0..2] of every seed; row = seed * 3 + k       (phase A)
//   list != null              : every view of the listed seeds; row = entry * V + t                    (phase B)
struct K1Work {
  const int* list;   // entry -> seed index (null: identity)
  int n;             // number of entries
  const int* sel;    // [n_seeds][3] wanted views (sel[3*seed] < 0: none), or null
};
enum { K1_COUNT = 0, K1_FILL = 1, K1_ANY = 2 };   // K1_ANY: only "is the list non-empty" (stops at a seed's first hit)

template <int MODE, bool ALL_PAIRS>   // ALL_PAIRS: list == null && sel == null resolved at compile time (the full sweep)
__global__ void __launch_bounds__(K1_THREADS) k1_sweep_kernel(const __grid_constant__ DevScene S, const __grid_constant__ K1Seeds seeds, const __grid_constant__ K1Work work,
                                                              int64_t* __restrict__ counts, unsigned char* __restrict__ flags,
                                                              const int64_t* __restrict__ off, eg3d_hit* __restrict__ hits) {
  constexpr bool FILL = MODE == K1_FILL;
  extern __shared__ __align__(128) unsigned char k1_smem[];
  float4 (*sbuf)[K1_CHUNK] = reinterpret_cast<float4 (*)[K1_CHUNK]>(k1_smem);
  float4 (*bbuf)[K1_GMAX] = reinterpret_cast<float4 (*)[K1_GMAX]>(k1_smem + sizeof(float4) * K1_STAGES * K1_CHUNK);
  uint32_t (*dbuf)[K1_GMAX] = reinterpret_cast<uint32_t (*)[K1_GMAX]>(k1_smem + sizeof(float4) * K1_STAGES * (K1_CHUNK + K1_GMAX));
  uint64_t* bars = reinterpret_cast<uint64_t*>(k1_smem + K1_STAGES * (K1_CHUNK * 16 + K1_GMAX * 16 + K1_GMAX * 4));
  const int tid = threadIdx.x, lane = tid & 31;
  const int t = blockIdx.y;                           // target view
  const int entry = blockIdx.x * K1_THREADS + tid;    // the entry whose line / counter this lane keeps
  const int ch0 = S.view_chunk_off[t];
  const int nchunks = S.view_chunk_off[t + 1] - ch0;

  // which pair is this lane's, is it wanted, and where does its result go
  bool valid = entry < work.n;
  int sidx = 0; size_t row = 0;
  if (ALL_PAIRS) { sidx = entry; row = (size_t)entry * S.V + t; }
  else if (valid) {
    sidx = work.list ? work.list[entry] : entry;
    if (work.sel) {
      const int* sl = work.sel + 3 * (size_t)sidx;
      const int k = sl[0] == t ? 0 : sl[1] == t ? 1 : sl[2] == t ? 2 : -1;
      if (k < 0 || sl[0] < 0) valid = false;
      row = (size_t)sidx * 3 + (size_t)(k < 0 ? 0 : k);
    } else row = (size_t)entry * S.V + t;
  }
  bool active = valid;
  int sv = -1; float2 p = make_float2(0.f, 0.f); float3 l = make_float3(0.f, 0.f, 0.f);
  if (valid) {
    sv = seeds.view[sidx]; p = seeds.xy[sidx];
    if (sv == t) active = false;
    else active = epiline(S, sv, t, p, l);
  }
  // a tile without a single pair to sweep (typical for the targeted forms) leaves before staging anything
  const bool any_active = ALL_PAIRS ? true : __syncthreads_or(active);   // the full sweep always has work
  if (any_active) {
    if (tid == 0) {
      for (int s = 0; s < K1_STAGES; s++) mbar_init(&bars[s], 1);
      mbar_fence_init();
    }
    __syncthreads();
  }
  int cnt = 0;
  int64_t obase = 0;
  if (FILL && valid) obase = off[row];
  const unsigned amask = __ballot_sync(0xffffffffu, active);

  auto issue = [&](int c) {
    const int4 ch = S.chunks[ch0 + c];               // (first segment, #segments, first group, #groups padded to x4)
    const uint32_t bytes = (uint32_t)ch.y * 16u, bbytes = (uint32_t)ch.w * 16u, dbytes = (uint32_t)ch.w * 4u;
    mbar_expect_tx(&bars[c % K1_STAGES], bytes + bbytes + dbytes);
    tma_bulk_g2s(&sbuf[c % K1_STAGES][0], S.seg + ch.x, bytes, &bars[c % K1_STAGES]);
    tma_bulk_g2s(&bbuf[c % K1_STAGES][0], S.grp_box + ch.z, bbytes, &bars[c % K1_STAGES]);
    tma_bulk_g2s(&dbuf[c % K1_STAGES][0], S.grp_desc + ch.z, dbytes, &bars[c % K1_STAGES]);
  };
  if (any_active && tid == 0 && nchunks > 0) issue(0);
  for (int c = 0; any_active && c < nchunks; c++) {
    if (tid == 0 && c + 1 < nchunks) issue(c + 1);
    mbar_wait(&bars[c % K1_STAGES], (uint32_t)((c / K1_STAGES) & 1));
    const float4* sb = sbuf[c % K1_STAGES];
    const float4* bb = bbuf[c % K1_STAGES];
    const uint32_t* db = dbuf[c % K1_STAGES];
    const int4 ch = S.chunks[ch0 + c];
    const int ng = ch.w;
    const uint2* sid = S.seg_id + ch.x;
    unsigned am = amask;
    if (MODE == K1_ANY) am &= ~__ballot_sync(0xffffffffu, cnt > 0);   // seeds that already have a hit are done
    while (am) {                                       // the warp's seeds, one after the other
      const int j = __ffs(am) - 1;
      am &= am - 1;
      const float la = __shfl_sync(0xffffffffu, l.x, j), lb = __shfl_sync(0xffffffffu, l.y, j), lc = __shfl_sync(0xffffffffu, l.z, j);
      const float aa = fabsf(la), ab = fabsf(lb);
      int cj = __shfl_sync(0xffffffffu, cnt, j);
      const int64_t oj = FILL ? __shfl_sync(0xffffffffu, obase, j) : 0;
      for (int gb = 0; gb < ng; gb += 32) {
        const int g = gb + lane;
        bool pass = false;
        if (g < ng) {
          const float4 bx = bb[g];                     // (cx, cy, ex, ey), inflated; padding groups have ex < 0
          pass = fabsf(la * bx.x + lb * bx.y + lc) <= aa * bx.z + ab * bx.w;
        }
        unsigned gm = __ballot_sync(0xffffffffu, pass);
        while (gm) {
          // K1_GPR surviving groups per step: lanes [k*K1_GROUP, (k+1)*K1_GROUP) take the k-th surviving group, so lane
          // order == (group ascending, segment ascending) == the reference's hit order
          const int which = lane / K1_GROUP, sub = lane % K1_GROUP;
          unsigned m2 = gm;
#pragma unroll
          for (int k = 0; k < K1_GPR - 1; k++) if (k < which) m2 &= m2 - 1;
          const unsigned pos = m2 ? (unsigned)(__ffs(m2) - 1) : 0xffffffffu;   // the (which+1)-th surviving group, if any
          bool hit = false; float hx = 0.f, hy = 0.f; int i = 0;
          if (pos != 0xffffffffu) {
            const uint32_t d = db[gb + (int)pos];                  // (offset in chunk << 6) | count
            i = (int)(d >> 6) + sub;
            if (sub < (int)(d & 63u)) {
              const float4 sg = sb[i];
              const float num = la * sg.x + lb * sg.y + lc;
              const float den = la * sg.z + lb * sg.w;
              if (den != 0) {
                const float tt = -num / den;
                if (tt >= 0 && tt <= 1) { hit = true; hx = sg.x + tt * sg.z; hy = sg.y + tt * sg.w; }
              }
            }
          }
          const unsigned hm = __ballot_sync(0xffffffffu, hit);
          if (FILL && hit) {
            const uint2 id = sid[i];
            eg3d_hit h; h.polyline = id.x; h.segment = id.y; h.x = hx; h.y = hy;
            hits[oj + cj + __popc(hm & ((1u << lane) - 1u))] = h;
          }
          cj += __popc(hm);
          if (MODE == K1_ANY && cj) break;
#pragma unroll
          for (int k = 0; k < K1_GPR; k++) gm &= gm - 1;          // drop the groups just handled
        }
        if (MODE == K1_ANY && cj) break;
      }
      if (lane == j) cnt = cj;
    }
    if (MODE == K1_ANY) {
      // every pair of the tile has its answer: drain the copy already in flight and stop staging
      if (!__syncthreads_or(active && cnt == 0)) {
        if (c + 1 < nchunks) mbar_wait(&bars[(c + 1) % K1_STAGES], (uint32_t)(((c + 1) / K1_STAGES) & 1));
        break;
      }
    } else __syncthreads();
  }
  if (valid) {
    if (sv == t) {  // the starting view holds the seed itself (polyline_matching.cpp:54-55)
      if (FILL) { eg3d_hit h; h.polyline = seeds.pl[sidx]; h.segment = seeds.seg[sidx]; h.x = p.x; h.y = p.y; hits[obase] = h; }
      cnt = 1;
    }
    if (MODE == K1_COUNT) counts[row] = cnt;
    if (MODE == K1_ANY) flags[row] = cnt > 0 ? 1 : 0;
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
