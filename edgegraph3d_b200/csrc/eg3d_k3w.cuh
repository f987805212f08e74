// eg3d_k3w.cuh — phase B of K3 (view expansion, triangulation.cpp:742-830, 960-973; plg_matching.cpp:797-914, 1011-1058,
// 1345-1412) as a CTA-resident WAVEFRONT of per-seed state machines.
//
// Why: the one-warp-per-seed form (k3b_expand_kernel, eg3d_k3.cuh) walks ~70 KB of branchy code per (seed, view) with 16
// unsynchronised warps per SM; the SM-level instruction cache (~32 KB) then misses on almost every line outside the
// Gauss-Newton loop and the GPC-level refill path saturates (profiles/r01_k3b_icache.md): 25 % issue-active, time follows
// the number of refills.  Here ONE CTA of 16 warps per SM keeps up to KW_MAX_SEEDS accepted seeds resident, each a small
// state record in shared memory plus its scratch arena in global memory, and cycles through four code REGIONS separated
// by CTA barriers.  In a region every warp runs the same few KB of code on whatever seeds are in one of the region's
// states (seeds are handed out dynamically), so the region's lines are fetched once per cycle for ~all resident seeds:
//   R1  view entry: epipolar hits of the central point, exact-safe pruning, queueing of the surviving Gauss-Newton problems;
//       also seed set-up (chain materialisation) and chain emission
//   R2  candidate selection + the polyline walks of add_view_to_3dpoint_and_sides (walk_geo), queueing the neighbours' solves
//   R3  commit (append observations, chain extension by follow_big) and the projection loop over the remaining chain points
//   RG  every queued Gauss-Newton problem of every resident seed, KW_G lanes per problem, cameras staged in shared memory —
//       the only place the FP64 observation loop exists, run by all 16 warps at once
// A seed advances by (at most) one Gauss-Newton round trip per cycle; it never waits for the other seeds' views.  Each
// state handler is the corresponding piece of expand_view_epc / add_view_finish / expand_view_main (eg3d_k3.cuh), cut at
// the gn_group calls; solves that the sequential code would not have started (a neighbour list beyond its first failure)
// have no side effects, so results are identical.
#pragma once
#include "eg3d_k3.cuh"

namespace eg3d {

#ifndef EG3D_K3W_THREADS
#define EG3D_K3W_THREADS 512
#endif
constexpr int KW_THREADS = EG3D_K3W_THREADS;       // 16 warps, one CTA per SM
constexpr int KW_CTAS_PER_SM = 512 / EG3D_K3W_THREADS;
#ifndef EG3D_K3W_MAX_SEEDS
#define EG3D_K3W_MAX_SEEDS 64
#endif
constexpr int KW_MAX_SEEDS = EG3D_K3W_MAX_SEEDS;      // resident seeds per CTA (<= 64)
constexpr int KW_G = 8;               // lanes per Gauss-Newton problem in RG (fixed: results do not depend on scheduling)

enum : int {
  KS_FREE = 0, KS_IDLE,
  KS_VIEW, KS_PRUNE, KS_EMIT,                          // R1
  KS_EPC_WAIT, KS_EPC_CAND, KS_MAIN_WAIT, KS_AVF_WAIT, // R2
  KS_AVF_FINISH, KS_MAIN_NEXT                          // R3
};
EG3D_D int kw_region_of(int st) { return st == KS_IDLE ? 0 : (st <= KS_EMIT ? 1 : (st <= KS_AVF_WAIT ? 2 : 3)); }

struct SeedSt {   // one resident seed (shared memory; identical in every lane once loaded)
  int state, seed, v, nh;
  long long hrow, h0;
  int sel0, sel1, sel2;
  int len, nslots, central, overflow;
  int enext, nq, P; unsigned okmask;        // epipolar-hit stage: next hit to prune, queue length, problems in flight, untried successes
  int matched, iv0, iv1;
  int cur, last, hi;                        // projection loop
  int origin, lo, acur, stage, g1, g2, n1, n2; unsigned nd1, nd2;   // add_view_finish in flight
  unsigned p_pl, p_seg; float p_x, p_y; float Xc0, Xc1, Xc2;
  int gq_n;                                 // Gauss-Newton problems queued for RG
};

EG3D_D void kw_load(Ctx& c, const SeedSt& s) {
  c.seed = s.seed; c.sv = 0; c.hrow = s.hrow; c.sel[0] = s.sel0; c.sel[1] = s.sel1; c.sel[2] = s.sel2;
  c.len = s.len; c.nslots = s.nslots; c.central = s.central; c.overflow = s.overflow != 0;
}
EG3D_D void kw_save(const Ctx& c, SeedSt& s) {
  s.sel0 = c.sel[0]; s.sel1 = c.sel[1]; s.sel2 = c.sel[2];
  s.len = c.len; s.nslots = c.nslots; s.central = c.central; s.overflow = c.overflow ? 1 : 0;
}

#define KWP_T0() const long long kwp_t0 = (c.A->prof ? clock64() : 0)
#define KWP_ADD(slot, t0) do { if (c.A->prof && c.lane == 0) { const unsigned long long d_ = (unsigned long long)(clock64() - (t0)); atomicAdd(&c.A->prof[slot], d_); atomicAdd(&c.A->prof[(slot) + 1], 1ull); atomicMax(&c.A->prof[(slot) + 2], d_); } } while (0)
// ---------------------------------------------------------------------------------------------------------------
// R1: seed set-up / chain emission / view entry / pruning of the central point's epipolar hits (expand_view_epc, first half)
static __device__ __noinline__ void kw_r1_body(Ctx& c, SeedSt& s, int n_acc);
static __device__ __noinline__ void kw_r1(Ctx& c, SeedSt& s, int n_acc) {
  KWP_T0();
  kw_r1_body(c, s, n_acc);
  KWP_ADD(42, kwp_t0);
}
static __device__ __noinline__ void kw_r1_body(Ctx& c, SeedSt& s, int n_acc) {
  const DevScene& S = *c.S; const K3Args& A = *c.A;
  const int V = S.V, lane = c.lane;
  for (;;) {
    if (s.state == KS_EMIT) { emit_chain(c, s.seed); s.state = KS_FREE; }
    if (s.state == KS_FREE) {
      int ri = 0;
      if (lane == 0) ri = atomicAdd(A.work_counter, 1);
      ri = __shfl_sync(0xffffffffu, ri, 0);
      if (ri >= n_acc) { s.state = KS_IDLE; s.gq_n = 0; return; }
      const int rank = A.pa_order ? A.pa_order[ri] : ri;
      const PaRec& r = A.pa_recs[rank];
      s.seed = r.seed; c.seed = r.seed;
      s.hrow = c.hrow = (long long)(A.b_compact ? rank : r.seed) * V;
      c.len = 0; c.nslots = 0; c.central = 0; c.overflow = false;
      c.sel[0] = c.sel[1] = c.sel[2] = -1;
      bool live = false;
      if (r.fn1 >= 0) { seed_phase_b(c, r, A.pa_pool + r.pool_off, A.pa_pool + r.pool_off + r.fn1); live = !c.overflow; }
      s.v = 0; s.gq_n = 0;
      s.state = live ? KS_VIEW : KS_EMIT;
      continue;
    }
    if (s.state == KS_VIEW) {
      if (c.overflow) { s.state = KS_EMIT; continue; }
      while (s.v < V && (s.v == c.sel[0] || s.v == c.sel[1] || s.v == c.sel[2])) s.v++;
      if (s.v >= V) { s.state = KS_EMIT; continue; }
      s.h0 = A.hit_off_b[c.hrow + s.v];
      s.nh = (int)(A.hit_off_b[c.hrow + s.v + 1] - s.h0);
      s.enext = 0; s.nq = 0; s.matched = 0; s.iv0 = 0; s.iv1 = 0; s.okmask = 0; s.P = 0;
      s.state = KS_PRUNE;
    }
    break;
  }
  if (s.state != KS_PRUNE) return;
  // triangulation.cpp:753-768 with the exact-safe 2-view bound (pair_cannot_fit) in front of the solves
  const int v = s.v;
  const eg3d_hit* epcs = A.hits_b + s.h0;
  int* hq = c.w.tq;
  const int cslot = slot_of(c, c.central);
  const int n = c.w.snobs[cslot];
  const size_t cb = (size_t)cslot * c.w.oc;
  {
    const double Tp = prune_radius(S.prm, n + 1);
    const int probe[4] = {0, n / 3, (2 * n) / 3, n - 1};
    while (s.nq < 32 && s.enext < s.nh) {
      const int e = s.enext + lane;
      bool pass = false;
      if (e < s.nh) {
        const float2 hp = make_float2(epcs[e].x, epcs[e].y);
        pass = true;
        for (int k = 0; k < 4 && pass; k++) {
          const int i = probe[k];
          const int vi = c.w.ov[cb + i];
          if (vi == v) continue;
          const size_t fi = (size_t)vi * S.V + v;
          if (pair_cannot_fit(S.Fp + fi * 9, S.Fph[fi], make_float2(c.w.ox[cb + i], c.w.oy[cb + i]), hp, Tp)) pass = false;
        }
      }
      const unsigned pm = __ballot_sync(0xffffffffu, pass);
      if (pass) hq[s.nq + __popc(pm & ((1u << lane) - 1u))] = e;
      s.nq += __popc(pm);
      s.enext += 32;
    }
    __syncwarp();
  }
  if (s.nq == 0) {                          // no (further) candidate: the view goes on with the projection loop
    s.cur = 0; s.last = -1; s.gq_n = 0;
    s.state = KS_MAIN_NEXT;
    return;
  }
  const int P = s.nq < 32 ? s.nq : 32;
  int e = 0, tail = 0;
  const int rest = s.nq - P;
  if (lane < P) e = hq[lane];
  if (lane < rest) tail = hq[P + lane];
  __syncwarp();
  if (lane < rest) hq[lane] = tail;
  if (lane < P) {
    const eg3d_hit h = epcs[e];
    c.w.pe[lane] = e;
    GnQ q; q.slot = cslot; q.ex = h.x; q.ey = h.y; q.pad = 0;
    c.w.gq[lane] = q;
  }
  __syncwarp();
  s.nq = rest; s.P = P; s.gq_n = P;
  s.state = KS_EPC_WAIT;
}

// ---------------------------------------------------------------------------------------------------------------
// R2: pick the next candidate, run the polyline walks of add_view_to_3dpoint_and_sides_plgp_matches_vector
// (plg_matching.cpp:1345-1412 -> :1011-1058 -> :866-914) and queue the neighbours' warm-started solves.
// The order of attempts of add_view_finish is kept as a stage number:
//   0: start side along pl.start (+ end side along pl.end, solved in the same batch)      1: start side along pl.end
//   2: end side along pl.start (after 1 succeeded)      3: end side along pl.end (after 1 failed)      4: end side along pl.start
static __device__ __noinline__ void kw_r2_body(Ctx& c, SeedSt& s);
static __device__ __noinline__ void kw_r2(Ctx& c, SeedSt& s) {
  KWP_T0();
  kw_r2_body(c, s);
  KWP_ADD(39, kwp_t0);
}
static __device__ __noinline__ void kw_r2_body(Ctx& c, SeedSt& s) {
  const DevScene& S = *c.S; const K3Args& A = *c.A;
  const int lane = c.lane;
  const int v = s.v;
  bool resumed = false;
  if (s.state == KS_EPC_WAIT) {
    GnR r; r.X[0] = r.X[1] = r.X[2] = 0.f; r.ok = 0;
    if (lane < s.P) { r = c.w.gr[lane]; c.w.er[lane] = r; }     // kept aside: the walks of a candidate reuse gq / gr
    s.okmask = __ballot_sync(0xffffffffu, r.ok != 0);
    __syncwarp();
    s.gq_n = 0;
    s.state = KS_EPC_CAND;
  }
  if (s.state == KS_EPC_CAND) {
    if (s.okmask == 0) { s.state = KS_PRUNE; return; }
    const int b = __ffs(s.okmask) - 1;
    s.okmask &= s.okmask - 1;
    const eg3d_hit h = (A.hits_b + s.h0)[c.w.pe[b]];
    const GnR r = c.w.er[b];
    s.p_pl = h.polyline; s.p_seg = h.segment; s.p_x = h.x; s.p_y = h.y;
    s.Xc0 = r.X[0]; s.Xc1 = r.X[1]; s.Xc2 = r.X[2];
    s.origin = 0; s.lo = 0; s.acur = c.central; s.hi = c.len;
  } else if (s.state == KS_MAIN_WAIT) {
    const GnR r = c.w.gr[0];
    s.gq_n = 0;
    if (!r.ok) { s.cur++; s.state = KS_MAIN_NEXT; return; }
    s.Xc0 = r.X[0]; s.Xc1 = r.X[1]; s.Xc2 = r.X[2];
    s.origin = 1; s.lo = s.last + 1; s.acur = s.cur;
  } else if (s.state == KS_AVF_WAIT) {
    resumed = true;
  } else return;

  Plg p; p.pl = s.p_pl; p.seg = s.p_seg; p.c = make_float2(s.p_x, s.p_y);
  const Pl pl = get_pl(S, v, p.pl);
  const int lo = s.lo, cur = s.acur, hi = s.hi;
  int stage;
  if (!resumed) {
    s.n1 = 0; s.n2 = 0; s.nd1 = 0; s.nd2 = 0; s.g1 = 0; s.g2 = 0;
    if (!(cur > lo)) { s.state = KS_AVF_FINISH; return; }
    stage = 0;
  } else stage = s.stage;
  for (;;) {
    int keep1 = 0, keep2 = 0;
    if (!resumed) {
      int g1 = 0, g2 = 0;
      if (stage == 0) { g1 = walk_geo(c, v, p, pl.start, true, lo, cur, hi, c.w.tmp1); if (g1 > 0 && cur < hi) g2 = walk_geo(c, v, p, pl.end, false, lo, cur, hi, c.w.tmp2); }
      else if (stage == 1) g1 = walk_geo(c, v, p, pl.end, true, lo, cur, hi, c.w.tmp1);
      else if (stage == 3) g2 = walk_geo(c, v, p, pl.end, false, lo, cur, hi, c.w.tmp2);
      else g2 = walk_geo(c, v, p, pl.start, false, lo, cur, hi, c.w.tmp2);      // stages 2 and 4
      if (g1 + g2 > 0) {
        for (int q = lane; q < g1 + g2; q += 32) {
          const bool s1 = q < g1;
          const int k = s1 ? q : q - g1;
          const NTmp& t = s1 ? c.w.tmp1[k] : c.w.tmp2[k];
          GnQ d; d.slot = slot_of(c, s1 ? cur - 1 - k : cur + 1 + k); d.ex = t.cx; d.ey = t.cy; d.pad = 0;
          c.w.gq[q] = d;
        }
        __syncwarp();
        s.g1 = g1; s.g2 = g2; s.stage = stage; s.gq_n = g1 + g2;
        s.state = KS_AVF_WAIT;
        return;
      }
    } else {
      // results of the batch queued for `stage`: each side's list is cut at its first failed solve
      const int g1 = s.g1, g2 = s.g2;
      keep1 = g1; keep2 = g2;
      for (int base = 0; base < g1 + g2; base += 32) {
        const int q = base + lane;
        const bool fail = q < g1 + g2 ? (c.w.gr[q].ok == 0) : false;
        const unsigned f1 = __ballot_sync(0xffffffffu, fail && q < g1), f2 = __ballot_sync(0xffffffffu, fail && q >= g1);
        if (f1) keep1 = min(keep1, base + __ffs(f1) - 1);
        if (f2) keep2 = min(keep2, base + __ffs(f2) - 1 - g1);
      }
      for (int q = lane; q < g1 + g2; q += 32) {
        const bool s1 = q < g1;
        const int k = s1 ? q : q - g1;
        if (k < (s1 ? keep1 : keep2)) {
          const GnR r = c.w.gr[q];
          NTmp& t = s1 ? c.w.tmp1[k] : c.w.tmp2[k];
          t.X[0] = r.X[0]; t.X[1] = r.X[1]; t.X[2] = r.X[2];
        }
      }
      __syncwarp();
      s.gq_n = 0;
      resumed = false;
    }
    bool done = false;
    switch (stage) {
      case 0: s.n1 = keep1; if (keep1 > 0) { s.n2 = keep2; s.nd1 = pl.start; s.nd2 = pl.end; done = true; } else stage = 1; break;
      case 1: s.n1 = keep1;
              if (keep1 > 0) { s.nd1 = pl.end; s.nd2 = pl.start; if (cur < hi) stage = 2; else done = true; }
              else if (cur < hi) stage = 3; else done = true;
              break;
      case 2: s.n2 = keep2; done = true; break;
      case 3: s.n2 = keep2; if (keep2 > 0) { s.nd2 = pl.end; s.nd1 = pl.start; done = true; } else stage = 4; break;
      default: s.n2 = keep2; if (keep2 > 0) { s.nd2 = pl.start; s.nd1 = pl.end; } done = true; break;
    }
    if (done) { s.state = KS_AVF_FINISH; return; }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// R3: the commit half of add_view_finish and the projection loop of expand_allpoints_to_other_view_using_plmap
// (triangulation.cpp:770-830).
static __device__ __noinline__ bool kw_avf_commit(Ctx& c, int v, const Plg& p, const float Xc[3], int cur, int n1, int n2, uint32_t nd1, uint32_t nd2, int& ns, int& ne) {
  if (cur > 0 && n1 == 0) return false;           // SWITCH_PLG_MATCHING_ADDPOINT_BOTHDIR_ONE (:1363-1368)
  if (cur < c.len - 1 && n2 == 0) return false;
  __syncwarp();
  slot_append(c, slot_of(c, cur), v, p.pl, p.seg, p.c.x, p.c.y, Xc);
  for (int q = c.lane; q < n1 + n2; q += 32) {     // the neighbours sit in distinct slots: one append per lane
    const bool s1 = q < n1;
    const NTmp t = s1 ? c.w.tmp1[q] : c.w.tmp2[q - n1];
    const int slot = slot_of(c, s1 ? cur - 1 - q : cur + 1 + (q - n1));
    const int n = c.w.snobs[slot];
    if (n >= c.w.oc) c.overflow = true;
    else {
      const size_t b = (size_t)slot * c.w.oc + n;
      c.w.ov[b] = v; c.w.opl[b] = p.pl; c.w.oseg[b] = t.seg; c.w.ox[b] = t.cx; c.w.oy[b] = t.cy;
      c.w.snobs[slot] = n + 1;
      c.w.sX[3 * slot] = t.X[0]; c.w.sX[3 * slot + 1] = t.X[1]; c.w.sX[3 * slot + 2] = t.X[2];
    }
  }
  c.overflow = __any_sync(0xffffffffu, c.overflow);
  __syncwarp();
  if (c.overflow) return false;
  ns = n1; ne = n2;
  if (n1 > 0 && n1 == cur) {
    if (c.lane == 0) c.w.sdirs[v] = nd1;
    __syncwarp();
    const long long tf = (c.A->prof ? clock64() : 0);
    const int added = follow_big(c, c.w.sdirs, true);
    KWP_ADD(36, tf);
    ns += added; cur += added;
  }
  if (n2 > 0 && n2 == (c.len - cur - 1)) {
    if (c.lane == 0) c.w.edirs[v] = nd2;
    __syncwarp();
    const long long tf = (c.A->prof ? clock64() : 0);
    ne += follow_big(c, c.w.edirs, false);
    KWP_ADD(36, tf);
  }
  return true;
}

static __device__ __noinline__ void kw_r3(Ctx& c, SeedSt& s) {
  const DevScene& S = *c.S;
  const int v = s.v;
  KWP_T0();
  if (s.state == KS_AVF_FINISH) {
    Plg p; p.pl = s.p_pl; p.seg = s.p_seg; p.c = make_float2(s.p_x, s.p_y);
    const float Xc[3] = {s.Xc0, s.Xc1, s.Xc2};
    int ns = 0, ne = 0;
    const bool ok = kw_avf_commit(c, v, p, Xc, s.acur, s.n1, s.n2, s.nd1, s.nd2, ns, ne);
    if (s.origin == 0) {
      const int cc = s.acur;
      if (ok) {
        s.matched = 1;
        if (ns > cc) { c.central = ns; s.iv0 = 0; s.iv1 = ns + ne; }
        else { s.iv0 = cc - ns; s.iv1 = cc + ne; }
      }
      if (c.overflow) { s.v++; s.state = KS_VIEW; return; }
      if (!ok) { s.state = KS_EPC_CAND; return; }
      s.cur = 0; s.last = -1;
    } else {
      if (ok) {
        if (ns > s.cur) { c.central = ns; s.cur = ns + ne; }
        else s.cur = s.cur + ne;
        s.last = s.cur;
      }
      if (c.overflow) { s.v++; s.state = KS_VIEW; return; }
      s.cur++;
    }
    s.state = KS_MAIN_NEXT;
    KWP_ADD(30, kwp_t0);
  }
  if (s.state != KS_MAIN_NEXT) return;
  const long long kwp_t1 = (c.A->prof ? clock64() : 0);
  for (; s.cur < c.len; s.cur++) {
    if (s.matched && s.cur == s.iv0) { s.cur = s.iv1; s.last = s.iv1; continue; }
    const int slot = slot_of(c, s.cur);
    const float2 q = project(S.P + 12 * v, c.w.sX[3 * slot], c.w.sX[3 * slot + 1], c.w.sX[3 * slot + 2]);
    uint32_t pl_id;
    if (!grid_unique_warp(S.g_expand, v, S.width, S.height, q, pl_id, c.lane)) continue;
    const Pl pl = get_pl(S, v, pl_id);
    Plg init; init.pl = pl_id;
    if (pl_distancesq_warp(pl, q, init.seg, init.c, c.lane) > S.prm.max_proj_distsq_expand) break;   // abandons the view
    s.hi = s.matched ? (s.cur <= s.iv0 ? s.iv0 : c.len) : c.len;
    s.p_pl = init.pl; s.p_seg = init.seg; s.p_x = init.c.x; s.p_y = init.c.y;
    if (c.lane == 0) { GnQ d; d.slot = slot; d.ex = init.c.x; d.ey = init.c.y; d.pad = 0; c.w.gq[0] = d; }
    __syncwarp();
    s.gq_n = 1;
    s.state = KS_MAIN_WAIT;
    KWP_ADD(33, kwp_t1);
    return;
  }
  s.v++; s.gq_n = 0;
  s.state = KS_VIEW;
  KWP_ADD(33, kwp_t1);
}

// Lanes per problem for a batch of n queued problems: a function of the batch only, so results do not depend on scheduling.
EG3D_D int kw_batch_width(int n) { return n >= 4 ? 8 : (n >= 2 ? 16 : 32); }

static __device__ __noinline__ void kw_gn_chunk(const DevScene& S, const WS& w, int vnew, int k0, int n_batch, int G, const double* __restrict__ P64, int lane) {
  const int sub = lane & (G - 1), grp = lane / G;
  const int k = k0 + grp;
  const bool active = k < n_batch;
  GnQ d; d.slot = 0; d.ex = 0.f; d.ey = 0.f; d.pad = 0;
  int n = 0;
  double X0 = 0, X1 = 0, X2 = 0;
  if (active) {
    d = w.gq[k];
    n = w.snobs[d.slot];
    X0 = w.sX[3 * d.slot]; X1 = w.sX[3 * d.slot + 1]; X2 = w.sX[3 * d.slot + 2];
  }
  const size_t ob = (size_t)d.slot * w.oc;
  const int* ov = w.ov + ob; const float* ox = w.ox + ob; const float* oy = w.oy + ob;
  const int ntot = n + 1;
  const int max_iters = S.prm.gn_max_iters;
  double last_mse = 0;
  bool running = active, failed = false;
  for (int it = 0; it < max_iters; it++) {
    if (!__any_sync(0xffffffffu, running)) break;
    GnAcc a = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (running) {
#pragma unroll 1
      for (int i = sub; i < ntot; i += G) {
        int vv; float px, py;
        if (i < n) { vv = ov[i]; px = ox[i]; py = oy[i]; } else { vv = vnew; px = d.ex; py = d.ey; }
        gn_accumulate_fast(P64 + 12 * vv, px, py, X0, X1, X2, a);
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
    if (running) {
      const double cur = a.mse / (ntot * 2);
      if (fabs(cur - last_mse) < S.prm.gn_stop) running = false;
      else {
        last_mse = cur;
        const double H[9] = {a.h00, a.h01, a.h02, a.h01, a.h11, a.h12, a.h02, a.h12, a.h22};
        const double dt = det3d(H);
        if (dt < S.prm.gn_det_min) { running = false; failed = true; }
        else {
          double Hi[9]; inv3d(H, dt, Hi);
          X0 += Hi[0] * a.g0 + Hi[1] * a.g1 + Hi[2] * a.g2;
          X1 += Hi[3] * a.g0 + Hi[4] * a.g1 + Hi[5] * a.g2;
          X2 += Hi[6] * a.g0 + Hi[7] * a.g1 + Hi[8] * a.g2;
        }
      }
    }
  }
  if (active && sub == 0) {
    GnR r; r.X[0] = (float)X0; r.X[1] = (float)X1; r.X[2] = (float)X2; r.ok = (!failed && last_mse < S.prm.gn_accept_mse) ? 1 : 0;
    w.gr[k] = r;
  }
}


// Solves queued problems until none is unclaimed: warps call this whenever they have no control task, so the FP64-bound
// solves fill the issue slots the latency-bound control regions leave empty.  pending[i] > 0 publishes seed i's batch
// (written last by the warp that queued it); cursor[i] hands out chunks; pending[i] drops to 0 when the batch is done.
static __device__ __noinline__ void kw_gn_fill(const DevScene& S, const K3Args& A, const SeedSt* st, int nS, int* pending, int* cursor, const double* __restrict__ P64, unsigned char* arena0, int lane, int rot) {
  volatile int* vpend = pending; volatile int* vcur = cursor;
  const volatile SeedSt* vst = st;
  for (;;) {
    const int ia = (lane + rot) & 63, ib = (lane + 32 + rot) & 63;
    const unsigned ma = __ballot_sync(0xffffffffu, ia < nS && vpend[ia] > 0 && vcur[ia] < vst[ia].gq_n);
    const unsigned mb = __ballot_sync(0xffffffffu, ib < nS && vpend[ib] > 0 && vcur[ib] < vst[ib].gq_n);
    if ((ma | mb) == 0) return;
    const int pick = ma ? __ffs(ma) - 1 : 32 + __ffs(mb) - 1;
    const int i = (pick + rot) & 63;
    __threadfence_block();
    const int nb = vst[i].gq_n;
    const int G = kw_batch_width(nb), ppw = 32 / G;
    int k0 = 0;
    if (lane == 0) k0 = atomicAdd(&cursor[i], ppw);
    k0 = __shfl_sync(0xffffffffu, k0, 0);
    if (k0 >= nb) continue;
    const WS w = make_ws(arena0 + (size_t)i * A.scratch_per_warp, S.V, A.capf, A.capc, A.oc);
    kw_gn_chunk(S, w, vst[i].v, k0, nb, G, P64, lane);
    __syncwarp();
    if (lane == 0) { __threadfence_block(); atomicSub(&pending[i], min(ppw, nb - k0)); }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// RG: every queued problem of every resident seed.  em_GaussNewton (triangulation.cpp:105-176) exactly as gn_group
// (eg3d_dev.cuh) evaluates it — one reciprocal of the depth, explicit FMAs, the ten sums butterfly-reduced over the
// problem's KW_G lanes — with the widened cameras read from shared memory.
static __device__ __noinline__ void kw_rg(const DevScene& S, const K3Args& A, const SeedSt* st, int nS, int* ctr, const double* __restrict__ P64, unsigned char* arena0, int lane) {
  constexpr int G = KW_G, PPW = 32 / G;
  int i0 = lane < nS ? st[lane].gq_n : 0, i1 = lane + 32 < nS ? st[lane + 32].gq_n : 0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, i0, o), b = __shfl_up_sync(0xffffffffu, i1, o);
    if (lane >= o) { i0 += a; i1 += b; }
  }
  i1 += __shfl_sync(0xffffffffu, i0, 31);
  const int T = __shfl_sync(0xffffffffu, i1, 31);
  const int sub = lane & (G - 1), grp = lane / G;
  const int max_iters = S.prm.gn_max_iters;
  const double gn_stop = S.prm.gn_stop, det_min = S.prm.gn_det_min, accept = S.prm.gn_accept_mse;
  for (;;) {
    int ch = 0;
    if (lane == 0) ch = atomicAdd(ctr, 1);
    ch = __shfl_sync(0xffffffffu, ch, 0);
    if (ch * PPW >= T) break;
    const int q = ch * PPW + grp;
    const bool active = q < T;
    int idx = 0;
#pragma unroll
    for (int g = 0; g < PPW; g++) {
      const int qq = ch * PPW + g;
      const int ii = __popc(__ballot_sync(0xffffffffu, i0 <= qq)) + __popc(__ballot_sync(0xffffffffu, i1 <= qq));
      if (grp == g) idx = ii;
    }
    const int e0 = __shfl_sync(0xffffffffu, i0, (idx - 1) & 31), e1 = __shfl_sync(0xffffffffu, i1, (idx - 33) & 31);
    if (!active) idx = 0;
    const int excl = idx == 0 ? 0 : (idx <= 32 ? e0 : e1);
    const int k = active ? q - excl : 0;
    const WS w = make_ws(arena0 + (size_t)idx * A.scratch_per_warp, S.V, A.capf, A.capc, A.oc);
    const int vnew = st[idx].v;
    GnQ d; d.slot = 0; d.ex = 0.f; d.ey = 0.f; d.pad = 0;
    int n = 0;
    double X0 = 0, X1 = 0, X2 = 0;
    if (active) {
      d = w.gq[k];
      n = w.snobs[d.slot];
      X0 = w.sX[3 * d.slot]; X1 = w.sX[3 * d.slot + 1]; X2 = w.sX[3 * d.slot + 2];
    }
    const size_t ob = (size_t)d.slot * w.oc;
    const int* ov = w.ov + ob; const float* ox = w.ox + ob; const float* oy = w.oy + ob;
    const int ntot = n + 1;
    double last_mse = 0;
    bool running = active, failed = false;
    for (int it = 0; it < max_iters; it++) {
      if (!__any_sync(0xffffffffu, running)) break;
      GnAcc a = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      if (running) {
#pragma unroll 1
        for (int i = sub; i < ntot; i += G) {
          int vv; float px, py;
          if (i < n) { vv = ov[i]; px = ox[i]; py = oy[i]; } else { vv = vnew; px = d.ex; py = d.ey; }
          gn_accumulate_fast(P64 + 12 * vv, px, py, X0, X1, X2, a);
        }
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
        const double cur = a.mse / (ntot * 2);
        if (fabs(cur - last_mse) < gn_stop) running = false;
        else {
          last_mse = cur;
          const double H[9] = {a.h00, a.h01, a.h02, a.h01, a.h11, a.h12, a.h02, a.h12, a.h22};
          const double dt = det3d(H);
          if (dt < det_min) { running = false; failed = true; }
          else {
            double Hi[9]; inv3d(H, dt, Hi);
            X0 += Hi[0] * a.g0 + Hi[1] * a.g1 + Hi[2] * a.g2;
            X1 += Hi[3] * a.g0 + Hi[4] * a.g1 + Hi[5] * a.g2;
            X2 += Hi[6] * a.g0 + Hi[7] * a.g1 + Hi[8] * a.g2;
          }
        }
      }
    }
    if (active && sub == 0) {
      GnR r; r.X[0] = (float)X0; r.X[1] = (float)X1; r.X[2] = (float)X2; r.ok = (!failed && last_mse < accept) ? 1 : 0;
      w.gr[k] = r;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KW_THREADS, KW_CTAS_PER_SM) k3w_expand_kernel(const __grid_constant__ DevScene S, const __grid_constant__ K3Args A, int nS, int p64_in_smem, int merge, int fill) {
  extern __shared__ double kw_smem_p64[];
  __shared__ SeedSt st[KW_MAX_SEEDS];
  __shared__ int ctr[4];
  __shared__ int pending[KW_MAX_SEEDS], cursor[KW_MAX_SEEDS];
  const int lane = threadIdx.x & 31;
  const int rot = ((threadIdx.x >> 5) * 4) & 63;
  const int n_acc = (int)A.pa_counters[0];
  unsigned char* arena0 = A.scratch + (size_t)blockIdx.x * nS * A.scratch_per_warp;
  const double* P64 = S.P64;
  if (p64_in_smem) {
    for (int i = threadIdx.x; i < S.V * 12; i += blockDim.x) kw_smem_p64[i] = S.P64[i];
    P64 = kw_smem_p64;
  }
  for (int i = threadIdx.x; i < KW_MAX_SEEDS; i += blockDim.x) { st[i].state = i < nS ? KS_FREE : KS_IDLE; st[i].gq_n = 0; st[i].v = 0; pending[i] = 0; cursor[i] = 0; }
  if (threadIdx.x < 4) ctr[threadIdx.x] = 0;
  __syncthreads();
  Ctx c;
  c.S = &S; c.A = &A; c.lane = lane;
#ifdef EG3D_K3_PROFILE
  for (int k = 0; k < 32; k++) c.pc[k] = 0;
#endif
  // Cycle: RG, R2, R3, R1.  Every consumer of Gauss-Newton results is an R2 state, so RG runs right before R2 and serves
  // the problems queued by R2, R3 and R1 of the previous cycle.  ctr[k]: hand-out counter of region k (0 = RG); each
  // region zeroes the counter of the region that follows it.
  long long t_reg[4] = {0, 0, 0, 0}, t_last = clock64(), n_cycles = 0, n_tasks[4] = {0, 0, 0, 0};
  for (;;) {
    if (threadIdx.x == 0) ctr[2] = 0;
    if (A.prof && threadIdx.x == 0) { int T = 0; for (int i = 0; i < nS; i++) T += st[i].gq_n; n_tasks[0] += T; }
    if (fill) kw_gn_fill(S, A, st, nS, pending, cursor, P64, arena0, lane, rot);
    else kw_rg(S, A, st, nS, &ctr[0], P64, arena0, lane);
    __syncthreads();
    if (A.prof && threadIdx.x == 0) { const long long t = clock64(); t_reg[0] += t - t_last; t_last = t; n_cycles++; }
    if (merge) {
      // one control region: every seed runs its handlers back to back until it queues Gauss-Newton problems (or ends)
      if (threadIdx.x == 0) ctr[0] = 0;
      if (A.prof && threadIdx.x == 0) { int T = 0; for (int i = 0; i < nS; i++) T += st[i].state != KS_IDLE; n_tasks[1] += T; }
      for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(&ctr[2], 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nS) break;
        if (st[i].state == KS_IDLE) continue;
        SeedSt s = st[i];
        c.w = make_ws(arena0 + (size_t)i * A.scratch_per_warp, S.V, A.capf, A.capc, A.oc);
        kw_load(c, s);
        for (;;) {
          const int region = kw_region_of(s.state);
          if (region == 1) kw_r1(c, s, n_acc);
          else if (region == 2) kw_r2(c, s);
          else if (region == 3) kw_r3(c, s);
          if (region == 0 || s.gq_n > 0 || s.state == KS_IDLE) break;
        }
        kw_save(c, s);
        __syncwarp();
        if (lane == 0) { st[i] = s; if (fill && s.gq_n > 0) { cursor[i] = 0; __threadfence_block(); pending[i] = s.gq_n; } }
        __syncwarp();
      }
    } else {
#pragma unroll 1
    for (int step = 0; step < 3; step++) {
      const int region = step == 0 ? 2 : (step == 1 ? 3 : 1);
      if (threadIdx.x == 0) ctr[region == 2 ? 3 : (region == 3 ? 1 : 0)] = 0;
      if (A.prof && threadIdx.x == 0) { int T = 0; for (int i = 0; i < nS; i++) T += kw_region_of(st[i].state) == region; n_tasks[region] += T; }
      for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(&ctr[region], 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nS) break;
        if (kw_region_of(st[i].state) != region) continue;
        SeedSt s = st[i];
        c.w = make_ws(arena0 + (size_t)i * A.scratch_per_warp, S.V, A.capf, A.capc, A.oc);
        kw_load(c, s);
        if (region == 1) kw_r1(c, s, n_acc);
        else if (region == 2) kw_r2(c, s);
        else kw_r3(c, s);
        kw_save(c, s);
        __syncwarp();
        if (lane == 0) { st[i] = s; if (fill && s.gq_n > 0) { cursor[i] = 0; __threadfence_block(); pending[i] = s.gq_n; } }
        __syncwarp();
      }
      if (fill) kw_gn_fill(S, A, st, nS, pending, cursor, P64, arena0, lane, rot);
      if (step < 2) {
        __syncthreads();
        if (A.prof && threadIdx.x == 0) { const long long t = clock64(); t_reg[region] += t - t_last; t_last = t; }
      }
    }
    }
    bool busy = false;
    for (int i = threadIdx.x; i < nS; i += blockDim.x) busy |= st[i].state != KS_IDLE;
    const bool go = __syncthreads_or(busy);
    if (A.prof && threadIdx.x == 0) { const long long t = clock64(); t_reg[1] += t - t_last; t_last = t; }
    if (!go) break;
  }
  if (A.prof && threadIdx.x == 0) {   // EG3D_K3_PROF: per-region clock cycles, cycles and tasks, summed over the CTAs ([16..27]); slowest CTA in [28]
    for (int k = 0; k < 4; k++) { atomicAdd(&A.prof[16 + k], (unsigned long long)t_reg[k]); atomicAdd(&A.prof[20 + k], (unsigned long long)n_tasks[k]); }
    atomicAdd(&A.prof[24], (unsigned long long)n_cycles);
    atomicMax(&A.prof[25], (unsigned long long)n_cycles);
    atomicMax(&A.prof[28], (unsigned long long)(t_reg[0] + t_reg[1] + t_reg[2] + t_reg[3]));
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Barrier-free form: the same handlers and the same Gauss-Newton evaluation, but the 16 warps of the CTA take work
// whenever they are free.  A seed is either being advanced by ONE warp (lock), or waiting for its queued problems
// (pending > 0: any warp may take chunks of them through `cursor`), or ready.  Control tasks come first (they feed the
// queue); a warp without a ready seed solves problems.  The FP64-bound solves fill the issue slots the latency-bound
// control code leaves empty.
__global__ void __launch_bounds__(KW_THREADS, KW_CTAS_PER_SM) k3w_async_kernel(const __grid_constant__ DevScene S, const __grid_constant__ K3Args A, int nS, int p64_in_smem) {
  extern __shared__ double kw_smem_p64[];
  __shared__ SeedSt st[KW_MAX_SEEDS];
  __shared__ int lock[KW_MAX_SEEDS], pending[KW_MAX_SEEDS], cursor[KW_MAX_SEEDS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_acc = (int)A.pa_counters[0];
  unsigned char* arena0 = A.scratch + (size_t)blockIdx.x * nS * A.scratch_per_warp;
  const double* P64 = S.P64;
  if (p64_in_smem) {
    for (int i = threadIdx.x; i < S.V * 12; i += blockDim.x) kw_smem_p64[i] = S.P64[i];
    P64 = kw_smem_p64;
  }
  for (int i = threadIdx.x; i < KW_MAX_SEEDS; i += blockDim.x) {
    st[i].state = i < nS ? KS_FREE : KS_IDLE; st[i].gq_n = 0; st[i].v = 0;
    lock[i] = 0; pending[i] = 0; cursor[i] = 0;
  }
  __syncthreads();
  volatile int* vlock = lock; volatile int* vpend = pending; volatile int* vcur = cursor;
  volatile SeedSt* vst = st;
  Ctx c;
  c.S = &S; c.A = &A; c.lane = lane;
#ifdef EG3D_K3_PROFILE
  for (int k = 0; k < 32; k++) c.pc[k] = 0;
#endif
  const int rot = (wid * 4) & 63;           // warps start their scans at different seeds
  int idle_spins = 0;
  for (;;) {
    // ---- scan the resident seeds (two per lane)
    const int ia = (lane + rot) & 63, ib = (lane + 32 + rot) & 63;
    const int sa = vst[ia].state, sb = vst[ib].state;
    const int la = vlock[ia], lb = vlock[ib];
    const int pa = vpend[ia], pb = vpend[ib];
    const int ga = vst[ia].gq_n - vcur[ia], gb = vst[ib].gq_n - vcur[ib];
    const unsigned ctl_a = __ballot_sync(0xffffffffu, ia < nS && sa != KS_IDLE && la == 0 && pa == 0);
    const unsigned ctl_b = __ballot_sync(0xffffffffu, ib < nS && sb != KS_IDLE && lb == 0 && pb == 0);
    const unsigned gn_a = __ballot_sync(0xffffffffu, ia < nS && la == 0 && pa > 0 && ga > 0);
    const unsigned gn_b = __ballot_sync(0xffffffffu, ib < nS && lb == 0 && pb > 0 && gb > 0);
    const unsigned live_a = __ballot_sync(0xffffffffu, ia < nS && sa != KS_IDLE), live_b = __ballot_sync(0xffffffffu, ib < nS && sb != KS_IDLE);
    if (ctl_a | ctl_b) {
      const int pick = ctl_a ? __ffs(ctl_a) - 1 : 32 + __ffs(ctl_b) - 1;
      const int i = (pick + rot) & 63;
      int got = 0;
      if (lane == 0) got = atomicCAS(&lock[i], 0, 1) == 0 ? 1 : 0;
      got = __shfl_sync(0xffffffffu, got, 0);
      if (!got) continue;
      __threadfence_block();
      if (vpend[i] != 0 || vst[i].state == KS_IDLE) { __syncwarp(); if (lane == 0) vlock[i] = 0; continue; }
      SeedSt s = st[i];
      c.w = make_ws(arena0 + (size_t)i * A.scratch_per_warp, S.V, A.capf, A.capc, A.oc);
      kw_load(c, s);
      for (;;) {
        const int region = kw_region_of(s.state);
        if (region == 1) kw_r1(c, s, n_acc);
        else if (region == 2) kw_r2(c, s);
        else if (region == 3) kw_r3(c, s);
        if (region == 0 || s.gq_n > 0 || s.state == KS_IDLE) break;
      }
      kw_save(c, s);
      __syncwarp();
      if (lane == 0) {
        st[i] = s;
        cursor[i] = 0; pending[i] = s.gq_n;
        __threadfence_block();
        vlock[i] = 0;
      }
      __syncwarp();
      idle_spins = 0;
      continue;
    }
    if (gn_a | gn_b) {
      const int pick = gn_a ? __ffs(gn_a) - 1 : 32 + __ffs(gn_b) - 1;
      const int i = (pick + rot) & 63;
      const int nb = vst[i].gq_n;
      const int G = kw_batch_width(nb), ppw = 32 / G;
      int k0 = 0;
      if (lane == 0) k0 = atomicAdd(&cursor[i], ppw);
      k0 = __shfl_sync(0xffffffffu, k0, 0);
      if (k0 >= nb) continue;
      const WS w = make_ws(arena0 + (size_t)i * A.scratch_per_warp, S.V, A.capf, A.capc, A.oc);
      kw_gn_chunk(S, w, vst[i].v, k0, nb, G, P64, lane);
      __syncwarp();
      if (lane == 0) { __threadfence_block(); atomicSub(&pending[i], min(ppw, nb - k0)); }
      idle_spins = 0;
      continue;
    }
    if ((live_a | live_b) == 0) break;
    if (++idle_spins > 4) __nanosleep(200);
  }
}

}  // namespace eg3d
