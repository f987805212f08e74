// eg3d_k3.cuh — K3: one warp per seed runs compute_3D_point_multiple_views_plg_following_expandallviews_vector
// (triangulation.cpp:1027-1088): view-triple selection, triple enumeration with the uniqueness test (:550-601),
// PLG following (plg_matching.cpp:51-370, 633-795, 1249-1287) and expansion to the remaining views
// (triangulation.cpp:742-830, 960-973; plg_matching.cpp:797-914, 1011-1058, 1345-1412).
//
// Execution model: the warp is a scalar processor with a 32-wide vector unit.  Control flow is warp-uniform (every
// lane evaluates the same sequential walk, so no broadcasts are needed), and the lanes fan out wherever the reference
// loops over independent work: triples (one 3-view DLT+GN per lane), the four direction combos, the epipolar hits of a
// view (one warm-started GN per lane, first success in order wins), the neighbours of a chain point (one GN per lane),
// the views of a chain point when a step is extended, and the observations of a single large GN.
// All per-seed state lives in a per-warp global scratch arena (L1/L2 resident); nothing is shared between warps, and a
// persistent grid pulls seeds from an atomic counter because per-seed cost varies by orders of magnitude.
#pragma once
#include "eg3d_dev.cuh"

namespace eg3d {

#ifndef EG3D_K3A_MIN_BLOCKS
#define EG3D_K3A_MIN_BLOCKS 12
#endif
// Phase B is bound by instruction-cache refills, not by occupancy: its hot path is ~65 KB of branchy scalar code, ncu
// shows the GPC-level instruction cache at 80 % of its peak request rate, and the kernel takes the same time with 8, 16,
// 28 or 32 resident warps per SM (profiles/r01_k3b_icache.md).  4 CTAs x 4 warps (128 registers) is used: as fast as any
// other shape, and with 16 warps per SM the per-lane stacks (1.4 KB) and scratch arenas fit the 126 MB L2, which cuts
// the kernel's DRAM traffic from 93 GB (7 CTAs) to 21 GB per launch.
// EG3D_K3B_SYNC=1 with EG3D_K3B_THREADS=640..1024, EG3D_K3B_MIN_BLOCKS=1 builds the lock-step form (one CTA per SM, a
// CTA barrier in front of each half of a view's expansion so that the warps share the lines they pull in): 35 % fewer
// instruction-cache requests but slower overall (barrier idling), kept for experiments.
#ifndef EG3D_K3A_FIRST_LOOKAHEAD
#define EG3D_K3A_FIRST_LOOKAHEAD 1
#endif
// Epipolar hits of a view are solved EG3D_EPC_BATCH at a time, in order: the first hit whose solve AND neighbour search succeed
// wins (triangulation.cpp:753-768), and on BASELINE configs[1] that is on average the third of ~9 survivors of the pruning, so
// solving all survivors at once (2 lanes each) solved mostly hits nobody asked about.  Measured (k3b ms for 2 / 4 / 6 / 8 / 12 / 16 / all):
// 183.5 / 185.9 / 180.2 / 176.2 / 183.0 / 180.2 / 186: smaller batches mean more trips through the non-loop code.
#ifndef EG3D_EPC_BATCH
#define EG3D_EPC_BATCH 8
#endif
#ifndef EG3D_EPC_BATCH_MIN_OBS
#define EG3D_EPC_BATCH_MIN_OBS 48
#endif
#ifndef EG3D_K3B_THREADS
#define EG3D_K3B_THREADS 128
#endif
#ifndef EG3D_K3B_MIN_BLOCKS
#define EG3D_K3B_MIN_BLOCKS 4
#endif
#ifndef EG3D_K3B_SYNC
#define EG3D_K3B_SYNC 0
#endif
#ifndef EG3D_K3B_BATCH
#define EG3D_K3B_BATCH 0   // > 0: lock-step form with this many seeds per warp (k3b_expand_kernel, batched variant)
#endif
constexpr int K3_THREADS = 128;  // phase A: 4 warps per CTA
constexpr int K3B_THREADS = EG3D_K3B_THREADS;

struct Pt3 {  // a 3-view point of the following phase (64 B)
  float X[3];
  uint8_t perm[4];    // perm[i] = index into sel[] of the i-th view of this point
  uint32_t pl[3], seg[3];
  float cx[3], cy[3];
};
struct NTmp { float X[3]; uint32_t seg; float cx, cy; };  // a neighbour candidate (polyline id is implied)

struct PaRec;
struct K3Args {
  int n_seeds;
  const int* seed_view; const uint32_t* seed_pl; const uint32_t* seed_seg; const float2* seed_xy;
  // Hit lists.  Phase A reads the three selected views of a seed: row seed*3 + k of (hit_off, hits) when a_compact, else
  // row seed*V + sel[k] of the full (seed, view) CSR.  Phase B reads every view of an accepted seed: row
  // (b_compact ? accepted rank : seed)*V + v of (hit_off_b, hits_b).
  const int64_t* hit_off; const eg3d_hit* hits; int a_compact;
  const int64_t* hit_off_b; const eg3d_hit* hits_b; int b_compact;
  const int* sel;           // [n_seeds][3] selected views (k3_select_views_kernel); sel[3*seed] < 0: fewer than 3 non-empty views
  int* acc_seed;            // [n_seeds] accepted seed of each phase-A record (rank order)
  int capf, capc, oc;       // capacities: follow points per list, chain points, observations per point
  int gn_cache;             // phase B: share the first Gauss-Newton iteration of a chain point's solves (gn_group); pays on rigs with many views
  unsigned char* scratch; size_t scratch_per_warp;
  int* work_counter;
  // unordered outputs (a warp reserves a contiguous range per seed)
  int64_t pt_cap, ob_cap;
  unsigned long long* out_counters;   // [0] points, [1] observations, [2] capacity overflows, [3] output overflows
  float* o_X; int* o_nobs; int64_t* o_obase;
  int* ob_view; uint32_t* ob_pl; uint32_t* ob_seg; float* ob_x; float* ob_y;
  int* seed_npts; int64_t* seed_pbase; int64_t* seed_nobs;
  unsigned long long* prof;   // optional [16] per-phase warp-cycle / event counters (null = off; profile builds only)
  unsigned long long* prof_seed;  // optional [n_seeds][3]: phase-B warp cycles, initial | final chain length (profile builds only)
  // phase A -> phase B hand-over: one record per accepted seed + its two 3-view point lists in a pool
  struct PaRec* pa_recs; Pt3* pa_pool; long long pa_pool_cap;
  unsigned long long* pa_counters;    // [0] accepted seeds, [1] pool entries used
  const int* pa_order;                // phase B visits pa_recs[pa_order[i]]: longest chains first (tail balance); null = as produced
};
struct PaRec {
  int seed; int sel[3]; int fn1, fn2;
  uint32_t fd1[3], fd2[3];
  Pt3 central;
  long long pool_off;
};

struct WS {  // per-warp scratch view
  Pt3 *tri, *D1, *D2, *fD1, *fD2;
  int* ov; uint32_t *opl, *oseg; float *ox, *oy;   // [capc][oc]
  float* sX; int* snobs; int* order;               // [capc][3], [capc], [capc]
  uint32_t *sdirs, *edirs;                         // [V]
  NTmp *tmp1, *tmp2;                               // [capc]
  int* idx;                                        // [oc] scratch index list (combination fallback)
  unsigned char* selmask;                          // [oc]
  int* tq;                                         // [3*64] queue of triples that survived pruning
  double* gcache; int* gtag;                       // [capc+1][10], [capc+1]: per-slot base sums of gn_group's shared first iteration
  int *fbs, *fbe, *fbm, *fbfs, *fbfe;                            // [oc], [oc], [4]: per-observation candidate counts of step_all_big at the chain's start / end slot + (slot, #observations evaluated) per side
  int capf, capc, oc;
};

inline __host__ __device__ size_t k3_align(size_t x) { return (x + 15) & ~(size_t)15; }
inline __host__ __device__ size_t k3_scratch_bytes(int V, int capf, int capc, int oc) {
  size_t b = 0;
  b += k3_align(sizeof(Pt3) * (size_t)capf * 12);
  b += k3_align(sizeof(int) * (size_t)(capc + 1) * oc) * 5;
  b += k3_align(sizeof(float) * 3 * capc) + k3_align(sizeof(int) * capc) * 2;
  b += k3_align(sizeof(uint32_t) * V) * 2;
  b += k3_align(sizeof(NTmp) * capc) * 2;
  b += k3_align(sizeof(int) * oc) + k3_align(oc);
  b += k3_align(sizeof(int) * 3 * 64);
  b += k3_align(sizeof(int) * oc) * 4 + k3_align(sizeof(int) * 4);
  b += k3_align(sizeof(double) * 10 * (size_t)(capc + 1)) + k3_align(sizeof(int) * (size_t)(capc + 1));
  return b;
}
EG3D_D WS make_ws(unsigned char* base, int V, int capf, int capc, int oc) {
  WS w; size_t o = 0;
  auto take = [&](size_t bytes) { unsigned char* p = base + o; o += k3_align(bytes); return p; };
  Pt3* p3 = (Pt3*)take(sizeof(Pt3) * (size_t)capf * 12);     // tri: 8 combo lists (first_extreme_dual), then D1, D2, fD1, fD2
  w.tri = p3; w.D1 = p3 + 8 * capf; w.D2 = p3 + 9 * capf; w.fD1 = p3 + 10 * capf; w.fD2 = p3 + 11 * capf;
  w.ov = (int*)take(sizeof(int) * (size_t)(capc + 1) * oc);
  w.opl = (uint32_t*)take(sizeof(int) * (size_t)(capc + 1) * oc);
  w.oseg = (uint32_t*)take(sizeof(int) * (size_t)(capc + 1) * oc);
  w.ox = (float*)take(sizeof(int) * (size_t)(capc + 1) * oc);
  w.oy = (float*)take(sizeof(int) * (size_t)(capc + 1) * oc);
  w.sX = (float*)take(sizeof(float) * 3 * capc);
  w.snobs = (int*)take(sizeof(int) * capc);
  w.order = (int*)take(sizeof(int) * capc);
  w.sdirs = (uint32_t*)take(sizeof(uint32_t) * V);
  w.edirs = (uint32_t*)take(sizeof(uint32_t) * V);
  w.tmp1 = (NTmp*)take(sizeof(NTmp) * capc);
  w.tmp2 = (NTmp*)take(sizeof(NTmp) * capc);
  w.idx = (int*)take(sizeof(int) * oc);
  w.selmask = (unsigned char*)take(oc);
  w.tq = (int*)take(sizeof(int) * 3 * 64);
  w.fbs = (int*)take(sizeof(int) * oc); w.fbe = (int*)take(sizeof(int) * oc); w.fbm = (int*)take(sizeof(int) * 4);
  w.fbfs = (int*)take(sizeof(int) * oc); w.fbfe = (int*)take(sizeof(int) * oc);
  w.gcache = (double*)take(sizeof(double) * 10 * (size_t)(capc + 1)); w.gtag = (int*)take(sizeof(int) * (size_t)(capc + 1));
  w.capf = capf; w.capc = capc; w.oc = oc;
  return w;
}

// ---------------------------------------------------------------------------------------------------------------
// per-seed context (registers; identical in every lane)
struct Ctx {
  const DevScene* S; const K3Args* A; WS w;
#ifdef EG3D_K3_PROFILE
  long long pc[32];   // profiling accumulators (profile builds only: build.sh -DEG3D_K3_PROFILE)
#endif
  int lane, seed, sv;
  long long hrow;  // phase B: first row of this seed in hit_off_b
  int sel[3];
  int len;         // chain length
  int nslots;      // slots in use
  int central;     // original_central_point_index
  bool overflow;
};

#ifdef EG3D_K3_PROFILE
#define K3P_BEGIN(t) long long t = clock64()
#define K3P_END(c, slot, t) (c).pc[slot] += clock64() - (t)
#define K3P_ADD(c, slot, v) (c).pc[slot] += (v)
#else
#define K3P_BEGIN(t) do {} while (0)
#define K3P_END(c, slot, t) do {} while (0)
#define K3P_ADD(c, slot, v) do {} while (0)
#endif

// 3-view compatible, plg_matching.cpp:51-132.  cur / next: (pl, seg, c) per view a,b,c in sel order.
struct Cur3 { uint32_t pl[3], seg[3]; float2 c[3]; };
// geometry of the 3-view step (plg_matching.cpp:66-108): 10 px along polyline a, epipolar walks on b and c
static __device__ __noinline__ bool geo3(const DevScene& S, const int ids[3], const Cur3& cur, const uint32_t dir[3], Cur3& next) {
  Pl pla = get_pl(S, ids[0], cur.pl[0]), plb = get_pl(S, ids[1], cur.pl[1]), plc = get_pl(S, ids[2], cur.pl[2]);
  bool reached;
  PlP ia; ia.seg = cur.seg[0]; ia.c = cur.c[0];
  PlP na = step_by_distance(pla, ia, dir[0], S.prm.follow_first_image_distance, reached);
  if (reached) return false;
  float3 l;
  if (!epiline(S, ids[0], ids[1], na.c, l)) return false;
  PlP ib; ib.seg = cur.seg[1]; ib.c = cur.c[1];
  PlP nb;
  if (!walk_line(plb, ib, dir[1], l, S.prm, false, nb)) return false;
  if (!epiline(S, ids[0], ids[2], na.c, l)) return false;
  PlP ic; ic.seg = cur.seg[2]; ic.c = cur.c[2];
  PlP nc;
  if (!walk_line(plc, ic, dir[2], l, S.prm, false, nc)) return false;
  next.pl[0] = cur.pl[0]; next.pl[1] = cur.pl[1]; next.pl[2] = cur.pl[2];
  next.seg[0] = na.seg; next.seg[1] = nb.seg; next.seg[2] = nc.seg;
  next.c[0] = na.c; next.c[1] = nb.c; next.c[2] = nc.c;
  return true;
}
// 3-view compatible, plg_matching.cpp:51-132
static __device__ __noinline__ bool step3(const DevScene& S, const int ids[3], const Cur3& cur, const uint32_t dir[3], Cur3& next, float X[3]) {
  if (!geo3(S, ids, cur, dir, next)) return false;
  float2 pts[3] = {next.c[0], next.c[1], next.c[2]};
  return est3(S, ids, pts, X);
}

EG3D_D void store_pt3(Pt3* dst, const Cur3& c, const float X[3]) {
  Pt3 p;
  p.X[0] = X[0]; p.X[1] = X[1]; p.X[2] = X[2];
  p.perm[0] = 0; p.perm[1] = 1; p.perm[2] = 2; p.perm[3] = 0;
#pragma unroll
  for (int i = 0; i < 3; i++) { p.pl[i] = c.pl[i]; p.seg[i] = c.seg[i]; p.cx[i] = c.c[i].x; p.cy[i] = c.c[i].y; }
  *dst = p;
}

// est3 of a stored 3-view candidate (views = sel[perm[i]]); writes X on success
EG3D_D bool est_pt3(const DevScene& S, const int sel[3], Pt3* p) {
  int v[3] = {sel[p->perm[0]], sel[p->perm[1]], sel[p->perm[2]]};
  float2 pt[3] = {make_float2(p->cx[0], p->cy[0]), make_float2(p->cx[1], p->cy[1]), make_float2(p->cx[2], p->cy[2])};
  // Exact-safe pruning, as for the triples of phase A (pair_cannot_fit): a followed candidate lies on view a's epipolar lines by
  // construction, but nothing ties its b and c observations to each other; when any view pair already costs more than the
  // acceptance budget the solve cannot be accepted — and a rejected solve tends to be the expensive one (no convergence: all 30
  // Gauss-Newton iterations after the DLT).  A rejected solve has no side effects, so skipping it changes nothing.
  {
    const double Tp = prune_radius(S.prm, 3);
    const size_t i01 = (size_t)v[0] * S.V + v[1], i02 = (size_t)v[0] * S.V + v[2], i12 = (size_t)v[1] * S.V + v[2];
    if (pair_cannot_fit(S.Fp + i12 * 9, S.Fph[i12], pt[1], pt[2], Tp) || pair_cannot_fit(S.Fp + i01 * 9, S.Fph[i01], pt[0], pt[1], Tp) ||
        pair_cannot_fit(S.Fp + i02 * 9, S.Fph[i02], pt[0], pt[2], Tp)) return false;
  }
  float X[3];
  if (!est3(S, v, pt, X)) return false;
  p->X[0] = X[0]; p->X[1] = X[1]; p->X[2] = X[2];
  return true;
}

// find_direction_given_first_extreme, plg_matching.cpp:142-203: the four (end_b, end_c) combos advance in lock-step and
// the last survivor wins.  A combo's fate depends only on its own chain, and the triangulated X of a step only gates
// validity (the next 2D step starts from the 2D points), so a combo's geometry can be walked ahead, the DLT+GN solves of a
// batch run one per lane, and each combo's lifetime L_k is the number of leading successes.  The lock-step loop
// `while (amount_of_valid > 1)` ends after round r* = L_(2) + 1 (second-largest lifetime + 1); the survivor (if its
// lifetime is larger) keeps exactly its first r* points.
//
// The caller tries first extreme A (pla.start) and, only if that yields nothing, first extreme B (pla.end)
// (plg_matching.cpp:325-370).  Almost every hypothesis that reaches this point is a wrong one whose combos all die in their
// first step for BOTH extremes (1-2 % of the calls end in an accepted seed), so the first step of both extremes is taken at
// once — lanes 0..3 the combos of A, lanes 4..7 those of B, one solve per lane — and only an extreme whose combos survive
// goes on, 8 steps per batch, A before B as in the reference; B's speculative first step has no side effects.  Returns the
// number of points copied into `dst` (0 = neither extreme works) and which extreme it was (`used`: 0 = A, 1 = B).
static __device__ __noinline__ int first_extreme_dual(Ctx& c, const Cur3& start, uint32_t dirA, uint32_t dirB, int& used, uint32_t dir_out[3], Pt3* dst) {
  const DevScene& S = *c.S;
  const int lane = c.lane;
  Pl plb = get_pl(S, c.sel[1], start.pl[1]), plc = get_pl(S, c.sel[2], start.pl[2]);
  uint32_t dir[3] = {(lane & 4) ? dirB : dirA, (lane & 2) ? plb.end : plb.start, (lane & 1) ? plc.end : plc.start};
  bool alive = lane < 8;
  Cur3 cur = start;
  int L = 0;                                  // successful steps so far (lanes 0..7)
  Pt3* mine = c.w.tri + (size_t)(lane & 7) * c.w.capf;
  used = 0;
  {   // first step of all eight combos
    bool got = false;
    if (alive) {
      Cur3 nx;
      if (geo3(S, c.sel, cur, dir, nx)) {
        float X0[3] = {0.f, 0.f, 0.f};
        store_pt3(mine, nx, X0);
        cur = nx; got = true;
      }
    }
    __syncwarp();
    bool ok = false;
    if (got) ok = est_pt3(S, c.sel, mine);
    __syncwarp();
    if (alive) { if (ok) L = 1; alive = ok; }   // a batch of one: the combo lives on iff its step exists and triangulates
  }
  constexpr int R = 8;
  for (int g = 0; g < 2; g++) {
    const unsigned gmask = 0xfu << (4 * g);
    const bool in_g = lane < 8 && (lane >> 2) == g;
    while (__popc(__ballot_sync(0xffffffffu, alive) & gmask) > 1) {
      int steps = 0;                            // geometric steps of this batch
      if (alive && in_g) {
        if (L + R > c.w.capf) c.overflow = true;
        else {
          for (; steps < R; steps++) {
            Cur3 nx;
            if (!geo3(S, c.sel, cur, dir, nx)) break;
            float X0[3] = {0.f, 0.f, 0.f};
            store_pt3(mine + L + steps, nx, X0);
            cur = nx;
          }
        }
      }
      if (__any_sync(0xffffffffu, c.overflow)) { c.overflow = true; return 0; }
      __syncwarp();
      // verification: lane (k*8 + r) solves candidate r of combo k of this extreme
      const int kk = lane >> 3, rr = lane & 7;
      const int gk = __shfl_sync(0xffffffffu, steps, 4 * g + kk), Lk = __shfl_sync(0xffffffffu, L, 4 * g + kk);
      bool ok = false;
      if (rr < gk) ok = est_pt3(S, c.sel, c.w.tri + (size_t)(4 * g + kk) * c.w.capf + Lk + rr);
      const unsigned okm = __ballot_sync(0xffffffffu, ok);
      __syncwarp();
      if (alive && in_g) {
        const unsigned mine_ok = (okm >> ((lane & 3) * 8)) & 0xffu;
        int lead = __ffs(~mine_ok) - 1;         // leading successes (8 when all ok)
        if (lead > steps) lead = steps;
        L += lead;
        if (lead < R) alive = false;            // geometry ended or a solve failed: lifetime is final
      }
    }
    // lifetimes of the four combos of this extreme
    const int L0 = __shfl_sync(0xffffffffu, L, 4 * g), L1 = __shfl_sync(0xffffffffu, L, 4 * g + 1), L2 = __shfl_sync(0xffffffffu, L, 4 * g + 2), L3 = __shfl_sync(0xffffffffu, L, 4 * g + 3);
    const int Ls[4] = {L0, L1, L2, L3};
    int win = 0;
    for (int k = 1; k < 4; k++) if (Ls[k] > Ls[win]) win = k;
    int second = -1;
    for (int k = 0; k < 4; k++) if (k != win && Ls[k] > second) second = Ls[k];
    if (Ls[win] <= second) continue;            // the best two die in the same round: no survivor for this extreme
    const int n = second + 1;                   // rounds executed by the reference's loop
    dir_out[0] = g ? dirB : dirA;
    dir_out[1] = __shfl_sync(0xffffffffu, dir[1], 4 * g + win);
    dir_out[2] = __shfl_sync(0xffffffffu, dir[2], 4 * g + win);
    __syncwarp();
    const Pt3* src = c.w.tri + (size_t)(4 * g + win) * c.w.capf;
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
    __syncwarp();
    used = g;
    return n;
  }
  return 0;
}

// Geometry of the all-view compatible (plg_matching.cpp:633-706) for a 3-view point and ONE driving view `si`:
// 10 px on the driving view, bounded epipolar walks on the other two.  Exactly three observations must survive
// (fewer => the attempt is skipped, :708-709); the combination fallback (:715-733) degenerates to the same 3-subset
// and cannot succeed, so a failed solve simply moves on to the next driving view.
static __device__ __noinline__ bool geo_all3(const DevScene& S, const int sel[3], const uint32_t dirs[3], const Pt3& cur, int si, Pt3& out) {
  const int k = cur.perm[si];
  const int sv = sel[k];
  Pl pls = get_pl(S, sv, cur.pl[si]);
  bool reached;
  PlP ip; ip.seg = cur.seg[si]; ip.c = make_float2(cur.cx[si], cur.cy[si]);
  PlP ns = step_by_distance(pls, ip, dirs[k], S.prm.follow_first_image_distance, reached);
  if (reached) return false;
  int nv = 1;
  out.perm[0] = (uint8_t)k; out.pl[0] = cur.pl[si]; out.seg[0] = ns.seg; out.cx[0] = ns.c.x; out.cy[0] = ns.c.y; out.perm[3] = 0;
  for (int i = 0; i < 3; i++) {
    if (i == si) continue;
    const int ki = cur.perm[i];
    const int vv = sel[ki];
    float3 l;
    if (!epiline(S, sv, vv, ns.c, l)) continue;
    Pl pl = get_pl(S, vv, cur.pl[i]);
    PlP iq; iq.seg = cur.seg[i]; iq.c = make_float2(cur.cx[i], cur.cy[i]);
    PlP np;
    if (walk_line(pl, iq, dirs[ki], l, S.prm, true, np)) {
      if (nv < 3) { out.perm[nv] = (uint8_t)ki; out.pl[nv] = cur.pl[i]; out.seg[nv] = np.seg; out.cx[nv] = np.c.x; out.cy[nv] = np.c.y; }
      nv++;
    }
  }
  return nv >= 3;
}

// follow_direction on a 3-view list (plg_matching.cpp:765-769).  Speculative form: the 2D chain is walked ahead with
// the first driving view whose geometry works at every step, the solves of up to 32 steps run one per lane, and the
// chain is cut at the first failed solve — where the reference's remaining driving views are then tried in order.
static __device__ __noinline__ void follow3(Ctx& c, const uint32_t dirs[3], Pt3* list, int& n) {
  const DevScene& S = *c.S;
  Pt3* cand = c.w.tri;                        // scratch (first_extreme is done with it)
  int* si_used = c.w.idx;
  while (true) {
    int K = 0;
    Pt3 cur = list[n - 1];
    for (; K < 32; K++) {
      Pt3 nx; int si = 0; bool got = false;
      for (; si < 3; si++) if (geo_all3(S, c.sel, dirs, cur, si, nx)) { got = true; break; }
      if (!got) break;
      nx.X[0] = nx.X[1] = nx.X[2] = 0.f;
      __syncwarp();
      if (c.lane == 0) { cand[K] = nx; si_used[K] = si; }
      cur = nx;
    }
    if (K == 0) break;
    __syncwarp();
    bool bad = false;
    if (c.lane < K) bad = !est_pt3(S, c.sel, cand + c.lane);
    unsigned fm = __ballot_sync(0xffffffffu, bad);
    const int f = fm ? (__ffs(fm) - 1) : K;
    __syncwarp();
    if (n + f > c.w.capf) { c.overflow = true; return; }
    for (int i = c.lane; i < f; i += 32) list[n + i] = cand[i];
    __syncwarp();
    n += f;
    if (f == K) { if (K < 32) break; else continue; }
    // the solve of step f failed for driving view si_used[f]: the reference goes on with the remaining driving views
    Pt3 base = list[n - 1];
    bool found = false;
    for (int si = si_used[f] + 1; si < 3 && !found; si++) {
      Pt3 nx;
      if (!geo_all3(S, c.sel, dirs, base, si, nx)) continue;
      __syncwarp();
      if (c.lane == 0) cand[0] = nx;
      __syncwarp();
      bool ok = est_pt3(S, c.sel, cand);      // uniform: every lane computes the same solve; identical stores
      __syncwarp();
      if (ok) {
        if (n >= c.w.capf) { c.overflow = true; return; }
        if (c.lane == 0) list[n] = cand[0];
        __syncwarp();
        n++; found = true;
      }
    }
    if (!found) break;
  }
}

// compatible_new_plg_point, plg_matching.cpp:1276-1287 (-> :1249-1270 -> :1060-1076 -> :325-370 -> :205-265).
// D1/D2 (and their counts) persist across the triples of a seed and are only overwritten when the corresponding
// direction is valid for THIS hypothesis.
static __device__ __noinline__ bool plg_compatible(Ctx& c, const Cur3& cand, int& n1, int& n2, uint32_t d1[3], uint32_t d2[3]) {
  const DevScene& S = *c.S;
  Pl pla = get_pl(S, c.sel[0], cand.pl[0]), plb = get_pl(S, c.sel[1], cand.pl[1]), plc = get_pl(S, c.sel[2], cand.pl[2]);
  bool d1ok = false, d2ok = false;
  uint32_t dir1[3];
  auto opposite = [&](const uint32_t d[3], uint32_t o[3]) {
    o[0] = pla.start == d[0] ? pla.end : pla.start;
    o[1] = plb.start == d[1] ? plb.end : plb.start;
    o[2] = plc.start == d[2] ? plc.end : plc.start;
  };
  int used = 0;
  const int n = first_extreme_dual(c, cand, pla.start, pla.end, used, dir1, c.w.D1);
  if (c.overflow) return false;
  if (n > 0) {
    d1ok = true; n1 = n;
    d1[0] = dir1[0]; d1[1] = dir1[1]; d1[2] = dir1[2];
    opposite(dir1, d2);
    if (used == 0) {                            // the pla.start extreme worked: one step towards the opposite extremes (:345-356)
      Cur3 nx; float X[3];
      if (step3(S, c.sel, cand, d2, nx, X)) {
        d2ok = true;
        __syncwarp();
        if (c.lane == 0) store_pt3(c.w.D2, nx, X);
        __syncwarp();
        n2 = 1;
      }
    }
  }
  if (d1ok) follow3(c, d1, c.w.D1, n1);
  if (d2ok) follow3(c, d2, c.w.D2, n2);
  if (c.overflow) return false;
  if (d1ok && n1 >= 2) return true;
  if (d2ok && n2 >= 2) return true;
  return false;
}

// ---------------------------------------------------------------------------------------------------------------
// expansion phase: chain of "big" points held in slots
EG3D_D int slot_of(const Ctx& c, int pos) { return c.w.order[pos]; }

template <bool GC = false>
EG3D_D ObsSrc slot_obs(const Ctx& c, int slot, int n, bool extra, int ev, float ex, float ey) {
  ObsSrc o; size_t b = (size_t)slot * c.w.oc;
  o.v = c.w.ov + b; o.x = c.w.ox + b; o.y = c.w.oy + b; o.n = n;
  o.has_extra = extra ? 1 : 0; o.ev = ev; o.ex = ex; o.ey = ey;
  if (GC) { o.cache = c.w.gcache + 10 * (size_t)slot; o.tag = c.w.gtag + slot; } else { o.cache = nullptr; o.tag = nullptr; }
  return o;
}

// append one observation to a slot (lane 0 writes; callers sync)
static __device__ __noinline__ void slot_append(Ctx& c, int slot, int view, uint32_t pl, uint32_t seg, float x, float y, const float X[3]) {
  int n = c.w.snobs[slot];
  if (n >= c.w.oc) { c.overflow = true; return; }
  if (c.lane == 0) {
    size_t b = (size_t)slot * c.w.oc + n;
    c.w.ov[b] = view; c.w.opl[b] = pl; c.w.oseg[b] = seg; c.w.ox[b] = x; c.w.oy[b] = y;
    c.w.snobs[slot] = n + 1;
    c.w.sX[3 * slot] = X[0]; c.w.sX[3 * slot + 1] = X[1]; c.w.sX[3 * slot + 2] = X[2];
  }
}

// em_estimate3Dpositions over the n observations of a slot (triangulation.cpp:178-250): DLT from (first arg-min view,
// last entry) + warp-cooperative GN.
static __device__ __noinline__ bool est_slot(Ctx& c, int slot, int n, float Xo[3]) {
  const DevScene& S = *c.S;
  const size_t b = (size_t)slot * c.w.oc;
  const int* ov = c.w.ov + b;
  int bestv = 0x7fffffff, besti = 0x7fffffff;
  for (int i = c.lane; i < n; i += 32) { int v = ov[i]; if (v < bestv) { bestv = v; besti = i; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int ov2 = __shfl_xor_sync(0xffffffffu, bestv, o), oi2 = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov2 < bestv || (ov2 == bestv && oi2 < besti)) { bestv = ov2; besti = oi2; }
  }
  int mi = besti, ma = n - 1;
  if (S.prm.dlt_wellposed == 1 && ov[ma] == ov[mi]) {
    for (int j = n - 1; j >= 0; j--) if (ov[j] != ov[mi]) { ma = j; break; }
  }
  float t4[4];
  dlt_null_opencv(S.P + 12 * ov[mi], S.P + 12 * ov[ma], make_float2(c.w.ox[b + mi], c.w.oy[b + mi]), make_float2(c.w.ox[b + ma], c.w.oy[b + ma]), t4);
  double X[3] = {(double)(t4[0] / t4[3]), (double)(t4[1] / t4[3]), (double)(t4[2] / t4[3])};
  if (!gn_seq_exact(S, slot_obs(c, slot, n, false, 0, 0.f, 0.f), c.lane, X)) return false;
  Xo[0] = (float)X[0]; Xo[1] = (float)X[1]; Xo[2] = (float)X[2];
  return true;
}

// next 3-combination in lexicographic order over [0, n); returns false past the last one
EG3D_D bool next_comb3(int& i, int& j, int& k, int n) {
  if (k + 1 < n) { k++; return true; }
  if (j + 2 < n) { j++; k = j + 1; return true; }
  if (i + 3 < n) { i++; j = i + 1; k = j + 1; return true; }
  return false;
}

// compute_3d_point_coords_combinations(min_combinations = 3), triangulation.cpp:1105-1158, on the observations of a
// slot.  On success the slot is compacted to the selected observations in their ORIGINAL order
// (plg_matching.cpp:720-732) and n / Xo are updated.
static __device__ __noinline__ bool combos_slot(Ctx& c, int slot, int& n, float Xo[3]) {
  const DevScene& S = *c.S;
  const size_t b = (size_t)slot * c.w.oc;
  int* ov = c.w.ov + b; float* ox = c.w.ox + b; float* oy = c.w.oy + b; uint32_t* opl = c.w.opl + b; uint32_t* oseg = c.w.oseg + b;
  // 1) first 3-subset (lexicographic = std::prev_permutation order of the selection mask) that triangulates
  int bi = 0, bj = 1, bk = 2;
  bool more = true, got = false;
  float X[3] = {0, 0, 0};
  int si = 0, sj = 0, sk = 0;
  while (more && !got) {
    int i = bi, j = bj, k = bk;
    bool mine = true;
    for (int s = 0; s < c.lane && mine; s++) mine = next_comb3(i, j, k, n);
    bool ok = false; float Xl[3] = {0, 0, 0};
    if (mine) {
      int v[3] = {ov[i], ov[j], ov[k]};
      float2 pt[3] = {make_float2(ox[i], oy[i]), make_float2(ox[j], oy[j]), make_float2(ox[k], oy[k])};
      ok = est3(S, v, pt, Xl);
    }
    unsigned m = __ballot_sync(0xffffffffu, ok);
    if (m) {
      int w = __ffs(m) - 1;
      si = __shfl_sync(0xffffffffu, i, w); sj = __shfl_sync(0xffffffffu, j, w); sk = __shfl_sync(0xffffffffu, k, w);
      X[0] = __shfl_sync(0xffffffffu, Xl[0], w); X[1] = __shfl_sync(0xffffffffu, Xl[1], w); X[2] = __shfl_sync(0xffffffffu, Xl[2], w);
      got = true;
    } else {
      for (int s = 0; s < 32 && more; s++) more = next_comb3(bi, bj, bk, n);
    }
  }
  if (!got) return false;
  // 2) greedily add the remaining observations in index order (warm-started GN each); the selected observations are
  //    kept as a compact copy in the spare slot (index capc) in the order they were selected
  __syncwarp();
  const size_t tb = (size_t)c.w.capc * c.w.oc;
  int* tv = c.w.ov + tb; float* tx = c.w.ox + tb; float* ty = c.w.oy + tb;
  for (int i = c.lane; i < n; i += 32) c.w.selmask[i] = (i == si || i == sj || i == sk) ? 1 : 0;
  if (c.lane == 0) {
    tv[0] = ov[si]; tx[0] = ox[si]; ty[0] = oy[si];
    tv[1] = ov[sj]; tx[1] = ox[sj]; ty[1] = oy[sj];
    tv[2] = ov[sk]; tx[2] = ox[sk]; ty[2] = oy[sk];
  }
  __syncwarp();
  int m = 3;
  for (int i = 0; i < n; i++) {
    if (c.w.selmask[i]) continue;
    ObsSrc obs; obs.v = tv; obs.x = tx; obs.y = ty; obs.n = m; obs.has_extra = 1; obs.ev = ov[i]; obs.ex = ox[i]; obs.ey = oy[i]; obs.cache = nullptr; obs.tag = nullptr;
    double Xd[3] = {X[0], X[1], X[2]};
    if (gn_seq_exact(S, obs, c.lane, Xd)) {
      X[0] = (float)Xd[0]; X[1] = (float)Xd[1]; X[2] = (float)Xd[2];
      __syncwarp();
      if (c.lane == 0) { c.w.selmask[i] = 1; tv[m] = ov[i]; tx[m] = ox[i]; ty[m] = oy[i]; }
      __syncwarp();
      m++;
    }
  }
  // 3) compact the slot by the mask, original order
  __syncwarp();
  int w = 0;
  for (int i = 0; i < n; i++) {
    if (c.w.selmask[i]) {
      if (w != i && c.lane == 0) { ov[w] = ov[i]; ox[w] = ox[i]; oy[w] = oy[i]; opl[w] = opl[i]; oseg[w] = oseg[i]; }
      w++;
    }
  }
  __syncwarp();
  n = w;
  Xo[0] = X[0]; Xo[1] = X[1]; Xo[2] = X[2];
  return true;
}

// all-view compatible (plg_matching.cpp:633-759) on a chain point with many views.  Builds the candidate in slot
// c.nslots (not yet claimed); returns true and leaves the new point there when a step is found.
//
// The reference re-evaluates every driving view of the extreme point each time a view has been added to it, although a
// driving view's candidate only ever GAINS observations: candidate(si) = the 10 px step on view si + one bounded epipolar
// walk per other observation i, and whether observation i contributes depends on (si, i) alone (the slot's observations are
// append-only and a view's direction entry is written once, before its observation can be driven).  So per side the
// number of observations each driving view collected is remembered (cnt[si]: -1 = cannot advance, >= 1 = count at the last
// call), a call only walks the NEW observations for the old driving views — all of them at once, one per lane — and the
// full candidate (with its DLT + Gauss-Newton) is rebuilt, in the reference's order, only for driving views that reach
// three observations.  A candidate with fewer than three observations is skipped by the reference too (:708-709), so
// the sequence of attempted solves, and with it every result, is unchanged.
static __device__ __noinline__ bool step_all_big(Ctx& c, const uint32_t* dirs, int cur_slot, int side) {
  const DevScene& S = *c.S;
  if (c.nslots >= c.w.capc) { c.overflow = true; return false; }
  const int t = c.nslots;
  const int n = c.w.snobs[cur_slot];
  const size_t cb = (size_t)cur_slot * c.w.oc, tb = (size_t)t * c.w.oc;
  int* cnt = side ? c.w.fbe : c.w.fbs;
  int* failc = side ? c.w.fbfe : c.w.fbfs;     // count at which the candidate's solves last failed (0 = never tried)
  int* meta = c.w.fbm + 2 * side;
  const int n_old = meta[0] == cur_slot ? min(meta[1], n) : 0;
  __syncwarp();
  // phase 1, one observation per lane: old driving views take the new observations into account, new ones are classified
  for (int sbase = 0; sbase < n; sbase += 32) {
    const int si = sbase + c.lane;
    if (si < n) {
      int k = si < n_old ? cnt[si] : 0;
      if (k != -1) {
        const int sv = c.w.ov[cb + si];
        Pl pls = get_pl(S, sv, c.w.opl[cb + si]);
        PlP ip; ip.seg = c.w.oseg[cb + si]; ip.c = make_float2(c.w.ox[cb + si], c.w.oy[cb + si]);
        bool reached;
        PlP ns = step_by_distance(pls, ip, dirs[sv], S.prm.follow_first_image_distance, reached);
        if (si >= n_old) { k = reached ? -1 : 0; failc[si] = 0; }   // 0: can advance, candidate not built yet
        else if (!reached) {
          for (int j = n_old; j < n; j++) {
            const int vv = c.w.ov[cb + j];
            float3 l;
            if (!epiline(S, sv, vv, ns.c, l)) continue;
            Pl pl = get_pl(S, vv, c.w.opl[cb + j]);
            PlP iq; iq.seg = c.w.oseg[cb + j]; iq.c = make_float2(c.w.ox[cb + j], c.w.oy[cb + j]);
            PlP np;
            if (walk_line(pl, iq, dirs[vv], l, S.prm, true, np)) k++;
          }
        }
        cnt[si] = k;
      } else cnt[si] = -1;
    }
  }
  __syncwarp();
  if (c.lane == 0) { meta[0] = cur_slot; meta[1] = n; }
  __syncwarp();
  // phase 2, in the reference's order: driving views whose candidate is new or has (at least) three observations
  for (int sbase = 0; sbase < n; sbase += 32) {
    const int si0 = sbase + c.lane;
    const int k0 = si0 < n ? cnt[si0] : -1;
    // a candidate whose DLT + Gauss-Newton (and combination fall-back) failed is a deterministic function of its observations:
    // it is tried again only once it has gained one
    unsigned cm = __ballot_sync(0xffffffffu, k0 == 0 || (k0 >= 3 && failc[si0 < n ? si0 : 0] != k0));
    while (cm) {
      const int si = sbase + __ffs(cm) - 1;
      cm &= cm - 1;
      const int sv = c.w.ov[cb + si];
      const uint32_t spl = c.w.opl[cb + si];
      Pl pls = get_pl(S, sv, spl);
      PlP ip; ip.seg = c.w.oseg[cb + si]; ip.c = make_float2(c.w.ox[cb + si], c.w.oy[cb + si]);
      bool reached;
      PlP ns = step_by_distance(pls, ip, dirs[sv], S.prm.follow_first_image_distance, reached);
      if (reached) continue;
      __syncwarp();
      if (c.lane == 0) { c.w.ov[tb] = sv; c.w.opl[tb] = spl; c.w.oseg[tb] = ns.seg; c.w.ox[tb] = ns.c.x; c.w.oy[tb] = ns.c.y; }
      int count = 1;
      for (int base = 0; base < n; base += 32) {
        const int i = base + c.lane;
        bool found = false; PlP np; int vv = 0; uint32_t ipl = 0;
        if (i < n && i != si) {
          vv = c.w.ov[cb + i]; ipl = c.w.opl[cb + i];
          float3 l;
          if (epiline(S, sv, vv, ns.c, l)) {
            Pl pl = get_pl(S, vv, ipl);
            PlP iq; iq.seg = c.w.oseg[cb + i]; iq.c = make_float2(c.w.ox[cb + i], c.w.oy[cb + i]);
            found = walk_line(pl, iq, dirs[vv], l, S.prm, true, np);
          }
        }
        unsigned m = __ballot_sync(0xffffffffu, found);
        if (found) {
          int pos = count + __popc(m & ((1u << c.lane) - 1u));
          c.w.ov[tb + pos] = vv; c.w.opl[tb + pos] = ipl; c.w.oseg[tb + pos] = np.seg; c.w.ox[tb + pos] = np.c.x; c.w.oy[tb + pos] = np.c.y;
        }
        count += __popc(m);
      }
      __syncwarp();
      if (c.lane == 0) cnt[si] = count;
      if (count < 3) continue;
      float X[3];
      K3P_ADD(c, 22, 1);
      bool valid = est_slot(c, t, count, X);
      if (!valid) { K3P_ADD(c, 23, 1); valid = combos_slot(c, t, count, X); }
      if (valid) {
        __syncwarp();
        if (c.lane == 0) { c.w.snobs[t] = count; c.w.sX[3 * t] = X[0]; c.w.sX[3 * t + 1] = X[1]; c.w.sX[3 * t + 2] = X[2]; }
        __syncwarp();
        return true;
      }
      if (c.lane == 0) failc[si] = count;
    }
  }
  return false;
}

// follow_direction_vector_start / _end, plg_matching.cpp:771-795.  Returns the number of points added.
static __device__ __noinline__ int follow_big(Ctx& c, const uint32_t* dirs, bool at_start) {
  int added = 0;
  K3P_ADD(c, 21, 1);
  while (true) {
    int cur_slot = slot_of(c, at_start ? 0 : c.len - 1);
    if (!step_all_big(c, dirs, cur_slot, at_start ? 0 : 1)) break;
    if (c.len >= c.w.capc) { c.overflow = true; break; }
    const int t = c.nslots++;
    __syncwarp();
    if (at_start) {
      // shift order right by one (descending so a single lane can do it safely)
      if (c.lane == 0) { for (int i = c.len; i > 0; i--) c.w.order[i] = c.w.order[i - 1]; c.w.order[0] = t; }
    } else if (c.lane == 0) c.w.order[c.len] = t;
    __syncwarp();
    c.len++;
    added++;
  }
  return added;
}

// compatible_direction_noupdate_vector, plg_matching.cpp:866-914, in two parts.  walk_geo: the polyline walk over the
// chain points on one side of `cur` (sequential, cheap, warp-cooperative steps); fills tmp[] with the new view's
// observation of every chain point reached and returns their number.
static __device__ __noinline__ int walk_geo(Ctx& c, int v, const Plg& p, uint32_t dir, bool towards_start, int lo, int cur, int hi, NTmp* tmp) {
  const DevScene& S = *c.S;
  Pl pl = get_pl(S, v, p.pl);
  PlP q; q.seg = p.seg; q.c = p.c;
  int cnt = 0;
  __syncwarp();
  // the chain points in walking order; their epipolar lines (from each point's FIRST observation, :797-806) are
  // independent of the walk and are computed 32 at a time, the walk itself is sequential with a 32-segment-wide step
  const int nwalk = towards_start ? cur - lo : hi - cur - 1;
  K3P_ADD(c, 27, 1); K3P_ADD(c, 28, nwalk);
  bool stop = false;
  for (int base = 0; base < nwalk && !stop; base += 32) {
    const int k = base + c.lane;
    float3 lk = make_float3(0.f, 0.f, 0.f); bool lok = false;
    if (k < nwalk) {
      const size_t b = (size_t)slot_of(c, towards_start ? cur - 1 - k : cur + 1 + k) * c.w.oc;
      lok = epiline(S, c.w.ov[b], v, make_float2(c.w.ox[b], c.w.oy[b]), lk);
    }
    const int nb = min(32, nwalk - base);
    for (int kk = 0; kk < nb; kk++) {
      if (!__shfl_sync(0xffffffffu, (int)lok, kk)) { stop = true; break; }
      const float3 l = make_float3(__shfl_sync(0xffffffffu, lk.x, kk), __shfl_sync(0xffffffffu, lk.y, kk), __shfl_sync(0xffffffffu, lk.z, kk));
      PlP nq;
      if (!walk_line_warp(pl, q, dir, l, S.prm, false, nq, c.lane)) { stop = true; break; }
      if (c.lane == 0) { tmp[cnt].seg = nq.seg; tmp[cnt].cx = nq.c.x; tmp[cnt].cy = nq.c.y; }
      cnt++;
      q = nq;
    }
  }
  __syncwarp();
  return cnt;
}

// walk_solve: the warm-started GNs of the chain points reached by walk_geo are independent, so the cnt1 points towards
// the start (tmp1) and the cnt2 points towards the end (tmp2) are solved together, up to 32 problems per gn_group call;
// each side's list is cut at its first failure (keep1 / keep2 = accepted neighbours, their X written to tmp[].X).
template <bool GC>
static __device__ __noinline__ void walk_solve(Ctx& c, int v, int cur, NTmp* tmp1, int cnt1, NTmp* tmp2, int cnt2, int& keep1, int& keep2) {
  const DevScene& S = *c.S;
  keep1 = cnt1; keep2 = cnt2;
  const int total = cnt1 + cnt2;
  for (int base = 0; base < total;) {
    const int P = min(32, total - base);
    const int G = gn_group_width(P);
    const int q = base + c.lane / G;                     // problem index in the concatenated list
    const bool active = (c.lane / G) < P;
    const bool side1 = q < cnt1;
    const int k = side1 ? q : q - cnt1;
    NTmp* tmp = side1 ? tmp1 : tmp2;
    bool fail = false;
    int slot = 0, n = 0; float ex = 0.f, ey = 0.f;
    double X[3] = {0, 0, 0};
    if (active) {
      slot = slot_of(c, side1 ? cur - 1 - k : cur + 1 + k);
      n = c.w.snobs[slot];
      X[0] = c.w.sX[3 * slot]; X[1] = c.w.sX[3 * slot + 1]; X[2] = c.w.sX[3 * slot + 2];
      ex = tmp[k].cx; ey = tmp[k].cy;
    }
    K3P_ADD(c, 16, 1); K3P_ADD(c, 17, P);
    const bool ok = gn_group<GC>(S, slot_obs<GC>(c, slot, n, true, v, ex, ey), active, G, c.lane, X);
    if (active) {
      if (ok) { if ((c.lane & (G - 1)) == 0) { tmp[k].X[0] = (float)X[0]; tmp[k].X[1] = (float)X[1]; tmp[k].X[2] = (float)X[2]; } }
      else fail = true;
    }
    const unsigned f1 = __ballot_sync(0xffffffffu, fail && side1), f2 = __ballot_sync(0xffffffffu, fail && !side1);
    if (f1) keep1 = min(keep1, base + (__ffs(f1) - 1) / G);
    if (f2) keep2 = min(keep2, base + (__ffs(f2) - 1) / G - cnt1);
    base += P;
    if (keep1 < cnt1 && (keep2 < cnt2 || cnt2 == 0) ) break;          // both lists are already cut
    if (keep1 < cnt1 && base < cnt1) base = cnt1;                      // the rest of side 1 is moot
    if (keep2 < cnt2 && base >= cnt1) break;                           // the rest of side 2 is moot
  }
  __syncwarp();
}

template <bool GC>
static __device__ __noinline__ int walk_dir(Ctx& c, int v, const Plg& p, uint32_t dir, bool towards_start, int lo, int cur, int hi, NTmp* tmp) {
  const int cnt = walk_geo(c, v, p, dir, towards_start, lo, cur, hi, tmp);
  if (cnt == 0) return 0;
  int k1 = 0, k2 = 0;
  if (towards_start) walk_solve<GC>(c, v, cur, tmp, cnt, tmp, 0, k1, k2);
  else walk_solve<GC>(c, v, cur, tmp, 0, tmp, cnt, k1, k2);
  return towards_start ? k1 : k2;
}

// add_view_to_3dpoint_and_sides_plgp_matches_vector after its first GN succeeded (plg_matching.cpp:1345-1412;
// neighbour search = find_directions_on_plg_known_3D_point_no_update_vector :1011-1058)
template <bool GC>
static __device__ __noinline__ bool add_view_finish(Ctx& c, int v, const Plg& p, const float Xc[3], int lo, int cur, int hi, int& ns, int& ne) {
  const DevScene& S = *c.S;
  Pl pl = get_pl(S, v, p.pl);
  int n1 = 0, n2 = 0; uint32_t nd1 = 0, nd2 = 0;
  if (cur > lo) {
    // Common case first: the start side follows pl.start and the end side pl.end.  Both walks are done before any solve,
    // and their neighbours are solved in ONE batch (a failed solve has no side effects, and when the start side yields
    // nothing the end-side work is simply discarded and the reference's order of attempts resumes below).
    const int g1 = walk_geo(c, v, p, pl.start, true, lo, cur, hi, c.w.tmp1);
    if (g1 > 0) {
      const int g2 = cur < hi ? walk_geo(c, v, p, pl.end, false, lo, cur, hi, c.w.tmp2) : 0;
      int k2 = 0;
      walk_solve<GC>(c, v, cur, c.w.tmp1, g1, c.w.tmp2, g2, n1, k2);
      if (n1 > 0) n2 = k2;
    }
    if (n1 > 0) {
      nd1 = pl.start; nd2 = pl.end;
    } else {
      n1 = walk_dir<GC>(c, v, p, pl.end, true, lo, cur, hi, c.w.tmp1);
      if (n1 > 0) {
        nd1 = pl.end; nd2 = pl.start;
        if (cur < hi) n2 = walk_dir<GC>(c, v, p, pl.start, false, lo, cur, hi, c.w.tmp2);
      } else if (cur < hi) {
        n2 = walk_dir<GC>(c, v, p, pl.end, false, lo, cur, hi, c.w.tmp2);
        if (n2 > 0) { nd2 = pl.end; nd1 = pl.start; }
        else {
          n2 = walk_dir<GC>(c, v, p, pl.start, false, lo, cur, hi, c.w.tmp2);
          if (n2 > 0) { nd2 = pl.start; nd1 = pl.end; }
        }
      }
    }
  }
  if (cur > 0 && n1 == 0) return false;           // SWITCH_PLG_MATCHING_ADDPOINT_BOTHDIR_ONE (:1363-1368)
  if (cur < c.len - 1 && n2 == 0) return false;
  // success: commit
  __syncwarp();
  slot_append(c, slot_of(c, cur), v, p.pl, p.seg, p.c.x, p.c.y, Xc);
  // the neighbours sit in distinct slots: one append per lane
  for (int q = c.lane; q < n1 + n2; q += 32) {
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
    int added = follow_big(c, c.w.sdirs, true);
    ns += added; cur += added;
  }
  if (n2 > 0 && n2 == (c.len - cur - 1)) {
    if (c.lane == 0) c.w.edirs[v] = nd2;
    __syncwarp();
    ne += follow_big(c, c.w.edirs, false);
  }
  return true;
}

// expand_allpoints_to_other_view_using_plmap, triangulation.cpp:742-830, in two halves (the phase-B kernel puts a CTA
// barrier in front of each): the epipolar hits on the central point (:753-768), then the projection loop over the chain
// points the first half did not reach (:770-830).
struct EvState { bool matched; int iv0, iv1; };
template <bool GC>
static __device__ __noinline__ void expand_view_epc(Ctx& c, int v, EvState& st) {
  const DevScene& S = *c.S;
  const int64_t h0 = c.A->hit_off_b[c.hrow + v];
  const int nh = (int)(c.A->hit_off_b[c.hrow + v + 1] - h0);
  const eg3d_hit* epcs = c.A->hits_b + h0;
  bool matched = false; int iv0 = 0, iv1 = 0;
  // the epipolar hits on the central point: one warm-started GN per lane, first complete success in order wins
  // Exact-safe pruning (pair_cannot_fit): a hit whose 2-view cost against one of the central point's observations
  // already exceeds the acceptance budget of the (n+1)-observation solve cannot be accepted; the survivors keep their
  // order (the first complete success wins, triangulation.cpp:753-768) and are solved in groups of lanes.
  int* hq = c.w.tq;                           // bounded queue (<= 63) of surviving hit indices, ascending
  int nq = 0, enext = 0;
  while (!matched) {
    // short observation lists (few views so far, or a rig with few views): one batch — a solve is cheap there and every extra
    // gn_group call is a trip through non-loop code (the packaged 25-view example ran 3-8 % slower with batches of 8)
    const int batch = c.w.snobs[slot_of(c, c.central)] >= EG3D_EPC_BATCH_MIN_OBS ? EG3D_EPC_BATCH : 32;
    K3P_BEGIN(te0);
    {
      const int cslot = slot_of(c, c.central);
      const int n = c.w.snobs[cslot];
      const size_t cb = (size_t)cslot * c.w.oc;
      const double Tp = prune_radius(S.prm, n + 1);
      const int probe[4] = {0, n / 3, (2 * n) / 3, n - 1};
      while (nq < batch && enext < nh) {
        const int e = enext + c.lane;
        bool pass = false;
        if (e < nh) {
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
        if (pass) hq[nq + __popc(pm & ((1u << c.lane) - 1u))] = e;
        nq += __popc(pm);
        enext += 32;
      }
      __syncwarp();
    }
    K3P_END(c, 3, te0);
    if (nq == 0) break;
    const int P = nq < batch ? nq : batch;
    K3P_ADD(c, 11, P);
    const int G = gn_group_width(P);
    const bool active = (c.lane / G) < P;
    const int e = active ? hq[c.lane / G] : 0;
    // keep the unsolved tail of the queue
    const int rest = nq - P;
    int tail = 0;
    if (c.lane < rest) tail = hq[P + c.lane];
    __syncwarp();
    if (c.lane < rest) hq[c.lane] = tail;
    nq = rest;
    __syncwarp();
    const int cslot = slot_of(c, c.central);
    const int n = c.w.snobs[cslot];
    float hx = 0.f, hy = 0.f;
    if (active) { eg3d_hit h = epcs[e]; hx = h.x; hy = h.y; }
    double X[3] = {c.w.sX[3 * cslot], c.w.sX[3 * cslot + 1], c.w.sX[3 * cslot + 2]};
    K3P_BEGIN(te1);
    K3P_ADD(c, 12, 1);
    bool ok = gn_group<GC>(S, slot_obs<GC>(c, cslot, n, true, v, hx, hy), active, G, c.lane, X);
    K3P_END(c, 4, te1);
    float Xe[3] = {(float)X[0], (float)X[1], (float)X[2]};
    unsigned m = __ballot_sync(0xffffffffu, ok && ((c.lane & (G - 1)) == 0));
    K3P_ADD(c, 29, __popc(m));                 // solves accepted in this batch
    while (m && !matched) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      K3P_ADD(c, 30, 1);                       // add_view_finish attempts from epipolar hits
      K3P_ADD(c, 31, b / G);                   // position of the attempted candidate within its batch
      eg3d_hit h = epcs[__shfl_sync(0xffffffffu, e, b)];
      float Xc[3] = {__shfl_sync(0xffffffffu, Xe[0], b), __shfl_sync(0xffffffffu, Xe[1], b), __shfl_sync(0xffffffffu, Xe[2], b)};
      Plg p; p.pl = h.polyline; p.seg = h.segment; p.c = make_float2(h.x, h.y);
      int ns = 0, ne = 0;
      const int cc = c.central;
      K3P_BEGIN(te2);
      const bool avf = add_view_finish<GC>(c, v, p, Xc, 0, cc, c.len, ns, ne);
      K3P_END(c, 5, te2);
      if (avf) {
        matched = true;
        if (ns > cc) { c.central = ns; iv0 = 0; iv1 = ns + ne; }
        else { iv0 = cc - ns; iv1 = cc + ne; }
      }
      if (c.overflow) break;
    }
    if (c.overflow) break;
  }
  st.matched = matched; st.iv0 = iv0; st.iv1 = iv1;
}
template <bool GC>
static __device__ __noinline__ void expand_view_main(Ctx& c, int v, const EvState& st) {
  const DevScene& S = *c.S;
  const bool matched = st.matched; const int iv0 = st.iv0, iv1 = st.iv1;
  int last = -1;
  K3P_BEGIN(te3);
  for (int cur = 0; cur < c.len; cur++) {
    if (matched && cur == iv0) { cur = iv1; last = iv1; continue; }
    K3P_ADD(c, 18, 1);
    const int slot = slot_of(c, cur);
    float2 q = project(S.P + 12 * v, c.w.sX[3 * slot], c.w.sX[3 * slot + 1], c.w.sX[3 * slot + 2]);
    uint32_t pl_id;
    if (!grid_unique_warp(S.g_expand, v, S.width, S.height, q, pl_id, c.lane)) continue;
    K3P_ADD(c, 19, 1);
    Pl pl = get_pl(S, v, pl_id);
    Plg init; init.pl = pl_id;
    if (pl_distancesq_warp(pl, q, init.seg, init.c, c.lane) > S.prm.max_proj_distsq_expand) return;  // abandons the view
    const int hi = matched ? (cur <= iv0 ? iv0 : c.len) : c.len;
    const int n = c.w.snobs[slot];
    double X[3] = {c.w.sX[3 * slot], c.w.sX[3 * slot + 1], c.w.sX[3 * slot + 2]};
    K3P_ADD(c, 13, 1);
    if (!gn_group<GC>(S, slot_obs<GC>(c, slot, n, true, v, init.c.x, init.c.y), true, 32, c.lane, X)) continue;
    float Xc[3] = {(float)X[0], (float)X[1], (float)X[2]};
    int ns = 0, ne = 0;
    K3P_ADD(c, 14, 1);
    if (add_view_finish<GC>(c, v, init, Xc, last + 1, cur, hi, ns, ne)) {
      K3P_ADD(c, 15, 1);
      if (ns > cur) { c.central = ns; cur = ns + ne; }
      else cur = cur + ne;
      last = cur;
    }
    if (c.overflow) return;
  }
  K3P_END(c, 6, te3);
}


// Phase A of a seed: view triple, pruned triple enumeration, PLG following.  Returns true when exactly one compatible
// hypothesis was found; its lists are left in c.w.fD1 / c.w.fD2 and the scalars in `r`.
static __device__ __noinline__ bool seed_phase_a(Ctx& c, PaRec& r) {
  const DevScene& S = *c.S; const K3Args& A = *c.A;
  const int V = S.V, lane = c.lane;
  // --- view triple (triangulation.cpp:1035-1066), chosen by k3_select_views_kernel
  const int* sl = A.sel + 3 * (size_t)c.seed;
  if (sl[0] < 0) return false;
  c.sel[0] = sl[0]; c.sel[1] = sl[1]; c.sel[2] = sl[2];
  size_t r0, r1, r2;
  if (A.a_compact) { r0 = (size_t)c.seed * 3; r1 = r0 + 1; r2 = r0 + 2; }
  else { r0 = (size_t)c.seed * V + c.sel[0]; r1 = (size_t)c.seed * V + c.sel[1]; r2 = (size_t)c.seed * V + c.sel[2]; }
  const eg3d_hit* h0 = A.hits + A.hit_off[r0]; const int n0 = (int)(A.hit_off[r0 + 1] - A.hit_off[r0]);
  const eg3d_hit* h1 = A.hits + A.hit_off[r1]; const int n1h = (int)(A.hit_off[r1 + 1] - A.hit_off[r1]);
  const eg3d_hit* h2 = A.hits + A.hit_off[r2]; const int n2h = (int)(A.hit_off[r2 + 1] - A.hit_off[r2]);
  // --- triple enumeration with the uniqueness test (triangulation.cpp:550-601): one triple per lane
  const long long total = (long long)n0 * n1h * n2h;
  bool found = false;
  int nD1 = 0, nD2 = 0, fn1 = 0, fn2 = 0;
  uint32_t d1[3] = {0, 0, 0}, d2[3] = {0, 0, 0}, fd1[3] = {0, 0, 0}, fd2[3] = {0, 0, 0};
  Cur3 fcentral; float fX[3] = {0, 0, 0};
  // Triples are visited in the reference's nested order (p0 outer, p2 inner).  Each lane first applies the exact-safe
  // 2-view bound (pair_cannot_fit) to its triple's three view pairs; survivors are queued IN ORDER and solved 32 at a
  // time, so the sequence of GN-valid hypotheses handed to plg_compatible is the reference's.
  const double Tp = prune_radius(S.prm, 3);
  const size_t i01 = (size_t)c.sel[0] * V + c.sel[1], i02 = (size_t)c.sel[0] * V + c.sel[2], i12 = (size_t)c.sel[1] * V + c.sel[2];
  const double* F01 = S.Fp + i01 * 9; const double* F02 = S.Fp + i02 * 9; const double* F12 = S.Fp + i12 * 9;
  const double h01 = S.Fph[i01], h02 = S.Fph[i02], h12 = S.Fph[i12];
  int* tq = c.w.tq;
  long long tnext = 0;
  int qn = 0;
  while (true) {
    K3P_BEGIN(tp0);
    while (qn < 32 && tnext < total) {
      const long long t = tnext + lane;
      bool pass = false;
      int i0 = 0, i1 = 0, i2 = 0;
      if (t < total) {
        if (total <= 0x7fffffffLL) {
          const unsigned tu = (unsigned)t, n12 = (unsigned)(n1h * n2h);
          i0 = (int)(tu / n12);
          unsigned rem = tu - (unsigned)i0 * n12;
          i1 = (int)(rem / (unsigned)n2h); i2 = (int)(rem - (unsigned)i1 * (unsigned)n2h);
        } else {
          i0 = (int)(t / ((long long)n1h * n2h));
          long long rem = t - (long long)i0 * n1h * n2h;
          i1 = (int)(rem / n2h); i2 = (int)(rem - (long long)i1 * n2h);
        }
        const float2 q0 = make_float2(h0[i0].x, h0[i0].y), q1 = make_float2(h1[i1].x, h1[i1].y), q2 = make_float2(h2[i2].x, h2[i2].y);
        pass = !(pair_cannot_fit(F02, h02, q0, q2, Tp) || pair_cannot_fit(F01, h01, q0, q1, Tp) || pair_cannot_fit(F12, h12, q1, q2, Tp));
      }
      const unsigned pm = __ballot_sync(0xffffffffu, pass);
      if (pass) { const int pos = qn + __popc(pm & ((1u << lane) - 1u)); tq[3 * pos] = i0; tq[3 * pos + 1] = i1; tq[3 * pos + 2] = i2; }
      qn += __popc(pm);
      tnext += 32;
    }
    K3P_END(c, 0, tp0);
    if (qn == 0) break;
    __syncwarp();
    const int take = qn < 32 ? qn : 32;
    K3P_ADD(c, 8, take); K3P_ADD(c, 9, 1);
    K3P_BEGIN(tp1);
    bool ok = false; float X[3] = {0, 0, 0};
    int i0 = 0, i1 = 0, i2 = 0;
    if (lane < take) {
      i0 = tq[3 * lane]; i1 = tq[3 * lane + 1]; i2 = tq[3 * lane + 2];
      float2 pts[3] = {make_float2(h0[i0].x, h0[i0].y), make_float2(h1[i1].x, h1[i1].y), make_float2(h2[i2].x, h2[i2].y)};
      ok = est3(S, c.sel, pts, X);
    }
    K3P_END(c, 1, tp1);
    // move the not-yet-solved tail of the queue to the front
    int r0 = 0, r1 = 0, r2 = 0;
    const int rest = qn - take;
    if (lane < rest) { r0 = tq[3 * (take + lane)]; r1 = tq[3 * (take + lane) + 1]; r2 = tq[3 * (take + lane) + 2]; }
    __syncwarp();
    if (lane < rest) { tq[3 * lane] = r0; tq[3 * lane + 1] = r1; tq[3 * lane + 2] = r2; }
    qn = rest;
    __syncwarp();
    unsigned m = __ballot_sync(0xffffffffu, ok);
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const int j0 = __shfl_sync(0xffffffffu, i0, b), j1 = __shfl_sync(0xffffffffu, i1, b), j2 = __shfl_sync(0xffffffffu, i2, b);
      float Xb[3] = {__shfl_sync(0xffffffffu, X[0], b), __shfl_sync(0xffffffffu, X[1], b), __shfl_sync(0xffffffffu, X[2], b)};
      Cur3 cand;
      eg3d_hit a0 = h0[j0], a1 = h1[j1], a2 = h2[j2];
      cand.pl[0] = a0.polyline; cand.seg[0] = a0.segment; cand.c[0] = make_float2(a0.x, a0.y);
      cand.pl[1] = a1.polyline; cand.seg[1] = a1.segment; cand.c[1] = make_float2(a1.x, a1.y);
      cand.pl[2] = a2.polyline; cand.seg[2] = a2.segment; cand.c[2] = make_float2(a2.x, a2.y);
      K3P_BEGIN(tp2);
      bool comp = plg_compatible(c, cand, nD1, nD2, d1, d2);
      K3P_END(c, 2, tp2); K3P_ADD(c, 10, 1);
      if (c.overflow) return false;
      if (comp) {
        if (found) return false;  // a second compatible triple: ambiguous, the seed yields nothing (:587-590)
        found = true;
        fn1 = nD1; fn2 = nD2;
        for (int k = 0; k < 3; k++) { fd1[k] = d1[k]; fd2[k] = d2[k]; fX[k] = Xb[k]; }
        fcentral = cand;
        __syncwarp();
        for (int i = lane; i < nD1; i += 32) c.w.fD1[i] = c.w.D1[i];
        for (int i = lane; i < nD2; i += 32) c.w.fD2[i] = c.w.D2[i];
        __syncwarp();
      }
    }
  }
  if (!found) return false;
  r.seed = c.seed; r.sel[0] = c.sel[0]; r.sel[1] = c.sel[1]; r.sel[2] = c.sel[2]; r.fn1 = fn1; r.fn2 = fn2;
  for (int k = 0; k < 3; k++) { r.fd1[k] = fd1[k]; r.fd2[k] = fd2[k]; }
  store_pt3(&r.central, fcentral, fX);
  r.pool_off = 0;
  return true;
}

// Phase B of an accepted seed, set-up: chain = reverse(D1) + central + D2.  The expansion to every other view
// (triangulation.cpp:960-973) is the kernel's view loop.
static __device__ __noinline__ void seed_phase_b(Ctx& c, const PaRec& r, const Pt3* l1, const Pt3* l2) {
  const DevScene& S = *c.S;
  const int V = S.V, lane = c.lane;
  const int fn1 = r.fn1, fn2 = r.fn2;
  c.sel[0] = r.sel[0]; c.sel[1] = r.sel[1]; c.sel[2] = r.sel[2];
  // --- chain = reverse(D1) + central + D2 (polyline_graph_2d.cpp:1298-1306)
  c.len = fn1 + 1 + fn2;
  if (c.len > c.w.capc) { c.overflow = true; return; }
  __syncwarp();
  for (int i = lane; i < c.len; i += 32) {     // one chain point per lane
    const Pt3& p = i < fn1 ? l1[fn1 - 1 - i] : (i == fn1 ? r.central : l2[i - fn1 - 1]);
    const size_t b = (size_t)i * c.w.oc;
    for (int k = 0; k < 3; k++) {
      c.w.ov[b + k] = c.sel[p.perm[k]]; c.w.opl[b + k] = p.pl[k]; c.w.oseg[b + k] = p.seg[k]; c.w.ox[b + k] = p.cx[k]; c.w.oy[b + k] = p.cy[k];
    }
    c.w.snobs[i] = 3;
    c.w.sX[3 * i] = p.X[0]; c.w.sX[3 * i + 1] = p.X[1]; c.w.sX[3 * i + 2] = p.X[2];
  }
  for (int i = lane; i < c.len; i += 32) c.w.order[i] = i;
  for (int v = lane; v < V; v += 32) { c.w.sdirs[v] = 0; c.w.edirs[v] = 0; }
  __syncwarp();
  if (lane == 0) for (int k = 0; k < 3; k++) { c.w.sdirs[c.sel[k]] = r.fd1[k]; c.w.edirs[c.sel[k]] = r.fd2[k]; }
  __syncwarp();
  if (lane < 4) c.w.fbm[lane] = -1;
  if (c.A->gn_cache) for (int i = lane; i <= c.w.capc; i += 32) c.w.gtag[i] = -1;
  __syncwarp();
  c.nslots = c.len;
  c.central = fn1;
}

__global__ void __launch_bounds__(K3_THREADS, EG3D_K3A_MIN_BLOCKS) k3a_hypothesis_kernel(const __grid_constant__ DevScene S, const __grid_constant__ K3Args A) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  Ctx c;
  c.S = &S; c.A = &A; c.lane = lane;
  c.w = make_ws(A.scratch + (size_t)warp * A.scratch_per_warp, S.V, A.capf, A.capc, A.oc);
  while (true) {
    int seed = 0;
    if (lane == 0) seed = atomicAdd(A.work_counter, 1);
    seed = __shfl_sync(0xffffffffu, seed, 0);
    if (seed >= A.n_seeds) break;
    c.seed = seed; c.sv = A.seed_view[seed];
    c.len = 0; c.nslots = 0; c.central = 0; c.overflow = false;
#ifdef EG3D_K3_PROFILE
    for (int k = 0; k < 32; k++) c.pc[k] = 0;
#endif
    K3P_BEGIN(tall);
    PaRec r;
    const bool found = seed_phase_a(c, r);
    K3P_END(c, 7, tall);
    __syncwarp();
    if (c.overflow) { if (lane == 0) atomicAdd(&A.out_counters[2], 1ull); }
    else if (found) {
      unsigned long long ri = 0, po = 0;
      if (lane == 0) { ri = atomicAdd(&A.pa_counters[0], 1ull); po = atomicAdd(&A.pa_counters[1], (unsigned long long)(r.fn1 + r.fn2)); }
      ri = __shfl_sync(0xffffffffu, ri, 0); po = __shfl_sync(0xffffffffu, po, 0);
      if ((long long)po + r.fn1 + r.fn2 > A.pa_pool_cap) { if (lane == 0) atomicAdd(&A.out_counters[3], 1ull); r.fn1 = -1; }
      r.pool_off = (long long)po;
      if (lane == 0) { A.pa_recs[ri] = r; A.acc_seed[ri] = seed; }
      if (r.fn1 >= 0) {
        for (int i = lane; i < r.fn1; i += 32) A.pa_pool[po + i] = c.w.fD1[i];
        for (int i = lane; i < r.fn2; i += 32) A.pa_pool[po + r.fn1 + i] = c.w.fD2[i];
      }
    }
#ifdef EG3D_K3_PROFILE
    if (A.prof && lane == 0) for (int k = 0; k < 32; k++) if (c.pc[k]) atomicAdd(&A.prof[k], (unsigned long long)c.pc[k]);
#endif
    __syncwarp();
  }
}

// View triple of every seed (triangulation.cpp:1035-1066): first view with hits, the starting view (or, when that is the
// first or last one, the middle non-empty view), last view with hits.  One warp per seed.  The "has hits" predicate comes
// from the any-hit flags of the lazy sweep (flags != null) or from the full (seed, view) CSR.
__global__ void k3_select_views_kernel(int n_seeds, int V, const int* __restrict__ seed_view, const unsigned char* __restrict__ flags,
                                       const int64_t* __restrict__ off, int* __restrict__ sel) {
  const int lane = threadIdx.x & 31;
  const int seed = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (seed >= n_seeds) return;
  const size_t row = (size_t)seed * V;
  auto nonempty = [&](int v) { return v < V && (flags ? flags[row + v] != 0 : off[row + v + 1] > off[row + v]); };
  int ne = 0, minv = -1, maxv = -1;
  for (int base = 0; base < V; base += 32) {
    const unsigned m = __ballot_sync(0xffffffffu, nonempty(base + lane));
    if (m) { if (minv < 0) minv = base + __ffs(m) - 1; maxv = base + 31 - __clz(m); ne += __popc(m); }
  }
  int s0 = -1, s1 = -1, s2 = -1;
  if (ne >= 3) {
    int mid = 0;
    const int want = ne / 2;
    int seen = 0;
    for (int base = 0; base < V; base += 32) {
      unsigned m = __ballot_sync(0xffffffffu, nonempty(base + lane));
      const int pc = __popc(m);
      if (seen + pc > want) {
        for (int k = 0; k < want - seen; k++) m &= m - 1;
        mid = base + __ffs(m) - 1;
        break;
      }
      seen += pc;
    }
    const int sv = seed_view[seed];
    s0 = minv; s1 = (sv == minv || sv == maxv) ? mid : sv; s2 = maxv;
  }
  if (lane == 0) { sel[3 * (size_t)seed] = s0; sel[3 * (size_t)seed + 1] = s1; sel[3 * (size_t)seed + 2] = s2; }
}

// Hands a finished chain to the unordered output: one contiguous range of points / observations per seed.
static __device__ __noinline__ void emit_chain(Ctx& c, int seed) {
  const K3Args& A = *c.A;
  const int lane = c.lane;
  __syncwarp();
  int npts = c.overflow ? 0 : c.len;
  long long nobs = 0;
  for (int i = lane; i < npts; i += 32) nobs += c.w.snobs[c.w.order[i]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nobs += __shfl_xor_sync(0xffffffffu, nobs, o);
  unsigned long long pbase = 0, obase = 0;
  if (lane == 0) {
    if (c.overflow) atomicAdd(&A.out_counters[2], 1ull);
    if (npts > 0) {
      pbase = atomicAdd(&A.out_counters[0], (unsigned long long)npts);
      obase = atomicAdd(&A.out_counters[1], (unsigned long long)nobs);
    }
  }
  pbase = __shfl_sync(0xffffffffu, pbase, 0); obase = __shfl_sync(0xffffffffu, obase, 0);
  if (npts > 0 && ((long long)pbase + npts > A.pt_cap || (long long)obase + nobs > A.ob_cap)) {
    if (lane == 0) atomicAdd(&A.out_counters[3], 1ull);
    npts = 0; nobs = 0;
  }
  if (lane == 0) { A.seed_npts[seed] = npts; A.seed_pbase[seed] = (int64_t)pbase; A.seed_nobs[seed] = nobs; }
  // write the chain: point headers by lane 0, observations coalesced
  long long ob = (long long)obase;
  for (int i = 0; i < npts; i++) {
    const int slot = c.w.order[i];
    const int n = c.w.snobs[slot];
    if (lane == 0) {
      A.o_X[3 * (pbase + i)] = c.w.sX[3 * slot]; A.o_X[3 * (pbase + i) + 1] = c.w.sX[3 * slot + 1]; A.o_X[3 * (pbase + i) + 2] = c.w.sX[3 * slot + 2];
      A.o_nobs[pbase + i] = n; A.o_obase[pbase + i] = ob;
    }
    const size_t b = (size_t)slot * c.w.oc;
    for (int k = lane; k < n; k += 32) {
      A.ob_view[ob + k] = c.w.ov[b + k]; A.ob_pl[ob + k] = c.w.opl[b + k]; A.ob_seg[ob + k] = c.w.oseg[b + k];
      A.ob_x[ob + k] = c.w.ox[b + k]; A.ob_y[ob + k] = c.w.oy[b + k];
    }
    ob += n;
  }
  __syncwarp();
}

// Work order of phase B.  Per-seed cost grows with the chain length (every view is tried against every chain point), and
// a few hundred seeds are 10-100x the median, so the accepted seeds are visited longest chain first; the order has no
// effect on the results (pack orders by seed).  Keys for the radix sort: descending length, unused tail entries last.
__global__ void k3_order_keys_kernel(int n, const unsigned long long* __restrict__ pa_counters, const PaRec* __restrict__ recs,
                                     unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned key = 0xffffffffu;
  if ((unsigned long long)i < pa_counters[0]) {
    const PaRec& r = recs[i];
    const int len = r.fn1 < 0 ? 0 : r.fn1 + r.fn2 + 1;
    key = 0xffffu - (unsigned)min(len, 0xffff);
  }
  keys[i] = key; vals[i] = i;
}

#if EG3D_K3B_BATCH > 0
// Lock-step phase B with EG3D_K3B_BATCH seeds per warp: one CTA per SM, every warp owns a small batch of accepted seeds
// (one scratch arena each, the few per-seed scalars parked in shared memory) and runs one half of a view's expansion for
// all of them before the CTA-wide barrier.  The warps of an SM then execute the same ~half of the hot code at the same
// time and every line they pull in is used by batch x warps seeds, instead of each warp cycling through the whole
// ~65 KB path on its own (profiles/r01_k3b_icache.md).
struct SeedState { int seed, sv, sel0, sel1, sel2, len, nslots, central, overflow, live, ran, matched, iv0, iv1, has; long long hrow; };
template <bool GC>
__global__ void __launch_bounds__(K3B_THREADS, 1) k3b_expand_kernel(const __grid_constant__ DevScene S, const __grid_constant__ K3Args A) {
  constexpr int KB = EG3D_K3B_BATCH;
  __shared__ SeedState ss[K3B_THREADS / 32][KB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  Ctx c;
  c.S = &S; c.A = &A; c.lane = lane;
  const int n_acc = (int)A.pa_counters[0];
  const int V = S.V;
  auto load = [&](int j) {
    const SeedState& q = ss[wib][j];
    c.seed = q.seed; c.sv = q.sv; c.sel[0] = q.sel0; c.sel[1] = q.sel1; c.sel[2] = q.sel2;
    c.len = q.len; c.nslots = q.nslots; c.central = q.central; c.overflow = q.overflow != 0; c.hrow = q.hrow;
    c.w = make_ws(A.scratch + ((size_t)warp * KB + j) * A.scratch_per_warp, V, A.capf, A.capc, A.oc);
  };
  auto store = [&](int j) {
    __syncwarp();
    if (lane == 0) { SeedState& q = ss[wib][j]; q.len = c.len; q.nslots = c.nslots; q.central = c.central; q.overflow = c.overflow ? 1 : 0; }
    __syncwarp();
  };
#ifdef EG3D_K3_PROFILE
  for (int k = 0; k < 32; k++) c.pc[k] = 0;
#endif
  while (true) {
    bool any = false;
    for (int j = 0; j < KB; j++) {
      int ri = 0;
      if (lane == 0) ri = atomicAdd(A.work_counter, 1);
      ri = __shfl_sync(0xffffffffu, ri, 0);
      const bool has = ri < n_acc;
      any |= has;
      __syncwarp();
      if (lane == 0) { SeedState& q = ss[wib][j]; q.has = has ? 1 : 0; q.live = 0; q.overflow = 0; q.len = 0; q.nslots = 0; q.central = 0; q.seed = 0; q.sel0 = q.sel1 = q.sel2 = -1; }
      __syncwarp();
      if (!has) continue;
      const int rank = A.pa_order ? A.pa_order[ri] : ri;
      const PaRec& r = A.pa_recs[rank];
      c.seed = r.seed; c.sv = A.seed_view[r.seed];
      c.hrow = (long long)(A.b_compact ? rank : r.seed) * V;
      c.len = 0; c.nslots = 0; c.central = 0; c.overflow = false;
      c.sel[0] = c.sel[1] = c.sel[2] = -1;
      c.w = make_ws(A.scratch + ((size_t)warp * KB + j) * A.scratch_per_warp, V, A.capf, A.capc, A.oc);
      bool live = false;
      if (r.fn1 >= 0) { seed_phase_b(c, r, A.pa_pool + r.pool_off, A.pa_pool + r.pool_off + r.fn1); live = !c.overflow; }
      __syncwarp();
      if (lane == 0) {
        SeedState& q = ss[wib][j];
        q.seed = c.seed; q.sv = c.sv; q.sel0 = c.sel[0]; q.sel1 = c.sel[1]; q.sel2 = c.sel[2]; q.hrow = c.hrow; q.live = live ? 1 : 0;
        q.len = c.len; q.nslots = c.nslots; q.central = c.central; q.overflow = c.overflow ? 1 : 0;
      }
      __syncwarp();
    }
    if (!__syncthreads_or(any)) break;
    // --- expansion to every other view (triangulation.cpp:960-973): the CTA's seeds, view by view, half by half
    for (int v = 0; v < V; v++) {
      __syncthreads();
      for (int j = 0; j < KB; j++) {
        const SeedState& q = ss[wib][j];
        const bool run = q.live && !q.overflow && v != q.sel0 && v != q.sel1 && v != q.sel2;
        if (run) {
          load(j);
          EvState st; st.matched = false; st.iv0 = 0; st.iv1 = 0;
          expand_view_epc<false>(c, v, st);
          store(j);
          if (lane == 0) { SeedState& w = ss[wib][j]; w.matched = st.matched ? 1 : 0; w.iv0 = st.iv0; w.iv1 = st.iv1; }
        }
        __syncwarp();
        if (lane == 0) ss[wib][j].ran = run ? 1 : 0;
        __syncwarp();
      }
      __syncthreads();
      for (int j = 0; j < KB; j++) {
        const SeedState& q = ss[wib][j];
        if (q.ran && !q.overflow) {
          load(j);
          EvState st; st.matched = q.matched != 0; st.iv0 = q.iv0; st.iv1 = q.iv1;
          expand_view_main<false>(c, v, st);
          store(j);
        }
      }
    }
    for (int j = 0; j < KB; j++) {
      if (!ss[wib][j].has) continue;
      load(j);
      emit_chain(c, c.seed);
    }
    __syncwarp();
  }
}
#else
// GC: the variant with the shared first Gauss-Newton iteration (gn_group); the host picks it for rigs with many views
template <bool GC>
__global__ void __launch_bounds__(K3B_THREADS, EG3D_K3B_MIN_BLOCKS) k3b_expand_kernel(const __grid_constant__ DevScene S, const __grid_constant__ K3Args A) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  Ctx c;
  c.S = &S; c.A = &A; c.lane = lane;
  c.w = make_ws(A.scratch + (size_t)warp * A.scratch_per_warp, S.V, A.capf, A.capc, A.oc);
  const int n_acc = (int)A.pa_counters[0];
  const int V = S.V;
#ifdef EG3D_K3_PROFILE
  unsigned long long t_warp0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_warp0));
#endif
  while (true) {
    int ri = 0;
    if (lane == 0) ri = atomicAdd(A.work_counter, 1);
    ri = __shfl_sync(0xffffffffu, ri, 0);
    const bool has = ri < n_acc;
#ifdef EG3D_K3_PROFILE
    if (!has && A.prof && lane == 0) {        // tail: when this warp ran out of seeds (ns since it started): sum / max over warps
      unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      atomicAdd(&A.prof[43], t1 - t_warp0); atomicMax(&A.prof[44], t1 - t_warp0); atomicAdd(&A.prof[45], 1ull);
    }
#endif
#if EG3D_K3B_SYNC
    if (!__syncthreads_or(has)) break;       // the CTA's warps take their seeds together and leave together
#else
    if (!has) break;
#endif
    int seed = 0;
    c.len = 0; c.nslots = 0; c.central = 0; c.overflow = false;
    c.sel[0] = c.sel[1] = c.sel[2] = -1;
#ifdef EG3D_K3_PROFILE
    for (int k = 0; k < 32; k++) c.pc[k] = 0;
#endif
    K3P_BEGIN(tall);
    bool live = false;
#ifdef EG3D_K3_PROFILE
    int len0 = 0;
#endif
    if (has) {
      const int rank = A.pa_order ? A.pa_order[ri] : ri;
      const PaRec& r = A.pa_recs[rank];
      seed = r.seed;
      c.seed = seed; c.sv = A.seed_view[seed];
      c.hrow = (long long)(A.b_compact ? rank : seed) * V;
      if (r.fn1 >= 0) {
        seed_phase_b(c, r, A.pa_pool + r.pool_off, A.pa_pool + r.pool_off + r.fn1);
        live = !c.overflow;
#ifdef EG3D_K3_PROFILE
        len0 = r.fn1 + r.fn2 + 1;
#endif
      }
    }
    // --- expansion to every other view (triangulation.cpp:960-973), the CTA's seeds view by view
    for (int v = 0; v < V; v++) {
      const bool run = live && !c.overflow && v != c.sel[0] && v != c.sel[1] && v != c.sel[2];
      EvState st; st.matched = false; st.iv0 = 0; st.iv1 = 0;
#if EG3D_K3B_SYNC
      __syncthreads();
#endif
      if (run) { K3P_ADD(c, 24, 1); expand_view_epc<GC>(c, v, st); if (st.matched) K3P_ADD(c, 25, 1); K3P_ADD(c, 26, c.len); }
#if EG3D_K3B_SYNC
      __syncthreads();
#endif
      if (run && !c.overflow) expand_view_main<GC>(c, v, st);
    }
    K3P_END(c, 7, tall);
    if (!has) continue;
#ifdef EG3D_K3_PROFILE
    if (A.prof && lane == 0) {
      for (int k = 0; k < 32; k++) if (c.pc[k]) atomicAdd(&A.prof[k], (unsigned long long)c.pc[k]);
      atomicMax(&A.prof[40], (unsigned long long)c.pc[7]);
      if (c.pc[7] > 50000000ll) atomicAdd(&A.prof[41], 1ull);
      if (c.pc[7] > 200000000ll) atomicAdd(&A.prof[42], 1ull);
      if (A.prof_seed) { A.prof_seed[3 * (size_t)seed] = (unsigned long long)c.pc[7]; A.prof_seed[3 * (size_t)seed + 1] = (unsigned long long)len0 | ((unsigned long long)c.len << 32);
                         A.prof_seed[3 * (size_t)seed + 2] = (unsigned long long)(A.hit_off_b[c.hrow + V] - A.hit_off_b[c.hrow]); }
    }
#endif
    emit_chain(c, seed);
  }
}

#endif  // EG3D_K3B_BATCH

}  // namespace eg3d
