// walk_conv.cuh -- argument blocks of the walk kernels and the warp-convergent alias-fold kernel (K6 v5).
//
// Included by walk.cu (inside its anonymous namespace) and, with SRW_EMU defined, by tests/emu/ where the
// same source is compiled for the host with one-lane "warps" so that the kernel's decisions can be checked
// against the CPU twin (oracle_alias_walk) without a GPU.  Everything CUDA-specific the kernel touches is
// listed in tests/emu/cuda_emu.h.
#pragma once
#include <stdint.h>

#include "layout.h"
#include "philox.cuh"

struct WalkArgs {
  const int64_t *__restrict__ off;
  const int32_t *__restrict__ col;        // sorted rows
  const AliasSlot *__restrict__ slot;     // Vose slots (weighted) or nullptr
  const int32_t *__restrict__ col_app;    // appearance-order rows (exact sampler)
  const float *__restrict__ w_app;
  const int32_t *__restrict__ vids;
  int64_t nv;
  uint64_t walker_first;
  int64_t n_walkers;
  int32_t stride;                         // walk_length + 2 (RW:103)
  uint32_t seed_lo, seed_hi;
  uint64_t t_ret, t_common, t_far;        // alias acceptance thresholds
  float p, q;                             // exact sampler (RW:112-113 .toFloat)
  int32_t u_mode;
  float u_const;
  int32_t *paths;
  int32_t *lens;
  unsigned long long *stats;              // [steps, proposals, member_tests, probes_log2]
};

enum : int { ST_EXTENT = 0, ST_PROPOSE = 1, ST_SEARCH = 2, ST_DONE = 3, ST_HASH = 4, ST_WAIT = 5 /* v5: a draw is pending */ };

__device__ __forceinline__ int ceil_log2_p1(int64_t d) { return d <= 0 ? 0 : 64 - __clzll(d); }

struct FoldArgs {
  const NbrEntry *__restrict__ ent;
  const int32_t *__restrict__ hash;
  double a, mp;              // a = 1/p - Mp > 0, Mp = max(1, 1/q)   (a = 0: no return component, plain rejection under mp)
  uint64_t t_ret, t_common, t_far;  // thresholds under the envelope (t_ret = 2^32 when the return edge is folded out)
};

// Alias-fold arguments from (p, q) as the walk narrows them (RW:112-113 .toFloat): envelope Mp = max(1, 1/q),
// a = 1/p - Mp, thresholds T(f) = f >= Mp ? 2^32 : floor(f/Mp * 2^32).  Returns true when folding applies
// (a > 0 and the caller asked for it); otherwise the caller installs the classic thresholds with a = 0.
inline bool srw_fold_args(double p, double q, bool want_fold, FoldArgs *f) {
  const double inv_p = 1.0 / (double)(float)p, inv_q = 1.0 / (double)(float)q;
  const double M = inv_q > 1.0 ? inv_q : 1.0;
  auto thr = [M](double v) -> uint64_t { return v >= M ? 4294967296ULL : (uint64_t)((v / M) * 4294967296.0); };
  f->a = inv_p - M; f->mp = M; f->t_ret = 4294967296ULL; f->t_common = thr(1.0); f->t_far = thr(inv_q);
  return f->a > 0.0 && want_fold;
}

// Peer-gather mode: the graph is cut into vertex ranges, shard s lives in the HBM of GPU s, and
// every GPU can address every shard (NVLink peer mappings).  A neighbour entry names the owner of the
// neighbour's row, so the kernel picks the base pointer per access and a walker never migrates: the
// "exchange" of the reference's super-step shuffle (RW:186-192) becomes 16/32-byte loads over NVLink.
struct PeerTable {
  int world;
  int64_t first[SRW_MAX_SHARDS + 1];          // first rank of every shard
  const int64_t *off[SRW_MAX_SHARDS];         // shard-local row offsets
  const NbrEntry *ent[SRW_MAX_SHARDS];
  const int32_t *hash[SRW_MAX_SHARDS];
};

constexpr int kStage = 16;

// ------------------------------------------------------------------------------------------
// K6 (v5): the same sampler, the same bits, laid out for WARP CONVERGENCE.  ncu on v4 (RMAT-26): 7.4 of 32
// lanes active per issued instruction and 60 % of the issue slots busy -- the compiler threads the arms of
// the v4 state machine straight into the draw code, so the ~70-instruction Philox block runs several times
// per iteration, each time for the few lanes that arrived together.  v5 gives every iteration three
// phases separated by warp barriers, and no lane leaves the loop before its warp does:
//   A  draw      every lane whose step/trial needs fresh bits runs Philox ONCE, together; the return-excess
//                component is resolved here (no memory access: the lane draws again next iteration);
//   B  load      one 16-byte gather per lane (neighbour entry, or the entry probed by a short-row search), or
//                one 32-byte bucket as a single 256-bit load -- the address is selected, the load is shared;
//   C  consume   accept / reject / next probe; the single push site and the single flush site follow.
// Row offsets are kept as u32 (the ABI caps nnz below 2^32).  VAR bit 0: loads carry L2::64B (the L2 fills
// 64 instead of 128 bytes per missing gather, profiles/README.md "gather_probe"); bit 1 (with bit 0): the gather does not allocate in L1.
// ------------------------------------------------------------------------------------------

template <int VAR>
__device__ __forceinline__ int4 gather16(const int4 *p) {
#ifdef SRW_EMU
  return *p;
#else
  int4 v;
  if ((VAR & 3) == 3) asm("ld.global.nc.L1::no_allocate.L2::64B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if (VAR & 1) asm("ld.global.nc.L2::64B.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else v = __ldg(p);
  return v;
#endif
}
template <int VAR>
__device__ __forceinline__ void gather32(const int4 *p, int4 &lo, int4 &hi) {   // p is 32-byte aligned: one 256-bit load
#ifdef SRW_EMU
  lo = p[0]; hi = p[1];
#else
  if ((VAR & 3) == 3)
    asm("ld.global.nc.L1::no_allocate.L2::64B.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(p));
  else if (VAR & 1)
    asm("ld.global.nc.L2::64B.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(p));
  else
    asm("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(p));
#endif
}

template <bool STATS, bool PEER, int VAR, int MINB = 4, bool IDS = false>
__global__ void __launch_bounds__(256, MINB) walk_fold_conv_kernel(WalkArgs a, FoldArgs f, const PeerTable pt) {
  __shared__ int32_t sbuf[kStage * 256];
  __shared__ const NbrEntry *s_ent[SRW_MAX_SHARDS];
  __shared__ const int32_t *s_hash[SRW_MAX_SHARDS];
  const int tid = threadIdx.x;
  if (PEER) {
    if (tid < SRW_MAX_SHARDS) { s_ent[tid] = pt.ent[tid]; s_hash[tid] = pt.hash[tid]; }
    __syncthreads();
  }
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + tid;
  const bool live = i < a.n_walkers;                     // dead lanes stay with their warp (full-mask barriers below)
  const uint64_t walker = a.walker_first + (uint64_t)(live ? i : 0);
  int32_t curr = (int32_t)(walker % (uint64_t)a.nv), prev = -1;
  int32_t *path = a.paths + (live ? i : 0) * a.stride;
  int32_t len = 0, staged = 0, flushed = 0;
  // the first flush stops at a 16-byte boundary of the row, every later one is whole int4 stores
  const int head = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(path) & 15u)) & 15u) >> 2);
  int lim = head ? kStage - 4 + head : kStage;
  auto flush = [&]() {
    int32_t *dst = path + flushed;
    if (staged == kStage && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {        // the common case: four 16-byte stores
#pragma unroll
      for (int j = 0; j < kStage; j += 4)
        *reinterpret_cast<int4 *>(dst + j) = make_int4(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid], sbuf[(j + 2) * 256 + tid], sbuf[(j + 3) * 256 + tid]);
    } else {
      int j = 0;
      if ((reinterpret_cast<uintptr_t>(dst) & 4) && staged >= 1) { dst[0] = sbuf[tid]; j = 1; }          // -> 8-byte aligned
      if ((reinterpret_cast<uintptr_t>(dst + j) & 8) && j + 1 < staged) {                                 // -> 16-byte aligned
        *reinterpret_cast<int2 *>(dst + j) = make_int2(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid]);
        j += 2;
      }
#pragma unroll 1
      for (; j + 3 < staged; j += 4)
        *reinterpret_cast<int4 *>(dst + j) = make_int4(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid], sbuf[(j + 2) * 256 + tid], sbuf[(j + 3) * 256 + tid]);
#pragma unroll 1
      for (; j + 1 < staged; j += 2) *reinterpret_cast<int2 *>(dst + j) = make_int2(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid]);
#pragma unroll 1
      for (; j < staged; ++j) dst[j] = sbuf[j * 256 + tid];
    }
    flushed += staged; staged = 0; lim = kStage;
  };
  uint32_t off = 0, poff = 0, xoff = 0, k = 0;
  uint32_t deg = 0, pdeg = 0, m = 1, xdeg = 0, xm = 1, trial = 0, lo = 0, hi = 0, y = 0, bkt = 0, pnb = 0;
  uint32_t cown = 0, pown = 0, xown = 0;   // PEER: shards that hold the rows of curr / prev / x
  int32_t x = 0;
  double ret_lhs = 0.0, ret_rhs = 0.0;
  unsigned long long n_prop = 0, n_mem = 0, n_log = 0;
  const uint64_t t_lo = f.t_common < f.t_far ? f.t_common : f.t_far;
  const uint64_t t_hi = f.t_common < f.t_far ? f.t_far : f.t_common;
  int state = ST_DONE;
  if (live) {
    int64_t e0, e1;                                             // the start vertex's extent: the only row-offset load
    if (PEER) {
      while ((int)cown + 1 < pt.world && (int64_t)curr >= pt.first[cown + 1]) cown++;
      const int64_t *o = pt.off[cown] + ((int64_t)curr - pt.first[cown]);
      e0 = __ldg(o); e1 = __ldg(o + 1);
    } else {
      e0 = __ldg(a.off + curr); e1 = __ldg(a.off + curr + 1);
    }
    off = (uint32_t)e0; deg = (uint32_t)(e1 - e0);
    if (IDS) curr = __ldg(a.vids + curr);                       // id space: from here on a vertex is its original id (entries and hash sets carry ids)
    sbuf[tid] = curr; staged = 1; len = 1;
    if (deg != 0 && len != a.stride) state = ST_WAIT;           // dead end (RW:59-62) / RW:103
  }

  while (__any_sync(0xffffffffu, state != ST_DONE)) {
    int32_t newv = 0;
    bool moved = false;
    // ---- A: draw ----
    if (state == ST_WAIT) {
      if (trial == 0 && len > 1) {                             // per step: P(return-excess component) = a*m / (Mp*deg + a*m)
        const double t1 = __dmul_rn(f.a, (double)m), t2 = __dmul_rn(f.mp, (double)deg);
        ret_lhs = __dadd_rn(t2, t1);
        ret_rhs = __dmul_rn(t1, 4294967296.0);
      }
      const Philox4 r = walker_rng(a.seed_lo, a.seed_hi, walker, (uint32_t)(len - 1), trial);
      if (len > 1 && __dmul_rn((double)r.y, ret_lhs) < ret_rhs) {   // return-excess component: always accepted, no memory access
        if (STATS) n_prop++;
        newv = prev; moved = true;
        const int32_t c = curr; curr = prev; prev = c;
        const uint32_t o = off; off = poff; poff = o;
        const uint32_t d = deg; deg = pdeg; pdeg = d;          // m unchanged: the same bundle of parallel edges
        const uint32_t w = cown; cown = pown; pown = w;
      } else {
        k = (uint32_t)__umul64hi(((uint64_t)r.x << 32) | (uint64_t)r.w, (uint64_t)deg);
        y = r.z;
        state = ST_PROPOSE;
      }
    }
    __syncwarp();
    // ---- B: one memory access per lane ----
    const int4 *P = nullptr;
    if (!moved) {
      if (state == ST_PROPOSE) P = reinterpret_cast<const int4 *>((PEER ? s_ent[cown] : f.ent) + ((uint64_t)off + k));
      else if (state == ST_HASH) P = reinterpret_cast<const int4 *>((PEER ? s_hash[pown] : f.hash) + ((uint64_t)(poff >> 2) + bkt) * 8);
      else if (state == ST_SEARCH) P = reinterpret_cast<const int4 *>((PEER ? s_ent[pown] : f.ent) + ((uint64_t)poff + ((lo + hi) >> 1)));
    }
    int4 q0 = make_int4(0, 0, 0, 0), q1 = make_int4(0, 0, 0, 0);
    if (state == ST_HASH) gather32<VAR>(P, q0, q1);
    else if (P) q0 = gather16<VAR>(P);
    __syncwarp();
    // ---- C: consume it ----
    int verdict = 0;           // 1 = accept entry x, 2 = reject (next trial)
    if (P) {
      if (state == ST_PROPOSE) {
        x = q0.x; xdeg = (uint32_t)q0.y; xoff = (uint32_t)q0.z;
        xown = (uint32_t)q0.w & 0xFFu;
        xm = (uint32_t)q0.w >> 8;
        if (STATS && len > 1) n_prop++;
        if (len == 1 || deg == 1) verdict = 1;                   // first-order step (RW:57) / single choice
        else if (x == prev) verdict = ((uint64_t)y < f.t_ret) ? 1 : 2;   // RS:36; folded: mass Mp of Mp, t_ret = 2^32
        else if ((uint64_t)y < t_lo) verdict = 1;
        else if ((uint64_t)y >= t_hi) verdict = 2;
        else {
          if (STATS) { n_mem++; n_log += ceil_log2_p1(pdeg); }
          pnb = srw_hash_buckets((int64_t)poff, pdeg);
          if (pnb) { bkt = __umulhi(srw_hash32((uint32_t)x), pnb); state = ST_HASH; }
          else { lo = 0; hi = pdeg; state = ST_SEARCH; }
        }
      } else if (state == ST_HASH) {
        const bool found = q0.x == x || q0.y == x || q0.z == x || q0.w == x || q1.x == x || q1.y == x || q1.z == x || q1.w == x;
        if (found) verdict = ((uint64_t)y < f.t_common) ? 1 : 2;           // RS:38
        else if (q1.w == -1) verdict = ((uint64_t)y < f.t_far) ? 1 : 2;    // RS:34
        else bkt = bkt + 1 == pnb ? 0 : bkt + 1;
      } else {
        const uint32_t mid = (lo + hi) >> 1;
        if (q0.x == x) verdict = ((uint64_t)y < f.t_common) ? 1 : 2;
        else {
          if (q0.x < x) lo = mid + 1; else hi = mid;
          if (lo >= hi) verdict = ((uint64_t)y < f.t_far) ? 1 : 2;
        }
      }
    }
    if (verdict == 1) {                                        // move along entry (x, xoff, xdeg, xm)
      newv = x; moved = true;
      prev = curr; poff = off; pdeg = deg; pown = cown;
      curr = x; off = xoff; deg = xdeg; m = xm; cown = xown;
    } else if (verdict == 2) {
      trial++;
      state = ST_WAIT;
    }
    if (moved) {                                               // RW:114, then RW:103 / RW:115-119
      sbuf[staged * 256 + tid] = newv;
      staged++; len++;
      if (staged == lim) flush();
      trial = 0;
      state = (len == a.stride || deg == 0) ? ST_DONE : ST_WAIT;
    }
  }
  if (live) {
    flush();
    a.lens[i] = len;
  }
  if (STATS) {
    atomicAdd(a.stats + 1, n_prop);
    atomicAdd(a.stats + 2, n_mem);
    atomicAdd(a.stats + 3, n_log);
  }
}


// ------------------------------------------------------------------------------------------
// K6 (v5, classic alias sampler: weighted or directed graphs, or (p, q) where folding does not apply).
// The decisions of walk_alias_hash_kernel (v3) in the convergent three-phase layout of walk_fold_conv_kernel:
//   A  draw      lanes in ST_WAIT run Philox together (draw(walker, step, trial), then trial++);
//   B  load      ST_EXTENT and ST_HASH take one 256-bit load (row descriptor / hash bucket), ST_PROPOSE one
//                16-byte Vose slot (weighted) or one 4-byte neighbour id, ST_SEARCH one 4-byte id;
//   C  consume   verdicts, the single push site (ids staged in shared memory, 16-byte stores).
// The hash set of prev is addressed from (poff, pdeg) alone (layout.h), so a lane carries two row extents and
// nothing else.  Same bits as v3 and as the CPU twin (oracle_alias_walk, fold = 0).
// ------------------------------------------------------------------------------------------
template <bool HAS_ALIAS, bool STATS, int VAR>
__global__ void __launch_bounds__(256, 4) walk_alias_conv_kernel(WalkArgs a, const RowMeta *__restrict__ meta,
                                                                 const int32_t *__restrict__ hash) {
  __shared__ int32_t sbuf[kStage * 256];
  const int tid = threadIdx.x;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + tid;
  const bool live = i < a.n_walkers;
  const uint64_t walker = a.walker_first + (uint64_t)(live ? i : 0);
  int32_t curr = (int32_t)(walker % (uint64_t)a.nv), prev = -1;
  int32_t *path = a.paths + (live ? i : 0) * a.stride;
  int32_t len = 0, staged = 0, flushed = 0;
  const int head = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(path) & 15u)) & 15u) >> 2);
  int lim = head ? kStage - 4 + head : kStage;
  auto flush = [&]() {
    int32_t *dst = path + flushed;
    if (staged == kStage && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < kStage; j += 4)
        *reinterpret_cast<int4 *>(dst + j) = make_int4(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid], sbuf[(j + 2) * 256 + tid], sbuf[(j + 3) * 256 + tid]);
    } else {
      int j = 0;
      if ((reinterpret_cast<uintptr_t>(dst) & 4) && staged >= 1) { dst[0] = sbuf[tid]; j = 1; }
      if ((reinterpret_cast<uintptr_t>(dst + j) & 8) && j + 1 < staged) {
        *reinterpret_cast<int2 *>(dst + j) = make_int2(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid]);
        j += 2;
      }
#pragma unroll 1
      for (; j + 3 < staged; j += 4)
        *reinterpret_cast<int4 *>(dst + j) = make_int4(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid], sbuf[(j + 2) * 256 + tid], sbuf[(j + 3) * 256 + tid]);
#pragma unroll 1
      for (; j + 1 < staged; j += 2) *reinterpret_cast<int2 *>(dst + j) = make_int2(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid]);
#pragma unroll 1
      for (; j < staged; ++j) dst[j] = sbuf[j * 256 + tid];
    }
    flushed += staged; staged = 0; lim = kStage;
  };
  int64_t off = 0, poff = 0;
  uint32_t deg = 0, pdeg = 0, trial = 0, lo = 0, hi = 0, y = 0, coin = 0, bkt = 0, pnb = 0;
  uint64_t k = 0;
  int32_t x = 0;
  unsigned long long n_prop = 0, n_mem = 0, n_log = 0;
  const uint64_t t_lo = a.t_common < a.t_far ? a.t_common : a.t_far;
  const uint64_t t_hi = a.t_common < a.t_far ? a.t_far : a.t_common;
  int state = ST_DONE;
  if (live) { sbuf[tid] = curr; staged = 1; len = 1; state = ST_EXTENT; }

  while (__any_sync(0xffffffffu, state != ST_DONE)) {
    // ---- A: draw ----
    if (state == ST_WAIT) {
      const Philox4 r = walker_rng(a.seed_lo, a.seed_hi, walker, (uint32_t)(len - 1), trial);
      trial++;
      k = __umul64hi(((uint64_t)r.x << 32) | (uint64_t)r.w, (uint64_t)deg);
      coin = r.y;
      y = r.z;
      state = ST_PROPOSE;
    }
    __syncwarp();
    // ---- B: one memory access per lane ----
    int4 q0 = make_int4(0, 0, 0, 0), q1 = make_int4(0, 0, 0, 0);
    int32_t v = 0;
    if (state == ST_EXTENT || state == ST_HASH) {
      const int4 *P = state == ST_EXTENT ? reinterpret_cast<const int4 *>(meta + curr)
                                         : reinterpret_cast<const int4 *>(hash + (srw_hash_first(poff) + (int64_t)bkt) * 8);
      gather32<VAR>(P, q0, q1);
    } else if (HAS_ALIAS && state == ST_PROPOSE) {
      q0 = gather16<VAR>(reinterpret_cast<const int4 *>(a.slot + off + (int64_t)k));
    } else if (state == ST_PROPOSE || state == ST_SEARCH) {
      v = __ldg(state == ST_PROPOSE ? a.col + off + (int64_t)k : a.col + poff + (int64_t)((lo + hi) >> 1));
    }
    __syncwarp();
    // ---- C: consume it ----
    int verdict = 0;           // 1 = accept x, 2 = reject (next trial)
    if (state == ST_EXTENT) {
      off = ((int64_t)(uint32_t)q0.x) | ((int64_t)q0.y << 32);
      deg = (uint32_t)q1.x;
      trial = 0;
      state = deg == 0 ? ST_DONE : ST_WAIT;                    // dead end (RW:59-62, RW:115-119)
    } else if (state == ST_PROPOSE) {
      if (HAS_ALIAS) x = (coin < (uint32_t)q0.x) ? q0.y : q0.z; else x = v;
      if (STATS && len > 1) n_prop++;
      if (len == 1 || deg == 1) verdict = 1;                   // first-order step (RW:57) / single choice
      else if (x == prev) verdict = ((uint64_t)y < a.t_ret) ? 1 : 2;    // RS:36  w/p
      else if ((uint64_t)y < t_lo) verdict = 1;
      else if ((uint64_t)y >= t_hi) verdict = 2;
      else {
        if (STATS) { n_mem++; n_log += ceil_log2_p1(pdeg); }
        pnb = srw_hash_buckets(poff, pdeg);
        if (pnb) { bkt = __umulhi(srw_hash32((uint32_t)x), pnb); state = ST_HASH; }
        else { lo = 0; hi = pdeg; state = ST_SEARCH; }
      }
    } else if (state == ST_HASH) {
      const bool found = q0.x == x || q0.y == x || q0.z == x || q0.w == x || q1.x == x || q1.y == x || q1.z == x || q1.w == x;
      if (found) verdict = ((uint64_t)y < a.t_common) ? 1 : 2;           // RS:38  x in N(prev): w
      else if (q1.w == -1) verdict = ((uint64_t)y < a.t_far) ? 1 : 2;    // bucket not full: x is absent (RS:34 w/q)
      else bkt = bkt + 1 == pnb ? 0 : bkt + 1;                           // full bucket: linear probing
    } else if (state == ST_SEARCH) {
      const uint32_t mid = (lo + hi) >> 1;
      if (v == x) verdict = ((uint64_t)y < a.t_common) ? 1 : 2;
      else {
        if (v < x) lo = mid + 1; else hi = mid;
        if (lo >= hi) verdict = ((uint64_t)y < a.t_far) ? 1 : 2;
      }
    }
    if (verdict == 1) {                                        // RW:114, then RW:103
      sbuf[staged * 256 + tid] = x;
      staged++; len++;
      if (staged == lim) flush();
      prev = curr; poff = off; pdeg = deg;
      curr = x;
      state = (len == a.stride) ? ST_DONE : ST_EXTENT;
    } else if (verdict == 2) {
      state = ST_WAIT;
    }
  }
  if (live) {
    flush();
    a.lens[i] = len;
  }
  if (STATS) {
    atomicAdd(a.stats + 1, n_prop);
    atomicAdd(a.stats + 2, n_mem);
    atomicAdd(a.stats + 3, n_log);
  }
}

// ------------------------------------------------------------------------------------------
// K6 (v5, WEIGHTED alias-fold: undirected weighted graphs with 1/p > max(1, 1/q)).  The folding of
// walk_fold_conv_kernel with (multiplicity, degree) replaced by (bundle weight wb, row weight sum W):
// a trial returns to prev with probability a*wb / (Mp*W + a*wb), else proposes from the Vose table under the
// envelope Mp.  r.y is the Vose coin, so ONE draw z = r.z decides both: with t1 = a*wb, t2 = Mp*W,
//     z*(t2 + t1) < t1*2^32                 -> return (no memory access; the walker already holds prev's row)
//     z*(t2 + t1) < t1*2^32 + t2*T(class)   -> the proposal is accepted (z rescaled to the rest of [0, 1))
// all in IEEE double, one rounding per operation (the CPU twin evaluates the same expressions).  One 32-byte
// AliasSlotW gather yields the neighbour AND its bundle weight whichever way the coin falls; W rides in the
// row descriptor.  Requests per step: 1 descriptor + ~1.9 slots + ~0.9 buckets instead of 1 + 3.7 + 1.8.
// ------------------------------------------------------------------------------------------
template <bool STATS, int VAR>
__global__ void __launch_bounds__(256, 4) walk_wfold_conv_kernel(WalkArgs a, FoldArgs f, const RowMeta *__restrict__ meta,
                                                                 const int32_t *__restrict__ hash, const AliasSlotW *__restrict__ slotw) {
  __shared__ int32_t sbuf[kStage * 256];
  const int tid = threadIdx.x;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + tid;
  const bool live = i < a.n_walkers;
  const uint64_t walker = a.walker_first + (uint64_t)(live ? i : 0);
  int32_t curr = (int32_t)(walker % (uint64_t)a.nv), prev = -1;
  int32_t *path = a.paths + (live ? i : 0) * a.stride;
  int32_t len = 0, staged = 0, flushed = 0;
  const int head = (int)(((16u - (uint32_t)(reinterpret_cast<uintptr_t>(path) & 15u)) & 15u) >> 2);
  int lim = head ? kStage - 4 + head : kStage;
  auto flush = [&]() {
    int32_t *dst = path + flushed;
    if (staged == kStage && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < kStage; j += 4)
        *reinterpret_cast<int4 *>(dst + j) = make_int4(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid], sbuf[(j + 2) * 256 + tid], sbuf[(j + 3) * 256 + tid]);
    } else {
      int j = 0;
      if ((reinterpret_cast<uintptr_t>(dst) & 4) && staged >= 1) { dst[0] = sbuf[tid]; j = 1; }
      if ((reinterpret_cast<uintptr_t>(dst + j) & 8) && j + 1 < staged) {
        *reinterpret_cast<int2 *>(dst + j) = make_int2(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid]);
        j += 2;
      }
#pragma unroll 1
      for (; j + 3 < staged; j += 4)
        *reinterpret_cast<int4 *>(dst + j) = make_int4(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid], sbuf[(j + 2) * 256 + tid], sbuf[(j + 3) * 256 + tid]);
#pragma unroll 1
      for (; j + 1 < staged; j += 2) *reinterpret_cast<int2 *>(dst + j) = make_int2(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid]);
#pragma unroll 1
      for (; j < staged; ++j) dst[j] = sbuf[j * 256 + tid];
    }
    flushed += staged; staged = 0; lim = kStage;
  };
  uint32_t off = 0, poff = 0, k = 0;            // row offsets fit 32 bits: the ABI caps nnz below 2^32
  uint32_t deg = 0, pdeg = 0, trial = 0, lo = 0, hi = 0, coin = 0, bkt = 0, pnb = 0;
  int32_t x = 0;
  double W = 0.0, pW = 0.0, wb = 0.0, xwb = 0.0, zl = 0.0;
  unsigned long long n_prop = 0, n_mem = 0, n_log = 0;
  int state = ST_DONE;
  if (live) { sbuf[tid] = curr; staged = 1; len = 1; state = ST_EXTENT; }

  while (__any_sync(0xffffffffu, state != ST_DONE)) {
    int32_t newv = 0;
    int moved = 0;             // 1 = accepted proposal (the new row's descriptor is needed), 2 = direct return
    // ---- A: draw ----
    if (state == ST_WAIT) {
      const Philox4 r = walker_rng(a.seed_lo, a.seed_hi, walker, (uint32_t)(len - 1), trial);
      trial++;
      bool ret = false;
      if (len > 1) {
        const double t1 = __dmul_rn(f.a, wb), t2 = __dmul_rn(f.mp, W);
        zl = __dmul_rn((double)r.z, __dadd_rn(t2, t1));
        ret = zl < __dmul_rn(t1, 4294967296.0);
      }
      if (ret) {                                               // return-excess component: always accepted, no memory access
        if (STATS) n_prop++;
        newv = prev; moved = 2;
        const int32_t c = curr; curr = prev; prev = c;
        const uint32_t o = off; off = poff; poff = o;
        const uint32_t d = deg; deg = pdeg; pdeg = d;
        const double w = W; W = pW; pW = w;                    // wb unchanged: the same bundle of parallel edges
      } else {
        k = (uint32_t)__umul64hi(((uint64_t)r.x << 32) | (uint64_t)r.w, (uint64_t)deg);
        coin = r.y;
        state = ST_PROPOSE;
      }
    }
    __syncwarp();
    // ---- B: one memory access per lane ----
    int4 q0 = make_int4(0, 0, 0, 0), q1 = make_int4(0, 0, 0, 0);
    int32_t v = 0;
    if (!moved) {
      if (state == ST_EXTENT || state == ST_HASH || state == ST_PROPOSE) {
        const int4 *P = state == ST_EXTENT ? reinterpret_cast<const int4 *>(meta + curr)
                      : state == ST_HASH   ? reinterpret_cast<const int4 *>(hash + ((uint64_t)(poff >> 2) + bkt) * 8)
                                           : reinterpret_cast<const int4 *>(slotw + ((uint64_t)off + k));
        gather32<VAR>(P, q0, q1);
      } else if (state == ST_SEARCH) {
        v = __ldg(a.col + ((uint64_t)poff + ((lo + hi) >> 1)));
      }
    }
    __syncwarp();
    // ---- C: consume it ----
    int verdict = 0;           // 1 = accept x, 2 = reject (next trial)
    if (!moved) {
      if (state == ST_EXTENT) {
        off = (uint32_t)q0.x;
        deg = (uint32_t)q1.x;
        W = __hiloint2double(q1.w, q1.z);
        trial = 0;
        state = deg == 0 ? ST_DONE : ST_WAIT;                  // dead end (RW:59-62); cannot happen after the first vertex (undirected)
      } else if (state == ST_PROPOSE) {
        const bool own = coin < (uint32_t)q0.x;
        x = own ? q0.y : q0.z;
        xwb = own ? __hiloint2double(q1.y, q1.x) : __hiloint2double(q1.w, q1.z);
        if (STATS && len > 1) n_prop++;
        if (len == 1 || deg == 1 || x == prev) verdict = 1;    // first-order step (RW:57) / single choice / envelope mass Mp of Mp
        else {
          const double t1 = __dmul_rn(f.a, wb), t2 = __dmul_rn(f.mp, W), r0 = __dmul_rn(t1, 4294967296.0);
          const double rc = __dadd_rn(r0, __dmul_rn(t2, (double)f.t_common)), rf = __dadd_rn(r0, __dmul_rn(t2, (double)f.t_far));
          const double rlo = rc < rf ? rc : rf, rhi = rc < rf ? rf : rc;
          if (zl < rlo) verdict = 1;
          else if (!(zl < rhi)) verdict = 2;
          else {
            if (STATS) { n_mem++; n_log += ceil_log2_p1(pdeg); }
            pnb = srw_hash_buckets((int64_t)poff, pdeg);
            if (pnb) { bkt = __umulhi(srw_hash32((uint32_t)x), pnb); state = ST_HASH; }
            else { lo = 0; hi = pdeg; state = ST_SEARCH; }
          }
        }
      } else if (state == ST_HASH || state == ST_SEARCH) {
        int member = -1;       // 1 = x in N(prev), 0 = absent, -1 = undecided (next probe)
        if (state == ST_HASH) {
          const bool found = q0.x == x || q0.y == x || q0.z == x || q0.w == x || q1.x == x || q1.y == x || q1.z == x || q1.w == x;
          if (found) member = 1;
          else if (q1.w == -1) member = 0;                     // bucket not full: x is absent
          else bkt = bkt + 1 == pnb ? 0 : bkt + 1;             // full bucket: linear probing
        } else {
          const uint32_t mid = (lo + hi) >> 1;
          if (v == x) member = 1;
          else {
            if (v < x) lo = mid + 1; else hi = mid;
            if (lo >= hi) member = 0;
          }
        }
        if (member >= 0) {
          const double t1 = __dmul_rn(f.a, wb), t2 = __dmul_rn(f.mp, W), r0 = __dmul_rn(t1, 4294967296.0);
          const double rr = __dadd_rn(r0, __dmul_rn(t2, (double)(member ? f.t_common : f.t_far)));   // RS:38 / RS:34
          verdict = zl < rr ? 1 : 2;
        }
      }
    }
    if (verdict == 1) {                                        // move along the proposal
      newv = x; moved = 1;
      prev = curr; poff = off; pdeg = deg; pW = W;
      curr = x; wb = xwb;
    } else if (verdict == 2) {
      state = ST_WAIT;
    }
    if (moved) {                                               // RW:114, then RW:103
      sbuf[staged * 256 + tid] = newv;
      staged++; len++;
      if (staged == lim) flush();
      trial = 0;
      state = (len == a.stride) ? ST_DONE : (moved == 1 ? ST_EXTENT : ST_WAIT);
    }
  }
  if (live) {
    flush();
    a.lens[i] = len;
  }
  if (STATS) {
    atomicAdd(a.stats + 1, n_prop);
    atomicAdd(a.stats + 2, n_mem);
    atomicAdd(a.stats + 3, n_log);
  }
}
