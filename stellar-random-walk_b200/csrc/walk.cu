// walk.cu -- K4/K5/K6: first step + second-order walk steps on the device.
//
// Replaces RandomWalk.initFirstStep (RW:51-66), the per-walker hot loop of RandomWalk.randomWalk
// (RW:95-139) and RandomSample (RS:5-63).  Two samplers:
//
//  * EXACT  (walk_exact_kernel): RS:12-62 literally -- float32 bias weights (w/p, w/q, RS:34-38),
//    float64 left-to-right sum and inverse-CDF scan with `acc >= u` and the edges.head fallback
//    (RS:14-24) over the file-appearance-order row.  The O(d_c*d_p) `exists` scan (RS:38) is replaced
//    by a binary search in the sorted row of prev (same truth value).  Bit-identical to the oracle.
//  * ALIAS  (walk_alias_kernel): one alias-table proposal per trial from the static weights, accepted
//    with probability f(x)/M where f is the RS:33-41 bias factor; distribution-equal to EXACT,
//    bit-identical to the CPU twin in oracle/ (oracle_alias_walk).
//
// One walker per thread; a draw is Philox4x32-10 keyed by (seed; walker, step, trial).
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <vector>

#include "philox.cuh"
#include "srw_internal.h"

namespace {

#include "walk_conv.cuh"   // WalkArgs, FoldArgs, PeerTable, the state enum and the convergent kernels (v5)
#include "walk_exact.cuh"  // row_contains / hash_contains and K5, the exact (bit-parity) sampler kernels



// alias proposal from row [off, off+deg): slot index from 64 random bits, Vose coin from r.y
template <bool HAS_ALIAS>
__device__ __forceinline__ int32_t propose(const WalkArgs &a, int64_t off, int64_t deg, const Philox4 &r) {
  const uint64_t R = ((uint64_t)r.x << 32) | (uint64_t)r.w;
  const int64_t k = (int64_t)__umul64hi(R, (uint64_t)deg);
  if (HAS_ALIAS) {
    const int4 raw = __ldg(reinterpret_cast<const int4 *>(a.slot + off + k));
    return (r.y < (uint32_t)raw.x) ? raw.y : raw.z;
  } else {
    return __ldg(a.col + off + k);
  }
}

// ------------------------------------------------------------------------------------------
// K6: alias sampler
// ------------------------------------------------------------------------------------------
template <bool HAS_ALIAS, bool STATS>
__global__ void __launch_bounds__(256) walk_alias_kernel(WalkArgs a) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= a.n_walkers) return;
  const uint64_t walker = a.walker_first + (uint64_t)i;
  int32_t curr = (int32_t)(walker % (uint64_t)a.nv);
  int32_t *path = a.paths + i * a.stride;
  path[0] = curr;
  int32_t len = 1;
  int64_t off = __ldg(a.off + curr), deg = __ldg(a.off + curr + 1) - off;
  unsigned long long n_prop = 0, n_mem = 0, n_log = 0;
  if (deg > 0) {
    // RW:51-66 first step: first-order draw, the proposal is the sample
    Philox4 r = walker_rng(a.seed_lo, a.seed_hi, walker, 0u, 0u);
    int32_t prev = curr;
    int64_t poff = off, pdeg = deg;
    curr = propose<HAS_ALIAS>(a, off, deg, r);
    path[len++] = curr;
    const uint64_t t_lo = a.t_common < a.t_far ? a.t_common : a.t_far;
    const uint64_t t_hi = a.t_common < a.t_far ? a.t_far : a.t_common;
    while (len != a.stride) {                                  // RW:103
      off = __ldg(a.off + curr);
      deg = __ldg(a.off + curr + 1) - off;
      if (deg <= 0) break;                                     // RW:115-119 dead end
      int32_t x;
      if (deg == 1) {
        x = __ldg(a.col + off);                                // single choice: any trial count accepts it
        if (STATS) n_prop++;
      } else {
        for (uint32_t trial = 0;; ++trial) {
          r = walker_rng(a.seed_lo, a.seed_hi, walker, (uint32_t)(len - 1), trial);
          x = propose<HAS_ALIAS>(a, off, deg, r);
          if (STATS) n_prop++;
          const uint64_t y = r.z;
          uint64_t t;
          if (x == prev) t = a.t_ret;                          // RS:36  w/p
          else if (y < t_lo) break;                            // below both bounds: accept without a test
          else if (y >= t_hi) continue;                        // above both bounds: reject without a test
          else {
            if (STATS) { n_mem++; n_log += ceil_log2_p1(pdeg); }
            t = row_contains(a.col, poff, pdeg, x) ? a.t_common : a.t_far;   // RS:38 w  |  RS:34 w/q
          }
          if (y < t) break;
        }
      }
      prev = curr; poff = off; pdeg = deg;
      curr = x;
      path[len++] = x;                                         // RW:114
    }
  }
  a.lens[i] = len;
  if (STATS) {
    atomicAdd(a.stats + 1, n_prop);
    atomicAdd(a.stats + 2, n_mem);
    atomicAdd(a.stats + 3, n_log);
  }
}

// ------------------------------------------------------------------------------------------
// K6 (v2): the same sampler as a per-lane state machine.  The v1 loop nest above leaves ~5 of 32
// lanes active (ncu: smsp__thread_inst_executed_per_inst_executed = 4.8) because rejection loops and
// binary searches of different lengths serialise inside a warp.  Here every lane performs exactly ONE
// dependent memory access per iteration of a single convergent loop -- a row-extent load, a proposal
// gather or a binary-search probe, whichever its walker needs next -- so a warp keeps 32 independent
// gathers in flight.  Decisions are the same pure functions of (seed; walker, step, trial): the
// output is bit-identical to v1 and to the CPU twin.
// ------------------------------------------------------------------------------------------

template <bool HAS_ALIAS, bool STATS>
__global__ void __launch_bounds__(256) walk_alias_sm_kernel(WalkArgs a) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= a.n_walkers) return;
  const uint64_t walker = a.walker_first + (uint64_t)i;
  int32_t curr = (int32_t)(walker % (uint64_t)a.nv), prev = -1;
  int32_t *path = a.paths + i * a.stride;
  path[0] = curr;
  int32_t len = 1;
  int64_t off = 0, poff = 0;
  uint32_t deg = 0, pdeg = 0, trial = 0, lo = 0, hi = 0, y = 0;
  int32_t x = 0;
  uint64_t k = 0;              // proposal slot of the pending trial
  uint32_t coin = 0;           // Vose coin of the pending trial
  int state = ST_EXTENT;
  unsigned long long n_prop = 0, n_mem = 0, n_log = 0;
  const uint64_t t_lo = a.t_common < a.t_far ? a.t_common : a.t_far;
  const uint64_t t_hi = a.t_common < a.t_far ? a.t_far : a.t_common;

  while (state != ST_DONE) {
    // ---- one memory access per lane ----
    int64_t e0 = 0, e1 = 0;
    int32_t v = 0, v_alias = 0;
    uint32_t thr = 0xFFFFFFFFu;
    if (state == ST_EXTENT) {
      e0 = __ldg(a.off + curr);
      e1 = __ldg(a.off + curr + 1);
    } else if (state == ST_PROPOSE) {
      if (HAS_ALIAS) {
        const int4 raw = __ldg(reinterpret_cast<const int4 *>(a.slot + off + (int64_t)k));
        thr = (uint32_t)raw.x; v = raw.y; v_alias = raw.z;
      } else {
        v = __ldg(a.col + off + (int64_t)k);
      }
    } else {
      v = __ldg(a.col + poff + (int64_t)((lo + hi) >> 1));
    }
    // ---- consume it ----
    int verdict = 0;           // 0 = nothing yet, 1 = accept x, 2 = reject (next trial)
    if (state == ST_EXTENT) {
      off = e0;
      deg = (uint32_t)(e1 - e0);
      if (deg == 0) { state = ST_DONE; continue; }              // RW:59-62 / RW:115-119 dead end
      trial = 0;
      verdict = 2;                                             // draw trial 0
    } else if (state == ST_PROPOSE) {
      x = (HAS_ALIAS && !(coin < thr)) ? v_alias : v;
      if (STATS && len > 1) n_prop++;
      if (len == 1 || deg == 1) verdict = 1;                   // first-order step (RW:57) / single choice
      else if (x == prev) verdict = ((uint64_t)y < a.t_ret) ? 1 : 2;    // RS:36  w/p
      else if ((uint64_t)y < t_lo) verdict = 1;                // below both bounds: accept without a test
      else if ((uint64_t)y >= t_hi) verdict = 2;               // above both bounds: reject without a test
      else {
        if (STATS) { n_mem++; n_log += ceil_log2_p1(pdeg); }
        lo = 0; hi = pdeg;
        state = ST_SEARCH;
      }
    } else {
      const uint32_t mid = (lo + hi) >> 1;
      if (v == x) verdict = ((uint64_t)y < a.t_common) ? 1 : 2;          // RS:38  x in N(prev): w
      else {
        if (v < x) lo = mid + 1; else hi = mid;
        if (lo >= hi) verdict = ((uint64_t)y < a.t_far) ? 1 : 2;         // RS:34  not a neighbour: w/q
      }
    }
    if (verdict == 1) {
      path[len++] = x;                                         // RW:114
      prev = curr; poff = off; pdeg = deg;
      curr = x;
      state = (len == a.stride) ? ST_DONE : ST_EXTENT;         // RW:103
    } else if (verdict == 2) {
      const Philox4 r = walker_rng(a.seed_lo, a.seed_hi, walker, (uint32_t)(len - 1), trial);
      trial++;
      k = __umul64hi(((uint64_t)r.x << 32) | (uint64_t)r.w, (uint64_t)deg);
      coin = r.y;
      y = r.z;
      state = ST_PROPOSE;
    }
  }
  a.lens[i] = len;
  if (STATS) {
    atomicAdd(a.stats + 1, n_prop);
    atomicAdd(a.stats + 2, n_mem);
    atomicAdd(a.stats + 3, n_log);
  }
}

// ------------------------------------------------------------------------------------------
// K6 (v3): v2's state machine over packed row descriptors and per-row neighbour hash sets.
// ncu on v2 (RMAT-24): 865 B of DRAM traffic per step, about half of it the ~10-probe binary search
// for "is x a neighbour of prev" (RS:38).  Here that test is one 32-byte bucket probe (rows longer
// than kHashMinDeg), and the row extent is one aligned 32-byte RowMeta load.  Same decisions, same bits.
// ------------------------------------------------------------------------------------------

template <bool HAS_ALIAS, bool STATS>
__global__ void __launch_bounds__(256) walk_alias_hash_kernel(WalkArgs a, const RowMeta *__restrict__ meta,
                                                              const int32_t *__restrict__ hash) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= a.n_walkers) return;
  const uint64_t walker = a.walker_first + (uint64_t)i;
  int32_t curr = (int32_t)(walker % (uint64_t)a.nv), prev = -1;
  int32_t *path = a.paths + i * a.stride;
  path[0] = curr;
  int32_t len = 1;
  int64_t off = 0, hoff = 0, poff = 0, phoff = 0;
  uint32_t deg = 0, nb = 0, pdeg = 0, pnb = 0, trial = 0, lo = 0, hi = 0, y = 0, coin = 0, bkt = 0;
  int32_t x = 0;
  uint64_t k = 0;
  int state = ST_EXTENT;
  unsigned long long n_prop = 0, n_mem = 0, n_log = 0;
  const uint64_t t_lo = a.t_common < a.t_far ? a.t_common : a.t_far;
  const uint64_t t_hi = a.t_common < a.t_far ? a.t_far : a.t_common;

  while (state != ST_DONE) {
    // ---- one memory access per lane ----
    int4 q0 = make_int4(0, 0, 0, 0), q1 = make_int4(0, 0, 0, 0);
    int32_t v = 0;
    if (state == ST_EXTENT) {
      const int4 *m = reinterpret_cast<const int4 *>(meta + curr);
      q0 = __ldg(m); q1 = __ldg(m + 1);
    } else if (state == ST_PROPOSE) {
      if (HAS_ALIAS) q0 = __ldg(reinterpret_cast<const int4 *>(a.slot + off + (int64_t)k));
      else v = __ldg(a.col + off + (int64_t)k);
    } else if (state == ST_HASH) {
      const int4 *b = reinterpret_cast<const int4 *>(hash + (phoff + (int64_t)bkt) * 8);
      q0 = __ldg(b); q1 = __ldg(b + 1);
    } else {
      v = __ldg(a.col + poff + (int64_t)((lo + hi) >> 1));
    }
    // ---- consume it ----
    int verdict = 0;           // 1 = accept x, 2 = reject (next trial)
    if (state == ST_EXTENT) {
      off = ((int64_t)(uint32_t)q0.x) | ((int64_t)q0.y << 32);
      hoff = ((int64_t)(uint32_t)q0.z) | ((int64_t)q0.w << 32);
      deg = (uint32_t)q1.x; nb = (uint32_t)q1.y;
      if (deg == 0) { state = ST_DONE; continue; }              // dead end (RW:59-62, RW:115-119)
      trial = 0;
      verdict = 2;
    } else if (state == ST_PROPOSE) {
      if (HAS_ALIAS) x = (coin < (uint32_t)q0.x) ? q0.y : q0.z; else x = v;
      if (STATS && len > 1) n_prop++;
      if (len == 1 || deg == 1) verdict = 1;                   // first-order step (RW:57) / single choice
      else if (x == prev) verdict = ((uint64_t)y < a.t_ret) ? 1 : 2;    // RS:36  w/p
      else if ((uint64_t)y < t_lo) verdict = 1;
      else if ((uint64_t)y >= t_hi) verdict = 2;
      else {
        if (STATS) { n_mem++; n_log += ceil_log2_p1(pdeg); }
        if (pnb) { bkt = __umulhi(srw_hash32((uint32_t)x), pnb); state = ST_HASH; }
        else { lo = 0; hi = pdeg; state = ST_SEARCH; }
      }
    } else if (state == ST_HASH) {
      const bool found = q0.x == x || q0.y == x || q0.z == x || q0.w == x || q1.x == x || q1.y == x || q1.z == x || q1.w == x;
      if (found) verdict = ((uint64_t)y < a.t_common) ? 1 : 2;           // RS:38  x in N(prev): w
      else if (q1.w == -1) verdict = ((uint64_t)y < a.t_far) ? 1 : 2;    // bucket not full: x is absent (RS:34 w/q)
      else bkt = bkt + 1 == pnb ? 0 : bkt + 1;                           // full bucket: linear probing
    } else {
      const uint32_t mid = (lo + hi) >> 1;
      if (v == x) verdict = ((uint64_t)y < a.t_common) ? 1 : 2;
      else {
        if (v < x) lo = mid + 1; else hi = mid;
        if (lo >= hi) verdict = ((uint64_t)y < a.t_far) ? 1 : 2;
      }
    }
    if (verdict == 1) {
      path[len++] = x;                                         // RW:114
      prev = curr; poff = off; pdeg = deg; phoff = hoff; pnb = nb;
      curr = x;
      state = (len == a.stride) ? ST_DONE : ST_EXTENT;         // RW:103
    } else if (verdict == 2) {
      const Philox4 r = walker_rng(a.seed_lo, a.seed_hi, walker, (uint32_t)(len - 1), trial);
      trial++;
      k = __umul64hi(((uint64_t)r.x << 32) | (uint64_t)r.w, (uint64_t)deg);
      coin = r.y;
      y = r.z;
      state = ST_PROPOSE;
    }
  }
  a.lens[i] = len;
  if (STATS) {
    atomicAdd(a.stats + 1, n_prop);
    atomicAdd(a.stats + 2, n_mem);
    atomicAdd(a.stats + 3, n_log);
  }
}

// ------------------------------------------------------------------------------------------
// K6 (v4, SRW_SAMPLER_ALIAS_FOLD): fewer memory requests per step.  ncu on v3 (RMAT-26): the kernel runs
// at the memory system's random-request ceiling (~46 G requests/s, profiles/README.md) with 6.1 requests
// per step: 3.7 proposals, 1 row descriptor, ~1 hash probe, 1 path write.  v4 removes most of them:
//   * fold: for 1/p > max(1, 1/q) the return edge's excess weight (1/p - Mp) * mult is its own mixture
//     component, picked with probability a*m / (Mp*deg + a*m) and always accepted; everything else is
//     rejection under the envelope Mp = max(1, 1/q) instead of 1/p  (3.7 -> ~1.8 proposals per step);
//   * the 16-byte neighbour entry carries deg/off/multiplicity of the neighbour: no row-descriptor load;
//   * the hash set of prev is addressed from (poff, pdeg) alone;
//   * path ids are staged in shared memory and flushed as 8-byte stores, 16 ids at a time.
// Defined for undirected, unweighted graphs (multiplicity of prev in N(curr) == multiplicity of the edge
// just taken); otherwise the launch falls back to v3.  CPU twin: oracle_alias_walk with cfg.fold = 1.
// ------------------------------------------------------------------------------------------
template <bool STATS, bool PEER, int MINB = 4>
__global__ void __launch_bounds__(256, MINB) walk_fold_kernel(WalkArgs a, FoldArgs f, const PeerTable pt) {
  __shared__ int32_t sbuf[kStage * 256];
  __shared__ const NbrEntry *s_ent[SRW_MAX_SHARDS];
  __shared__ const int32_t *s_hash[SRW_MAX_SHARDS];
  const int tid = threadIdx.x;
  if (PEER) {
    if (tid < SRW_MAX_SHARDS) { s_ent[tid] = pt.ent[tid]; s_hash[tid] = pt.hash[tid]; }
    __syncthreads();
  }
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + tid;
  if (i >= a.n_walkers) return;
  const uint64_t walker = a.walker_first + (uint64_t)i;
  int32_t curr = (int32_t)(walker % (uint64_t)a.nv), prev = -1;
  int32_t *path = a.paths + i * a.stride;
  const bool vec2 = ((a.stride & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.paths) & 7) == 0);
  int32_t len = 0, staged = 0, flushed = 0;
  auto flush = [&]() {
    int32_t *dst = path + flushed;
    int j = 0;
    if (vec2) for (; j + 1 < staged; j += 2) *reinterpret_cast<int2 *>(dst + j) = make_int2(sbuf[j * 256 + tid], sbuf[(j + 1) * 256 + tid]);
    for (; j < staged; ++j) dst[j] = sbuf[j * 256 + tid];
    flushed += staged; staged = 0;
  };
  auto push = [&](int32_t v) {
    sbuf[staged * 256 + tid] = v;
    staged++; len++;
    if (staged == kStage) flush();
  };
  push(curr);
  int64_t off = 0, poff = 0, xoff = 0;
  uint32_t deg = 0, pdeg = 0, m = 1, xdeg = 0, xm = 1, trial = 0, lo = 0, hi = 0, y = 0, bkt = 0, pnb = 0;
  uint32_t cown = 0, pown = 0, xown = 0;   // PEER: shards that hold the rows of curr / prev / x
  int32_t x = 0;
  uint64_t k = 0;
  double ret_lhs = 0.0, ret_rhs = 0.0;
  int state = ST_EXTENT;      // only for the start vertex
  unsigned long long n_prop = 0, n_mem = 0, n_log = 0;
  const uint64_t t_lo = f.t_common < f.t_far ? f.t_common : f.t_far;
  const uint64_t t_hi = f.t_common < f.t_far ? f.t_far : f.t_common;

  while (state != ST_DONE) {
    // ---- one memory access per lane ----
    int4 q0 = make_int4(0, 0, 0, 0), q1 = make_int4(0, 0, 0, 0);
    int64_t e0 = 0, e1 = 0;
    int32_t v = 0;
    if (state == ST_EXTENT) {
      if (PEER) {
        while ((int)cown + 1 < pt.world && (int64_t)curr >= pt.first[cown + 1]) cown++;
        const int64_t *o = pt.off[cown] + ((int64_t)curr - pt.first[cown]);
        e0 = __ldg(o); e1 = __ldg(o + 1);
      } else {
        e0 = __ldg(a.off + curr); e1 = __ldg(a.off + curr + 1);
      }
    } else if (state == ST_PROPOSE) {
      q0 = __ldg(reinterpret_cast<const int4 *>((PEER ? s_ent[cown] : f.ent) + off + (int64_t)k));
    } else if (state == ST_HASH) {
      const int4 *b = reinterpret_cast<const int4 *>((PEER ? s_hash[pown] : f.hash) + (srw_hash_first(poff) + (int64_t)bkt) * 8);
      q0 = __ldg(b); q1 = __ldg(b + 1);
    } else {
      v = __ldg(&(PEER ? s_ent[pown] : f.ent)[poff + (int64_t)((lo + hi) >> 1)].x);
    }
    // ---- consume it ----
    int verdict = 0;           // 1 = accept entry x, 2 = reject (next trial), 3 = new step: draw trial 0, 4 = direct return
    if (state == ST_EXTENT) {
      off = e0; deg = (uint32_t)(e1 - e0);
      if (deg == 0) { state = ST_DONE; continue; }              // dead end (RW:59-62)
      verdict = 3;
    } else if (state == ST_PROPOSE) {
      x = q0.x; xdeg = (uint32_t)q0.y;
      xoff = (int64_t)(uint32_t)q0.z;
      xown = (uint32_t)q0.w & 0xFFu;
      xm = (uint32_t)q0.w >> 8;
      if (STATS && len > 1) n_prop++;
      if (len == 1 || deg == 1) verdict = 1;                   // first-order step (RW:57) / single choice
      else if (x == prev) verdict = ((uint64_t)y < f.t_ret) ? 1 : 2;   // RS:36; folded: mass Mp of Mp, t_ret = 2^32
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
      if (found) verdict = ((uint64_t)y < f.t_common) ? 1 : 2;           // RS:38
      else if (q1.w == -1) verdict = ((uint64_t)y < f.t_far) ? 1 : 2;    // RS:34
      else bkt = bkt + 1 == pnb ? 0 : bkt + 1;
    } else {
      const uint32_t mid = (lo + hi) >> 1;
      if (v == x) verdict = ((uint64_t)y < f.t_common) ? 1 : 2;
      else {
        if (v < x) lo = mid + 1; else hi = mid;
        if (lo >= hi) verdict = ((uint64_t)y < f.t_far) ? 1 : 2;
      }
    }
    bool draw = false;
    if (verdict == 1) {                                        // move along entry (x, xoff, xdeg, xm)
      push(x);                                                 // RW:114
      prev = curr; poff = off; pdeg = deg; pown = cown;
      curr = x; off = xoff; deg = xdeg; m = xm; cown = xown;
      verdict = 3;
    }
    if (verdict == 3) {                                        // a new step starts at curr
      if (len == a.stride || deg == 0) { state = ST_DONE; continue; }   // RW:103 / RW:115-119
      trial = 0;
      draw = true;
    } else if (verdict == 2) {
      trial++;
      draw = true;
    }
    while (draw) {
      if (trial == 0 && len > 1) {                             // per step: P(return-excess component) = a*m / (Mp*deg + a*m)
        const double t1 = __dmul_rn(f.a, (double)m), t2 = __dmul_rn(f.mp, (double)deg);
        ret_lhs = __dadd_rn(t2, t1);
        ret_rhs = __dmul_rn(t1, 4294967296.0);
      }
      const Philox4 r = walker_rng(a.seed_lo, a.seed_hi, walker, (uint32_t)(len - 1), trial);
      if (len > 1 && __dmul_rn((double)r.y, ret_lhs) < ret_rhs) {   // return-excess component: always accepted, no memory access
        if (STATS) n_prop++;
        push(prev);
        const int32_t c = curr; curr = prev; prev = c;
        const int64_t o = off; off = poff; poff = o;
        const uint32_t d = deg; deg = pdeg; pdeg = d;          // m unchanged: the same bundle of parallel edges
        const uint32_t w = cown; cown = pown; pown = w;
        if (len == a.stride) { state = ST_DONE; break; }
        trial = 0;
        continue;                                              // draw trial 0 of the next step
      }
      k = __umul64hi(((uint64_t)r.x << 32) | (uint64_t)r.w, (uint64_t)deg);
      y = r.z;
      state = ST_PROPOSE;
      draw = false;
    }
  }
  flush();
  a.lens[i] = len;
  if (STATS) {
    atomicAdd(a.stats + 1, n_prop);
    atomicAdd(a.stats + 2, n_mem);
    atomicAdd(a.stats + 3, n_log);
  }
}

// ranks -> original vertex ids, and the step count
__global__ void finalize_paths_kernel(int64_t n_walkers, int32_t stride, const int32_t *__restrict__ vids,
                                      const int32_t *__restrict__ lens, int32_t *paths, unsigned long long *stats) {
  unsigned long long steps = 0;
  // one warp per path row: coalesced, no 64-bit division
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t wk = warp; wk < n_walkers; wk += n_warps) {
    const int32_t len = __ldg(lens + wk);
    int32_t *row = paths + wk * stride;
    for (int32_t k = lane; k < stride; k += 32) row[k] = k < len ? __ldg(vids + row[k]) : -1;
    if (lane == 0) steps += (unsigned long long)(len - 1);
  }
  // warp-reduce then one atomic per warp
  for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
  if ((threadIdx.x & 31) == 0 && steps) atomicAdd(stats, steps);
}

// The same pass over the matrix as ONE contiguous array of int2 (even stride: rows are 8-byte aligned and adjacent),
// four independent 8-byte loads and eight id gathers in flight per thread instead of one dependent chain per lane.
// row = e2 / half by a 64-bit multiply-high (magic = ceil(2^64 / half), exact for e2 * half < 2^64).
__global__ void __launch_bounds__(256) finalize_paths_flat_kernel(int64_t n_pairs, uint32_t half, uint64_t magic, const int32_t *__restrict__ vids,
                                                                  const int32_t *__restrict__ lens, int2 *paths, unsigned long long *stats) {
  constexpr int U = 4;
  unsigned long long steps = 0;
  const int64_t base = (int64_t)blockIdx.x * (256 * U) + threadIdx.x;
  int2 v[U];
  int32_t len[U];
  uint32_t c[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t e = base + (int64_t)u * 256;
    v[u] = make_int2(0, 0); len[u] = 0; c[u] = 0;
    if (e < n_pairs) {
      v[u] = paths[e];
      const uint64_t r = __umul64hi((uint64_t)e, magic);
      c[u] = (uint32_t)((uint64_t)e - r * half) * 2u;
      len[u] = __ldg(lens + r);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t e = base + (int64_t)u * 256;
    if (e < n_pairs) {
      int2 o;
      o.x = (int32_t)c[u] < len[u] ? __ldg(vids + v[u].x) : -1;
      o.y = (int32_t)c[u] + 1 < len[u] ? __ldg(vids + v[u].y) : -1;
      paths[e] = o;
      if (c[u] == 0) steps += (unsigned long long)(len[u] - 1);
    }
  }
  for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
  if ((threadIdx.x & 31) == 0 && steps) atomicAdd(stats, steps);
}

// id-space walks need no translation: count the steps and pad the rows of walkers that stopped early (dead ends) with -1
__global__ void count_steps_kernel(int64_t n_walkers, int32_t stride, const int32_t *__restrict__ lens, int32_t *paths, unsigned long long *stats) {
  unsigned long long steps = 0;
  for (int64_t wk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; wk < n_walkers; wk += (int64_t)gridDim.x * blockDim.x) {
    const int32_t len = __ldg(lens + wk);
    steps += (unsigned long long)(len - 1);
    for (int32_t k = len; k < stride; ++k) paths[wk * stride + k] = -1;
  }
  for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
  if ((threadIdx.x & 31) == 0 && steps) atomicAdd(stats, steps);
}

// ---- KAT kernels: one thread, same device functions as the exact walk ----
__global__ void kat_sample_kernel(int64_t n, const float *w, float u, int64_t *out) {
  *out = cdf_pick(n, u, [&](int64_t j) { return w[j]; });
}
__global__ void kat_second_order_kernel(float p, float q, int32_t prev, int64_t np, const int32_t *pd_sorted, int64_t nc,
                                        const int32_t *cd, const float *cw, float u, float *w_out, int64_t *k_out) {
  for (int64_t j = 0; j < nc; ++j)
    w_out[j] = biased_weight(p, q, prev, cd[j], cw[j], row_contains(pd_sorted, 0, np, cd[j]));
  if (k_out) *k_out = cdf_pick(nc, u, [&](int64_t j) { return w_out[j]; });
}
__global__ void kat_philox_kernel(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
  Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

thread_local srw_walk_info t_info = {};
thread_local int t_collect_stats = 0;

// Per-thread launch context, created once: no cudaMalloc / cudaFree / event creation on the call
// path (those take driver-wide locks and serialise against other tools using the driver).
struct LaunchCtx {
  int device = -1;
  cudaEvent_t a = nullptr, b = nullptr, done = nullptr;
  bool pending = false;                    // kernels enqueued, srw_walk_wait not yet called
  unsigned long long *d_stats = nullptr;   // [4]
  unsigned long long *h_stats = nullptr;   // pinned [4]
  srw_status init(int dev) {
    if (device == dev) return SRW_OK;
    release();
    SRW_CUDA(cudaEventCreate(&a));
    SRW_CUDA(cudaEventCreate(&b));
    SRW_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    SRW_CUDA(cudaMalloc(&d_stats, 4 * sizeof(unsigned long long)));
    SRW_CUDA(cudaMallocHost(&h_stats, 4 * sizeof(unsigned long long)));
    device = dev;
    return SRW_OK;
  }
  void release() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    if (done) cudaEventDestroy(done);
    if (d_stats) cudaFree(d_stats);
    if (h_stats) cudaFreeHost(h_stats);
    a = b = done = nullptr; d_stats = h_stats = nullptr; device = -1; pending = false;
  }
};
thread_local LaunchCtx t_ctx;

}  // namespace

// Enqueues the walk kernel on l.stream and the rank -> id pass (+ the 32-byte statistics copy) on `fin`: the same stream
// for the blocking call, the library's high-priority finalisation stream for the asynchronous one -- there the caller's
// stream is free for the next round's walk as soon as this round's walk kernel has been launched.
static srw_status walk_enqueue(const srw_graph *g, const srw_params *p, const WalkLaunch &l, LaunchCtx &ev, cudaStream_t fin, bool own_fin) {
  if (p->walk_length < 0 || l.n_walkers < 0) { srw_set_error("walkLength and the walker count must be >= 0"); return SRW_ERR_ARG; }
  if (!(p->p > 0.0) || !(p->q > 0.0)) { srw_set_error("p and q must be > 0"); return SRW_ERR_ARG; }
  ev.pending = false;
  const bool peer = g->shard_world > 1;
  if (peer) {
    // one shard of several: only the peer-gather walk runs through this entry point (the tuple-exchange
    // walk is driven per super-step through srw_shard_step)
    for (int r = 0; r < g->shard_world; ++r)
      if (!g->peer_attached[r]) { srw_set_error("this handle is shard %d of %d and shard %d is not attached: attach every peer (srw_shard_attach_*) for the peer-gather walk, or use the srw_shard_* super-step calls", g->shard_rank, g->shard_world, r); return SRW_ERR_ARG; }
    if (p->sampler == SRW_SAMPLER_EXACT || g->directed || g->has_alias) { srw_set_error("the peer-gather walk covers undirected, unweighted graphs with --sampler alias|fold"); return SRW_ERR_UNSUPPORTED; }
  }
  if (l.n_walkers == 0 || g->nv == 0) return SRW_OK;
  const bool exact = p->sampler == SRW_SAMPLER_EXACT;
  if (p->sampler != SRW_SAMPLER_EXACT && p->sampler != SRW_SAMPLER_ALIAS && p->sampler != SRW_SAMPLER_ALIAS_FOLD) { srw_set_error("unknown sampler %d", p->sampler); return SRW_ERR_ARG; }
  if (exact && g->nnz > 0 && !g->d_col_app) { srw_set_error("graph was built without SRW_BUILD_EXACT"); return SRW_ERR_ARG; }
  if (!exact && p->u_mode == SRW_U_CONST) { srw_set_error("the constant-u generator is defined for --sampler exact only"); return SRW_ERR_ARG; }
  SRW_CUDA(cudaSetDevice(g->device));
  WalkArgs a{};
  a.off = g->d_off; a.col = g->d_col; a.slot = g->d_slot; a.col_app = g->d_col_app; a.w_app = g->d_w_app; a.vids = g->d_vids;
  a.nv = g->nv; a.walker_first = l.walker_first; a.n_walkers = l.n_walkers; a.stride = p->walk_length + 2;
  a.seed_lo = (uint32_t)p->seed; a.seed_hi = (uint32_t)(p->seed >> 32);
  srw_alias_thresholds(p->p, p->q, &a.t_ret, &a.t_common, &a.t_far);
  a.p = (float)p->p; a.q = (float)p->q; a.u_mode = p->u_mode; a.u_const = p->u_const;
  a.paths = l.d_paths; a.lens = l.d_lens;
  SRW_TRY(ev.init(g->device));
  unsigned long long *d_stats = ev.d_stats;
  SRW_CUDA(cudaMemsetAsync(d_stats, 0, 4 * sizeof(unsigned long long), l.stream));
  a.stats = d_stats;
  SRW_CUDA(cudaEventRecord(ev.a, l.stream));
  bool ids = false;                      // the walk kernel wrote original ids (id-space fold): no translation below
  if (exact) {
    const char *ex = getenv("SRW_EXACT");                                       // A/B switch (read per launch): thread | warp | cert | cert2 (default)
    if (ex && !strcmp(ex, "thread")) walk_exact_kernel<<<(unsigned)((l.n_walkers + 127) / 128), 128, 0, l.stream>>>(a);
    else if (ex && !strcmp(ex, "warp")) walk_exact_warp_kernel<<<(unsigned)((l.n_walkers + 7) / 8), 256, 0, l.stream>>>(a, g->d_hash);
    else if (ex && !strcmp(ex, "cert")) walk_exact_cert_kernel<<<(unsigned)((l.n_walkers + 7) / 8), 256, 0, l.stream>>>(a, g->d_hash);
    else walk_exact_cert2_kernel<<<(unsigned)((l.n_walkers + 7) / 8), 256, 0, l.stream>>>(a, g->d_hash);
  } else {
    const unsigned grid = (unsigned)((l.n_walkers + 255) / 256);
    const bool st = t_collect_stats != 0;
    static const bool use_v1 = getenv("SRW_KERNEL") && !strcmp(getenv("SRW_KERNEL"), "v1");   // A/B switches
    static const bool use_v2 = getenv("SRW_KERNEL") && !strcmp(getenv("SRW_KERNEL"), "v2");
    static const bool use_v3 = getenv("SRW_KERNEL") && !strcmp(getenv("SRW_KERNEL"), "v3");
    // SRW_SAMPLER_ALIAS_FOLD: undirected + unweighted + 1/p > max(1, 1/q), else the classic sampler
    // (the CPU twin applies the same rule, oracle_alias_walk)
    const int occ = getenv("SRW_FOLD_OCC") ? atoi(getenv("SRW_FOLD_OCC")) : 0;   // A/B (read per launch): resident blocks per SM the fold kernel is compiled for
    FoldArgs f{};
    bool fold = false;
    if ((peer || p->sampler == SRW_SAMPLER_ALIAS_FOLD) && (peer || (g->d_ent && g->d_hash)) && !g->directed && !g->has_alias) {
      fold = srw_fold_args(p->p, p->q, p->sampler == SRW_SAMPLER_ALIAS_FOLD, &f);
      f.ent = g->d_ent; f.hash = g->d_hash;
      if (peer && !fold) {
        // classic rejection under M = max(1/p, 1, 1/q) through the same kernel: no return component
        f.a = 0.0; f.mp = 1.0; f.t_ret = a.t_ret; f.t_common = a.t_common; f.t_far = a.t_far;
      }
    }
    // SRW_SAMPLER_ALIAS_FOLD on a weighted undirected graph: the same folding over bundle weights (walk_wfold_conv_kernel);
    // the CPU twin applies the same rule
    FoldArgs wf{};
    const bool wfold = !peer && p->sampler == SRW_SAMPLER_ALIAS_FOLD && g->has_alias && !g->directed && g->d_slotw && g->d_meta &&
                       g->d_hash && !use_v1 && !use_v2 && !use_v3 && srw_fold_args(p->p, p->q, true, &wf);
    // A/B (read per launch, so one process can time every variant on one graph): SRW_FOLD=v4 runs the
    // pre-convergence kernel; SRW_FOLD_VAR bit 0 = L2::64B loads in v5; SRW_FOLD_OCC = 5 | 6 blocks per SM
    const bool fold_v4 = getenv("SRW_FOLD") && !strcmp(getenv("SRW_FOLD"), "v4");
    constexpr int kFoldVarDefault = 1;   // v5 load flavour when SRW_FOLD_VAR is unset: L2::64B gathers (half the DRAM traffic at the same speed, profiles/)
    const int fold_var = getenv("SRW_FOLD_VAR") ? atoi(getenv("SRW_FOLD_VAR")) : kFoldVarDefault;
    ids = fold && !peer && !fold_v4 && g->ent_ids && g->d_hash_id;
    if (g->ent_ids && fold && !peer && !ids) { /* SRW_FOLD=v4 on an id-space handle */ srw_set_error("this graph was built in id space (SRW_FOLD_IDS): only the v5 alias-fold kernel can walk its neighbour entries"); return SRW_ERR_UNSUPPORTED; }
    if ((peer || fold) && !fold_v4) {
      PeerTable pt{};
      if (peer) {
        pt.world = g->shard_world;
        for (int r = 0; r <= g->shard_world; ++r) pt.first[r] = g->bounds[(size_t)r];
        for (int r = 0; r < g->shard_world; ++r) { pt.off[r] = g->peer_off[r]; pt.ent[r] = g->peer_ent[r]; pt.hash[r] = g->peer_hash[r]; }
      }
      const bool v64 = (fold_var & 1) != 0;
      if (peer) {
        if (st) walk_fold_conv_kernel<true, true, 0><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else if (v64) walk_fold_conv_kernel<false, true, 1><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else walk_fold_conv_kernel<false, true, 0><<<grid, 256, 0, l.stream>>>(a, f, pt);
      } else {
        if (ids) {                        // id space: entries and hash sets carry original ids, the walk emits ids (no rank -> id pass)
          f.hash = g->d_hash_id;
          if (st) walk_fold_conv_kernel<true, false, 0, 4, true><<<grid, 256, 0, l.stream>>>(a, f, pt);
          else if (v64) walk_fold_conv_kernel<false, false, 1, 4, true><<<grid, 256, 0, l.stream>>>(a, f, pt);
          else walk_fold_conv_kernel<false, false, 0, 4, true><<<grid, 256, 0, l.stream>>>(a, f, pt);
        }
        else if (st) walk_fold_conv_kernel<true, false, 0><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else if (v64 && occ == 5) walk_fold_conv_kernel<false, false, 1, 5><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else if (v64 && occ == 6) walk_fold_conv_kernel<false, false, 1, 6><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else if (v64) walk_fold_conv_kernel<false, false, 1><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else if (occ == 5) walk_fold_conv_kernel<false, false, 0, 5><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else if (occ == 6) walk_fold_conv_kernel<false, false, 0, 6><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else walk_fold_conv_kernel<false, false, 0><<<grid, 256, 0, l.stream>>>(a, f, pt);
      }
    } else if (peer) {
      PeerTable pt{};
      pt.world = g->shard_world;
      for (int r = 0; r <= g->shard_world; ++r) pt.first[r] = g->bounds[(size_t)r];
      for (int r = 0; r < g->shard_world; ++r) { pt.off[r] = g->peer_off[r]; pt.ent[r] = g->peer_ent[r]; pt.hash[r] = g->peer_hash[r]; }
      if (st) walk_fold_kernel<true, true><<<grid, 256, 0, l.stream>>>(a, f, pt);
      else if (occ == 2) walk_fold_kernel<false, true, 2><<<grid, 256, 0, l.stream>>>(a, f, pt);
      else if (occ == 5) walk_fold_kernel<false, true, 5><<<grid, 256, 0, l.stream>>>(a, f, pt);
      else if (occ == 6) walk_fold_kernel<false, true, 6><<<grid, 256, 0, l.stream>>>(a, f, pt);
      else if (occ == 8) walk_fold_kernel<false, true, 8><<<grid, 256, 0, l.stream>>>(a, f, pt);
      else walk_fold_kernel<false, true><<<grid, 256, 0, l.stream>>>(a, f, pt);
    } else if (fold) {
      const PeerTable pt{};
      if (st) walk_fold_kernel<true, false><<<grid, 256, 0, l.stream>>>(a, f, pt);
      else if (occ == 2) walk_fold_kernel<false, false, 2><<<grid, 256, 0, l.stream>>>(a, f, pt);
      else if (occ == 5) walk_fold_kernel<false, false, 5><<<grid, 256, 0, l.stream>>>(a, f, pt);
      else if (occ == 6) walk_fold_kernel<false, false, 6><<<grid, 256, 0, l.stream>>>(a, f, pt);
      else if (occ == 8) walk_fold_kernel<false, false, 8><<<grid, 256, 0, l.stream>>>(a, f, pt);
      else walk_fold_kernel<false, false><<<grid, 256, 0, l.stream>>>(a, f, pt);
    } else if (wfold) {
      // v5, weighted alias-fold: bundle weights in 32-byte slots, row weight sums in the descriptors
      const bool v64 = (fold_var & 1) != 0;
      if (st) walk_wfold_conv_kernel<true, 0><<<grid, 256, 0, l.stream>>>(a, wf, g->d_meta, g->d_hash, g->d_slotw);
      else if (v64) walk_wfold_conv_kernel<false, 1><<<grid, 256, 0, l.stream>>>(a, wf, g->d_meta, g->d_hash, g->d_slotw);
      else walk_wfold_conv_kernel<false, 0><<<grid, 256, 0, l.stream>>>(a, wf, g->d_meta, g->d_hash, g->d_slotw);
    } else if (!use_v1 && !use_v2 && !use_v3 && g->d_meta && g->d_hash) {
      // v5: the classic alias sampler in the warp-convergent layout (walk_conv.cuh)
      const RowMeta *mt = g->d_meta;
      const int32_t *hs = g->d_hash;
      const bool v64 = (fold_var & 1) != 0;
      if (g->has_alias) {
        if (st) walk_alias_conv_kernel<true, true, 0><<<grid, 256, 0, l.stream>>>(a, mt, hs);
        else if (v64) walk_alias_conv_kernel<true, false, 1><<<grid, 256, 0, l.stream>>>(a, mt, hs);
        else walk_alias_conv_kernel<true, false, 0><<<grid, 256, 0, l.stream>>>(a, mt, hs);
      } else {
        if (st) walk_alias_conv_kernel<false, true, 0><<<grid, 256, 0, l.stream>>>(a, mt, hs);
        else if (v64) walk_alias_conv_kernel<false, false, 1><<<grid, 256, 0, l.stream>>>(a, mt, hs);
        else walk_alias_conv_kernel<false, false, 0><<<grid, 256, 0, l.stream>>>(a, mt, hs);
      }
    } else if (!use_v1 && !use_v2 && g->d_meta) {
      const RowMeta *mt = g->d_meta;
      const int32_t *hs = g->d_hash;
      if (g->has_alias) { if (st) walk_alias_hash_kernel<true, true><<<grid, 256, 0, l.stream>>>(a, mt, hs); else walk_alias_hash_kernel<true, false><<<grid, 256, 0, l.stream>>>(a, mt, hs); }
      else              { if (st) walk_alias_hash_kernel<false, true><<<grid, 256, 0, l.stream>>>(a, mt, hs); else walk_alias_hash_kernel<false, false><<<grid, 256, 0, l.stream>>>(a, mt, hs); }
    } else if (!use_v1) {
      if (g->has_alias) { if (st) walk_alias_sm_kernel<true, true><<<grid, 256, 0, l.stream>>>(a); else walk_alias_sm_kernel<true, false><<<grid, 256, 0, l.stream>>>(a); }
      else              { if (st) walk_alias_sm_kernel<false, true><<<grid, 256, 0, l.stream>>>(a); else walk_alias_sm_kernel<false, false><<<grid, 256, 0, l.stream>>>(a); }
    } else if (g->has_alias) { if (st) walk_alias_kernel<true, true><<<grid, 256, 0, l.stream>>>(a); else walk_alias_kernel<true, false><<<grid, 256, 0, l.stream>>>(a); }
    else              { if (st) walk_alias_kernel<false, true><<<grid, 256, 0, l.stream>>>(a); else walk_alias_kernel<false, false><<<grid, 256, 0, l.stream>>>(a); }
  }
  SRW_CUDA(cudaEventRecord(ev.b, l.stream));
  if (own_fin) SRW_CUDA(cudaStreamWaitEvent(fin, ev.b, 0));
  {
    const int64_t total = l.n_walkers * (int64_t)a.stride;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    const bool flat = (a.stride & 1) == 0 && (reinterpret_cast<uintptr_t>(l.d_paths) & 7) == 0 && !getenv("SRW_FINALIZE_ROWS");
    if (ids) {
      count_steps_kernel<<<(unsigned)std::min<int64_t>((l.n_walkers + 255) / 256, 148 * 8), 256, 0, fin>>>(l.n_walkers, a.stride, l.d_lens, l.d_paths, d_stats);
    } else if (flat && total > 0) {
      const uint32_t half = (uint32_t)a.stride / 2;
      const uint64_t magic = ~0ULL / half + 1;                  // ceil(2^64 / half) (half >= 1; half == 1: wraps to 0, handled below)
      const int64_t n_pairs = total / 2;
      if (half == 1) finalize_paths_kernel<<<(unsigned)blocks, 256, 0, fin>>>(l.n_walkers, a.stride, g->d_vids, l.d_lens, l.d_paths, d_stats);
      else finalize_paths_flat_kernel<<<(unsigned)((n_pairs + 1023) / 1024), 256, 0, fin>>>(n_pairs, half, magic, g->d_vids, l.d_lens, reinterpret_cast<int2 *>(l.d_paths), d_stats);
    } else {
      finalize_paths_kernel<<<(unsigned)blocks, 256, 0, fin>>>(l.n_walkers, a.stride, g->d_vids, l.d_lens, l.d_paths, d_stats);
    }
  }
  SRW_CUDA(cudaMemcpyAsync(ev.h_stats, d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, fin));
  SRW_CUDA(cudaEventRecord(ev.done, fin));
  SRW_CUDA(cudaGetLastError());
  ev.pending = true;
  return SRW_OK;
}

// Blocks until an enqueued walk has finished and reports it.
static srw_status walk_finish(LaunchCtx &ev, srw_walk_info *out) {
  srw_walk_info wi{};
  if (ev.pending) {
    SRW_CUDA(cudaEventSynchronize(ev.done));
    SRW_CUDA(cudaGetLastError());
    float ms = 0.f;
    SRW_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    wi.kernel_ms = ms;
    wi.kernel_launches = 2;
    wi.steps = (int64_t)ev.h_stats[0];
    wi.proposals = (int64_t)ev.h_stats[1];
    wi.member_tests = (int64_t)ev.h_stats[2];
    wi.probes_log2 = (int64_t)ev.h_stats[3];
    ev.pending = false;
  }
  if (out) *out = wi;
  return SRW_OK;
}

srw_status srw_walk_launch(const srw_graph *g, const srw_params *p, const WalkLaunch &l) {
  t_info = srw_walk_info{};
  SRW_TRY(walk_enqueue(g, p, l, t_ctx, l.stream, false));
  return walk_finish(t_ctx, &t_info);
}

// ---- asynchronous rounds: tickets from a small pool, one high-priority finalisation stream per device ----
struct srw_walk_ticket {
  LaunchCtx ctx;
};
namespace {
std::mutex g_ticket_mu;
std::vector<srw_walk_ticket *> g_ticket_pool;
cudaStream_t g_fin_stream[64] = {};
srw_status fin_stream_for(int dev, cudaStream_t *out) {
  if (dev < 0 || dev >= 64) { srw_set_error("device index %d out of range", dev); return SRW_ERR_ARG; }
  std::lock_guard<std::mutex> lk(g_ticket_mu);
  if (!g_fin_stream[dev]) {
    int lo = 0, hi = 0;
    SRW_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SRW_CUDA(cudaStreamCreateWithPriority(&g_fin_stream[dev], cudaStreamNonBlocking, hi));
  }
  *out = g_fin_stream[dev];
  return SRW_OK;
}
}  // namespace

extern "C" srw_status srw_walk_device_async(const srw_graph *g, const srw_params *params, uint64_t walker_first, int64_t n_walkers,
                                            int32_t *d_paths, int32_t *d_lens, void *stream, srw_walk_ticket **ticket) {
  SRW_TRY(srw_require_device());
  if (!g || !params || !ticket || (n_walkers > 0 && (!d_paths || !d_lens))) { srw_set_error("srw_walk_device_async: bad argument"); return SRW_ERR_ARG; }
  SRW_CUDA(cudaSetDevice(g->device));
  cudaStream_t fin;
  SRW_TRY(fin_stream_for(g->device, &fin));
  srw_walk_ticket *t = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_ticket_mu);
    for (size_t i = 0; i < g_ticket_pool.size(); ++i)
      if (g_ticket_pool[i]->ctx.device == g->device || g_ticket_pool[i]->ctx.device < 0) { t = g_ticket_pool[i]; g_ticket_pool.erase(g_ticket_pool.begin() + (long)i); break; }
  }
  if (!t) t = new srw_walk_ticket();
  WalkLaunch l{walker_first, n_walkers, d_paths, d_lens, (cudaStream_t)stream};
  const srw_status st = walk_enqueue(g, params, l, t->ctx, fin, true);
  if (st != SRW_OK) {
    std::lock_guard<std::mutex> lk(g_ticket_mu);
    g_ticket_pool.push_back(t);
    return st;
  }
  *ticket = t;
  return SRW_OK;
}

extern "C" srw_status srw_walk_wait(srw_walk_ticket *ticket, srw_walk_info *info) {
  if (!ticket) return SRW_ERR_ARG;
  const srw_status st = walk_finish(ticket->ctx, info);
  std::lock_guard<std::mutex> lk(g_ticket_mu);
  g_ticket_pool.push_back(ticket);
  return st;
}

void srw_set_walk_info(double kernel_ms, int64_t launches, int64_t steps, int64_t proposals, int64_t member_tests, int64_t probes_log2) {
  t_info.kernel_ms = kernel_ms; t_info.kernel_launches = launches; t_info.steps = steps;
  t_info.proposals = proposals; t_info.member_tests = member_tests; t_info.probes_log2 = probes_log2;
}

extern "C" srw_status srw_last_walk_info(srw_walk_info *out) {
  if (!out) return SRW_ERR_ARG;
  *out = t_info;
  return SRW_OK;
}
extern "C" srw_status srw_walk_collect_stats(int enable) {
  t_collect_stats = enable;
  return SRW_OK;
}

// ---- KAT entry points (RS:12-62 on the device) ----
namespace {
template <class T>
struct DevArr {
  T *p = nullptr;
  ~DevArr() { if (p) cudaFree(p); }
  cudaError_t upload(const T *h, int64_t n) {
    cudaError_t e = cudaMalloc(&p, (size_t)(n > 0 ? n : 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    return (n > 0 && h) ? cudaMemcpy(p, h, (size_t)n * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
  }
};
}  // namespace

extern "C" srw_status srw_sample(int64_t n, const int32_t *h_dst, const float *h_w, float u, int32_t *dst_out, float *w_out) {
  SRW_TRY(srw_require_device());
  if (n <= 0 || !h_dst || !h_w) { srw_set_error("srw_sample: empty edge array (reference: edges.head on empty throws)"); return SRW_ERR_ARG; }
  DevArr<float> w; DevArr<int64_t> k;
  SRW_CUDA(w.upload(h_w, n));
  SRW_CUDA(k.upload(nullptr, 1));
  kat_sample_kernel<<<1, 1>>>(n, w.p, u, k.p);
  int64_t hk = 0;
  SRW_CUDA(cudaMemcpy(&hk, k.p, 8, cudaMemcpyDeviceToHost));
  if (dst_out) *dst_out = h_dst[hk];
  if (w_out) *w_out = h_w[hk];
  return SRW_OK;
}

static srw_status second_order(float p, float q, int32_t prev, int64_t np, const int32_t *h_pdst, int64_t nc,
                               const int32_t *h_cdst, const float *h_cw, float u, float *h_out, int64_t *k_out) {
  SRW_TRY(srw_require_device());
  if (nc <= 0 || !h_cdst || !h_cw) { srw_set_error("second-order sample: empty currNeighbors"); return SRW_ERR_ARG; }
  std::vector<int32_t> ps(h_pdst, h_pdst + (np > 0 ? np : 0));
  std::sort(ps.begin(), ps.end());
  DevArr<int32_t> pd, cd; DevArr<float> cw, wo; DevArr<int64_t> k;
  SRW_CUDA(pd.upload(ps.data(), np));
  SRW_CUDA(cd.upload(h_cdst, nc));
  SRW_CUDA(cw.upload(h_cw, nc));
  SRW_CUDA(wo.upload(nullptr, nc));
  SRW_CUDA(k.upload(nullptr, 1));
  kat_second_order_kernel<<<1, 1>>>(p, q, prev, np, pd.p, nc, cd.p, cw.p, u, wo.p, k_out ? k.p : nullptr);
  SRW_CUDA(cudaMemcpy(h_out, wo.p, (size_t)nc * 4, cudaMemcpyDeviceToHost));
  if (k_out) SRW_CUDA(cudaMemcpy(k_out, k.p, 8, cudaMemcpyDeviceToHost));
  return SRW_OK;
}

extern "C" srw_status srw_second_order_weights(float p, float q, int32_t prev, int64_t np, const int32_t *h_pdst, int64_t nc,
                                               const int32_t *h_cdst, const float *h_cw, float *h_out) {
  if (!h_out) return SRW_ERR_ARG;
  return second_order(p, q, prev, np, h_pdst, nc, h_cdst, h_cw, 0.f, h_out, nullptr);
}
extern "C" srw_status srw_second_order_sample(float p, float q, int32_t prev, int64_t np, const int32_t *h_pdst, int64_t nc,
                                              const int32_t *h_cdst, const float *h_cw, float u, int32_t *dst_out, float *w_out) {
  std::vector<float> w((size_t)(nc > 0 ? nc : 1));
  int64_t k = 0;
  SRW_TRY(second_order(p, q, prev, np, h_pdst, nc, h_cdst, h_cw, u, w.data(), &k));
  if (dst_out) *dst_out = h_cdst[k];
  if (w_out) *w_out = w[k];     // the BIASED weight (T-RS:76)
  return SRW_OK;
}
extern "C" srw_status srw_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  SRW_TRY(srw_require_device());
  DevArr<uint32_t> c, k, o;
  SRW_CUDA(c.upload(ctr, 4)); SRW_CUDA(k.upload(key, 2)); SRW_CUDA(o.upload(nullptr, 4));
  kat_philox_kernel<<<1, 1>>>(c.p, k.p, o.p);
  SRW_CUDA(cudaMemcpy(out, o.p, 16, cudaMemcpyDeviceToHost));
  return SRW_OK;
}
