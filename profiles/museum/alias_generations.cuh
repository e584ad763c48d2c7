// alias_generations.cuh -- MUSEUM, not product code: the alias-sampler kernel generations that preceded the warp-convergent
// kernels of csrc/walk_conv.cuh (v1 loop nest, v2 per-lane state machine, v3 hash-set membership + packed row descriptor,
// v4 alias-fold state machine).  Kept buildable (walk_museum.cu -> libsrw_museum.so, a TEST-ONLY library) because the A/B
// history in profiles/README.md refers to them and tests/test_gpu_parity.py checks that every generation produces the bits of
// the product kernel.  Nothing in libsrw.so includes this file.
#pragma once

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

