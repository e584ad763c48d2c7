// migrate.cuh -- K7/K8 fused: one super-step of the vertex-range-sharded walk in which the WALKERS MIGRATE and the
// step kernel itself performs the exchange (reference seam RW:91-162 super-step loop, RW:186-192 shuffle,
// URW:103-112 routing key = current vertex).
//
// Every GPU owns the rows of one contiguous vertex range (16-byte neighbour entries + per-row hash sets, layout.h).  A
// walker lives on owner(curr); a sampled step to x moves it to owner(x) as a 32-byte tuple that the kernel stores STRAIGHT
// INTO THE DESTINATION GPU'S INBOX over NVLink (peer pointers: symmetric memory between processes, plain cudaMalloc inside
// one process).  Inboxes are double-buffered; super-step s reads buffer s&1 and fills the peers' buffer (s+1)&1.  Each
// (source, destination) pair owns a fixed region of the destination inbox, so slots are claimed with LOCAL atomics only
// (32 slots at a time per warp and destination) and nothing on the data path waits for a remote round trip; a full region
// spills to a local queue and is re-sent in the next super-step.  The per-region counts are published to the destination
// by the last warp of the kernel; a barrier between super-steps (NCCL all-reduce of the tuple count, which is also the
// RW:162 termination test; CUDA events inside one process) is the only collective.
//
// The second-order membership test d(t, x) = 1 (RS:38) needs N(prev), which lives on owner(prev).  It is answered where
// the walker already is or is going anyway:
//   * a REPLICATED edge filter (one 64-bit Bloom word per probe, 16 bits per undirected edge = 1 byte per adjacency entry
//     against 24 bytes of sharded rows) says "definitely not adjacent" for ~99.6 % of the non-adjacent pairs;
//   * "maybe adjacent" is verified EXACTLY at owner(x) by the symmetric test t in N(x) in x's own hash set (undirected:
//     x in N(t) <=> t in N(x)).  If accepted the walker is already where its next step happens; only a false positive
//     of the filter (or a rejected member when q < 1) costs a bounce back to owner(curr).
// => ~(1 - 1/world) hops per step instead of ~2.8 with the test at owner(prev) (round 1), identical decisions: every draw
// is the pure function Philox(seed; walker, step, trial) of the single-GPU kernel (walk_fold_conv_kernel) and of the CPU twin.
//
// Path entries go to the walker's HOME GPU (home(v) = v mod world, as shard.cu) as peer stores into its path matrix -- but not
// one by one: a 4-byte store into a matrix far larger than L2 costs a DRAM read-modify-write on the home GPU, about as much as
// one of the step's own gathers (measured: the first version ran at 8.5e9 steps/s per GPU against 13.3e9 with local paths).
// The walker therefore CARRIES up to three decided entries (in registers while resident, in the third 16-byte word of its
// 48-byte tuple when it migrates) and stores them as ONE aligned 16-byte chunk when the fourth arrives: chunk boundaries are
// the multiples of 4 of the GLOBAL int index row * stride + pos, so no row padding is needed and the chunk phase of a row (2
// bits) rides in the tuple.  Undirected, unweighted graphs; samplers alias / alias-fold (same thresholds as walk_conv.cuh).
//
// The loop is the warp-convergent three-phase layout of walk_conv.cuh (draw / one access per lane / consume) plus a refill
// phase (a lane whose walker left or finished takes the next inbox tuple: lanes never idle to the end of the warp's longest
// walk) and a send phase.  Compiled for the host by tests/emu (warp_emu.h) to check the logic against the CPU twin.
#pragma once
#include <stdint.h>

#include "layout.h"
#include "philox.cuh"
#include "walk_conv.cuh"

enum : uint32_t { MIG_SETTLED = 0, MIG_PENDING = 1, MIG_NOP = 2, MIG_KIND_MASK = 3, MIG_NEEDEXT = 0x10, MIG_PHASE_SHIFT = 5 /* bits 5-6: (row * stride) & 3 */ };
enum : int { MS_EMPTY = 0, MS_LOAD, MS_LOADEXT, MS_EXTENT, MS_WAIT, MS_PROPOSE, MS_BLOOM, MS_HASH, MS_SEARCH };

constexpr int kMigChunk = 32;        // inbox slots a warp claims at a time per destination
constexpr int kMigClaim = 64;        // inbox items a warp claims at a time
constexpr int kMigMaxDest = SRW_MAX_SHARDS + 1;   // peers + the local spill region

struct MigArgs {
  // this shard's rows
  const int64_t *__restrict__ off;        // [rows + 1] shard-local offsets
  const NbrEntry *__restrict__ ent;       // [nnz_local]
  const int32_t *__restrict__ hash;       // per-row hash sets of neighbour RANKS, placement derived from (off, deg)
  const unsigned long long *__restrict__ bloom;   // replicated edge filter
  uint64_t bloom_words;
  int64_t nv, row_first, row_last;
  int world, rank;
  int64_t bounds[SRW_MAX_SHARDS + 1];
  // sampler (walk_conv.cuh FoldArgs semantics)
  double a, mp;
  uint64_t t_ret, t_common, t_far;
  uint32_t seed_lo, seed_hi;
  int32_t stride;
  uint64_t walker_base;                   // round_first * nv: batch-local walker w is global walker walker_base + w
  int64_t n_rounds;
  // inbox of THIS super-step: regions 0..world-1 (filled by the peers), region `world` (local spill), then n_seed virtual seeds
  const int4 *__restrict__ in_base;       // 3 x int4 per slot (MigTuple)
  const int4 *__restrict__ in_ext;        // 1 x int4 per slot
  const unsigned long long *__restrict__ in_cnt;   // [world + 1] slots used per region (published by the senders)
  int64_t seg_cap, spill_cap;             // slots per peer region / in the spill region
  int64_t n_seed;                         // virtual seeds of THIS super-step: seed j is walker number seed_first + j * seed_step of the
  int64_t seed_first, seed_step;          // rows_local * n_rounds walkers this shard starts (super-steps 0 and 1 take every other one:
                                          // a shard's whole population leaves in one super-step, so on two shards an uneven start would
                                          // slosh back and forth for the whole walk; staggering the seeds damps that mode at once)
  // destinations: region `rank` of every peer's NEXT inbox (index world = own spill region)
  int4 *out_base[kMigMaxDest];
  int4 *out_ext[kMigMaxDest];
  unsigned long long *out_cnt_pub[kMigMaxDest];   // where the slot count of that region is published (peer memory)
  int32_t *home_paths[SRW_MAX_SHARDS];    // path matrix of every home GPU: [n_rounds * home_rows[h]][stride]
  int64_t home_rows[SRW_MAX_SHARDS];      // vertices v with v mod world == h
  // local scratch (device memory of this GPU)
  unsigned long long *cursor;             // inbox work cursor
  unsigned long long *out_cnt;            // [world + 1] slots claimed per destination region
  unsigned long long *done_warps;
  unsigned long long *stats;              // [0] slots sent this super-step (written by the last warp), [1] steps, [2] proposals, [3] tests, [4] exact tests, [5] spills, [6] error flags
};

struct MigTuple {            // 48 bytes: three 16-byte words
  uint32_t walker;           // batch-local
  int32_t prev, curr;
  uint32_t off, deg;         // row extent of curr inside owner(curr)'s arrays (invalid when MIG_NEEDEXT)
  uint32_t m_kind;           // [31:8] parallel edges curr-prev, [7:0] kind | flags | chunk phase of the walker's path row
  uint32_t trial;
  uint32_t len;              // ids already in the path
  int32_t carry[3];          // decided path entries not yet stored: positions len - n .. len - 1, n = mig_carried(phase, len)
  uint32_t pad;
};
struct MigExt {              // 16 bytes, only for MIG_PENDING: the proposal under test
  int32_t x;
  uint32_t xoff, xdeg;
  uint32_t xm_own;           // [31:8] parallel edges curr-x, [7:0] owner(x)
};

// number of path entries a walker with `len` ids holds back (positions >= 1 only: position 0 is written by the home GPU)
__device__ __forceinline__ uint32_t mig_carried(uint32_t phase, uint32_t len) {
  if (len <= 1) return 0;
  const uint32_t in_chunk = ((phase + len - 1u) & 3u) + 1u;        // entries of the chunk that position len - 1 belongs to, up to it
  const uint32_t n = in_chunk == 4u ? 0u : in_chunk;               // a complete chunk was stored when its fourth entry arrived
  return n < len - 1u ? n : len - 1u;
}

__device__ __forceinline__ int mig_owner(const MigArgs &a, int32_t v) {
  int o = 0;
  while (o + 1 < a.world && (int64_t)v >= a.bounds[o + 1]) o++;
  return o;
}

#ifdef SRW_EMU
static inline void __threadfence_system() {}
static inline void __threadfence() {}
static inline unsigned mig_reduce_or(unsigned v) {
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#else
__device__ __forceinline__ unsigned mig_reduce_or(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
#endif

template <bool STATS>
__global__ void __launch_bounds__(256, 3) mig_step_kernel(const MigArgs a) {
  // per-warp send state: open chunk (first slot, slots used) per destination region
  __shared__ unsigned long long s_chunk[8][kMigMaxDest];
  __shared__ unsigned int s_used[8][kMigMaxDest];
  __shared__ unsigned long long s_pre[8][kMigMaxDest + 2];   // prefix of the inbox regions (+ seeds)
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int W = a.world, me = a.rank;
  unsigned long long *chunk = s_chunk[wib];
  unsigned int *used = s_used[wib];
  unsigned long long *pre = s_pre[wib];
  if (lane <= W) { chunk[lane] = 0; used[lane] = kMigChunk; }
  if (lane == 0) {
    unsigned long long acc = 0;
    for (int r = 0; r <= W; ++r) { pre[r] = acc; acc += a.in_cnt ? a.in_cnt[r] : 0ull; }
    pre[W + 1] = acc;
    pre[W + 2] = acc + (unsigned long long)a.n_seed;
  }
  __syncwarp();
  const unsigned long long total = pre[W + 2];
  const bool acc_member = a.t_common > a.t_far, acc_non = a.t_far > a.t_common;   // verdict of a test that HAD to run (t_lo <= y < t_hi)
  const uint64_t t_lo = a.t_common < a.t_far ? a.t_common : a.t_far;
  const uint64_t t_hi = a.t_common < a.t_far ? a.t_far : a.t_common;

  // lane state: one walker
  uint32_t walker = 0, off = 0, deg = 0, m = 1, trial = 0, len = 0, poff = 0, pdeg = 0;
  int32_t prev = -1, curr = 0, x = 0;
  uint32_t xoff = 0, xdeg = 0, xm = 1, xown = 0, cown = 0, k = 0, y = 0, bkt = 0, pnb = 0, lo = 0, hi = 0;
  uint32_t phase = 0;                     // (row * stride) & 3 of the walker's home path row
  int32_t c0 = 0, c1 = 0, c2 = 0;         // carried path entries, oldest first
  bool pvalid = false;
  unsigned long long item = 0;
  uint64_t bmask = 0, bword = 0;
  int st = MS_EMPTY;
  unsigned long long w_next = 0, w_end = 0;
  bool exhausted = total == 0;
  unsigned long long n_steps = 0, n_prop = 0, n_test = 0, n_exact = 0, n_spill = 0, n_err = 0;

  for (;;) {
    // ---- R: refill empty lanes from the inbox ----
    const unsigned em = __ballot_sync(0xffffffffu, st == MS_EMPTY);
    if (em && !exhausted) {
      if (w_next >= w_end) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.cursor, (unsigned long long)kMigClaim);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) { exhausted = true; w_next = w_end = 0; }
        else { w_next = base; w_end = base + kMigClaim < total ? base + kMigClaim : total; }
      }
      const unsigned long long mine = w_next + (unsigned long long)__popc(em & lt);
      if (st == MS_EMPTY && mine < w_end) {
        item = mine;
        if (item >= pre[W + 1]) {                        // a virtual seed: walker (round, row) of this shard, path = [v]
          const unsigned long long j = (unsigned long long)a.seed_first + (item - pre[W + 1]) * (unsigned long long)a.seed_step;
          const int64_t rows = a.row_last - a.row_first;
          const int64_t round = (int64_t)(j / (unsigned long long)rows), row = (int64_t)(j % (unsigned long long)rows);
          curr = (int32_t)(a.row_first + row); prev = -1;
          walker = (uint32_t)((unsigned long long)round * (unsigned long long)a.nv + (unsigned long long)curr);
          m = 1; trial = 0; len = 1; cown = (uint32_t)me; pvalid = false;
          {
            const uint32_t h = (uint32_t)curr % (uint32_t)W;
            const int64_t prow = round * a.home_rows[h] + (int64_t)((uint32_t)curr / (uint32_t)W);
            phase = (uint32_t)((prow * a.stride) & 3);
          }
          st = MS_EXTENT;
        } else {
          st = MS_LOAD;
        }
      }
      const unsigned long long adv = w_next + (unsigned long long)__popc(em);
      w_next = adv < w_end ? adv : w_end;
    }
    if (!__any_sync(0xffffffffu, st != MS_EMPTY)) {
      if (exhausted) break;
      continue;
    }

    int send = -1;                 // destination region of the tuple this lane emits in this iteration
    uint32_t send_kind = MIG_SETTLED;
    bool moved = false;
    int32_t newv = 0;
    bool needext = false;
    // ---- A: draw ----
    if (st == MS_WAIT) {
      const Philox4 r = walker_rng(a.seed_lo, a.seed_hi, a.walker_base + (uint64_t)walker, len - 1u, trial);
      bool ret = false;
      if (len > 1) {                                     // P(return-excess component) = a*m / (Mp*deg + a*m)   (walk_conv.cuh)
        const double t1 = __dmul_rn(a.a, (double)m), t2 = __dmul_rn(a.mp, (double)deg);
        ret = __dmul_rn((double)r.y, __dadd_rn(t2, t1)) < __dmul_rn(t1, 4294967296.0);
      }
      if (ret) {                                         // always accepted, no memory access
        if (STATS) n_prop++;
        newv = prev; moved = true;
        const int32_t c = curr; curr = prev; prev = c;
        const uint32_t o = off, d = deg;
        if (pvalid) { off = poff; deg = pdeg; } else needext = true;
        poff = o; pdeg = d; pvalid = true;               // m unchanged: the same bundle of parallel edges
        cown = (uint32_t)mig_owner(a, curr);
      } else {
        k = (uint32_t)__umul64hi(((uint64_t)r.x << 32) | (uint64_t)r.w, (uint64_t)deg);
        y = r.z;
        st = MS_PROPOSE;
      }
    }
    __syncwarp();
    // ---- B: one memory access per lane ----
    int4 q0 = make_int4(0, 0, 0, 0), q1 = make_int4(0, 0, 0, 0), q2 = make_int4(0, 0, 0, 0);
    int64_t e0 = 0, e1 = 0;
    unsigned long long bw = 0;
    if (!moved) {
      if (st == MS_LOAD) {
        unsigned long long slot = 0;
        int r = 0;
        while (r < W && item >= pre[r + 1]) r++;
        slot = (unsigned long long)r * (unsigned long long)a.seg_cap + (item - pre[r]);
        const int4 *tp = a.in_base + 3 * slot;
        q0 = gather16<0>(tp); q1 = gather16<0>(tp + 1); q2 = gather16<0>(tp + 2);
        item = slot;                                      // MS_LOADEXT reads the same slot
      } else if (st == MS_HASH) {
        gather32<1>(reinterpret_cast<const int4 *>(a.hash + ((uint64_t)(xoff >> 2) + bkt) * 8), q0, q1);
      } else if (st == MS_LOADEXT) {
        q0 = gather16<0>(a.in_ext + item);
      } else if (st == MS_PROPOSE) {
        q0 = gather16<1>(reinterpret_cast<const int4 *>(a.ent + ((uint64_t)off + k)));
      } else if (st == MS_SEARCH) {
        q0 = gather16<1>(reinterpret_cast<const int4 *>(a.ent + ((uint64_t)xoff + ((lo + hi) >> 1))));
      } else if (st == MS_EXTENT) {
        const int64_t *o = a.off + ((int64_t)curr - a.row_first);
        e0 = __ldg(o); e1 = __ldg(o + 1);
      } else if (st == MS_BLOOM) {
        bw = __ldg(a.bloom + bword);
      }
    }
    __syncwarp();
    // ---- C: consume ----
    int verdict = 0;               // 1 = accept x, 2 = reject (next trial)
    int member = -1;
    if (!moved) {
      if (st == MS_LOAD) {
        walker = (uint32_t)q0.x; prev = q0.y; curr = q0.z; off = (uint32_t)q0.w;
        deg = (uint32_t)q1.x; m = (uint32_t)q1.y >> 8; trial = (uint32_t)q1.z; len = (uint32_t)q1.w;
        phase = ((uint32_t)q1.y >> MIG_PHASE_SHIFT) & 3u;
        c0 = q2.x; c1 = q2.y; c2 = q2.z;
        const uint32_t kind = (uint32_t)q1.y & MIG_KIND_MASK;
        pvalid = false;
        if (kind == MIG_NOP) st = MS_EMPTY;
        else if (kind == MIG_PENDING) st = MS_LOADEXT;
        else {
          cown = (uint32_t)mig_owner(a, curr);
          if ((int)cown != me) { send = (int)cown; send_kind = (uint32_t)q1.y & (MIG_KIND_MASK | MIG_NEEDEXT); }    // spilled last super-step: forward as it is
          else st = ((uint32_t)q1.y & MIG_NEEDEXT) ? MS_EXTENT : MS_WAIT;
        }
      } else if (st == MS_LOADEXT) {
        x = q0.x; xoff = (uint32_t)q0.y; xdeg = (uint32_t)q0.z; xm = (uint32_t)q0.w >> 8; xown = (uint32_t)q0.w & 0xFFu;
        cown = (uint32_t)mig_owner(a, curr);
        if ((int)xown != me) { send = (int)xown; send_kind = MIG_PENDING; }                 // spilled: forward
        else {                                                                              // the exact test t in N(x), in x's own row
          if (STATS) n_exact++;
          pnb = srw_hash_buckets((int64_t)xoff, xdeg);
          if (pnb) { bkt = __umulhi(srw_hash32((uint32_t)prev), pnb); st = MS_HASH; }
          else { lo = 0; hi = xdeg; st = MS_SEARCH; }
        }
      } else if (st == MS_EXTENT) {
        off = (uint32_t)e0; deg = (uint32_t)(e1 - e0);
        if (deg == 0) { n_err |= 1; st = MS_EMPTY; }      // cannot happen on an undirected graph (every vertex has an entry)
        else st = MS_WAIT;
      } else if (st == MS_PROPOSE) {
        x = q0.x; xdeg = (uint32_t)q0.y; xoff = (uint32_t)q0.z;
        xown = (uint32_t)q0.w & 0xFFu; xm = (uint32_t)q0.w >> 8;
        if (STATS && len > 1) n_prop++;
        if (len == 1 || deg == 1) verdict = 1;                             // first-order step (RW:57) / single choice
        else if (x == prev) verdict = ((uint64_t)y < a.t_ret) ? 1 : 2;     // RS:36
        else if ((uint64_t)y < t_lo) verdict = 1;
        else if ((uint64_t)y >= t_hi) verdict = 2;
        else {                                                             // RS:38 needs d(prev, x): ask the replicated filter first
          if (STATS) n_test++;
          srw_bloom_probe(prev, x, a.bloom_words, &bword, &bmask);
          st = MS_BLOOM;
        }
      } else if (st == MS_BLOOM) {
        if ((bw & bmask) != bmask) member = 0;                             // definitely not adjacent
        else if ((int)xown != me) { send = (int)xown; send_kind = MIG_PENDING; }   // verify where the walker would go anyway
        else {
          if (STATS) n_exact++;
          pnb = srw_hash_buckets((int64_t)xoff, xdeg);
          if (pnb) { bkt = __umulhi(srw_hash32((uint32_t)prev), pnb); st = MS_HASH; }
          else { lo = 0; hi = xdeg; st = MS_SEARCH; }
        }
      } else if (st == MS_HASH) {
        const int32_t t = prev;
        const bool found = q0.x == t || q0.y == t || q0.z == t || q0.w == t || q1.x == t || q1.y == t || q1.z == t || q1.w == t;
        if (found) member = 1;
        else if (q1.w == -1) member = 0;
        else bkt = bkt + 1 == pnb ? 0 : bkt + 1;
      } else if (st == MS_SEARCH) {
        const uint32_t mid = (lo + hi) >> 1;
        if (q0.x == prev) member = 1;
        else {
          if (q0.x < prev) lo = mid + 1; else hi = mid;
          if (lo >= hi) member = 0;
        }
      }
      if (member >= 0) verdict = (member ? acc_member : acc_non) ? 1 : 2;
    }
    if (verdict == 1) {                                    // move along entry (x, xoff, xdeg, xm, xown)
      newv = x; moved = true;
      prev = curr; poff = off; pdeg = deg; pvalid = true;
      curr = x; off = xoff; deg = xdeg; m = xm; cown = xown;
    } else if (verdict == 2) {
      trial++;
      if ((int)cown == me) st = MS_WAIT;
      else { send = (int)cown; send_kind = MIG_SETTLED; }  // the test ran at owner(x): back to the row of curr
    }
    if (moved) {                                           // RW:114: the step is decided -> the walker's home path row, four entries at a time
      const uint32_t pos = len, in_chunk = (phase + pos) & 3u;          // position of newv; its place in its 16-byte chunk
      const uint32_t have = mig_carried(phase, len);                    // entries carried so far (all of this chunk)
      if (in_chunk == 3u || (int32_t)(pos + 1u) == a.stride) {
        // the chunk is complete (or the path ends): store carried + newv, positions pos - have .. pos
        const uint32_t v0 = walker % (uint32_t)a.nv, rnd = walker / (uint32_t)a.nv;
        const uint32_t h = v0 % (uint32_t)W;
        int32_t *dst = a.home_paths[h] + ((int64_t)rnd * a.home_rows[h] + (int64_t)(v0 / (uint32_t)W)) * a.stride + (pos - have);
        if (have == 3u && in_chunk == 3u) *reinterpret_cast<int4 *>(dst) = make_int4(c0, c1, c2, newv);
        else if (have == 0u) dst[0] = newv;
        else if (have == 1u) { dst[0] = c0; dst[1] = newv; }
        else if (have == 2u) { dst[0] = c0; dst[1] = c1; dst[2] = newv; }
        else { dst[0] = c0; dst[1] = c1; dst[2] = c2; dst[3] = newv; }
      } else if (have == 0u) c0 = newv;
      else if (have == 1u) c1 = newv;
      else c2 = newv;
      len++; trial = 0;
      if (STATS) n_steps++;
      if ((int32_t)len == a.stride) st = MS_EMPTY;         // RW:103,132
      else if ((int)cown == me) st = needext ? MS_EXTENT : MS_WAIT;
      else { send = (int)cown; send_kind = MIG_SETTLED | (needext ? MIG_NEEDEXT : 0u); }
    }
    // ---- D: sends (tuples to the next inbox of their destination) ----
    unsigned dmask = mig_reduce_or(send >= 0 ? 1u << send : 0u);        // destinations some lane sends to in this iteration
    while (dmask) {
      {
        const int d = __ffs(dmask) - 1;
        dmask &= dmask - 1;
        const unsigned sm = __ballot_sync(0xffffffffu, send == d);
        if (!sm) continue;
        const unsigned n = (unsigned)__popc(sm);
        unsigned u = used[d];
        unsigned long long cb = chunk[d];
        bool full = false;
        __syncwarp();
        if (u + n > (unsigned)kMigChunk) {
          // close the open chunk (pad with NOPs) and claim the next one
          if (u + (unsigned)lane < (unsigned)kMigChunk) a.out_base[d][3 * (cb + u + (unsigned)lane) + 1] = make_int4(0, (int)MIG_NOP, 0, 0);
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(a.out_cnt + d, (unsigned long long)kMigChunk);
          base = __shfl_sync(0xffffffffu, base, 0);
          const unsigned long long cap = d == W ? (unsigned long long)a.spill_cap : (unsigned long long)a.seg_cap;
          if (base + kMigChunk > cap) {
            full = true;
            if (lane == 0) { used[d] = kMigChunk; atomicAdd(a.out_cnt + d, (unsigned long long)(0ull - (unsigned long long)kMigChunk)); }
          } else {
            cb = base; u = 0;
            if (lane == 0) chunk[d] = base;
          }
        }
        if (full) {
          if (send == d) {
            if (d == W) { n_err |= 2; send = -1; st = MS_EMPTY; }   // the spill region is sized for every walker of the batch: cannot happen
            else { send = W; n_spill++; }                          // region full: park locally, forwarded next super-step
          }
          if (d != W) dmask |= 1u << W;                          // (warp-uniform: sm != 0) the spill region comes last
          __syncwarp();
          continue;
        }
        if (send == d) {
          const unsigned long long slot = cb + u + (unsigned)__popc(sm & lt);
          int4 *p = a.out_base[d] + 3 * slot;
          p[0] = make_int4((int)walker, prev, curr, (int)off);
          p[1] = make_int4((int)deg, (int)((m << 8) | (phase << MIG_PHASE_SHIFT) | send_kind), (int)trial, (int)len);
          p[2] = make_int4(c0, c1, c2, 0);
          if ((send_kind & MIG_KIND_MASK) == MIG_PENDING) a.out_ext[d][slot] = make_int4(x, (int)xoff, (int)xdeg, (int)((xm << 8) | xown));
          st = MS_EMPTY;
        }
        if (lane == 0) used[d] = u + n;
        __syncwarp();
      }
    }
  }
  // pad the open chunks, then hand the counts over
  for (int d = 0; d <= W; ++d) {
    const unsigned u = used[d];
    if (u + (unsigned)lane < (unsigned)kMigChunk) a.out_base[d][3 * (chunk[d] + u + (unsigned)lane) + 1] = make_int4(0, (int)MIG_NOP, 0, 0);
  }
  if (STATS) {
    for (int o = 16; o > 0; o >>= 1) {
      n_steps += __shfl_down_sync(0xffffffffu, n_steps, o); n_prop += __shfl_down_sync(0xffffffffu, n_prop, o);
      n_test += __shfl_down_sync(0xffffffffu, n_test, o); n_exact += __shfl_down_sync(0xffffffffu, n_exact, o);
    }
    if (lane == 0) {
      if (n_steps) atomicAdd(a.stats + 1, n_steps);
      if (n_prop) atomicAdd(a.stats + 2, n_prop);
      if (n_test) atomicAdd(a.stats + 3, n_test);
      if (n_exact) atomicAdd(a.stats + 4, n_exact);
    }
  }
  if (n_spill) atomicAdd(a.stats + 5, n_spill);
  if (n_err) atomicAdd(a.stats + 6, n_err);
  __threadfence_system();
  __syncwarp();
  unsigned long long fin = 0;
  if (lane == 0) fin = atomicAdd(a.done_warps, 1ull);
  fin = __shfl_sync(0xffffffffu, fin, 0);
  if (fin + 1 == (unsigned long long)gridDim.x * (blockDim.x >> 5)) {
    // last warp of the grid: every region's slot count goes to its destination; local counters are reset for the next launch
    __threadfence();
    unsigned long long c = 0;
    if (lane <= W) {
      c = atomicAdd(a.out_cnt + lane, 0ull);
      *a.out_cnt_pub[lane] = c;
      a.out_cnt[lane] = 0;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if (lane == 0) { a.stats[0] = c; *a.cursor = 0; *a.done_warps = 0; }
    __threadfence_system();
  }
}
