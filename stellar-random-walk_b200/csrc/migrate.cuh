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
// NVLink moves ~1e10 store REQUESTS per second and GPU whatever their size (measured: 2, 4 and 8 GPUs all levelled at ~230 GB/s
// of 16-byte peer stores per GPU while the same kernel ran twice as fast with every shard on one device), so a departing walker
// is first collected in SHARED MEMORY -- kMigStage tuples per warp and destination -- and a full stage leaves as one coalesced
// copy: three runs of 256 contiguous bytes.
//
// Path entries go to the walker's HOME GPU (home(v) = v mod world, as shard.cu) as peer stores into its path matrix -- but not
// one by one: a 4-byte store into a matrix far larger than L2 costs a DRAM read-modify-write on the home GPU, about as much as
// one of the step's own gathers (measured: the first version ran at 8.5e9 steps/s per GPU against 13.3e9 with local paths).
// The walker therefore CARRIES up to three decided entries (in registers while resident, in the third 16-byte word of its
// 48-byte tuple when it migrates) and stores them as ONE aligned 16-byte chunk when the fourth arrives: chunk boundaries are
// the multiples of 4 of the GLOBAL int index row * stride + pos, so no row padding is needed and the chunk phase of a row (2
// bits) rides in the tuple.  Undirected, unweighted graphs; samplers alias / alias-fold (same thresholds as walk_conv.cuh).
//
// The loop is warp-convergent like walk_conv.cuh (every stage is entered by the whole warp) with a refill stage (a lane whose
// walker left or finished takes the next inbox tuple: lanes never idle to the end of the warp's longest walk) and a send stage --
// but a pass over the code advances a lane by a WHOLE trial (arrival, draw, neighbour entry, filter word as dependent loads),
// because this kernel is bound by instruction issue, not by memory requests (profiles/README.md).  Compiled for the host by
// tests/emu (warp_emu.h) to check the logic against the CPU twin.
#pragma once
#include <stdint.h>

#include "layout.h"
#include "philox.cuh"
#include "walk_conv.cuh"

enum : uint32_t { MIG_SETTLED = 0, MIG_PENDING = 1, MIG_NOP = 2, MIG_KIND_MASK = 3, MIG_NEEDEXT = 4, MIG_FWD = 8 /* spilled: route again */,
                  MIG_POWN_SHIFT = 4 /* bits 4-7: owner(prev) */, MIG_M_SHIFT = 8 /* bits 8-31: parallel edges curr-prev */ };
enum : int { MS_EMPTY = 0, MS_LOAD, MS_EXTENT, MS_TRIAL, MS_EXACT };

constexpr int kMigChunk = 64;        // inbox slots a warp claims at a time per destination (a multiple of the 32-slot block of mig_word)
constexpr int kMigStage = 16;        // tuples a warp collects in shared memory per destination before ONE coalesced flush over NVLink
constexpr int kMigClaim = 128;       // inbox items a warp claims at a time
constexpr int kMigMaxDest = SRW_MAX_SHARDS + 1;   // peers + the local spill region
constexpr uint32_t kMigRowMask = 0x0FFFFFFFu;     // MigTuple::home_row: [31:28] home shard, [27:0] path row on it
constexpr uint32_t kMigHub = 0xFFu;               // owner byte of a REPLICATED row (table-mapped shards): the row is wherever the walker is

struct MigArgs {
  // Table-mapped shards (template flag VCUT; all NULL for plain vertex ranges).  (1) The VCut shard map: owner(v) =
  // getPartition(v) mod world comes from the partition-id column of the edge file instead of `bounds`, so a shard's rows are not
  // a contiguous range of ranks.  (2) Replicated hub rows: the rows of the highest-degree vertices are on EVERY shard (owner
  // kMigHub), so a step onto a hub does not migrate -- under the degree-biased walk a row is visited in proportion to its length,
  // so replicating the rows that hold a fraction f of the entries keeps a fraction f of the steps local on top of the 1 / world.
  const MigExt *__restrict__ ext;         // [nv] row extent of every vertex inside its owner's arrays (hub rows: the same on every shard)
  const uint8_t *__restrict__ owner;      // [nv] owner(v), kMigHub for a replicated row
  const int32_t *__restrict__ lverts;     // [rows_local] the vertices this shard starts walkers for, ascending (seed order)
  int64_t rows_local;
  // this shard's rows
  const int64_t *__restrict__ off;        // [rows + 1] shard-local offsets
  const NbrEntry *__restrict__ ent;       // [nnz_local]
  const int32_t *__restrict__ hash;       // per-row hash sets of neighbour RANKS, placement derived from (off, deg)
  const unsigned long long *__restrict__ bloom;   // replicated edge filter
  uint32_t bloom_words;
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
  const unsigned long long *__restrict__ in_cnt;   // [world + 1] slots used per region (published by the senders)
  int64_t seg_cap, spill_cap;             // slots per peer region / in the spill region (world * seg_cap + spill_cap < 2^32)
  int64_t n_seed;                         // virtual seeds of THIS super-step: seed j is walker number seed_first + j * seed_step of the
  int64_t seed_first, seed_step;          // rows_local * n_rounds walkers this shard starts (super-steps 0 and 1 take every other one:
                                          // a shard's whole population leaves in one super-step, so on two shards an uneven start would
                                          // slosh back and forth for the whole walk; staggering the seeds damps that mode at once)
  // destinations: region `rank` of every peer's NEXT inbox (index world = own spill region)
  int4 *out_base[kMigMaxDest];
  unsigned long long *out_cnt_pub[kMigMaxDest];   // where the slot count of that region is published (peer memory)
  int32_t *home_paths[SRW_MAX_SHARDS];    // path matrix of every home GPU: [n_rounds * home_rows[h]][stride]
  int64_t home_rows[SRW_MAX_SHARDS];      // vertices v with v mod world == h
  // local scratch (device memory of this GPU)
  unsigned long long *cursor;             // inbox work cursor (low 32 bits used)
  unsigned long long *out_cnt;            // [world + 1] slots claimed per destination region
  unsigned long long *done_warps;
  int debug;                              // measurement switches (SRW_MIG_DEBUG): 1 = drop the path stores, 2 = path stores go to the local GPU
  unsigned long long *stats;              // [0] slots sent this super-step (written by the last warp), [1] steps, [2] proposals, [3] tests, [4] exact tests, [5] spills, [6] error flags, [7] exact tests that found the edge
};

struct MigTuple {            // 48 bytes: three 16-byte words
  uint32_t walker;           // batch-local
  int32_t prev, curr;
  uint32_t off, deg;         // row extent of curr inside owner(curr)'s arrays (invalid when MIG_NEEDEXT)
  uint32_t m_kind;           // [31:8] parallel edges curr-prev, [7:4] owner(prev), [3:0] kind | MIG_NEEDEXT | MIG_FWD
  uint32_t trial;
  uint32_t len;              // ids already in the path
  int32_t carry[3];          // decided path entries not yet stored: positions len - n .. len - 1, n = mig_carried(phase, len)
  uint32_t home_row;         // [31:28] home shard of the walker, [27:0] its path row there (no division on the hot path)
};
// A MIG_PENDING tuple (a proposal x whose adjacency to prev is verified at owner(x)) reuses two words: `off` holds x and `deg`
// holds [31:8] parallel edges curr-x, [7:4] owner(curr), [3:0] owner(x).  Row extents are not carried: owner(x) reads x's from its
// own row table before the test, and a rejected walker goes back to owner(curr) with MIG_NEEDEXT.

// number of path entries a walker with `len` ids holds back (positions >= 1 only: position 0 is written by the home GPU);
// phase = (row * stride) & 3: chunk boundaries are the multiples of 4 of the GLOBAL int index row * stride + pos
__device__ __forceinline__ uint32_t mig_carried(uint32_t phase, uint32_t len) {
  if (len <= 1) return 0;
  const uint32_t in_chunk = ((phase + len - 1u) & 3u) + 1u;        // entries of the chunk that position len - 1 belongs to, up to it
  const uint32_t n = in_chunk == 4u ? 0u : in_chunk;               // a complete chunk was stored when its fourth entry arrived
  return n < len - 1u ? n : len - 1u;
}

// Inbox slots are grouped in blocks of 32 (= the chunks senders claim); inside a block the three 16-byte words of the 32 tuples
// are stored word by word (word k of slot j at block * 96 + k * 32 + j, in int4 units).  Lanes that send together hold
// consecutive slots, so each of their three store instructions writes ONE contiguous run of 16-byte words -- NVLink carries
// a few large write packets instead of a 16-byte packet per lane and word -- and lanes that refill together read the same way.
__device__ __forceinline__ uint64_t mig_word(uint32_t slot, uint32_t k) { return (uint64_t)(slot >> 5) * 96u + k * 32u + (slot & 31u); }

// A warp stages tuples for every destination but its own rank: world - 1 stages (at 8 shards the eighth would push four resident
// blocks from the 100 KB shared-memory configuration into the 132 KB one, i.e. cost 32 KB of L1 -- and L1 capacity is what holds the
// kernel's outstanding gathers, profiles/README.md)
__device__ __forceinline__ int mig_sidx(int dest, int me) { return dest - (dest > me ? 1 : 0); }
__device__ __forceinline__ uint32_t mig_here(uint32_t owner, uint32_t me) { return owner == kMigHub ? me : owner; }
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
static inline unsigned mig_atomic_add32(unsigned long long *p, unsigned v) { return __atomic_fetch_add(reinterpret_cast<unsigned *>(p), v, __ATOMIC_RELAXED); }
#else
__device__ __forceinline__ unsigned mig_reduce_or(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
__device__ __forceinline__ unsigned mig_atomic_add32(unsigned long long *p, unsigned v) { return atomicAdd(reinterpret_cast<unsigned *>(p), v); }
#endif

// The whole warp copies the n staged tuples of destination d (stage: [3 words][kMigStage], word by word) into the open chunk of d's
// inbox region: per word ONE run of n contiguous 16-byte stores (mig_word keeps a stage inside one 32-slot block).  A new chunk is
// claimed from the region's counter when the open one is used up; a region that is full diverts the stage to the local spill region
// with MIG_FWD set (routed again in the next super-step).
template <int STAGE, class Args>
__device__ __forceinline__ void mig_flush(const Args &a, int d, int n, int4 *stage, unsigned int *chunk, unsigned int *fill, int lane,
                                           unsigned int &n_spill, unsigned int &n_err) {
  const int W = a.world;
  int dest = d;
  uint32_t fwd_flag = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    unsigned f = fill[dest];
    unsigned cb = chunk[dest];
    __syncwarp();
    bool ok = true;
    if (f + (unsigned)STAGE > (unsigned)kMigChunk) {              // no room in the open chunk: claim the next one
      // (the tail of the old chunk, if any, is padded at the end of the kernel only when it is the last one: a chunk is used in
      // whole stages of STAGE slots, and kMigChunk is a multiple of STAGE, so a used-up chunk has no tail)
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.out_cnt + dest, (unsigned long long)kMigChunk);
      base = __shfl_sync(0xffffffffu, base, 0);
      const unsigned long long cap = dest == W ? (unsigned long long)a.spill_cap : (unsigned long long)a.seg_cap;
      if (base + kMigChunk > cap) {
        if (lane == 0) atomicAdd(a.out_cnt + dest, (unsigned long long)(0ull - (unsigned long long)kMigChunk));
        ok = false;
      } else {
        cb = (unsigned int)base; f = 0;
        if (lane == 0) chunk[dest] = cb;
      }
    }
    if (ok) {
      int4 *out = a.out_base[dest];
      const int4 *sp = stage + mig_sidx(d, a.rank) * (3 * STAGE);
      for (int e = lane; e < 3 * STAGE; e += 32) {
        const int k = e / STAGE, j = e % STAGE;
        if (j < n) {
          int4 v = sp[e];
          if (k == 1) v.y |= (int)fwd_flag;
          out[mig_word(cb + f + (unsigned)j, (uint32_t)k)] = v;
        } else if (k == 1) {
          out[mig_word(cb + f + (unsigned)j, 1)] = make_int4(0, (int)MIG_NOP, 0, 0);     // a partial stage (end of the kernel): the rest of its slots
        }
      }
      if (lane == 0) fill[dest] = f + (unsigned)STAGE;
      __syncwarp();
      return;
    }
    if (dest == W) { if (lane == 0) n_err |= 2; return; }       // the spill region is sized for every walker of the batch: cannot happen
    if (lane == 0) n_spill += (unsigned)n;
    dest = W; fwd_flag = MIG_FWD;                                  // region full: park the stage locally
  }
}

template <bool STATS, int MINB = 4, bool VCUT = false, int STAGE = kMigStage, int FLAGS = 0>
__global__ void __launch_bounds__(256, MINB) mig_step_kernel(const MigArgs a) {
  static_assert(kMigChunk % STAGE == 0 && STAGE <= 32, "a chunk is used in whole stages");
  constexpr int GV = (FLAGS & 2) ? 3 : 1;       // random gathers: L2::64B; FLAGS bit 1: and no L1 allocation (measurement knob)
  // per-warp send state: open chunk (first slot, slots used) per destination region; prefix of the inbox regions (+ seeds)
  __shared__ unsigned int s_chunk[8][kMigMaxDest];     // open chunk of the destination region: first slot ...
  __shared__ unsigned int s_fill[8][kMigMaxDest];      // ... and slots of it already written (kMigChunk = none open)
  __shared__ unsigned int s_used[8][kMigMaxDest];      // tuples in the stage
#ifdef SRW_EMU
  static int4 mig_dyn[8 * kMigMaxDest * STAGE * 3];
#else
  extern __shared__ int4 mig_dyn[];                    // [8 warps][world - 1][3 words][STAGE] staged tuples
#endif
  __shared__ unsigned int s_pre[8][kMigMaxDest + 2];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int W = a.world, me = a.rank;
  unsigned int *chunk = s_chunk[wib];
  unsigned int *used = s_used[wib];
  unsigned int *fill = s_fill[wib];
  unsigned int *pre = s_pre[wib];
  int4 *stage = mig_dyn + (size_t)wib * (W - 1) * (3 * STAGE);
  if (lane <= W) { chunk[lane] = 0; fill[lane] = kMigChunk; used[lane] = 0; }
  if (lane == 0) {
    unsigned int acc = 0;
    for (int r = 0; r <= W; ++r) { pre[r] = acc; acc += a.in_cnt ? (unsigned int)a.in_cnt[r] : 0u; }
    pre[W + 1] = acc;
    pre[W + 2] = acc + (unsigned int)a.n_seed;
  }
  __syncwarp();
  const unsigned int total = pre[W + 2], seed0 = pre[W + 1];
  const bool acc_member = a.t_common > a.t_far, acc_non = a.t_far > a.t_common;   // verdict of a test that HAD to run (t_lo <= y < t_hi)
  const uint64_t t_lo = a.t_common < a.t_far ? a.t_common : a.t_far;
  const uint64_t t_hi = a.t_common < a.t_far ? a.t_far : a.t_common;
  const uint32_t stride = (uint32_t)a.stride;

  // lane state: one walker
  uint32_t walker = 0, off = 0, deg = 0, m = 1, trial = 0, len = 0;
  int32_t prev = -1, curr = 0, x = 0;
  uint32_t xoff = 0, xdeg = 0, xm = 1, xown = 0, cown = 0, pown = 0, k = 0, y = 0, bkt = 0, pnb = 0, lo = 0, hi = 0;
  uint32_t hrow = 0;                      // home shard | path row
  int32_t c0 = 0, c1 = 0, c2 = 0;         // carried path entries, oldest first
  bool fwd = false;
  uint32_t item = 0;
  int st = MS_EMPTY;
  uint32_t w_next = 0, w_end = 0, w_seg = 0;
  bool exhausted = total == 0;
  unsigned long long n_steps = 0, n_prop = 0, n_test = 0, n_exact = 0, n_hit = 0;
  unsigned int n_spill = 0, n_err = 0;

  for (;;) {
    // ---- R: refill empty lanes from the inbox ----
    const unsigned em = __ballot_sync(0xffffffffu, st == MS_EMPTY);
    if (em && !exhausted) {
      if (w_next >= w_end) {
        unsigned int base = 0;
        if (lane == 0) base = mig_atomic_add32(a.cursor, (unsigned)kMigClaim);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= total) { exhausted = true; w_next = w_end = 0; }
        else {
          w_next = base; w_end = total - base > (unsigned)kMigClaim ? base + kMigClaim : total;
          w_seg = 0;
          while ((int)w_seg <= W && base >= pre[w_seg + 1]) w_seg++;       // region of the first claimed item (W + 1 = seeds)
#ifndef SRW_EMU
          if (FLAGS & 1) {
            // The lanes consume these items over the next passes, one dependent 48-byte load each (6-9 % of the kernel's stall samples
            // sat on its first use): pull the claim's inbox lines into L2 now -- every lane asks for the three words of four items.
            for (unsigned int it = base + (unsigned)lane; it < w_end && it < seed0; it += 32) {
              uint32_t r = w_seg;
              while ((int)r < W && it >= pre[r + 1]) r++;
              const uint32_t slot = r * (uint32_t)a.seg_cap + (it - pre[r]);
              asm volatile("prefetch.global.L2 [%0];" ::"l"(a.in_base + mig_word(slot, 0)));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(a.in_base + mig_word(slot, 1)));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(a.in_base + mig_word(slot, 2)));
            }
          }
#endif
        }
      }
      const unsigned int mine = w_next + (unsigned)__popc(em & lt);
      if (st == MS_EMPTY && mine < w_end) {
        if (mine >= seed0) {                             // a virtual seed: walker (round, row) of this shard, path = [v]
          const unsigned long long j = (unsigned long long)a.seed_first + (unsigned long long)(mine - seed0) * (unsigned long long)a.seed_step;
          const int64_t rows = VCUT ? a.rows_local : a.row_last - a.row_first;
          const int64_t round = (int64_t)(j / (unsigned long long)rows), row = (int64_t)(j % (unsigned long long)rows);
          curr = VCUT ? __ldg(a.lverts + row) : (int32_t)(a.row_first + row); prev = -1;
          walker = (uint32_t)((unsigned long long)round * (unsigned long long)a.nv + (unsigned long long)curr);
          m = 1; trial = 0; len = 1; cown = (uint32_t)me; pown = 0;
          const uint32_t h = (uint32_t)curr % (uint32_t)W;
          const uint32_t prow = (uint32_t)(round * a.home_rows[h]) + (uint32_t)curr / (uint32_t)W;
          hrow = (h << 28) | prow;
          st = MS_EXTENT;
        } else {
          uint32_t r = w_seg;
          while ((int)r < W && mine >= pre[r + 1]) r++;
          item = r * (uint32_t)a.seg_cap + (mine - pre[r]);              // the slot
          st = MS_LOAD;
        }
      }
      const unsigned int adv = w_next + (unsigned)__popc(em);
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
    // The kernel is bound by instruction issue, not by memory (ncu: 60 % of the issue slots busy, 12 of 32 lanes active per
    // instruction, DRAM 27 % busy), so an iteration is FAT: a lane runs a whole trial -- [arrive: tuple load] -> draw -> neighbour
    // entry -> [filter word] -- as dependent loads inside ONE pass over the code, instead of one access per pass.  Every stage is
    // entered by the whole warp (or skipped by the whole warp when no lane needs it).
    // ---- S0: arrivals (tuple -> lane state), then row extents for seeds / returns that did not carry one ----
    if (__any_sync(0xffffffffu, st == MS_LOAD)) {
      bool pend = false;
      if (st == MS_LOAD) {
        const int4 q0 = gather16<0>(a.in_base + mig_word(item, 0)), q1 = gather16<0>(a.in_base + mig_word(item, 1)), q2 = gather16<0>(a.in_base + mig_word(item, 2));
        walker = (uint32_t)q0.x; prev = q0.y; curr = q0.z; off = (uint32_t)q0.w;
        deg = (uint32_t)q1.x; m = (uint32_t)q1.y >> MIG_M_SHIFT; trial = (uint32_t)q1.z; len = (uint32_t)q1.w;
        pown = ((uint32_t)q1.y >> MIG_POWN_SHIFT) & 15u;
        c0 = q2.x; c1 = q2.y; c2 = q2.z; hrow = (uint32_t)q2.w;
        const uint32_t kind = (uint32_t)q1.y & MIG_KIND_MASK;
        fwd = ((uint32_t)q1.y & MIG_FWD) != 0;
        cown = (uint32_t)me;
        if (kind == MIG_NOP) st = MS_EMPTY;
        else if (kind == MIG_PENDING) pend = true;
        else if (fwd && (cown = VCUT ? mig_here((uint32_t)__ldg(a.owner + curr), (uint32_t)me) : (uint32_t)mig_owner(a, curr)) != (uint32_t)me) {           // spilled last super-step: forward as it is
          send = (int)cown; send_kind = (uint32_t)q1.y & (MIG_KIND_MASK | MIG_NEEDEXT);
        } else st = ((uint32_t)q1.y & MIG_NEEDEXT) ? MS_EXTENT : MS_TRIAL;
      }
      if (__any_sync(0xffffffffu, pend)) {
        if (pend) {                                                        // a proposal under test arrives: t in N(x)? in x's own row
          x = (int32_t)off; xm = deg >> 8; cown = (deg >> 4) & 15u; xown = deg & 15u;       // (see MIG_PENDING: `off` / `deg` carry x and its tags)
          if (fwd && (int)xown != me) { send = (int)xown; send_kind = MIG_PENDING; }        // spilled: forward
          else {
            if (STATS) n_exact++;
            if (VCUT) { const MigExt e = a.ext[x]; xoff = e.off; xdeg = e.deg; }
            else {
              const int64_t *o = a.off + ((int64_t)x - a.row_first);       // x's row extent from this shard's own row table
              const int64_t e0 = __ldg(o), e1 = __ldg(o + 1);
              xoff = (uint32_t)e0; xdeg = (uint32_t)(e1 - e0);
            }
            pnb = srw_hash_buckets((int64_t)xoff, xdeg);
            if (pnb) bkt = __umulhi(srw_hash32((uint32_t)prev), pnb); else { lo = 0; hi = xdeg; }
            st = MS_EXACT;
          }
        }
      }
    }
    if (__any_sync(0xffffffffu, st == MS_EXTENT)) {
      if (st == MS_EXTENT) {
        if (VCUT) { const MigExt e = a.ext[curr]; off = e.off; deg = e.deg; }
        else {
          const int64_t *o = a.off + ((int64_t)curr - a.row_first);
          const int64_t e0 = __ldg(o), e1 = __ldg(o + 1);
          off = (uint32_t)e0; deg = (uint32_t)(e1 - e0);
        }
        if (deg == 0) { n_err |= 1; st = MS_EMPTY; }      // cannot happen on an undirected graph (every vertex has an entry)
        else st = MS_TRIAL;
      }
    }
    // ---- A: draw ----
    bool prop = false;
    if (st == MS_TRIAL && send < 0) {
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
        const uint32_t w = cown;
        needext = true;                                  // the row extent of prev is re-read at its owner (returns are rare: ~1/deg)
        cown = pown; pown = w;                           // m unchanged: the same bundle of parallel edges
      } else {
        k = (uint32_t)__umul64hi(((uint64_t)r.x << 32) | (uint64_t)r.w, (uint64_t)deg);
        y = r.z;
        prop = true;
      }
    }
    __syncwarp();
    // ---- B1 / C1: the neighbour entry of the proposal; a probe of an exact test under way ----
    int verdict = 0;               // 1 = accept x, 2 = reject (next trial)
    int member = -1;
    bool need_test = false;
    uint32_t bword = 0;
    uint64_t bmask = 0;
    if (prop) {
      const int4 q0 = gather16<GV>(reinterpret_cast<const int4 *>(a.ent + ((uint64_t)off + k)));
      x = q0.x; xdeg = (uint32_t)q0.y; xoff = (uint32_t)q0.z;
      xown = (uint32_t)q0.w & 0xFFu; xm = (uint32_t)q0.w >> 8;
      if (VCUT) xown = mig_here(xown, (uint32_t)me);                     // a replicated hub row: the walker stays where it is
      if (STATS && len > 1) n_prop++;
      if (len == 1 || deg == 1) verdict = 1;                             // first-order step (RW:57) / single choice
      else if (x == prev) verdict = ((uint64_t)y < a.t_ret) ? 1 : 2;     // RS:36
      else if ((uint64_t)y < t_lo) verdict = 1;
      else if ((uint64_t)y >= t_hi) verdict = 2;
      else {                                                             // RS:38 needs d(prev, x): ask the replicated filter first
        if (STATS) n_test++;
        srw_bloom_probe(prev, x, a.bloom_words, &bword, &bmask);
        need_test = true;
      }
    } else if (st == MS_EXACT && send < 0) {
      if (pnb) {
        int4 q0, q1;
        gather32<GV>(reinterpret_cast<const int4 *>(a.hash + ((uint64_t)(xoff >> 2) + bkt) * 8), q0, q1);
        const int32_t t = prev;
        const bool found = q0.x == t || q0.y == t || q0.z == t || q0.w == t || q1.x == t || q1.y == t || q1.z == t || q1.w == t;
        if (found) member = 1;
        else if (q1.w == -1) member = 0;
        else bkt = bkt + 1 == pnb ? 0 : bkt + 1;
      } else {
        const uint32_t mid = (lo + hi) >> 1;
        const int4 q0 = gather16<GV>(reinterpret_cast<const int4 *>(a.ent + ((uint64_t)xoff + mid)));
        if (q0.x == prev) member = 1;
        else {
          if (q0.x < prev) lo = mid + 1; else hi = mid;
          if (lo >= hi) member = 0;
        }
      }
    }
    // ---- B2 / C2: the filter word ----
    if (__any_sync(0xffffffffu, need_test)) {
      if (need_test) {
        unsigned long long bw;
#ifdef SRW_EMU
        bw = a.bloom[bword];
#else
        if (FLAGS & 2) asm("ld.global.nc.L1::no_allocate.L2::64B.u64 %0, [%1];" : "=l"(bw) : "l"(a.bloom + bword));
        else asm("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(bw) : "l"(a.bloom + bword));     // a missing probe fills 64 bytes, not a 128-byte line
#endif
        if ((bw & bmask) != bmask) member = 0;                             // definitely not adjacent
        else if ((int)xown != me) { send = (int)xown; send_kind = MIG_PENDING; }   // verify where the walker would go anyway
        else {                                                             // exact test in x's row, here, from the next pass on
          if (STATS) n_exact++;
          pnb = srw_hash_buckets((int64_t)xoff, xdeg);
          if (pnb) bkt = __umulhi(srw_hash32((uint32_t)prev), pnb); else { lo = 0; hi = xdeg; }
          st = MS_EXACT;
        }
      }
    }
    if (STATS && member == 1) n_hit++;
    if (member >= 0) verdict = (member ? acc_member : acc_non) ? 1 : 2;
    if (verdict == 1) {                                    // move along entry (x, xoff, xdeg, xm, xown)
      newv = x; moved = true;
      prev = curr; pown = cown;
      curr = x; off = xoff; deg = xdeg; m = xm; cown = xown;
    } else if (verdict == 2) {
      trial++;
      if ((int)cown == me) st = MS_TRIAL;
      else { send = (int)cown; send_kind = MIG_SETTLED | MIG_NEEDEXT; }  // the test ran at owner(x): back to the row of curr, whose extent is re-read there
    }
    if (moved) {                                           // RW:114: the step is decided -> the walker's home path row, four entries at a time
      const uint32_t phase = ((hrow & kMigRowMask) * stride) & 3u;
      const uint32_t pos = len, in_chunk = (phase + pos) & 3u;          // position of newv; its place in its 16-byte chunk
      const uint32_t have = mig_carried(phase, len);                    // entries carried so far (all of this chunk)
      if (in_chunk == 3u || pos + 1u == stride) {
        // the chunk is complete (or the path ends): store carried + newv, positions pos - have .. pos
        int32_t *dst = a.home_paths[a.debug == 2 ? (uint32_t)me : hrow >> 28] + ((uint64_t)(hrow & kMigRowMask) * stride + (pos - have));
        if (a.debug == 1) {
        } else if (have == 3u && in_chunk == 3u) *reinterpret_cast<int4 *>(dst) = make_int4(c0, c1, c2, newv);
        else if (have == 0u) dst[0] = newv;
        else if (have == 1u) { dst[0] = c0; dst[1] = newv; }
        else if (have == 2u) { dst[0] = c0; dst[1] = c1; dst[2] = newv; }
        else { dst[0] = c0; dst[1] = c1; dst[2] = c2; dst[3] = newv; }
      } else if (have == 0u) c0 = newv;
      else if (have == 1u) c1 = newv;
      else c2 = newv;
      len++; trial = 0;
      if (STATS) n_steps++;
      if (len == stride) st = MS_EMPTY;                    // RW:103,132
      else if ((int)cown == me) st = needext ? MS_EXTENT : MS_TRIAL;
      else { send = (int)cown; send_kind = MIG_SETTLED | (needext ? MIG_NEEDEXT : 0u); }
    }
    // ---- D: sends ----
    // A departing walker is written into the warp's shared-memory stage of its destination (one shared-memory atomic per lane);
    // a stage that fills up (STAGE tuples) is flushed by the whole warp into the destination's inbox region.
    if (__any_sync(0xffffffffu, send >= 0)) {
      const uint32_t w_off = (send_kind & MIG_KIND_MASK) == MIG_PENDING ? (uint32_t)x : off;
      const uint32_t w_deg = (send_kind & MIG_KIND_MASK) == MIG_PENDING ? ((xm << 8) | ((cown & 15u) << 4) | (xown & 15u)) : deg;
      const int4 t0 = make_int4((int)walker, prev, curr, (int)w_off);
      const int4 t1 = make_int4((int)w_deg, (int)((m << MIG_M_SHIFT) | ((pown & 15u) << MIG_POWN_SHIFT) | send_kind), (int)trial, (int)len);
      const int4 t2 = make_int4(c0, c1, c2, (int)hrow);
      unsigned pos = (unsigned)STAGE;
      if (send >= 0) {
        pos = atomicAdd(&used[send], 1u);
        if (pos < (unsigned)STAGE) { int4 *sp = stage + mig_sidx(send, me) * (3 * STAGE) + pos; sp[0] = t0; sp[STAGE] = t1; sp[2 * STAGE] = t2; }
      }
      __syncwarp();
      unsigned ovf = __ballot_sync(0xffffffffu, send >= 0 && pos >= (unsigned)STAGE);
      while (ovf) {                                              // the stage of destination d is full: flush it, then retry
        const int d = __shfl_sync(0xffffffffu, send, __ffs(ovf) - 1);
        mig_flush<STAGE>(a, d, STAGE, stage, chunk, fill, lane, n_spill, n_err);
        if (lane == 0) used[d] = 0;
        __syncwarp();
        if (send == d && pos >= (unsigned)STAGE) {
          pos = atomicAdd(&used[d], 1u);
          if (pos < (unsigned)STAGE) { int4 *sp = stage + mig_sidx(d, me) * (3 * STAGE) + pos; sp[0] = t0; sp[STAGE] = t1; sp[2 * STAGE] = t2; }
        }
        __syncwarp();
        ovf = __ballot_sync(0xffffffffu, send >= 0 && pos >= (unsigned)STAGE);
      }
      if (send >= 0) st = MS_EMPTY;
    }
  }
  // flush what is staged, pad the open chunks with NOPs, then hand the counts over
  for (int d = 0; d < W; ++d) {
    const unsigned n = used[d];
    __syncwarp();
    if (n) mig_flush<STAGE>(a, d, (int)n, stage, chunk, fill, lane, n_spill, n_err);
  }
  __syncwarp();
  for (int d = 0; d <= W; ++d) {
    const unsigned f = fill[d];          // slots of the open chunk in use (kMigChunk: no open chunk)
    for (unsigned j = f + (unsigned)lane; j < (unsigned)kMigChunk; j += 32) a.out_base[d][mig_word(chunk[d] + j, 1)] = make_int4(0, (int)MIG_NOP, 0, 0);
  }
  if (STATS) {
    for (int o = 16; o > 0; o >>= 1) {
      n_steps += __shfl_down_sync(0xffffffffu, n_steps, o); n_prop += __shfl_down_sync(0xffffffffu, n_prop, o);
      n_test += __shfl_down_sync(0xffffffffu, n_test, o); n_exact += __shfl_down_sync(0xffffffffu, n_exact, o);
      n_hit += __shfl_down_sync(0xffffffffu, n_hit, o);
    }
    if (lane == 0) {
      if (n_steps) atomicAdd(a.stats + 1, n_steps);
      if (n_prop) atomicAdd(a.stats + 2, n_prop);
      if (n_test) atomicAdd(a.stats + 3, n_test);
      if (n_exact) atomicAdd(a.stats + 4, n_exact);
      if (n_hit) atomicAdd(a.stats + 7, n_hit);
    }
  }
  if (n_spill) atomicAdd(a.stats + 5, (unsigned long long)n_spill);
  if (n_err) atomicAdd(a.stats + 6, (unsigned long long)n_err);
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
