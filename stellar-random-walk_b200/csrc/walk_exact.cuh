// walk_exact.cuh -- K5: the exact (bit-parity) sampler, RS:12-62 literally, in three kernels (one thread per walker, one warp
// per walker with in-order float64 chains, one warp per walker with a CERTIFIED parallel search), plus the membership helpers
// the older alias kernels share.  Included by walk.cu (inside its anonymous namespace) and, with SRW_EMU defined, by
// tests/emu/ where the warp kernels run on the host under a lockstep 32-lane warp emulator (tests/emu/warp_emu.h).
#pragma once
#include <stdint.h>

#include "layout.h"
#include "philox.cuh"

// x in sorted row [lo, lo+n)?
__device__ __forceinline__ bool row_contains(const int32_t *__restrict__ col, int64_t lo, int64_t n, int32_t x) {
  int64_t a = 0, b = n;
  while (a < b) {
    const int64_t m = (a + b) >> 1;
    if (__ldg(col + lo + m) < x) a = m + 1; else b = m;
  }
  return a < n && __ldg(col + lo + a) == x;
}

// ------------------------------------------------------------------------------------------
// K5: exact sampler (also used by the KAT entry points)
// ------------------------------------------------------------------------------------------
// RS:12-25 over weights produced by `wf(i)`: two passes, float64 accumulation, left to right.
template <class WF>
__device__ __forceinline__ int64_t cdf_pick(int64_t n, float u, WF wf) {
  double sum = 0.0;
  for (int64_t i = 0; i < n; ++i) sum = __dadd_rn(sum, (double)wf(i));     // RS:14
  double acc = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    acc = __dadd_rn(acc, __ddiv_rn((double)wf(i), sum));                   // RS:19
    if (acc >= (double)u) return i;                                        // RS:20
  }
  return 0;                                                                // RS:24 edges.head
}

// RS:33-41 for one neighbour
__device__ __forceinline__ float biased_weight(float p, float q, int32_t prev, int32_t dst, float w, bool in_prev_row) {
  float un = __fdiv_rn(w, q);
  if (dst == prev) un = __fdiv_rn(w, p);
  else if (in_prev_row) un = w;
  return un;
}

__device__ __forceinline__ float draw_u(const WalkArgs &a, uint64_t walker, uint32_t step) {
  if (a.u_mode == SRW_U_CONST) return a.u_const;
  return u01_from_bits(walker_rng(a.seed_lo, a.seed_hi, walker, step, 0u).x);
}

// (walk_exact_kernel, the one-thread-per-walker original, lives in profiles/museum/exact_generations.cuh)

// ------------------------------------------------------------------------------------------
// K5 (v2): the exact sampler with one WARP per walker.  RS:27-44 is embarrassingly parallel over the
// neighbours of curr (one membership test each), RS:14 / RS:19 are not: float64 addition does not
// associate, so the sum and the running CDF stay strictly left-to-right.  Each lane therefore computes the
// biased weight (and, in pass 2, the float64 quotient w'/sum) of one neighbour per 32-wide chunk -- a hash
// probe in prev's neighbour set, or a binary search in its sorted row when no set was built -- and the 32
// values are then folded in order through warp shuffles, every lane carrying the same accumulator.
// Bit-identical to walk_exact_kernel and to the oracle; O(d_c/32) chunk rounds per pass instead of O(d_c).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool hash_contains(const int32_t *__restrict__ hash, int64_t poff, uint32_t pnb, int32_t x) {
  uint32_t b = __umulhi(srw_hash32((uint32_t)x), pnb);
  for (;;) {
    const int4 *q = reinterpret_cast<const int4 *>(hash + (srw_hash_first(poff) + (int64_t)b) * 8);
    const int4 q0 = __ldg(q), q1 = __ldg(q + 1);
    if (q0.x == x || q0.y == x || q0.z == x || q0.w == x || q1.x == x || q1.y == x || q1.z == x || q1.w == x) return true;
    if (q1.w == -1) return false;
    b = b + 1 == pnb ? 0 : b + 1;
  }
}

// (walk_exact_warp_kernel: profiles/museum/exact_generations.cuh)

// ------------------------------------------------------------------------------------------
// K5 (v3): the exact sampler with a CERTIFIED parallel inverse-CDF search.  Still bit-identical to RS:12-25,
// but the two float64 chains of the reference (sum, then acc += w/sum) are not replayed element by element
// unless they have to be.  For non-negative weights, ANY summation order of k terms is within
// k * 2^-53 * (exact sum) of the exact sum (Higham, gamma_k), and fl(w/sum) is within 2^-53 relative of w/sum.
// Hence, with S = a parallel (tree) sum of the row and P_k = a parallel prefix sum,
//        | acc_k(reference, sequential) - P_k / S |  <=  (3k + 2) * 2^-53 * (1 + tiny)
// and the reference's answer "first k with acc_k >= u" is decided by comparing P_k with (u -+ delta) * S,
// delta = (4n + 64) * 2^-52, whenever no prefix falls inside the +-delta band around u.  The first prefix
// certainly above the band is then the reference's pick -- every earlier one is certainly below.  If some
// earlier prefix lands inside the band (probability ~ n * 2 * delta per step, < 2^-8 for a million-entry
// row), or a weight is negative / non-finite, or the sum is not a positive finite number, the step is
// replayed with the in-order chains of walk_exact_warp_kernel.  O(d_c / 32) warp scans per step instead of
// d_c dependent float64 additions.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_scan_incl(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = __dadd_rn(v, t);
  }
  return v;
}

// (walk_exact_cert_kernel: profiles/museum/exact_generations.cuh)


// ------------------------------------------------------------------------------------------
// K5 (v4): walk_exact_cert_kernel with two savings on long rows, same bits (the certification argument above
// holds for ANY summation order of the non-negative weights, and the reference's acc is monotone in k):
//  * COMMON LIST.  The bias rule needs "is d a neighbour of prev" for every neighbour d of curr (RS:38).  When curr's
//    row is much longer than prev's (the usual case at a hub: the walk alternates hub -> leaf -> hub), the few common
//    neighbours are found once per step from prev's side -- deg(prev) probes of curr's hash set -- and kept in shared
//    memory; the pass over curr's row then compares against that short list instead of probing a hash set per element.
//  * TWO LEVELS.  Rows of >= kTwoLevelMinDeg entries are summed as 32 contiguous groups; the group-end prefixes locate
//    the group of the pick (every earlier group end certainly below u, hence every prefix inside it), and only that
//    group is scanned: deg * (1 + 1/32) weight evaluations per step instead of deg * ~1.5.
// Anything uncertain (a prefix inside the +-delta band, odd weights) is replayed in order, as before.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kTwoLevelMinDeg = 2048;
constexpr int kCommonCap = 32;

__global__ void __launch_bounds__(256) walk_exact_cert2_kernel(WalkArgs a, const int32_t *__restrict__ hash) {
  __shared__ int32_t s_common[8][kCommonCap];
  const int lane = threadIdx.x & 31, wib = (threadIdx.x >> 5) & 7;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;     // one warp per walker
  if (i >= a.n_walkers) return;
  const uint64_t walker = a.walker_first + (uint64_t)i;
  int32_t curr = (int32_t)(walker % (uint64_t)a.nv), prev = -1;
  int32_t *path = a.paths + i * a.stride;
  if (lane == 0) path[0] = curr;
  int32_t len = 1;
  int64_t poff = 0;
  uint32_t pdeg = 0;
  unsigned long long n_replay = 0;
  const bool biased = a.p != 1.0f || a.q != 1.0f;
  int32_t *common = s_common[wib];
  while (len != a.stride) {                                                     // RW:103
    const int64_t off = __ldg(a.off + curr);
    const uint32_t deg = (uint32_t)(__ldg(a.off + curr + 1) - off);
    if (deg == 0) break;                                                        // RW:59-62 / RW:115-119
    const int32_t *cd = a.col_app + off;
    const float *cw = a.w_app + off;
    const float u = draw_u(a, walker, (uint32_t)(len - 1));
    const bool second = len > 1;
    const uint32_t pnb = (second && biased && hash) ? srw_hash_buckets(poff, pdeg) : 0u;
    // ---- common list: N(prev) /\ N(curr) from prev's side ----
    int n_common = -1;                                                          // -1: membership is probed per element
    if (second && biased && hash && deg >= 64u && pdeg <= 1024u && (uint64_t)deg >= 4ull * (uint64_t)pdeg) {
      const uint32_t cnb = srw_hash_buckets(off, deg);                          // deg >= 8: curr's row has a hash set
      int cnt = 0;
      for (uint32_t base = 0; base < pdeg; base += 32) {
        const uint32_t j = base + lane;
        bool m = false;
        int32_t y = -1;
        if (j < pdeg) {
          y = __ldg(a.col + poff + j);
          const bool dup = j > 0 && __ldg(a.col + poff + j - 1) == y;           // parallel edges: the sorted row repeats y
          m = !dup && hash_contains(hash, off, cnb, y);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, m);
        if (m) {
          const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
          if (pos < kCommonCap) common[pos] = y;
        }
        cnt += __popc(bal);
      }
      __syncwarp();
      if (cnt <= kCommonCap) n_common = cnt;
    }
    auto weight_of = [&](uint32_t j) -> float {                                 // RS:33-41 for neighbour j (first step: RS:12 plain weights)
      const float w = __ldg(cw + j);
      if (!second) return w;
      const int32_t d = __ldg(cd + j);
      bool in_prev = false;
      if (biased && d != prev) {
        if (n_common >= 0) { for (int c = 0; c < n_common; ++c) in_prev |= common[c] == d; }
        else in_prev = pnb ? hash_contains(hash, poff, pnb, d) : row_contains(a.col, poff, pdeg, d);
      }
      return biased_weight(a.p, a.q, prev, d, w, in_prev);
    };
    // scans [b0, b1) in chunks of 32 with `carry` = the prefix before b0: the pick, or -1 (a prefix inside the band, or
    // no prefix of the range certainly above it); `whole_row`: reaching the end certainly below u means RS:24 edges.head
    auto scan_range = [&](uint32_t b0, uint32_t b1, double carry, double t_lo, double t_hi, bool whole_row) -> int64_t {
      for (uint32_t base = b0; base < b1; base += 32) {
        const uint32_t j = base + lane;
        const double wv = j < b1 ? (double)weight_of(j) : 0.0;
        const double P = __dadd_rn(carry, warp_scan_incl(wv, lane));
        const unsigned valid = (b1 - base >= 32u) ? 0xffffffffu : ((1u << (b1 - base)) - 1u);
        const unsigned hi = __ballot_sync(0xffffffffu, P >= t_hi) & valid;
        const unsigned band = __ballot_sync(0xffffffffu, P > t_lo) & valid;     // includes the hi lanes
        const unsigned below_first_hi = hi ? ((1u << (__ffs(hi) - 1)) - 1u) : 0xffffffffu;
        if (band & ~hi & below_first_hi) return -1;                              // a prefix inside the band: replay in order
        if (hi) return (int64_t)base + (__ffs(hi) - 1);
        carry = __shfl_sync(0xffffffffu, P, 31);
        if (base + 32 >= b1 && whole_row) return 0;                              // never reached u, certainly: RS:24 edges.head
      }
      return -1;
    };
    int64_t pick = -1;
    // ---- certified parallel search ----
    if (deg < kTwoLevelMinDeg) {
      double part = 0.0;
      bool bad = false;
      for (uint32_t j = lane; j < deg; j += 32) {
        const float wv = weight_of(j);
        bad |= !(wv >= 0.0f) || !(wv <= 3.0e38f);
        part = __dadd_rn(part, (double)wv);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part = __dadd_rn(part, __shfl_xor_sync(0xffffffffu, part, o));
      bad = __any_sync(0xffffffffu, bad) || !(part > 0.0) || !(part <= 1.0e300);
      if (!bad) {
        const double delta = (4.0 * (double)deg + 64.0) * 2.220446049250313e-16;        // 2^-52
        const double uu = (double)u;
        pick = scan_range(0u, deg, 0.0, (uu - delta) * part, (uu + delta) * part * (1.0 + 1e-15), true);
      }
    } else {
      const uint32_t gs = ((deg + 1023u) / 1024u) * 32u;                        // 32 groups of gs entries (a multiple of 32) cover the row
      double gsum = 0.0;                                                         // lane g keeps the sum of group g
      bool bad = false;
      for (uint32_t g = 0; g < 32u; ++g) {
        const uint64_t b0 = (uint64_t)g * gs;
        const uint32_t b1 = (uint32_t)(b0 + gs < (uint64_t)deg ? b0 + gs : (uint64_t)deg);
        double part = 0.0;
        for (uint64_t j = b0 + (uint64_t)lane; j < (uint64_t)b1; j += 32) {
          const float wv = weight_of((uint32_t)j);
          bad |= !(wv >= 0.0f) || !(wv <= 3.0e38f);
          part = __dadd_rn(part, (double)wv);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part = __dadd_rn(part, __shfl_xor_sync(0xffffffffu, part, o));
        if (lane == (int)g) gsum = part;
      }
      const double Pg = warp_scan_incl(gsum, lane);                             // prefix at the end of every group
      const double S = __shfl_sync(0xffffffffu, Pg, 31);
      bad = __any_sync(0xffffffffu, bad) || !(S > 0.0) || !(S <= 1.0e300);
      if (!bad) {
        const double delta = (4.0 * (double)deg + 64.0) * 2.220446049250313e-16;
        const double uu = (double)u;
        const double t_hi = (uu + delta) * S * (1.0 + 1e-15), t_lo = (uu - delta) * S;
        const unsigned hi = __ballot_sync(0xffffffffu, Pg >= t_hi);
        const unsigned band = __ballot_sync(0xffffffffu, Pg > t_lo);
        if (hi) {
          const int g = __ffs(hi) - 1;                                           // first group whose end is certainly above u
          const double before = __shfl_sync(0xffffffffu, Pg, g > 0 ? g - 1 : 0);
          if (!(band & ~hi & ((1u << g) - 1u))) {                               // every earlier group end certainly below u
            const uint64_t b0 = (uint64_t)g * gs;
            const uint32_t b1 = (uint32_t)(b0 + gs < (uint64_t)deg ? b0 + gs : (uint64_t)deg);
            pick = scan_range((uint32_t)b0, b1, g > 0 ? before : 0.0, t_lo, t_hi, false);
          }
        }
      }
    }
    if (pick < 0) {
      // ---- in-order replay (RS:14, RS:16-22 literally) ----
      n_replay++;
      double sum = 0.0;
      for (uint32_t base = 0; base < deg; base += 32) {
        const uint32_t j = base + lane, n = min(32u, deg - base);
        const float wv = j < deg ? weight_of(j) : 0.0f;
        for (uint32_t l = 0; l < n; ++l) sum = __dadd_rn(sum, (double)__shfl_sync(0xffffffffu, wv, (int)l));
      }
      double acc = 0.0;
      pick = 0;                                                                 // RS:24 edges.head
      bool found = false;
      for (uint32_t base = 0; base < deg && !found; base += 32) {
        const uint32_t j = base + lane, n = min(32u, deg - base);
        const double qv = j < deg ? __ddiv_rn((double)weight_of(j), sum) : 0.0;
        for (uint32_t l = 0; l < n; ++l) {
          acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, qv, (int)l));
          if (acc >= (double)u) { pick = base + l; found = true; break; }
        }
      }
    }
    const int32_t nxt = __ldg(cd + pick);
    if (lane == 0) path[len] = nxt;                                             // RW:114
    len++;
    prev = curr; poff = off; pdeg = deg;
    curr = nxt;
    __syncwarp();                                                               // the common list is rewritten next step
  }
  if (lane == 0) {
    a.lens[i] = len;
    if (n_replay) atomicAdd(a.stats + 2, n_replay);                             // reported as member_tests: in-order replays
  }
}
