// exact_generations.cuh -- MUSEUM, not product code: the exact (bit-parity) sampler kernels that preceded
// walk_exact_cert2_kernel (csrc/walk_exact.cuh): one thread per walker, one warp per walker with in-order shuffled folds, and
// the first certified parallel search.  Built into the test-only libsrw_museum.so (walk_museum.cu); see alias_generations.cuh.
#pragma once

__global__ void __launch_bounds__(128) walk_exact_kernel(WalkArgs a) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= a.n_walkers) return;
  const uint64_t walker = a.walker_first + (uint64_t)i;
  int32_t curr = (int32_t)(walker % (uint64_t)a.nv);
  int32_t *path = a.paths + i * a.stride;
  path[0] = curr;
  int32_t len = 1;
  int64_t off = a.off[curr], deg = a.off[curr + 1] - off;
  if (deg > 0) {
    const float *w0 = a.w_app + off;
    int64_t k = cdf_pick(deg, draw_u(a, walker, 0u), [&](int64_t j) { return w0[j]; });   // RW:57
    int32_t prev = curr;
    int64_t poff = off, pdeg = deg;
    curr = a.col_app[off + k];
    path[len++] = curr;
    while (len != a.stride) {
      off = a.off[curr];
      deg = a.off[curr + 1] - off;
      if (deg <= 0) break;
      const int32_t *cd = a.col_app + off;
      const float *cw = a.w_app + off;
      const float u = draw_u(a, walker, (uint32_t)(len - 1));
      k = cdf_pick(deg, u, [&](int64_t j) {
        const int32_t d = cd[j];
        const bool need = (d != prev) && (a.p != 1.0f || a.q != 1.0f);
        return biased_weight(a.p, a.q, prev, d, cw[j], need ? row_contains(a.col, poff, pdeg, d) : false);
      });
      prev = curr; poff = off; pdeg = deg;
      curr = cd[k];
      path[len++] = curr;
    }
  }
  a.lens[i] = len;
}

__global__ void __launch_bounds__(256) walk_exact_warp_kernel(WalkArgs a, const int32_t *__restrict__ hash) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;     // one warp per walker
  if (i >= a.n_walkers) return;
  const uint64_t walker = a.walker_first + (uint64_t)i;
  int32_t curr = (int32_t)(walker % (uint64_t)a.nv), prev = -1;
  int32_t *path = a.paths + i * a.stride;
  if (lane == 0) path[0] = curr;
  int32_t len = 1;
  int64_t poff = 0;
  uint32_t pdeg = 0;
  const bool biased = a.p != 1.0f || a.q != 1.0f;
  while (len != a.stride) {                                                     // RW:103
    const int64_t off = __ldg(a.off + curr);
    const uint32_t deg = (uint32_t)(__ldg(a.off + curr + 1) - off);
    if (deg == 0) break;                                                        // RW:59-62 / RW:115-119
    const int32_t *cd = a.col_app + off;
    const float *cw = a.w_app + off;
    const float u = draw_u(a, walker, (uint32_t)(len - 1));
    const bool second = len > 1;
    const uint32_t pnb = (second && biased && hash) ? srw_hash_buckets(poff, pdeg) : 0u;
    auto weight_of = [&](uint32_t j) -> float {                                 // RS:33-41 for neighbour j (first step: RS:12 plain weights)
      const float w = __ldg(cw + j);
      if (!second) return w;
      const int32_t d = __ldg(cd + j);
      bool in_prev = false;
      if (biased && d != prev) in_prev = pnb ? hash_contains(hash, poff, pnb, d) : row_contains(a.col, poff, pdeg, d);
      return biased_weight(a.p, a.q, prev, d, w, in_prev);
    };
    // pass 1 (RS:14): sum, strictly left to right
    double sum = 0.0;
    for (uint32_t base = 0; base < deg; base += 32) {
      const uint32_t j = base + lane, n = min(32u, deg - base);
      const float wv = j < deg ? weight_of(j) : 0.0f;
      for (uint32_t l = 0; l < n; ++l) sum = __dadd_rn(sum, (double)__shfl_sync(0xffffffffu, wv, (int)l));
    }
    // pass 2 (RS:16-22): acc += w / sum; first index with acc >= u
    double acc = 0.0;
    int64_t pick = 0;                                                           // RS:24 edges.head
    bool found = false;
    for (uint32_t base = 0; base < deg && !found; base += 32) {
      const uint32_t j = base + lane, n = min(32u, deg - base);
      const double qv = j < deg ? __ddiv_rn((double)weight_of(j), sum) : 0.0;
      for (uint32_t l = 0; l < n; ++l) {
        acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, qv, (int)l));
        if (acc >= (double)u) { pick = base + l; found = true; break; }
      }
    }
    const int32_t nxt = __ldg(cd + pick);
    if (lane == 0) path[len] = nxt;                                             // RW:114
    len++;
    prev = curr; poff = off; pdeg = deg;
    curr = nxt;
  }
  if (lane == 0) a.lens[i] = len;
}

__global__ void __launch_bounds__(256) walk_exact_cert_kernel(WalkArgs a, const int32_t *__restrict__ hash) {
  const int lane = threadIdx.x & 31;
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
  while (len != a.stride) {                                                     // RW:103
    const int64_t off = __ldg(a.off + curr);
    const uint32_t deg = (uint32_t)(__ldg(a.off + curr + 1) - off);
    if (deg == 0) break;                                                        // RW:59-62 / RW:115-119
    const int32_t *cd = a.col_app + off;
    const float *cw = a.w_app + off;
    const float u = draw_u(a, walker, (uint32_t)(len - 1));
    const bool second = len > 1;
    const uint32_t pnb = (second && biased && hash) ? srw_hash_buckets(poff, pdeg) : 0u;
    auto weight_of = [&](uint32_t j) -> float {                                 // RS:33-41 for neighbour j (first step: RS:12 plain weights)
      const float w = __ldg(cw + j);
      if (!second) return w;
      const int32_t d = __ldg(cd + j);
      bool in_prev = false;
      if (biased && d != prev) in_prev = pnb ? hash_contains(hash, poff, pnb, d) : row_contains(a.col, poff, pdeg, d);
      return biased_weight(a.p, a.q, prev, d, w, in_prev);
    };
    int64_t pick = -1;
    // ---- certified parallel search ----
    {
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
        const double t_hi = (uu + delta) * part * (1.0 + 1e-15), t_lo = (uu - delta) * part;   // t_lo < 0: every prefix is above it
        double carry = 0.0;
        for (uint32_t base = 0; base < deg; base += 32) {
          const uint32_t j = base + lane;
          const double wv = j < deg ? (double)weight_of(j) : 0.0;
          const double P = __dadd_rn(carry, warp_scan_incl(wv, lane));
          const unsigned valid = (deg - base >= 32u) ? 0xffffffffu : ((1u << (deg - base)) - 1u);
          const unsigned hi = __ballot_sync(0xffffffffu, P >= t_hi) & valid;
          const unsigned band = __ballot_sync(0xffffffffu, P > t_lo) & valid;     // includes the hi lanes
          const unsigned below_first_hi = hi ? ((1u << (__ffs(hi) - 1)) - 1u) : 0xffffffffu;
          if (band & ~hi & below_first_hi) break;                                  // a prefix inside the band: replay in order
          if (hi) { pick = base + (__ffs(hi) - 1); break; }
          carry = __shfl_sync(0xffffffffu, P, 31);
          if (base + 32 >= deg) pick = 0;                                          // never reached u, certainly: RS:24 edges.head
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
  }
  if (lane == 0) {
    a.lens[i] = len;
    if (n_replay) atomicAdd(a.stats + 2, n_replay);                             // reported as member_tests: in-order replays
  }
}
