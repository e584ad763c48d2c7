// emu_exact.cpp -- TEST INFRASTRUCTURE: the product's warp-cooperative exact-sampler kernels (csrc/walk_exact.cuh) on the host
// under the lockstep warp emulator, over an appearance-order CSR + the sorted rows / hash sets handed in by the test
// (laid out as graph_build.cu does).  The test compares the paths with the oracle (the reference algorithm).
#include "warp_emu.h"

#include "../../include/srw.h"
#include "../../stellar-random-walk_b200/csrc/walk_conv.cuh"   // WalkArgs
#include "../../stellar-random-walk_b200/csrc/walk_exact.cuh"
#include "../../profiles/museum/exact_generations.cuh"   // kernels 0 and 1 below: superseded generations, kept as a cross-check

namespace {
void hash_insert(std::vector<int32_t> &hash, int64_t off, uint32_t deg, int32_t x) {
  const uint32_t nb = srw_hash_buckets(off, deg);
  if (!nb) return;
  uint32_t b = __umulhi(srw_hash32((uint32_t)x), nb);
  for (;;) {
    int32_t *bucket = hash.data() + (srw_hash_first(off) + b) * 8;
    for (int s = 0; s < 8; ++s) {
      if (bucket[s] == x) return;
      if (bucket[s] == -1) { bucket[s] = x; return; }
    }
    b = b + 1 == nb ? 0 : b + 1;
  }
}
}  // namespace

// kernel: 0 = walk_exact_warp_kernel (in-order chains), 1 = walk_exact_cert_kernel, 2 = walk_exact_cert2_kernel.
// off/col_app/w_app: appearance-order rows (ranks); col_sorted: the same rows sorted (membership).  u_const < 0: Philox draws.
extern "C" int emu_exact_walk(int kernel, int64_t nv, const int64_t *off, const int32_t *col_app, const float *w_app,
                              const int32_t *col_sorted, float p, float q, uint64_t seed, float u_const, int32_t walk_length,
                              uint64_t walker_first, int64_t n_walkers, int32_t *paths, int32_t *lens, int use_hash,
                              unsigned long long *stats_out) {
  const int64_t nnz = off[nv];
  std::vector<int32_t> hash((size_t)(((nnz >> 2) + 1) * 8), -1);
  for (int64_t r = 0; r < nv; ++r)
    for (int64_t e = off[r]; e < off[r + 1]; ++e) hash_insert(hash, off[r], (uint32_t)(off[r + 1] - off[r]), col_sorted[e]);
  WalkArgs a{};
  a.off = off; a.col = col_sorted; a.col_app = col_app; a.w_app = w_app; a.nv = nv;
  a.walker_first = walker_first; a.n_walkers = n_walkers; a.stride = walk_length + 2;
  a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  a.p = p; a.q = q;
  a.u_mode = u_const >= 0.0f ? SRW_U_CONST : SRW_U_PHILOX; a.u_const = u_const;
  a.paths = paths; a.lens = lens;
  unsigned long long st[4] = {0, 0, 0, 0};
  a.stats = st;
  const int32_t *h = use_hash ? hash.data() : nullptr;
  const int64_t n_warps = ((n_walkers + 7) / 8) * 8;      // whole blocks, as the launch does
  emu_launch_warps(n_warps, [&] {
    if (kernel == 0) walk_exact_warp_kernel(a, h);
    else if (kernel == 1) walk_exact_cert_kernel(a, h);
    else walk_exact_cert2_kernel(a, h);
  });
  if (stats_out) memcpy(stats_out, st, sizeof(st));
  return 0;
}
