// emu_walk.cpp -- TEST INFRASTRUCTURE: runs the product's walk_fold_conv_kernel SOURCE on the host
// (tests/emu/cuda_emu.h) over a CSR handed in by the test, after laying the rows out exactly as
// graph_build.cu does (16-byte neighbour entries, derived-placement hash sets, optional vertex-range
// shards for the PEER variant).  The test compares the paths with the CPU twin (oracle_alias_walk).
// Two builds: the default one drives the kernel one lane at a time (cuda_emu.h: fast, a "warp" is one lane); with
// -DSRW_EMU_WARP the same kernels run under the lockstep 32-lane warp emulator (warp_emu.h), where lanes of a warp finish
// at different times and the __any_sync loop / dead-lane logic is exercised for real.
#ifdef SRW_EMU_WARP
#include "warp_emu.h"
#else
#include "cuda_emu.h"
#endif

#include <functional>
#include <vector>

#include "../../stellar-random-walk_b200/csrc/walk_conv.cuh"

namespace {

// runs `kernel` for every thread of a grid of 256-thread blocks covering n_walkers walkers
void run_grid(int64_t n_walkers, int extra, const std::function<void()> &kernel) {
  const int64_t n_blocks = (n_walkers + 255) / 256;
#ifdef SRW_EMU_WARP
  (void)extra;
  emu_launch_warps(n_blocks * 8, kernel);
#else
  emu_extra_iters = extra;
  blockDim.x = 256; blockDim.y = blockDim.z = 1;
  for (int64_t b = 0; b < n_blocks; ++b) {
    blockIdx.x = (unsigned)b;
    for (unsigned t = 0; t < 256; ++t) {
      threadIdx.x = t;
      emu_linger = 0;
      kernel();
    }
  }
#endif
}

struct ShardRows {
  std::vector<int64_t> off;        // shard-local offsets, rows+1
  std::vector<NbrEntry> ent;
  std::vector<int32_t> hash;       // 8-slot buckets, -1 = empty
};

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

// Lays the graph out for `shards` vertex ranges (bounds[shards+1], first rank of every range; shards == 1:
// the unsharded layout) and walks n_walkers walkers.  var: VAR template value; extra: see cuda_emu.h.
// fold != 0: alias-fold arguments from (p, q); fold == 0: classic thresholds through the same kernel.
extern "C" int emu_fold_walk_ids(int64_t nv, const int64_t *off, const int32_t *col, const uint32_t *mult, const int32_t *vids,
                                 double p, double q, uint64_t seed, int32_t walk_length, uint64_t walker_first, int64_t n_walkers,
                                 int32_t *paths, int32_t *lens, int var, int extra);

extern "C" int emu_fold_walk(int64_t nv, const int64_t *off, const int32_t *col, const uint32_t *mult, int shards,
                             const int64_t *bounds, double p, double q, int fold, uint64_t t_ret, uint64_t t_common,
                             uint64_t t_far, uint64_t seed, int32_t walk_length, uint64_t walker_first, int64_t n_walkers,
                             int32_t *paths, int32_t *lens, int var, int extra, int stats, unsigned long long *stats_out) {
  if (shards < 1 || shards > SRW_MAX_SHARDS) return -1;
  std::vector<ShardRows> sh((size_t)shards);
  std::vector<int64_t> base((size_t)shards);
  for (int s = 0; s < shards; ++s) {
    const int64_t r0 = bounds[s], r1 = bounds[s + 1];
    base[(size_t)s] = off[r0];
    ShardRows &R = sh[(size_t)s];
    R.off.resize((size_t)(r1 - r0 + 1));
    for (int64_t r = r0; r <= r1; ++r) R.off[(size_t)(r - r0)] = off[r] - off[r0];
    const int64_t n = off[r1] - off[r0];
    R.ent.resize((size_t)n);
    R.hash.assign((size_t)(((n >> 2) + 1) * 8), -1);
  }
  auto owner_of = [&](int64_t v) { int o = 0; while (o + 1 < shards && v >= bounds[o + 1]) o++; return o; };
  for (int s = 0; s < shards; ++s) {
    ShardRows &R = sh[(size_t)s];
    for (int64_t r = bounds[s]; r < bounds[s + 1]; ++r) {
      const int64_t lo = off[r] - base[(size_t)s];
      const uint32_t deg = (uint32_t)(off[r + 1] - off[r]);
      for (int64_t e = off[r]; e < off[r + 1]; ++e) {
        const int32_t x = col[e];
        const int ox = owner_of(x);
        NbrEntry ne;
        ne.x = x; ne.deg = (uint32_t)(off[x + 1] - off[x]); ne.off_lo = (uint32_t)(off[x] - base[(size_t)ox]);
        ne.off_hi_mult = (uint32_t)ox | (mult[e] << 8);
        R.ent[(size_t)(e - base[(size_t)s])] = ne;
        hash_insert(R.hash, lo, deg, x);
      }
    }
  }
  WalkArgs a{};
  a.off = sh[0].off.data(); a.nv = nv; a.walker_first = walker_first; a.n_walkers = n_walkers; a.stride = walk_length + 2;
  a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  a.paths = paths; a.lens = lens;
  unsigned long long st[4] = {0, 0, 0, 0};
  a.stats = st;
  FoldArgs f{};
  const bool folded = srw_fold_args(p, q, fold != 0, &f);
  if (!folded) { f.a = 0.0; f.mp = 1.0; f.t_ret = t_ret; f.t_common = t_common; f.t_far = t_far; }
  f.ent = sh[0].ent.data(); f.hash = sh[0].hash.data();
  PeerTable pt{};
  pt.world = shards;
  for (int s = 0; s <= shards; ++s) pt.first[s] = bounds[s];
  for (int s = 0; s < shards; ++s) { pt.off[s] = sh[(size_t)s].off.data(); pt.ent[s] = sh[(size_t)s].ent.data(); pt.hash[s] = sh[(size_t)s].hash.data(); }
  const bool peer = shards > 1;
  auto launch = [&](const WalkArgs &aa) {
    run_grid(n_walkers, extra, [&] {
      if (peer) {
        if (var & 1) walk_fold_conv_kernel<false, true, 1>(aa, f, pt);
        else if (stats) walk_fold_conv_kernel<true, true, 0>(aa, f, pt);
        else walk_fold_conv_kernel<false, true, 0>(aa, f, pt);
      } else {
        if (var & 1) walk_fold_conv_kernel<false, false, 1>(aa, f, pt);
        else if (stats) walk_fold_conv_kernel<true, false, 0>(aa, f, pt);
        else walk_fold_conv_kernel<false, false, 0>(aa, f, pt);
      }
    });
  };
#ifndef SRW_EMU_WARP
  if (peer) {                 // one lane at a time: a first pass (of the SAME instantiation: `static` stands in for __shared__) in
    WalkArgs a0 = a;          // which every lane only publishes the shard tables (s_ent / s_hash)
    a0.n_walkers = 0;
    launch(a0);
  }
#endif
  launch(a);
  if (stats_out) memcpy(stats_out, st, sizeof(st));
  return folded ? 1 : 0;
}

// The classic alias sampler (walk_alias_conv_kernel): sorted CSR + optional Vose tables (thr / row-relative alias
// index, as the twin's oa_view returns them) laid out as graph_build.cu does: RowMeta per row, hash sets placed
// from the row extent, 16-byte AliasSlot per entry.
extern "C" int emu_alias_walk(int64_t nv, const int64_t *off, const int32_t *col, const uint32_t *thr, const uint32_t *alias,
                              uint64_t t_ret, uint64_t t_common, uint64_t t_far, uint64_t seed, int32_t walk_length,
                              uint64_t walker_first, int64_t n_walkers, int32_t *paths, int32_t *lens, int var, int extra,
                              unsigned long long *stats_out) {
  const int64_t nnz = off[nv];
  std::vector<RowMeta> meta((size_t)nv);
  std::vector<int32_t> hash((size_t)(((nnz >> 2) + 1) * 8), -1);
  std::vector<AliasSlot> slot;
  for (int64_t r = 0; r < nv; ++r) {
    RowMeta m;
    m.off = off[r]; m.deg = (uint32_t)(off[r + 1] - off[r]);
    m.hoff = srw_hash_first(m.off); m.nb = srw_hash_buckets(m.off, m.deg); m.w_sum = (double)m.deg;
    meta[(size_t)r] = m;
    for (int64_t e = off[r]; e < off[r + 1]; ++e) hash_insert(hash, m.off, m.deg, col[e]);
  }
  if (thr) {
    slot.resize((size_t)nnz);
    for (int64_t r = 0; r < nv; ++r)
      for (int64_t e = off[r]; e < off[r + 1]; ++e) {
        AliasSlot s;
        s.thr = thr[e]; s.own = col[e]; s.alias_index = alias[e]; s.alias_vertex = col[off[r] + alias[e]];
        slot[(size_t)e] = s;
      }
  }
  WalkArgs a{};
  a.off = off; a.col = col; a.slot = thr ? slot.data() : nullptr; a.nv = nv; a.walker_first = walker_first; a.n_walkers = n_walkers;
  a.stride = walk_length + 2; a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  a.t_ret = t_ret; a.t_common = t_common; a.t_far = t_far;
  a.paths = paths; a.lens = lens;
  unsigned long long st[4] = {0, 0, 0, 0};
  a.stats = st;
  run_grid(n_walkers, extra, [&] {
    if (thr) {
      if (var & 1) walk_alias_conv_kernel<true, false, 1>(a, meta.data(), hash.data());
      else walk_alias_conv_kernel<true, true, 0>(a, meta.data(), hash.data());
    } else {
      if (var & 1) walk_alias_conv_kernel<false, false, 1>(a, meta.data(), hash.data());
      else walk_alias_conv_kernel<false, true, 0>(a, meta.data(), hash.data());
    }
  });
  if (stats_out) memcpy(stats_out, st, sizeof(st));
  return 0;
}

// The weighted alias-fold sampler (walk_wfold_conv_kernel): Vose tables, row weight sums and bundle weights as the
// twin computes them (oa_view / oa_wsum / oa_wbundle), laid out as graph_build.cu does (RowMeta.w_sum, AliasSlotW).
extern "C" int emu_wfold_walk(int64_t nv, const int64_t *off, const int32_t *col, const uint32_t *thr, const uint32_t *alias,
                              const double *wsum, const double *wb, double p, double q, uint64_t seed, int32_t walk_length,
                              uint64_t walker_first, int64_t n_walkers, int32_t *paths, int32_t *lens, int var, int extra,
                              unsigned long long *stats_out) {
  const int64_t nnz = off[nv];
  std::vector<RowMeta> meta((size_t)nv);
  std::vector<int32_t> hash((size_t)(((nnz >> 2) + 1) * 8), -1);
  std::vector<AliasSlotW> slot((size_t)nnz);
  for (int64_t r = 0; r < nv; ++r) {
    RowMeta m;
    m.off = off[r]; m.deg = (uint32_t)(off[r + 1] - off[r]);
    m.hoff = srw_hash_first(m.off); m.nb = srw_hash_buckets(m.off, m.deg); m.w_sum = wsum[r];
    meta[(size_t)r] = m;
    for (int64_t e = off[r]; e < off[r + 1]; ++e) {
      hash_insert(hash, m.off, m.deg, col[e]);
      AliasSlotW s;
      s.thr = thr[e]; s.own = col[e]; s.alias_index = alias[e]; s.alias_vertex = col[off[r] + alias[e]];
      s.wb_own = wb[e]; s.wb_alias = wb[off[r] + alias[e]];
      slot[(size_t)e] = s;
    }
  }
  WalkArgs a{};
  a.off = off; a.col = col; a.nv = nv; a.walker_first = walker_first; a.n_walkers = n_walkers;
  a.stride = walk_length + 2; a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  a.paths = paths; a.lens = lens;
  unsigned long long st[4] = {0, 0, 0, 0};
  a.stats = st;
  FoldArgs f{};
  if (!srw_fold_args(p, q, true, &f)) return -2;
  run_grid(n_walkers, extra, [&] {
    if (var & 1) walk_wfold_conv_kernel<false, 1>(a, f, meta.data(), hash.data(), slot.data());
    else walk_wfold_conv_kernel<true, 0>(a, f, meta.data(), hash.data(), slot.data());
  });
  if (stats_out) memcpy(stats_out, st, sizeof(st));
  return 0;
}


// ID SPACE (graph_build.cu, SRW_FOLD_IDS): entries and hash sets carry ORIGINAL vertex ids, the kernel (IDS = true) emits ids.
extern "C" int emu_fold_walk_ids(int64_t nv, const int64_t *off, const int32_t *col, const uint32_t *mult, const int32_t *vids,
                                 double p, double q, uint64_t seed, int32_t walk_length, uint64_t walker_first, int64_t n_walkers,
                                 int32_t *paths, int32_t *lens, int var, int extra) {
  const int64_t nnz = off[nv];
  std::vector<NbrEntry> ent((size_t)nnz);
  std::vector<int32_t> hash((size_t)(((nnz >> 2) + 1) * 8), -1);
  for (int64_t r = 0; r < nv; ++r)
    for (int64_t e = off[r]; e < off[r + 1]; ++e) {
      const int32_t x = col[e];
      NbrEntry ne;
      ne.x = vids[x]; ne.deg = (uint32_t)(off[x + 1] - off[x]); ne.off_lo = (uint32_t)off[x]; ne.off_hi_mult = mult[e] << 8;
      ent[(size_t)e] = ne;
      hash_insert(hash, off[r], (uint32_t)(off[r + 1] - off[r]), vids[x]);
    }
  WalkArgs a{};
  a.off = off; a.vids = vids; a.nv = nv; a.walker_first = walker_first; a.n_walkers = n_walkers; a.stride = walk_length + 2;
  a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  a.paths = paths; a.lens = lens;
  unsigned long long st[4] = {0, 0, 0, 0};
  a.stats = st;
  FoldArgs f{};
  if (!srw_fold_args(p, q, true, &f)) return -2;
  f.ent = ent.data(); f.hash = hash.data();
  PeerTable pt{};
  run_grid(n_walkers, extra, [&] {
    if (var & 1) walk_fold_conv_kernel<false, false, 1, 4, true>(a, f, pt);
    else walk_fold_conv_kernel<false, false, 0, 4, true>(a, f, pt);
  });
  return 0;
}
