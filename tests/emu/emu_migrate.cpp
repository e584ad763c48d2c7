// emu_migrate.cpp -- TEST INFRASTRUCTURE: runs the product's mig_step_kernel SOURCE (csrc/migrate.cuh: the migrating-walker
// super-step of the sharded walk) on the host under the lockstep 32-lane warp emulator (warp_emu.h), for `shards` vertex
// ranges laid out as graph_build.cu / migrate.cu do (16-byte neighbour entries with owner and owner-local offsets, per-row
// hash sets, the replicated edge filter, double-buffered inboxes with one region per source and a spill region, home path
// rows).  "GPUs" are emulated one after another inside a super-step, which is exactly the ordering the barrier between
// super-steps guarantees.  The test compares the assembled paths with the CPU twin (oracle_alias_walk).
#include "warp_emu.h"
static emu_dim3 gridDim = {1, 1, 1};
static inline unsigned int atomicAdd(unsigned int *p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

#include <vector>

#include "../../stellar-random-walk_b200/csrc/migrate.cuh"

namespace {
struct Shard {
  std::vector<int64_t> off;
  std::vector<NbrEntry> ent;
  std::vector<int32_t> hash;
  std::vector<int4> base[2];
  unsigned long long cnt[2][kMigMaxDest];
  std::vector<int32_t> paths;
  unsigned long long scratch[2 + kMigMaxDest + 9];
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

// out_paths: [n_rounds * nv][stride] vertex RANKS in global walker order (round, start rank); stats_out[8]:
// [0] super-steps, [1] steps, [2] proposals, [3] membership tests, [4] exact tests, [5] spills, [6] error flags, [7] tuples sent
extern "C" int emu_migrate_walk(int64_t nv, const int64_t *off, const int32_t *col, const uint32_t *mult, int shards, const int64_t *bounds,
                                double p, double q, int fold, uint64_t t_ret, uint64_t t_common, uint64_t t_far, uint64_t seed,
                                int32_t walk_length, int64_t round_first, int64_t n_rounds, int64_t seg_cap, int bloom_bits, int grid_blocks,
                                int32_t *out_paths, unsigned long long *stats_out, const uint8_t *owner_map /* NULL = vertex ranges; else the VCut shard map */,
                                uint32_t hub_deg /* 0 = none; else rows of at least this degree are replicated on every shard */) {
  if (shards < 1 || shards > SRW_MAX_SHARDS) return -1;
  const int W = shards;
  const int32_t stride = walk_length + 2;
  const int64_t nnz = off[nv];
  auto owner_of = [&](int64_t v) { if (owner_map) return (int)owner_map[v]; int o = 0; while (o + 1 < W && v >= bounds[o + 1]) o++; return o; };
  // table-mapped shards (graph_build.cu): every shard lays its arrays out as [hub rows | own rows]; a hub row (routing owner
  // kMigHub) sits at the same offset on every shard.  Seeds belong to the TRUE owner of a vertex, hub or not.
  const bool mapped = owner_map != nullptr || hub_deg > 0;
  std::vector<std::vector<int32_t>> rows((size_t)W), seeds((size_t)W);
  std::vector<int32_t> hubs;
  std::vector<MigExt> ext((size_t)nv);
  std::vector<uint8_t> own((size_t)nv);
  int64_t hub_entries = 0;
  for (int64_t v = 0; v < nv; ++v) {
    const int o = owner_of(v);
    if (o < 0 || o >= W) return -2;
    seeds[(size_t)o].push_back((int32_t)v);
    const uint32_t deg = (uint32_t)(off[v + 1] - off[v]);
    if (hub_deg > 0 && deg >= hub_deg) { own[(size_t)v] = (uint8_t)kMigHub; hubs.push_back((int32_t)v); ext[(size_t)v].off = (uint32_t)hub_entries; ext[(size_t)v].deg = deg; hub_entries += deg; }
    else own[(size_t)v] = (uint8_t)o;
  }
  {
    std::vector<int64_t> fill((size_t)W, hub_entries);
    for (int s = 0; s < W; ++s) rows[(size_t)s] = hubs;
    for (int64_t v = 0; v < nv; ++v) {
      if (own[(size_t)v] == (uint8_t)kMigHub) continue;
      const int o = own[(size_t)v];
      rows[(size_t)o].push_back((int32_t)v);
      ext[(size_t)v].off = (uint32_t)fill[(size_t)o]; ext[(size_t)v].deg = (uint32_t)(off[v + 1] - off[v]);
      fill[(size_t)o] += off[v + 1] - off[v];
    }
  }
  // the replicated filter
  uint32_t bloom_words = (uint32_t)((nnz / 2 * bloom_bits + 63) / 64);
  if (bloom_words < 4) bloom_words = 4;
  std::vector<unsigned long long> bloom((size_t)bloom_words, 0ull);
  for (int64_t r = 0; r < nv; ++r)
    for (int64_t e = off[r]; e < off[r + 1]; ++e) {
      uint32_t word;
      uint64_t mask;
      srw_bloom_probe((int32_t)r, col[e], bloom_words, &word, &mask);
      bloom[(size_t)word] |= mask;
    }
  const int64_t spill_cap = (nv * n_rounds + (int64_t)grid_blocks * 8 * kMigChunk + 64 + 31) & ~(int64_t)31;
  if (seg_cap <= 0) seg_cap = (nv * n_rounds + W - 1) / W + (int64_t)grid_blocks * 8 * kMigChunk + 64;
  seg_cap = (seg_cap + 31) & ~(int64_t)31;       // whole 32-slot blocks (mig_word)
  const int64_t slots = (int64_t)W * seg_cap + spill_cap;
  std::vector<Shard> sh((size_t)W);
  for (int s = 0; s < W; ++s) {
    Shard &R = sh[(size_t)s];
    const std::vector<int32_t> &lv = rows[(size_t)s];
    R.off.resize(lv.size() + 1);
    int64_t n = 0;
    for (size_t i = 0; i < lv.size(); ++i) { R.off[i] = n; n += off[lv[i] + 1] - off[lv[i]]; }
    R.off[lv.size()] = n;
    R.ent.resize((size_t)n);
    R.hash.assign((size_t)(((n >> 2) + 1) * 8), -1);
    for (int b = 0; b < 2; ++b) { R.base[b].assign((size_t)slots * 3, make_int4(-1, -1, -1, -1)); }
    memset(R.cnt, 0, sizeof(R.cnt));
    memset(R.scratch, 0, sizeof(R.scratch));
    const int64_t hrows = (nv - s + W - 1) / W;
    R.paths.assign((size_t)(hrows * n_rounds * stride + 1), -7);
    for (int64_t i = 0; i < hrows * n_rounds; ++i) R.paths[(size_t)(i * stride)] = (int32_t)(s + (i % hrows) * W);
    for (size_t i = 0; i < lv.size(); ++i) {
      const int64_t r = lv[i], lo = R.off[i];
      if ((uint32_t)lo != ext[(size_t)r].off) return -4;
      const uint32_t deg = (uint32_t)(off[r + 1] - off[r]);
      for (int64_t e = off[r]; e < off[r + 1]; ++e) {
        const int32_t x = col[e];
        NbrEntry ne;
        ne.x = x; ne.deg = ext[(size_t)x].deg; ne.off_lo = ext[(size_t)x].off;
        ne.off_hi_mult = (uint32_t)own[(size_t)x] | (mult[e] << 8);
        R.ent[(size_t)(lo + (e - off[r]))] = ne;
        hash_insert(R.hash, lo, deg, x);
      }
    }
  }
  FoldArgs f{};
  const bool folded = srw_fold_args(p, q, fold != 0, &f);
  if (!folded) { f.a = 0.0; f.mp = 1.0; f.t_ret = t_ret; f.t_common = t_common; f.t_far = t_far; }
  unsigned long long tuples = 0;
  int64_t steps = 0;
  for (int64_t s = 0;; ++s) {
    const int cur = (int)(s & 1), nxt = cur ^ 1;
    for (int r = 0; r < W; ++r) {
      Shard &R = sh[(size_t)r];
      MigArgs a{};
      a.off = R.off.data(); a.ent = R.ent.data(); a.hash = R.hash.data(); a.bloom = bloom.data(); a.bloom_words = bloom_words;
      a.nv = nv; a.world = W; a.rank = r;
      if (mapped) { a.ext = ext.data(); a.owner = own.data(); a.lverts = seeds[(size_t)r].data(); a.rows_local = (int64_t)seeds[(size_t)r].size(); }
      else {
        a.row_first = bounds[r]; a.row_last = bounds[r + 1];
        for (int k = 0; k <= W; ++k) a.bounds[k] = bounds[k];
      }
      a.a = f.a; a.mp = f.mp; a.t_ret = f.t_ret; a.t_common = f.t_common; a.t_far = f.t_far;
      a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32); a.stride = stride;
      a.walker_base = (uint64_t)round_first * (uint64_t)nv; a.n_rounds = n_rounds;
      a.in_base = R.base[cur].data(); a.in_cnt = R.cnt[cur];
      a.seg_cap = seg_cap; a.spill_cap = spill_cap;
      {
        const int64_t n_seeds = (int64_t)seeds[(size_t)r].size() * n_rounds;
        a.seed_step = 2; a.seed_first = s;
        a.n_seed = s == 0 ? (n_seeds + 1) / 2 : s == 1 ? n_seeds / 2 : 0;
      }
      for (int d = 0; d <= W; ++d) {
        Shard &D = d == W ? R : sh[(size_t)d];
        const int64_t first = d == W ? (int64_t)W * seg_cap : (int64_t)r * seg_cap;
        a.out_base[d] = D.base[nxt].data() + 3 * first;
        a.out_cnt_pub[d] = &D.cnt[nxt][d == W ? W : r];
      }
      for (int h = 0; h < W; ++h) { a.home_paths[h] = sh[(size_t)h].paths.data(); a.home_rows[h] = (nv - h + W - 1) / W; }
      a.cursor = R.scratch; a.done_warps = R.scratch + 1; a.out_cnt = R.scratch + 2; a.stats = R.scratch + 2 + kMigMaxDest;
      gridDim.x = (unsigned)grid_blocks;
      // the product's default variants (migrate.cu mig_variant): 16 staged tuples per warp and destination up to 4 shards, 8 beyond
      if (mapped && W > 4) emu_launch_warps((int64_t)grid_blocks * 8, [&] { mig_step_kernel<true, 4, true, 8>(a); });
      else if (mapped) emu_launch_warps((int64_t)grid_blocks * 8, [&] { mig_step_kernel<true, 4, true>(a); });
      else if (W > 4) emu_launch_warps((int64_t)grid_blocks * 8, [&] { mig_step_kernel<true, 4, false, 8>(a); });
      else emu_launch_warps((int64_t)grid_blocks * 8, [&] { mig_step_kernel<true>(a); });
    }
    unsigned long long sent = 0;
    for (int r = 0; r < W; ++r) sent += sh[(size_t)r].scratch[2 + kMigMaxDest];
    tuples += sent;
    steps = s + 1;
    if (sent == 0 && s >= 1) break;      // super-step 1 still seeds walkers
    if (s > 100000) return -3;
  }
  // assemble in global walker order
  for (int64_t rnd = 0; rnd < n_rounds; ++rnd)
    for (int64_t v = 0; v < nv; ++v) {
      const int h = (int)(v % W);
      const int64_t hrows = (nv - h + W - 1) / W;
      memcpy(out_paths + (rnd * nv + v) * stride, sh[(size_t)h].paths.data() + (rnd * hrows + v / W) * stride, (size_t)stride * 4);
    }
  if (stats_out) {
    memset(stats_out, 0, 64);
    stats_out[0] = (unsigned long long)steps;
    for (int r = 0; r < W; ++r)
      for (int k = 1; k <= 6; ++k) {
        if (k == 6) stats_out[k] |= sh[(size_t)r].scratch[2 + kMigMaxDest + k];
        else stats_out[k] += sh[(size_t)r].scratch[2 + kMigMaxDest + k];
      }
    stats_out[7] = tuples;
  }
  return folded ? 1 : 0;
}
