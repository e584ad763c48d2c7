// migrate.cu -- host side of the migrating-walker sharded walk (kernel: migrate.cuh; SURVEY 8(e), reference seam RW:91-162,
// RW:186-192, URW:103-112).  One srw_mig context per rank owns the rank's part of the exchange: a peer-visible BLOCK
// (double-buffered inbox + this rank's home path rows), the local counters, and the pointer tables into every peer's block.
// The block is allocated by the caller because how memory becomes peer-visible is the runtime's business: a symmetric-memory
// allocation between processes (sharded.py: torch.distributed._symmetric_memory), plain cudaMalloc memory with peer access
// enabled inside one process (srw_walk with num_gpus > 1, the in-process tests).
#include <cuda_runtime.h>

#include <vector>

#include "srw_internal.h"

namespace {
#include "migrate.cuh"

inline int64_t up256(int64_t x) { return (x + 255) & ~(int64_t)255; }

struct MigLayout {
  int64_t slots;        // per inbox buffer: world * seg_cap + spill_cap
  int64_t o_cnt, o_base[2], o_paths, total;
};
MigLayout mig_layout(int world, int64_t seg_cap, int64_t spill_cap, int64_t path_rows, int32_t stride) {
  MigLayout L;
  L.slots = (int64_t)world * seg_cap + spill_cap;
  int64_t o = 0;
  L.o_cnt = o; o += up256(2 * kMigMaxDest * 8);
  for (int b = 0; b < 2; ++b) { L.o_base[b] = o; o += up256(L.slots * 48); }
  L.o_paths = o; o += up256(path_rows * (int64_t)stride * 4);
  L.total = o;
  return L;
}

__global__ void mig_init_paths_kernel(int64_t rows_per_round, int64_t n_rounds, int32_t stride, int rank, int world, int32_t *paths, int32_t *lens) {
  const int64_t total = rows_per_round * n_rounds;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    paths[i * stride] = (int32_t)(rank + (i % rows_per_round) * world);     // RW:84-87 path = Array(vId), as a rank
    lens[i] = stride;                                                     // undirected graph: no dead ends (RW:115-119 never fires)
  }
}

// ranks -> original vertex ids over the home rows (one warp per row); *steps_out += sum(len - 1)
__global__ void mig_finalize_kernel(int64_t n_rows, int32_t stride, const int32_t *__restrict__ vids, const int32_t *__restrict__ lens,
                                    int32_t *paths, unsigned long long *steps_out) {
  unsigned long long steps = 0;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += n_warps) {
    const int32_t len = lens[r];
    int32_t *row = paths + r * stride;
    for (int32_t k = lane; k < stride; k += 32) row[k] = k < len ? __ldg(vids + row[k]) : -1;
    if (lane == 0) steps += (unsigned long long)(len > 0 ? len - 1 : 0);
  }
  for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
  if (lane == 0 && steps) atomicAdd(steps_out, steps);
}
}  // namespace

struct srw_mig {
  const srw_graph *g = nullptr;
  srw_params prm;
  int world = 1, rank = 0;
  int64_t n_rounds = 0 /* capacity */, n_active = 0 /* rounds of the current batch */, seg_cap = 0, spill_cap = 0, home_rows = 0, home_rows_max = 0, round_first = 0;
  int32_t stride = 0;
  MigLayout L;
  char *block = nullptr;
  char *peers[SRW_MAX_SHARDS] = {};
  unsigned long long *d_scratch = nullptr;     // cursor, done, out_cnt[kMigMaxDest], stats[8]
  int32_t *d_lens = nullptr;
  unsigned grid = 0;
  bool stats = false;
  int minb = 4, stage = kMigStage;
  int flags = 0;
  void *attr_kern = nullptr;
  MigArgs base;                                // everything that does not change between super-steps
};

namespace {
constexpr int kScratchWords = 2 + kMigMaxDest + 8 + 1;   // cursor, done, out_cnt[], stats[8], finish scratch

srw_status mig_check(const srw_graph *g, const srw_params *p) {
  if (!g || !p) { srw_set_error("srw_mig: null graph or params"); return SRW_ERR_ARG; }
  if (g->directed) { srw_set_error("the migrating sharded walk needs an undirected graph (the membership test runs at owner(x): t in N(x))"); return SRW_ERR_UNSUPPORTED; }
  // (a shard without rows -- a partition id no edge carries -- holds no row arrays)
  if ((g->nnz > 0 && (!g->d_ent || !g->d_hash)) || !g->d_bloom) { srw_set_error("the migrating sharded walk needs an unweighted shard built with SRW_BUILD_ALIAS | SRW_BUILD_MIGRATE"); return SRW_ERR_UNSUPPORTED; }
  if (p->sampler == SRW_SAMPLER_EXACT) { srw_set_error("the sharded walk implements --sampler alias | fold"); return SRW_ERR_UNSUPPORTED; }
  if (p->walk_length < 0 || p->walk_length > 65000) { srw_set_error("sharded walk: walkLength must be in [0, 65000]"); return SRW_ERR_ARG; }
  if (!(p->p > 0.0) || !(p->q > 0.0)) { srw_set_error("p and q must be > 0"); return SRW_ERR_ARG; }
  if (p->u_mode != SRW_U_PHILOX) { srw_set_error("the sharded walk draws from Philox only"); return SRW_ERR_UNSUPPORTED; }
  return SRW_OK;
}
int64_t default_seg_cap(const srw_graph *g, int64_t n_rounds, unsigned grid) {
  const int64_t n = g->nv * n_rounds, W = g->shard_world;
  // Every walker resident on a source LEAVES during a super-step (it keeps stepping until it does), to one of the W - 1 other
  // shards: a balanced source of n / W walkers sends n / (W (W - 1)) to each.  Regions hold 1.5 n / (W - 1) -- W / 1.5 times the
  // balanced flow, enough for the first super-step, where the shard with most VERTICES seeds far more than n / W walkers --
  // capped at n (nobody can send more than every walker).  The slack covers the NOP padding of the warps' open chunks.  A region
  // that still fills up spills locally (correct, one super-step later).
  int64_t cap = W > 1 ? (3 * n + 2 * (W - 1) - 1) / (2 * (W - 1)) : 0;
  if (cap > n) cap = n;
  return (cap + (int64_t)grid * 8 * kMigChunk + 1024 + 31) & ~(int64_t)31;      // whole 32-slot blocks (mig_word)
}
// Default: 4 blocks per SM; 16 staged tuples per warp and destination up to 4 shards, 8 beyond.  What matters is the shared-memory
// CONFIGURATION the four resident blocks force on the SM: 8 destinations x 16 tuples were 49 KB per block = the 228 KB configuration =
// 28 KB of L1, too little to hold the outstanding gathers of 1024 threads (+27 % at N = 8 with 8-tuple stages, profiles/README.md).
void mig_variant(int world, int *minb, int *stage) {
  *minb = 4; *stage = world > 4 ? 8 : kMigStage;
  const char *e = getenv("SRW_MIG_VARIANT");
  if (e) { int a = 0, b = 0; if (sscanf(e, "%d,%d", &a, &b) == 2 && a > 0 && b > 0) { *minb = a; *stage = b; } }
}
unsigned mig_grid() {
  const char *e = getenv("SRW_MIG_BLOCKS");
  if (e && atoi(e) > 0) return (unsigned)atoi(e);
  int minb, stage;
  mig_variant(1, &minb, &stage);
  return 148u * (unsigned)minb;     // persistent: one wave at the blocks per SM the kernel is compiled for
}
}  // namespace

extern "C" srw_status srw_mig_block_bytes(const srw_graph *g, const srw_params *p, int64_t n_rounds, int64_t seg_cap, int64_t *bytes) {
  SRW_TRY(mig_check(g, p));
  if (n_rounds < 1 || !bytes) { srw_set_error("srw_mig_block_bytes: bad argument"); return SRW_ERR_ARG; }
  const unsigned grid = mig_grid();
  if (seg_cap <= 0) seg_cap = default_seg_cap(g, n_rounds, grid);
  seg_cap = (seg_cap + 31) & ~(int64_t)31;
  const int64_t spill = (g->nv * n_rounds + (int64_t)grid * 8 * kMigChunk + 1024 + 31) & ~(int64_t)31;
  const int64_t hmax = (g->nv + g->shard_world - 1) / g->shard_world;
  *bytes = mig_layout(g->shard_world, seg_cap, spill, hmax * n_rounds, p->walk_length + 2).total;
  return SRW_OK;
}

extern "C" srw_status srw_mig_create(const srw_graph *g, const srw_params *p, int64_t n_rounds, int64_t seg_cap, void *d_block_self,
                                     void *const *d_block_peers, srw_mig **out) {
  SRW_TRY(srw_require_device());
  SRW_TRY(mig_check(g, p));
  if (n_rounds < 1 || !d_block_self || !d_block_peers || !out) { srw_set_error("srw_mig_create: bad argument"); return SRW_ERR_ARG; }
  if ((uint64_t)g->nv * (uint64_t)n_rounds >= (1ull << 32)) { srw_set_error("srw_mig_create: %lld rounds x %lld vertices do not fit the 32-bit batch-local walker id; use smaller batches", (long long)n_rounds, (long long)g->nv); return SRW_ERR_ARG; }
  SRW_CUDA(cudaSetDevice(g->device));
  srw_mig *m = new srw_mig();
  m->g = g; m->prm = *p; m->world = g->shard_world; m->rank = g->shard_rank; m->n_rounds = n_rounds;
  m->grid = mig_grid();
  mig_variant(m->world, &m->minb, &m->stage);
  m->flags = getenv("SRW_MIG_FLAGS") ? atoi(getenv("SRW_MIG_FLAGS")) : 0;
  m->seg_cap = ((seg_cap > 0 ? seg_cap : default_seg_cap(g, n_rounds, m->grid)) + 31) & ~(int64_t)31;
  if (m->seg_cap < kMigChunk) m->seg_cap = kMigChunk;      // a region must hold at least one chunk, or nothing is ever delivered
  m->spill_cap = (g->nv * n_rounds + (int64_t)m->grid * 8 * kMigChunk + 1024 + 31) & ~(int64_t)31;
  m->stride = p->walk_length + 2;
  m->home_rows = (g->nv - m->rank + m->world - 1) / m->world;
  m->home_rows_max = (g->nv + m->world - 1) / m->world;
  m->L = mig_layout(m->world, m->seg_cap, m->spill_cap, m->home_rows_max * n_rounds, m->stride);
  if (m->L.slots >= ((int64_t)1 << 32) || m->home_rows_max * n_rounds > (int64_t)kMigRowMask) {
    srw_set_error("srw_mig_create: a batch of %lld rounds needs %lld inbox slots / %lld path rows per rank; the tuple format holds 2^32 / 2^28: use smaller batches",
                  (long long)n_rounds, (long long)m->L.slots, (long long)(m->home_rows_max * n_rounds));
    delete m;
    return SRW_ERR_ARG;
  }
  m->block = (char *)d_block_self;
  for (int r = 0; r < m->world; ++r) m->peers[r] = r == m->rank ? m->block : (char *)d_block_peers[r];
  for (int r = 0; r < m->world; ++r)
    if (!m->peers[r]) { srw_set_error("srw_mig_create: no block for rank %d", r); delete m; return SRW_ERR_ARG; }
  if (cudaMalloc(&m->d_scratch, kScratchWords * 8) != cudaSuccess || cudaMalloc(&m->d_lens, (size_t)(m->home_rows * n_rounds + 1) * 4) != cudaSuccess) {
    srw_set_error("srw_mig_create: out of device memory");
    cudaFree(m->d_scratch); delete m;
    return SRW_ERR_CUDA;
  }
  SRW_CUDA(cudaMemset(m->d_scratch, 0, kScratchWords * 8));
  SRW_CUDA(cudaMemset(m->block + m->L.o_cnt, 0, 2 * kMigMaxDest * 8));
  MigArgs &a = m->base;
  memset(&a, 0, sizeof(a));
  a.off = g->d_off; a.ent = g->d_ent; a.hash = g->d_hash; a.bloom = (const unsigned long long *)g->d_bloom; a.bloom_words = (uint32_t)g->bloom_words;
  a.nv = g->nv; a.row_first = g->row_first; a.row_last = g->row_last; a.world = m->world; a.rank = m->rank;
  if (g->vcut) { a.ext = g->d_ext; a.owner = g->d_owner; a.lverts = g->d_lverts; a.rows_local = g->seed_rows; }
  for (int r = 0; r <= m->world; ++r) a.bounds[r] = g->bounds[(size_t)r];
  FoldArgs f;
  const bool folded = srw_fold_args(p->p, p->q, p->sampler == SRW_SAMPLER_ALIAS_FOLD, &f);
  if (folded) { a.a = f.a; a.mp = f.mp; a.t_ret = f.t_ret; a.t_common = f.t_common; a.t_far = f.t_far; }
  else { a.a = 0.0; a.mp = 1.0; srw_alias_thresholds(p->p, p->q, &a.t_ret, &a.t_common, &a.t_far); }
  a.seed_lo = (uint32_t)p->seed; a.seed_hi = (uint32_t)(p->seed >> 32);
  a.stride = m->stride; a.n_rounds = n_rounds;
  a.seg_cap = m->seg_cap; a.spill_cap = m->spill_cap;
  for (int h = 0; h < m->world; ++h) {
    a.home_paths[h] = (int32_t *)(m->peers[h] + m->L.o_paths);
    a.home_rows[h] = (g->nv - h + m->world - 1) / m->world;
  }
  a.debug = getenv("SRW_MIG_DEBUG") ? atoi(getenv("SRW_MIG_DEBUG")) : 0;     // timing experiments only: the paths are wrong with it
  a.cursor = m->d_scratch; a.done_warps = m->d_scratch + 1; a.out_cnt = m->d_scratch + 2; a.stats = m->d_scratch + 2 + kMigMaxDest;
  *out = m;
  return SRW_OK;
}

extern "C" void srw_mig_free(srw_mig *m) {
  if (!m) return;
  cudaFree(m->d_scratch); cudaFree(m->d_lens);
  delete m;
}

extern "C" srw_status srw_mig_collect_stats(srw_mig *m, int enable) {
  if (!m) return SRW_ERR_ARG;
  m->stats = enable != 0;
  return SRW_OK;
}

// Start a batch: rounds [round_first, round_first + n_rounds).  Counters zeroed, both inbox count vectors zeroed, home rows = [v].
// Every rank must have finished srw_mig_begin (barrier) before any rank runs super-step 0: peers write into this block.
extern "C" srw_status srw_mig_begin(srw_mig *m, int64_t round_first, int64_t n_rounds, void *stream_) {
  SRW_TRY(srw_require_device());
  if (!m || n_rounds < 1 || n_rounds > m->n_rounds) { srw_set_error("srw_mig_begin: the context holds batches of 1..%lld rounds", m ? (long long)m->n_rounds : 0LL); return SRW_ERR_ARG; }
  m->n_active = n_rounds;
  cudaStream_t stream = (cudaStream_t)stream_;
  SRW_CUDA(cudaSetDevice(m->g->device));
  m->round_first = round_first;
  SRW_CUDA(cudaMemsetAsync(m->d_scratch, 0, kScratchWords * 8, stream));
  SRW_CUDA(cudaMemsetAsync(m->block + m->L.o_cnt, 0, 2 * kMigMaxDest * 8, stream));
  const int64_t total = m->home_rows * m->n_active;
  if (total > 0) {
    int64_t b = (total + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    mig_init_paths_kernel<<<(unsigned)b, 256, 0, stream>>>(m->home_rows, m->n_active, m->stride, m->rank, m->world,
                                                           (int32_t *)(m->block + m->L.o_paths), m->d_lens);
  }
  SRW_CUDA(cudaGetLastError());
  return SRW_OK;
}

// One super-step on this rank (asynchronous on `stream`): consume inbox buffer s&1, fill every peer's buffer (s+1)&1.  When
// d_sent is not NULL the number of inbox slots this rank filled (0 on every rank <=> no walker is left, RW:162) is also
// copied there -- a device word the caller all-reduces (the barrier between super-steps).
extern "C" srw_status srw_mig_superstep(srw_mig *m, int64_t s, unsigned long long *d_sent, void *stream_) {
  if (!m || s < 0) { srw_set_error("srw_mig_superstep: bad argument"); return SRW_ERR_ARG; }
  cudaStream_t stream = (cudaStream_t)stream_;
  SRW_CUDA(cudaSetDevice(m->g->device));
  MigArgs a = m->base;
  const int cur = (int)(s & 1), nxt = cur ^ 1;
  a.walker_base = (uint64_t)m->round_first * (uint64_t)m->g->nv;
  a.in_base = (const int4 *)(m->block + m->L.o_base[cur]);
  a.in_cnt = (const unsigned long long *)(m->block + m->L.o_cnt) + cur * kMigMaxDest;
  a.n_rounds = m->n_active;
  {
    // seeds: every other walker of this shard in super-step 0, the rest in super-step 1 (see MigArgs::seed_step)
    const int64_t seeds = (m->g->vcut ? m->g->seed_rows : m->g->row_last - m->g->row_first) * m->n_active;
    a.seed_step = 2; a.seed_first = s;
    a.n_seed = s == 0 ? (seeds + 1) / 2 : s == 1 ? seeds / 2 : 0;
  }
  for (int d = 0; d <= m->world; ++d) {
    char *blk = d == m->world ? m->block : m->peers[d];
    const int64_t first = d == m->world ? (int64_t)m->world * m->seg_cap : (int64_t)m->rank * m->seg_cap;
    a.out_base[d] = (int4 *)(blk + m->L.o_base[nxt]) + 3 * first;
    a.out_cnt_pub[d] = (unsigned long long *)(blk + m->L.o_cnt) + nxt * kMigMaxDest + (d == m->world ? m->world : m->rank);
  }
  // kernel variant = (resident blocks per SM the kernel is compiled for, tuples staged per warp and destination): (4, 16) by default;
  // SRW_MIG_VARIANT=minb,stage selects another one (measurement knob: the kernel is latency-bound, profiles/README.md)
  const size_t dyn = (size_t)8 * (m->world > 1 ? m->world - 1 : 1) * 3 * m->stage * sizeof(int4);      // the warps' stages (migrate.cuh): one per OTHER rank
  void (*kern)(const MigArgs) = nullptr;
#define MIG_PICK(MB, ST)                                                                                                              \
  if (m->minb == MB && m->stage == ST)                                                                                                \
    kern = m->g->vcut ? (m->stats ? mig_step_kernel<true, MB, true, ST> : mig_step_kernel<false, MB, true, ST>)                       \
                      : (m->stats ? mig_step_kernel<true, MB, false, ST> : mig_step_kernel<false, MB, false, ST>);
  MIG_PICK(4, 16) MIG_PICK(4, 8) MIG_PICK(5, 8) MIG_PICK(6, 8) MIG_PICK(4, 4) MIG_PICK(3, 8) MIG_PICK(4, 32)
#undef MIG_PICK
#define MIG_PICK_F(ST, FL)                                                                                                            \
  if (m->minb == 4 && m->stage == ST && m->flags == FL)                                                                               \
    kern = m->g->vcut ? (m->stats ? mig_step_kernel<true, 4, true, ST, FL> : mig_step_kernel<false, 4, true, ST, FL>)                 \
                      : (m->stats ? mig_step_kernel<true, 4, false, ST, FL> : mig_step_kernel<false, 4, false, ST, FL>);
  // SRW_MIG_FLAGS (measurement knobs): 1 = the inbox lines of a claim are prefetched into L2, 2 = random gathers do not allocate in L1
  MIG_PICK_F(8, 1) MIG_PICK_F(16, 1) MIG_PICK_F(8, 2) MIG_PICK_F(16, 2)
#undef MIG_PICK_F
  if (!kern) { srw_set_error("SRW_MIG_VARIANT: no kernel variant (%d blocks per SM, %d staged tuples)", m->minb, m->stage); return SRW_ERR_ARG; }
  if (m->attr_kern != (void *)kern) {
    SRW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    m->attr_kern = (void *)kern;
  }
  kern<<<m->grid, 256, dyn, stream>>>(a);
  SRW_CUDA(cudaGetLastError());
  if (d_sent) SRW_CUDA(cudaMemcpyAsync(d_sent, a.stats, 8, cudaMemcpyDeviceToDevice, stream));
  return SRW_OK;
}

// counters of the batch so far (synchronises the stream): [0] slots sent in the last super-step, [1] steps, [2] proposals,
// [3] membership tests, [4] exact (hash-set) tests, [5] spills, [6] error flags, [7] exact tests that found the edge
extern "C" srw_status srw_mig_counters(srw_mig *m, int64_t *h_out8, void *stream_) {
  if (!m || !h_out8) return SRW_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  SRW_CUDA(cudaSetDevice(m->g->device));
  unsigned long long h[8];
  SRW_CUDA(cudaMemcpyAsync(h, m->base.stats, 64, cudaMemcpyDeviceToHost, stream));
  SRW_CUDA(cudaStreamSynchronize(stream));
  for (int i = 0; i < 8; ++i) h_out8[i] = (int64_t)h[i];
  if (h[6]) { srw_set_error("migrating walk: device error flags 0x%llx (1 = vertex without a row, 2 = spill region overflow)", h[6]); return SRW_ERR_CUDA; }
  return SRW_OK;
}

// End of a batch (after the super-step in which no rank sent anything): ranks -> vertex ids over this rank's home rows.
// *d_paths is [home_rows * n_rounds][walk_length + 2] inside the block (valid until the next srw_mig_begin), row
// (round - round_first) * home_rows + v / world for the walker that started at vertex rank v = rank + (row % home_rows) * world.
extern "C" srw_status srw_mig_finish(srw_mig *m, int32_t **d_paths, int32_t *d_lens_out, int64_t *n_rows, int64_t *steps, void *stream_) {
  if (!m) return SRW_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  SRW_CUDA(cudaSetDevice(m->g->device));
  const int64_t rows = m->home_rows * m->n_active;
  int32_t *paths = (int32_t *)(m->block + m->L.o_paths);
  unsigned long long *d_steps = m->base.stats + 8;
  SRW_CUDA(cudaMemsetAsync(d_steps, 0, 8, stream));
  if (rows > 0) {
    int64_t b = (rows * 32 + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
    mig_finalize_kernel<<<(unsigned)b, 256, 0, stream>>>(rows, m->stride, m->g->d_vids, m->d_lens, paths, d_steps);
  }
  unsigned long long h = 0;
  SRW_CUDA(cudaMemcpyAsync(&h, d_steps, 8, cudaMemcpyDeviceToHost, stream));
  SRW_CUDA(cudaStreamSynchronize(stream));
  SRW_CUDA(cudaGetLastError());
  if (d_paths) *d_paths = paths;
  if (d_lens_out && rows > 0) SRW_CUDA(cudaMemcpy(d_lens_out, m->d_lens, (size_t)rows * 4, cudaMemcpyDeviceToDevice));
  if (n_rows) *n_rows = rows;
  if (steps) *steps = (int64_t)h;
  return SRW_OK;
}

extern "C" srw_status srw_mig_info(const srw_mig *m, int64_t *seg_cap, int64_t *spill_cap, int64_t *home_rows, int64_t *block_bytes) {
  if (!m) return SRW_ERR_ARG;
  if (seg_cap) *seg_cap = m->seg_cap;
  if (spill_cap) *spill_cap = m->spill_cap;
  if (home_rows) *home_rows = m->home_rows;
  if (block_bytes) *block_bytes = m->L.total;
  return SRW_OK;
}
