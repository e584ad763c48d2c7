// graph_build.cu -- K1/K2/K3: edge list -> CSR in HBM (+ neighbour-sorted copy, + Vose alias slots).
//
// Replaces the reference's Spark adjacency build (URW:35-43 `reduceByKey(_ ++ _)`, VRW:35-49) and
// GraphMap.addVertex (GM:23-64): per-vertex neighbour arrays in FILE-APPEARANCE order (for one input
// line the src-side entry precedes the dst-side entry), duplicates and self-loops kept, every id seen
// in the file is a vertex (|V| = distinct ids, RW:23; a vertex without out-neighbours has an empty row,
// GM:50-52).  Vertices are ranked by ascending id; ranks index every device array.
//
// Layout produced (all in HBM, sized for 64-bit row offsets -- RMAT-26 has 2^31 adjacency entries,
// beyond the reference's Int offsetCounter, GM:18):
//   d_off[nv+1] int64, d_col_app/d_w_app [nnz] (appearance order, exact sampler),
//   d_col[nnz] (ascending per row: membership search + unweighted proposals),
//   d_slot[nnz] 16-byte Vose slots (weighted graphs only).
//
// Sorting/scanning uses CUB device primitives (a CUDA toolkit library, like cuBLAS for GEMMs); the
// build runs once per graph and is timed separately from the walk.
#include <cub/cub.cuh>

#include <chrono>

#include "philox.cuh"
#include "srw_internal.h"

namespace {

constexpr int kThreads = 256;
inline unsigned grid_for(int64_t n, int threads = kThreads) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  return (unsigned)b;
}

struct DevBuf {  // RAII scratch
  void *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T *as() { return (T *)p; }
  cudaError_t alloc(size_t bytes) { if (p) { cudaFree(p); p = nullptr; } return cudaMalloc(&p, bytes ? bytes : 16); }
  void *release() { void *q = p; p = nullptr; return q; }
};

__device__ __forceinline__ int32_t rank_of(const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ wordrank,
                                           int32_t id_min, int32_t id) {
  const uint32_t rel = (uint32_t)id - (uint32_t)id_min;
  const uint32_t w = rel >> 5, b = rel & 31u;
  return (int32_t)(wordrank[w] + __popc(bitmap[w] & ((1u << b) - 1u)));
}

__global__ void k_mark(int64_t n, const int32_t *__restrict__ ids, int32_t id_min, uint32_t *bitmap) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t rel = (uint32_t)ids[i] - (uint32_t)id_min;
    const uint32_t bit = 1u << (rel & 31u);
    if (!(bitmap[rel >> 5] & bit)) atomicOr(&bitmap[rel >> 5], bit);
  }
}
__global__ void k_popc(uint64_t words, const uint32_t *__restrict__ bitmap, uint32_t *cnt) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x)
    cnt[i] = __popc(bitmap[i]);
}
__global__ void k_vids(uint64_t words, const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ wordrank,
                       int32_t id_min, int32_t *vids) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t bits = bitmap[i];
    uint32_t r = wordrank[i];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      vids[r++] = (int32_t)((uint32_t)id_min + (uint32_t)(i * 32 + b));
    }
  }
}
// adjacency entries in appearance order: undirected entry 2e = (src -> dst), 2e+1 = (dst -> src)
// (URW:38); directed entry e = (src -> dst) (URW:36).  Also counts degrees.
__global__ void k_entries(int64_t n, const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int directed,
                          const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ wordrank, int32_t id_min,
                          uint32_t *ent_row, uint32_t *ent_col, uint32_t *deg) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t rs = rank_of(bitmap, wordrank, id_min, src[e]);
    const uint32_t rd = rank_of(bitmap, wordrank, id_min, dst[e]);
    if (directed) {
      ent_row[e] = rs; ent_col[e] = rd;
      atomicAdd(&deg[rs], 1u);
    } else {
      ent_row[2 * e] = rs; ent_col[2 * e] = rd;
      ent_row[2 * e + 1] = rd; ent_col[2 * e + 1] = rs;
      atomicAdd(&deg[rs], 1u);
      atomicAdd(&deg[rd], 1u);
    }
  }
}
// ---- sharded build: every rank scans the whole edge list, counts global degrees, and
// keeps only the adjacency entries whose row lies in its vertex range [row_first, row_last) ----
__global__ void k_degrees(int64_t n, const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int directed,
                          const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ wordrank, int32_t id_min,
                          uint32_t *deg) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    atomicAdd(&deg[rank_of(bitmap, wordrank, id_min, src[e])], 1u);
    if (!directed) atomicAdd(&deg[rank_of(bitmap, wordrank, id_min, dst[e])], 1u);
  }
}
struct ShardBoundsLite { int world; int64_t first[SRW_MAX_SHARDS + 1]; };
__device__ __forceinline__ uint32_t warp_reserve(bool want, unsigned long long *cursor) {
  const unsigned m = __ballot_sync(0xffffffffu, want);
  if (!m) return 0;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned long long base = 0;
  if (lane == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  return (uint32_t)(base + __popc(m & ((1u << lane) - 1u)));
}
// compaction order is arbitrary; the global entry index (2e / 2e+1, or e when directed) is kept so
// that a sort restores file-appearance order
__global__ void k_entries_range(int64_t n, const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int directed,
                                const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ wordrank, int32_t id_min,
                                uint32_t row_first, uint32_t row_last, uint32_t *ent_row, uint32_t *ent_col,
                                uint32_t *ent_gidx, unsigned long long *cursor) {
  const int64_t n_pad = (n + 31) & ~(int64_t)31;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_pad; e += (int64_t)gridDim.x * blockDim.x) {
    uint32_t rs = 0, rd = 0;
    const bool live = e < n;
    if (live) { rs = rank_of(bitmap, wordrank, id_min, src[e]); rd = rank_of(bitmap, wordrank, id_min, dst[e]); }
    const bool k0 = live && rs >= row_first && rs < row_last;
    const bool k1 = live && !directed && rd >= row_first && rd < row_last;
    const uint32_t p0 = warp_reserve(k0, cursor);
    if (k0) { ent_row[p0] = rs - row_first; ent_col[p0] = rd; ent_gidx[p0] = (uint32_t)(directed ? e : 2 * e); }
    const uint32_t p1 = warp_reserve(k1, cursor);
    if (k1) { ent_row[p1] = rd - row_first; ent_col[p1] = rs; ent_gidx[p1] = (uint32_t)(2 * e + 1); }
  }
}
// ---- table-mapped shards: the VCut shard map (VRW:23-26,121-134; GM:31,66-68: owner(v) = getPartition(v) mod world) and/or
// REPLICATED HUB ROWS (VRW:43-54 replicates a vertex's adjacency into every partition it has an edge in; here the rows of the
// highest-degree vertices are kept by EVERY shard, so a walker that steps onto a hub does not migrate) ----
constexpr uint32_t kHubOwner = 0xFFu;        // d_owner[v] / NbrEntry owner byte: "this row is on every shard"
__global__ void k_map_owner_pid(int64_t nv, const int32_t *__restrict__ vpid, int world, uint8_t *owner_true) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = vpid[v];                      // (-1: a vertex no edge leads to -- cannot happen on an undirected graph)
    owner_true[v] = (uint8_t)(((int)(p % world) + world) % world);     // Spark HashPartitioner: nonNegativeMod(pid.hashCode, numPartitions)
  }
}
__global__ void k_map_owner_ranges(int64_t nv, ShardBoundsLite sb, uint8_t *owner_true) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    int o = 0;
    while (o + 1 < sb.world && v >= sb.first[o + 1]) o++;
    owner_true[v] = (uint8_t)o;
  }
}
// sorted[] ascending degrees, pre[] their inclusive prefix sums: rows i .. nv-1 hold pre[nv-1] - pre[i-1] entries.  The cut is the
// smallest i whose tail fits the budget; a degree value is wholly in or wholly out (a partial run of equal degrees at the cut is dropped).
__global__ void k_hub_threshold(int64_t nv, const uint32_t *__restrict__ sorted, const int64_t *__restrict__ pre, int64_t budget, uint32_t *hub_deg) {
  const int64_t total = pre[nv - 1], need = total - budget;      // tail(i) <= budget  <=>  pre[i-1] >= need
  int64_t lo = 0, hi = nv;                                       // first index j with pre[j] >= need; the cut is i = j + 1 (i = 0 if need <= 0)
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (pre[mid] >= need) hi = mid; else lo = mid + 1; }
  const int64_t i = need <= 0 ? 0 : lo + 1;
  uint32_t T = 0xFFFFFFFFu;
  if (i < nv) {
    T = sorted[i];
    if (i > 0 && sorted[i - 1] == T) T++;
    if (T < 2) T = 2;
  }
  *hub_deg = T;
}
// first rank of shard r = lower bound of r * total / world in the exclusive prefix sums moff[0 .. nv] (one thread per boundary)
__global__ void k_range_bounds(int64_t nv, const int64_t *__restrict__ moff, int world, int64_t *bounds) {
  const int r = threadIdx.x;
  if (r < 1 || r >= world) return;
  const int64_t target = (int64_t)((__int128)moff[nv] * r / world);
  int64_t lo = 0, hi = nv + 1;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (moff[mid] >= target) hi = mid; else lo = mid + 1; }
  bounds[r] = lo;
}
// degree with the hub rows masked out (ranges are balanced on the entries that are NOT replicated)
__global__ void k_map_masked_deg(int64_t nv, const uint32_t *__restrict__ deg, uint32_t hub_deg, uint32_t *out) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v <= nv; v += (int64_t)gridDim.x * blockDim.x)
    out[v] = v < nv && deg[v] < hub_deg ? deg[v] : 0u;
}
// routing owner (kHubOwner for replicated rows) and the sort key that groups the rows: 0 = hubs, o + 1 = shard o
__global__ void k_map_keys(int64_t nv, const uint32_t *__restrict__ deg, uint32_t hub_deg, const uint8_t *__restrict__ owner_true,
                           uint8_t *owner, uint32_t *key) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    const bool hub = deg[v] >= hub_deg;
    owner[v] = hub ? (uint8_t)kHubOwner : owner_true[v];
    key[v] = hub ? 0u : (uint32_t)owner_true[v] + 1u;
  }
}
// first position of every group in the sorted key array (groups that exist)
__global__ void k_map_group_first(int64_t nv, const uint32_t *__restrict__ key, long long *first) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x)
    if (i == 0 || key[i] != key[i - 1]) first[key[i]] = i;
}
struct MapGroups { int64_t first[SRW_MAX_SHARDS + 2]; int64_t base[SRW_MAX_SHARDS + 2]; };
// position i of the (group, vertex)-sorted order holds vertex perm[i].  Every shard lays its arrays out as [hub rows | own rows]:
// a hub row starts at poff[i] (the same on every shard), an own row at hub_entries + poff[i] - base[group] inside its owner's arrays
__global__ void k_map_ext(int64_t nv, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ key, const int64_t *__restrict__ poff,
                          MapGroups gr, MigExt *ext, uint32_t *lrow) {
  const int64_t hub_rows = gr.first[1], hub_entries = gr.base[1];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = perm[i], k = key[i];
    MigExt e;
    e.deg = (uint32_t)(poff[i + 1] - poff[i]);
    if (k == 0) { e.off = (uint32_t)poff[i]; lrow[v] = (uint32_t)i; }
    else { e.off = (uint32_t)(hub_entries + poff[i] - gr.base[k]); lrow[v] = (uint32_t)(hub_rows + i - gr.first[k]); }
    ext[v] = e;
  }
}
// local row offsets of shard `rank`: the hub rows, then its own rows
__global__ void k_map_local_off(int64_t hub_rows, int64_t own_rows, int64_t own_first, int64_t own_base, int64_t hub_entries,
                                const int64_t *__restrict__ poff, int64_t *off) {
  const int64_t rows = hub_rows + own_rows;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= rows; r += (int64_t)gridDim.x * blockDim.x)
    off[r] = r < hub_rows ? poff[r] : hub_entries + poff[own_first + (r - hub_rows)] - own_base;
}
__global__ void k_entries_owner(int64_t n, const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int directed,
                                const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ wordrank, int32_t id_min,
                                const uint8_t *__restrict__ owner, const uint32_t *__restrict__ lrow, int rank, uint32_t *ent_row,
                                uint32_t *ent_col, uint32_t *ent_gidx, unsigned long long *cursor) {
  const int64_t n_pad = (n + 31) & ~(int64_t)31;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_pad; e += (int64_t)gridDim.x * blockDim.x) {
    uint32_t rs = 0, rd = 0;
    const bool live = e < n;
    if (live) { rs = rank_of(bitmap, wordrank, id_min, src[e]); rd = rank_of(bitmap, wordrank, id_min, dst[e]); }
    const bool k0 = live && (owner[rs] == rank || owner[rs] == kHubOwner);
    const bool k1 = live && !directed && (owner[rd] == rank || owner[rd] == kHubOwner);
    const uint32_t p0 = warp_reserve(k0, cursor);
    if (k0) { ent_row[p0] = lrow[rs]; ent_col[p0] = rd; ent_gidx[p0] = (uint32_t)(directed ? e : 2 * e); }
    const uint32_t p1 = warp_reserve(k1, cursor);
    if (k1) { ent_row[p1] = lrow[rd]; ent_col[p1] = rs; ent_gidx[p1] = (uint32_t)(2 * e + 1); }
  }
}
struct IsOwner {
  const uint8_t *owner_true; int rank;
  __host__ __device__ bool operator()(uint32_t v) const { return owner_true[v] == rank; }
};
// SRW_BUILD_MIGRATE: every undirected input edge {rank(src), rank(dst)} sets kMigBloomK bits of one 64-bit word
__global__ void k_bloom_insert(int64_t n, const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                               const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ wordrank, int32_t id_min,
                               unsigned long long *bloom, uint32_t n_words) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    uint32_t word;
    uint64_t mask;
    srw_bloom_probe(rank_of(bitmap, wordrank, id_min, src[e]), rank_of(bitmap, wordrank, id_min, dst[e]), n_words, &word, &mask);
    atomicOr(bloom + word, (unsigned long long)mask);
  }
}
__global__ void k_rebase_off(int64_t rows, const int64_t *__restrict__ goff, int64_t row_first, int64_t *off) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= rows; r += (int64_t)gridDim.x * blockDim.x)
    off[r] = goff[row_first + r] - goff[row_first];
}

__global__ void k_pack_rc(int64_t n, const uint32_t *__restrict__ row, const uint32_t *__restrict__ col, int cbits, unsigned long long *key) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    key[i] = ((unsigned long long)row[i] << cbits) | (unsigned long long)col[i];
}
__global__ void k_unpack_rc(int64_t n, const unsigned long long *__restrict__ key, int cbits, uint32_t *row, uint32_t *col) {
  const unsigned long long mask = (1ULL << cbits) - 1ULL;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    row[i] = (uint32_t)(key[i] >> cbits); col[i] = (uint32_t)(key[i] & mask);
  }
}
__global__ void k_iota(int64_t n, uint32_t *a) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = (uint32_t)i;
}
__global__ void k_gather_u32(int64_t n, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ table, uint32_t *out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = table[idx[i]];
}
// entry index -> input edge weight (shift = 1 when each edge made two entries)
// (gidx: local position -> global entry index, NULL = identity)
__global__ void k_gather_w(int64_t n, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ gidx,
                           const float *__restrict__ w, int shift, float *out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t p = idx[i];
    out[i] = w ? w[(gidx ? gidx[p] : p) >> shift] : 1.0f;
  }
}
__global__ void k_any_non_unit(int64_t n, const float *__restrict__ w, int *flag) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (w[i] != 1.0f) *flag = 1;
}
// GM:31 vertexPartitionMap.put(dst, pId): the partition id of the last input line (file order) in
// which the vertex is a neighbour.  (The reference's winner depends on hash-partition order: unpinned.)
__global__ void k_vpid_last(int64_t n, const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int directed,
                            const uint32_t *__restrict__ bitmap, const uint32_t *__restrict__ wordrank, int32_t id_min,
                            unsigned long long *last_line) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    atomicMax(&last_line[rank_of(bitmap, wordrank, id_min, dst[e])], (unsigned long long)(e + 1));
    if (!directed) atomicMax(&last_line[rank_of(bitmap, wordrank, id_min, src[e])], (unsigned long long)(e + 1));
  }
}
__global__ void k_vpid_fill(int64_t nv, const unsigned long long *__restrict__ last_line, const int32_t *__restrict__ pid,
                            int32_t *vpid) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x)
    vpid[v] = last_line[v] ? pid[last_line[v] - 1] : -1;
}

// K3: Vose alias tables by the in-order sweep (two cursors, no work lists) -- built IN PARALLEL over all adjacency entries.
// The sweep is defined in exact integer arithmetic (oracle/srw_oracle.c alias_row is the sequential statement):
//   W    = the row's weight sum, 32 strided partial sums combined by a butterfly (k_row_wsum: one warp per row);
//   t[k] = floor((w[k] * n / W) * 2^32): the scaled weight in units of 2^-32; light iff t < 2^32, else heavy with excess t - 2^32.
// With E_k = the summed excess of the row's first k heavy items and D_m = the summed deficit (2^32 - t) of its first m light items,
// the sweep's state (heavy k active, m light items filled) has residual 2^32 + E_k - D_m, so
//   * light item m + 1 is filled by the FIRST heavy k with E_k >= D_m           (threshold t, alias = that heavy item);
//   * heavy item k retires at the FIRST m with D_m > E_k, if a next heavy exists (threshold 2^32 + E_k - D_m, alias = the next heavy);
//   * everything the sweep never reaches keeps threshold 2^32 - 1 and itself as alias.
// Integer sums do not depend on the order of evaluation: E and D are segmented prefix sums (cub::DeviceScan::InclusiveSumByKey), the
// two "first" searches are binary searches over the row's compacted heavy / light lists -- one thread per entry, no O(max degree)
// chain on a single lane (round 1's builder walked a million-entry hub row with one thread).
__device__ __forceinline__ uint64_t alias_scaled(float w, double dn, double W) {
  return __double2ull_rz(__dmul_rn(__ddiv_rn(__dmul_rn((double)w, dn), W), 4294967296.0));
}
__global__ void k_row_wsum(int64_t rows, const int64_t *__restrict__ off, const float *__restrict__ w, double *wsum, RowMeta *meta) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += n_warps) {
    const int64_t o = off[r], n = off[r + 1] - o;
    double s = 0.0;
    for (int64_t k = lane; k < n; k += 32) s = __dadd_rn(s, (double)w[o + k]);
    for (int d = 16; d > 0; d >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, d));
    if (lane == 0) { wsum[r] = s; if (meta && n > 0) meta[r].w_sum = s; }
  }
}
__global__ void k_alias_classify(int64_t nnz, const uint32_t *__restrict__ row_of, const int64_t *__restrict__ off, const float *__restrict__ w,
                                 const double *__restrict__ wsum, unsigned long long *exc, unsigned long long *def, uint8_t *heavy) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r = row_of[i];
    const uint64_t t = alias_scaled(w[i], (double)(off[r + 1] - off[r]), wsum[r]);
    const bool h = t >= 4294967296ULL;
    heavy[i] = h ? 1 : 0;
    exc[i] = h ? t - 4294967296ULL : 0ULL;
    def[i] = h ? 0ULL : 4294967296ULL - t;
  }
}
__global__ void k_alias_totals(int64_t nnz, const uint8_t *__restrict__ heavy, uint32_t *pos_h, uint32_t *pos_l) {
  pos_h[nnz] = pos_h[nnz - 1] + (heavy[nnz - 1] ? 1u : 0u);
  pos_l[nnz] = pos_l[nnz - 1] + (heavy[nnz - 1] ? 0u : 1u);
}
struct IsHeavy { __host__ __device__ uint32_t operator()(uint8_t h) const { return h ? 1u : 0u; } };
struct IsLight { __host__ __device__ uint32_t operator()(uint8_t h) const { return h ? 0u : 1u; } };
// pos_h / pos_l: exclusive counts of heavy / light entries before entry i (over the whole array, [nnz + 1])
__global__ void k_alias_lists(int64_t nnz, const uint8_t *__restrict__ heavy, const uint32_t *__restrict__ pos_h, const uint32_t *__restrict__ pos_l,
                              uint32_t *list_h, uint32_t *list_l) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    if (heavy[i]) list_h[pos_h[i]] = (uint32_t)i; else list_l[pos_l[i]] = (uint32_t)i;
  }
}
__global__ void k_alias_assign(int64_t nnz, const uint32_t *__restrict__ row_of, const int64_t *__restrict__ off, const int32_t *__restrict__ col,
                               const float *__restrict__ w, const double *__restrict__ wsum, const uint8_t *__restrict__ heavy,
                               const unsigned long long *__restrict__ E, const unsigned long long *__restrict__ D,
                               const uint32_t *__restrict__ pos_h, const uint32_t *__restrict__ pos_l, const uint32_t *__restrict__ list_h,
                               const uint32_t *__restrict__ list_l, AliasSlot *slot) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r = row_of[i];
    const int64_t o = off[r], n = off[r + 1] - o;
    AliasSlot s;
    s.thr = 0xFFFFFFFFu; s.own = col[i]; s.alias_vertex = col[i]; s.alias_index = (uint32_t)(i - o);
    const uint32_t hb = pos_h[o], he = pos_h[o + n], lb = pos_l[o], le = pos_l[o + n];     // the row's heavy / light items in the lists
    if (!heavy[i]) {
      const uint64_t t = alias_scaled(w[i], (double)n, wsum[r]);
      const unsigned long long d_before = D[i] - (4294967296ULL - t);      // D_m: the deficit of the light items before this one
      uint32_t lo = hb, hi = he;                                            // first heavy k with E_k >= D_m
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (E[list_h[mid]] >= d_before) hi = mid; else lo = mid + 1; }
      if (lo < he) { const uint32_t j = list_h[lo]; s.thr = (uint32_t)t; s.alias_vertex = col[j]; s.alias_index = (uint32_t)(j - o); }
    } else {
      const uint32_t x = pos_h[i];
      if (x + 1 < he) {                                                     // a next heavy item exists
        const unsigned long long e_k = E[i];
        uint32_t lo = lb, hi = le;                                          // first m with D_m > E_k
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (D[list_l[mid]] > e_k) hi = mid; else lo = mid + 1; }
        if (lo < le) {
          const uint32_t j2 = list_h[x + 1];
          s.thr = (uint32_t)(4294967296ULL + e_k - D[list_l[lo]]);
          s.alias_vertex = col[j2]; s.alias_index = (uint32_t)(j2 - o);
        }
      }
    }
    slot[i] = s;
  }
}

// Weighted alias-fold: bundle weight of every entry (double sum, in row order, over the run of equal neighbours) ...
__global__ void k_bundle_weights(int64_t nnz, const uint32_t *__restrict__ row_of, const int32_t *__restrict__ col,
                                 const int64_t *__restrict__ off, const float *__restrict__ w, double *wb) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lo = off[row_of[i]], hi = off[row_of[i] + 1];
    const int32_t x = col[i];
    int64_t a = i, b = i;
    while (a > lo && col[a - 1] == x) a--;
    while (b + 1 < hi && col[b + 1] == x) b++;
    double sum = 0.0;
    for (int64_t t = a; t <= b; ++t) sum = __dadd_rn(sum, (double)w[t]);
    wb[i] = sum;
  }
}
// ... and the 32-byte slots that carry it for both outcomes of the Vose coin
__global__ void k_slots_w(int64_t nnz, const uint32_t *__restrict__ row_of, const int64_t *__restrict__ off,
                          const AliasSlot *__restrict__ slot, const double *__restrict__ wb, AliasSlotW *out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const AliasSlot s = slot[i];
    AliasSlotW o;
    o.thr = s.thr; o.own = s.own; o.alias_vertex = s.alias_vertex; o.alias_index = s.alias_index;
    o.wb_own = wb[i];
    o.wb_alias = wb[off[row_of[i]] + (int64_t)s.alias_index];
    out[i] = o;
  }
}

// ---- per-row neighbour hash sets + packed row descriptors ----
__global__ void k_row_meta(int64_t rows, const int64_t *__restrict__ off, RowMeta *meta) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    RowMeta m;
    m.off = off[r]; m.deg = (uint32_t)(off[r + 1] - off[r]);
    m.hoff = srw_hash_first(m.off); m.nb = srw_hash_buckets(m.off, m.deg); m.w_sum = (double)m.deg;
    meta[r] = m;
  }
}
// 16-byte neighbour entries: (x, deg(x), off(x) inside owner(x)'s arrays, owner(x), multiplicity of x in this row).
// Unsharded: goff == nullptr, owner 0, offsets from `off`.  Sharded: `off` holds this shard's local offsets
// (row_of is a local row), goff the global ones, sb the vertex ranges and the global offset of every range.
struct ShardBounds {
  int world;
  int64_t first[SRW_MAX_SHARDS + 1];   // first rank of every shard
  int64_t base[SRW_MAX_SHARDS];        // global offset of the shard's first entry
};
__global__ void k_nbr_entries(int64_t nnz, const uint32_t *__restrict__ row_of, const int32_t *__restrict__ col,
                              const int64_t *__restrict__ off, const int64_t *__restrict__ goff, ShardBounds sb,
                              const MigExt *__restrict__ ext, const uint8_t *__restrict__ own, NbrEntry *ent, int *overflow) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lo = off[row_of[i]], hi = off[row_of[i] + 1];
    const int32_t x = col[i];
    int64_t a = i, b = i;
    while (a > lo && col[a - 1] == x) a--;
    while (b + 1 < hi && col[b + 1] == x) b++;
    const int64_t mult = b - a + 1;
    int64_t xo, xd;
    uint32_t owner = 0;
    if (ext) {                       // VCut shard map
      owner = own[x]; xo = ext[x].off; xd = ext[x].deg;
    } else if (goff) {
      while ((int)owner + 1 < sb.world && (int64_t)x >= sb.first[owner + 1]) owner++;
      xo = goff[x] - sb.base[owner];
      xd = goff[x + 1] - goff[x];
    } else {
      xo = off[x];
      xd = off[x + 1] - xo;
    }
    if (mult >= (1 << 24) || xo >= ((int64_t)1 << 32)) *overflow = 1;
    NbrEntry e;
    e.x = x; e.deg = (uint32_t)xd; e.off_lo = (uint32_t)xo; e.off_hi_mult = owner | ((uint32_t)mult << 8);
    ent[i] = e;
  }
}
// Insert x into the hash set of a row (buckets [hoff, hoff + nb), 8 slots each, filled from slot 0).  The bucket is READ first
// (one 32-byte L2 load) and the compare-and-swap goes straight to its first free slot: ~1 load + ~1 atomic per entry instead of a
// chain of ~4 dependent atomics through the occupied slots (the build's top kernel at RMAT-26: 0.40 s).
__device__ __forceinline__ void hash_set_insert(int32_t *hash, int64_t hoff, uint32_t nb, int32_t x) {
  uint32_t b = __umulhi(srw_hash32((uint32_t)x), nb);
  for (;;) {
    int32_t *bucket = hash + (hoff + b) * 8;
    const int4 q0 = __ldcg(reinterpret_cast<const int4 *>(bucket)), q1 = __ldcg(reinterpret_cast<const int4 *>(bucket) + 1);
    const int32_t v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    int s = 8;
    bool dup = false;
#pragma unroll
    for (int k = 7; k >= 0; --k) { dup |= v[k] == x; if (v[k] == -1) s = k; }      // s = first free slot
    if (dup) return;                                   // a parallel edge already did
    for (; s < 8; ++s) {                               // (slots are never emptied: a bucket that looked full is full)
      const int32_t old = atomicCAS(bucket + s, -1, x);
      if (old == -1 || old == x) return;               // inserted, or a parallel edge did meanwhile
    }
    b = b + 1 == nb ? 0 : b + 1;
  }
}
// one thread per adjacency entry (row keys come from the sort that produced d_col)
__global__ void k_hash_insert(int64_t nnz, const uint32_t *__restrict__ row_of, const int32_t *__restrict__ col,
                              const RowMeta *__restrict__ meta, int32_t *hash) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const RowMeta m = meta[row_of[i]];
    if (m.nb == 0) continue;
    hash_set_insert(hash, m.hoff, m.nb, col[i]);
  }
}

// id-space variant of the two kernels above: hash sets of ORIGINAL ids, entries relabelled in place
__global__ void k_hash_insert_ids(int64_t nnz, const uint32_t *__restrict__ row_of, const int32_t *__restrict__ col,
                                  const RowMeta *__restrict__ meta, const int32_t *__restrict__ vids, int32_t *hash) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const RowMeta m = meta[row_of[i]];
    if (m.nb == 0) continue;
    hash_set_insert(hash, m.hoff, m.nb, vids[col[i]]);
  }
}
__global__ void k_ent_relabel(int64_t nnz, const int32_t *__restrict__ vids, NbrEntry *ent) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) ent[i].x = vids[ent[i].x];
}

struct CastU32ToI64 {
  __host__ __device__ int64_t operator()(uint32_t x) const { return (int64_t)x; }
};

int bits_for(int64_t nv) {
  int b = 1;
  while (((int64_t)1 << b) < nv) b++;
  return b;
}

srw_status sort_pairs(uint32_t *&k_in, uint32_t *&k_out, uint32_t *&v_in, uint32_t *&v_out, int64_t n, int end_bit) {
  cub::DoubleBuffer<uint32_t> dk(k_in, k_out), dv(v_in, v_out);
  size_t tmp_bytes = 0;
  SRW_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n, 0, end_bit));
  DevBuf tmp;
  SRW_CUDA(tmp.alloc(tmp_bytes));
  SRW_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, n, 0, end_bit));
  SRW_CUDA(cudaDeviceSynchronize());
  k_in = dk.Current(); k_out = dk.Alternate();
  v_in = dv.Current(); v_out = dv.Alternate();
  return SRW_OK;
}

template <class T>
srw_status reduce_minmax(const T *d, int64_t n, T *mn, T *mx) {
  DevBuf out, tmp;
  SRW_CUDA(out.alloc(2 * sizeof(T)));
  size_t b1 = 0, b2 = 0;
  SRW_CUDA(cub::DeviceReduce::Min(nullptr, b1, d, out.as<T>(), n));
  SRW_CUDA(cub::DeviceReduce::Max(nullptr, b2, d, out.as<T>() + 1, n));
  SRW_CUDA(tmp.alloc(b1 > b2 ? b1 : b2));
  SRW_CUDA(cub::DeviceReduce::Min(tmp.p, b1, d, out.as<T>(), n));
  SRW_CUDA(cub::DeviceReduce::Max(tmp.p, b2, d, out.as<T>() + 1, n));
  T h[2];
  SRW_CUDA(cudaMemcpy(h, out.p, 2 * sizeof(T), cudaMemcpyDeviceToHost));
  *mn = h[0]; *mx = h[1];
  return SRW_OK;
}

}  // namespace

void srw_alias_thresholds(double p, double q, uint64_t *t_ret, uint64_t *t_common, uint64_t *t_far) {
  // accept a proposal x with probability f(x)/M, f = 1/p (x == prev), 1 (x in N(prev)), 1/q (else):
  // the RS:33-41 bias rule; p and q narrowed to float32 first as in RW:112-113.
  const double inv_p = 1.0 / (double)(float)p, inv_q = 1.0 / (double)(float)q;
  double M = inv_p > 1.0 ? inv_p : 1.0;
  if (inv_q > M) M = inv_q;
  auto thr = [M](double f) -> uint64_t { return f >= M ? 4294967296ULL : (uint64_t)((f / M) * 4294967296.0); };
  *t_ret = thr(inv_p);
  *t_common = thr(1.0);
  *t_far = thr(inv_q);
}

static srw_status build_impl(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w, const int32_t *d_pid,
                             int directed, unsigned flags, int64_t n_extra, const int32_t *d_extra, srw_graph *g, int shard_rank = 0,
                             int shard_world = 1, bool vcut = false, double hub_fraction = 0.0) {
  g->directed = directed != 0;
  g->flags = flags;
  SRW_CUDA(cudaGetDevice(&g->device));
  // build-time breakdown (srw_graph_build_profile): host clock around device-synchronised phases
  auto t_last = std::chrono::steady_clock::now();
  g->build_profile = "{";
  auto phase = [&](const char *name) {
    cudaDeviceSynchronize();
    const auto now = std::chrono::steady_clock::now();
    char buf[96];
    snprintf(buf, sizeof buf, "%s\"%s\": %.3f", g->build_profile.size() > 1 ? ", " : "", name,
             std::chrono::duration<double, std::milli>(now - t_last).count());
    g->build_profile += buf;
    t_last = now;
  };
  struct CloseProfile { std::string &s; ~CloseProfile() { s += "}"; } } close_profile{g->build_profile};
  if (n + n_extra <= 0) {  // empty graph
    SRW_CUDA(cudaMalloc(&g->d_off, sizeof(int64_t)));
    SRW_CUDA(cudaMemset(g->d_off, 0, sizeof(int64_t)));
    return SRW_OK;
  }
  int64_t nnz = directed ? n : 2 * n;   // global; becomes this shard's count below
  // an entry IS its (row, column) pair when there are no weights to carry and no appearance order to keep (K2 below)
  const bool keys_only = !d_w && !(flags & SRW_BUILD_EXACT);
  // Row offsets inside the neighbour entries are 32-bit PER SHARD: one handle holds at most 2^32 - 1 adjacency entries, a graph
  // beyond that (RMAT-27: 2^32) is sharded.  (The sharded build numbers its entries globally with 32 bits only to restore the
  // appearance order, which the keys-only path does not need.)
  if (nnz >= ((int64_t)1 << 32) && !(shard_world > 1 && keys_only)) {
    srw_set_error("more than 2^32-1 adjacency entries in one handle are not supported: shard the graph (unweighted, SRW_BUILD_ALIAS)");
    return SRW_ERR_UNSUPPORTED;
  }

  // ---- id range and presence bitmap ----
  int32_t mn = INT32_MAX, mx = INT32_MIN, a, b;
  if (n > 0) {
    SRW_TRY(reduce_minmax(d_src, n, &a, &b)); mn = a < mn ? a : mn; mx = b > mx ? b : mx;
    SRW_TRY(reduce_minmax(d_dst, n, &a, &b)); mn = a < mn ? a : mn; mx = b > mx ? b : mx;
  }
  if (n_extra > 0) { SRW_TRY(reduce_minmax(d_extra, n_extra, &a, &b)); mn = a < mn ? a : mn; mx = b > mx ? b : mx; }
  g->id_min = mn;
  const uint64_t range = (uint64_t)((int64_t)mx - (int64_t)mn) + 1;
  const uint64_t words = (range + 31) / 32;
  g->id_words = words;
  SRW_CUDA(cudaMalloc(&g->d_bitmap, words * 4));
  SRW_CUDA(cudaMalloc(&g->d_wordrank, (words + 1) * 4));
  SRW_CUDA(cudaMemset(g->d_bitmap, 0, words * 4));
  const unsigned gmax = 148 * 16;
  auto grid = [&](int64_t m) { unsigned x = grid_for(m); return x > gmax ? gmax : x; };
  if (n > 0) {
    k_mark<<<grid(n), kThreads>>>(n, d_src, mn, g->d_bitmap);
    k_mark<<<grid(n), kThreads>>>(n, d_dst, mn, g->d_bitmap);
  }
  if (n_extra > 0) k_mark<<<grid(n_extra), kThreads>>>(n_extra, d_extra, mn, g->d_bitmap);
  {
    DevBuf cnt, tmp;
    SRW_CUDA(cnt.alloc((words + 1) * 4));
    SRW_CUDA(cudaMemset(cnt.p, 0, (words + 1) * 4));
    k_popc<<<grid((int64_t)words), kThreads>>>(words, g->d_bitmap, cnt.as<uint32_t>());
    size_t tb = 0;
    SRW_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.as<uint32_t>(), g->d_wordrank, (int64_t)(words + 1)));
    SRW_CUDA(tmp.alloc(tb));
    SRW_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.as<uint32_t>(), g->d_wordrank, (int64_t)(words + 1)));
    uint32_t nv32 = 0;
    SRW_CUDA(cudaMemcpy(&nv32, g->d_wordrank + words, 4, cudaMemcpyDeviceToHost));
    g->nv = nv32;
  }
  const int64_t nv = g->nv;
  g->nnz = nnz;
  SRW_CUDA(cudaMalloc(&g->d_vids, (size_t)nv * 4));
  k_vids<<<grid((int64_t)words), kThreads>>>(words, g->d_bitmap, g->d_wordrank, mn, g->d_vids);
  phase("id_bitmap_rank");

  // ---- adjacency entries + degrees -> row offsets ----
  const bool sharded = shard_world > 1;
  DevBuf ent_row, ent_col, ent_gidx, deg;
  SRW_CUDA(deg.alloc((size_t)(nv + 1) * 4));
  SRW_CUDA(cudaMemset(deg.p, 0, (size_t)(nv + 1) * 4));
  auto scan_degrees = [&](int64_t *d_out) -> srw_status {
    cub::TransformInputIterator<int64_t, CastU32ToI64, uint32_t *> it(deg.as<uint32_t>(), CastU32ToI64());
    size_t tb = 0;
    DevBuf tmp;
    SRW_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, d_out, nv + 1));
    SRW_CUDA(tmp.alloc(tb));
    SRW_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, it, d_out, nv + 1));
    SRW_CUDA(cudaDeviceSynchronize());
    return SRW_OK;
  };
  auto build_vpid = [&]() -> srw_status {   // GM:21,31
    if (!d_pid || g->d_vpid) return SRW_OK;
    DevBuf last;
    SRW_CUDA(last.alloc((size_t)nv * 8));
    SRW_CUDA(cudaMemset(last.p, 0, (size_t)nv * 8));
    k_vpid_last<<<grid(n), kThreads>>>(n, d_src, d_dst, directed, g->d_bitmap, g->d_wordrank, mn, last.as<unsigned long long>());
    SRW_CUDA(cudaMalloc(&g->d_vpid, (size_t)nv * 4));
    k_vpid_fill<<<grid(nv), kThreads>>>(nv, last.as<unsigned long long>(), d_pid, g->d_vpid);
    SRW_CUDA(cudaDeviceSynchronize());
    g->has_pid = true;
    return SRW_OK;
  };
  g->shard_rank = shard_rank; g->shard_world = shard_world;
  g->row_first = 0; g->row_last = nv; g->nnz_global = nnz;
  g->bounds.assign((size_t)shard_world + 1, 0);
  g->bounds[(size_t)shard_world] = nv;
  int64_t nrows = nv;
  DevBuf goff;                 // sharded: global row offsets (kept for the neighbour entries below)
  ShardBounds sb{};
  sb.world = shard_world;
  if (!sharded) {
    SRW_CUDA(ent_row.alloc((size_t)nnz * 4));
    SRW_CUDA(ent_col.alloc((size_t)nnz * 4));
    if (n > 0)
      k_entries<<<grid(n), kThreads>>>(n, d_src, d_dst, directed, g->d_bitmap, g->d_wordrank, mn, ent_row.as<uint32_t>(),
                                       ent_col.as<uint32_t>(), deg.as<uint32_t>());
    SRW_CUDA(cudaMalloc(&g->d_off, (size_t)(nv + 1) * 8));
    SRW_TRY(scan_degrees(g->d_off));
  } else if (vcut) {
    // Table-mapped shards.  owner_true(v): the VCut shard map -- getPartition(v) mod world (GM:66-68 via VRW:126), every rank derives
    // the same map from the partition-id column -- or, without the column, edge-balanced vertex ranges.  Rows whose degree reaches
    // hub_deg are REPLICATED: every shard keeps them in front of its own rows, and their routing owner is kHubOwner ("wherever
    // the walker is"); ranges are balanced on the entries that are not replicated.
    if (n > 0) k_degrees<<<grid(n), kThreads>>>(n, d_src, d_dst, directed, g->d_bitmap, g->d_wordrank, mn, deg.as<uint32_t>());
    SRW_TRY(build_vpid());
    uint32_t hub_deg = 0xFFFFFFFFu;                 // no row reaches it: nothing replicated
    if (hub_fraction > 0.0 && nnz > 0) {
      // the smallest degree threshold whose rows hold at most hub_fraction of the entries: degrees sorted ascending, prefix-summed; one
      // device thread finds the cut (nothing but the 4-byte answer leaves the device)
      DevBuf dk0, dk1, dv0, dv1, pre, ans, tmp;
      SRW_CUDA(dk0.alloc((size_t)nv * 4)); SRW_CUDA(dk1.alloc((size_t)nv * 4)); SRW_CUDA(dv0.alloc((size_t)nv * 4)); SRW_CUDA(dv1.alloc((size_t)nv * 4));
      SRW_CUDA(cudaMemcpy(dk0.p, deg.p, (size_t)nv * 4, cudaMemcpyDeviceToDevice));
      uint32_t *a_in = dk0.as<uint32_t>(), *a_out = dk1.as<uint32_t>(), *b_in = dv0.as<uint32_t>(), *b_out = dv1.as<uint32_t>();
      SRW_TRY(sort_pairs(a_in, a_out, b_in, b_out, nv, 32));        // ascending (values unused)
      SRW_CUDA(pre.alloc((size_t)nv * 8)); SRW_CUDA(ans.alloc(4));
      {
        cub::TransformInputIterator<int64_t, CastU32ToI64, uint32_t *> it(a_in, CastU32ToI64());
        size_t tb = 0;
        SRW_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb, it, pre.as<int64_t>(), nv));
        SRW_CUDA(tmp.alloc(tb));
        SRW_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, tb, it, pre.as<int64_t>(), nv));
      }
      const int64_t budget = (int64_t)(hub_fraction * (double)nnz);
      k_hub_threshold<<<1, 1>>>(nv, a_in, pre.as<int64_t>(), budget, ans.as<uint32_t>());
      SRW_CUDA(cudaMemcpy(&hub_deg, ans.p, 4, cudaMemcpyDeviceToHost));
    }
    DevBuf otrue, key0, key1, val0, val1, pdeg, poff, lrow;
    SRW_CUDA(otrue.alloc((size_t)nv));
    if (d_pid) k_map_owner_pid<<<grid(nv), kThreads>>>(nv, g->d_vpid, shard_world, otrue.as<uint8_t>());
    else {
      // edge-balanced contiguous vertex ranges over the entries that are not replicated (same on every rank)
      DevBuf mdeg, moff;
      SRW_CUDA(mdeg.alloc((size_t)(nv + 1) * 4)); SRW_CUDA(moff.alloc((size_t)(nv + 1) * 8));
      k_map_masked_deg<<<grid(nv + 1), kThreads>>>(nv, deg.as<uint32_t>(), hub_deg, mdeg.as<uint32_t>());
      {
        cub::TransformInputIterator<int64_t, CastU32ToI64, uint32_t *> it(mdeg.as<uint32_t>(), CastU32ToI64());
        size_t tb = 0;
        DevBuf tmp;
        SRW_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, moff.as<int64_t>(), nv + 1));
        SRW_CUDA(tmp.alloc(tb));
        SRW_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, it, moff.as<int64_t>(), nv + 1));
      }
      ShardBoundsLite bl{};
      bl.world = shard_world;
      {
        DevBuf db;
        SRW_CUDA(db.alloc((SRW_MAX_SHARDS + 1) * 8));
        k_range_bounds<<<1, 32>>>(nv, moff.as<int64_t>(), shard_world, db.as<int64_t>());
        int64_t h_b[SRW_MAX_SHARDS + 1];
        SRW_CUDA(cudaMemcpy(h_b, db.p, sizeof h_b, cudaMemcpyDeviceToHost));
        bl.first[0] = 0; bl.first[shard_world] = nv;
        for (int r = 1; r < shard_world; ++r) {
          int64_t bnd = h_b[r];
          if (bnd > nv) bnd = nv;
          if (bnd < bl.first[r - 1]) bnd = bl.first[r - 1];
          bl.first[r] = bnd;
        }
      }
      k_map_owner_ranges<<<grid(nv), kThreads>>>(nv, bl, otrue.as<uint8_t>());
    }
    SRW_CUDA(key0.alloc((size_t)nv * 4)); SRW_CUDA(key1.alloc((size_t)nv * 4)); SRW_CUDA(val0.alloc((size_t)nv * 4)); SRW_CUDA(val1.alloc((size_t)nv * 4));
    SRW_CUDA(cudaMalloc(&g->d_owner, (size_t)nv));
    k_map_keys<<<grid(nv), kThreads>>>(nv, deg.as<uint32_t>(), hub_deg, otrue.as<uint8_t>(), g->d_owner, key0.as<uint32_t>());
    k_iota<<<grid(nv), kThreads>>>(nv, val0.as<uint32_t>());
    uint32_t *ok_in = key0.as<uint32_t>(), *ok_out = key1.as<uint32_t>(), *ov_in = val0.as<uint32_t>(), *ov_out = val1.as<uint32_t>();
    SRW_TRY(sort_pairs(ok_in, ok_out, ov_in, ov_out, nv, bits_for(shard_world + 1) + 1));      // stable: ascending vertex inside a group
    SRW_CUDA(pdeg.alloc((size_t)(nv + 1) * 4));
    SRW_CUDA(cudaMemset(pdeg.p, 0, (size_t)(nv + 1) * 4));
    k_gather_u32<<<grid(nv), kThreads>>>(nv, ov_in, deg.as<uint32_t>(), pdeg.as<uint32_t>());
    SRW_CUDA(poff.alloc((size_t)(nv + 1) * 8));
    {
      cub::TransformInputIterator<int64_t, CastU32ToI64, uint32_t *> it(pdeg.as<uint32_t>(), CastU32ToI64());
      size_t tb = 0;
      DevBuf tmp;
      SRW_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, it, poff.as<int64_t>(), nv + 1));
      SRW_CUDA(tmp.alloc(tb));
      SRW_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, it, poff.as<int64_t>(), nv + 1));
      SRW_CUDA(cudaDeviceSynchronize());
    }
    MapGroups gr{};
    const int n_groups = shard_world + 1;            // group 0 = hubs, group o + 1 = shard o
    {
      DevBuf first;
      SRW_CUDA(first.alloc((SRW_MAX_SHARDS + 2) * 8));
      SRW_CUDA(cudaMemset(first.p, 0xFF, (SRW_MAX_SHARDS + 2) * 8));
      k_map_group_first<<<grid(nv), kThreads>>>(nv, ok_in, first.as<long long>());
      long long h_first[SRW_MAX_SHARDS + 2];
      SRW_CUDA(cudaMemcpy(h_first, first.p, sizeof h_first, cudaMemcpyDeviceToHost));
      gr.first[n_groups] = nv;
      for (int o = n_groups - 1; o >= 0; --o) gr.first[o] = h_first[o] >= 0 ? h_first[o] : gr.first[o + 1];    // empty group
      for (int o = 0; o <= n_groups; ++o) SRW_CUDA(cudaMemcpy(&gr.base[o], poff.as<int64_t>() + gr.first[o], 8, cudaMemcpyDeviceToHost));
    }
    SRW_CUDA(cudaMalloc(&g->d_ext, (size_t)nv * sizeof(MigExt)));
    SRW_CUDA(lrow.alloc((size_t)nv * 4));
    k_map_ext<<<grid(nv), kThreads>>>(nv, ov_in, ok_in, poff.as<int64_t>(), gr, g->d_ext, lrow.as<uint32_t>());
    const int64_t hub_rows = gr.first[1], hub_entries = gr.base[1];
    const int64_t own_rows = gr.first[shard_rank + 2] - gr.first[shard_rank + 1], own_entries = gr.base[shard_rank + 2] - gr.base[shard_rank + 1];
    nrows = hub_rows + own_rows;
    nnz = hub_entries + own_entries;
    g->nnz = nnz; g->vcut = true; g->hub_rows = hub_rows; g->hub_entries = hub_entries; g->hub_deg = hub_deg;
    g->row_first = 0; g->row_last = nrows;            // LOCAL row count: the rows are not a rank range (d_owner / d_ext say which)
    for (int r = 0; r <= shard_world; ++r) g->bounds[(size_t)r] = gr.first[r + 1] - gr.first[1];          // own-row group sizes as a prefix sum
    if (nnz >= ((int64_t)1 << 32)) { srw_set_error("shard %d would hold %lld adjacency entries (limit 2^32 - 1 per shard)", shard_rank, (long long)nnz); return SRW_ERR_UNSUPPORTED; }
    {
      // seeds: the vertices this shard starts walkers for = owner_true(v) == rank (replicated or not), ascending
      DevBuf iota, cnt, tmp;
      SRW_CUDA(iota.alloc((size_t)nv * 4)); SRW_CUDA(cnt.alloc(8));
      k_iota<<<grid(nv), kThreads>>>(nv, iota.as<uint32_t>());
      SRW_CUDA(cudaMalloc(&g->d_lverts, (size_t)nv * 4));
      size_t tb = 0;
      IsOwner pred{otrue.as<uint8_t>(), shard_rank};
      SRW_CUDA(cub::DeviceSelect::If(nullptr, tb, iota.as<uint32_t>(), (uint32_t *)g->d_lverts, cnt.as<long long>(), nv, pred));
      SRW_CUDA(tmp.alloc(tb));
      SRW_CUDA(cub::DeviceSelect::If(tmp.p, tb, iota.as<uint32_t>(), (uint32_t *)g->d_lverts, cnt.as<long long>(), nv, pred));
      long long h_cnt = 0;
      SRW_CUDA(cudaMemcpy(&h_cnt, cnt.p, 8, cudaMemcpyDeviceToHost));
      g->seed_rows = h_cnt;
    }
    SRW_CUDA(cudaMalloc(&g->d_off, (size_t)(nrows + 1) * 8));
    k_map_local_off<<<grid(nrows + 1), kThreads>>>(hub_rows, own_rows, gr.first[shard_rank + 1], gr.base[shard_rank + 1], hub_entries, poff.as<int64_t>(), g->d_off);
    phase("shard_map_tables");
    DevBuf raw_row, raw_col, raw_gidx, cursor, ka, va;
    SRW_CUDA(raw_row.alloc((size_t)nnz * 4)); SRW_CUDA(raw_col.alloc((size_t)nnz * 4)); SRW_CUDA(raw_gidx.alloc((size_t)nnz * 4));
    SRW_CUDA(cursor.alloc(8));
    SRW_CUDA(cudaMemset(cursor.p, 0, 8));
    if (n > 0)
      k_entries_owner<<<grid(n), kThreads>>>(n, d_src, d_dst, directed, g->d_bitmap, g->d_wordrank, mn, g->d_owner, lrow.as<uint32_t>(), shard_rank,
                                             raw_row.as<uint32_t>(), raw_col.as<uint32_t>(), raw_gidx.as<uint32_t>(), cursor.as<unsigned long long>());
    SRW_CUDA(cudaDeviceSynchronize());
    if (nnz > 0 && keys_only) {
      // nothing depends on the appearance order: the compacted entries are used as they are
      ent_row.p = raw_row.release(); ent_col.p = raw_col.release();
    } else if (nnz > 0) {
      SRW_CUDA(ka.alloc((size_t)nnz * 4)); SRW_CUDA(va.alloc((size_t)nnz * 4));
      DevBuf vb;
      SRW_CUDA(vb.alloc((size_t)nnz * 4));
      uint32_t *k_in = raw_gidx.as<uint32_t>(), *k_out = ka.as<uint32_t>(), *v_in = va.as<uint32_t>(), *v_out = vb.as<uint32_t>();
      k_iota<<<grid(nnz), kThreads>>>(nnz, v_in);
      SRW_TRY(sort_pairs(k_in, k_out, v_in, v_out, nnz, 32));
      SRW_CUDA(ent_row.alloc((size_t)nnz * 4)); SRW_CUDA(ent_col.alloc((size_t)nnz * 4)); SRW_CUDA(ent_gidx.alloc((size_t)nnz * 4));
      k_gather_u32<<<grid(nnz), kThreads>>>(nnz, v_in, raw_row.as<uint32_t>(), ent_row.as<uint32_t>());
      k_gather_u32<<<grid(nnz), kThreads>>>(nnz, v_in, raw_col.as<uint32_t>(), ent_col.as<uint32_t>());
      SRW_CUDA(cudaMemcpy(ent_gidx.p, k_in, (size_t)nnz * 4, cudaMemcpyDeviceToDevice));
      SRW_CUDA(cudaDeviceSynchronize());
    }
  } else {
    // global degrees -> global offsets -> edge-balanced contiguous vertex ranges (same on every rank)
    if (n > 0) k_degrees<<<grid(n), kThreads>>>(n, d_src, d_dst, directed, g->d_bitmap, g->d_wordrank, mn, deg.as<uint32_t>());
    SRW_CUDA(goff.alloc((size_t)(nv + 1) * 8));
    SRW_TRY(scan_degrees(goff.as<int64_t>()));
    std::vector<int64_t> h_goff((size_t)nv + 1);
    SRW_CUDA(cudaMemcpy(h_goff.data(), goff.p, (size_t)(nv + 1) * 8, cudaMemcpyDeviceToHost));
    for (int r = 1; r < shard_world; ++r) {
      const int64_t target = (int64_t)((__int128)nnz * r / shard_world);
      g->bounds[(size_t)r] = (int64_t)(std::lower_bound(h_goff.begin(), h_goff.end(), target) - h_goff.begin());
      if (g->bounds[(size_t)r] > nv) g->bounds[(size_t)r] = nv;
      if (g->bounds[(size_t)r] < g->bounds[(size_t)r - 1]) g->bounds[(size_t)r] = g->bounds[(size_t)r - 1];
    }
    for (int r = 0; r <= shard_world; ++r) sb.first[r] = g->bounds[(size_t)r];
    for (int r = 0; r < shard_world; ++r) sb.base[r] = h_goff[(size_t)g->bounds[(size_t)r]];
    g->row_first = g->bounds[(size_t)shard_rank];
    g->row_last = g->bounds[(size_t)shard_rank + 1];
    nrows = g->row_last - g->row_first;
    nnz = h_goff[(size_t)g->row_last] - h_goff[(size_t)g->row_first];     // this shard's entries
    g->nnz = nnz;
    SRW_CUDA(cudaMalloc(&g->d_off, (size_t)(nrows + 1) * 8));
    k_rebase_off<<<grid(nrows + 1), kThreads>>>(nrows, goff.as<int64_t>(), g->row_first, g->d_off);
    DevBuf raw_row, raw_col, raw_gidx, cursor, ka, va;
    SRW_CUDA(raw_row.alloc((size_t)nnz * 4)); SRW_CUDA(raw_col.alloc((size_t)nnz * 4)); SRW_CUDA(raw_gidx.alloc((size_t)nnz * 4));
    SRW_CUDA(cursor.alloc(8));
    SRW_CUDA(cudaMemset(cursor.p, 0, 8));
    if (n > 0)
      k_entries_range<<<grid(n), kThreads>>>(n, d_src, d_dst, directed, g->d_bitmap, g->d_wordrank, mn, (uint32_t)g->row_first,
                                             (uint32_t)g->row_last, raw_row.as<uint32_t>(), raw_col.as<uint32_t>(),
                                             raw_gidx.as<uint32_t>(), cursor.as<unsigned long long>());
    SRW_CUDA(cudaDeviceSynchronize());
    if (nnz > 0 && keys_only) {
      // nothing depends on the appearance order: the compacted entries are used as they are
      ent_row.p = raw_row.release(); ent_col.p = raw_col.release();
    } else if (nnz > 0) {
      // restore file-appearance order: sort the kept entries by their global entry index
      SRW_CUDA(ka.alloc((size_t)nnz * 4)); SRW_CUDA(va.alloc((size_t)nnz * 4));
      DevBuf vb;
      SRW_CUDA(vb.alloc((size_t)nnz * 4));
      uint32_t *k_in = raw_gidx.as<uint32_t>(), *k_out = ka.as<uint32_t>(), *v_in = va.as<uint32_t>(), *v_out = vb.as<uint32_t>();
      k_iota<<<grid(nnz), kThreads>>>(nnz, v_in);
      SRW_TRY(sort_pairs(k_in, k_out, v_in, v_out, nnz, 32));
      SRW_CUDA(ent_row.alloc((size_t)nnz * 4)); SRW_CUDA(ent_col.alloc((size_t)nnz * 4)); SRW_CUDA(ent_gidx.alloc((size_t)nnz * 4));
      k_gather_u32<<<grid(nnz), kThreads>>>(nnz, v_in, raw_row.as<uint32_t>(), ent_row.as<uint32_t>());
      k_gather_u32<<<grid(nnz), kThreads>>>(nnz, v_in, raw_col.as<uint32_t>(), ent_col.as<uint32_t>());
      SRW_CUDA(cudaMemcpy(ent_gidx.p, k_in, (size_t)nnz * 4, cudaMemcpyDeviceToDevice));
      SRW_CUDA(cudaDeviceSynchronize());
    }
  }
  const uint32_t *gidx = sharded ? ent_gidx.as<uint32_t>() : nullptr;
  deg.alloc(0);
  phase(sharded ? "k_entries_range_sort" : "k_entries_scan");
  if ((flags & SRW_BUILD_MIGRATE) && !directed && n > 0) {
    // replicated on every shard: 16 bits per input edge by default (1 byte per adjacency entry; false positives ~0.4 %)
    const char *eb = getenv("SRW_BLOOM_BITS");
    const int64_t bits = eb && atoi(eb) > 0 ? atoi(eb) : 16;
    g->bloom_words = (uint64_t)((n * bits + 63) / 64);
    if (g->bloom_words < 64) g->bloom_words = 64;
    if (g->bloom_words > 0xFFFFFFFFull) g->bloom_words = 0xFFFFFFFFull;     // the probe indexes words with 32 bits (32 GB of filter)
    SRW_CUDA(cudaMalloc(&g->d_bloom, g->bloom_words * 8));
    SRW_CUDA(cudaMemset(g->d_bloom, 0, g->bloom_words * 8));
    k_bloom_insert<<<grid(n), kThreads>>>(n, d_src, d_dst, g->d_bitmap, g->d_wordrank, mn, g->d_bloom, (uint32_t)g->bloom_words);
    SRW_CUDA(cudaDeviceSynchronize());
    phase("k_bloom_insert");
  }

  SRW_TRY(build_vpid());
  if (nnz == 0) return SRW_OK;

  const int rbits = bits_for(nrows), cbits = bits_for(nv);
  const int wshift = directed ? 0 : 1;
  DevBuf kb0, kb1, vb0, vb1;
  SRW_CUDA(kb0.alloc((size_t)nnz * 4));
  if (!keys_only) { SRW_CUDA(kb1.alloc((size_t)nnz * 4)); SRW_CUDA(vb0.alloc((size_t)nnz * 4)); SRW_CUDA(vb1.alloc((size_t)nnz * 4)); }
  uint32_t *k_in = kb0.as<uint32_t>(), *k_out = kb1.as<uint32_t>(), *v_in = vb0.as<uint32_t>(), *v_out = vb1.as<uint32_t>();

  // ---- K1: appearance-order rows = stable sort of the entries by row ----
  if (flags & SRW_BUILD_EXACT) {
    SRW_CUDA(cudaMemcpy(k_in, ent_row.p, (size_t)nnz * 4, cudaMemcpyDeviceToDevice));
    k_iota<<<grid(nnz), kThreads>>>(nnz, v_in);
    SRW_TRY(sort_pairs(k_in, k_out, v_in, v_out, nnz, rbits));
    SRW_CUDA(cudaMalloc(&g->d_col_app, (size_t)nnz * 4));
    SRW_CUDA(cudaMalloc(&g->d_w_app, (size_t)nnz * 4));
    k_gather_u32<<<grid(nnz), kThreads>>>(nnz, v_in, ent_col.as<uint32_t>(), (uint32_t *)g->d_col_app);
    k_gather_w<<<grid(nnz), kThreads>>>(nnz, v_in, gidx, d_w, wshift, g->d_w_app);
    SRW_CUDA(cudaDeviceSynchronize());
    phase("appearance_rows_sort");
  }

  // ---- K2: neighbour-sorted rows ----
  SRW_CUDA(cudaMalloc(&g->d_col, (size_t)nnz * 4));
  if (keys_only) {
    // No weights to carry along: an entry IS its (row, column) pair, so ONE keys-only radix sort of the packed 64-bit key
    // row << cbits | col (rbits + cbits significant bits) replaces the two stable pair sorts + three gathers below
    // (RMAT-26: 0.51 s -> see config.build_ms_per_phase).  Equal keys are identical entries: stability is moot.
    DevBuf p0, p1;
    SRW_CUDA(p0.alloc((size_t)nnz * 8)); SRW_CUDA(p1.alloc((size_t)nnz * 8));
    k_pack_rc<<<grid(nnz), kThreads>>>(nnz, ent_row.as<uint32_t>(), ent_col.as<uint32_t>(), cbits, p0.as<unsigned long long>());
    SRW_CUDA(cudaDeviceSynchronize());
    ent_row.alloc(0); ent_col.alloc(0);
    cub::DoubleBuffer<unsigned long long> dk(p0.as<unsigned long long>(), p1.as<unsigned long long>());
    size_t tb = 0;
    SRW_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, dk, nnz, 0, rbits + cbits));
    DevBuf tmp;
    SRW_CUDA(tmp.alloc(tb));
    SRW_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, dk, nnz, 0, rbits + cbits));
    k_unpack_rc<<<grid(nnz), kThreads>>>(nnz, dk.Current(), cbits, k_in, (uint32_t *)g->d_col);
    SRW_CUDA(cudaDeviceSynchronize());
    v_in = nullptr;                        // no permutation exists on this path (it is only needed to carry weights)
    phase("sorted_rows_1_key_sort");
  } else {
    // stable sort by column, then stable sort by row, carrying the entry index (the weights follow it)
    SRW_CUDA(cudaMemcpy(k_in, ent_col.p, (size_t)nnz * 4, cudaMemcpyDeviceToDevice));
    k_iota<<<grid(nnz), kThreads>>>(nnz, v_in);
    SRW_TRY(sort_pairs(k_in, k_out, v_in, v_out, nnz, cbits));
    k_gather_u32<<<grid(nnz), kThreads>>>(nnz, v_in, ent_row.as<uint32_t>(), k_in);
    SRW_TRY(sort_pairs(k_in, k_out, v_in, v_out, nnz, rbits));
    k_gather_u32<<<grid(nnz), kThreads>>>(nnz, v_in, ent_col.as<uint32_t>(), (uint32_t *)g->d_col);
    SRW_CUDA(cudaDeviceSynchronize());
    ent_row.alloc(0); ent_col.alloc(0);
    phase("sorted_rows_2_sorts");
  }
  (void)ent_gidx;

  bool weighted_graph = false;     // any weight != 1.0f (RS semantics are weight-relative; 1.0f rows need no table)
  if (d_w) {
    DevBuf flag;
    SRW_CUDA(flag.alloc(4));
    SRW_CUDA(cudaMemset(flag.p, 0, 4));
    k_any_non_unit<<<grid(n), kThreads>>>(n, d_w, flag.as<int>());
    int h = 0;
    SRW_CUDA(cudaMemcpy(&h, flag.p, 4, cudaMemcpyDeviceToHost));
    weighted_graph = h != 0;
  }

  // ---- packed row descriptors + neighbour hash sets (membership test of the alias sampler) ----
  // SRW_BUILD_LEAN: only what the id-space alias-fold kernel reads survives the build (d_off, d_vids, d_ent, d_hash_id)
  const bool ids_ok = !sharded && !(flags & SRW_BUILD_MIGRATE) /* the migrating walk routes by rank */ && g->id_min >= 0 /* -1 marks an empty hash slot */ &&
                      !(getenv("SRW_FOLD_IDS") && atoi(getenv("SRW_FOLD_IDS")) == 0);
  const bool lean = (flags & SRW_BUILD_LEAN) && (flags & SRW_BUILD_ALIAS) && !(flags & SRW_BUILD_EXACT) && !weighted_graph && !directed && ids_ok;
  if (flags & SRW_BUILD_ALIAS) {
    g->hash_buckets = (nnz >> 2) + 1;
    SRW_CUDA(cudaMalloc(&g->d_meta, (size_t)(nrows ? nrows : 1) * sizeof(RowMeta)));
    k_row_meta<<<grid(nrows), kThreads>>>(nrows, g->d_off, g->d_meta);
    if (!lean) {
      SRW_CUDA(cudaMalloc(&g->d_hash, (size_t)g->hash_buckets * 32));
      SRW_CUDA(cudaMemset(g->d_hash, 0xFF, (size_t)g->hash_buckets * 32));
      k_hash_insert<<<grid(nnz), kThreads>>>(nnz, k_in, g->d_col, g->d_meta, g->d_hash);   // k_in = row keys in d_col order
      SRW_CUDA(cudaDeviceSynchronize());
      phase("k_hash_insert");
    }
    if (!weighted_graph) {
      DevBuf ovf;
      SRW_CUDA(ovf.alloc(4));
      SRW_CUDA(cudaMemset(ovf.p, 0, 4));
      SRW_CUDA(cudaMalloc(&g->d_ent, (size_t)nnz * sizeof(NbrEntry)));
      k_nbr_entries<<<grid(nnz), kThreads>>>(nnz, k_in, g->d_col, g->d_off, sharded && !vcut ? goff.as<int64_t>() : nullptr, sb,
                                             vcut ? g->d_ext : nullptr, vcut ? g->d_owner : nullptr, g->d_ent, ovf.as<int>());
      int h = 0;
      SRW_CUDA(cudaMemcpy(&h, ovf.p, 4, cudaMemcpyDeviceToHost));
      if (h) { cudaFree(g->d_ent); g->d_ent = nullptr; }    // absurd multiplicities: fold sampler unavailable
      phase("k_nbr_entries");
      if (lean && !g->d_ent) { srw_set_error("SRW_BUILD_LEAN: an edge multiplicity or a row offset does not fit the neighbour entry; build without the flag"); return SRW_ERR_UNSUPPORTED; }
      // ID SPACE (default; SRW_FOLD_IDS=0 at build time keeps the entries in rank space for A/B runs): the fold kernel treats a neighbour as an opaque label -- it compares it with prev,
      // hashes it and appends it to the path -- so the entries can carry ORIGINAL VERTEX IDS and the walk emits ids directly:
      // the rank -> id pass over the path matrix disappears (measured at RMAT-26, profiles/r1_fold_ids_ab.jsonl: 186.8 -> 174.2 ms per round).  Costs a second, id-labelled copy
      // of the hash sets (the rank-labelled one serves the other kernels).  Ranks ascend with ids, so rows stay sorted.
      if (g->d_ent && ids_ok) {
        if (cudaMalloc(&g->d_hash_id, (size_t)g->hash_buckets * 32) == cudaSuccess) {
          SRW_CUDA(cudaMemset(g->d_hash_id, 0xFF, (size_t)g->hash_buckets * 32));
          k_hash_insert_ids<<<grid(nnz), kThreads>>>(nnz, k_in, g->d_col, g->d_meta, g->d_vids, g->d_hash_id);
          k_ent_relabel<<<grid(nnz), kThreads>>>(nnz, g->d_vids, g->d_ent);
          SRW_CUDA(cudaDeviceSynchronize());
          g->ent_ids = true;
          phase("k_hash_insert_ids");
        } else {
          cudaGetLastError();
          g->d_hash_id = nullptr;
          if (lean) { srw_set_error("SRW_BUILD_LEAN: not enough device memory for the hash sets"); return SRW_ERR_CUDA; }
        }
      }
      if (lean) {                       // the fold kernel reads d_off, d_vids, d_ent and d_hash_id: nothing else is kept
        cudaFree(g->d_col); g->d_col = nullptr;
        cudaFree(g->d_meta); g->d_meta = nullptr;
        g->lean = true;
      }
    }
  }

  // ---- K3: Vose slots over the sorted rows (weighted graphs only) ----
  if ((flags & SRW_BUILD_ALIAS) && weighted_graph) {
    {
      DevBuf ws;
      SRW_CUDA(ws.alloc((size_t)nnz * 4));
      k_gather_w<<<grid(nnz), kThreads>>>(nnz, v_in, gidx, d_w, wshift, ws.as<float>());
      SRW_CUDA(cudaMalloc(&g->d_slot, (size_t)nnz * sizeof(AliasSlot)));
      {
        DevBuf wsum, exc, def, hv, E, D, pos_h, pos_l, list, tmp;
        SRW_CUDA(wsum.alloc((size_t)(nrows ? nrows : 1) * 8));
        k_row_wsum<<<grid(nrows * 32), kThreads>>>(nrows, g->d_off, ws.as<float>(), wsum.as<double>(), g->d_meta);
        SRW_CUDA(exc.alloc((size_t)nnz * 8)); SRW_CUDA(def.alloc((size_t)nnz * 8)); SRW_CUDA(hv.alloc((size_t)nnz));
        k_alias_classify<<<grid(nnz), kThreads>>>(nnz, k_in, g->d_off, ws.as<float>(), wsum.as<double>(), exc.as<unsigned long long>(),
                                                  def.as<unsigned long long>(), hv.as<uint8_t>());
        SRW_CUDA(E.alloc((size_t)nnz * 8)); SRW_CUDA(D.alloc((size_t)nnz * 8));
        SRW_CUDA(pos_h.alloc((size_t)(nnz + 1) * 4)); SRW_CUDA(pos_l.alloc((size_t)(nnz + 1) * 4));
        size_t tb = 0, tb2 = 0;
        SRW_CUDA(cub::DeviceScan::InclusiveSumByKey(nullptr, tb, k_in, exc.as<unsigned long long>(), E.as<unsigned long long>(), nnz));
        cub::TransformInputIterator<uint32_t, IsHeavy, uint8_t *> it_h(hv.as<uint8_t>(), IsHeavy());
        cub::TransformInputIterator<uint32_t, IsLight, uint8_t *> it_l(hv.as<uint8_t>(), IsLight());
        SRW_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, it_h, pos_h.as<uint32_t>(), nnz));
        SRW_CUDA(tmp.alloc(tb > tb2 ? tb : tb2));
        SRW_CUDA(cub::DeviceScan::InclusiveSumByKey(tmp.p, tb, k_in, exc.as<unsigned long long>(), E.as<unsigned long long>(), nnz));
        SRW_CUDA(cub::DeviceScan::InclusiveSumByKey(tmp.p, tb, k_in, def.as<unsigned long long>(), D.as<unsigned long long>(), nnz));
        SRW_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, it_h, pos_h.as<uint32_t>(), nnz));
        SRW_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, it_l, pos_l.as<uint32_t>(), nnz));
        k_alias_totals<<<1, 1>>>(nnz, hv.as<uint8_t>(), pos_h.as<uint32_t>(), pos_l.as<uint32_t>());
        exc.alloc(0); def.alloc(0);
        SRW_CUDA(list.alloc((size_t)nnz * 4 + 8));         // heavy list from the front, light list behind it
        uint32_t n_heavy = 0;
        SRW_CUDA(cudaMemcpy(&n_heavy, pos_h.as<uint32_t>() + nnz, 4, cudaMemcpyDeviceToHost));
        uint32_t *list_h = list.as<uint32_t>(), *list_l = list.as<uint32_t>() + n_heavy;
        k_alias_lists<<<grid(nnz), kThreads>>>(nnz, hv.as<uint8_t>(), pos_h.as<uint32_t>(), pos_l.as<uint32_t>(), list_h, list_l);
        k_alias_assign<<<grid(nnz), kThreads>>>(nnz, k_in, g->d_off, g->d_col, ws.as<float>(), wsum.as<double>(), hv.as<uint8_t>(),
                                                E.as<unsigned long long>(), D.as<unsigned long long>(), pos_h.as<uint32_t>(), pos_l.as<uint32_t>(),
                                                list_h, list_l, g->d_slot);
        SRW_CUDA(cudaDeviceSynchronize());
      }
      g->has_alias = true;
      phase("k_alias_parallel_vose");
      // weighted alias-fold (undirected, unsharded): 32-byte slots with the bundle weights.  48 bytes per entry in all:
      // HBM capacity is spent to keep a proposal at ONE memory request.
      if (!directed && !sharded && g->d_meta && !getenv("SRW_NO_WFOLD")) {
        DevBuf wb;
        SRW_CUDA(wb.alloc((size_t)nnz * 8));
        k_bundle_weights<<<grid(nnz), kThreads>>>(nnz, k_in, g->d_col, g->d_off, ws.as<float>(), wb.as<double>());
        if (cudaMalloc(&g->d_slotw, (size_t)nnz * sizeof(AliasSlotW)) == cudaSuccess) {
          k_slots_w<<<grid(nnz), kThreads>>>(nnz, k_in, g->d_off, g->d_slot, wb.as<double>(), g->d_slotw);
          SRW_CUDA(cudaDeviceSynchronize());
          phase("bundle_weights_wide_slots");
        } else {
          cudaGetLastError();            // not enough HBM for the wide slots: the classic alias sampler still runs
          g->d_slotw = nullptr;
        }
      }
    }
  }
  SRW_CUDA(cudaGetLastError());
  if (vcut) g->device_bytes += nv * 13;
  g->device_bytes += (int64_t)(words * 8 + (size_t)nv * 12 + (g->d_col ? (size_t)nnz * 4 : 0)) + (g->d_col_app ? nnz * 8 : 0) +
                    (g->d_slot ? nnz * 16 : 0) + (g->d_slotw ? nnz * 32 : 0) + (g->d_vpid ? nv * 4 : 0) + (g->d_meta ? nrows * 32 : 0) + (g->d_hash ? g->hash_buckets * 32 : 0) + (g->d_hash_id ? g->hash_buckets * 32 : 0) + (g->d_ent ? nnz * 16 : 0) + (int64_t)g->bloom_words * 8;
  return SRW_OK;
}

namespace {
__global__ void k_ent_ranks(int64_t n, const NbrEntry *__restrict__ ent, const uint32_t *__restrict__ bitmap,
                            const uint32_t *__restrict__ wordrank, int32_t id_min, int32_t *out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = rank_of(bitmap, wordrank, id_min, ent[i].x);
}
}  // namespace
srw_status srw_ent_ranks_to_host(const srw_graph *g, int32_t *h_col) {
  SRW_CUDA(cudaSetDevice(g->device));
  const int64_t chunk = (int64_t)1 << 26;
  DevBuf tmp;
  SRW_CUDA(tmp.alloc((size_t)std::min<int64_t>(chunk, g->nnz) * 4));
  for (int64_t at = 0; at < g->nnz; at += chunk) {
    const int64_t m = std::min<int64_t>(chunk, g->nnz - at);
    k_ent_ranks<<<grid_for(m) > 148 * 16 ? 148 * 16 : grid_for(m), kThreads>>>(m, g->d_ent + at, g->d_bitmap, g->d_wordrank, g->id_min, tmp.as<int32_t>());
    SRW_CUDA(cudaMemcpy(h_col + at, tmp.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
  }
  return SRW_OK;
}

srw_status srw_build_graph_device(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w,
                                  const int32_t *d_pid, int directed, unsigned flags, srw_graph **out) {
  SRW_TRY(srw_require_device());
  if (n < 0 || !out || (n > 0 && (!d_src || !d_dst))) { srw_set_error("srw_graph_from_device_edges: bad argument"); return SRW_ERR_ARG; }
  srw_graph *g = new srw_graph();
  srw_status s = build_impl(n, d_src, d_dst, d_w, d_pid, directed, flags ? flags : SRW_BUILD_ALL, 0, nullptr, g);
  if (s != SRW_OK) { srw_graph_free(g); return s; }
  *out = g;
  return SRW_OK;
}

srw_status srw_build_graph_rows(int64_t n_rows, const int32_t *h_vids, const int64_t *h_row_off, const int64_t *h_row_len,
                                const int32_t *h_dst, const int32_t *h_pid, const float *h_w, unsigned flags,
                                srw_graph **out) {
  SRW_TRY(srw_require_device());
  // rows -> directed edge list in row order; added vertices without neighbours become extra ids
  std::vector<int32_t> src, dst, pid;
  std::vector<float> w;
  for (int64_t r = 0; r < n_rows; ++r)
    for (int64_t k = 0; k < h_row_len[r]; ++k) {
      src.push_back(h_vids[r]);
      dst.push_back(h_dst[h_row_off[r] + k]);
      w.push_back(h_w ? h_w[h_row_off[r] + k] : 1.0f);
      if (h_pid) pid.push_back(h_pid[h_row_off[r] + k]);
    }
  const int64_t n = (int64_t)src.size();
  DevBuf ds, dd, dw, dp, dx;
  SRW_CUDA(ds.alloc(n * 4)); SRW_CUDA(dd.alloc(n * 4)); SRW_CUDA(dw.alloc(n * 4)); SRW_CUDA(dx.alloc(n_rows * 4));
  SRW_CUDA(cudaMemcpy(ds.p, src.data(), n * 4, cudaMemcpyHostToDevice));
  SRW_CUDA(cudaMemcpy(dd.p, dst.data(), n * 4, cudaMemcpyHostToDevice));
  SRW_CUDA(cudaMemcpy(dw.p, w.data(), n * 4, cudaMemcpyHostToDevice));
  SRW_CUDA(cudaMemcpy(dx.p, h_vids, n_rows * 4, cudaMemcpyHostToDevice));
  if (h_pid) { SRW_CUDA(dp.alloc(n * 4)); SRW_CUDA(cudaMemcpy(dp.p, pid.data(), n * 4, cudaMemcpyHostToDevice)); }
  srw_graph *g = new srw_graph();
  srw_status s = build_impl(n, ds.as<int32_t>(), dd.as<int32_t>(), dw.as<float>(), h_pid ? dp.as<int32_t>() : nullptr, 1,
                            flags ? flags : SRW_BUILD_ALL, n_rows, dx.as<int32_t>(), g);
  if (s != SRW_OK) { srw_graph_free(g); return s; }
  *out = g;
  return SRW_OK;
}

srw_status srw_build_graph_device_sharded(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w, int directed,
                                          unsigned flags, int rank, int world, srw_graph **out, const int32_t *d_pid, double hub_fraction) {
  SRW_TRY(srw_require_device());
  if (n < 0 || !out || (n > 0 && (!d_src || !d_dst)) || world < 1 || world > SRW_MAX_SHARDS || rank < 0 || rank >= world) {
    srw_set_error("srw_graph_from_device_edges_sharded: bad argument");
    return SRW_ERR_ARG;
  }
  srw_graph *g = new srw_graph();
  // d_pid != NULL: the VCut shard map -- owner(v) = getPartition(v) mod world instead of edge-balanced vertex ranges;
  // hub_fraction > 0: the highest-degree rows holding up to that share of the entries are replicated on every shard
  if (!(hub_fraction >= 0.0) || hub_fraction > 0.95) { srw_set_error("hub_fraction must be in [0, 0.95]"); delete g; return SRW_ERR_ARG; }
  const bool mapped = world > 1 && (d_pid != nullptr || hub_fraction > 0.0);
  srw_status s = build_impl(n, d_src, d_dst, d_w, d_pid, directed, flags ? flags : SRW_BUILD_ALIAS, 0, nullptr, g, rank, world, mapped, hub_fraction);
  if (s != SRW_OK) { srw_graph_free(g); return s; }
  g->peer_off[rank] = g->d_off; g->peer_ent[rank] = g->d_ent; g->peer_hash[rank] = g->d_hash;
  g->peer_attached[rank] = true;
  *out = g;
  return SRW_OK;
}
