// Internal structures of libsrw (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/srw.h"

void srw_set_error(const char *fmt, ...);

#define SRW_CUDA(call)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      srw_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return SRW_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)
#define SRW_TRY(call)                 \
  do {                                \
    srw_status s_ = (call);           \
    if (s_ != SRW_OK) return s_;      \
  } while (0)

srw_status srw_require_device();

struct srw_edges {
  std::vector<int32_t> src, dst, pid;
  std::vector<float> w;
  bool has_pid = false;
};

#include "layout.h"

struct srw_graph {
  int device = 0;
  int64_t nv = 0, nnz = 0;
  bool directed = false, has_alias = false, has_pid = false;
  unsigned flags = 0;
  // id <-> rank
  int32_t id_min = 0;
  uint64_t id_words = 0;
  uint32_t *d_bitmap = nullptr, *d_wordrank = nullptr;
  int32_t *d_vids = nullptr;      // [nv] ascending original ids
  int64_t *d_off = nullptr;       // [nv+1]
  int32_t *d_col_app = nullptr;   // [nnz] neighbour ranks, file-appearance order   (SRW_BUILD_EXACT)
  float *d_w_app = nullptr;       // [nnz]
  int32_t *d_col = nullptr;       // [nnz] neighbour ranks, ascending per row (membership + unweighted proposals)
  AliasSlot *d_slot = nullptr;    // [nnz] iff has_alias                              (SRW_BUILD_ALIAS)
  AliasSlotW *d_slotw = nullptr;  // [nnz] weighted, undirected, unsharded graphs: slots + bundle weights (weighted alias-fold)
  int32_t *d_vpid = nullptr;      // [nv] GM:21 vertexPartitionMap (-1 = absent)
  struct RowMeta *d_meta = nullptr;  // [rows] packed row descriptor: one 32-byte load per step   (SRW_BUILD_ALIAS)
  int32_t *d_hash = nullptr;         // per-row neighbour hash sets, 8-slot (32-byte) buckets, -1 = empty
  int64_t hash_buckets = 0;
  NbrEntry *d_ent = nullptr;         // [nnz] unweighted, unsharded graphs (fold sampler)
  int32_t *d_hash_id = nullptr;      // id-space fold (SRW_FOLD_IDS): hash sets of original ids; d_ent[].x then holds ids too
  bool ent_ids = false;
  bool lean = false;                 // SRW_BUILD_LEAN took effect: d_col, d_hash and d_meta were dropped
  std::string build_profile;         // srw_graph_build_profile
  unsigned long long *d_bloom = nullptr;   // SRW_BUILD_MIGRATE: replicated edge filter of the migrating sharded walk (migrate.cuh), 64-bit words
  uint64_t bloom_words = 0;
  int64_t device_bytes = 0;
  mutable std::vector<int32_t> h_vids;  // lazy host copies for the query entry points
  mutable std::vector<int64_t> h_off;
  // vertex-range shard: this handle holds rows [row_first, row_last) of a graph whose
  // ranks run to nv; d_off / d_col / d_slot are indexed by (rank - row_first); neighbour ids stay global
  int shard_rank = 0, shard_world = 1;
  int64_t row_first = 0, row_last = 0, nnz_global = 0;
  std::vector<int64_t> bounds;          // [world+1] first rank of every shard
  // VCut shard map: owner(v) = getPartition(v) mod world; this shard's rows are the vertices d_lverts[0 .. row_last)
  // (row_first = 0, row_last = the LOCAL row count); d_ext / d_owner are replicated on every shard
  bool vcut = false;                    // table-mapped shard: VCut shard map and/or replicated hub rows
  MigExt *d_ext = nullptr;              // [nv] (offset inside owner(v)'s arrays, degree)
  uint8_t *d_owner = nullptr;           // [nv] routing owner; 0xFF = hub row, replicated on every shard
  int32_t *d_lverts = nullptr;          // [seed_rows] vertices this shard starts walkers for, ascending
  int64_t seed_rows = 0, hub_rows = 0, hub_entries = 0;   // local arrays = [hub rows | own rows]
  uint32_t hub_deg = 0xFFFFFFFFu;       // rows of at least this degree are replicated
  struct ShardScratch *scratch = nullptr;
  // peer-gather mode (SURVEY 8(e) "NVLink peer loads"): the row arrays of every shard, addressable from
  // this device -- own pointers for shard_rank, cudaIpcOpenMemHandle mappings (or same-process pointers)
  // for the others.  The walk kernel then reads remote rows directly; no walker ever migrates.
  const int64_t *peer_off[SRW_MAX_SHARDS] = {};
  const NbrEntry *peer_ent[SRW_MAX_SHARDS] = {};
  const int32_t *peer_hash[SRW_MAX_SHARDS] = {};
  bool peer_attached[SRW_MAX_SHARDS] = {};
  bool peer_ipc[SRW_MAX_SHARDS] = {};   // mapping opened with cudaIpcOpenMemHandle (closed on free)
  bool rows_external = false;           // d_off / d_ent / d_hash live in a caller-owned block (srw_shard_rows_relocate)
  // multi-GPU container (multi.cu, `--gpus N` in one process): shards[r] is the vertex-range shard on device r; the container
  // itself holds no device arrays.  `multi` caches the exchange blocks / contexts of the last walk.
  std::vector<srw_graph *> shards;
  struct MultiWalk *multi = nullptr;
};
void srw_multi_free(struct MultiWalk *w);
extern bool g_srw_log_supersteps;     // srw_main: print the reference's per-super-step `Unfinished Walkers: N` line (RW:154)
int64_t srw_last_short_paths();        // paths of this thread's last srw_walk_save that ended at a dead end (RW:115-119 `Zero Neighbors`)
srw_status srw_build_graph_device_multi(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w, int directed,
                                        int num_gpus, srw_graph **out, const int32_t *d_pid = nullptr, double hub_fraction = -1.0);
// rounds [round_first, round_first + n_rounds) over a container graph, delivered on device 0 in walker order
srw_status srw_multi_walk_rounds(const srw_graph *g, const srw_params *p, int64_t round_first, int64_t n_rounds, int32_t *d_paths0,
                                 int32_t *d_lens0, srw_walk_info *info);

srw_status srw_build_graph_device_sharded(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w, int directed,
                                          unsigned flags, int rank, int world, srw_graph **out, const int32_t *d_pid = nullptr,
                                          double hub_fraction = 0.0);
void srw_shard_scratch_free(struct ShardScratch *s);

struct srw_paths {
  int64_t n_paths = 0, n_steps = 0;
  int32_t stride = 0;
  std::vector<int32_t> ids;       // ragged
  std::vector<int64_t> offsets;
};

struct srw_graphmap {
  std::vector<int32_t> vids;            // insertion order
  std::unordered_set<int32_t> seen;
  std::vector<int64_t> row_off;         // per inserted vertex
  std::vector<int64_t> row_len;
  std::vector<int32_t> dst, pid;
  std::vector<float> w;
  bool any_pid = false;
};

// walk.cu
struct WalkLaunch {
  uint64_t walker_first;
  int64_t n_walkers;
  int32_t *d_paths, *d_lens;
  cudaStream_t stream;
};
srw_status srw_walk_launch(const srw_graph *g, const srw_params *p, const WalkLaunch &l);
void srw_set_walk_info(double kernel_ms, int64_t launches, int64_t steps, int64_t proposals, int64_t member_tests, int64_t probes_log2);

// graph_build.cu
srw_status srw_build_graph_device(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w,
                                  const int32_t *d_pid, int directed, unsigned flags, srw_graph **out);
// adjacency rows given explicitly (GraphMap builder): entries already grouped by row in row order
srw_status srw_build_graph_rows(int64_t n_rows, const int32_t *h_vids, const int64_t *h_row_off, const int64_t *h_row_len,
                                const int32_t *h_dst, const int32_t *h_pid, const float *h_w, unsigned flags,
                                srw_graph **out);
// lean handles: neighbour ranks of every entry (d_ent[].x holds ids) copied to the host in chunks
srw_status srw_ent_ranks_to_host(const srw_graph *g, int32_t *h_col);
void srw_alias_thresholds(double p, double q, uint64_t *t_ret, uint64_t *t_common, uint64_t *t_far);

// text_io.cu
srw_status srw_parse_text_device(const char *h_text, size_t len, int weighted, int partitioned, int64_t *n_out, int32_t **d_src,
                                 int32_t **d_dst, float **d_w, int32_t **d_pid);
srw_status srw_graph_load_device(const srw_params *params, unsigned flags, srw_graph **out);
