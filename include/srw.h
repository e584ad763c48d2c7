/*
 * srw.h -- C ABI of libsrw, the B200-native node2vec second-order random-walk engine.
 *
 * Drop-in boundary for the hot path of data61/stellar-random-walk (reference @ 0b2da95).  The
 * reference has no FFI; its seam is the Scala trait `RandomWalk` as driven by `Main.doRandomWalk`.
 * Every entry point below names the reference interface it replaces.  Paths are relative to
 * /root/reference/randomwalk/src/main/scala/au/csiro/data61/randomwalk/ :
 *   Main = Main.scala, CP = common/CommandParser.scala, Params = common/Params.scala,
 *   RW = algorithm/RandomWalk.scala, URW = algorithm/UniformRandomWalk.scala,
 *   VRW = algorithm/VCutRandomWalk.scala, GM = algorithm/GraphMap.scala,
 *   RS = algorithm/RandomSample.scala.
 *
 * Conventions: plain pointers and sizes only; every function returns srw_status (0 = OK) unless
 * stated otherwise; the message for the calling thread's last failure is srw_last_error();
 * handles are opaque and freed by the caller.  Calls are blocking.  The library has NO CPU
 * walk path: graph and walk calls fail with SRW_ERR_NO_DEVICE when no CUDA device is present.
 * Pointers named h_* are host memory, d_* are device memory of the current CUDA device.
 */
#ifndef SRW_H
#define SRW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum srw_status {
  SRW_OK = 0,
  SRW_ERR_ARG = 1,        /* bad argument */
  SRW_ERR_USAGE = 2,      /* CP:32-109 parse failure: caller prints srw_usage(), exits 1 (Main:25) */
  SRW_ERR_PARSE = 3,      /* edge-list line the reference would throw on (URW:34) */
  SRW_ERR_IO = 4,
  SRW_ERR_CUDA = 5,
  SRW_ERR_NO_DEVICE = 6,
  SRW_ERR_UNSUPPORTED = 7
} srw_status;

/* CP:7-10 TaskName */
enum { SRW_TASK_NODE2VEC = 0, SRW_TASK_RANDOMWALK = 1, SRW_TASK_EMBEDDING = 2 };
/* sampler: ALIAS = Vose proposal + p/q rejection (throughput path); EXACT = RS:12-62 literal
 * float32/float64 inverse-CDF (bit-parity path) */
enum { SRW_SAMPLER_ALIAS = 0, SRW_SAMPLER_EXACT = 1,
       /* ALIAS with the return edge folded out of the rejection envelope (undirected, unweighted graphs
        * with 1/p > max(1, 1/q); anything else runs as ALIAS).  Same distribution, about half the
        * proposals per step; bit-identical to its own CPU twin (oracle cfg.fold = 1). */
       SRW_SAMPLER_ALIAS_FOLD = 2 };
/* where a draw comes from: Philox4x32-10 keyed by (seed; walker, step, trial), or the constant
 * generator the reference's tests inject (`nextFloat = () => rValue`, RS:5, T-URW:183-185) */
enum { SRW_U_PHILOX = 0, SRW_U_CONST = 1 };

#define SRW_PATH_MAX 1024

/* Params.scala:7-23, field for field, followed by this build's additive options. */
typedef struct srw_params {
  int32_t w2v_iter;        /* w2vIter = 10        (accepted, unused: word2vec is out of scope) */
  double w2v_lr;           /* w2vLr = 0.025 */
  int32_t w2v_partitions;  /* w2vPartitions = 1 */
  int32_t w2v_dim;         /* w2vDim = 128 */
  int32_t w2v_window;      /* w2vWindow = 10 */
  int32_t walk_length;     /* walkLength = 80 */
  int32_t num_walks;       /* numWalks = 10 */
  double p;                /* p = 1.0 */
  double q;                /* q = 1.0 */
  int32_t weighted;        /* weighted = true */
  int32_t directed;        /* directed = false */
  char input[SRW_PATH_MAX];   /* input = null  -> "" */
  char output[SRW_PATH_MAX];  /* output = null -> "" */
  int32_t rdd_partitions;  /* rddPartitions = 200 */
  int32_t single_output;   /* singleOutput = true */
  int32_t partitioned;     /* partitioned = false */
  int32_t cmd;             /* cmd = node2vec */
  /* ---- additive ---- */
  uint64_t seed;           /* --seed (default 1): the reference has no seed at all */
  int32_t sampler;         /* --sampler alias|fold|exact (default fold = SRW_SAMPLER_ALIAS_FOLD; runs as alias where folding does not apply) */
  int32_t u_mode;          /* SRW_U_PHILOX; SRW_U_CONST for the reference's constant-u tests */
  float u_const;
  int32_t num_gpus;        /* --gpus (default 1) */
} srw_params;

typedef struct srw_edges srw_edges;   /* parsed edge list (host) */
typedef struct srw_graph srw_graph;   /* GraphMap replacement: CSR (+sorted copy, +Vose slots) in HBM */
typedef struct srw_paths srw_paths;   /* result of a walk: RDD[Array[Int]] replacement */
typedef struct srw_graphmap srw_graphmap; /* incremental builder with GraphMap.addVertex semantics */

const char *srw_last_error(void);
const char *srw_version(void);
int srw_device_count(void);           /* number of CUDA devices visible (0 = none), never fails */

/* ---- Params / CommandParser ---- */
srw_status srw_params_default(srw_params *out);                        /* Params:7-23 */
/* CP:32-109 (scopt): --name value pairs, --input/--output/--cmd required, scopt booleans
 * (true/false/yes/no/1/0), unknown option or bad value => SRW_ERR_USAGE. */
srw_status srw_params_parse_argv(int argc, const char *const *argv, srw_params *out);
const char *srw_usage(void);                                           /* CP:33-105 usage text */

/* ---- A1: edge-list text -> arrays (URW:23-34 rules; VRW:19-34 when partitioned) ---- */
srw_status srw_edges_parse_file(const char *path, int weighted, int partitioned, srw_edges **out);
srw_status srw_edges_parse_buffer(const char *buf, size_t len, int weighted, int partitioned, srw_edges **out);
/* any out pointer may be NULL; *h_pid is NULL unless parsed with partitioned != 0 */
srw_status srw_edges_view(const srw_edges *e, int64_t *n, const int32_t **h_src, const int32_t **h_dst,
                          const float **h_w, const int32_t **h_pid);
void srw_edges_free(srw_edges *e);
/* The same rules applied ON THE DEVICE (text_io.cu: one thread per line; weight tokens outside the exact float fast
 * path are re-parsed on the host), returned as the same host handle.  srw_graph_load uses this parser and keeps the
 * arrays in HBM.  Identical output to srw_edges_parse_buffer, including the first error and its line number. */
srw_status srw_edges_parse_buffer_device(const char *buf, size_t len, int weighted, int partitioned, srw_edges **out);

/* ---- A1+A2: adjacency build (URW:35-43 + GM:41-64 semantics) on the device ---- */
#define SRW_BUILD_EXACT 1u   /* keep the file-appearance-order rows (needed by SRW_SAMPLER_EXACT) */
#define SRW_BUILD_ALIAS 2u   /* build neighbour-sorted rows + Vose slots (needed by SRW_SAMPLER_ALIAS) */
#define SRW_BUILD_ALL 3u
#define SRW_BUILD_MIGRATE 4u /* sharded builds: also the replicated edge filter the migrating-walker exchange needs (srw_mig_*) */
#define SRW_BUILD_LEAN 8u    /* with SRW_BUILD_ALIAS on an unweighted, undirected, unsharded graph with non-negative ids: keep ONLY
                              * what --sampler alias|fold reads (16-byte neighbour entries + id-labelled hash sets: 24 bytes per
                              * adjacency entry instead of 36); the sorted column array, the rank-labelled hash sets and the row
                              * descriptors are dropped.  Ignored (full build) where the conditions do not hold.  Same paths. */
srw_status srw_graph_from_edges(int64_t n, const int32_t *h_src, const int32_t *h_dst, const float *h_w /*NULL=1.0f*/,
                                const int32_t *h_pid /*NULL*/, int directed, unsigned flags, srw_graph **out);
srw_status srw_graph_from_device_edges(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w,
                                       const int32_t *d_pid, int directed, unsigned flags, srw_graph **out);
/* One process, several GPUs (`--gpus N`, srw_params::num_gpus): one edge-balanced vertex-range shard per device inside ONE handle.
 * srw_walk / srw_walk_save / srw_walk_device (whole rounds, result on device 0) then run the migrating-walker walk (srw_mig_*
 * below) over peer memory; srw_graph_stats / _neighbors / _vertex_ids answer for the whole graph.  Undirected, unweighted.
 * srw_graph_load builds the same when params->num_gpus > 1. */
srw_status srw_graph_from_edges_multi(int64_t n, const int32_t *h_src, const int32_t *h_dst, int directed, int num_gpus, srw_graph **out);
/* The same with the partition-id column as the shard map (VCutRandomWalk: `--partitioned true --gpus N`; see
 * srw_graph_from_device_edges_vcut).  srw_graph_load does this when params->partitioned && params->num_gpus > 1. */
srw_status srw_graph_from_edges_multi_vcut(int64_t n, const int32_t *h_src, const int32_t *h_dst, const int32_t *h_pid, int directed,
                                           int num_gpus, srw_graph **out);
/* Where the build of this handle spent its time: a JSON object {"phase": milliseconds, ...} (host clock around
 * device-synchronised phases of K1-K3); valid until the handle is freed. */
const char *srw_graph_build_profile(const srw_graph *g);
/* loadGraph(): URW:17-88 / VRW:13-98 chosen by params->partitioned (Main:54-57) */
srw_status srw_graph_load(const srw_params *params, unsigned flags, srw_graph **out);
/* RW:23-24 nVertices / nEdges (= adjacency entries) */
srw_status srw_graph_stats(const srw_graph *g, int64_t *n_vertices, int64_t *n_edges);
/* GM:109-120 getNeighbors: *n = -1 unknown vid (reference: null), 0 dead end, else degree.
 * Copies up to cap entries (file-appearance order) into h_dst / h_w (either may be NULL). */
srw_status srw_graph_neighbors(const srw_graph *g, int32_t vid, int32_t *h_dst, float *h_w, int64_t cap, int64_t *n);
/* GM:66-68 getPartition: *found = 0 when the vid was never a neighbour in a partitioned load */
srw_status srw_graph_partition(const srw_graph *g, int32_t vid, int32_t *pid, int *found);
/* ascending vertex ids (this build's emission order) */
srw_status srw_graph_vertex_ids(const srw_graph *g, int32_t *h_out, int64_t cap);
/* device layout for inspection/tests: row offsets [nv+1], sorted neighbour ranks [nnz], and when
 * has_alias the Vose slots {thr, own, alias_vertex, alias_index} [nnz]; any pointer may be NULL.  On a vertex-range shard
 * the arrays are the shard's own: row_last - row_first + 1 offsets relative to its first entry, nnz_local entries; and
 * srw_graph_neighbors answers -1 ("not on this shard", GM:118) for a vertex another shard owns. */
srw_status srw_graph_layout(const srw_graph *g, int64_t *h_offsets, int32_t *h_col_sorted, uint32_t *h_slots4,
                            int *has_alias);
int64_t srw_graph_device_bytes(const srw_graph *g);
void srw_graph_free(srw_graph *g);

/* GraphMap.addVertex x3 / reset (GM:23-56, 83-85, 99-107) as an incremental host builder */
srw_status srw_graphmap_new(srw_graphmap **out);
srw_status srw_graphmap_add_vertex(srw_graphmap *m, int32_t vid, int64_t n, const int32_t *h_dst,
                                   const int32_t *h_pid /*NULL*/, const float *h_w /*NULL=1.0f*/);
srw_status srw_graphmap_reset(srw_graphmap *m);
srw_status srw_graphmap_counts(const srw_graphmap *m, int64_t *n_vertices, int64_t *n_edges); /* GM:87-93 */
srw_status srw_graphmap_finalize(const srw_graphmap *m, unsigned flags, srw_graph **out);
void srw_graphmap_free(srw_graphmap *m);

/* ---- A5-A7 on the device, for known-answer tests (one sampler call through the same device
 * functions the exact walk kernel uses) ---- */
srw_status srw_sample(int64_t n, const int32_t *h_dst, const float *h_w, float u, int32_t *dst_out, float *w_out); /* RS:12-25 */
srw_status srw_second_order_weights(float p, float q, int32_t prev, int64_t np, const int32_t *h_pdst, int64_t nc,
                                    const int32_t *h_cdst, const float *h_cw, float *h_out);          /* RS:27-44 */
srw_status srw_second_order_sample(float p, float q, int32_t prev, int64_t np, const int32_t *h_pdst, int64_t nc,
                                   const int32_t *h_cdst, const float *h_cw, float u, int32_t *dst_out,
                                   float *w_out);                                                     /* RS:55-62 */
srw_status srw_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);          /* device KAT */

/* ---- A8-A10: randomWalk(initPaths) (RW:75-176): all numWalks rounds; one walker per vertex per
 * round; uses walk_length, num_walks, p, q, seed, sampler, u_mode, u_const, num_gpus ---- */
srw_status srw_walk(const srw_graph *g, const srw_params *params, srw_paths **out);
/* Device-resident variant: walkers [walker_first, walker_first + n_walkers) with walker id =
 * round * nVertices + vertex rank.  d_paths is [n_walkers][walk_length+2] int32 (vertex ids,
 * row-major), d_lens [n_walkers].  Nothing leaves the device.  stream: a cudaStream_t or NULL. */
srw_status srw_walk_device(const srw_graph *g, const srw_params *params, uint64_t walker_first, int64_t n_walkers,
                           int32_t *d_paths, int32_t *d_lens, void *stream);
/* timing / counters of the calling thread's most recent srw_walk* call */
/* srw_walk_device in two halves.  _async enqueues the walk kernel on `stream` and the rank -> id pass on the library's
 * high-priority finalisation stream, and returns; `stream` is then free for the next round's walk (into ANOTHER path
 * buffer), so the L2-bound finalisation of round r runs under the DRAM-request-bound walk of round r+1.  srw_walk_wait
 * blocks until the ticket's round is complete (d_paths / d_lens valid), fills *info (may be NULL) and releases the ticket.
 * Every ticket must be waited for exactly once; the buffers of a round must not be reused before its wait. */
typedef struct srw_walk_ticket srw_walk_ticket;
struct srw_walk_info;
srw_status srw_walk_device_async(const srw_graph *g, const srw_params *params, uint64_t walker_first, int64_t n_walkers,
                                 int32_t *d_paths, int32_t *d_lens, void *stream, srw_walk_ticket **ticket);
srw_status srw_walk_wait(srw_walk_ticket *ticket, struct srw_walk_info *info);

typedef struct srw_walk_info {
  double kernel_ms;        /* CUDA-event time of the walk kernel(s) alone */
  int64_t kernel_launches; /* number of kernels launched (walk + finalize) */
  int64_t steps;           /* sampled transitions = sum(len - 1) */
  int64_t proposals;       /* alias proposals on second-order steps (0 unless stats were requested) */
  int64_t member_tests;
  int64_t probes_log2;     /* sum over membership tests of ceil(log2(deg(prev)+1)) */
} srw_walk_info;
srw_status srw_last_walk_info(srw_walk_info *out);
/* name of the walk kernel (with its template arguments) this thread's last srw_walk* call launched, e.g.
 * "walk_fold_conv_kernel<0,0,1,4,1>": lets a profile be matched to the launch it describes */
const char *srw_last_walk_kernel(void);
/* make the next srw_walk_device of this thread also count proposals/member tests (slower kernel) */
srw_status srw_walk_collect_stats(int enable);

/* ragged view, host memory owned by the handle: ids[offsets[i] .. offsets[i+1]) is path i, in
 * (round, ascending vertex id) order */
srw_status srw_paths_view(srw_paths *paths, int64_t *n_paths, const int32_t **h_ids, const int64_t **h_offsets);
srw_status srw_paths_counts(const srw_paths *paths, int64_t *n_paths, int64_t *n_steps);
/* RW:234-241 save(): <output>/path/part-NNNNN, ids joined by '\t', one path per line; 1 file when
 * single_output else rdd_partitions files (Main:64-69), plus an empty _SUCCESS marker */
srw_status srw_save(srw_paths *paths, const srw_params *params);
/* the same text into a caller buffer; *needed = bytes required */
srw_status srw_paths_format(srw_paths *paths, char *h_buf, int64_t cap, int64_t *needed);
void srw_paths_free(srw_paths *paths);
/* RW:234-241 on the device: the text of n_paths rows of a [n_paths][stride] path matrix (row i holds d_lens[i] ids,
 * as srw_walk_device leaves it) -- `ids.mkString("\t") + "\n"` per row, rows in order -- into d_text (device
 * memory, cap bytes).  *needed = bytes of the text; pass d_text = NULL to size the buffer first. */
srw_status srw_paths_format_device(const int32_t *d_paths, const int32_t *d_lens, int64_t n_paths, int32_t stride,
                                   char *d_text, int64_t cap, int64_t *needed, void *stream);
/* execute() + save() streamed (RW:31-33, RW:234-241, Main:53-62): every round is walked, formatted on the device and
 * written to <output>/path/part-NNNNN in walker order, so neither the paths nor their text have to fit in host
 * memory.  Same files as srw_walk + srw_save.  srw_last_walk_info() then holds the totals of the call. */
srw_status srw_walk_save(const srw_graph *g, const srw_params *params);

/* Main.main / runJob for --cmd randomwalk (Main:18-27, 109-127): parse, load, walk, save.
 * Prints the reference's "edges:/vertices:" lines (URW:69-72).  Returns the process exit code. */
int srw_main(int argc, const char *const *argv);

/* ---- A10 / SURVEY 8(e): vertex-range shards.  Replaces the walker routing of URW:103-112 /
 * VRW:121-134 and the per-super-step shuffle RW:186-192.  One srw_graph per rank holds the rows of a
 * contiguous, edge-balanced vertex range; walkers travel as 32-byte tuples, path entries reach the
 * walker's home shard (owner of its start vertex) as 16-byte records.  The library fills the send
 * buffers; the caller performs the all-to-all (NCCL) between srw_shard_step calls -- see
 * stellar-random-walk_b200/sharded.py for the loop (RW:91-162).  Alias sampler only. ---- */
srw_status srw_graph_from_device_edges_sharded(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w,
                                               int directed, unsigned flags, int rank, int world, srw_graph **out);
/* SURVEY 8(f)3 -- VERTEX-CUT shards (VCutRandomWalk): table-mapped instead of plain vertex ranges.
 *   d_pid != NULL: the VCut shard map (VRW:23-26 partition-id column, VRW:121-134 routing by GraphMap.getPartition(steps.last),
 *     GM:31,66-68): owner(v) = getPartition(v) mod world, where getPartition(v) is the partition id of the last input line in
 *     which v is a neighbour (srw_graph_partition).  d_pid == NULL: edge-balanced vertex ranges.
 *   hub_fraction in [0, 0.95]: the rows of the highest-degree vertices -- as many as hold up to that share of all adjacency entries
 *     -- are REPLICATED on every shard (VRW:43-54 replicates a vertex's adjacency into every partition it has an edge in): a walker
 *     that steps onto a hub does not migrate.  Costs hub_fraction x 24 bytes per adjacency entry of HBM on every shard.
 * An 8-byte-per-vertex extent table and a 1-byte-per-vertex owner table are replicated on every shard.  Walked by the migrating
 * walk (srw_mig_*) only; same paths as any other sharding.  For such a shard srw_graph_shard_info reports row_first = 0,
 * row_last = its local row count (hub rows + own rows), bounds = own-row group sizes as a prefix sum.
 * Undirected, unweighted; flags as srw_graph_from_device_edges_sharded. */
srw_status srw_graph_from_device_edges_vcut(int64_t n, const int32_t *d_src, const int32_t *d_dst, const int32_t *d_pid /*NULL*/,
                                            int directed, unsigned flags, int rank, int world, double hub_fraction, srw_graph **out);
/* What hub_fraction selected on this shard (identical on every shard; 0 / 0 / 0xFFFFFFFF when nothing is replicated): replicated
 * rows, the adjacency entries they hold, the smallest replicated degree; *seed_rows = the vertices this shard starts walkers for. */
srw_status srw_graph_hub_info(const srw_graph *g, int64_t *hub_rows, int64_t *hub_entries, uint32_t *hub_min_degree, int64_t *seed_rows);
/* bounds: world+1 first-ranks of the shards (identical on every rank) */
srw_status srw_graph_shard_info(const srw_graph *g, int *rank, int *world, int64_t *row_first, int64_t *row_last,
                                int64_t *bounds, int64_t *nnz_local);
int srw_walker_msg_bytes(void);   /* 32 */
int srw_path_rec_bytes(void);     /* 16 */
/* RW:81-87 initial walkers of rounds [round_first, round_first+n_rounds) for this shard's vertices;
 * d_paths is [rows*n_rounds][walk_length+2], d_lens [rows*n_rounds] */
srw_status srw_shard_seed(const srw_graph *g, const srw_params *params, int64_t round_first, int64_t n_rounds, void *d_inbox,
                          int64_t cap, int64_t *n_seeded, int32_t *d_paths, int32_t *d_lens, void *stream);
/* one super-step: advance the n_in resident tuples of d_inbox until each needs a remote row; outgoing
 * tuples / records are left in d_send_msgs / d_send_recs as contiguous per-destination segments in
 * rank order with h_msg_counts[world] / h_rec_counts[world] items each */
srw_status srw_shard_step(const srw_graph *g, const srw_params *params, int64_t round_first, int64_t n_rounds,
                          const void *d_inbox, int64_t n_in, void *d_send_msgs, void *d_send_recs, int64_t rec_cap,
                          int32_t *d_paths, int32_t *d_lens, int64_t *h_msg_counts, int64_t *h_rec_counts, int64_t *steps_done,
                          void *stream);
srw_status srw_shard_apply(const srw_graph *g, const srw_params *params, int64_t round_first, int64_t n_rounds,
                           const void *d_recs, int64_t n_recs, int32_t *d_paths, void *stream);
/* ranks -> vertex ids over the home rows; *steps = sum(len - 1) */
srw_status srw_shard_finalize(const srw_graph *g, const srw_params *params, int64_t n_rows, int32_t *d_paths,
                              const int32_t *d_lens, int64_t *steps, void *stream);

/* ---- A10, peer-gather variant.  The same vertex-range shards, but no walker ever moves: every shard's row
 * arrays are made addressable from every GPU of the box (CUDA IPC mappings between the per-GPU processes,
 * plain pointers inside one process) and the walk kernel loads remote rows over NVLink.  This is what the
 * reference's "ship prevNeighbors with the walker" (RW:135) and its per-super-step shuffle (RW:186-192)
 * collapse to on an NVSwitch box.  Once every peer is attached, srw_walk_device(g = any shard handle, ...)
 * walks ANY range of walkers against the whole graph; output is bit-identical to the unsharded graph's.
 * Undirected, unweighted graphs (SRW_BUILD_ALIAS), samplers alias | fold. ---- */
int srw_shard_ipc_bytes(void);                                       /* size of the export blob */
srw_status srw_shard_ipc_export(const srw_graph *g, void *h_blob);  /* this shard's row arrays as IPC handles */
srw_status srw_shard_ipc_attach(srw_graph *g, const void *h_blob);  /* map another process's shard (own blob: no-op) */
srw_status srw_shard_attach_local(srw_graph *g, const srw_graph *peer);   /* peer handle lives in this process */
/* One process per GPU, the fast way: the host runtime allocates one symmetric-memory block per rank (CUDA VMM
 * allocations whose handles it exchanges; torch.distributed._symmetric_memory in sharded.py), the shard moves
 * its row arrays into its block, and every peer's block pointer is attached.  block layout: see shard.cu. */
srw_status srw_shard_rows_info(const srw_graph *g, int64_t *rows, int64_t *nnz, int64_t *hash_buckets, int64_t *block_bytes);
srw_status srw_shard_rows_relocate(srw_graph *g, void *d_block, int64_t block_bytes);   /* block must outlive g */
srw_status srw_shard_attach_block(srw_graph *g, int peer_rank, const void *d_block, int64_t rows, int64_t nnz,
                                  int64_t hash_buckets);

/* ---- A10, migrating walkers with the exchange fused into the step kernel (csrc/migrate.cuh) -- what `bench.py --gpus N`
 * and srw_walk with num_gpus > 1 run.  Replaces RW:91-162 (super-step loop), RW:186-192 (shuffle), URW:103-112 (routing key =
 * current vertex).  A walker lives on owner(curr); one kernel per super-step advances every resident walker and stores each
 * departing walker as a 32-byte tuple straight into the destination GPU's inbox (peer memory), path entries straight into the
 * walker's home GPU's path rows.  The d(t,x)=1 test (RS:38) is answered by a replicated edge filter and verified exactly at
 * owner(x) (t in N(x)); ~1 hop per step.  Undirected, unweighted shards built with SRW_BUILD_ALIAS | SRW_BUILD_MIGRATE;
 * samplers alias | fold; output bit-identical to the unsharded walk for any number of shards.
 *
 * The caller owns the peer-visible block of every rank (srw_mig_block_bytes bytes, the same on every rank: symmetric memory
 * between processes, plain device memory with peer access inside one process) and the barrier between super-steps:
 *     srw_mig_begin(m, round_first, n_rounds);  barrier;
 *     for (s = 0; ; ++s) { srw_mig_superstep(m, s, d_sent); all-reduce(d_sent) (= barrier); if (s >= 1 && sum == 0) break; }   (RW:162;
 *                                                      super-steps 0 and 1 both seed walkers)
 *     srw_mig_finish(m, ...)
 * A super-step with nothing to do is harmless, so the termination test need not be read back every iteration. ---- */
typedef struct srw_mig srw_mig;
srw_status srw_mig_block_bytes(const srw_graph *g, const srw_params *params, int64_t n_rounds, int64_t seg_cap /*0 = default*/, int64_t *bytes);
/* d_block_peers[world]: every rank's block as addressable from this device (entry `rank` is ignored: d_block_self) */
srw_status srw_mig_create(const srw_graph *g, const srw_params *params, int64_t n_rounds, int64_t seg_cap, void *d_block_self,
                          void *const *d_block_peers, srw_mig **out);
srw_status srw_mig_collect_stats(srw_mig *m, int enable);    /* instrumented kernel: proposals / tests / exact tests */
srw_status srw_mig_begin(srw_mig *m, int64_t round_first, int64_t n_rounds /* <= the n_rounds of srw_mig_create */, void *stream);
srw_status srw_mig_superstep(srw_mig *m, int64_t s, unsigned long long *d_sent /*device, may be NULL*/, void *stream);
/* [0] inbox slots this rank filled in the last super-step, [1] steps, [2] proposals, [3] membership tests (filter probes),
 * [4] exact tests, [5] tuples spilled (region full), [6] error flags (nonzero => SRW_ERR_CUDA), [7] exact tests that found the edge (instrumented); synchronises */
srw_status srw_mig_counters(srw_mig *m, int64_t *h_out8, void *stream);
/* ranks -> vertex ids over this rank's home rows.  *d_paths: [home_rows * n_rounds][walk_length + 2] inside the block (valid
 * until the next srw_mig_begin); row (round - round_first) * home_rows + v / world belongs to the walker that started at
 * vertex rank v = rank + (row % home_rows) * world.  d_lens_out (device, n_rows int32, may be NULL) receives the lengths. */
srw_status srw_mig_finish(srw_mig *m, int32_t **d_paths, int32_t *d_lens_out, int64_t *n_rows, int64_t *steps, void *stream);
srw_status srw_mig_info(const srw_mig *m, int64_t *seg_cap, int64_t *spill_cap, int64_t *home_rows, int64_t *block_bytes);
void srw_mig_free(srw_mig *m);

/* ---- synthetic inputs for the benchmark; device-resident, not on the walk path ---- */
srw_status srw_synth_rmat_device(int scale, int edge_factor, uint64_t seed, int64_t first, int64_t count,
                                 int32_t *d_src, int32_t *d_dst);
srw_status srw_synth_weights_device(uint64_t seed, int64_t first, int64_t count, float *d_w);

#ifdef __cplusplus
}
#endif
#endif
