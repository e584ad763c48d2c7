/*
 * srw_oracle.h -- CPU ORACLE for the node2vec second-order random-walk hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / CPU baseline -- never as the thing shipped.
 *
 * It is a plain-C restatement of the reference algorithm (data61/stellar-random-walk @ 0b2da95).
 * The reference is Scala-on-Spark and cannot be compiled or run in this image (no JVM, no
 * Spark), so the restatement is pinned on the reference's own known-answer tests
 * (RandomSampleTest, GraphMapTest, the karate/testgraph counts and the constant-u walk
 * scenarios) -- see tests/test_oracle_golden.py.  What those tests do NOT pin -- the JDK
 * random stream, Spark's output line order, neighbour order on a cluster -- is "parity
 * unpinned" and is defined here (Philox4x32-10 counter RNG, (round, ascending vertex id)
 * emission order, file-appearance neighbour order).
 *
 * Abbreviations for citations (paths relative to
 * /root/reference/randomwalk/src/main/scala/au/csiro/data61/randomwalk/):
 *   RS  = algorithm/RandomSample.scala      GM  = algorithm/GraphMap.scala
 *   RW  = algorithm/RandomWalk.scala        URW = algorithm/UniformRandomWalk.scala
 *   VRW = algorithm/VCutRandomWalk.scala
 */
#ifndef SRW_ORACLE_H
#define SRW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Philox4x32-10 (replacement for scala.util.Random.nextFloat, RW:9,52,76 / RS:5) ---- */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* u on the same 2^-24 grid as java.util.Random.nextFloat: (r0 >> 8) * 2^-24 */
float oracle_u01(uint64_t seed, uint64_t walker, uint32_t step);

/* ---- GraphMap restatement (GM:11-121) ---- */
typedef struct og_graph og_graph;
og_graph *og_new(void);
void og_free(og_graph *g);
void og_reset(og_graph *g);                                                   /* GM:99-107 */
void og_add_vertex(og_graph *g, int32_t vid, const int32_t *dst, const float *w, int64_t n); /* GM:41-56 */
void og_add_vertex_pid(og_graph *g, int32_t vid, const int32_t *dst, const int32_t *pid,
                       const float *w, int64_t n);                            /* GM:23-39 */
void og_add_vertex_empty(og_graph *g, int32_t vid);                           /* GM:83-85 */
/* GM:109-120: returns -1 for an unknown vid (reference: null), 0 for a dead end, else the degree.
 * The reference returns a COPY; the oracle hands out pointers into its own arrays. */
int64_t og_neighbors(const og_graph *g, int32_t vid, const int32_t **dst, const float **w);
int og_partition(const og_graph *g, int32_t vid, int32_t *pid);               /* GM:66-68, 1 if present */
int64_t og_num_vertices(const og_graph *g);                                   /* GM:87-89 */
int64_t og_num_edges(const og_graph *g);                                      /* GM:91-93 */
/* ascending list of all vertex ids (this build's emission order; reference order is arbitrary) */
int64_t og_vertex_ids(const og_graph *g, int32_t *out, int64_t cap);

/* ---- loaders: URW:23-43 (partitioned=0) and VRW:19-54 (partitioned=1) ----
 * Returns 0, or the 1-based number of the first line the reference would throw on
 * (NumberFormatException / ArrayIndexOutOfBounds), with a message in err. */
int64_t og_load_text(og_graph *g, const char *buf, size_t len, int weighted, int directed,
                     int partitioned, char *err, size_t errcap);
/* same adjacency semantics from parsed arrays (w may be NULL = 1.0f, pid may be NULL) */
void og_load_edges(og_graph *g, int64_t n, const int32_t *src, const int32_t *dst, const float *w,
                   const int32_t *pid, int directed);

/* ---- RandomSample restatement (RS:5-63) ---- */
/* RS:12-25; returns the index of the chosen edge (n must be > 0) */
int64_t oracle_sample(int64_t n, const float *w, float u);
/* RS:27-44; out[i] = biased float32 weight of currNeighbors[i] */
void oracle_second_order_weights(float p, float q, int32_t prev, int64_t np, const int32_t *pdst,
                                 int64_t nc, const int32_t *cdst, const float *cw, float *out);
/* RS:55-62; returns chosen index into currNeighbors, *w_out = its biased weight (T-RS:76) */
int64_t oracle_second_order_sample(float p, float q, int32_t prev, int64_t np, const int32_t *pdst,
                                   int64_t nc, const int32_t *cdst, const float *cw, float u,
                                   float *w_out);

/* ---- walk driver restatement (RW:51-66 first step, RW:75-176 loop, local[*] semantics) ---- */
enum { ORACLE_U_CONST = 0, ORACLE_U_PHILOX = 1 };
typedef struct oracle_walk_cfg {
  int32_t walk_length;   /* Params.walkLength; a full path has walk_length + 2 ids (RW:103,132) */
  int32_t num_walks;     /* Params.numWalks (RW:82) */
  double p, q;           /* Params.p/q; narrowed with (float) exactly like RW:112-113 */
  int32_t u_mode;        /* ORACLE_U_CONST: every draw = u_const (the reference tests' generator) */
  float u_const;
  uint64_t seed;         /* ORACLE_U_PHILOX: u = oracle_u01(seed, walker, step) */
  int32_t threads;       /* OpenMP threads over walkers (<=0: all) */
  /* optional walker subset for bounded CPU-baseline timing: walkers w with
   * (w % sample_mod) == 0 only (sample_mod <= 1: all walkers) */
  int64_t sample_mod;
  /* alias twin only: 1 = "alias-fold" (the product's SRW_SAMPLER_ALIAS_FOLD): on an undirected,
   * unweighted graph with 1/p > max(1, 1/q) the return edge's excess weight is sampled as its own
   * mixture component, so the rejection envelope is max(1, 1/q) instead of 1/p.  Ignored (classic
   * rejection) when the graph or (p, q) do not qualify.  On a WEIGHTED undirected graph the same folding applies with
   * (multiplicity, degree) replaced by (bundle weight, row weight sum), and the accept draw r[2] doubles as the
   * component draw (rescaled), because r[1] is the Vose coin. */
  int32_t fold;
} oracle_walk_cfg;

/* Runs all rounds.  Paths are emitted in (round, ascending vertex id) order into a ragged array:
 * ids[offsets[i] .. offsets[i+1]).  Returns the number of paths, or -1 when ids_cap is too small
 * (needed size in offsets[0]).  offsets must hold n_paths+1 entries = num_walks*|V| + 1. */
int64_t oracle_walk(const og_graph *g, const oracle_walk_cfg *cfg, int32_t *ids, int64_t ids_cap,
                    int64_t *offsets);

/* RW:234-241: one line per path, ids joined by '\t', '\n' terminated.  Returns bytes written
 * (or needed when cap is too small). */
int64_t oracle_format_paths(int64_t n_paths, const int32_t *ids, const int64_t *offsets, char *out,
                            int64_t cap);

/* ---- CPU twin of the product's ALIAS sampler (NOT the reference's algorithm) ----
 * Used for parity level P2: the CUDA alias-mode output must equal this bit for bit. */
typedef struct oa_graph oa_graph;   /* dense, neighbour-sorted CSR + Vose tables built from an og_graph */
oa_graph *oa_build(const og_graph *g);
void oa_free(oa_graph *a);
int64_t oa_num_vertices(const oa_graph *a);
int oa_has_alias(const oa_graph *a);
/* views for comparing against the device build (any pointer may be NULL) */
void oa_view(const oa_graph *a, const int32_t **vids, const int64_t **offsets, const int32_t **col,
             const float **w, const uint32_t **thr, const uint32_t **alias);
typedef struct oracle_alias_stats {
  int64_t steps;        /* sampled transitions */
  int64_t proposals;    /* alias proposals drawn (second-order steps only) */
  int64_t probes_log2;  /* sum over proposals of ceil(log2(deg(prev)+1)) where a membership test ran */
  int64_t member_tests;
} oracle_alias_stats;
int64_t oracle_alias_walk(const oa_graph *a, const oracle_walk_cfg *cfg, int32_t *ids, int64_t ids_cap,
                          int64_t *offsets, oracle_alias_stats *stats);
/* acceptance thresholds, shared definition: T(f) = f>=M ? 2^32 : floor(f/M * 2^32) */
void oracle_alias_thresholds(double p, double q, uint64_t *t_ret, uint64_t *t_common, uint64_t *t_far);
/* fold mode: envelope Mp = max(1, 1/q); *a = 1/p - Mp (> 0 iff folding applies) */
void oracle_fold_thresholds(double p, double q, uint64_t *t_common, uint64_t *t_far, double *a, double *mp);
/* per-entry multiplicity (number of parallel edges to the same neighbour), sorted-row order */
const uint32_t *oa_mult(const oa_graph *a);
/* weighted graphs: per-row weight sum W (sequential double sum, the Vose build's) and per-entry bundle weight */
const double *oa_wsum(const oa_graph *a);
const double *oa_wbundle(const oa_graph *a);
void oa_set_directed(oa_graph *a, int directed);   /* folding needs mult(prev->curr) == mult(curr->prev) */

/* ---- CPU-baseline helpers (bench.py only): dense CSR (vid == rank), no GraphMap hash lookups and
 * no row copies -- both omissions favour the CPU.  The walk itself is walk_one's algorithm
 * (RS:12-62 + RW:103-133) on the CSR rows. ---- */
/* RMAT edges [first, first+count): C twin of stellar-random-walk_b200/synth.py::rmat_edges */
void oracle_rmat_edges(int scale, uint64_t seed, int64_t first, int64_t count, int32_t *src, int32_t *dst, int threads);
/* undirected CSR over ids [0, n_ids): offsets[n_ids+1], col[2*n_edges].  Parallel counting build;
 * neighbour order within a row is arbitrary (timing only -- not a parity input). */
void oracle_csr_build(int64_t n_ids, int64_t n_edges, const int32_t *src, const int32_t *dst, int64_t *offsets,
                      int32_t *col, int threads);
/* Walks a strided sample of walkers (start vertices with degree > 0, stride `sample_stride`) for at
 * most budget_s seconds on `threads` OpenMP threads.  w may be NULL (all 1.0f).  Returns the number
 * of sampled transitions; *elapsed_s = wall time, *walkers_done = walkers finished. */
int64_t oracle_walk_csr_timed(int64_t nv, const int64_t *offsets, const int32_t *col, const float *w,
                              const oracle_walk_cfg *cfg, int64_t sample_stride, int64_t sample_phase,
                              double budget_s, double *elapsed_s, int64_t *walkers_done, uint64_t *checksum);
/* the product's alias-fold sampler (the decisions of oracle_alias_walk with fold = 1) over a dense, neighbour-sorted,
 * unweighted CSR: bench.py's "optimised CPU twin" figure.  paths_out (optional): [n_samples][walk_length + 2], -1 padded. */
int64_t oracle_fold_walk_csr_timed(int64_t nv, const int64_t *offsets, const int32_t *col, const oracle_walk_cfg *cfg,
                                   int64_t sample_stride, int64_t sample_phase, double budget_s, double *elapsed_s,
                                   int64_t *walkers_done, uint64_t *checksum, int32_t *paths_out);
int oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
