/*
 * srw_oracle.c -- CPU ORACLE (test infrastructure only; see srw_oracle.h for the rules).
 *
 * Parity status: pinned on the reference's own KATs (T-RS, T-GM, karate/testgraph loads and
 * the constant-u walk scenarios).  The RNG stream, output order and cluster neighbour order are
 * "parity unpinned" in the reference itself and are defined by this file.
 *
 * Every function cites the reference lines it restates.  Keep it literal: same loop order, same
 * float widths, same `edges.head` fallback.
 */
#include "srw_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11; constants as published in Random123)                   */
/* ------------------------------------------------------------------------------------------ */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline void walker_rng(uint64_t seed, uint64_t walker, uint32_t step, uint32_t trial,
                              uint32_t out[4]) {
  uint32_t ctr[4] = {(uint32_t)walker, (uint32_t)(walker >> 32), step, trial};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  oracle_philox4x32_10(ctr, key, out);
}

/* java.util.Random.nextFloat() = next(24) / (float)(1 << 24): a float on the 2^-24 grid in [0,1).
 * Same grid here, bits from Philox instead of the JDK LCG (stream parity unpinned, see header). */
float oracle_u01(uint64_t seed, uint64_t walker, uint32_t step) {
  uint32_t r[4];
  walker_rng(seed, walker, step, 0u, r);
  return (float)(r[0] >> 8) * (1.0f / 16777216.0f);
}

/* ------------------------------------------------------------------------------------------ */
/* small int32 -> int64 open-addressing map (stands in for scala.collection.mutable.HashMap)   */
/* ------------------------------------------------------------------------------------------ */
typedef struct imap {
  int32_t *keys;
  int64_t *vals;
  uint8_t *used;
  uint64_t cap, size;
} imap;

static uint64_t hash32(int32_t k) {
  uint64_t x = (uint32_t)k;
  x = (x ^ (x >> 16)) * 0x45d9f3bULL;
  x = (x ^ (x >> 16)) * 0x45d9f3bULL;
  return x ^ (x >> 16);
}
static void imap_init(imap *m, uint64_t cap) {
  m->cap = 16;
  while (m->cap < cap) m->cap <<= 1;
  m->size = 0;
  m->keys = (int32_t *)malloc(m->cap * sizeof(int32_t));
  m->vals = (int64_t *)malloc(m->cap * sizeof(int64_t));
  m->used = (uint8_t *)calloc(m->cap, 1);
}
static void imap_destroy(imap *m) {
  free(m->keys); free(m->vals); free(m->used);
  memset(m, 0, sizeof(*m));
}
static int64_t *imap_find(const imap *m, int32_t k) {
  uint64_t i = hash32(k) & (m->cap - 1);
  while (m->used[i]) {
    if (m->keys[i] == k) return &m->vals[i];
    i = (i + 1) & (m->cap - 1);
  }
  return NULL;
}
static void imap_put(imap *m, int32_t k, int64_t v);
static void imap_grow(imap *m) {
  imap n;
  imap_init(&n, m->cap * 2);
  for (uint64_t i = 0; i < m->cap; ++i)
    if (m->used[i]) imap_put(&n, m->keys[i], m->vals[i]);
  imap_destroy(m);
  *m = n;
}
static void imap_put(imap *m, int32_t k, int64_t v) {
  if ((m->size + 1) * 10 > m->cap * 6) imap_grow(m);
  uint64_t i = hash32(k) & (m->cap - 1);
  while (m->used[i]) {
    if (m->keys[i] == k) { m->vals[i] = v; return; }
    i = (i + 1) & (m->cap - 1);
  }
  m->used[i] = 1; m->keys[i] = k; m->vals[i] = v; m->size++;
}

/* ------------------------------------------------------------------------------------------ */
/* GraphMap (GM:11-121)                                                                        */
/* ------------------------------------------------------------------------------------------ */
struct og_graph {
  imap src_vertex_map;        /* GM:13 srcVertexMap: vid -> row index, -1 = no out-neighbours */
  int64_t *offsets, *lengths; /* GM:14-15 */
  int64_t rows, rows_cap;     /* GM:17 indexCounter */
  int32_t *edge_dst;          /* GM:16 edges._1 */
  float *edge_w;              /* GM:16 edges._2 */
  int64_t offset_counter, edges_cap; /* GM:18 */
  imap vertex_partition_map;  /* GM:21 */
};

og_graph *og_new(void) {
  og_graph *g = (og_graph *)calloc(1, sizeof(og_graph));
  imap_init(&g->src_vertex_map, 64);
  imap_init(&g->vertex_partition_map, 16);
  return g;
}
void og_reset(og_graph *g) { /* GM:99-107 */
  imap_destroy(&g->src_vertex_map);
  imap_destroy(&g->vertex_partition_map);
  imap_init(&g->src_vertex_map, 64);
  imap_init(&g->vertex_partition_map, 16);
  g->rows = 0;
  g->offset_counter = 0;
}
void og_free(og_graph *g) {
  if (!g) return;
  imap_destroy(&g->src_vertex_map);
  imap_destroy(&g->vertex_partition_map);
  free(g->offsets); free(g->lengths); free(g->edge_dst); free(g->edge_w);
  free(g);
}
static void og_reserve_edges(og_graph *g, int64_t extra) {
  if (g->offset_counter + extra > g->edges_cap) {
    int64_t c = g->edges_cap ? g->edges_cap : 1024;
    while (c < g->offset_counter + extra) c *= 2;
    g->edge_dst = (int32_t *)realloc(g->edge_dst, (size_t)c * sizeof(int32_t));
    g->edge_w = (float *)realloc(g->edge_w, (size_t)c * sizeof(float));
    g->edges_cap = c;
  }
}
/* GM:58-64 updateIndices */
static void og_update_indices(og_graph *g, int32_t vid, int64_t out_degree) {
  if (g->rows == g->rows_cap) {
    g->rows_cap = g->rows_cap ? g->rows_cap * 2 : 1024;
    g->offsets = (int64_t *)realloc(g->offsets, (size_t)g->rows_cap * sizeof(int64_t));
    g->lengths = (int64_t *)realloc(g->lengths, (size_t)g->rows_cap * sizeof(int64_t));
  }
  imap_put(&g->src_vertex_map, vid, g->rows);
  g->offsets[g->rows] = g->offset_counter;
  g->lengths[g->rows] = out_degree;
  g->rows++;
}
void og_add_vertex_empty(og_graph *g, int32_t vid) { /* GM:83-85: put(vId, -1), overwrites */
  imap_put(&g->src_vertex_map, vid, -1);
}
void og_add_vertex(og_graph *g, int32_t vid, const int32_t *dst, const float *w, int64_t n) {
  /* GM:41-56: the first insertion of a vid wins (GM:42,54) */
  if (imap_find(&g->src_vertex_map, vid)) return;
  if (n > 0) {
    og_update_indices(g, vid, n);
    og_reserve_edges(g, n);
    for (int64_t i = 0; i < n; ++i) {
      g->edge_dst[g->offset_counter] = dst[i];
      g->edge_w[g->offset_counter] = w ? w[i] : 1.0f;
      g->offset_counter++;
    }
  } else {
    og_add_vertex_empty(g, vid);
  }
}
void og_add_vertex_pid(og_graph *g, int32_t vid, const int32_t *dst, const int32_t *pid,
                       const float *w, int64_t n) {
  /* GM:23-39: as above, and records dst -> partition id of the edge leading to it (GM:31) */
  if (imap_find(&g->src_vertex_map, vid)) return;
  if (n > 0) {
    og_update_indices(g, vid, n);
    og_reserve_edges(g, n);
    for (int64_t i = 0; i < n; ++i) {
      g->edge_dst[g->offset_counter] = dst[i];
      g->edge_w[g->offset_counter] = w ? w[i] : 1.0f;
      g->offset_counter++;
      imap_put(&g->vertex_partition_map, dst[i], pid[i]);
    }
  } else {
    og_add_vertex_empty(g, vid);
  }
}
int64_t og_neighbors(const og_graph *g, int32_t vid, const int32_t **dst, const float **w) {
  /* GM:109-120 */
  const int64_t *idx = imap_find(&g->src_vertex_map, vid);
  if (!idx) return -1;             /* case None => null */
  if (*idx == -1) return 0;        /* Array.empty */
  if (dst) *dst = g->edge_dst + g->offsets[*idx];
  if (w) *w = g->edge_w + g->offsets[*idx];
  return g->lengths[*idx];
}
int og_partition(const og_graph *g, int32_t vid, int32_t *pid) { /* GM:66-68 */
  const int64_t *v = imap_find(&g->vertex_partition_map, vid);
  if (!v) return 0;
  if (pid) *pid = (int32_t)*v;
  return 1;
}
int64_t og_num_vertices(const og_graph *g) { return (int64_t)g->src_vertex_map.size; } /* GM:87 */
int64_t og_num_edges(const og_graph *g) { return g->offset_counter; }                 /* GM:91 */

static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}
int64_t og_vertex_ids(const og_graph *g, int32_t *out, int64_t cap) {
  int64_t n = 0;
  const imap *m = &g->src_vertex_map;
  for (uint64_t i = 0; i < m->cap; ++i)
    if (m->used[i]) {
      if (n < cap) out[n] = m->keys[i];
      n++;
    }
  if (n <= cap) qsort(out, (size_t)n, sizeof(int32_t), cmp_i32);
  return n;
}

/* ------------------------------------------------------------------------------------------ */
/* adjacency construction shared by both loaders: `reduceByKey(_ ++ _)` in record order        */
/* (URW:35-41; VRW:35-49).  Entry order within a vertex = file-appearance order; for one line   */
/* the src-side entry precedes the dst-side entry.  Duplicates and self-loops are kept.        */
/* ------------------------------------------------------------------------------------------ */
void og_load_edges(og_graph *g, int64_t n, const int32_t *src, const int32_t *dst, const float *w,
                   const int32_t *pid, int directed) {
  imap tmp;                                    /* vid -> first-seen temp index */
  imap_init(&tmp, (uint64_t)(n > 16 ? n / 2 : 16));
  int64_t nv = 0, cap = 1024;
  int32_t *vids = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
  int64_t *deg = (int64_t *)calloc((size_t)cap, sizeof(int64_t));
#define TMP_INDEX(V, OUT)                                                      \
  do {                                                                         \
    int64_t *f_ = imap_find(&tmp, (V));                                        \
    if (f_) (OUT) = *f_;                                                       \
    else {                                                                     \
      if (nv == cap) {                                                         \
        cap *= 2;                                                              \
        vids = (int32_t *)realloc(vids, (size_t)cap * sizeof(int32_t));        \
        deg = (int64_t *)realloc(deg, (size_t)cap * sizeof(int64_t));          \
        memset(deg + nv, 0, (size_t)(cap - nv) * sizeof(int64_t));             \
      }                                                                        \
      imap_put(&tmp, (V), nv); vids[nv] = (V); (OUT) = nv++;                   \
    }                                                                          \
  } while (0)
  for (int64_t i = 0; i < n; ++i) {
    int64_t a, b;
    TMP_INDEX(src[i], a);
    TMP_INDEX(dst[i], b);
    deg[a]++;                       /* (src, Array((dst, weight)))  URW:36,38 */
    if (!directed) deg[b]++;        /* (dst, Array((src, weight)))  URW:38; directed: (dst, empty) URW:36 */
  }
  int64_t *off = (int64_t *)malloc((size_t)(nv + 1) * sizeof(int64_t));
  off[0] = 0;
  for (int64_t v = 0; v < nv; ++v) off[v + 1] = off[v] + deg[v];
  int64_t tot = off[nv];
  int32_t *adst = (int32_t *)malloc((size_t)(tot ? tot : 1) * sizeof(int32_t));
  int32_t *apid = (int32_t *)malloc((size_t)(tot ? tot : 1) * sizeof(int32_t));
  float *aw = (float *)malloc((size_t)(tot ? tot : 1) * sizeof(float));
  int64_t *fill = (int64_t *)malloc((size_t)(nv ? nv : 1) * sizeof(int64_t));
  memcpy(fill, off, (size_t)nv * sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i) {
    int64_t a = *imap_find(&tmp, src[i]), b = *imap_find(&tmp, dst[i]);
    float wi = w ? w[i] : 1.0f;
    int32_t pi = pid ? pid[i] : 0;
    int64_t k = fill[a]++;
    adst[k] = dst[i]; aw[k] = wi; apid[k] = pi;
    if (!directed) {
      k = fill[b]++;
      adst[k] = src[i]; aw[k] = wi; apid[k] = pi;
    }
  }
  /* buildRoutingTable: GraphMap.addVertex per vertex (URW:90-101 / VRW:107-119).  Vertex
   * insertion order does not matter (T-GM:27-32); first-seen order is used. */
  for (int64_t v = 0; v < nv; ++v) {
    if (pid) og_add_vertex_pid(g, vids[v], adst + off[v], apid + off[v], aw + off[v], deg[v]);
    else     og_add_vertex(g, vids[v], adst + off[v], aw + off[v], deg[v]);
  }
  /* GM:31 `vertexPartitionMap.put(dst, pId)`: the surviving value depends on the order in which
   * Spark hands vertices to addVertex (hash-partition order: unpinned).  This build pins it to FILE
   * order: the partition id of the last input line in which the vertex is a neighbour. */
  if (pid)
    for (int64_t i = 0; i < n; ++i) {
      imap_put(&g->vertex_partition_map, dst[i], pid[i]);
      if (!directed) imap_put(&g->vertex_partition_map, src[i], pid[i]);
    }
#undef TMP_INDEX
  free(fill); free(aw); free(apid); free(adst); free(off); free(deg); free(vids);
  imap_destroy(&tmp);
}

/* ---- text parsing with the JVM's rules ---- */
static int is_java_ws(char c) { /* \s = [ \t\n\x0B\f\r] */
  return c == ' ' || c == '\t' || c == '\n' || c == '\x0B' || c == '\f' || c == '\r';
}
/* java.lang.Integer.parseInt: [+-]?[0-9]+ within int32 */
static int parse_java_int(const char *s, size_t n, int32_t *out) {
  size_t i = 0;
  int neg = 0;
  if (n == 0) return 0;
  if (s[0] == '-' || s[0] == '+') { neg = (s[0] == '-'); i = 1; }
  if (i == n) return 0;
  int64_t v = 0;
  for (; i < n; ++i) {
    if (s[i] < '0' || s[i] > '9') return 0;
    v = v * 10 + (s[i] - '0');
    if (v > 2147483648LL) return 0;
  }
  if (neg) v = -v;
  if (v > 2147483647LL || v < -2147483648LL) return 0;
  *out = (int32_t)v;
  return 1;
}
/* java.lang.Float.parseFloat grammar (decimal / hex / NaN / Infinity, optional fFdD suffix) */
static int parse_java_float(const char *s, size_t n, float *out) {
  char buf[128];
  if (n == 0 || n >= sizeof(buf)) return 0;
  size_t i = 0;
  if (s[i] == '+' || s[i] == '-') i++;
  size_t body = i;
  if (n - body == 3 && memcmp(s + body, "NaN", 3) == 0) { *out = NAN; return 1; }
  if (n - body == 8 && memcmp(s + body, "Infinity", 8) == 0) {
    *out = (s[0] == '-') ? -INFINITY : INFINITY;
    return 1;
  }
  size_t end = n;
  if (end > body && (s[end - 1] == 'f' || s[end - 1] == 'F' || s[end - 1] == 'd' || s[end - 1] == 'D'))
    end--;
  int hex = (end - i >= 2 && s[i] == '0' && (s[i + 1] == 'x' || s[i + 1] == 'X'));
  /* a trailing d/f of a hex literal without exponent is a hex digit: Java requires p-exponent */
  size_t j = i;
  int digits = 0, exp_seen = 0;
  if (hex) {
    end = n;
    if (s[end - 1] == 'f' || s[end - 1] == 'F' || s[end - 1] == 'd' || s[end - 1] == 'D') {
      /* suffix only legal after the exponent digits */
      size_t k = end - 1;
      while (k > i && s[k - 1] >= '0' && s[k - 1] <= '9') k--;
      if (k > i && (s[k - 1] == '+' || s[k - 1] == '-')) k--;
      if (k > i && (s[k - 1] == 'p' || s[k - 1] == 'P') && k < end - 1) end--;
    }
    j = i + 2;
    int dot = 0;
    for (; j < end; ++j) {
      char c = s[j];
      if ((c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F')) digits++;
      else if (c == '.' && !dot) dot = 1;
      else break;
    }
    if (!digits || j >= end || (s[j] != 'p' && s[j] != 'P')) return 0;
    j++;
    if (j < end && (s[j] == '+' || s[j] == '-')) j++;
    size_t e0 = j;
    while (j < end && s[j] >= '0' && s[j] <= '9') j++;
    if (j == e0 || j != end) return 0;
  } else {
    int dot = 0;
    for (; j < end; ++j) {
      char c = s[j];
      if (c >= '0' && c <= '9') digits++;
      else if (c == '.' && !dot) dot = 1;
      else break;
    }
    if (!digits) return 0;
    if (j < end && (s[j] == 'e' || s[j] == 'E')) {
      exp_seen = 1;
      j++;
      if (j < end && (s[j] == '+' || s[j] == '-')) j++;
      size_t e0 = j;
      while (j < end && s[j] >= '0' && s[j] <= '9') j++;
      if (j == e0) return 0;
    }
    (void)exp_seen;
    if (j != end) return 0;
  }
  memcpy(buf, s, end);
  buf[end] = 0;
  char *ep = NULL;
  float v = strtof(buf, &ep);
  if (ep != buf + end) return 0;
  *out = v;
  return 1;
}

int64_t og_load_text(og_graph *g, const char *buf, size_t len, int weighted, int directed,
                     int partitioned, char *err, size_t errcap) {
  int64_t cap = 1024, n = 0, line_no = 0;
  int32_t *src = (int32_t *)malloc((size_t)cap * 4), *dst = (int32_t *)malloc((size_t)cap * 4);
  int32_t *pid = (int32_t *)malloc((size_t)cap * 4);
  float *w = (float *)malloc((size_t)cap * 4);
  int64_t bad = 0;
  size_t pos = 0;
  while (pos < len) {
    size_t e = pos;
    while (e < len && buf[e] != '\n' && buf[e] != '\r') e++;
    size_t next = e;
    if (next < len) next += (buf[e] == '\r' && e + 1 < len && buf[e + 1] == '\n') ? 2 : 1;
    line_no++;
    /* triplet.split("\\s+")  (URW:26 / VRW:21): leading whitespace yields an empty first token,
     * trailing empty tokens are dropped; an empty line yields [""] */
    const char *tok[64];
    size_t tlen[64];
    int nt = 0;
    size_t i = pos;
    if (i == e || is_java_ws(buf[i])) { tok[0] = buf + i; tlen[0] = 0; nt = 1; }
    while (i < e) {
      while (i < e && is_java_ws(buf[i])) i++;
      if (i == e) break;
      size_t s0 = i;
      while (i < e && !is_java_ws(buf[i])) i++;
      if (nt < 64) { tok[nt] = buf + s0; tlen[nt] = i - s0; }
      nt++;
    }
    if (nt > 64) nt = 64;
    int32_t s = 0, d = 0, pi = 0;
    float wi = 1.0f;
    /* (parts.head.toInt, parts(1).toInt)  URW:34 / VRW:34 */
    if (!parse_java_int(tok[0], tlen[0], &s)) {
      if (err) snprintf(err, errcap, "line %lld: NumberFormatException for src \"%.*s\"", (long long)line_no, (int)tlen[0], tok[0]);
      bad = line_no; break;
    }
    if (nt < 2) {
      if (err) snprintf(err, errcap, "line %lld: ArrayIndexOutOfBoundsException: 1", (long long)line_no);
      bad = line_no; break;
    }
    if (!parse_java_int(tok[1], tlen[1], &d)) {
      if (err) snprintf(err, errcap, "line %lld: NumberFormatException for dst \"%.*s\"", (long long)line_no, (int)tlen[1], tok[1]);
      bad = line_no; break;
    }
    if (!partitioned) {
      /* URW:29-32: weight = parts.last.toFloat iff weighted && parts.length > 2, else/unparsable 1.0f */
      if (weighted && nt > 2 && !parse_java_float(tok[nt - 1], tlen[nt - 1], &wi)) wi = 1.0f;
    } else {
      /* VRW:23-26: pid = parts(2).toInt iff partitioned && parts.length > 2, else random
       * (Random.nextInt(rddPartitions): unpinned; this build uses 0) */
      if (nt > 2 && !parse_java_int(tok[2], tlen[2], &pi)) pi = 0;
      /* VRW:29-32: weight only when there are more than 3 columns */
      if (weighted && nt > 3 && !parse_java_float(tok[nt - 1], tlen[nt - 1], &wi)) wi = 1.0f;
    }
    if (n == cap) {
      cap *= 2;
      src = (int32_t *)realloc(src, (size_t)cap * 4); dst = (int32_t *)realloc(dst, (size_t)cap * 4);
      pid = (int32_t *)realloc(pid, (size_t)cap * 4); w = (float *)realloc(w, (size_t)cap * 4);
    }
    src[n] = s; dst[n] = d; w[n] = wi; pid[n] = pi; n++;
    pos = next;
  }
  if (!bad) og_load_edges(g, n, src, dst, w, partitioned ? pid : NULL, directed);
  free(src); free(dst); free(pid); free(w);
  return bad;
}

/* ------------------------------------------------------------------------------------------ */
/* RandomSample (RS:5-63)                                                                      */
/* ------------------------------------------------------------------------------------------ */
int64_t oracle_sample(int64_t n, const float *w, float u) {
  /* RS:14  val sum = edges.foldLeft(0.0) { case (w1, (_, w2)) => w1 + w2 }   (Double + Float) */
  double sum = 0.0;
  for (int64_t i = 0; i < n; ++i) sum = sum + (double)w[i];
  /* RS:16-22  acc += w / sum (Float / Double -> Double);  if (acc >= p) return   (Double >= Float) */
  double acc = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    acc += (double)w[i] / sum;
    if (acc >= (double)u) return i;
  }
  return 0; /* RS:24 edges.head */
}

void oracle_second_order_weights(float p, float q, int32_t prev, int64_t np, const int32_t *pdst,
                                 int64_t nc, const int32_t *cdst, const float *cw, float *out) {
  /* RS:33-41 */
  for (int64_t i = 0; i < nc; ++i) {
    int32_t dst_id = cdst[i];
    float w = cw[i];
    float unnorm = w / q;                       /* RS:34 float32 division */
    if (dst_id == prev) unnorm = w / p;         /* RS:36 */
    else {
      int exists = 0;                           /* RS:38 prevNeighbors.exists(_._1 == dstId) */
      for (int64_t j = 0; j < np; ++j)
        if (pdst[j] == dst_id) { exists = 1; break; }
      if (exists) unnorm = w;
    }
    out[i] = unnorm;
  }
}

int64_t oracle_second_order_sample(float p, float q, int32_t prev, int64_t np, const int32_t *pdst,
                                   int64_t nc, const int32_t *cdst, const float *cw, float u,
                                   float *w_out) {
  /* RS:60-61 */
  float stackbuf[256];
  float *nw = nc <= 256 ? stackbuf : (float *)malloc((size_t)nc * sizeof(float));
  oracle_second_order_weights(p, q, prev, np, pdst, nc, cdst, cw, nw);
  int64_t k = oracle_sample(nc, nw, u);
  if (w_out) *w_out = nw[k];
  if (nw != stackbuf) free(nw);
  return k;
}

/* ------------------------------------------------------------------------------------------ */
/* walk driver: RW:51-66 + RW:75-176 with local[*] semantics (one GraphMap holds every vertex,  */
/* so RW:121-129 never fires) == the test helper T-URW:293-321.                                 */
/* ------------------------------------------------------------------------------------------ */
static int32_t walk_one(const og_graph *g, const oracle_walk_cfg *cfg, int32_t start, uint64_t walker,
                        int32_t *path) {
  const float p = (float)cfg->p, q = (float)cfg->q;       /* RW:112-113 .toFloat */
  const int32_t full = cfg->walk_length + 2;               /* RW:103,132 */
  int32_t len = 0;
  path[len++] = start;
  const int32_t *nd; const float *nw;
  int64_t deg = og_neighbors(g, start, &nd, &nw);
  if (deg <= 0) return len;                                /* RW:59-62 dead end: path = [v] */
  float u = cfg->u_mode == ORACLE_U_CONST ? cfg->u_const : oracle_u01(cfg->seed, walker, 0u);
  path[len++] = nd[oracle_sample(deg, nw, u)];             /* RW:57-58 */
  while (len != full) {                                    /* RW:103 */
    int32_t curr = path[len - 1], prev = path[len - 2];
    const int32_t *cd; const float *cw;
    int64_t dc = og_neighbors(g, curr, &cd, &cw);          /* RW:104 */
    if (dc <= 0) break;                                    /* RW:115-119 dead end */
    const int32_t *pd; const float *pw;
    int64_t dp = og_neighbors(g, prev, &pd, &pw);          /* RW:106-109 */
    if (dp < 0) dp = 0;
    u = cfg->u_mode == ORACLE_U_CONST ? cfg->u_const : oracle_u01(cfg->seed, walker, (uint32_t)(len - 1));
    int64_t k = oracle_second_order_sample(p, q, prev, dp, pd, dc, cd, cw, u, NULL); /* RW:112-113 */
    path[len++] = cd[k];                                   /* RW:114 */
  }
  return len;
}

int64_t oracle_walk(const og_graph *g, const oracle_walk_cfg *cfg, int32_t *ids, int64_t ids_cap,
                    int64_t *offsets) {
  int64_t nv = og_num_vertices(g);
  int32_t *vids = (int32_t *)malloc((size_t)(nv ? nv : 1) * sizeof(int32_t));
  og_vertex_ids(g, vids, nv);
  const int32_t stride = cfg->walk_length + 2;
  int32_t *tmp = (int32_t *)malloc((size_t)(nv ? nv : 1) * (size_t)stride * sizeof(int32_t));
  int32_t *lens = (int32_t *)malloc((size_t)(nv ? nv : 1) * sizeof(int32_t));
  int64_t n_paths = 0, total = 0;
  int overflow = 0;
  const int64_t mod = cfg->sample_mod > 1 ? cfg->sample_mod : 1;
#ifdef _OPENMP
  int nthreads = cfg->threads > 0 ? cfg->threads : omp_get_max_threads();
#endif
  for (int32_t round = 0; round < cfg->num_walks; ++round) {   /* RW:82 */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
#endif
    for (int64_t v = 0; v < nv; ++v) {
      uint64_t walker = (uint64_t)round * (uint64_t)nv + (uint64_t)v;
      if (walker % (uint64_t)mod) { lens[v] = 0; continue; }
      lens[v] = walk_one(g, cfg, vids[v], walker, tmp + (size_t)v * stride);
    }
    for (int64_t v = 0; v < nv; ++v) {
      if (!lens[v]) continue;
      offsets[n_paths] = total;
      if (total + lens[v] <= ids_cap) memcpy(ids + total, tmp + (size_t)v * stride, (size_t)lens[v] * 4);
      else overflow = 1;
      total += lens[v];
      n_paths++;
    }
  }
  offsets[n_paths] = total;
  free(lens); free(tmp); free(vids);
  if (overflow) { offsets[0] = total; return -1; }
  return n_paths;
}

int64_t oracle_format_paths(int64_t n_paths, const int32_t *ids, const int64_t *offsets, char *out,
                            int64_t cap) {
  /* RW:234-241: path.mkString("\t"), one path per line */
  int64_t pos = 0;
  char num[16];
  for (int64_t i = 0; i < n_paths; ++i) {
    for (int64_t k = offsets[i]; k < offsets[i + 1]; ++k) {
      int m = snprintf(num, sizeof(num), "%d", ids[k]);
      if (k > offsets[i]) { if (pos < cap) out[pos] = '\t'; pos++; }
      if (pos + m <= cap) memcpy(out + pos, num, (size_t)m);
      pos += m;
    }
    if (pos < cap) out[pos] = '\n';
    pos++;
  }
  return pos;
}

/* ------------------------------------------------------------------------------------------ */
/* CPU twin of the product's alias sampler (definition in DESIGN.md "alias mode")              */
/* ------------------------------------------------------------------------------------------ */
struct oa_graph {
  int64_t nv;
  int32_t *vids;      /* ascending; dense index = rank */
  int64_t *offsets;   /* nv+1 */
  int32_t *col;       /* dense neighbour index, each row sorted ascending (ties: appearance order) */
  float *w;
  uint32_t *thr;      /* Vose bucket threshold (NULL when every weight is 1.0f) */
  uint32_t *alias;    /* row-relative alias slot */
  uint32_t *mult;     /* parallel-edge multiplicity of every entry */
  double *wsum;       /* per row: sequential double sum of the weights (the W of the Vose build); weighted graphs */
  double *wb;         /* per entry: double sum, in row order, of the weights of all parallel entries to the same neighbour */
  int has_alias;
  int directed;       /* set by the caller: folding is defined for undirected graphs only */
};

typedef struct { int32_t c; float w; int64_t pos; } oa_ent;
static int cmp_ent(const void *a, const void *b) {
  const oa_ent *x = (const oa_ent *)a, *y = (const oa_ent *)b;
  if (x->c != y->c) return (x->c > y->c) - (x->c < y->c);
  return (x->pos > y->pos) - (x->pos < y->pos);
}
static int64_t rank_of(const int32_t *vids, int64_t nv, int32_t v) {
  int64_t lo = 0, hi = nv;
  while (lo < hi) { int64_t m = (lo + hi) >> 1; if (vids[m] < v) lo = m + 1; else hi = m; }
  return lo;
}

/* Vose alias table by the in-order "sweep" (two cursors, no work lists), in EXACT INTEGER arithmetic so that the result does not
 * depend on the order of evaluation and the device can build it in parallel (graph_build.cu K3):
 *   W      = the row's weight sum: 32 strided partial sums (partial l adds w[l], w[l+32], ... in order, in double) combined by a
 *            butterfly (distance 16, 8, 4, 2, 1) -- the one floating-point reduction, in a fixed shape;
 *   t[k]   = floor((((double)w[k] * (double)n) / W) * 2^32) as a 64-bit integer: the scaled weight in units of 2^-32;
 *            k is light iff t[k] < 2^32;
 *   sweep  : i walks the light items upward, j the heavy ones; r is heavy j's residual (an integer). */
static double alias_row_wsum(int64_t n, const float *w) {
  double part[32], nx[32];
  for (int l = 0; l < 32; ++l) part[l] = 0.0;
  for (int64_t k = 0; k < n; ++k) part[k & 31] = part[k & 31] + (double)w[k];
  for (int o = 16; o > 0; o >>= 1) {
    for (int l = 0; l < 32; ++l) nx[l] = part[l] + part[l ^ o];
    for (int l = 0; l < 32; ++l) part[l] = nx[l];
  }
  return part[0];
}
static double alias_row(int64_t n, const float *w, uint32_t *thr, uint32_t *alias) {
  const double W = alias_row_wsum(n, w);
  for (int64_t k = 0; k < n; ++k) { thr[k] = 0xFFFFFFFFu; alias[k] = (uint32_t)k; }
  const double dn = (double)n;
  const uint64_t ONE = 4294967296ULL;
#define SCALED(k) ((uint64_t)((((double)w[(k)] * dn) / W) * 4294967296.0))
  int64_t i = 0, j = 0;
  while (i < n && !(SCALED(i) < ONE)) i++;
  while (j < n && (SCALED(j) < ONE)) j++;
  if (j >= n) return W;
  uint64_t r = SCALED(j);
  while (j < n) {
    if (!(r < ONE)) {
      if (i >= n) break;
      const uint64_t ti = SCALED(i);
      thr[i] = (uint32_t)ti;
      alias[i] = (uint32_t)j;
      r = (r + ti) - ONE;
      i++;
      while (i < n && !(SCALED(i) < ONE)) i++;
    } else {
      int64_t j2 = j + 1;
      while (j2 < n && (SCALED(j2) < ONE)) j2++;
      if (j2 >= n) break;
      thr[j] = (uint32_t)r;
      alias[j] = (uint32_t)j2;
      r = (r + SCALED(j2)) - ONE;
      j = j2;
    }
  }
#undef SCALED
  return W;
}

oa_graph *oa_build(const og_graph *g) {
  oa_graph *a = (oa_graph *)calloc(1, sizeof(oa_graph));
  a->nv = og_num_vertices(g);
  a->vids = (int32_t *)malloc((size_t)(a->nv ? a->nv : 1) * 4);
  og_vertex_ids(g, a->vids, a->nv);
  a->offsets = (int64_t *)malloc((size_t)(a->nv + 1) * 8);
  a->offsets[0] = 0;
  for (int64_t v = 0; v < a->nv; ++v) {
    int64_t d = og_neighbors(g, a->vids[v], NULL, NULL);
    a->offsets[v + 1] = a->offsets[v] + (d > 0 ? d : 0);
  }
  int64_t nnz = a->offsets[a->nv];
  a->col = (int32_t *)malloc((size_t)(nnz ? nnz : 1) * 4);
  a->w = (float *)malloc((size_t)(nnz ? nnz : 1) * 4);
  int has = 0;
  for (int64_t v = 0; v < a->nv; ++v) {
    const int32_t *nd; const float *nw;
    int64_t d = og_neighbors(g, a->vids[v], &nd, &nw);
    if (d <= 0) continue;
    oa_ent *e = (oa_ent *)malloc((size_t)d * sizeof(oa_ent));
    for (int64_t k = 0; k < d; ++k) {
      e[k].c = (int32_t)rank_of(a->vids, a->nv, nd[k]);
      e[k].w = nw[k]; e[k].pos = k;
      if (nw[k] != 1.0f) has = 1;
    }
    qsort(e, (size_t)d, sizeof(oa_ent), cmp_ent);
    for (int64_t k = 0; k < d; ++k) { a->col[a->offsets[v] + k] = e[k].c; a->w[a->offsets[v] + k] = e[k].w; }
    free(e);
  }
  a->has_alias = has;
  a->mult = (uint32_t *)malloc((size_t)(nnz ? nnz : 1) * 4);
  for (int64_t v = 0; v < a->nv; ++v) {
    const int64_t lo = a->offsets[v], hi = a->offsets[v + 1];
    for (int64_t i = lo; i < hi;) {
      int64_t j = i;
      while (j < hi && a->col[j] == a->col[i]) j++;
      for (int64_t k = i; k < j; ++k) a->mult[k] = (uint32_t)(j - i);
      i = j;
    }
  }
  if (has) {
    a->thr = (uint32_t *)malloc((size_t)(nnz ? nnz : 1) * 4);
    a->alias = (uint32_t *)malloc((size_t)(nnz ? nnz : 1) * 4);
    a->wsum = (double *)calloc((size_t)(a->nv ? a->nv : 1), 8);
    a->wb = (double *)malloc((size_t)(nnz ? nnz : 1) * 8);
    for (int64_t v = 0; v < a->nv; ++v) {
      int64_t d = a->offsets[v + 1] - a->offsets[v];
      if (d > 0) a->wsum[v] = alias_row(d, a->w + a->offsets[v], a->thr + a->offsets[v], a->alias + a->offsets[v]);
      const int64_t lo = a->offsets[v], hi = a->offsets[v + 1];
      for (int64_t i = lo; i < hi;) {                 /* bundle weight: the run of equal neighbours, summed in row order */
        int64_t j = i;
        double sum = 0.0;
        while (j < hi && a->col[j] == a->col[i]) { sum = sum + (double)a->w[j]; j++; }
        for (int64_t k = i; k < j; ++k) a->wb[k] = sum;
        i = j;
      }
    }
  }
  return a;
}
void oa_free(oa_graph *a) {
  if (!a) return;
  free(a->vids); free(a->offsets); free(a->col); free(a->w); free(a->thr); free(a->alias); free(a->mult); free(a->wsum); free(a->wb); free(a);
}
int64_t oa_num_vertices(const oa_graph *a) { return a->nv; }
int oa_has_alias(const oa_graph *a) { return a->has_alias; }
const uint32_t *oa_mult(const oa_graph *a) { return a->mult; }
const double *oa_wsum(const oa_graph *a) { return a->wsum; }
const double *oa_wbundle(const oa_graph *a) { return a->wb; }
void oa_set_directed(oa_graph *a, int directed) { a->directed = directed; }
void oa_view(const oa_graph *a, const int32_t **vids, const int64_t **offsets, const int32_t **col,
             const float **w, const uint32_t **thr, const uint32_t **alias) {
  if (vids) *vids = a->vids;
  if (offsets) *offsets = a->offsets;
  if (col) *col = a->col;
  if (w) *w = a->w;
  if (thr) *thr = a->thr;
  if (alias) *alias = a->alias;
}

void oracle_alias_thresholds(double p, double q, uint64_t *t_ret, uint64_t *t_common, uint64_t *t_far) {
  const double inv_p = 1.0 / (double)(float)p, inv_q = 1.0 / (double)(float)q;
  double M = inv_p > 1.0 ? inv_p : 1.0;
  if (inv_q > M) M = inv_q;
#define THR(f) ((f) >= M ? 4294967296ULL : (uint64_t)(((f) / M) * 4294967296.0))
  *t_ret = THR(inv_p);
  *t_common = THR(1.0);
  *t_far = THR(inv_q);
#undef THR
}

void oracle_fold_thresholds(double p, double q, uint64_t *t_common, uint64_t *t_far, double *a, double *mp) {
  const double inv_p = 1.0 / (double)(float)p, inv_q = 1.0 / (double)(float)q;
  const double M = inv_q > 1.0 ? inv_q : 1.0;
#define THR(f) ((f) >= M ? 4294967296ULL : (uint64_t)(((f) / M) * 4294967296.0))
  *t_common = THR(1.0);
  *t_far = THR(inv_q);
#undef THR
  *a = inv_p - M;
  *mp = M;
}

static inline uint64_t mulhi64(uint64_t a, uint64_t b) {
  return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
}
static inline int64_t alias_pick(const oa_graph *a, int64_t off, int64_t deg, const uint32_t r[4]) {
  uint64_t R = ((uint64_t)r[0] << 32) | (uint64_t)r[3];
  int64_t k = (int64_t)mulhi64(R, (uint64_t)deg);
  if (a->has_alias && !(r[1] < a->thr[off + k])) k = (int64_t)a->alias[off + k];
  return k;
}
static inline int row_contains(const int32_t *row, int64_t n, int32_t x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) { int64_t m = (lo + hi) >> 1; if (row[m] < x) lo = m + 1; else hi = m; }
  return lo < n && row[lo] == x;
}
static inline int ceil_log2_p1(int64_t d) { /* ceil(log2(d+1)) */
  int l = 0;
  while (((int64_t)1 << l) < d + 1) l++;
  return l;
}

int64_t oracle_alias_walk(const oa_graph *a, const oracle_walk_cfg *cfg, int32_t *ids, int64_t ids_cap,
                          int64_t *offsets, oracle_alias_stats *stats) {
  const int64_t nv = a->nv;
  const int32_t stride = cfg->walk_length + 2;
  uint64_t t_ret, t_common, t_far;
  oracle_alias_thresholds(cfg->p, cfg->q, &t_ret, &t_common, &t_far);
  double fold_a = 0.0, fold_mp = 1.0;
  int fold = 0;
  int wfold = 0;             /* weighted alias-fold: the return draw and the accept draw share r[2] (see below) */
  if (cfg->fold && !a->directed) {
    uint64_t fc, ff;
    oracle_fold_thresholds(cfg->p, cfg->q, &fc, &ff, &fold_a, &fold_mp);
    if (fold_a > 0.0) { fold = 1; wfold = a->has_alias; t_common = fc; t_far = ff; t_ret = 4294967296ULL; }
  }
  int32_t *tmp = (int32_t *)malloc((size_t)(nv ? nv : 1) * (size_t)stride * 4);
  int32_t *lens = (int32_t *)malloc((size_t)(nv ? nv : 1) * 4);
  int64_t n_paths = 0, total = 0, st_steps = 0, st_prop = 0, st_log = 0, st_mem = 0;
  int overflow = 0;
  const int64_t mod = cfg->sample_mod > 1 ? cfg->sample_mod : 1;
#ifdef _OPENMP
  int nthreads = cfg->threads > 0 ? cfg->threads : omp_get_max_threads();
#endif
  for (int32_t round = 0; round < cfg->num_walks; ++round) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64) num_threads(nthreads) reduction(+ : st_steps, st_prop, st_log, st_mem)
#endif
    for (int64_t v = 0; v < nv; ++v) {
      uint64_t walker = (uint64_t)round * (uint64_t)nv + (uint64_t)v;
      if (walker % (uint64_t)mod) { lens[v] = 0; continue; }
      int32_t *path = tmp + (size_t)v * stride;
      int32_t len = 0;
      path[len++] = (int32_t)v;
      int64_t off = a->offsets[v], deg = a->offsets[v + 1] - off;
      uint32_t r[4];
      if (deg > 0) {
        walker_rng(cfg->seed, walker, 0u, 0u, r);           /* first-order step: proposal accepted */
        int64_t kk = alias_pick(a, off, deg, r);
        uint32_t m_ret = a->mult[off + kk];                  /* parallel edges start<->first (undirected: symmetric) */
        double w_ret = wfold ? a->wb[off + kk] : 0.0;        /* their total weight (weighted fold) */
        path[len++] = a->col[off + kk];
        st_steps++;
        while (len != stride) {
          int32_t curr = path[len - 1], prev = path[len - 2];
          off = a->offsets[curr]; deg = a->offsets[curr + 1] - off;
          if (deg <= 0) break;
          const int64_t poff = a->offsets[prev], pdeg = a->offsets[prev + 1] - poff;
          int32_t x = -1;
          /* fold: one trial picks the return-excess component with probability
           * a*m / (Mp*deg + a*m) (always accepted), else a uniform entry under the envelope Mp */
          /* division-free form: return iff r1 * (Mp*deg + a*m) < a*m * 2^32 (IEEE double, one
           * rounding per operation; the kernel evaluates the same expression) */
          double ret_lhs = 0.0, ret_rhs = 0.0, wf_t2 = 0.0;
          if (fold && !wfold) {
            const double t1 = fold_a * (double)m_ret, t2 = fold_mp * (double)deg;
            ret_lhs = t2 + t1;
            ret_rhs = t1 * 4294967296.0;
          }
          if (wfold) {
            /* weighted: mass of the return-excess component a*wb against the envelope Mp*W(curr).  ONE draw z = r[2]
             * decides both: z*den < t1*2^32 -> return; else the proposal of class c is accepted iff
             * z*den < t1*2^32 + t2*T_c  (z rescaled to the non-return part of [0,1); r[1] stays the Vose coin). */
            const double t1 = fold_a * w_ret;
            wf_t2 = fold_mp * a->wsum[curr];
            ret_lhs = wf_t2 + t1;
            ret_rhs = t1 * 4294967296.0;
          }
          for (uint32_t trial = 0;; ++trial) {
            walker_rng(cfg->seed, walker, (uint32_t)(len - 1), trial, r);
            if (wfold) {
              const double zl = (double)r[2] * ret_lhs;
              if (zl < ret_rhs) { x = prev; kk = -1; st_prop++; break; }
              kk = alias_pick(a, off, deg, r);
              x = a->col[off + kk];
              st_prop++;
              if (x == prev || deg == 1) break;                 /* mass Mp of Mp under the envelope: always accepted */
              const double rc = ret_rhs + wf_t2 * (double)t_common, rf = ret_rhs + wf_t2 * (double)t_far;
              const double rlo = rc < rf ? rc : rf, rhi = rc < rf ? rf : rc;
              if (zl < rlo) break;
              if (!(zl < rhi)) continue;
              st_mem++; st_log += ceil_log2_p1(pdeg);
              if (zl < (row_contains(a->col + poff, pdeg, x) ? rc : rf)) break;
              continue;
            }
            if (fold && (double)r[1] * ret_lhs < ret_rhs) { x = prev; kk = -1; st_prop++; break; }
            kk = alias_pick(a, off, deg, r);
            x = a->col[off + kk];
            st_prop++;
            uint64_t t;
            if (x == prev) t = t_ret;
            else if (t_common == t_far) t = t_far;
            else {
              /* count a membership test only where the outcome can matter (same rule as the kernel) */
              uint64_t lo = t_common < t_far ? t_common : t_far, hi = t_common < t_far ? t_far : t_common;
              if ((uint64_t)r[2] >= lo && (uint64_t)r[2] < hi) { st_mem++; st_log += ceil_log2_p1(pdeg); }
              t = row_contains(a->col + poff, pdeg, x) ? t_common : t_far;
            }
            if ((uint64_t)r[2] < t) break;
          }
          if (wfold && kk >= 0) w_ret = a->wb[off + kk];         /* a direct return keeps the bundle (symmetric) */
          if (fold) {
            if (kk >= 0) m_ret = a->mult[off + kk];
            else {                                            /* direct return: multiplicity of prev in N(curr) */
              int64_t lo = 0, hi = deg;
              while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a->col[off + mid] < x) lo = mid + 1; else hi = mid; }
              m_ret = a->mult[off + lo];
            }
          }
          path[len++] = x;
          st_steps++;
        }
      }
      lens[v] = len;
    }
    for (int64_t v = 0; v < nv; ++v) {
      if (!lens[v]) continue;
      offsets[n_paths] = total;
      if (total + lens[v] <= ids_cap) {
        for (int32_t k = 0; k < lens[v]; ++k) ids[total + k] = a->vids[tmp[(size_t)v * stride + k]];
      } else overflow = 1;
      total += lens[v];
      n_paths++;
    }
  }
  offsets[n_paths] = total;
  if (stats) { stats->steps = st_steps; stats->proposals = st_prop; stats->probes_log2 = st_log; stats->member_tests = st_mem; }
  free(lens); free(tmp);
  if (overflow) { offsets[0] = total; return -1; }
  return n_paths;
}

/* ------------------------------------------------------------------------------------------ */
/* CPU-baseline helpers (bench.py only)                                                        */
/* ------------------------------------------------------------------------------------------ */
#include <time.h>
static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_rmat_edges(int scale, uint64_t seed, int64_t first, int64_t count, int32_t *src, int32_t *dst, int threads) {
  const uint32_t A = (uint32_t)(0.57 * 4294967296.0), AB = (uint32_t)((0.57 + 0.19) * 4294967296.0),
                 ABC = (uint32_t)((0.57 + 0.19 + 0.19) * 4294967296.0);
  (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
  for (int64_t i = 0; i < count; ++i) {
    const uint64_t e = (uint64_t)(first + i);
    uint32_t s = 0, d = 0;
    for (int blk = 0; blk * 4 < scale; ++blk) {
      uint32_t ctr[4] = {(uint32_t)e, (uint32_t)(e >> 32), (uint32_t)blk, 0x524D4154u}, key[2] = {(uint32_t)seed, 0u}, r[4];
      oracle_philox4x32_10(ctr, key, r);
      for (int k = 0; k < 4 && blk * 4 + k < scale; ++k) {
        const uint32_t x = r[k];
        s = (s << 1) | (x >= AB ? 1u : 0u);
        d = (d << 1) | (((x >= A && x < AB) || x >= ABC) ? 1u : 0u);
      }
    }
    src[i] = (int32_t)s;
    dst[i] = (int32_t)d;
  }
}

void oracle_csr_build(int64_t n_ids, int64_t n_edges, const int32_t *src, const int32_t *dst, int64_t *offsets,
                      int32_t *col, int threads) {
  (void)threads;
  int64_t *cursor = (int64_t *)calloc((size_t)n_ids + 1, sizeof(int64_t));
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
  for (int64_t e = 0; e < n_edges; ++e) {
#ifdef _OPENMP
#pragma omp atomic
#endif
    cursor[src[e]]++;
#ifdef _OPENMP
#pragma omp atomic
#endif
    cursor[dst[e]]++;
  }
  offsets[0] = 0;
  for (int64_t v = 0; v < n_ids; ++v) offsets[v + 1] = offsets[v] + cursor[v];
  memcpy(cursor, offsets, (size_t)n_ids * sizeof(int64_t));
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : omp_get_max_threads())
#endif
  for (int64_t e = 0; e < n_edges; ++e) {
    int64_t k;
#ifdef _OPENMP
#pragma omp atomic capture
#endif
    k = cursor[src[e]]++;
    col[k] = dst[e];
#ifdef _OPENMP
#pragma omp atomic capture
#endif
    k = cursor[dst[e]]++;
    col[k] = src[e];
  }
  free(cursor);
}

int64_t oracle_walk_csr_timed(int64_t nv, const int64_t *offsets, const int32_t *col, const float *w,
                              const oracle_walk_cfg *cfg, int64_t sample_stride, int64_t sample_phase,
                              double budget_s, double *elapsed_s, int64_t *walkers_done, uint64_t *checksum) {
  const float p = (float)cfg->p, q = (float)cfg->q;
  const int32_t full = cfg->walk_length + 2;
  if (sample_stride < 1) sample_stride = 1;
  const int64_t n_samples = (nv - sample_phase + sample_stride - 1) / sample_stride;
  int64_t steps = 0, done = 0;
  uint64_t sum = 0;
  const double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel num_threads(cfg->threads > 0 ? cfg->threads : omp_get_max_threads()) reduction(+ : steps, done, sum)
#endif
  {
    int32_t *path = (int32_t *)malloc((size_t)full * sizeof(int32_t));
    float *nw = NULL, *ones = NULL;
    int64_t nw_cap = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (int64_t sidx = 0; sidx < n_samples; ++sidx) {
      if (now_s() - t0 > budget_s) continue;
      const int64_t v = sample_phase + sidx * sample_stride;
      const uint64_t walker = (uint64_t)v;
      int32_t len = 0;
      path[len++] = (int32_t)v;
      int64_t off = offsets[v], deg = offsets[v + 1] - off;
      if (deg > 0) {
        if (deg > nw_cap) { nw_cap = deg * 2; nw = (float *)realloc(nw, (size_t)nw_cap * 4); ones = (float *)realloc(ones, (size_t)nw_cap * 4); for (int64_t k = 0; k < nw_cap; ++k) ones[k] = 1.0f; }
        path[len++] = col[off + oracle_sample(deg, w ? w + off : ones, oracle_u01(cfg->seed, walker, 0u))];
        steps++;
        while (len != full) {
          const int32_t curr = path[len - 1], prev = path[len - 2];
          off = offsets[curr]; deg = offsets[curr + 1] - off;
          if (deg <= 0) break;
          if (deg > nw_cap) { nw_cap = deg * 2; nw = (float *)realloc(nw, (size_t)nw_cap * 4); ones = (float *)realloc(ones, (size_t)nw_cap * 4); for (int64_t k = 0; k < nw_cap; ++k) ones[k] = 1.0f; }
          const int64_t poff = offsets[prev], pdeg = offsets[prev + 1] - poff;
          /* RS:27-44, in slices of neighbours so that one hub-by-hub step (d_c*d_p ~ 1e12 compares on
           * RMAT-26) cannot overrun the time budget: an unfinished step is abandoned and not counted */
          int timed_out = 0;
          for (int64_t c0 = 0; c0 < deg; c0 += 256) {
            const int64_t c1 = c0 + 256 < deg ? c0 + 256 : deg;
            oracle_second_order_weights(p, q, prev, pdeg, col + poff, c1 - c0, col + off + c0, (w ? w + off : ones) + c0, nw + c0);
            if (now_s() - t0 > budget_s) { timed_out = 1; break; }
          }
          if (timed_out) break;
          const float u = oracle_u01(cfg->seed, walker, (uint32_t)(len - 1));
          path[len++] = col[off + oracle_sample(deg, nw, u)];                                                /* RS:12-25 */
          steps++;
          if (now_s() - t0 > budget_s) break;
        }
      }
      for (int32_t k = 0; k < len; ++k) sum += (uint64_t)(uint32_t)path[k] * (uint64_t)(k + 1);
      done++;
    }
    free(path); free(nw); free(ones);
  }
  if (elapsed_s) *elapsed_s = now_s() - t0;
  if (walkers_done) *walkers_done = done;
  if (checksum) *checksum = sum;
  return steps;
}

/* The product's alias-fold algorithm (NOT the reference's: see oracle_alias_walk, whose decisions this repeats) over a
 * dense, neighbour-sorted, UNWEIGHTED CSR -- bench.py's "optimised CPU twin" figure: the same sampler
 * the GPU kernel runs, on the host cores, membership by binary search, multiplicity by scanning the run of equal
 * neighbours.  Walker v of round 0 for every sampled start vertex; stops at the time budget. */
int64_t oracle_fold_walk_csr_timed(int64_t nv, const int64_t *offsets, const int32_t *col, const oracle_walk_cfg *cfg,
                                   int64_t sample_stride, int64_t sample_phase, double budget_s, double *elapsed_s,
                                   int64_t *walkers_done, uint64_t *checksum, int32_t *paths_out) {
  const int32_t full = cfg->walk_length + 2;
  uint64_t t_ret, t_common, t_far;
  oracle_alias_thresholds(cfg->p, cfg->q, &t_ret, &t_common, &t_far);
  double fold_a = 0.0, fold_mp = 1.0;
  int fold = 0;
  {
    uint64_t fc, ff;
    oracle_fold_thresholds(cfg->p, cfg->q, &fc, &ff, &fold_a, &fold_mp);
    if (fold_a > 0.0) { fold = 1; t_common = fc; t_far = ff; t_ret = 4294967296ULL; }
  }
  if (sample_stride < 1) sample_stride = 1;
  const int64_t n_samples = (nv - sample_phase + sample_stride - 1) / sample_stride;
  int64_t steps = 0, done = 0;
  uint64_t sum = 0;
  const double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel num_threads(cfg->threads > 0 ? cfg->threads : omp_get_max_threads()) reduction(+ : steps, done, sum)
#endif
  {
    int32_t *path = (int32_t *)malloc((size_t)full * sizeof(int32_t));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
    for (int64_t sidx = 0; sidx < n_samples; ++sidx) {
      if (now_s() - t0 > budget_s) continue;
      const int64_t v = sample_phase + sidx * sample_stride;
      const uint64_t walker = (uint64_t)v;
      int32_t len = 0;
      path[len++] = (int32_t)v;
      int64_t off = offsets[v], deg = offsets[v + 1] - off;
      uint32_t r[4];
      if (deg > 0) {
        walker_rng(cfg->seed, walker, 0u, 0u, r);
        int64_t kk = (int64_t)mulhi64(((uint64_t)r[0] << 32) | (uint64_t)r[3], (uint64_t)deg);
        /* multiplicity of the chosen neighbour in this (sorted) row */
#define RUN_LEN(o, d, k, out)                                                         \
  do {                                                                                \
    int64_t a_ = (k), b_ = (k);                                                       \
    while (a_ > 0 && col[(o) + a_ - 1] == col[(o) + (k)]) a_--;                       \
    while (b_ + 1 < (d) && col[(o) + b_ + 1] == col[(o) + (k)]) b_++;                 \
    (out) = (uint32_t)(b_ - a_ + 1);                                                  \
  } while (0)
        uint32_t m_ret;
        RUN_LEN(off, deg, kk, m_ret);
        path[len++] = col[off + kk];
        steps++;
        while (len != full) {
          const int32_t curr = path[len - 1], prev = path[len - 2];
          off = offsets[curr]; deg = offsets[curr + 1] - off;
          if (deg <= 0) break;
          const int64_t poff = offsets[prev], pdeg = offsets[prev + 1] - poff;
          int32_t x = -1;
          double ret_lhs = 0.0, ret_rhs = 0.0;
          if (fold) {
            const double t1 = fold_a * (double)m_ret, t2 = fold_mp * (double)deg;
            ret_lhs = t2 + t1;
            ret_rhs = t1 * 4294967296.0;
          }
          for (uint32_t trial = 0;; ++trial) {
            walker_rng(cfg->seed, walker, (uint32_t)(len - 1), trial, r);
            if (fold && (double)r[1] * ret_lhs < ret_rhs) { x = prev; kk = -1; break; }
            kk = (int64_t)mulhi64(((uint64_t)r[0] << 32) | (uint64_t)r[3], (uint64_t)deg);
            x = col[off + kk];
            uint64_t t;
            if (x == prev) t = t_ret;
            else if (t_common == t_far) t = t_far;
            else {
              const uint64_t lo = t_common < t_far ? t_common : t_far, hi = t_common < t_far ? t_far : t_common;
              if ((uint64_t)r[2] < lo) break;                       /* accepted whatever the class */
              if ((uint64_t)r[2] >= hi) continue;                   /* rejected whatever the class */
              t = row_contains(col + poff, pdeg, x) ? t_common : t_far;
            }
            if ((uint64_t)r[2] < t) break;
          }
          if (fold && kk >= 0) RUN_LEN(off, deg, kk, m_ret);        /* a direct return keeps the bundle (symmetric) */
          path[len++] = x;
          steps++;
        }
#undef RUN_LEN
      }
      for (int32_t k = 0; k < len; ++k) sum += (uint64_t)(uint32_t)path[k] * (uint64_t)(k + 1);
      if (paths_out) {
        for (int32_t k = 0; k < full; ++k) paths_out[sidx * full + k] = k < len ? path[k] : -1;
      }
      done++;
    }
    free(path);
  }
  if (elapsed_s) *elapsed_s = now_s() - t0;
  if (walkers_done) *walkers_done = done;
  if (checksum) *checksum = sum;
  return steps;
}
