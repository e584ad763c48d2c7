// srw_host.cpp -- host-side mirror of the reference's boundary for the walk path:
//   Params defaults (Params.scala:7-23), the scopt option table (CommandParser.scala:32-109),
//   edge-list line parsing with the JVM's rules (URW:26-34, VRW:21-34), the GraphMap.addVertex
//   first-wins builder (GM:23-56) and save() (RW:234-241).  No device code here.
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "srw_internal.h"

static thread_local std::string t_error;

void srw_set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_error = buf;
}

extern "C" const char *srw_last_error(void) { return t_error.c_str(); }
extern "C" const char *srw_version(void) { return "stellar-rw-b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------
// Params / CommandParser
// ------------------------------------------------------------------------------------------
extern "C" srw_status srw_params_default(srw_params *o) {
  if (!o) return SRW_ERR_ARG;
  memset(o, 0, sizeof(*o));
  o->w2v_iter = 10; o->w2v_lr = 0.025; o->w2v_partitions = 1; o->w2v_dim = 128; o->w2v_window = 10;  // Params:7-11
  o->walk_length = 80; o->num_walks = 10; o->p = 1.0; o->q = 1.0;                                      // Params:12-15
  o->weighted = 1; o->directed = 0;                                                                    // Params:16-17
  o->rdd_partitions = 200; o->single_output = 1; o->partitioned = 0; o->cmd = SRW_TASK_NODE2VEC;       // Params:20-23
  o->seed = 1; o->sampler = SRW_SAMPLER_ALIAS_FOLD; o->u_mode = SRW_U_PHILOX; o->u_const = 0.f; o->num_gpus = 1;
  return SRW_OK;
}

extern "C" const char *srw_usage(void) {
  // CP:33-105 (scopt renders "--name <value>" + the .text string)
  return "Main\n"
         "Usage: 2nd Order Random Walk + Word2Vec [options]\n\n"
         "  --walkLength <value>     walkLength: 80\n"
         "  --numWalks <value>       numWalks: 10\n"
         "  --p <value>              return parameter p: 1.0\n"
         "  --q <value>              in-out parameter q: 1.0\n"
         "  --rddPartitions <value>  Number of RDD partitions in running Random Walk and Word2vec: 200\n"
         "  --weighted <value>       weighted: true\n"
         "  --directed <value>       directed: false\n"
         "  --singleOutput <value>   generate single output file: true\n"
         "  --w2vPartitions <value>  Number of partitions in word2vec: 1\n"
         "  --input <value>          Input edge file path: empty\n"
         "  --output <value>         Output path: empty\n"
         "  --cmd <value>            command: node2vec\n"
         "  --partitioned <value>    Whether the graph is partitioned: false\n"
         "  --lr <value>             Learning rate in word2vec: 0.025\n"
         "  --iter <value>           Number of iterations in word2vec: 10\n"
         "  --dim <value>            Number of dimensions in word2vec: 128\n"
         "  --window <value>         Window size in word2vec: 10\n"
         "  --seed <value>           [b200] Philox seed: 1\n"
         "  --sampler <value>        [b200] fold | alias | exact: fold\n"
         "  --gpus <value>           [b200] number of GPUs: 1\n";
}

static bool parse_int_arg(const char *s, int32_t *out) {  // scopt Read[Int] = _.toInt
  char *ep = nullptr;
  errno = 0;
  if (!*s) return false;
  long long v = strtoll(s, &ep, 10);
  if (errno || *ep || v > INT32_MAX || v < INT32_MIN) return false;
  for (const char *c = s; *c; ++c)
    if (!((*c >= '0' && *c <= '9') || ((*c == '-' || *c == '+') && c == s))) return false;
  *out = (int32_t)v;
  return true;
}
static bool parse_double_arg(const char *s, double *out) {  // scopt Read[Double] = _.toDouble
  char *ep = nullptr;
  if (!*s) return false;
  double v = strtod(s, &ep);
  if (*ep && !((*ep == 'd' || *ep == 'D' || *ep == 'f' || *ep == 'F') && !ep[1])) return false;
  *out = v;
  return true;
}
static bool parse_bool_arg(const char *s, int32_t *out) {  // scopt Read[Boolean]
  std::string v(s);
  std::transform(v.begin(), v.end(), v.begin(), ::tolower);
  if (v == "true" || v == "yes" || v == "1") { *out = 1; return true; }
  if (v == "false" || v == "no" || v == "0") { *out = 0; return true; }
  return false;
}

extern "C" srw_status srw_params_parse_argv(int argc, const char *const *argv, srw_params *o) {
  if (!o || (argc > 0 && !argv)) return SRW_ERR_ARG;
  srw_params_default(o);
  bool has_in = false, has_out = false, has_cmd = false;
  for (int i = 0; i < argc; ++i) {
    std::string a = argv[i];
    if (a.rfind("--", 0) != 0) { srw_set_error("Error: Unknown argument '%s'", a.c_str()); return SRW_ERR_USAGE; }
    std::string name = a.substr(2), val;
    size_t eq = name.find('=');   // scopt also accepts --name=value
    if (eq != std::string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); }
    {
      // scopt looks the option up before it reads a value: `--bogus` alone is "Unknown option --bogus" (CP:32-109, scopt 3.7)
      static const char *const known[] = {"walkLength", "numWalks", "p", "q", "rddPartitions", "weighted", "directed", "singleOutput", "w2vPartitions",
                                          "input", "output", "cmd", "partitioned", "lr", "iter", "dim", "window", "seed", "sampler", "gpus"};
      bool found = false;
      for (const char *k : known) found = found || name == k;
      if (!found) { srw_set_error("Error: Unknown option --%s", name.c_str()); return SRW_ERR_USAGE; }
    }
    if (eq == std::string::npos) {
      if (i + 1 >= argc) { srw_set_error("Error: Missing value after '%s'", a.c_str()); return SRW_ERR_USAGE; }
      val = argv[++i];
    }
    const char *v = val.c_str();
    bool ok = true;
    if (name == "walkLength") ok = parse_int_arg(v, &o->walk_length);                  // CP:34-36
    else if (name == "numWalks") ok = parse_int_arg(v, &o->num_walks);                 // CP:37-39
    else if (name == "p") ok = parse_double_arg(v, &o->p);                             // CP:40-42
    else if (name == "q") ok = parse_double_arg(v, &o->q);                             // CP:43-45
    else if (name == "rddPartitions") ok = parse_int_arg(v, &o->rdd_partitions);       // CP:46-51
    else if (name == "weighted") ok = parse_bool_arg(v, &o->weighted);                 // CP:52-54
    else if (name == "directed") ok = parse_bool_arg(v, &o->directed);                 // CP:55-57
    else if (name == "singleOutput") ok = parse_bool_arg(v, &o->single_output);        // CP:58-60
    else if (name == "w2vPartitions") ok = parse_int_arg(v, &o->w2v_partitions);       // CP:61-63
    else if (name == "input") { ok = val.size() < SRW_PATH_MAX; if (ok) { strcpy(o->input, v); has_in = true; } }    // CP:64-67
    else if (name == "output") { ok = val.size() < SRW_PATH_MAX; if (ok) { strcpy(o->output, v); has_out = true; } } // CP:68-71
    else if (name == "cmd") {                                                          // CP:72-75 TaskName.withName
      has_cmd = true;
      if (val == "node2vec") o->cmd = SRW_TASK_NODE2VEC;
      else if (val == "randomwalk") o->cmd = SRW_TASK_RANDOMWALK;
      else if (val == "embedding") o->cmd = SRW_TASK_EMBEDDING;
      else ok = false;
    }
    else if (name == "partitioned") ok = parse_bool_arg(v, &o->partitioned);           // CP:76-78
    else if (name == "lr") ok = parse_double_arg(v, &o->w2v_lr);                       // CP:79-81
    else if (name == "iter") ok = parse_int_arg(v, &o->w2v_iter);                      // CP:82-84
    else if (name == "dim") ok = parse_int_arg(v, &o->w2v_dim);                        // CP:85-87
    else if (name == "window") ok = parse_int_arg(v, &o->w2v_window);                  // CP:88-90
    else if (name == "seed") { char *ep; errno = 0; o->seed = strtoull(v, &ep, 10); ok = !errno && *v && !*ep; }
    else if (name == "sampler") {
      if (val == "alias") o->sampler = SRW_SAMPLER_ALIAS;
      else if (val == "fold") o->sampler = SRW_SAMPLER_ALIAS_FOLD;
      else if (val == "exact") o->sampler = SRW_SAMPLER_EXACT;
      else ok = false;
    }
    else if (name == "gpus") ok = parse_int_arg(v, &o->num_gpus) && o->num_gpus >= 1;
    else { srw_set_error("Error: Unknown option --%s", name.c_str()); return SRW_ERR_USAGE; }
    if (!ok) { srw_set_error("Error: Option --%s failed when given '%s'", name.c_str(), v); return SRW_ERR_USAGE; }
  }
  if (!has_in) { srw_set_error("Error: Missing option --input"); return SRW_ERR_USAGE; }
  if (!has_out) { srw_set_error("Error: Missing option --output"); return SRW_ERR_USAGE; }
  if (!has_cmd) { srw_set_error("Error: Missing option --cmd"); return SRW_ERR_USAGE; }
  return SRW_OK;
}

// ------------------------------------------------------------------------------------------
// A1: edge-list text
// ------------------------------------------------------------------------------------------
static inline bool java_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\x0B' || c == '\f' || c == '\r'; }

// java.lang.Integer.parseInt
static bool java_int(const char *s, size_t n, int32_t *out) {
  if (n == 0) return false;
  size_t i = 0;
  bool neg = false;
  if (s[0] == '-' || s[0] == '+') { neg = s[0] == '-'; i = 1; }
  if (i == n) return false;
  int64_t v = 0;
  for (; i < n; ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    v = v * 10 + (s[i] - '0');
    if (v > 2147483648LL) return false;
  }
  if (neg) v = -v;
  if (v > INT32_MAX || v < INT32_MIN) return false;
  *out = (int32_t)v;
  return true;
}
// java.lang.Float.parseFloat: [sign] (NaN | Infinity | decimal | hex-with-p-exponent) [fFdD]
static bool java_float(const char *s, size_t n, float *out) {
  if (n == 0 || n > 100) return false;
  std::string t(s, n);
  size_t b = (t[0] == '+' || t[0] == '-') ? 1 : 0;
  if (t.compare(b, std::string::npos, "NaN") == 0) { *out = NAN; return true; }
  if (t.compare(b, std::string::npos, "Infinity") == 0) { *out = t[0] == '-' ? -INFINITY : INFINITY; return true; }
  const bool hex = t.size() >= b + 2 && t[b] == '0' && (t[b + 1] == 'x' || t[b + 1] == 'X');
  const bool has_p = t.find_first_of("pP") != std::string::npos;
  if (hex && !has_p) return false;
  char last = t.back();
  if ((last == 'f' || last == 'F' || last == 'd' || last == 'D') && (!hex || has_p)) t.pop_back();
  if (t.size() <= b) return false;
  // reject what strtof accepts but Java does not (inf/nan spellings, leading/trailing junk)
  for (size_t i = b; i < t.size(); ++i) {
    char c = t[i];
    bool okc = (c >= '0' && c <= '9') || c == '.' || c == '+' || c == '-' || c == 'e' || c == 'E';
    if (hex) okc = okc || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F') || c == 'x' || c == 'X' || c == 'p' || c == 'P';
    if (!okc) return false;
  }
  bool digit = false;
  for (size_t i = b; i < t.size(); ++i) digit = digit || (t[i] >= '0' && t[i] <= '9');
  if (!digit) return false;
  char *ep = nullptr;
  float v = strtof(t.c_str(), &ep);
  if (*ep) return false;
  *out = v;
  return true;
}

extern "C" srw_status srw_edges_parse_buffer(const char *buf, size_t len, int weighted, int partitioned, srw_edges **out) {
  if (!out || (len && !buf)) return SRW_ERR_ARG;
  srw_edges *E = new srw_edges();
  E->has_pid = partitioned != 0;
  size_t pos = 0;
  int64_t line_no = 0;
  std::vector<std::pair<const char *, size_t>> tok;
  while (pos < len) {
    size_t e = pos;
    while (e < len && buf[e] != '\n' && buf[e] != '\r') e++;
    size_t next = e;
    if (next < len) next += (buf[e] == '\r' && e + 1 < len && buf[e + 1] == '\n') ? 2 : 1;
    line_no++;
    // triplet.split("\\s+") (URW:26): leading whitespace -> empty first token; empty line -> [""]
    tok.clear();
    size_t i = pos;
    if (i == e || java_ws(buf[i])) tok.push_back({buf + i, 0});
    while (i < e) {
      while (i < e && java_ws(buf[i])) i++;
      if (i == e) break;
      size_t s0 = i;
      while (i < e && !java_ws(buf[i])) i++;
      tok.push_back({buf + s0, i - s0});
    }
    int32_t s = 0, d = 0, pid = 0;
    float w = 1.0f;
    if (!java_int(tok[0].first, tok[0].second, &s)) {
      srw_set_error("line %lld: NumberFormatException: For input string: \"%.*s\"", (long long)line_no, (int)tok[0].second, tok[0].first);
      delete E; return SRW_ERR_PARSE;
    }
    if (tok.size() < 2) { srw_set_error("line %lld: ArrayIndexOutOfBoundsException: 1", (long long)line_no); delete E; return SRW_ERR_PARSE; }
    if (!java_int(tok[1].first, tok[1].second, &d)) {
      srw_set_error("line %lld: NumberFormatException: For input string: \"%.*s\"", (long long)line_no, (int)tok[1].second, tok[1].first);
      delete E; return SRW_ERR_PARSE;
    }
    if (!partitioned) {
      if (weighted && tok.size() > 2 && !java_float(tok.back().first, tok.back().second, &w)) w = 1.0f;   // URW:29-32
    } else {
      if (tok.size() > 2 && !java_int(tok[2].first, tok[2].second, &pid)) pid = 0;                           // VRW:23-26 (random in the reference)
      if (weighted && tok.size() > 3 && !java_float(tok.back().first, tok.back().second, &w)) w = 1.0f;    // VRW:29-32
      E->pid.push_back(pid);
    }
    E->src.push_back(s); E->dst.push_back(d); E->w.push_back(w);
    pos = next;
  }
  *out = E;
  return SRW_OK;
}

extern "C" srw_status srw_edges_parse_file(const char *path, int weighted, int partitioned, srw_edges **out) {
  if (!path || !out) return SRW_ERR_ARG;
  FILE *f = fopen(path, "rb");
  if (!f) { srw_set_error("Input path does not exist: %s", path); return SRW_ERR_IO; }
  std::string data;
  char chunk[1 << 16];
  size_t r;
  while ((r = fread(chunk, 1, sizeof(chunk), f)) > 0) data.append(chunk, r);
  fclose(f);
  return srw_edges_parse_buffer(data.data(), data.size(), weighted, partitioned, out);
}

extern "C" srw_status srw_edges_view(const srw_edges *e, int64_t *n, const int32_t **h_src, const int32_t **h_dst,
                                     const float **h_w, const int32_t **h_pid) {
  if (!e) return SRW_ERR_ARG;
  if (n) *n = (int64_t)e->src.size();
  if (h_src) *h_src = e->src.data();
  if (h_dst) *h_dst = e->dst.data();
  if (h_w) *h_w = e->w.data();
  if (h_pid) *h_pid = e->has_pid ? e->pid.data() : nullptr;
  return SRW_OK;
}
extern "C" void srw_edges_free(srw_edges *e) { delete e; }

// ------------------------------------------------------------------------------------------
// GraphMap.addVertex builder (GM:23-56, 83-85): first insertion of a vid wins
// ------------------------------------------------------------------------------------------
extern "C" srw_status srw_graphmap_new(srw_graphmap **out) {
  if (!out) return SRW_ERR_ARG;
  *out = new srw_graphmap();
  return SRW_OK;
}
extern "C" srw_status srw_graphmap_add_vertex(srw_graphmap *m, int32_t vid, int64_t n, const int32_t *h_dst,
                                              const int32_t *h_pid, const float *h_w) {
  if (!m || n < 0 || (n > 0 && !h_dst)) return SRW_ERR_ARG;
  if (!m->seen.insert(vid).second) return SRW_OK;                                          // GM:42,54 case Some(value)
  m->vids.push_back(vid);
  m->row_off.push_back((int64_t)m->dst.size());
  m->row_len.push_back(n);
  for (int64_t i = 0; i < n; ++i) {
    m->dst.push_back(h_dst[i]);
    m->w.push_back(h_w ? h_w[i] : 1.0f);
    m->pid.push_back(h_pid ? h_pid[i] : 0);
  }
  if (h_pid && n > 0) m->any_pid = true;
  return SRW_OK;
}
extern "C" srw_status srw_graphmap_reset(srw_graphmap *m) {   // GM:99-107
  if (!m) return SRW_ERR_ARG;
  *m = srw_graphmap();
  return SRW_OK;
}
extern "C" srw_status srw_graphmap_counts(const srw_graphmap *m, int64_t *nv, int64_t *ne) {
  if (!m) return SRW_ERR_ARG;
  if (nv) *nv = (int64_t)m->vids.size();    // GM:87 srcVertexMap.size
  if (ne) *ne = (int64_t)m->dst.size();     // GM:91 offsetCounter
  return SRW_OK;
}
extern "C" srw_status srw_graphmap_finalize(const srw_graphmap *m, unsigned flags, srw_graph **out) {
  if (!m || !out) return SRW_ERR_ARG;
  return srw_build_graph_rows((int64_t)m->vids.size(), m->vids.data(), m->row_off.data(), m->row_len.data(), m->dst.data(),
                              m->any_pid ? m->pid.data() : nullptr, m->w.data(), flags, out);
}
extern "C" void srw_graphmap_free(srw_graphmap *m) { delete m; }

// ------------------------------------------------------------------------------------------
// A11: save (RW:234-241) -- path.mkString("\t") per line under <output>/path/part-NNNNN
// ------------------------------------------------------------------------------------------
static void append_int(std::string &s, int32_t v) {
  char b[16];
  int n = snprintf(b, sizeof(b), "%d", v);
  s.append(b, (size_t)n);
}
static void format_range(const srw_paths *p, int64_t first, int64_t last, std::string &s) {
  for (int64_t i = first; i < last; ++i) {
    for (int64_t k = p->offsets[i]; k < p->offsets[i + 1]; ++k) {
      if (k > p->offsets[i]) s.push_back('\t');
      append_int(s, p->ids[k]);
    }
    s.push_back('\n');
  }
}
extern "C" srw_status srw_paths_format(srw_paths *p, char *h_buf, int64_t cap, int64_t *needed) {
  if (!p) return SRW_ERR_ARG;
  std::string s;
  format_range(p, 0, p->n_paths, s);
  if (needed) *needed = (int64_t)s.size();
  if (h_buf && cap >= (int64_t)s.size()) memcpy(h_buf, s.data(), s.size());
  return SRW_OK;
}
static int mkdir_p(const std::string &dir) {
  std::string cur;
  for (size_t i = 0; i <= dir.size(); ++i) {
    if (i == dir.size() || dir[i] == '/') {
      if (!cur.empty() && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) return -1;
    }
    if (i < dir.size()) cur.push_back(dir[i]);
  }
  return 0;
}
extern "C" srw_status srw_save(srw_paths *p, const srw_params *params) {
  if (!p || !params || !params->output[0]) { srw_set_error("srw_save: no output path"); return SRW_ERR_ARG; }
  const std::string dir = std::string(params->output) + "/path";          // Property.pathSuffix
  struct stat st;
  if (stat(dir.c_str(), &st) == 0) {   // Hadoop saveAsTextFile refuses an existing directory
    srw_set_error("FileAlreadyExistsException: Output directory %s already exists", dir.c_str());
    return SRW_ERR_IO;
  }
  if (mkdir_p(dir) != 0) { srw_set_error("cannot create %s: %s", dir.c_str(), strerror(errno)); return SRW_ERR_IO; }
  int parts = params->single_output ? 1 : params->rdd_partitions;          // Main:64-69
  if (parts < 1) parts = 1;
  for (int k = 0; k < parts; ++k) {
    // repartition(n) spreads lines arbitrarily (RW:240); contiguous blocks here
    const int64_t first = p->n_paths * k / parts, last = p->n_paths * (k + 1) / parts;
    std::string s;
    format_range(p, first, last, s);
    char name[32];
    snprintf(name, sizeof(name), "/part-%05d", k);
    FILE *f = fopen((dir + name).c_str(), "wb");
    if (!f) { srw_set_error("cannot write %s%s: %s", dir.c_str(), name, strerror(errno)); return SRW_ERR_IO; }
    const bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
    if (fclose(f) != 0 || !ok) { srw_set_error("short write to %s%s", dir.c_str(), name); return SRW_ERR_IO; }
  }
  FILE *f = fopen((dir + "/_SUCCESS").c_str(), "wb");
  if (f) fclose(f);
  return SRW_OK;
}

extern "C" srw_status srw_paths_view(srw_paths *p, int64_t *n_paths, const int32_t **h_ids, const int64_t **h_offsets) {
  if (!p) return SRW_ERR_ARG;
  if (n_paths) *n_paths = p->n_paths;
  if (h_ids) *h_ids = p->ids.data();
  if (h_offsets) *h_offsets = p->offsets.data();
  return SRW_OK;
}
extern "C" srw_status srw_paths_counts(const srw_paths *p, int64_t *n_paths, int64_t *n_steps) {
  if (!p) return SRW_ERR_ARG;
  if (n_paths) *n_paths = p->n_paths;
  if (n_steps) *n_steps = p->n_steps;
  return SRW_OK;
}
extern "C" void srw_paths_free(srw_paths *p) { delete p; }
