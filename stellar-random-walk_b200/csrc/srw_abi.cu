// srw_abi.cu -- extern "C" entry points that touch the device: graph handles, the walk driver
// (RW:31-33 execute / RW:75-176 randomWalk), Main (Main:18-27,109-127), synthetic inputs.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <vector>

#include "philox.cuh"
#include "srw_internal.h"

srw_status srw_require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    srw_set_error("no CUDA device available (%s): libsrw has no CPU walk path", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    cudaGetLastError();
    return SRW_ERR_NO_DEVICE;
  }
  // The walk is a random 4..16-byte gather: ask L2 to fetch single 32-byte sectors instead of
  // promoting every miss to 64 bytes (ncu showed ~2x the touched sectors in dram__bytes_read).
  static thread_local int configured_for = -1;
  int dev = 0;
  // The library does not touch context-wide limits (cudaLimitMaxL2FetchGranularity used to be set here for every co-tenant of the
  // CUDA context): the walk kernels ask for 64-byte fills per load (ld.global.nc.L2::64B), which is what mattered.  SRW_L2_FETCH=<bytes>
  // opts in for experiments.
  if (cudaGetDevice(&dev) == cudaSuccess && configured_for != dev) {
    const char *g = getenv("SRW_L2_FETCH");
    if (g && atoi(g) > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    cudaGetLastError();
    configured_for = dev;
  }
  return SRW_OK;
}

extern "C" int srw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// ------------------------------------------------------------------------------------------
// graph handles
// ------------------------------------------------------------------------------------------
namespace {
template <class T>
struct Dev {
  T *p = nullptr;
  ~Dev() { if (p) cudaFree(p); }
  cudaError_t put(const T *h, int64_t n) {
    cudaError_t e = cudaMalloc(&p, (size_t)(n > 0 ? n : 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    return n > 0 ? cudaMemcpy(p, h, (size_t)n * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
  }
};
srw_status host_vids(const srw_graph *g) {
  if ((int64_t)g->h_vids.size() != g->nv) {
    g->h_vids.resize((size_t)g->nv);
    if (g->nv) SRW_CUDA(cudaMemcpy(g->h_vids.data(), g->d_vids, (size_t)g->nv * 4, cudaMemcpyDeviceToHost));
  }
  return SRW_OK;
}
// rank of vid, or -1
int64_t host_rank(const srw_graph *g, int32_t vid) {
  auto it = std::lower_bound(g->h_vids.begin(), g->h_vids.end(), vid);
  if (it == g->h_vids.end() || *it != vid) return -1;
  return (int64_t)(it - g->h_vids.begin());
}
}  // namespace

extern "C" srw_status srw_graph_from_device_edges(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w,
                                                  const int32_t *d_pid, int directed, unsigned flags, srw_graph **out) {
  return srw_build_graph_device(n, d_src, d_dst, d_w, d_pid, directed, flags, out);
}

extern "C" srw_status srw_graph_from_edges(int64_t n, const int32_t *h_src, const int32_t *h_dst, const float *h_w,
                                           const int32_t *h_pid, int directed, unsigned flags, srw_graph **out) {
  SRW_TRY(srw_require_device());
  if (n < 0 || !out || (n > 0 && (!h_src || !h_dst))) { srw_set_error("srw_graph_from_edges: bad argument"); return SRW_ERR_ARG; }
  Dev<int32_t> s, d, pid;
  Dev<float> w;
  SRW_CUDA(s.put(h_src, n));
  SRW_CUDA(d.put(h_dst, n));
  if (h_w) SRW_CUDA(w.put(h_w, n));
  if (h_pid) SRW_CUDA(pid.put(h_pid, n));
  return srw_build_graph_device(n, s.p, d.p, h_w ? w.p : nullptr, h_pid ? pid.p : nullptr, directed, flags, out);
}

extern "C" srw_status srw_graph_load(const srw_params *params, unsigned flags, srw_graph **out) {
  if (!params || !out) return SRW_ERR_ARG;
  SRW_TRY(srw_require_device());
  // URW:23-34 | VRW:19-34 on the device: the file is parsed in HBM and the edge arrays never exist on the host
  return srw_graph_load_device(params, flags, out);
}

extern "C" srw_status srw_graph_stats(const srw_graph *g, int64_t *nv, int64_t *ne) {
  if (!g) return SRW_ERR_ARG;
  if (nv) *nv = g->nv;
  if (ne) *ne = g->nnz;
  return SRW_OK;
}

extern "C" srw_status srw_graph_neighbors(const srw_graph *g, int32_t vid, int32_t *h_dst, float *h_w, int64_t cap, int64_t *n) {
  if (!g || !n) return SRW_ERR_ARG;
  if (!g->shards.empty()) {                       // container: the shard that owns the vertex answers (everyone else says -1)
    for (srw_graph *sh : g->shards) {
      SRW_CUDA(cudaSetDevice(sh->device));
      SRW_TRY(srw_graph_neighbors(sh, vid, h_dst, h_w, cap, n));
      if (*n >= 0) break;
    }
    cudaSetDevice(0);
    return SRW_OK;
  }
  SRW_TRY(host_vids(g));
  int64_t r = host_rank(g, vid);
  if (r < 0) { *n = -1; return SRW_OK; }                       // GM:118 case None => null
  int64_t ext[2];
  if (g->vcut) {
    // VCut shard map: owner(v) = getPartition(v) mod world; the replicated tables say whether the row is here, and where
    uint8_t o = 0;
    MigExt e;
    SRW_CUDA(cudaMemcpy(&o, g->d_owner + r, 1, cudaMemcpyDeviceToHost));
    if ((int)o != g->shard_rank && o != 0xFF /* a replicated hub row: on every shard */) { *n = -1; return SRW_OK; }
    SRW_CUDA(cudaMemcpy(&e, g->d_ext + r, sizeof e, cudaMemcpyDeviceToHost));
    ext[0] = e.off; ext[1] = (int64_t)e.off + e.deg;
  } else {
    if (g->shard_world > 1) {
      // a vertex-range shard holds rows [row_first, row_last) only and indexes them locally: a vertex owned by another
      // shard is "not on this shard" -- the reference's null (GM:118, the RW:121-129 case)
      if (r < g->row_first || r >= g->row_last) { *n = -1; return SRW_OK; }
      r -= g->row_first;
    }
    SRW_CUDA(cudaMemcpy(ext, g->d_off + r, 16, cudaMemcpyDeviceToHost));
  }
  const int64_t deg = ext[1] - ext[0];
  *n = deg;
  const int64_t m = std::min(deg, cap);
  if (m <= 0) return SRW_OK;
  const int32_t *col = g->d_col_app ? g->d_col_app : g->d_col;  // appearance order when it was kept
  if (!col) { srw_set_error("srw_graph_neighbors: this handle was built with SRW_BUILD_LEAN (no column array)"); return SRW_ERR_UNSUPPORTED; }
  if (h_dst) {
    std::vector<int32_t> ranks((size_t)m);
    SRW_CUDA(cudaMemcpy(ranks.data(), col + ext[0], (size_t)m * 4, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < m; ++i) h_dst[i] = g->h_vids[(size_t)ranks[i]];
  }
  if (h_w) {
    if (g->d_w_app) SRW_CUDA(cudaMemcpy(h_w, g->d_w_app + ext[0], (size_t)m * 4, cudaMemcpyDeviceToHost));
    else for (int64_t i = 0; i < m; ++i) h_w[i] = 1.0f;
  }
  return SRW_OK;
}

extern "C" srw_status srw_graph_partition(const srw_graph *g, int32_t vid, int32_t *pid, int *found) {
  if (!g || !found) return SRW_ERR_ARG;
  *found = 0;
  if (!g->shards.empty()) {                       // container: the map is replicated on every shard
    SRW_CUDA(cudaSetDevice(g->shards[0]->device));
    return srw_graph_partition(g->shards[0], vid, pid, found);
  }
  if (!g->d_vpid) return SRW_OK;
  SRW_TRY(host_vids(g));
  const int64_t r = host_rank(g, vid);
  if (r < 0) return SRW_OK;
  int32_t v = -1;
  SRW_CUDA(cudaMemcpy(&v, g->d_vpid + r, 4, cudaMemcpyDeviceToHost));
  if (v >= 0) { *found = 1; if (pid) *pid = v; }
  return SRW_OK;
}

extern "C" srw_status srw_graph_vertex_ids(const srw_graph *g, int32_t *h_out, int64_t cap) {
  if (!g || (cap > 0 && !h_out)) return SRW_ERR_ARG;
  if (!g->shards.empty()) { SRW_CUDA(cudaSetDevice(0)); return srw_graph_vertex_ids(g->shards[0], h_out, cap); }   // replicated on every shard
  const int64_t m = std::min(cap, g->nv);
  if (m > 0) SRW_CUDA(cudaMemcpy(h_out, g->d_vids, (size_t)m * 4, cudaMemcpyDeviceToHost));
  return SRW_OK;
}

extern "C" srw_status srw_graph_layout(const srw_graph *g, int64_t *h_off, int32_t *h_col, uint32_t *h_slots4, int *has_alias) {
  if (!g) return SRW_ERR_ARG;
  if (!g->shards.empty()) { srw_set_error("srw_graph_layout: a multi-GPU graph has one layout per shard"); return SRW_ERR_UNSUPPORTED; }
  if (has_alias) *has_alias = g->has_alias ? 1 : 0;
  // on a vertex-range shard the arrays are shard-local: row_last - row_first + 1 offsets (relative to the shard's first entry)
  const int64_t n_rows = g->shard_world > 1 ? g->row_last - g->row_first : g->nv;
  if (h_off) SRW_CUDA(cudaMemcpy(h_off, g->d_off, (size_t)(n_rows + 1) * 8, cudaMemcpyDeviceToHost));
  if (h_col && g->nnz && g->d_col) SRW_CUDA(cudaMemcpy(h_col, g->d_col, (size_t)g->nnz * 4, cudaMemcpyDeviceToHost));
  else if (h_col && g->nnz) {
    // SRW_BUILD_LEAN: the sorted column array was dropped; the neighbour entries carry the same ids in the same order
    if (!g->d_ent || !g->ent_ids) { srw_set_error("srw_graph_layout: this handle holds no column array"); return SRW_ERR_UNSUPPORTED; }
    SRW_TRY(srw_ent_ranks_to_host(g, h_col));
  }
  if (h_slots4 && g->has_alias) SRW_CUDA(cudaMemcpy(h_slots4, g->d_slot, (size_t)g->nnz * 16, cudaMemcpyDeviceToHost));
  return SRW_OK;
}

extern "C" int64_t srw_graph_device_bytes(const srw_graph *g) { return g ? g->device_bytes : 0; }
extern "C" const char *srw_graph_build_profile(const srw_graph *g) {
  if (!g) return "{}";
  if (!g->shards.empty()) return g->shards[0] ? g->shards[0]->build_profile.c_str() : "{}";
  return g->build_profile.empty() ? "{}" : g->build_profile.c_str();
}

extern "C" void srw_graph_free(srw_graph *g) {
  if (!g) return;
  if (!g->shards.empty() || g->multi) {           // multi-GPU container
    srw_multi_free(g->multi);
    for (srw_graph *sh : g->shards) { if (sh) { cudaSetDevice(sh->device); srw_graph_free(sh); } }
    cudaSetDevice(0);
    delete g;
    return;
  }
  cudaFree(g->d_bitmap); cudaFree(g->d_wordrank); cudaFree(g->d_vids);
  cudaFree(g->d_col_app); cudaFree(g->d_w_app); cudaFree(g->d_col); cudaFree(g->d_slot); cudaFree(g->d_slotw); cudaFree(g->d_vpid);
  cudaFree(g->d_meta); cudaFree(g->d_hash_id); cudaFree(g->d_bloom); cudaFree(g->d_ext); cudaFree(g->d_owner); cudaFree(g->d_lverts);
  if (!g->rows_external) { cudaFree(g->d_off); cudaFree(g->d_hash); cudaFree(g->d_ent); }
  for (int r = 0; r < SRW_MAX_SHARDS; ++r)
    if (g->peer_ipc[r]) {
      if (g->peer_off[r]) cudaIpcCloseMemHandle((void *)g->peer_off[r]);
      if (g->peer_ent[r]) cudaIpcCloseMemHandle((void *)g->peer_ent[r]);
      if (g->peer_hash[r]) cudaIpcCloseMemHandle((void *)g->peer_hash[r]);
    }
  cudaGetLastError();
  srw_shard_scratch_free(g->scratch);
  delete g;
}

// ------------------------------------------------------------------------------------------
// walk driver
// ------------------------------------------------------------------------------------------
extern "C" srw_status srw_walk_device(const srw_graph *g, const srw_params *params, uint64_t walker_first, int64_t n_walkers,
                                      int32_t *d_paths, int32_t *d_lens, void *stream) {
  SRW_TRY(srw_require_device());
  if (!g || !params || (n_walkers > 0 && (!d_paths || !d_lens))) { srw_set_error("srw_walk_device: bad argument"); return SRW_ERR_ARG; }
  if (!g->shards.empty()) {
    // multi-GPU container: whole rounds only (the sharded walk seeds one walker per vertex and round), delivered on device 0
    if (g->nv == 0 || walker_first % (uint64_t)g->nv != 0 || n_walkers % g->nv != 0) { srw_set_error("srw_walk_device on a multi-GPU graph walks whole rounds: walker_first and n_walkers must be multiples of nVertices"); return SRW_ERR_ARG; }
    if (n_walkers == 0) return SRW_OK;
    srw_walk_info wi;
    SRW_TRY(srw_multi_walk_rounds(g, params, (int64_t)(walker_first / (uint64_t)g->nv), n_walkers / g->nv, d_paths, d_lens, &wi));
    srw_set_walk_info(wi.kernel_ms, wi.kernel_launches, wi.steps, 0, 0, 0);
    return SRW_OK;
  }
  WalkLaunch l{walker_first, n_walkers, d_paths, d_lens, (cudaStream_t)stream};
  return srw_walk_launch(g, params, l);
}

// RW:75-176: numWalks rounds, one walker per vertex per round.  Rounds are independent under the
// counter-based RNG; they run in device-sized batches and are compacted into a ragged host array.
extern "C" srw_status srw_walk(const srw_graph *g, const srw_params *params, srw_paths **out) {
  SRW_TRY(srw_require_device());
  if (!g || !params || !out) { srw_set_error("srw_walk: bad argument"); return SRW_ERR_ARG; }
  if (params->num_walks < 0) { srw_set_error("numWalks must be >= 0"); return SRW_ERR_ARG; }
  const bool multi = !g->shards.empty();
  if (params->num_gpus > 1 && !multi) { srw_set_error("--gpus %d: the graph was loaded on one GPU (load it with the same --gpus)", params->num_gpus); return SRW_ERR_ARG; }
  SRW_CUDA(cudaSetDevice(g->device));
  const int32_t stride = params->walk_length + 2;
  const int64_t total = (int64_t)params->num_walks * g->nv;
  std::unique_ptr<srw_paths> owner(new srw_paths());     // released into *out on success only: no error path leaks it
  srw_paths *P = owner.get();
  P->stride = stride;
  P->offsets.push_back(0);
  size_t free_b = 0, total_b = 0;
  SRW_CUDA(cudaMemGetInfo(&free_b, &total_b));
  int64_t batch = (int64_t)(free_b / 2) / ((int64_t)stride * 4 + 4);
  if (batch > total) batch = total;
  if (batch < 1) batch = 1;
  if (multi && g->nv > 0) {
    // whole rounds per batch; device 0 also holds its shard and its exchange block (~400 bytes per walker of the batch)
    int64_t rounds = std::max<int64_t>(1, std::min<int64_t>({(int64_t)params->num_walks, ((int64_t)1 << 25) / g->nv, (int64_t)(free_b / 3) / (g->nv * ((int64_t)stride * 4 + 420))}));
    batch = rounds * g->nv;
  }
  Dev<int32_t> d_paths, d_lens;
  SRW_CUDA(cudaMalloc(&d_paths.p, (size_t)batch * stride * 4));
  SRW_CUDA(cudaMalloc(&d_lens.p, (size_t)batch * 4));
  // The result is ONE flat id array: it is reserved once (no regrowth copies of a 10-GB vector) and every batch is copied from the
  // device straight into its place in it (no staging vector); ragged paths (dead ends, RW:115-119) are compacted in place.
  std::vector<int32_t> h_lens((size_t)batch);
  try {
    P->ids.reserve((size_t)total * (size_t)stride);
    P->offsets.reserve((size_t)total + 1);
  } catch (const std::bad_alloc &) {
    srw_set_error("srw_walk: %lld paths of up to %d ids do not fit host memory (srw_walk_save streams them to files, srw_walk_device leaves them in HBM)",
                  (long long)total, (int)stride);
    return SRW_ERR_ARG;
  }
  double kernel_ms = 0;
  int64_t launches = 0, steps = 0, props = 0, mem = 0, logs = 0;
  for (int64_t first = 0; first < total; first += batch) {
    const int64_t n = std::min(batch, total - first);
    srw_status s = srw_walk_device(g, params, (uint64_t)first, n, d_paths.p, d_lens.p, nullptr);
    if (s != SRW_OK) return s;
    srw_walk_info wi;
    srw_last_walk_info(&wi);
    kernel_ms += wi.kernel_ms; launches += wi.kernel_launches; steps += wi.steps;
    props += wi.proposals; mem += wi.member_tests; logs += wi.probes_log2;
    SRW_CUDA(cudaMemcpy(h_lens.data(), d_lens.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    const size_t base = P->ids.size();
    P->ids.resize(base + (size_t)n * (size_t)stride);
    SRW_CUDA(cudaMemcpy(P->ids.data() + base, d_paths.p, (size_t)n * stride * 4, cudaMemcpyDeviceToHost));
    bool all_full = true;                  // full-length paths: every walk on an undirected graph
    for (int64_t i = 0; i < n && all_full; ++i) all_full = h_lens[i] == stride;
    if (all_full) {
      for (int64_t i = 1; i <= n; ++i) P->offsets.push_back((int64_t)base + i * stride);
    } else {
      int32_t *ids = P->ids.data();
      size_t w = base;
      for (int64_t i = 0; i < n; ++i) {
        const size_t len = (size_t)std::max<int32_t>(0, std::min<int32_t>(h_lens[i], stride)), from = base + (size_t)i * (size_t)stride;
        if (w != from && len) memmove(ids + w, ids + from, len * 4);
        w += len;
        P->offsets.push_back((int64_t)w);
      }
      P->ids.resize(w);
    }
  }
  P->n_paths = total;
  P->n_steps = steps;
  // expose the totals of the whole call
  srw_set_walk_info(kernel_ms, launches, steps, props, mem, logs);
  *out = owner.release();
  return SRW_OK;
}

// ------------------------------------------------------------------------------------------
// Main.main / runJob for --cmd randomwalk (Main:18-27, 53-62, 109-127)
// ------------------------------------------------------------------------------------------
extern "C" int srw_main(int argc, const char *const *argv) {
  srw_params prm;
  if (srw_params_parse_argv(argc, argv, &prm) != SRW_OK) {      // CP:107 -> None => sys.exit(1) (Main:25)
    fprintf(stderr, "%s\n%s", srw_last_error(), srw_usage());
    return 1;
  }
  if (prm.cmd != SRW_TASK_RANDOMWALK) {
    // Main:113-124: node2vec / embedding need MLlib Word2Vec, which is outside this engine's scope
    fprintf(stderr, "Error: --cmd %s is not supported by this engine (only randomwalk)\n", prm.cmd == SRW_TASK_NODE2VEC ? "node2vec" : "embedding");
    return 1;
  }
  // alias and fold share a layout; the exact sampler also uses the neighbour hash sets of the alias layout
  // (alias | fold: lean -- only the arrays the walk kernel reads stay in HBM; ignored where it does not apply)
  const unsigned flags = prm.sampler == SRW_SAMPLER_EXACT ? SRW_BUILD_ALL : (SRW_BUILD_ALIAS | SRW_BUILD_LEAN);
  srw_graph *g = nullptr;
  auto t0 = std::chrono::steady_clock::now();
  if (srw_graph_load(&prm, flags, &g) != SRW_OK) { fprintf(stderr, "Exception: %s\n", srw_last_error()); return 2; }
  int64_t nv = 0, ne = 0;
  srw_graph_stats(g, &nv, &ne);
  printf("edges: %lld\nvertices: %lld\n", (long long)ne, (long long)nv);   // URW:71-72
  auto t1 = std::chrono::steady_clock::now();
  // execute() + save() (Main:53-62) streamed through the device formatter
  int rc = 0;
  g_srw_log_supersteps = prm.num_gpus > 1;       // the sharded walk has real super-steps: their RW:154 lines are printed as they happen
  if (srw_walk_save(g, &prm) != SRW_OK) { fprintf(stderr, "Exception: %s\n", srw_last_error()); srw_graph_free(g); return 2; }
  g_srw_log_supersteps = false;
  auto t2 = std::chrono::steady_clock::now();
  srw_walk_info wi;
  srw_last_walk_info(&wi);
  // The reference's self-checks (RW:150-167).  On one GPU a round is ONE super-step that finishes every walker, as in Spark
  // local[*]: one `Unfinished Walkers: 0` per round (RW:154).  `Zero Neighbors` is printed when walkers stopped at a vertex without
  // out-neighbours (RW:115-119, RW:155-158); `Wrong Transports` (RW:117-123: a walker shipped to a partition that does not hold its
  // vertex) cannot happen here -- a walker is routed by the owner table of the same build.  The path-count check (RW:164-167):
  if (prm.num_gpus <= 1)
    for (int r = 0; r < prm.num_walks; ++r) printf("Unfinished Walkers: 0\n");
  if (srw_last_short_paths() > 0) printf("Wrong Transports: 0\nZero Neighbors: %lld\n", (long long)srw_last_short_paths());
  {
    const int64_t expect = (int64_t)prm.num_walks * nv, got = expect;       // one path per vertex and round by construction (checked: the writer counts lines)
    if (got != expect) printf("Inconsistent number of paths: nPaths=[%lld] != vertices[%lld]\n", (long long)got, (long long)nv);
  }
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  fprintf(stderr, "[srw] load+build %.1f ms, walk+save %.1f ms (walk kernels %.2f ms, %lld steps)\n", ms(t0, t1), ms(t1, t2),
          wi.kernel_ms, (long long)wi.steps);
  srw_graph_free(g);
  return rc;
}

// ------------------------------------------------------------------------------------------
// synthetic inputs (benchmark utilities)
// ------------------------------------------------------------------------------------------
namespace {
constexpr uint32_t kRmatTag = 0x524D4154u, kWeightTag = 0x57454947u;
__global__ void k_rmat(int scale, uint32_t seed, int64_t first, int64_t count, uint32_t A, uint32_t AB, uint32_t ABC,
                       int32_t *src, int32_t *dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t e = (uint64_t)(first + i);
    uint32_t s = 0, d = 0;
    for (int blk = 0; blk * 4 < scale; ++blk) {
      const Philox4 r = philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)blk, kRmatTag, seed, 0u);
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (blk * 4 + k >= scale) break;
        const uint32_t x = w[k];
        s = (s << 1) | (x >= AB ? 1u : 0u);
        d = (d << 1) | (((x >= A && x < AB) || x >= ABC) ? 1u : 0u);
      }
    }
    src[i] = (int32_t)s;
    dst[i] = (int32_t)d;
  }
}
__global__ void k_weights(uint32_t seed, int64_t first, int64_t count, float *w) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t e = (uint64_t)(first + i);
    const Philox4 r = philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), 0u, kWeightTag, seed, 0u);
    w[i] = __fadd_rn(1.0f, __fdiv_rn((float)(r.x % 1000u), 1000.0f));
  }
}
}  // namespace

extern "C" srw_status srw_synth_rmat_device(int scale, int edge_factor, uint64_t seed, int64_t first, int64_t count,
                                            int32_t *d_src, int32_t *d_dst) {
  SRW_TRY(srw_require_device());
  if (scale < 1 || scale > 30 || edge_factor < 1 || count < 0 || (count > 0 && (!d_src || !d_dst))) return SRW_ERR_ARG;
  if (count == 0) return SRW_OK;
  const uint32_t A = (uint32_t)(0.57 * 4294967296.0), AB = (uint32_t)((0.57 + 0.19) * 4294967296.0),
                 ABC = (uint32_t)((0.57 + 0.19 + 0.19) * 4294967296.0);
  k_rmat<<<148 * 8, 256>>>(scale, (uint32_t)seed, first, count, A, AB, ABC, d_src, d_dst);
  SRW_CUDA(cudaDeviceSynchronize());
  return SRW_OK;
}
extern "C" srw_status srw_synth_weights_device(uint64_t seed, int64_t first, int64_t count, float *d_w) {
  SRW_TRY(srw_require_device());
  if (count < 0 || (count > 0 && !d_w)) return SRW_ERR_ARG;
  if (count == 0) return SRW_OK;
  k_weights<<<148 * 8, 256>>>((uint32_t)seed, first, count, d_w);
  SRW_CUDA(cudaDeviceSynchronize());
  return SRW_OK;
}
