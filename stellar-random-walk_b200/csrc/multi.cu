// multi.cu -- the sharded walk behind the ordinary entry points: ONE process driving the GPUs of the box.
//
// `--gpus N` (srw_params::num_gpus > 1) makes srw_graph_load / srw_graph_from_edges_multi build one vertex-range shard per
// device (Main:54-57 picks the partitioned walker from Params in the same way) inside a container handle, and srw_walk /
// srw_walk_save walk it with the migrating-walker kernel of migrate.cuh: the reference's super-step loop (RW:91-162) with the
// shuffle (RW:186-192) done by the step kernel's own peer stores.  Inside one process peer memory is plain cudaMalloc memory
// with peer access enabled, and the barrier between super-steps is a set of CUDA events every device's stream waits for.
// One process per GPU (torchrun, NCCL barrier) drives the same srw_mig_* calls from sharded.py.
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "srw_internal.h"

struct srw_mig;
extern "C" {
srw_status srw_mig_block_bytes(const srw_graph *g, const srw_params *p, int64_t n_rounds, int64_t seg_cap, int64_t *bytes);
srw_status srw_mig_create(const srw_graph *g, const srw_params *p, int64_t n_rounds, int64_t seg_cap, void *d_block_self,
                          void *const *d_block_peers, srw_mig **out);
srw_status srw_mig_begin(srw_mig *m, int64_t round_first, int64_t n_rounds, void *stream);
srw_status srw_mig_superstep(srw_mig *m, int64_t s, unsigned long long *d_sent, void *stream);
srw_status srw_mig_counters(srw_mig *m, int64_t *h_out8, void *stream);
srw_status srw_mig_finish(srw_mig *m, int32_t **d_paths, int32_t *d_lens_out, int64_t *n_rows, int64_t *steps, void *stream);
void srw_mig_free(srw_mig *m);
}

struct MultiWalk {
  int world = 0;
  int64_t n_rounds = 0;                 // capacity of the contexts
  srw_params key;                       // the parameters the contexts were created for
  std::vector<srw_mig *> ctx;
  std::vector<void *> blocks;
  std::vector<cudaStream_t> streams;
  std::vector<cudaEvent_t> ev;
};

namespace {
struct GatherArgs {
  int world;
  const int32_t *paths[SRW_MAX_SHARDS];   // home path rows of every shard (peer pointers)
  int64_t home_rows[SRW_MAX_SHARDS];
};
// walker order on device 0: row (rnd, v) <- home shard v mod W, row rnd * home_rows + v / W
__global__ void multi_gather_kernel(GatherArgs a, int64_t nv, int64_t n_rounds, int32_t stride, int32_t *out, int32_t *lens) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < nv * n_rounds; i += n_warps) {
    const int64_t rnd = i / nv, v = i % nv;
    const int h = (int)(v % a.world);
    const int32_t *src = a.paths[h] + (rnd * a.home_rows[h] + v / a.world) * stride;
    int32_t *dst = out + i * stride;
    for (int32_t k = lane; k < stride; k += 32) dst[k] = src[k];
    if (lane == 0) lens[i] = stride;
  }
}

void multi_release(MultiWalk *w) {
  if (!w) return;
  for (size_t d = 0; d < w->ctx.size(); ++d) {
    cudaSetDevice((int)d);
    if (w->ctx[d]) srw_mig_free(w->ctx[d]);
    if (d < w->blocks.size() && w->blocks[d]) cudaFree(w->blocks[d]);
    if (d < w->streams.size() && w->streams[d]) cudaStreamDestroy(w->streams[d]);
    if (d < w->ev.size() && w->ev[d]) cudaEventDestroy(w->ev[d]);
  }
  cudaGetLastError();
  delete w;
}

bool same_walk(const srw_params &a, const srw_params &b) {
  return a.walk_length == b.walk_length && a.p == b.p && a.q == b.q && a.seed == b.seed && a.sampler == b.sampler;
}

srw_status multi_prepare(const srw_graph *g, const srw_params *p, int64_t n_rounds, MultiWalk **out) {
  srw_graph *mg = const_cast<srw_graph *>(g);
  MultiWalk *w = mg->multi;
  if (w && w->n_rounds >= n_rounds && same_walk(w->key, *p)) { *out = w; return SRW_OK; }
  multi_release(w);
  mg->multi = nullptr;
  const int W = (int)g->shards.size();
  w = new MultiWalk();
  w->world = W; w->n_rounds = n_rounds; w->key = *p;
  w->ctx.assign((size_t)W, nullptr); w->blocks.assign((size_t)W, nullptr); w->streams.assign((size_t)W, nullptr); w->ev.assign((size_t)W, nullptr);
  srw_status rc = SRW_OK;
  int64_t bytes = 0;
  rc = srw_mig_block_bytes(g->shards[0], p, n_rounds, 0, &bytes);
  for (int d = 0; d < W && rc == SRW_OK; ++d) {
    if (cudaSetDevice(d) != cudaSuccess || cudaMalloc(&w->blocks[(size_t)d], (size_t)bytes) != cudaSuccess ||
        cudaStreamCreateWithFlags(&w->streams[(size_t)d], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&w->ev[(size_t)d], cudaEventDisableTiming) != cudaSuccess) {
      srw_set_error("multi-GPU walk: cannot allocate the %.1f GB exchange block on device %d: %s", bytes / 1e9, d, cudaGetErrorString(cudaGetLastError()));
      rc = SRW_ERR_CUDA;
    }
  }
  for (int d = 0; d < W && rc == SRW_OK; ++d) rc = srw_mig_create(g->shards[(size_t)d], p, n_rounds, 0, w->blocks[(size_t)d], w->blocks.data(), &w->ctx[(size_t)d]);
  if (rc != SRW_OK) { multi_release(w); return rc; }
  mg->multi = w;
  *out = w;
  return SRW_OK;
}

// every device's stream waits for every device's work so far (the barrier between super-steps)
srw_status multi_barrier(MultiWalk *w) {
  for (int d = 0; d < w->world; ++d) { SRW_CUDA(cudaSetDevice(d)); SRW_CUDA(cudaEventRecord(w->ev[(size_t)d], w->streams[(size_t)d])); }
  for (int d = 0; d < w->world; ++d) {
    SRW_CUDA(cudaSetDevice(d));
    for (int e = 0; e < w->world; ++e)
      if (e != d) SRW_CUDA(cudaStreamWaitEvent(w->streams[(size_t)d], w->ev[(size_t)e], 0));
  }
  return SRW_OK;
}
}  // namespace

void srw_multi_free(MultiWalk *w) { multi_release(w); }
bool g_srw_log_supersteps = false;

// Rounds [round_first, round_first + n_rounds) of the walk over a multi-GPU container graph, delivered on device 0 in walker
// order: d_paths0 [n_rounds * nv][walk_length + 2] vertex ids, d_lens0 [n_rounds * nv].  Blocking.
srw_status srw_multi_walk_rounds(const srw_graph *g, const srw_params *p, int64_t round_first, int64_t n_rounds, int32_t *d_paths0,
                                 int32_t *d_lens0, srw_walk_info *info) {
  if (!g || g->shards.empty() || !p || n_rounds < 1) { srw_set_error("srw_multi_walk_rounds: bad argument"); return SRW_ERR_ARG; }
  MultiWalk *w = nullptr;
  SRW_TRY(multi_prepare(g, p, n_rounds, &w));
  const int W = w->world;
  cudaEvent_t t0, t1;
  SRW_CUDA(cudaSetDevice(0));
  SRW_CUDA(cudaEventCreate(&t0)); SRW_CUDA(cudaEventCreate(&t1));
  for (int d = 0; d < W; ++d) SRW_TRY(srw_mig_begin(w->ctx[(size_t)d], round_first, n_rounds, w->streams[(size_t)d]));
  SRW_TRY(multi_barrier(w));                 // peers store into a block from super-step 0 on: every block is initialised first
  SRW_CUDA(cudaSetDevice(0));
  SRW_CUDA(cudaEventRecord(t0, w->streams[0]));
  int64_t super_steps = 0;
  for (int64_t s = 0;; ++s) {
    for (int d = 0; d < W; ++d) SRW_TRY(srw_mig_superstep(w->ctx[(size_t)d], s, nullptr, w->streams[(size_t)d]));
    SRW_TRY(multi_barrier(w));
    super_steps = s + 1;
    // RW:162 `remainingWalkers != 0`, read back every few super-steps (an empty super-step is harmless); super-step 1 still seeds
    if (s >= 1 && ((s & 3) == 3 || W == 1)) {
      int64_t sent = 0, c8[8];
      for (int d = 0; d < W; ++d) { SRW_TRY(srw_mig_counters(w->ctx[(size_t)d], c8, w->streams[(size_t)d])); sent += c8[0]; }
      if (g_srw_log_supersteps) printf("Unfinished Walkers: %lld\n", (long long)sent);      // RW:154 (inbox slots in flight, read back every 4th super-step)
      if (sent == 0) break;
    }
    if (s > (int64_t)1 << 20) { srw_set_error("multi-GPU walk did not terminate"); return SRW_ERR_CUDA; }
  }
  SRW_CUDA(cudaSetDevice(0));
  SRW_CUDA(cudaEventRecord(t1, w->streams[0]));
  GatherArgs ga{};
  ga.world = W;
  int64_t steps = 0;
  for (int d = 0; d < W; ++d) {
    int32_t *pp = nullptr;
    int64_t rows = 0, st = 0;
    SRW_TRY(srw_mig_finish(w->ctx[(size_t)d], &pp, nullptr, &rows, &st, w->streams[(size_t)d]));     // synchronises the device's stream
    ga.paths[d] = pp;
    ga.home_rows[d] = (g->nv - d + W - 1) / W;
    steps += st;
  }
  SRW_CUDA(cudaSetDevice(0));
  const int64_t total = g->nv * n_rounds;
  if (total > 0) {
    int64_t b = (total * 32 + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
    multi_gather_kernel<<<(unsigned)b, 256, 0, w->streams[0]>>>(ga, g->nv, n_rounds, p->walk_length + 2, d_paths0, d_lens0);
  }
  SRW_CUDA(cudaStreamSynchronize(w->streams[0]));
  SRW_CUDA(cudaGetLastError());
  float ms = 0.f;
  cudaEventElapsedTime(&ms, t0, t1);
  cudaEventDestroy(t0); cudaEventDestroy(t1);
  if (info) { *info = srw_walk_info{}; info->kernel_ms = ms; info->kernel_launches = super_steps * W + W + 1; info->steps = steps; }
  return SRW_OK;
}

// Builds one vertex-range shard per device from an edge list resident on the CURRENT device.
// d_pid != NULL (`--partitioned true`): the partition-id column is the shard map (VCut: owner(v) = getPartition(v) mod num_gpus).
srw_status srw_build_graph_device_multi(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w, int directed,
                                        int num_gpus, srw_graph **out, const int32_t *d_pid, double hub_fraction) {
  SRW_TRY(srw_require_device());
  // replicated hub rows (srw.h, srw_graph_from_device_edges_vcut): the rows holding up to this share of the adjacency entries are
  // kept by every shard.  Default 0.5 (12 extra bytes per adjacency entry of the whole graph on every GPU); SRW_HUB_FRACTION overrides.
  if (hub_fraction < 0.0) {
    const char *e = getenv("SRW_HUB_FRACTION");
    hub_fraction = e ? atof(e) : 0.5;
    if (!(hub_fraction >= 0.0)) hub_fraction = 0.0;
    if (hub_fraction > 0.95) hub_fraction = 0.95;
  }
  int have = 0;
  cudaGetDeviceCount(&have);
  if (num_gpus < 2 || num_gpus > SRW_MAX_SHARDS || num_gpus > have) { srw_set_error("--gpus %d: this process sees %d CUDA device(s) (at most %d shards)", num_gpus, have, SRW_MAX_SHARDS); return SRW_ERR_ARG; }
  if (directed) { srw_set_error("--gpus > 1 walks undirected graphs (the second-order test runs at owner(x) as t in N(x))"); return SRW_ERR_UNSUPPORTED; }
  int dev0 = 0;
  SRW_CUDA(cudaGetDevice(&dev0));
  if (dev0 != 0) { srw_set_error("multi-GPU build: the edge list must live on device 0"); return SRW_ERR_ARG; }
  for (int a = 0; a < num_gpus; ++a)
    for (int b = 0; b < num_gpus; ++b) {
      if (a == b) continue;
      int can = 0;
      SRW_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
      if (!can) { srw_set_error("device %d cannot address device %d: --gpus needs peer access (NVLink)", a, b); return SRW_ERR_UNSUPPORTED; }
      SRW_CUDA(cudaSetDevice(a));
      const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SRW_CUDA(e);
      cudaGetLastError();
    }
  srw_graph *c = new srw_graph();
  c->device = 0; c->directed = false; c->shard_world = num_gpus; c->shard_rank = -1;
  srw_status rc = SRW_OK;
  for (int d = 0; d < num_gpus && rc == SRW_OK; ++d) {
    cudaSetDevice(d);
    int32_t *s = nullptr, *t = nullptr, *pv = nullptr;
    float *wv = nullptr;
    if (d == 0) { s = const_cast<int32_t *>(d_src); t = const_cast<int32_t *>(d_dst); wv = const_cast<float *>(d_w); pv = const_cast<int32_t *>(d_pid); }
    else if (n > 0) {
      if (cudaMalloc(&s, (size_t)n * 4) != cudaSuccess || cudaMalloc(&t, (size_t)n * 4) != cudaSuccess || (d_w && cudaMalloc(&wv, (size_t)n * 4) != cudaSuccess) ||
          (d_pid && cudaMalloc(&pv, (size_t)n * 4) != cudaSuccess) ||
          cudaMemcpyPeer(s, d, d_src, 0, (size_t)n * 4) != cudaSuccess || cudaMemcpyPeer(t, d, d_dst, 0, (size_t)n * 4) != cudaSuccess ||
          (d_w && cudaMemcpyPeer(wv, d, d_w, 0, (size_t)n * 4) != cudaSuccess) || (d_pid && cudaMemcpyPeer(pv, d, d_pid, 0, (size_t)n * 4) != cudaSuccess)) {
        srw_set_error("multi-GPU build: copying the edge list to device %d failed: %s", d, cudaGetErrorString(cudaGetLastError()));
        rc = SRW_ERR_CUDA;
      }
    }
    srw_graph *sh = nullptr;
    if (rc == SRW_OK) rc = srw_build_graph_device_sharded(n, s, t, wv, 0, SRW_BUILD_ALIAS | SRW_BUILD_MIGRATE, d, num_gpus, &sh, n > 0 ? pv : nullptr, hub_fraction);
    if (d != 0) { cudaFree(s); cudaFree(t); cudaFree(wv); cudaFree(pv); }
    if (rc == SRW_OK) {
      c->shards.push_back(sh);
      if (sh->nnz > 0 && !sh->d_ent) { srw_set_error("--gpus > 1 walks unweighted graphs (weighted rows carry Vose slots, which the sharded walk does not read yet)"); rc = SRW_ERR_UNSUPPORTED; }
    }
  }
  cudaSetDevice(0);
  if (rc != SRW_OK) { srw_graph_free(c); return rc; }
  c->nv = c->shards[0]->nv; c->nnz = c->shards[0]->nnz_global; c->nnz_global = c->nnz;
  c->bounds = c->shards[0]->bounds; c->row_first = 0; c->row_last = c->nv;
  c->id_min = c->shards[0]->id_min; c->vcut = c->shards[0]->vcut; c->has_pid = c->shards[0]->has_pid;
  for (auto *sh : c->shards) c->device_bytes += sh->device_bytes;
  *out = c;
  return SRW_OK;
}

static srw_status from_edges_multi(int64_t n, const int32_t *h_src, const int32_t *h_dst, const int32_t *h_pid, int directed, int num_gpus,
                                   srw_graph **out) {
  SRW_TRY(srw_require_device());
  if (n < 0 || !out || (n > 0 && (!h_src || !h_dst))) { srw_set_error("srw_graph_from_edges_multi: bad argument"); return SRW_ERR_ARG; }
  SRW_CUDA(cudaSetDevice(0));
  int32_t *s = nullptr, *d = nullptr, *pp = nullptr;
  SRW_CUDA(cudaMalloc(&s, (size_t)(n ? n : 1) * 4));
  if (cudaMalloc(&d, (size_t)(n ? n : 1) * 4) != cudaSuccess || (h_pid && cudaMalloc(&pp, (size_t)(n ? n : 1) * 4) != cudaSuccess)) {
    cudaFree(s); cudaFree(d); srw_set_error("out of device memory"); return SRW_ERR_CUDA;
  }
  srw_status rc = SRW_OK;
  if (cudaMemcpy(s, h_src, (size_t)n * 4, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(d, h_dst, (size_t)n * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      (h_pid && cudaMemcpy(pp, h_pid, (size_t)n * 4, cudaMemcpyHostToDevice) != cudaSuccess)) {
    srw_set_error("srw_graph_from_edges_multi: H2D copy failed");
    rc = SRW_ERR_CUDA;
  }
  if (rc == SRW_OK) rc = srw_build_graph_device_multi(n, s, d, nullptr, directed, num_gpus, out, pp, -1.0);
  cudaFree(s); cudaFree(d); cudaFree(pp);
  return rc;
}
extern "C" srw_status srw_graph_from_edges_multi(int64_t n, const int32_t *h_src, const int32_t *h_dst, int directed, int num_gpus,
                                                 srw_graph **out) {
  return from_edges_multi(n, h_src, h_dst, nullptr, directed, num_gpus, out);
}
extern "C" srw_status srw_graph_from_edges_multi_vcut(int64_t n, const int32_t *h_src, const int32_t *h_dst, const int32_t *h_pid, int directed,
                                                      int num_gpus, srw_graph **out) {
  if (!h_pid && n > 0) { srw_set_error("srw_graph_from_edges_multi_vcut: the partition-id column is the shard map"); return SRW_ERR_ARG; }
  return from_edges_multi(n, h_src, h_dst, h_pid, directed, num_gpus, out);
}
