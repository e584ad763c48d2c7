// walk.cu -- K4/K5/K6: first step + second-order walk steps on the device.
//
// Replaces RandomWalk.initFirstStep (RW:51-66), the per-walker hot loop of RandomWalk.randomWalk
// (RW:95-139) and RandomSample (RS:5-63).  Two samplers:
//
//  * EXACT  (walk_exact_kernel): RS:12-62 literally -- float32 bias weights (w/p, w/q, RS:34-38),
//    float64 left-to-right sum and inverse-CDF scan with `acc >= u` and the edges.head fallback
//    (RS:14-24) over the file-appearance-order row.  The O(d_c*d_p) `exists` scan (RS:38) is replaced
//    by a binary search in the sorted row of prev (same truth value).  Bit-identical to the oracle.
//  * ALIAS  (walk_alias_kernel): one alias-table proposal per trial from the static weights, accepted
//    with probability f(x)/M where f is the RS:33-41 bias factor; distribution-equal to EXACT,
//    bit-identical to the CPU twin in oracle/ (oracle_alias_walk).
//
// One walker per thread; a draw is Philox4x32-10 keyed by (seed; walker, step, trial).
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <vector>

#include "philox.cuh"
#include "srw_internal.h"

namespace {

#include "walk_conv.cuh"   // WalkArgs, FoldArgs, PeerTable, the state enum and the convergent kernels (v5)
#include "walk_exact.cuh"  // row_contains / hash_contains and K5, the exact (bit-parity) sampler kernels



// ranks -> original vertex ids, and the step count
__global__ void finalize_paths_kernel(int64_t n_walkers, int32_t stride, const int32_t *__restrict__ vids,
                                      const int32_t *__restrict__ lens, int32_t *paths, unsigned long long *stats) {
  unsigned long long steps = 0;
  // one warp per path row: coalesced, no 64-bit division
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t wk = warp; wk < n_walkers; wk += n_warps) {
    const int32_t len = __ldg(lens + wk);
    int32_t *row = paths + wk * stride;
    for (int32_t k = lane; k < stride; k += 32) row[k] = k < len ? __ldg(vids + row[k]) : -1;
    if (lane == 0) steps += (unsigned long long)(len - 1);
  }
  // warp-reduce then one atomic per warp
  for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
  if ((threadIdx.x & 31) == 0 && steps) atomicAdd(stats, steps);
}

// The same pass over the matrix as ONE contiguous array of int2 (even stride: rows are 8-byte aligned and adjacent),
// four independent 8-byte loads and eight id gathers in flight per thread instead of one dependent chain per lane.
// row = e2 / half by a 64-bit multiply-high (magic = ceil(2^64 / half), exact for e2 * half < 2^64).
__global__ void __launch_bounds__(256) finalize_paths_flat_kernel(int64_t n_pairs, uint32_t half, uint64_t magic, const int32_t *__restrict__ vids,
                                                                  const int32_t *__restrict__ lens, int2 *paths, unsigned long long *stats) {
  constexpr int U = 4;
  unsigned long long steps = 0;
  const int64_t base = (int64_t)blockIdx.x * (256 * U) + threadIdx.x;
  int2 v[U];
  int32_t len[U];
  uint32_t c[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t e = base + (int64_t)u * 256;
    v[u] = make_int2(0, 0); len[u] = 0; c[u] = 0;
    if (e < n_pairs) {
      v[u] = paths[e];
      const uint64_t r = __umul64hi((uint64_t)e, magic);
      c[u] = (uint32_t)((uint64_t)e - r * half) * 2u;
      len[u] = __ldg(lens + r);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t e = base + (int64_t)u * 256;
    if (e < n_pairs) {
      int2 o;
      o.x = (int32_t)c[u] < len[u] ? __ldg(vids + v[u].x) : -1;
      o.y = (int32_t)c[u] + 1 < len[u] ? __ldg(vids + v[u].y) : -1;
      paths[e] = o;
      if (c[u] == 0) steps += (unsigned long long)(len[u] - 1);
    }
  }
  for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
  if ((threadIdx.x & 31) == 0 && steps) atomicAdd(stats, steps);
}

// id-space walks need no translation: count the steps and pad the rows of walkers that stopped early (dead ends) with -1
__global__ void count_steps_kernel(int64_t n_walkers, int32_t stride, const int32_t *__restrict__ lens, int32_t *paths, unsigned long long *stats) {
  unsigned long long steps = 0;
  for (int64_t wk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; wk < n_walkers; wk += (int64_t)gridDim.x * blockDim.x) {
    const int32_t len = __ldg(lens + wk);
    steps += (unsigned long long)(len - 1);
    for (int32_t k = len; k < stride; ++k) paths[wk * stride + k] = -1;
  }
  for (int o = 16; o > 0; o >>= 1) steps += __shfl_down_sync(0xffffffffu, steps, o);
  if ((threadIdx.x & 31) == 0 && steps) atomicAdd(stats, steps);
}

// ---- KAT kernels: one thread, same device functions as the exact walk ----
__global__ void kat_sample_kernel(int64_t n, const float *w, float u, int64_t *out) {
  *out = cdf_pick(n, u, [&](int64_t j) { return w[j]; });
}
__global__ void kat_second_order_kernel(float p, float q, int32_t prev, int64_t np, const int32_t *pd_sorted, int64_t nc,
                                        const int32_t *cd, const float *cw, float u, float *w_out, int64_t *k_out) {
  for (int64_t j = 0; j < nc; ++j)
    w_out[j] = biased_weight(p, q, prev, cd[j], cw[j], row_contains(pd_sorted, 0, np, cd[j]));
  if (k_out) *k_out = cdf_pick(nc, u, [&](int64_t j) { return w_out[j]; });
}
__global__ void kat_philox_kernel(const uint32_t *ctr, const uint32_t *key, uint32_t *out) {
  Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

thread_local srw_walk_info t_info = {};
thread_local int t_collect_stats = 0;
thread_local const char *t_kernel = "";   // the walk kernel (with its template arguments) of this thread's last launch

// Per-thread launch context, created once: no cudaMalloc / cudaFree / event creation on the call
// path (those take driver-wide locks and serialise against other tools using the driver).
struct LaunchCtx {
  int device = -1;
  cudaEvent_t a = nullptr, b = nullptr, done = nullptr;
  bool pending = false;                    // kernels enqueued, srw_walk_wait not yet called
  unsigned long long *d_stats = nullptr;   // [4]
  unsigned long long *h_stats = nullptr;   // pinned [4]
  srw_status init(int dev) {
    if (device == dev) return SRW_OK;
    release();
    SRW_CUDA(cudaEventCreate(&a));
    SRW_CUDA(cudaEventCreate(&b));
    SRW_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    SRW_CUDA(cudaMalloc(&d_stats, 4 * sizeof(unsigned long long)));
    SRW_CUDA(cudaMallocHost(&h_stats, 4 * sizeof(unsigned long long)));
    device = dev;
    return SRW_OK;
  }
  void release() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    if (done) cudaEventDestroy(done);
    if (d_stats) cudaFree(d_stats);
    if (h_stats) cudaFreeHost(h_stats);
    a = b = done = nullptr; d_stats = h_stats = nullptr; device = -1; pending = false;
  }
};
thread_local LaunchCtx t_ctx;

}  // namespace

// Enqueues the walk kernel on l.stream and the rank -> id pass (+ the 32-byte statistics copy) on `fin`: the same stream
// for the blocking call, the library's high-priority finalisation stream for the asynchronous one -- there the caller's
// stream is free for the next round's walk as soon as this round's walk kernel has been launched.
static srw_status walk_enqueue(const srw_graph *g, const srw_params *p, const WalkLaunch &l, LaunchCtx &ev, cudaStream_t fin, bool own_fin) {
  if (p->walk_length < 0 || l.n_walkers < 0) { srw_set_error("walkLength and the walker count must be >= 0"); return SRW_ERR_ARG; }
  if (!(p->p > 0.0) || !(p->q > 0.0)) { srw_set_error("p and q must be > 0"); return SRW_ERR_ARG; }
  ev.pending = false;
  const bool peer = g->shard_world > 1;
  if (peer && g->vcut) { srw_set_error("a shard built from the partition-id column (VCut shard map) is walked by the migrating walk (srw_mig_*, --gpus N) only"); return SRW_ERR_UNSUPPORTED; }
  if (peer) {
    // one shard of several: only the peer-gather walk runs through this entry point (the tuple-exchange
    // walk is driven per super-step through srw_shard_step)
    for (int r = 0; r < g->shard_world; ++r)
      if (!g->peer_attached[r]) { srw_set_error("this handle is shard %d of %d and shard %d is not attached: attach every peer (srw_shard_attach_*) for the peer-gather walk, or use the srw_shard_* super-step calls", g->shard_rank, g->shard_world, r); return SRW_ERR_ARG; }
    if (p->sampler == SRW_SAMPLER_EXACT || g->directed || g->has_alias) { srw_set_error("the peer-gather walk covers undirected, unweighted graphs with --sampler alias|fold"); return SRW_ERR_UNSUPPORTED; }
  }
  if (l.n_walkers == 0 || g->nv == 0) return SRW_OK;
  const bool exact = p->sampler == SRW_SAMPLER_EXACT;
  if (p->sampler != SRW_SAMPLER_EXACT && p->sampler != SRW_SAMPLER_ALIAS && p->sampler != SRW_SAMPLER_ALIAS_FOLD) { srw_set_error("unknown sampler %d", p->sampler); return SRW_ERR_ARG; }
  if (exact && g->nnz > 0 && !g->d_col_app) { srw_set_error("graph was built without SRW_BUILD_EXACT"); return SRW_ERR_ARG; }
  if (!exact && p->u_mode == SRW_U_CONST) { srw_set_error("the constant-u generator is defined for --sampler exact only"); return SRW_ERR_ARG; }
  SRW_CUDA(cudaSetDevice(g->device));
  WalkArgs a{};
  a.off = g->d_off; a.col = g->d_col; a.slot = g->d_slot; a.col_app = g->d_col_app; a.w_app = g->d_w_app; a.vids = g->d_vids;
  a.nv = g->nv; a.walker_first = l.walker_first; a.n_walkers = l.n_walkers; a.stride = p->walk_length + 2;
  a.seed_lo = (uint32_t)p->seed; a.seed_hi = (uint32_t)(p->seed >> 32);
  srw_alias_thresholds(p->p, p->q, &a.t_ret, &a.t_common, &a.t_far);
  a.p = (float)p->p; a.q = (float)p->q; a.u_mode = p->u_mode; a.u_const = p->u_const;
  a.paths = l.d_paths; a.lens = l.d_lens;
  SRW_TRY(ev.init(g->device));
  unsigned long long *d_stats = ev.d_stats;
  SRW_CUDA(cudaMemsetAsync(d_stats, 0, 4 * sizeof(unsigned long long), l.stream));
  a.stats = d_stats;
  SRW_CUDA(cudaEventRecord(ev.a, l.stream));
  bool ids = false;                      // the walk kernel wrote original ids (id-space fold): no translation below
  // ONE kernel per sampler (+ its instrumented variant).  Gathers carry L2::64B (VAR = 1): a missing 16/32-byte gather fills 64
  // instead of 128 bytes at the same request rate (profiles/README.md).  Earlier kernel generations live in profiles/museum/.
  if (exact) {
    t_kernel = "walk_exact_cert2_kernel";
    walk_exact_cert2_kernel<<<(unsigned)((l.n_walkers + 7) / 8), 256, 0, l.stream>>>(a, g->d_hash);
  } else {
    const unsigned grid = (unsigned)((l.n_walkers + 255) / 256);
    const bool st = t_collect_stats != 0;
    // SRW_SAMPLER_ALIAS_FOLD: undirected + 1/p > max(1, 1/q), else the classic sampler (the CPU twin applies the same rule,
    // oracle_alias_walk); unweighted graphs fold over multiplicities (walk_fold_conv_kernel), weighted ones over bundle weights
    FoldArgs f{};
    bool fold = false;
    // a SRW_BUILD_LEAN handle holds only d_ent + d_hash_id: every alias-class walk on it runs through the id-space fold kernel
    const bool lean = g->lean && g->d_ent && g->d_hash_id && g->ent_ids;
    if ((peer || lean || p->sampler == SRW_SAMPLER_ALIAS_FOLD) && (peer || lean || (g->d_ent && g->d_hash)) && !g->directed && !g->has_alias) {
      fold = srw_fold_args(p->p, p->q, p->sampler == SRW_SAMPLER_ALIAS_FOLD, &f);
      f.ent = g->d_ent; f.hash = g->d_hash;
      if ((peer || lean) && !fold) {
        // classic rejection under M = max(1/p, 1, 1/q) through the same kernel: no return component
        f.a = 0.0; f.mp = 1.0; f.t_ret = a.t_ret; f.t_common = a.t_common; f.t_far = a.t_far;
        if (lean) fold = true;
      }
    }
    FoldArgs wf{};
    const bool wfold = !peer && p->sampler == SRW_SAMPLER_ALIAS_FOLD && g->has_alias && !g->directed && g->d_slotw && g->d_meta &&
                       g->d_hash && srw_fold_args(p->p, p->q, true, &wf);
    ids = fold && !peer && g->ent_ids && g->d_hash_id;
    if (g->ent_ids && !ids && !g->has_alias && g->d_ent) {
      // the neighbour entries of this handle carry original ids: only the id-space fold kernel may read them, and the classic
      // kernel below does not (it reads d_col / d_hash, which stay in rank space)
      fold = false;
    }
    if (peer || fold) {
      PeerTable pt{};
      if (peer) {
        pt.world = g->shard_world;
        for (int r = 0; r <= g->shard_world; ++r) pt.first[r] = g->bounds[(size_t)r];
        for (int r = 0; r < g->shard_world; ++r) { pt.off[r] = g->peer_off[r]; pt.ent[r] = g->peer_ent[r]; pt.hash[r] = g->peer_hash[r]; }
        t_kernel = st ? "walk_fold_conv_kernel<1,1,0,4,0>" : "walk_fold_conv_kernel<0,1,1,4,0>";
        if (st) walk_fold_conv_kernel<true, true, 0><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else walk_fold_conv_kernel<false, true, 1><<<grid, 256, 0, l.stream>>>(a, f, pt);
      } else if (ids) {                   // id space: entries and hash sets carry original ids, the walk emits ids (no rank -> id pass)
        f.hash = g->d_hash_id;
        t_kernel = st ? "walk_fold_conv_kernel<1,0,0,4,1>" : "walk_fold_conv_kernel<0,0,1,4,1>";
        if (st) walk_fold_conv_kernel<true, false, 0, 4, true><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else walk_fold_conv_kernel<false, false, 1, 4, true><<<grid, 256, 0, l.stream>>>(a, f, pt);
      } else {
        t_kernel = st ? "walk_fold_conv_kernel<1,0,0,4,0>" : "walk_fold_conv_kernel<0,0,1,4,0>";
        if (st) walk_fold_conv_kernel<true, false, 0><<<grid, 256, 0, l.stream>>>(a, f, pt);
        else walk_fold_conv_kernel<false, false, 1><<<grid, 256, 0, l.stream>>>(a, f, pt);
      }
    } else if (wfold) {
      // weighted alias-fold: bundle weights in 32-byte slots, row weight sums in the descriptors
      t_kernel = st ? "walk_wfold_conv_kernel<1,0>" : "walk_wfold_conv_kernel<0,1>";
      if (st) walk_wfold_conv_kernel<true, 0><<<grid, 256, 0, l.stream>>>(a, wf, g->d_meta, g->d_hash, g->d_slotw);
      else walk_wfold_conv_kernel<false, 1><<<grid, 256, 0, l.stream>>>(a, wf, g->d_meta, g->d_hash, g->d_slotw);
    } else {
      // the classic alias sampler (weighted or directed graphs, or (p, q) where folding does not apply)
      if (!g->d_meta || !g->d_hash) { srw_set_error("graph was built without SRW_BUILD_ALIAS"); return SRW_ERR_ARG; }
      const RowMeta *mt = g->d_meta;
      const int32_t *hs = g->d_hash;
      if (g->has_alias) {
        t_kernel = st ? "walk_alias_conv_kernel<1,1,0>" : "walk_alias_conv_kernel<1,0,1>";
        if (st) walk_alias_conv_kernel<true, true, 0><<<grid, 256, 0, l.stream>>>(a, mt, hs);
        else walk_alias_conv_kernel<true, false, 1><<<grid, 256, 0, l.stream>>>(a, mt, hs);
      } else {
        t_kernel = st ? "walk_alias_conv_kernel<0,1,0>" : "walk_alias_conv_kernel<0,0,1>";
        if (st) walk_alias_conv_kernel<false, true, 0><<<grid, 256, 0, l.stream>>>(a, mt, hs);
        else walk_alias_conv_kernel<false, false, 1><<<grid, 256, 0, l.stream>>>(a, mt, hs);
      }
    }
  }
  SRW_CUDA(cudaEventRecord(ev.b, l.stream));
  if (own_fin) SRW_CUDA(cudaStreamWaitEvent(fin, ev.b, 0));
  {
    const int64_t total = l.n_walkers * (int64_t)a.stride;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    const bool flat = (a.stride & 1) == 0 && (reinterpret_cast<uintptr_t>(l.d_paths) & 7) == 0;
    if (ids) {
      count_steps_kernel<<<(unsigned)std::min<int64_t>((l.n_walkers + 255) / 256, 148 * 8), 256, 0, fin>>>(l.n_walkers, a.stride, l.d_lens, l.d_paths, d_stats);
    } else if (flat && total > 0) {
      const uint32_t half = (uint32_t)a.stride / 2;
      const uint64_t magic = ~0ULL / half + 1;                  // ceil(2^64 / half) (half >= 1; half == 1: wraps to 0, handled below)
      const int64_t n_pairs = total / 2;
      if (half == 1) finalize_paths_kernel<<<(unsigned)blocks, 256, 0, fin>>>(l.n_walkers, a.stride, g->d_vids, l.d_lens, l.d_paths, d_stats);
      else finalize_paths_flat_kernel<<<(unsigned)((n_pairs + 1023) / 1024), 256, 0, fin>>>(n_pairs, half, magic, g->d_vids, l.d_lens, reinterpret_cast<int2 *>(l.d_paths), d_stats);
    } else {
      finalize_paths_kernel<<<(unsigned)blocks, 256, 0, fin>>>(l.n_walkers, a.stride, g->d_vids, l.d_lens, l.d_paths, d_stats);
    }
  }
  SRW_CUDA(cudaMemcpyAsync(ev.h_stats, d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, fin));
  SRW_CUDA(cudaEventRecord(ev.done, fin));
  SRW_CUDA(cudaGetLastError());
  ev.pending = true;
  return SRW_OK;
}

// Blocks until an enqueued walk has finished and reports it.
static srw_status walk_finish(LaunchCtx &ev, srw_walk_info *out) {
  srw_walk_info wi{};
  if (ev.pending) {
    SRW_CUDA(cudaEventSynchronize(ev.done));
    SRW_CUDA(cudaGetLastError());
    float ms = 0.f;
    SRW_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    wi.kernel_ms = ms;
    wi.kernel_launches = 2;
    wi.steps = (int64_t)ev.h_stats[0];
    wi.proposals = (int64_t)ev.h_stats[1];
    wi.member_tests = (int64_t)ev.h_stats[2];
    wi.probes_log2 = (int64_t)ev.h_stats[3];
    ev.pending = false;
  }
  if (out) *out = wi;
  return SRW_OK;
}

srw_status srw_walk_launch(const srw_graph *g, const srw_params *p, const WalkLaunch &l) {
  t_info = srw_walk_info{};
  SRW_TRY(walk_enqueue(g, p, l, t_ctx, l.stream, false));
  return walk_finish(t_ctx, &t_info);
}

// ---- asynchronous rounds: tickets from a small pool, one high-priority finalisation stream per device ----
struct srw_walk_ticket {
  LaunchCtx ctx;
};
namespace {
std::mutex g_ticket_mu;
std::vector<srw_walk_ticket *> g_ticket_pool;
cudaStream_t g_fin_stream[64] = {};
srw_status fin_stream_for(int dev, cudaStream_t *out) {
  if (dev < 0 || dev >= 64) { srw_set_error("device index %d out of range", dev); return SRW_ERR_ARG; }
  std::lock_guard<std::mutex> lk(g_ticket_mu);
  if (!g_fin_stream[dev]) {
    int lo = 0, hi = 0;
    SRW_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SRW_CUDA(cudaStreamCreateWithPriority(&g_fin_stream[dev], cudaStreamNonBlocking, hi));
  }
  *out = g_fin_stream[dev];
  return SRW_OK;
}
}  // namespace

extern "C" srw_status srw_walk_device_async(const srw_graph *g, const srw_params *params, uint64_t walker_first, int64_t n_walkers,
                                            int32_t *d_paths, int32_t *d_lens, void *stream, srw_walk_ticket **ticket) {
  SRW_TRY(srw_require_device());
  if (!g || !params || !ticket || (n_walkers > 0 && (!d_paths || !d_lens))) { srw_set_error("srw_walk_device_async: bad argument"); return SRW_ERR_ARG; }
  SRW_CUDA(cudaSetDevice(g->device));
  cudaStream_t fin;
  SRW_TRY(fin_stream_for(g->device, &fin));
  srw_walk_ticket *t = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_ticket_mu);
    for (size_t i = 0; i < g_ticket_pool.size(); ++i)
      if (g_ticket_pool[i]->ctx.device == g->device || g_ticket_pool[i]->ctx.device < 0) { t = g_ticket_pool[i]; g_ticket_pool.erase(g_ticket_pool.begin() + (long)i); break; }
  }
  if (!t) t = new srw_walk_ticket();
  WalkLaunch l{walker_first, n_walkers, d_paths, d_lens, (cudaStream_t)stream};
  const srw_status st = walk_enqueue(g, params, l, t->ctx, fin, true);
  if (st != SRW_OK) {
    std::lock_guard<std::mutex> lk(g_ticket_mu);
    g_ticket_pool.push_back(t);
    return st;
  }
  *ticket = t;
  return SRW_OK;
}

extern "C" srw_status srw_walk_wait(srw_walk_ticket *ticket, srw_walk_info *info) {
  if (!ticket) return SRW_ERR_ARG;
  const srw_status st = walk_finish(ticket->ctx, info);
  std::lock_guard<std::mutex> lk(g_ticket_mu);
  g_ticket_pool.push_back(ticket);
  return st;
}

void srw_set_walk_info(double kernel_ms, int64_t launches, int64_t steps, int64_t proposals, int64_t member_tests, int64_t probes_log2) {
  t_info.kernel_ms = kernel_ms; t_info.kernel_launches = launches; t_info.steps = steps;
  t_info.proposals = proposals; t_info.member_tests = member_tests; t_info.probes_log2 = probes_log2;
}

extern "C" srw_status srw_last_walk_info(srw_walk_info *out) {
  if (!out) return SRW_ERR_ARG;
  *out = t_info;
  return SRW_OK;
}
extern "C" const char *srw_last_walk_kernel(void) { return t_kernel; }

extern "C" srw_status srw_walk_collect_stats(int enable) {
  t_collect_stats = enable;
  return SRW_OK;
}

// ---- KAT entry points (RS:12-62 on the device) ----
namespace {
template <class T>
struct DevArr {
  T *p = nullptr;
  ~DevArr() { if (p) cudaFree(p); }
  cudaError_t upload(const T *h, int64_t n) {
    cudaError_t e = cudaMalloc(&p, (size_t)(n > 0 ? n : 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    return (n > 0 && h) ? cudaMemcpy(p, h, (size_t)n * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
  }
};
}  // namespace

extern "C" srw_status srw_sample(int64_t n, const int32_t *h_dst, const float *h_w, float u, int32_t *dst_out, float *w_out) {
  SRW_TRY(srw_require_device());
  if (n <= 0 || !h_dst || !h_w) { srw_set_error("srw_sample: empty edge array (reference: edges.head on empty throws)"); return SRW_ERR_ARG; }
  DevArr<float> w; DevArr<int64_t> k;
  SRW_CUDA(w.upload(h_w, n));
  SRW_CUDA(k.upload(nullptr, 1));
  kat_sample_kernel<<<1, 1>>>(n, w.p, u, k.p);
  int64_t hk = 0;
  SRW_CUDA(cudaMemcpy(&hk, k.p, 8, cudaMemcpyDeviceToHost));
  if (dst_out) *dst_out = h_dst[hk];
  if (w_out) *w_out = h_w[hk];
  return SRW_OK;
}

static srw_status second_order(float p, float q, int32_t prev, int64_t np, const int32_t *h_pdst, int64_t nc,
                               const int32_t *h_cdst, const float *h_cw, float u, float *h_out, int64_t *k_out) {
  SRW_TRY(srw_require_device());
  if (nc <= 0 || !h_cdst || !h_cw) { srw_set_error("second-order sample: empty currNeighbors"); return SRW_ERR_ARG; }
  std::vector<int32_t> ps(h_pdst, h_pdst + (np > 0 ? np : 0));
  std::sort(ps.begin(), ps.end());
  DevArr<int32_t> pd, cd; DevArr<float> cw, wo; DevArr<int64_t> k;
  SRW_CUDA(pd.upload(ps.data(), np));
  SRW_CUDA(cd.upload(h_cdst, nc));
  SRW_CUDA(cw.upload(h_cw, nc));
  SRW_CUDA(wo.upload(nullptr, nc));
  SRW_CUDA(k.upload(nullptr, 1));
  kat_second_order_kernel<<<1, 1>>>(p, q, prev, np, pd.p, nc, cd.p, cw.p, u, wo.p, k_out ? k.p : nullptr);
  SRW_CUDA(cudaMemcpy(h_out, wo.p, (size_t)nc * 4, cudaMemcpyDeviceToHost));
  if (k_out) SRW_CUDA(cudaMemcpy(k_out, k.p, 8, cudaMemcpyDeviceToHost));
  return SRW_OK;
}

extern "C" srw_status srw_second_order_weights(float p, float q, int32_t prev, int64_t np, const int32_t *h_pdst, int64_t nc,
                                               const int32_t *h_cdst, const float *h_cw, float *h_out) {
  if (!h_out) return SRW_ERR_ARG;
  return second_order(p, q, prev, np, h_pdst, nc, h_cdst, h_cw, 0.f, h_out, nullptr);
}
extern "C" srw_status srw_second_order_sample(float p, float q, int32_t prev, int64_t np, const int32_t *h_pdst, int64_t nc,
                                              const int32_t *h_cdst, const float *h_cw, float u, int32_t *dst_out, float *w_out) {
  std::vector<float> w((size_t)(nc > 0 ? nc : 1));
  int64_t k = 0;
  SRW_TRY(second_order(p, q, prev, np, h_pdst, nc, h_cdst, h_cw, u, w.data(), &k));
  if (dst_out) *dst_out = h_cdst[k];
  if (w_out) *w_out = w[k];     // the BIASED weight (T-RS:76)
  return SRW_OK;
}
extern "C" srw_status srw_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  SRW_TRY(srw_require_device());
  DevArr<uint32_t> c, k, o;
  SRW_CUDA(c.upload(ctr, 4)); SRW_CUDA(k.upload(key, 2)); SRW_CUDA(o.upload(nullptr, 4));
  kat_philox_kernel<<<1, 1>>>(c.p, k.p, o.p);
  SRW_CUDA(cudaMemcpy(out, o.p, 16, cudaMemcpyDeviceToHost));
  return SRW_OK;
}
