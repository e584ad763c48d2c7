// shard.cu -- K7/K8 device side: the walk over a vertex-range-sharded graph.
//
// Replaces the reference's walker routing (URW:103-112 / VRW:121-134 re-key, RW:186-192 shuffle,
// RW:200-216 finished/unfinished split).  The reference ships (path, prevNeighbors, completed) per
// walker and super-step (RW:135); here a walker is a fixed 32-byte tuple and the path is assembled on
// the walker's HOME shard (the owner of its start vertex) from 16-byte (walker, position, vertex)
// records.  Each rank runs the same per-lane state machine as walk_alias_sm_kernel on the walkers
// resident in its inbox and stops a walker the moment it needs a row another rank owns:
//
//   MSG_STEP   -> owner(curr): load row extent of curr, draw proposal `trial`
//   MSG_SEARCH -> owner(prev): is x in N(prev)?  accept -> owner(x), reject -> back to owner(curr)
//   MSG_FIN    -> home: final path length
//
// Every decision is the pure function of (seed; walker, step, trial) used on one GPU, so the emitted
// paths are identical for any number of shards and any arrival order.  The exchange itself (counts,
// all-to-all over NCCL/NVLink, termination) is driven by the host (sharded.py) on buffers this file
// fills; no collective is issued from here.
#include <cuda_runtime.h>

#include <vector>

#include "philox.cuh"
#include "srw_internal.h"

namespace {

struct __align__(16) WalkerMsg {   // 32 bytes
  uint64_t walker;
  int32_t curr, prev, x;
  uint32_t y, trial;
  uint16_t len;     // ids already in the path
  uint8_t state;
  uint8_t pad;
};
struct __align__(16) PathRec {     // 16 bytes
  uint64_t walker;
  int32_t pos;
  int32_t value;
};
static_assert(sizeof(WalkerMsg) == 32 && sizeof(PathRec) == 16, "wire formats");

enum : int { MSG_STEP = 0, MSG_SEARCH = 1, MSG_FIN = 2 };
enum : int { L_EXT_CURR = 0, L_PROPOSE = 1, L_EXT_PREV = 2, L_SEARCH = 3, L_EXIT = 4, L_HASH = 5 };

struct ShardArgs {
  const int64_t *__restrict__ off;
  const int32_t *__restrict__ col;
  const AliasSlot *__restrict__ slot;
  const int32_t *__restrict__ hash;   // per-row neighbour hash sets (derived placement, srw_internal.h)
  int64_t nv, row_first, row_last;
  int world, rank;
  int64_t bounds[SRW_MAX_SHARDS + 1];
  int32_t stride;
  uint32_t seed_lo, seed_hi;
  uint64_t t_ret, t_common, t_far;
  int64_t round_first, n_rounds;
  const WalkerMsg *inbox;
  int64_t n_in;
  WalkerMsg *stage_msgs;     // [n_in]
  int32_t *stage_dest;       // [n_in], -1 = nothing to send
  PathRec *stage_recs;       // [rec_cap]
  int32_t *rec_dest;
  int64_t rec_cap;
  unsigned long long *counters;   // [0] record cursor, [1] steps, [2 .. 2+world) msgs per dest, [2+world .. 2+2*world) records per dest
  int32_t *paths, *lens;
};

__device__ __forceinline__ int owner_of(const ShardArgs &a, int32_t v) {
  int o = 0;
  while (o + 1 < a.world && (int64_t)v >= a.bounds[o + 1]) o++;
  return o;
}
// Path rows are homed round-robin (home(v) = v mod world) so that every rank stores ~|V|/world rows per
// round even though the edge-balanced vertex ranges hold very different numbers of vertices.
__device__ __forceinline__ int64_t home_rows(const ShardArgs &a) { return (a.nv - a.rank + a.world - 1) / a.world; }
__device__ __forceinline__ int64_t home_row(const ShardArgs &a, uint64_t walker) {
  const int64_t round = (int64_t)(walker / (uint64_t)a.nv), v = (int64_t)(walker % (uint64_t)a.nv);
  return (round - a.round_first) * home_rows(a) + v / a.world;
}

template <bool HAS_ALIAS>
__global__ void __launch_bounds__(256) shard_walk_kernel(ShardArgs a) {
  // per-block tallies (message / record counts per destination, steps): same-address global atomics
  // from every lane serialise (~1 per ns chip-wide) and dominated the first version of this kernel
  __shared__ unsigned int s_msg[SRW_MAX_SHARDS], s_rec[SRW_MAX_SHARDS];
  __shared__ unsigned long long s_steps;
  if (threadIdx.x < SRW_MAX_SHARDS) { s_msg[threadIdx.x] = 0; s_rec[threadIdx.x] = 0; }
  if (threadIdx.x == 0) s_steps = 0;
  __syncthreads();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool live = i < a.n_in;
  WalkerMsg m;
  if (live) m = a.inbox[i];
  else { m.walker = 0; m.curr = m.prev = m.x = 0; m.y = m.trial = 0; m.len = 0; m.state = MSG_FIN + 1; m.pad = 0; }
  const uint64_t walker = m.walker;
  int32_t curr = m.curr, prev = m.prev, x = m.x;
  uint32_t y = m.y, trial = m.trial, coin = 0, lo = 0, hi = 0, deg = 0, pdeg = 0, pnb = 0, bkt = 0;
  int32_t len = m.len;
  int64_t off = 0, poff = 0;
  uint64_t k = 0;
  bool have_ext = false;
  const int home = (int)((walker % (uint64_t)a.nv) % (uint64_t)a.world);
  const uint64_t t_lo = a.t_common < a.t_far ? a.t_common : a.t_far;
  const uint64_t t_hi = a.t_common < a.t_far ? a.t_far : a.t_common;
  unsigned long long steps = 0;

  int out_dest = -1;                 // at most one outgoing tuple per walker and super-step
  WalkerMsg out = m;
  auto emit = [&](int dest, int state) {
    out.walker = walker; out.curr = curr; out.prev = prev; out.x = x; out.y = y; out.trial = trial;
    out.len = (uint16_t)len; out.state = (uint8_t)state; out.pad = 0;
    out_dest = dest;
  };

  int st;
  if (!live) st = L_EXIT;
  else if (m.state == MSG_FIN) {     // home: RW:132 completed path
    a.lens[home_row(a, walker)] = len;
    st = L_EXIT;
  } else if (m.state == MSG_STEP) {
    const int co = owner_of(a, curr);          // freshly seeded walkers start on their home rank
    if (co == a.rank) st = L_EXT_CURR;
    else { emit(co, MSG_STEP); st = L_EXIT; }
  } else {
    st = L_EXT_PREV;
  }

  while (st != L_EXIT) {
    // ---- one memory access per lane ----
    int64_t e0 = 0, e1 = 0;
    int32_t v = 0, v_alias = 0;
    uint32_t thr = 0xFFFFFFFFu;
    int4 h0 = make_int4(0, 0, 0, 0), h1 = make_int4(0, 0, 0, 0);
    if (st == L_HASH) {
      const int4 *b = reinterpret_cast<const int4 *>(a.hash + (srw_hash_first(poff) + (int64_t)bkt) * 8);
      h0 = __ldg(b); h1 = __ldg(b + 1);
    } else if (st == L_EXT_CURR) {
      e0 = __ldg(a.off + (curr - a.row_first));
      e1 = __ldg(a.off + (curr - a.row_first) + 1);
    } else if (st == L_EXT_PREV) {
      e0 = __ldg(a.off + (prev - a.row_first));
      e1 = __ldg(a.off + (prev - a.row_first) + 1);
    } else if (st == L_PROPOSE) {
      if (HAS_ALIAS) {
        const int4 raw = __ldg(reinterpret_cast<const int4 *>(a.slot + off + (int64_t)k));
        thr = (uint32_t)raw.x; v = raw.y; v_alias = raw.z;
      } else {
        v = __ldg(a.col + off + (int64_t)k);
      }
    } else {
      v = __ldg(a.col + poff + (int64_t)((lo + hi) >> 1));
    }
    // ---- consume it ----
    int verdict = 0;   // 1 accept x, 2 reject (next trial), 3 draw the pending trial
    if (st == L_EXT_CURR) {
      off = e0; deg = (uint32_t)(e1 - e0); have_ext = true;
      if (deg == 0) {                                               // dead end (RW:59-62, RW:115-119)
        if (home == a.rank) a.lens[home_row(a, walker)] = len; else emit(home, MSG_FIN);
        st = L_EXIT;
        continue;
      }
      verdict = 3;
    } else if (st == L_EXT_PREV) {
      poff = e0; pdeg = (uint32_t)(e1 - e0);
      lo = 0; hi = pdeg;
      pnb = a.hash ? srw_hash_buckets(poff, pdeg) : 0u;
      if (pnb) { bkt = __umulhi(srw_hash32((uint32_t)x), pnb); st = L_HASH; } else st = L_SEARCH;
      if (pdeg == 0) verdict = ((uint64_t)y < a.t_far) ? 1 : 2;     // directed: prev may have no out-row here
    } else if (st == L_PROPOSE) {
      x = (HAS_ALIAS && !(coin < thr)) ? v_alias : v;
      if (len == 1 || deg == 1) verdict = 1;                        // first-order step / single choice
      else if (x == prev) verdict = ((uint64_t)y < a.t_ret) ? 1 : 2;
      else if ((uint64_t)y < t_lo) verdict = 1;
      else if ((uint64_t)y >= t_hi) verdict = 2;
      else {
        const int po = owner_of(a, prev);
        if (po == a.rank) st = L_EXT_PREV;
        else { emit(po, MSG_SEARCH); st = L_EXIT; continue; }
      }
    } else if (st == L_HASH) {
      const bool found = h0.x == x || h0.y == x || h0.z == x || h0.w == x || h1.x == x || h1.y == x || h1.z == x || h1.w == x;
      if (found) verdict = ((uint64_t)y < a.t_common) ? 1 : 2;
      else if (h1.w == -1) verdict = ((uint64_t)y < a.t_far) ? 1 : 2;
      else bkt = bkt + 1 == pnb ? 0 : bkt + 1;
    } else {
      const uint32_t mid = (lo + hi) >> 1;
      if (v == x) verdict = ((uint64_t)y < a.t_common) ? 1 : 2;
      else {
        if (v < x) lo = mid + 1; else hi = mid;
        if (lo >= hi) verdict = ((uint64_t)y < a.t_far) ? 1 : 2;
      }
    }
    if (verdict == 1) {
      // the step (walker, len) -> x is decided: deliver it to the home shard's path row
      if (home == a.rank) {
        a.paths[home_row(a, walker) * a.stride + len] = x;
      } else {
        // warp-aggregated slot reservation: one global atomic per group of lanes that reach this point together
        const unsigned act = __activemask();
        const int lane = threadIdx.x & 31, leader = __ffs(act) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(a.counters + 0, (unsigned long long)__popc(act));
        base = __shfl_sync(act, base, leader);
        const unsigned long long slot = base + __popc(act & ((1u << lane) - 1u));
        if (slot >= (unsigned long long)a.rec_cap) {
          // record buffer full: park the walker on this rank in its pre-decision state; the next
          // super-step recomputes the same decision (pure function of walker/step/trial)
          emit(a.rank, st == L_PROPOSE ? MSG_STEP : MSG_SEARCH);   // L_SEARCH / L_HASH / L_EXT_PREV re-run the membership test
          st = L_EXIT;
          continue;
        }
        PathRec r; r.walker = walker; r.pos = len; r.value = x;
        a.stage_recs[slot] = r;
        a.rec_dest[slot] = home;
        atomicAdd(&s_rec[home], 1u);
      }
      steps++;
      len++;
      prev = curr; curr = x; trial = 0; have_ext = false;
      if (len == a.stride) {                                        // RW:103,132
        if (home == a.rank) a.lens[home_row(a, walker)] = len; else emit(home, MSG_FIN);
        st = L_EXIT;
      } else {
        const int co = owner_of(a, curr);
        if (co == a.rank) st = L_EXT_CURR;
        else { emit(co, MSG_STEP); st = L_EXIT; }
      }
    } else if (verdict == 2) {
      trial++;
      const int co = owner_of(a, curr);
      if (co != a.rank) { emit(co, MSG_STEP); st = L_EXIT; }
      else if (!have_ext) st = L_EXT_CURR;
      else verdict = 3;
    }
    if (verdict == 3) {
      const Philox4 r = walker_rng(a.seed_lo, a.seed_hi, walker, (uint32_t)(len - 1), trial);
      k = __umul64hi(((uint64_t)r.x << 32) | (uint64_t)r.w, (uint64_t)deg);
      coin = r.y;
      y = r.z;
      st = L_PROPOSE;
    }
  }
  if (live) {
    a.stage_dest[i] = out_dest;
    if (out_dest >= 0) {
      a.stage_msgs[i] = out;
      atomicAdd(&s_msg[out_dest], 1u);
    }
    if (steps) atomicAdd(&s_steps, steps);
  }
  __syncthreads();
  if (threadIdx.x < a.world) {
    if (s_msg[threadIdx.x]) atomicAdd(a.counters + 2 + threadIdx.x, (unsigned long long)s_msg[threadIdx.x]);
    if (s_rec[threadIdx.x]) atomicAdd(a.counters + 2 + a.world + threadIdx.x, (unsigned long long)s_rec[threadIdx.x]);
  }
  if (threadIdx.x == 0 && s_steps) atomicAdd(a.counters + 1, s_steps);
}

__global__ void shard_seed_kernel(ShardArgs a, WalkerMsg *inbox) {
  const int64_t rows = home_rows(a), total = rows * a.n_rounds;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t round = a.round_first + i / rows, v = a.rank + (i % rows) * a.world;
    WalkerMsg m;
    m.walker = (uint64_t)round * (uint64_t)a.nv + (uint64_t)v;
    m.curr = (int32_t)v; m.prev = -1; m.x = 0; m.y = 0; m.trial = 0; m.len = 1; m.state = MSG_STEP; m.pad = 0;
    inbox[i] = m;
    a.paths[i * a.stride] = (int32_t)v;    // RW:84-87 path = Array(vId)
    a.lens[i] = 1;
  }
}

// stage -> contiguous per-destination segments (K7 bucket-by-owner)
template <class T>
__global__ void shard_scatter_kernel(int64_t n, const T *__restrict__ stage, const int32_t *__restrict__ dest,
                                     const unsigned long long *__restrict__ seg_first, unsigned long long *seg_cursor, T *out) {
  // per tile of blockDim items: shared histogram -> one global reservation per destination -> local ranks
  __shared__ unsigned int s_cnt[SRW_MAX_SHARDS], s_pos[SRW_MAX_SHARDS];
  __shared__ unsigned long long s_base[SRW_MAX_SHARDS];
  const int64_t tiles = (n + blockDim.x - 1) / blockDim.x;
  for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
    if (threadIdx.x < SRW_MAX_SHARDS) { s_cnt[threadIdx.x] = 0; s_pos[threadIdx.x] = 0; }
    __syncthreads();
    const int64_t i = t * blockDim.x + threadIdx.x;
    const int d = i < n ? dest[i] : -1;
    if (d >= 0) atomicAdd(&s_cnt[d], 1u);
    __syncthreads();
    if (threadIdx.x < SRW_MAX_SHARDS && s_cnt[threadIdx.x])
      s_base[threadIdx.x] = seg_first[threadIdx.x] + atomicAdd(seg_cursor + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
    __syncthreads();
    if (d >= 0) out[s_base[d] + atomicAdd(&s_pos[d], 1u)] = stage[i];
    __syncthreads();
  }
}

__global__ void shard_apply_kernel(ShardArgs a, const PathRec *__restrict__ recs, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const PathRec r = recs[i];
    a.paths[home_row(a, r.walker) * a.stride + r.pos] = r.value;
  }
}

}  // namespace

struct ShardScratch {
  void *stage_msgs = nullptr; int32_t *stage_dest = nullptr; int64_t msg_cap = 0;
  void *stage_recs = nullptr; int32_t *rec_dest = nullptr; int64_t rec_cap = 0;
  unsigned long long *d_counters = nullptr, *h_counters = nullptr;   // [2 + 4*MAX]
};
void srw_shard_scratch_free(ShardScratch *s) {
  if (!s) return;
  cudaFree(s->stage_msgs); cudaFree(s->stage_dest); cudaFree(s->stage_recs); cudaFree(s->rec_dest);
  cudaFree(s->d_counters); cudaFreeHost(s->h_counters);
  delete s;
}

namespace {
constexpr int kCounters = 2 + 4 * SRW_MAX_SHARDS;

srw_status fill_args(const srw_graph *g, const srw_params *p, int64_t round_first, int64_t n_rounds, ShardArgs *a) {
  if (!g || !p) { srw_set_error("shard call: null graph or params"); return SRW_ERR_ARG; }
  if (g->vcut) { srw_set_error("a shard built from the partition-id column (VCut shard map) is walked by the migrating walk (srw_mig_*) only"); return SRW_ERR_UNSUPPORTED; }
  // SRW_SAMPLER_ALIAS_FOLD (the default of srw_params_default) runs as the classic alias sampler here, as it does wherever
  // folding does not apply: the same distribution, the classic thresholds (the migrating walk, migrate.cu, implements the fold)
  if (p->sampler == SRW_SAMPLER_EXACT) { srw_set_error("the sharded walk implements --sampler alias (fold runs as alias), not exact"); return SRW_ERR_UNSUPPORTED; }
  if (p->walk_length < 0 || p->walk_length > 65000) { srw_set_error("sharded walk: walkLength must be in [0, 65000]"); return SRW_ERR_ARG; }
  if (!(p->p > 0.0) || !(p->q > 0.0)) { srw_set_error("p and q must be > 0"); return SRW_ERR_ARG; }
  SRW_CUDA(cudaSetDevice(g->device));
  memset(a, 0, sizeof(*a));
  a->off = g->d_off; a->col = g->d_col; a->slot = g->d_slot; a->hash = g->d_hash;
  a->nv = g->nv; a->row_first = g->row_first; a->row_last = g->row_last; a->world = g->shard_world; a->rank = g->shard_rank;
  for (int r = 0; r <= g->shard_world; ++r) a->bounds[r] = g->bounds[(size_t)r];
  a->stride = p->walk_length + 2;
  a->seed_lo = (uint32_t)p->seed; a->seed_hi = (uint32_t)(p->seed >> 32);
  srw_alias_thresholds(p->p, p->q, &a->t_ret, &a->t_common, &a->t_far);
  a->round_first = round_first; a->n_rounds = n_rounds;
  return SRW_OK;
}

srw_status ensure_scratch(const srw_graph *g, int64_t n_in, int64_t rec_cap, ShardScratch **out) {
  srw_graph *mg = const_cast<srw_graph *>(g);
  if (!mg->scratch) {
    mg->scratch = new ShardScratch();
    SRW_CUDA(cudaMalloc(&mg->scratch->d_counters, kCounters * 8));
    SRW_CUDA(cudaMallocHost(&mg->scratch->h_counters, kCounters * 8));
  }
  ShardScratch *s = mg->scratch;
  if (n_in > s->msg_cap) {
    cudaFree(s->stage_msgs); cudaFree(s->stage_dest);
    s->msg_cap = n_in + n_in / 4 + 1024;
    SRW_CUDA(cudaMalloc(&s->stage_msgs, (size_t)s->msg_cap * sizeof(WalkerMsg)));
    SRW_CUDA(cudaMalloc(&s->stage_dest, (size_t)s->msg_cap * 4));
  }
  if (rec_cap > s->rec_cap) {
    cudaFree(s->stage_recs); cudaFree(s->rec_dest);
    s->rec_cap = rec_cap;
    SRW_CUDA(cudaMalloc(&s->stage_recs, (size_t)s->rec_cap * sizeof(PathRec)));
    SRW_CUDA(cudaMalloc(&s->rec_dest, (size_t)s->rec_cap * 4));
  }
  *out = s;
  return SRW_OK;
}
inline unsigned blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > 148 * 32) b = 148 * 32;
  return (unsigned)(b < 1 ? 1 : b);
}
}  // namespace

extern "C" srw_status srw_graph_from_device_edges_sharded(int64_t n, const int32_t *d_src, const int32_t *d_dst, const float *d_w,
                                                          int directed, unsigned flags, int rank, int world, srw_graph **out) {
  return srw_build_graph_device_sharded(n, d_src, d_dst, d_w, directed, flags, rank, world, out);
}
extern "C" srw_status srw_graph_from_device_edges_vcut(int64_t n, const int32_t *d_src, const int32_t *d_dst, const int32_t *d_pid,
                                                       int directed, unsigned flags, int rank, int world, double hub_fraction, srw_graph **out) {
  return srw_build_graph_device_sharded(n, d_src, d_dst, nullptr, directed, flags, rank, world, out, d_pid, hub_fraction);
}

extern "C" srw_status srw_graph_shard_info(const srw_graph *g, int *rank, int *world, int64_t *row_first, int64_t *row_last,
                                           int64_t *bounds, int64_t *nnz_local) {
  if (!g) return SRW_ERR_ARG;
  if (rank) *rank = g->shard_rank;
  if (world) *world = g->shard_world;
  if (row_first) *row_first = g->row_first;
  if (row_last) *row_last = g->row_last;
  if (nnz_local) *nnz_local = g->nnz;
  if (bounds) for (int r = 0; r <= g->shard_world; ++r) bounds[r] = r < (int)g->bounds.size() ? g->bounds[(size_t)r] : g->nv;
  return SRW_OK;
}

// replicated hub rows of a table-mapped shard (0 / 0 / 0xFFFFFFFF on any other handle); seed_rows = vertices this shard starts walkers for
extern "C" srw_status srw_graph_hub_info(const srw_graph *g, int64_t *hub_rows, int64_t *hub_entries, uint32_t *hub_min_degree, int64_t *seed_rows) {
  if (!g) return SRW_ERR_ARG;
  const srw_graph *s = g->shards.empty() ? g : g->shards[0];
  if (hub_rows) *hub_rows = s->hub_rows;
  if (hub_entries) *hub_entries = s->hub_entries;
  if (hub_min_degree) *hub_min_degree = s->hub_deg;
  if (seed_rows) *seed_rows = s->vcut ? s->seed_rows : s->row_last - s->row_first;
  return SRW_OK;
}

extern "C" int srw_walker_msg_bytes(void) { return (int)sizeof(WalkerMsg); }
extern "C" int srw_path_rec_bytes(void) { return (int)sizeof(PathRec); }

// RW:81-87 + URW:81-87: one walker per homed vertex (v mod world == rank) and round, path = [v]
extern "C" srw_status srw_shard_seed(const srw_graph *g, const srw_params *params, int64_t round_first, int64_t n_rounds,
                                     void *d_inbox, int64_t cap, int64_t *n_seeded, int32_t *d_paths, int32_t *d_lens, void *stream) {
  SRW_TRY(srw_require_device());
  ShardArgs a;
  SRW_TRY(fill_args(g, params, round_first, n_rounds, &a));
  const int64_t total = ((g->nv - g->shard_rank + g->shard_world - 1) / g->shard_world) * n_rounds;
  if (total > cap) { srw_set_error("srw_shard_seed: inbox capacity %lld < %lld walkers", (long long)cap, (long long)total); return SRW_ERR_ARG; }
  a.paths = d_paths; a.lens = d_lens;
  if (total > 0) shard_seed_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(a, (WalkerMsg *)d_inbox);
  SRW_CUDA(cudaGetLastError());
  if (n_seeded) *n_seeded = total;
  return SRW_OK;
}

// One super-step on this rank: advance every resident walker until it needs a remote row, then
// bucket the outgoing tuples and path records by destination rank (contiguous segments, rank order).
extern "C" srw_status srw_shard_step(const srw_graph *g, const srw_params *params, int64_t round_first, int64_t n_rounds,
                                     const void *d_inbox, int64_t n_in, void *d_send_msgs, void *d_send_recs, int64_t rec_cap,
                                     int32_t *d_paths, int32_t *d_lens, int64_t *h_msg_counts, int64_t *h_rec_counts,
                                     int64_t *steps_done, void *stream_) {
  SRW_TRY(srw_require_device());
  cudaStream_t stream = (cudaStream_t)stream_;
  ShardArgs a;
  SRW_TRY(fill_args(g, params, round_first, n_rounds, &a));
  ShardScratch *s = nullptr;
  SRW_TRY(ensure_scratch(g, n_in, rec_cap, &s));
  const int W = g->shard_world;
  a.inbox = (const WalkerMsg *)d_inbox; a.n_in = n_in;
  a.stage_msgs = (WalkerMsg *)s->stage_msgs; a.stage_dest = s->stage_dest;
  a.stage_recs = (PathRec *)s->stage_recs; a.rec_dest = s->rec_dest; a.rec_cap = rec_cap;
  a.counters = s->d_counters; a.paths = d_paths; a.lens = d_lens;
  SRW_CUDA(cudaMemsetAsync(s->d_counters, 0, kCounters * 8, stream));
  if (n_in > 0) {
    const unsigned grid = (unsigned)((n_in + 255) / 256);
    if (g->has_alias) shard_walk_kernel<true><<<grid, 256, 0, stream>>>(a);
    else shard_walk_kernel<false><<<grid, 256, 0, stream>>>(a);
  }
  SRW_CUDA(cudaMemcpyAsync(s->h_counters, s->d_counters, kCounters * 8, cudaMemcpyDeviceToHost, stream));
  SRW_CUDA(cudaStreamSynchronize(stream));
  SRW_CUDA(cudaGetLastError());
  const int64_t n_recs = (int64_t)(s->h_counters[0] < (unsigned long long)rec_cap ? s->h_counters[0] : (unsigned long long)rec_cap);
  if (steps_done) *steps_done = (int64_t)s->h_counters[1];
  // segment starts (msgs and records), cursors zeroed
  unsigned long long seg[4 * SRW_MAX_SHARDS];
  unsigned long long accm = 0, accr = 0;
  for (int r = 0; r < W; ++r) {
    h_msg_counts[r] = (int64_t)s->h_counters[2 + r];
    h_rec_counts[r] = (int64_t)s->h_counters[2 + W + r];
    seg[r] = accm; accm += s->h_counters[2 + r];
    seg[SRW_MAX_SHARDS + r] = accr; accr += s->h_counters[2 + W + r];
    seg[2 * SRW_MAX_SHARDS + r] = 0; seg[3 * SRW_MAX_SHARDS + r] = 0;
  }
  SRW_CUDA(cudaMemcpyAsync(s->d_counters + 2, seg, sizeof(seg), cudaMemcpyHostToDevice, stream));
  unsigned long long *d_seg = s->d_counters + 2;
  if (n_in > 0 && accm > 0)
    shard_scatter_kernel<WalkerMsg><<<blocks_for(n_in), 256, 0, stream>>>(n_in, (const WalkerMsg *)s->stage_msgs, s->stage_dest, d_seg,
                                                                          d_seg + 2 * SRW_MAX_SHARDS, (WalkerMsg *)d_send_msgs);
  if (n_recs > 0)
    shard_scatter_kernel<PathRec><<<blocks_for(n_recs), 256, 0, stream>>>(n_recs, (const PathRec *)s->stage_recs, s->rec_dest,
                                                                          d_seg + SRW_MAX_SHARDS, d_seg + 3 * SRW_MAX_SHARDS,
                                                                          (PathRec *)d_send_recs);
  SRW_CUDA(cudaStreamSynchronize(stream));
  SRW_CUDA(cudaGetLastError());
  return SRW_OK;
}

// home shard: write received (walker, position, vertex) records into the path rows
extern "C" srw_status srw_shard_apply(const srw_graph *g, const srw_params *params, int64_t round_first, int64_t n_rounds,
                                      const void *d_recs, int64_t n_recs, int32_t *d_paths, void *stream) {
  SRW_TRY(srw_require_device());
  ShardArgs a;
  SRW_TRY(fill_args(g, params, round_first, n_rounds, &a));
  a.paths = d_paths;
  if (n_recs > 0) shard_apply_kernel<<<blocks_for(n_recs), 256, 0, (cudaStream_t)stream>>>(a, (const PathRec *)d_recs, n_recs);
  SRW_CUDA(cudaGetLastError());
  return SRW_OK;
}

namespace {
// ranks -> original vertex ids over this shard's home rows (one warp per row)
__global__ void shard_finalize_kernel(int64_t n_rows, int32_t stride, const int32_t *__restrict__ vids, const int32_t *__restrict__ lens,
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

// RW:168 totalPaths: this shard's home rows are complete; translate ranks to vertex ids
extern "C" srw_status srw_shard_finalize(const srw_graph *g, const srw_params *params, int64_t n_rows, int32_t *d_paths,
                                         const int32_t *d_lens, int64_t *steps, void *stream_) {
  SRW_TRY(srw_require_device());
  if (!g || !params) return SRW_ERR_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  ShardScratch *s = nullptr;
  SRW_TRY(ensure_scratch(g, 0, 0, &s));
  SRW_CUDA(cudaMemsetAsync(s->d_counters, 0, 8, stream));
  if (n_rows > 0)
    shard_finalize_kernel<<<blocks_for(n_rows * 32), 256, 0, stream>>>(n_rows, params->walk_length + 2, g->d_vids, d_lens, d_paths, s->d_counters);
  SRW_CUDA(cudaMemcpyAsync(s->h_counters, s->d_counters, 8, cudaMemcpyDeviceToHost, stream));
  SRW_CUDA(cudaStreamSynchronize(stream));
  SRW_CUDA(cudaGetLastError());
  if (steps) *steps = (int64_t)s->h_counters[0];
  return SRW_OK;
}

// ------------------------------------------------------------------------------------------
// Peer-gather mode: instead of moving walkers to the rows (super-steps above), make every shard's rows
// addressable from every GPU and let the walk kernel (walk.cu, walk_fold_kernel<PEER>) load remote rows
// over NVLink.  One process per GPU exchanges CUDA IPC handles of the three row arrays; shards living in
// one process (tests, or one process driving several GPUs) hand over plain pointers.
// ------------------------------------------------------------------------------------------
namespace {
struct ShardIpcBlob {
  uint32_t magic;
  int32_t rank, world, device;
  int64_t rows, nnz, hash_buckets;
  uint32_t has_off, has_ent, has_hash, pad;
  cudaIpcMemHandle_t off, ent, hash;
};
constexpr uint32_t kIpcMagic = 0x53525749u;   // "SRWI"
}  // namespace

extern "C" int srw_shard_ipc_bytes(void) { return (int)sizeof(ShardIpcBlob); }

extern "C" srw_status srw_shard_ipc_export(const srw_graph *g, void *h_blob) {
  SRW_TRY(srw_require_device());
  if (!g || !h_blob || g->shard_world < 1) { srw_set_error("srw_shard_ipc_export: bad argument"); return SRW_ERR_ARG; }
  SRW_CUDA(cudaSetDevice(g->device));
  ShardIpcBlob b;
  memset(&b, 0, sizeof(b));
  b.magic = kIpcMagic; b.rank = g->shard_rank; b.world = g->shard_world; b.device = g->device;
  b.rows = g->row_last - g->row_first; b.nnz = g->nnz; b.hash_buckets = g->hash_buckets;
  if (g->d_off) { SRW_CUDA(cudaIpcGetMemHandle(&b.off, g->d_off)); b.has_off = 1; }
  if (g->d_ent) { SRW_CUDA(cudaIpcGetMemHandle(&b.ent, g->d_ent)); b.has_ent = 1; }
  if (g->d_hash) { SRW_CUDA(cudaIpcGetMemHandle(&b.hash, g->d_hash)); b.has_hash = 1; }
  memcpy(h_blob, &b, sizeof(b));
  return SRW_OK;
}

extern "C" srw_status srw_shard_ipc_attach(srw_graph *g, const void *h_blob) {
  SRW_TRY(srw_require_device());
  if (!g || !h_blob) { srw_set_error("srw_shard_ipc_attach: bad argument"); return SRW_ERR_ARG; }
  ShardIpcBlob b;
  memcpy(&b, h_blob, sizeof(b));
  if (b.magic != kIpcMagic || b.world != g->shard_world || b.rank < 0 || b.rank >= g->shard_world) {
    srw_set_error("srw_shard_ipc_attach: blob does not describe a shard of this graph (world %d)", g->shard_world);
    return SRW_ERR_ARG;
  }
  if (b.rank == g->shard_rank || g->peer_attached[b.rank]) return SRW_OK;     // own rows / already mapped
  if (b.nnz > 0 && !b.has_ent) { srw_set_error("shard %d has no neighbour entries (weighted graph?): the peer-gather walk needs an unweighted SRW_BUILD_ALIAS build", b.rank); return SRW_ERR_UNSUPPORTED; }
  SRW_CUDA(cudaSetDevice(g->device));
  void *p = nullptr;
  if (b.has_off) { SRW_CUDA(cudaIpcOpenMemHandle(&p, b.off, cudaIpcMemLazyEnablePeerAccess)); g->peer_off[b.rank] = (const int64_t *)p; }
  if (b.has_ent) { SRW_CUDA(cudaIpcOpenMemHandle(&p, b.ent, cudaIpcMemLazyEnablePeerAccess)); g->peer_ent[b.rank] = (const NbrEntry *)p; }
  if (b.has_hash) { SRW_CUDA(cudaIpcOpenMemHandle(&p, b.hash, cudaIpcMemLazyEnablePeerAccess)); g->peer_hash[b.rank] = (const int32_t *)p; }
  g->peer_ipc[b.rank] = true;
  g->peer_attached[b.rank] = true;
  return SRW_OK;
}

extern "C" srw_status srw_shard_attach_local(srw_graph *g, const srw_graph *peer) {
  SRW_TRY(srw_require_device());
  if (!g || !peer || peer->shard_world != g->shard_world || peer->nv != g->nv || peer->bounds != g->bounds) {
    srw_set_error("srw_shard_attach_local: not two shards of the same graph");
    return SRW_ERR_ARG;
  }
  if (g->vcut) { srw_set_error("peer-gather needs vertex-range shards (this one follows the partition-id column)"); return SRW_ERR_UNSUPPORTED; }
  const int r = peer->shard_rank;
  if (r == g->shard_rank) return SRW_OK;
  if (peer->nnz > 0 && !peer->d_ent) { srw_set_error("shard %d has no neighbour entries (weighted graph?): the peer-gather walk needs an unweighted SRW_BUILD_ALIAS build", r); return SRW_ERR_UNSUPPORTED; }
  if (peer->device != g->device) {
    int can = 0;
    SRW_CUDA(cudaDeviceCanAccessPeer(&can, g->device, peer->device));
    if (!can) { srw_set_error("device %d cannot address device %d", g->device, peer->device); return SRW_ERR_UNSUPPORTED; }
    SRW_CUDA(cudaSetDevice(g->device));
    cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) SRW_CUDA(e);
    cudaGetLastError();
  }
  g->peer_off[r] = peer->d_off; g->peer_ent[r] = peer->d_ent; g->peer_hash[r] = peer->d_hash;
  g->peer_ipc[r] = false;
  g->peer_attached[r] = true;
  return SRW_OK;
}

// ---- the same, over a caller-provided block per shard.  One process per GPU: the block is a symmetric-memory
// allocation (CUDA VMM handles exchanged by the host runtime -- torch.distributed._symmetric_memory in
// sharded.py), whose peer mappings are ordinary NVLink peer pointers.  (Legacy cudaIpc mappings of
// cudaMalloc memory, srw_shard_ipc_* above, were measured ~35x slower for random 16-byte loads on this
// platform: profiles/README.md.)  Layout of a block, every part 256-byte aligned:
//   [ off: (rows + 1) * 8 ][ ent: nnz * 16 ][ hash: hash_buckets * 32 ]
namespace {
inline int64_t up256(int64_t x) { return (x + 255) & ~(int64_t)255; }
inline void block_layout(int64_t rows, int64_t nnz, int64_t hash_buckets, int64_t *o_ent, int64_t *o_hash, int64_t *total) {
  const int64_t e = up256((rows + 1) * 8), h = e + up256(nnz * (int64_t)sizeof(NbrEntry));
  if (o_ent) *o_ent = e;
  if (o_hash) *o_hash = h;
  if (total) *total = h + up256(hash_buckets * 32);
}
}  // namespace

extern "C" srw_status srw_shard_rows_info(const srw_graph *g, int64_t *rows, int64_t *nnz, int64_t *hash_buckets, int64_t *block_bytes) {
  if (!g) return SRW_ERR_ARG;
  const int64_t r = g->row_last - g->row_first, hb = g->d_hash ? g->hash_buckets : 0;
  if (rows) *rows = r;
  if (nnz) *nnz = g->nnz;
  if (hash_buckets) *hash_buckets = hb;
  block_layout(r, g->nnz, hb, nullptr, nullptr, block_bytes);
  return SRW_OK;
}

extern "C" srw_status srw_shard_rows_relocate(srw_graph *g, void *d_block, int64_t block_bytes) {
  SRW_TRY(srw_require_device());
  if (!g || !d_block || g->rows_external) { srw_set_error("srw_shard_rows_relocate: bad argument (or already relocated)"); return SRW_ERR_ARG; }
  if (g->nnz > 0 && (!g->d_ent || !g->d_hash)) { srw_set_error("this shard has no neighbour entries (weighted graph?): the peer-gather walk needs an unweighted SRW_BUILD_ALIAS build"); return SRW_ERR_UNSUPPORTED; }
  const int64_t rows = g->row_last - g->row_first, hb = g->d_hash ? g->hash_buckets : 0;
  int64_t o_ent, o_hash, total;
  block_layout(rows, g->nnz, hb, &o_ent, &o_hash, &total);
  if (block_bytes < total) { srw_set_error("srw_shard_rows_relocate: block of %lld bytes < %lld needed", (long long)block_bytes, (long long)total); return SRW_ERR_ARG; }
  SRW_CUDA(cudaSetDevice(g->device));
  char *b = (char *)d_block;
  SRW_CUDA(cudaMemcpy(b, g->d_off, (size_t)(rows + 1) * 8, cudaMemcpyDeviceToDevice));
  if (g->d_ent) SRW_CUDA(cudaMemcpy(b + o_ent, g->d_ent, (size_t)g->nnz * sizeof(NbrEntry), cudaMemcpyDeviceToDevice));
  if (g->d_hash) SRW_CUDA(cudaMemcpy(b + o_hash, g->d_hash, (size_t)hb * 32, cudaMemcpyDeviceToDevice));
  SRW_CUDA(cudaDeviceSynchronize());
  cudaFree(g->d_off); cudaFree(g->d_ent); cudaFree(g->d_hash);
  g->d_off = (int64_t *)b;
  g->d_ent = g->nnz > 0 ? (NbrEntry *)(b + o_ent) : nullptr;
  g->d_hash = hb > 0 ? (int32_t *)(b + o_hash) : nullptr;
  g->rows_external = true;
  const int r = g->shard_rank;
  g->peer_off[r] = g->d_off; g->peer_ent[r] = g->d_ent; g->peer_hash[r] = g->d_hash;
  return SRW_OK;
}

extern "C" srw_status srw_shard_attach_block(srw_graph *g, int peer_rank, const void *d_block, int64_t rows, int64_t nnz,
                                             int64_t hash_buckets) {
  if (!g || !d_block || peer_rank < 0 || peer_rank >= g->shard_world) { srw_set_error("srw_shard_attach_block: bad argument"); return SRW_ERR_ARG; }
  if (peer_rank == g->shard_rank) return SRW_OK;
  if (rows != g->bounds[(size_t)peer_rank + 1] - g->bounds[(size_t)peer_rank]) { srw_set_error("srw_shard_attach_block: shard %d has %lld rows, expected %lld", peer_rank, (long long)rows, (long long)(g->bounds[(size_t)peer_rank + 1] - g->bounds[(size_t)peer_rank])); return SRW_ERR_ARG; }
  int64_t o_ent, o_hash;
  block_layout(rows, nnz, hash_buckets, &o_ent, &o_hash, nullptr);
  const char *b = (const char *)d_block;
  g->peer_off[peer_rank] = (const int64_t *)b;
  g->peer_ent[peer_rank] = nnz > 0 ? (const NbrEntry *)(b + o_ent) : nullptr;
  g->peer_hash[peer_rank] = hash_buckets > 0 ? (const int32_t *)(b + o_hash) : nullptr;
  g->peer_ipc[peer_rank] = false;
  g->peer_attached[peer_rank] = true;
  return SRW_OK;
}
