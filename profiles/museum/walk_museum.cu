// walk_museum.cu -- TEST-ONLY library (libsrw_museum.so): launches the superseded kernel generations kept in this directory on
// a graph handle of the product library, so that tests/test_gpu_parity.py can check "every generation produces the product
// kernel's bits" and profiles/run_ab.py can time them, without any of it living in libsrw.so.
//   build: bash profiles/museum/build.sh        (nvcc, links against ../../stellar-random-walk_b200/libsrw.so)
// Variants: alias_v1 | alias_v2 | alias_v3 (classic alias sampler), fold_v4 (alias-fold, rank-space handles only),
//           exact_thread | exact_warp | exact_cert.
// Output: vertex RANKS (the product's rank -> id pass is not repeated here), d_lens as the product leaves them.
#include <cuda_runtime.h>
#include <string.h>

#include "../../stellar-random-walk_b200/csrc/philox.cuh"
#include "../../stellar-random-walk_b200/csrc/srw_internal.h"

namespace {
#include "../../stellar-random-walk_b200/csrc/walk_conv.cuh"
#include "../../stellar-random-walk_b200/csrc/walk_exact.cuh"
#include "alias_generations.cuh"
#include "exact_generations.cuh"
}  // namespace

extern "C" srw_status srw_museum_walk(const srw_graph *g, const srw_params *p, const char *variant, uint64_t walker_first,
                                      int64_t n_walkers, int32_t *d_paths, int32_t *d_lens) {
  if (!g || !p || !variant || g->shard_world > 1) return SRW_ERR_ARG;
  if (n_walkers <= 0 || g->nv == 0) return SRW_OK;
  SRW_CUDA(cudaSetDevice(g->device));
  WalkArgs a{};
  a.off = g->d_off; a.col = g->d_col; a.slot = g->d_slot; a.col_app = g->d_col_app; a.w_app = g->d_w_app; a.vids = g->d_vids;
  a.nv = g->nv; a.walker_first = walker_first; a.n_walkers = n_walkers; a.stride = p->walk_length + 2;
  a.seed_lo = (uint32_t)p->seed; a.seed_hi = (uint32_t)(p->seed >> 32);
  srw_alias_thresholds(p->p, p->q, &a.t_ret, &a.t_common, &a.t_far);
  a.p = (float)p->p; a.q = (float)p->q; a.u_mode = p->u_mode; a.u_const = p->u_const;
  a.paths = d_paths; a.lens = d_lens;
  unsigned long long *d_stats = nullptr;
  SRW_CUDA(cudaMalloc(&d_stats, 32));
  SRW_CUDA(cudaMemset(d_stats, 0, 32));
  a.stats = d_stats;
  const unsigned grid = (unsigned)((n_walkers + 255) / 256);
  srw_status rc = SRW_OK;
  if (!strcmp(variant, "exact_thread")) walk_exact_kernel<<<(unsigned)((n_walkers + 127) / 128), 128>>>(a);
  else if (!strcmp(variant, "exact_warp")) walk_exact_warp_kernel<<<(unsigned)((n_walkers + 7) / 8), 256>>>(a, g->d_hash);
  else if (!strcmp(variant, "exact_cert")) walk_exact_cert_kernel<<<(unsigned)((n_walkers + 7) / 8), 256>>>(a, g->d_hash);
  else if (!strcmp(variant, "alias_v1")) { if (g->has_alias) walk_alias_kernel<true, false><<<grid, 256>>>(a); else walk_alias_kernel<false, false><<<grid, 256>>>(a); }
  else if (!strcmp(variant, "alias_v2")) { if (g->has_alias) walk_alias_sm_kernel<true, false><<<grid, 256>>>(a); else walk_alias_sm_kernel<false, false><<<grid, 256>>>(a); }
  else if (!strcmp(variant, "alias_v3")) {
    if (!g->d_meta || !g->d_hash) rc = SRW_ERR_ARG;
    else if (g->has_alias) walk_alias_hash_kernel<true, false><<<grid, 256>>>(a, g->d_meta, g->d_hash);
    else walk_alias_hash_kernel<false, false><<<grid, 256>>>(a, g->d_meta, g->d_hash);
  } else if (!strcmp(variant, "fold_v4")) {
    FoldArgs f{};
    if (!g->d_ent || !g->d_hash || g->ent_ids || g->directed || g->has_alias || !srw_fold_args(p->p, p->q, true, &f)) rc = SRW_ERR_UNSUPPORTED;
    else {
      f.ent = g->d_ent; f.hash = g->d_hash;
      const PeerTable pt{};
      walk_fold_kernel<false, false><<<grid, 256>>>(a, f, pt);
    }
  } else rc = SRW_ERR_ARG;
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(d_stats);
  if (rc != SRW_OK) return rc;
  if (e != cudaSuccess) return SRW_ERR_CUDA;
  return SRW_OK;
}
