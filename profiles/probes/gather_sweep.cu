// gather_sweep.cu -- WHERE does the random-gather ceiling of the walk kernel live?
//
// The walk kernel's dominant access is a data-dependent 16-byte (neighbour entry) or 32-byte (hash bucket) gather with
// at most ONE load in flight per lane.  This probe measures the achievable rate of exactly that access as a function of
//   (1) the table footprint: 32 MiB (L2-resident) ... 64 GiB (the walk's footprint class) -- a rate that is flat over the
//       footprint lives on the SM side (L1TEX -> XBAR request path, occupancy x latency); a rate that drops once the
//       table leaves L2 and then stays flat lives in DRAM (row activations); a second drop beyond the TLB reach would be
//       address translation;
//   (2) the concurrency: independent loads (8 in flight per thread, 2048 threads/SM) vs a DEPENDENT chain (the next index
//       is a hash of the loaded word: one load in flight per thread, like a walker) at 1024 and 2048 threads per SM.
// Every point runs >= ~50 ms.  Plain run: CUDA-event rates.  Under ncu (profiles/scripts/r2_call_a.sh): L2 hit rate,
// L2 requests / sectors, XBAR request-cycle utilisation, DRAM bytes and throughput per point.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_sweep gather_sweep.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void fill_kernel(uint4 *t, uint64_t n16) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t a = mix((uint32_t)i), b = mix((uint32_t)(i >> 32) + a);
    t[i] = make_uint4(a, b, a ^ b, a + b);
  }
}

__device__ __forceinline__ uint4 ld16(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// independent: 8 gathers in flight per thread
__global__ void __launch_bounds__(256, 8) gather_indep(const uint4 *table, uint64_t n16, int per_thread, uint32_t *sink) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0, s = mix(t + 1);
  for (int k = 0; k < per_thread; k += 8) {
    uint4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s = mix(s + 0x9E3779B9u * (j + 1));
      v[j] = ld16(table + __umul64hi(((uint64_t)s << 32) | mix(s), n16));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j].x ^ v[j].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

// dependent chain: the next index is a hash of the word just loaded -- one load in flight per thread (a walker).
// MINB = resident blocks per SM the launch bounds allow (4 -> 1024 threads/SM like the walk kernel, 8 -> 2048).
template <int MINB>
__global__ void __launch_bounds__(256, MINB) gather_chain(const uint4 *table, uint64_t n16, int per_thread, uint32_t *sink) {
  __shared__ uint32_t pad[MINB == 4 ? 12 * 1024 : 1];   // 48 KB per block at MINB = 4: caps residency at 4 blocks per SM
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t s = mix(t + 1);
  if (MINB == 4 && per_thread < 0) pad[threadIdx.x] = s;
  for (int k = 0; k < per_thread; ++k) {
    const uint4 v = ld16(table + __umul64hi(((uint64_t)s << 32) | mix(s), n16));
    s = mix(s ^ v.x) + v.w + (uint32_t)k;
  }
  if (s == 0x12345678u) *sink = s + (MINB == 4 ? pad[0] : 0);
}

template <class F>
static void timed(const char *name, double gathers, size_t mib, F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch();                       // warm-up (also what ncu's -s skips are counted against: 2 launches per point)
  cudaEventRecord(a);
  launch();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  printf("{\"probe\": \"%s\", \"table_mib\": %zu, \"ms\": %.3f, \"gathers_per_s\": %.4e, \"err\": \"%s\"}\n", name, mib, ms,
         gathers / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
  fflush(stdout);
  cudaEventDestroy(a); cudaEventDestroy(b);
}

int main(int argc, char **argv) {
  size_t max_gib = argc > 1 ? (size_t)atoi(argv[1]) : 64;
  const double scale = argc > 2 ? atof(argv[2]) : 1.0;     // < 1: shorter points (ncu replays)
  const size_t only_mib = argc > 3 ? (size_t)atoll(argv[3]) : 0;   // one table size only (bench.py: the ceiling beside the walk)
  uint4 *table;
  uint32_t *sink;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  while ((max_gib << 30) + (2ull << 30) > free_b && max_gib > 1) max_gib >>= 1;
  if (cudaMalloc(&table, max_gib << 30) != cudaSuccess) { printf("cudaMalloc of %zu GiB failed\n", max_gib); return 1; }
  cudaMalloc(&sink, 4);
  fill_kernel<<<148 * 16, 256>>>(table, (max_gib << 30) / 16);
  cudaDeviceSynchronize();
  const size_t sizes_mib[] = {32, 128, 512, 2048, 8192, 32768, 65536};
  for (size_t mib : sizes_mib) {
    if (mib > (max_gib << 10)) break;
    if (only_mib && mib != only_mib) continue;
    const uint64_t n16 = (mib << 20) / 16;
    {
      const int blocks = 148 * 8 * 4, per_thread = (int)(3072 * scale) & ~7;          // 1.2 M threads x 3072 = 3.7 G gathers
      timed("indep8_2048thr", (double)blocks * 256 * per_thread, mib, [&] { gather_indep<<<blocks, 256>>>(table, n16, per_thread, sink); });
    }
    {
      const int blocks = 148 * 8, per_thread = (int)(8192 * scale);                   // persistent: one wave, 2048 threads/SM
      timed("chain_2048thr", (double)blocks * 256 * per_thread, mib, [&] { gather_chain<8><<<blocks, 256>>>(table, n16, per_thread, sink); });
    }
    {
      const int blocks = 148 * 4, per_thread = (int)(12288 * scale);                  // one wave, 1024 threads/SM (the walk kernel's residency)
      timed("chain_1024thr", (double)blocks * 256 * per_thread, mib, [&] { gather_chain<4><<<blocks, 256>>>(table, n16, per_thread, sink); });
    }
  }
  return 0;
}
