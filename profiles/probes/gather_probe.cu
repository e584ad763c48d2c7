// gather_probe.cu -- how many DRAM bytes does one random 4-byte gather cost on B200, per load flavour?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu
// Run:   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum ./gather_probe [granularity]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE>
__device__ __forceinline__ uint32_t load4(const uint32_t *p) {
  uint32_t v;
  if (MODE == 0) v = __ldg(p);
  else if (MODE == 1) asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (MODE == 2) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (MODE == 3) asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (MODE == 5) asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (MODE == 6) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
  else asm volatile("ld.global.nc.L1::evict_first.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <int MODE>
__global__ void gather4(const uint32_t *table, uint64_t n_words, int per_thread, uint32_t *sink) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0, s = mix(t + 1);
  for (int k = 0; k < per_thread; k += 8) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s = mix(s + 0x9E3779B9u * (j + 1));
      const uint64_t i = __umul64hi(((uint64_t)s << 32) | mix(s), n_words);
      v[j] = load4<MODE>(table + i);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j];
  }
  if (acc == 0x12345678u) *sink = acc;
}

// 32-byte aligned sector read as 2 x 16 B
__global__ void gather32(const uint4 *table, uint64_t n_sectors, int per_thread, uint32_t *sink) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0, s = mix(t + 1);
  for (int k = 0; k < per_thread; k += 4) {
    uint4 a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s = mix(s + 0x9E3779B9u * (j + 1));
      const uint64_t i = __umul64hi(((uint64_t)s << 32) | mix(s), n_sectors);
      a[j] = __ldg(table + 2 * i); b[j] = __ldg(table + 2 * i + 1);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) acc += a[j].x ^ b[j].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

// W x 32 bytes from one random (W*32)-byte aligned address: 256-bit loads (LDG.E.256), optionally with the L2::64B hint.
// Does a 64-byte (or 128-byte) gather cost one request slot or several?
template <int W, bool L64>
__global__ void gather_wide(const uint4 *table, uint64_t n_units, int per_thread, uint32_t *sink) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t acc = 0, s = mix(t + 1);
  for (int k = 0; k < per_thread; k += 4) {
    uint32_t a[4][W];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s = mix(s + 0x9E3779B9u * (j + 1));
      const uint64_t i = __umul64hi(((uint64_t)s << 32) | mix(s), n_units);
      const uint4 *p = table + 2 * W * i;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
        if (L64) asm volatile("ld.global.nc.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "l"(p + 2 * w));
        else asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "l"(p + 2 * w));
        a[j][w] = r0 ^ r7;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int w = 0; w < W; ++w) acc += a[j][w];
  }
  if (acc == 0x12345678u) *sink = acc;
}
template <int W, bool L64>
void run_wide(const uint32_t *table, uint64_t bytes, uint32_t *sink, const char *name) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int per_thread = 64, threads = 256, blocks = 148 * 8 * 4;
  gather_wide<W, L64><<<blocks, threads>>>((const uint4 *)table, bytes / (32 * W), per_thread, sink);
  cudaEventRecord(a);
  gather_wide<W, L64><<<blocks, threads>>>((const uint4 *)table, bytes / (32 * W), per_thread, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  printf("%-34s %8.3f ms  %7.2f G gathers/s  err=%s\n", name, ms, (double)blocks * threads * per_thread / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int MODE>
float run(const uint32_t *table, uint64_t n_words, uint32_t *sink, const char *name) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int per_thread = 64, threads = 256, blocks = 148 * 8 * 4;
  gather4<MODE><<<blocks, threads>>>(table, n_words, per_thread, sink);
  cudaEventRecord(a);
  gather4<MODE><<<blocks, threads>>>(table, n_words, per_thread, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double n = (double)blocks * threads * per_thread;
  printf("%-34s %8.3f ms  %7.2f G gathers/s  err=%s\n", name, ms, n / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  return ms;
}

int main(int argc, char **argv) {
  const size_t gran = argc > 1 ? (size_t)atoi(argv[1]) : 0;
  const size_t gb = argc > 2 ? (size_t)atoi(argv[2]) : 8;
  if (gran) printf("cudaLimitMaxL2FetchGranularity <- %zu: %s\n", gran, cudaGetErrorString(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran)));
  size_t got = 0;
  cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
  printf("L2 fetch granularity limit = %zu, table = %zu GiB\n", got, gb);
  const uint64_t bytes = gb << 30;
  uint32_t *table, *sink;
  cudaMalloc(&table, bytes);
  cudaMemset(table, 1, bytes);
  cudaMalloc(&sink, 4);
  run<0>(table, bytes / 4, sink, "__ldg (ld.global.nc)");
  run<1>(table, bytes / 4, sink, "ld.global.ca");
  run<2>(table, bytes / 4, sink, "ld.global.cg");
  run<3>(table, bytes / 4, sink, "ld.global.cs");
  run<4>(table, bytes / 4, sink, "ld.global.nc.L1::no_allocate");
  run<5>(table, bytes / 4, sink, "ld.global.nc.L2::64B");
  run<6>(table, bytes / 4, sink, "ld.global.cv");
  run<7>(table, bytes / 4, sink, "ld.global.nc.L1::evict_first");
  {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int per_thread = 64, threads = 256, blocks = 148 * 8 * 4;
    gather32<<<blocks, threads>>>((const uint4 *)table, bytes / 32, per_thread, sink);
    cudaEventRecord(a);
    gather32<<<blocks, threads>>>((const uint4 *)table, bytes / 32, per_thread, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("%-34s %8.3f ms  %7.2f G gathers/s\n", "32B sector (2 x ldg.128)", ms, (double)blocks * threads * per_thread / ms / 1e6);
  }
  run_wide<1, false>(table, bytes, sink, "32B (1 x ldg.256)");
  run_wide<1, true>(table, bytes, sink, "32B (1 x ldg.256.L2::64B)");
  run_wide<2, false>(table, bytes, sink, "64B aligned (2 x ldg.256)");
  run_wide<2, true>(table, bytes, sink, "64B aligned (2 x ldg.256.L2::64B)");
  run_wide<4, false>(table, bytes, sink, "128B line (4 x ldg.256)");
  return 0;
}
