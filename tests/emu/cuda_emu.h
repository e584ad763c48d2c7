// cuda_emu.h -- TEST INFRASTRUCTURE.  The handful of CUDA built-ins the walk kernels in
// stellar-random-walk_b200/csrc/walk_conv.cuh use, restated for a host compile (g++), so that the kernel
// SOURCE can be executed lane by lane on the CPU and compared with the CPU twin before any GPU time is
// spent.  A "warp" here is one lane; __any_sync keeps a finished lane in its loop for a few more iterations
// (as a finished lane on the device stays with its warp), which checks that a DONE lane is inert.
// Never linked into libsrw.so; only tests/ builds it.
#pragma once
#include <stdint.h>
#include <string.h>

#define SRW_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)

struct emu_dim3 { unsigned x, y, z; };
static emu_dim3 threadIdx, blockIdx, blockDim;
static int emu_extra_iters = 0;       // iterations a finished lane is still driven through the loop
static int emu_linger = 0;

struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 v = {x, y, z, w}; return v; }

template <class T> static inline T __ldg(const T *p) { return *p; }
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
// IEEE double, one rounding per operation (the file is compiled with -ffp-contract=off)
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static inline bool __any_sync(unsigned, bool pred) {
  if (pred) { emu_linger = emu_extra_iters; return true; }
  return emu_linger-- > 0;
}
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint64_t)(uint32_t)lo;
  double d;
  memcpy(&d, &u, 8);
  return d;
}
