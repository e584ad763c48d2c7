// warp_emu.h -- TEST INFRASTRUCTURE.  A lockstep 32-lane warp for running the product's WARP-COOPERATIVE kernels
// (csrc/walk_exact.cuh: shuffles, ballots) on the host: one std::thread per lane, every full-mask collective is
// two phases of a 32-party barrier around a shared slot array.  Blocks are 256 threads = 8 warps that never talk to
// each other in these kernels, so warps are emulated one after another.  Slow (microseconds per collective) and
// only meant for the small graphs of the CPU test-suite; the parity tests proper run the same source on the device.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <barrier>
#include <functional>
#include <thread>
#include <vector>

#define SRW_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __launch_bounds__(...)

struct emu_dim3 { unsigned x, y, z; };
static thread_local emu_dim3 threadIdx, blockIdx;
static emu_dim3 blockDim = {256, 1, 1};

struct EmuWarp {
  std::barrier<> bar{32};
  uint64_t slot[32];
};
static EmuWarp *g_warp = nullptr;
static thread_local int t_lane = 0;

struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 v = {x, y, z, w}; return v; }

template <class T> static inline uint64_t emu_bits(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, "slot"); memcpy(&u, &v, sizeof(T)); return u; }
template <class T> static inline T emu_from(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

// every collective: publish, wait, read, wait (so that the slots can be reused by the next collective)
template <class T, class F> static inline T emu_collective(T v, F pick) {
  g_warp->slot[t_lane] = emu_bits(v);
  g_warp->bar.arrive_and_wait();
  const T r = pick();
  g_warp->bar.arrive_and_wait();
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_collective(v, [&] { return emu_from<T>(g_warp->slot[src & 31]); }); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emu_collective(v, [&] { return t_lane >= d ? emu_from<T>(g_warp->slot[t_lane - d]) : v; }); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_collective(v, [&] { return t_lane + d < 32 ? emu_from<T>(g_warp->slot[t_lane + d]) : v; }); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_collective(v, [&] { return emu_from<T>(g_warp->slot[(t_lane ^ m) & 31]); }); }
static inline unsigned __ballot_sync(unsigned, bool p) {
  return emu_collective<uint32_t>(p ? 1u : 0u, [&] { unsigned m = 0; for (int l = 0; l < 32; ++l) m |= (unsigned)(g_warp->slot[l] & 1u) << l; return m; });
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline void __syncwarp() { g_warp->bar.arrive_and_wait(); }
static inline void __syncthreads() { g_warp->bar.arrive_and_wait(); }   // warps of a block run one after another: a warp barrier orders what matters
static inline int __ffs(unsigned x) { return x ? __builtin_ctz(x) + 1 : 0; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }

template <class T> static inline T __ldg(const T *p) { return *p; }
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
// IEEE, one rounding per operation (compiled with -ffp-contract=off; SSE2 float/double arithmetic)
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
using std::min;
using std::max;
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// Runs `kernel()` for n_warps warps of a grid of 256-thread blocks: lane threads are created once and walk the warps together.
static inline void emu_launch_warps(int64_t n_warps, const std::function<void()> &kernel) {
  EmuWarp warp;
  g_warp = &warp;
  std::vector<std::thread> lanes;
  for (int l = 0; l < 32; ++l)
    lanes.emplace_back([&, l] {
      t_lane = l;
      for (int64_t w = 0; w < n_warps; ++w) {
        blockIdx.x = (unsigned)(w / 8); blockIdx.y = blockIdx.z = 0;
        threadIdx.x = (unsigned)((w % 8) * 32 + l); threadIdx.y = threadIdx.z = 0;
        kernel();
        warp.bar.arrive_and_wait();
      }
    });
  for (auto &t : lanes) t.join();
  g_warp = nullptr;
}
static inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint64_t)(uint32_t)lo;
  return emu_from<double>(u);
}
