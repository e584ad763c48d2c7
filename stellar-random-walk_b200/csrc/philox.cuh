// Philox4x32-10 counter RNG (Salmon et al., SC'11), host + device.  Replaces the reference's
// JVM-global scala.util.Random.nextFloat (RW:9,52,76; RS:5): a draw is a pure function of
// (seed; walker, step, trial), so results do not depend on scheduling, GPU count or arrival order.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SRW_HD __host__ __device__ __forceinline__
#else
#define SRW_HD inline
#endif

struct Philox4 {
  uint32_t x, y, z, w;
};

SRW_HD uint32_t srw_mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

SRW_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = srw_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = srw_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 o = {c0, c1, c2, c3};
  return o;
}

// the walker stream: key = seed, counter = (walker_lo, walker_hi, step, trial)
SRW_HD Philox4 walker_rng(uint32_t seed_lo, uint32_t seed_hi, uint64_t walker, uint32_t step, uint32_t trial) {
  return philox4x32_10((uint32_t)walker, (uint32_t)(walker >> 32), step, trial, seed_lo, seed_hi);
}

// a float on the same 2^-24 grid as java.util.Random.nextFloat (next(24) / 2^24)
SRW_HD float u01_from_bits(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
