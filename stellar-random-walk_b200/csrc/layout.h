// HBM record layouts shared by the build kernels, the walk kernels and the host-emulated kernel tests
// (tests/emu): plain structs and placement functions, no CUDA runtime dependency.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SRW_LAYOUT_HD __host__ __device__ inline
#define SRW_ALIGN(n) __align__(n)
#else
#define SRW_LAYOUT_HD inline
#define SRW_ALIGN(n) __attribute__((aligned(n)))
#endif

// Vose slot, one 16-byte gather per proposal: own and alias target vertex are both inline.
struct SRW_ALIGN(16) AliasSlot {
  uint32_t thr;           // take `own` iff r < thr
  int32_t own;            // neighbour rank stored at this slot
  int32_t alias_vertex;   // neighbour rank of the alias slot
  uint32_t alias_index;   // row-relative index of the alias slot
};

// Vose slot of the WEIGHTED alias-fold sampler (undirected weighted graphs): the 16 bytes above plus, for whichever
// neighbour the coin picks, the total weight of the parallel edges to it ("bundle weight", double sum in row order).  The
// walker carries it to the next step, where it is the mass of the return edge (symmetric on an undirected graph).
struct SRW_ALIGN(32) AliasSlotW {
  uint32_t thr;
  int32_t own;
  int32_t alias_vertex;
  uint32_t alias_index;
  double wb_own;
  double wb_alias;
};

// Row descriptor read once per walk step.  Rows with more than kHashMinDeg neighbours also own a hash
// set of their neighbour ids (`nb` buckets of 8 slots starting at bucket `hoff`): the d(t,x)=1
// membership test of node2vec becomes ~1 sector instead of a ~log2(deg) binary search.
struct SRW_ALIGN(32) RowMeta {
  int64_t off;     // first entry in d_col / d_slot
  int64_t hoff;    // first bucket in d_hash
  uint32_t deg;
  uint32_t nb;     // 0: no hash set (short row: binary search in d_col)
  double w_sum;    // weighted graphs: sequential double sum of the row's weights (the W of the Vose build); else deg
};
// Hash-set placement is DERIVED from the row extent, so nothing but (off, deg) is needed to probe it:
// row r owns buckets [off >> 2, (off + deg) >> 2) -- about deg/4 buckets of 8 slots (load <= ~0.5).
// Rows shorter than kHashMinDeg have no set (8 * nb >= 2*deg - 6 >= deg + 1 needs deg >= 7).
constexpr uint32_t kHashMinDeg = 8;
SRW_LAYOUT_HD int64_t srw_hash_first(int64_t off) { return off >> 2; }
SRW_LAYOUT_HD uint32_t srw_hash_buckets(int64_t off, uint32_t deg) {
  return deg >= kHashMinDeg ? (uint32_t)(((off + (int64_t)deg) >> 2) - (off >> 2)) : 0u;
}
// Neighbour entry of the fold sampler (unweighted graphs): one 16-byte gather yields the neighbour AND
// its row extent AND the multiplicity of the edge, so a walk step needs no separate row-descriptor load.
struct SRW_ALIGN(16) NbrEntry {
  int32_t x;             // neighbour rank
  uint32_t deg;          // deg(x)
  uint32_t off_lo;       // row offset of x inside its owner's arrays (< 2^32: the ABI caps nnz at 2^32 - 1)
  uint32_t off_hi_mult;  // [7:0] owner shard of x (0 on an unsharded graph), [31:8] number of parallel edges to x in this row
};
SRW_LAYOUT_HD uint32_t srw_hash32(uint32_t x) {
  x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12;
  return x;
}

#define SRW_MAX_SHARDS 16

// VCut shard map (VRW:23-26,121-134, GM:31,66-68): owner(v) = getPartition(v) mod world comes from the partition-id
// column of the edge file, so a shard's rows are not a contiguous rank range; this table -- replicated on every shard, 8 bytes
// per vertex -- says where the row of ANY vertex lives inside its owner's arrays.
struct SRW_ALIGN(8) MigExt { uint32_t off, deg; };

// The replicated edge filter of the migrating sharded walk (migrate.cuh): one 64-bit Bloom word per probe, kMigBloomK bits
// per undirected edge {a, b} of vertex ranks -- two in each 32-bit half of the word, all from 32-bit arithmetic (the probe sits
// on the walk kernel's instruction-bound path).  The word index and the bit positions come from two INDEPENDENT 32-bit hashes
// of the pair: deriving both from one 32-bit value would make every pair that collides with an inserted edge in those 32 bits a
// false positive -- 1e9 edges cover a quarter of 2^32 (measured: 6 % false positives at RMAT-24 with a single hash, 0.4 % with two).
// 16 bits per edge: ~0.3 % false positives.
constexpr int kMigBloomK = 4;
// word index (< n_words < 2^32) and bit mask of the pair (order-free)
SRW_LAYOUT_HD void srw_bloom_probe(int32_t a, int32_t b, uint32_t n_words, uint32_t *word, uint64_t *mask) {
  const uint32_t lo = (uint32_t)(a < b ? a : b), hi = (uint32_t)(a < b ? b : a);
  const uint32_t h = srw_hash32(lo ^ srw_hash32(hi + 0x9E3779B9u));
  *word = (uint32_t)(((uint64_t)h * (uint64_t)n_words) >> 32);
  const uint32_t g = srw_hash32((hi * 0x85EBCA6Bu) ^ srw_hash32(lo + 0x68E31DA4u));
  const uint32_t m_lo = (1u << (g & 31u)) | (1u << ((g >> 5) & 31u)), m_hi = (1u << ((g >> 10) & 31u)) | (1u << ((g >> 15) & 31u));
  *mask = ((uint64_t)m_hi << 32) | (uint64_t)m_lo;
}
