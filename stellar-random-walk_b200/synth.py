"""Deterministic synthetic edge lists -- numpy twins of the device generators
in csrc/synth.cu.  Used by tests and bench.py only; not on the walk path.

RMAT: edge e draws one Philox4x32-10 word per level, key = (gen_seed, 0), counter =
(e_lo, e_hi, level // 4, 0x524d4154), word = level % 4; quadrant by integer thresholds of
(a, b, c, d) = (0.57, 0.19, 0.19, 0.05).  No dedup, self-loops kept (Graph500 style, no permutation).
"""
import numpy as np

_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 over uint64-held 32-bit lanes."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & _MASK for x in (c0, c1, c2, c3))
    k0 = np.uint64(k0 & 0xFFFFFFFF)
    k1 = np.uint64(k1 & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_M0) * c0
        p1 = np.uint64(_M1) * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & _MASK
        n1 = p1 & _MASK
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & _MASK
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return c0, c1, c2, c3


RMAT_TAG = 0x524D4154
WEIGHT_TAG = 0x57454947
RMAT_A = int(0.57 * 2 ** 32)
RMAT_AB = int((0.57 + 0.19) * 2 ** 32)
RMAT_ABC = int((0.57 + 0.19 + 0.19) * 2 ** 32)


def rmat_edges(scale, edge_factor=16, seed=42, first=0, count=None):
    """Edges [first, first+count) of the RMAT instance; returns (src, dst) int32."""
    n_edges = edge_factor << scale
    if count is None:
        count = n_edges - first
    e = np.arange(first, first + count, dtype=np.uint64)
    lo, hi = e & _MASK, e >> np.uint64(32)
    src = np.zeros(count, dtype=np.uint64)
    dst = np.zeros(count, dtype=np.uint64)
    for blk in range((scale + 3) // 4):
        words = philox4x32_10(lo, hi, np.full(count, blk, np.uint64), np.full(count, RMAT_TAG, np.uint64), seed, 0)
        for k in range(4):
            level = blk * 4 + k
            if level >= scale:
                break
            r = words[k]
            sbit = (r >= np.uint64(RMAT_AB)).astype(np.uint64)
            dbit = (((r >= np.uint64(RMAT_A)) & (r < np.uint64(RMAT_AB))) | (r >= np.uint64(RMAT_ABC))).astype(np.uint64)
            src = (src << np.uint64(1)) | sbit
            dst = (dst << np.uint64(1)) | dbit
    return src.astype(np.int32), dst.astype(np.int32)


def edge_weights(n_edges, seed=43, first=0):
    """float32 weight 1 + (philox(seed, e) mod 1000) / 1000 per input edge (config C3)."""
    e = np.arange(first, first + n_edges, dtype=np.uint64)
    r0, _, _, _ = philox4x32_10(e & _MASK, e >> np.uint64(32), np.zeros(n_edges, np.uint64),
                                np.full(n_edges, WEIGHT_TAG, np.uint64), seed, 0)
    return (np.float32(1.0) + (r0 % np.uint64(1000)).astype(np.float32) / np.float32(1000.0)).astype(np.float32)


ZIPF_TAG = 0x5A495046
ZIPF_DST_TAG = 0x5A445354


def zipf_stub_counts(n_vertices, cap=1000000, seed=7):
    """Out-stub count per vertex: min(cap, floor(1 / (1 - U))), U = r / 2^32 -> integer 2^32 // (2^32 - r)
    (P(count >= k) = 1/k: Zipf exponent 2 on the pmf; config C5)."""
    v = np.arange(n_vertices, dtype=np.uint64)
    r, _, _, _ = philox4x32_10(v & _MASK, v >> np.uint64(32), np.zeros(n_vertices, np.uint64),
                               np.full(n_vertices, ZIPF_TAG, np.uint64), seed, 0)
    cnt = np.uint64(1 << 32) // (np.uint64(1 << 32) - r)
    return np.minimum(cnt, np.uint64(cap)).astype(np.int64)


def zipf_edges(n_vertices, cap=1000000, seed=7):
    """Edge e of vertex v's stub block: (v, uniform target) -- self-loops and duplicates kept."""
    cnt = zipf_stub_counts(n_vertices, cap, seed)
    n_edges = int(cnt.sum())
    src = np.repeat(np.arange(n_vertices, dtype=np.int32), cnt)
    e = np.arange(n_edges, dtype=np.uint64)
    r, _, _, _ = philox4x32_10(e & _MASK, e >> np.uint64(32), np.zeros(n_edges, np.uint64),
                               np.full(n_edges, ZIPF_DST_TAG, np.uint64), seed, 0)
    dst = ((r * np.uint64(n_vertices)) >> np.uint64(32)).astype(np.int32)
    return src, dst


def edges_to_text(src, dst, w=None, pid=None):
    cols = [src, dst]
    if pid is not None:
        cols.append(pid)
    lines = []
    for i in range(len(src)):
        parts = [str(int(c[i])) for c in cols]
        if w is not None:
            parts.append(repr(float(np.float32(w[i]))))
        lines.append(" ".join(parts))
    return "\n".join(lines) + "\n"
