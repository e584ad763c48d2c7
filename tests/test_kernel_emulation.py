"""The walk kernel SOURCE (csrc/walk_conv.cuh, walk_fold_conv_kernel) compiled for the host with
tests/emu/cuda_emu.h and run lane by lane against the CPU twin (oracle_alias_walk).  This is a CPU-side check
of the kernel's decision logic -- row layout, hash placement, thresholds, the draw/load/consume phases, the
staged path stores, the vertex-range (PEER) addressing -- so that a logic slip is caught before GPU time is
spent; the parity tests proper are the `-m gpu` tests, which run the same source on the device.
"""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from conftest import KARATE, ROOT

synth = importlib.import_module("stellar-random-walk_b200.synth")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libsrw_emu.so")


@pytest.fixture(scope="module", params=["lane", "warp"])
def emu(request):
    """lane: one lane at a time (cuda_emu.h).  warp: the lockstep 32-lane warp emulator (warp_emu.h), where the lanes of a
    warp finish at different times -- the `__any_sync` loop and the inert-finished-lane logic run for real."""
    warp = request.param == "warp"
    so = os.path.join(EMU_DIR, "libsrw_emu_warp.so" if warp else "libsrw_emu.so")
    srcs = [os.path.join(EMU_DIR, "emu_walk.cpp"), os.path.join(EMU_DIR, "warp_emu.h" if warp else "cuda_emu.h")]
    csrc = os.path.join(ROOT, "stellar-random-walk_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in ("walk_conv.cuh", "layout.h", "philox.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        flags = ["-std=c++20", "-DSRW_EMU_WARP", "-pthread"] if warp else ["-std=c++17"]
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC"] + flags + [srcs[0], "-o", so])
    EMU_SO = so
    lib = C.CDLL(EMU_SO)
    lib.emu_fold_walk.restype = C.c_int
    lib.emu_fold_walk.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                  C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.emu_fold_walk_ids.restype = C.c_int
    lib.emu_fold_walk_ids.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_uint64, C.c_int32,
                                      C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.emu_alias_walk.restype = C.c_int
    lib.emu_alias_walk.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                   C.c_int32, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.emu_wfold_walk.restype = C.c_int
    lib.emu_wfold_walk.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                   C.c_uint64, C.c_int32, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return lib


def _mult(col, offsets):
    """Number of parallel edges to the same neighbour, per entry of the sorted rows (graph_build.cu k_nbr_entries)."""
    m = np.ones(len(col), dtype=np.uint32)
    for r in range(len(offsets) - 1):
        lo, hi = int(offsets[r]), int(offsets[r + 1])
        if hi - lo > 1:
            _, inv, cnt = np.unique(col[lo:hi], return_inverse=True, return_counts=True)
            m[lo:hi] = cnt[inv]
    return m


def _bounds(offsets, shards):
    """Edge-balanced vertex ranges on the degree prefix sum (sharded.py / graph_build.cu rule is not needed here:
    any monotone cut must give the same paths)."""
    nv, nnz = len(offsets) - 1, int(offsets[-1])
    b = [0]
    for s in range(1, shards):
        b.append(int(np.searchsorted(offsets, nnz * s // shards, side="left")))
    b.append(nv)
    return np.maximum.accumulate(np.array(b, dtype=np.int64))


def _emu_walk(emu, oracle, tw, *, walk_length, p, q, seed, fold, shards=1, var=0, extra=2, first=0, n=None, rounds=1, stats=False):
    v = tw.view()
    off = np.ascontiguousarray(v["offsets"], np.int64)
    col = np.ascontiguousarray(v["col"], np.int32)
    mult = _mult(col, off)
    nv = len(off) - 1
    n = nv * rounds - first if n is None else n
    stride = walk_length + 2
    paths = np.full((n, stride), -7, np.int32)
    lens = np.zeros(n, np.int32)
    bounds = _bounds(off, shards)
    t_ret, t_common, t_far = oracle.alias_thresholds(p, q)
    st = np.zeros(4, np.uint64)
    rc = emu.emu_fold_walk(nv, off.ctypes.data, col.ctypes.data, mult.ctypes.data, shards, bounds.ctypes.data, p, q, int(fold),
                           t_ret, t_common, t_far, seed, walk_length, first, n, paths.ctypes.data, lens.ctypes.data, var, extra,
                           int(stats), st.ctypes.data)
    assert rc >= 0
    vids = v["vids"]
    out = [vids[paths[i, :lens[i]]].tolist() for i in range(n)]
    # nothing may be written past a path's end
    for i in range(n):
        assert (paths[i, lens[i]:] == -7).all()
    return out, rc, st


def _twin_paths(oracle, tw, **kw):
    ids, offs, st = tw.walk(**kw)
    return oracle.paths_as_lists(ids, offs), st


def _rmat_twin(oracle, scale, ef, seed=42):
    s, d = synth.rmat_edges(scale, ef, seed=seed)
    return oracle.AliasGraph(oracle.Graph().load_edges(s, d))


@pytest.mark.parametrize("p,q", [(0.5, 2.0), (0.25, 4.0), (0.5, 0.5), (1.0, 1.0), (2.0, 0.5)])
def test_emulated_kernel_equals_twin_karate(emu, oracle, p, q):
    tw = oracle.AliasGraph(oracle.Graph().load_file(KARATE))
    want, _ = _twin_paths(oracle, tw, walk_length=30, num_walks=3, p=p, q=q, seed=11, fold=1)
    got, folded, _ = _emu_walk(emu, oracle, tw, walk_length=30, p=p, q=q, seed=11, fold=True, rounds=3)
    assert folded == (1 if 1.0 / p > max(1.0, 1.0 / q) else 0)
    assert got == want


@pytest.mark.parametrize("walk_length", [0, 1, 2, 13, 14, 15, 16, 17, 80])
def test_emulated_kernel_path_staging_boundaries(emu, oracle, walk_length):
    """Stage of 16 ids, first flush cut at a 16-byte boundary: every stride parity and every tail length."""
    tw = _rmat_twin(oracle, 8, 8)
    want, _ = _twin_paths(oracle, tw, walk_length=walk_length, num_walks=2, p=0.5, q=2.0, seed=5, fold=1)
    got, _, _ = _emu_walk(emu, oracle, tw, walk_length=walk_length, p=0.5, q=2.0, seed=5, fold=True, rounds=2)
    assert got == want


@pytest.mark.parametrize("shards", [1, 2, 3, 8])
@pytest.mark.parametrize("fold", [True, False])
def test_emulated_kernel_equals_twin_rmat_sharded(emu, oracle, shards, fold):
    """PEER addressing (shard-local offsets, owner byte in the entry) must not change a single id; fold=False
    drives the classic thresholds through the same kernel (what a peer-gather walk does when 1/p <= max(1, 1/q))."""
    tw = _rmat_twin(oracle, 10, 8)
    p, q = (0.5, 2.0) if fold else (2.0, 0.5)
    want, wst = _twin_paths(oracle, tw, walk_length=40, num_walks=1, p=p, q=q, seed=9, fold=1)
    got, _, st = _emu_walk(emu, oracle, tw, walk_length=40, p=p, q=q, seed=9, fold=True, shards=shards, stats=True)
    assert got == want
    if fold:   # (classic thresholds: the kernel takes a degree-1 row's only entry at once, the twin counts its rejected trials)
        assert int(st[1]) == wst.proposals and int(st[2]) == wst.member_tests and int(st[3]) == wst.probes_log2


def test_emulated_kernel_walker_window_and_lingering_lanes(emu, oracle):
    """A launch over a window of walkers (walker_first, n not a multiple of the block) gives the same paths as the
    full launch; finished lanes driven through extra iterations stay inert; the L2::64B load flavour is the same code."""
    tw = _rmat_twin(oracle, 9, 8)
    want, _ = _twin_paths(oracle, tw, walk_length=20, num_walks=2, p=0.5, q=2.0, seed=3, fold=1)
    nv = tw.nv
    first, n = nv // 3, nv + 77
    for extra, var in ((0, 0), (5, 0), (3, 1)):
        got, _, _ = _emu_walk(emu, oracle, tw, walk_length=20, p=0.5, q=2.0, seed=3, fold=True, first=first, n=n, extra=extra, var=var)
        assert got == want[first:first + n]


def test_emulated_kernel_hub_and_multi_edges(emu, oracle):
    """Zipf hubs (long rows -> hash probes with full buckets), parallel edges (multiplicity in the return component)
    and self-loops."""
    zs, zd = synth.zipf_edges(2048, seed=7, cap=600)
    extra_s = np.array([0, 0, 0, 5, 5, 9], np.int32)
    extra_d = np.array([1, 1, 1, 5, 6, 9], np.int32)
    s, d = np.concatenate([zs, extra_s]), np.concatenate([zd, extra_d])
    tw = oracle.AliasGraph(oracle.Graph().load_edges(s, d))
    for p, q in ((0.25, 4.0), (0.5, 2.0)):
        want, _ = _twin_paths(oracle, tw, walk_length=25, num_walks=1, p=p, q=q, seed=21, fold=1)
        got, _, _ = _emu_walk(emu, oracle, tw, walk_length=25, p=p, q=q, seed=21, fold=True, shards=2)
        assert got == want


# ---- the classic alias sampler in the convergent layout (walk_alias_conv_kernel) ----
def _emu_alias_walk(emu, oracle, tw, *, walk_length, p, q, seed, var=0, extra=2, rounds=1):
    v = tw.view()
    off = np.ascontiguousarray(v["offsets"], np.int64)
    col = np.ascontiguousarray(v["col"], np.int32)
    thr = np.ascontiguousarray(v["thr"], np.uint32) if "thr" in v else None
    al = np.ascontiguousarray(v["alias"], np.uint32) if "alias" in v else None
    nv = len(off) - 1
    n = nv * rounds
    stride = walk_length + 2
    paths = np.full((n, stride), -7, np.int32)
    lens = np.zeros(n, np.int32)
    t_ret, t_common, t_far = oracle.alias_thresholds(p, q)
    st = np.zeros(4, np.uint64)
    rc = emu.emu_alias_walk(nv, off.ctypes.data, col.ctypes.data, None if thr is None else thr.ctypes.data,
                            None if al is None else al.ctypes.data, t_ret, t_common, t_far, seed, walk_length, 0, n,
                            paths.ctypes.data, lens.ctypes.data, var, extra, st.ctypes.data)
    assert rc == 0
    for i in range(n):
        assert (paths[i, lens[i]:] == -7).all()
    vids = v["vids"]
    return [vids[paths[i, :lens[i]]].tolist() for i in range(n)], st


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("directed", [False, True])
@pytest.mark.parametrize("p,q", [(0.5, 2.0), (2.0, 0.5), (1.0, 1.0), (0.25, 4.0)])
def test_emulated_alias_kernel_equals_twin(emu, oracle, weighted, directed, p, q):
    s, d = synth.rmat_edges(9, 8, seed=42)
    w = synth.edge_weights(len(s), seed=43) if weighted else None
    tw = oracle.AliasGraph(oracle.Graph().load_edges(s, d, w, directed=directed), directed=directed)
    assert tw.has_alias == weighted
    for wl in (30, 13):
        want, wst = _twin_paths(oracle, tw, walk_length=wl, num_walks=2, p=p, q=q, seed=17)
        for var, extra in ((0, 2), (1, 0)):
            got, st = _emu_alias_walk(emu, oracle, tw, walk_length=wl, p=p, q=q, seed=17, var=var, extra=extra, rounds=2)
            assert got == want


def test_emulated_alias_kernel_hubs(emu, oracle):
    zs, zd = synth.zipf_edges(2048, seed=7, cap=600)
    w = synth.edge_weights(len(zs), seed=5)
    for ww in (None, w):
        tw = oracle.AliasGraph(oracle.Graph().load_edges(zs, zd, ww))
        want, _ = _twin_paths(oracle, tw, walk_length=25, num_walks=1, p=0.25, q=4.0, seed=21)
        got, _ = _emu_alias_walk(emu, oracle, tw, walk_length=25, p=0.25, q=4.0, seed=21)
        assert got == want


# ---- the weighted alias-fold sampler (walk_wfold_conv_kernel) ----
def _emu_wfold_walk(emu, oracle, tw, *, walk_length, p, q, seed, var=0, extra=2, rounds=1):
    v = tw.view()
    off = np.ascontiguousarray(v["offsets"], np.int64)
    col = np.ascontiguousarray(v["col"], np.int32)
    thr = np.ascontiguousarray(v["thr"], np.uint32)
    al = np.ascontiguousarray(v["alias"], np.uint32)
    nv, nnz = len(off) - 1, int(off[-1])
    L = oracle.lib()
    L.oa_wbundle.restype = C.POINTER(C.c_double)
    L.oa_wsum.restype = C.POINTER(C.c_double)
    wb = np.ctypeslib.as_array(L.oa_wbundle(tw.h), (nnz,)).copy()
    ws = np.ctypeslib.as_array(L.oa_wsum(tw.h), (nv,)).copy()
    n = nv * rounds
    stride = walk_length + 2
    paths = np.full((n, stride), -7, np.int32)
    lens = np.zeros(n, np.int32)
    st = np.zeros(4, np.uint64)
    rc = emu.emu_wfold_walk(nv, off.ctypes.data, col.ctypes.data, thr.ctypes.data, al.ctypes.data, ws.ctypes.data, wb.ctypes.data,
                            p, q, seed, walk_length, 0, n, paths.ctypes.data, lens.ctypes.data, var, extra, st.ctypes.data)
    assert rc == 0
    for i in range(n):
        assert (paths[i, lens[i]:] == -7).all()
    vids = v["vids"]
    return [vids[paths[i, :lens[i]]].tolist() for i in range(n)], st


@pytest.mark.parametrize("p,q", [(0.5, 2.0), (0.25, 4.0), (0.1, 0.5), (0.5, 1.0)])
def test_emulated_weighted_fold_kernel_equals_twin(emu, oracle, p, q):
    s, d = synth.rmat_edges(9, 8, seed=42)
    w = synth.edge_weights(len(s), seed=43)
    tw = oracle.AliasGraph(oracle.Graph().load_edges(s, d, w))
    assert tw.has_alias
    for wl in (30, 13):
        want, wst = _twin_paths(oracle, tw, walk_length=wl, num_walks=2, p=p, q=q, seed=17, fold=1)
        classic, cst = _twin_paths(oracle, tw, walk_length=wl, num_walks=2, p=p, q=q, seed=17, fold=0)
        assert want != classic and wst.proposals < cst.proposals          # the fold really is a different, cheaper sampler
        for var, extra in ((0, 2), (1, 0)):
            got, st = _emu_wfold_walk(emu, oracle, tw, walk_length=wl, p=p, q=q, seed=17, var=var, extra=extra, rounds=2)
            assert got == want
            if var == 0:
                assert int(st[1]) == wst.proposals and int(st[2]) == wst.member_tests


def test_emulated_weighted_fold_kernel_hubs_and_bundles(emu, oracle):
    """Zipf hubs plus parallel edges of different weights (bundle weight != any single weight) and self-loops."""
    zs, zd = synth.zipf_edges(2048, seed=7, cap=600)
    w = synth.edge_weights(len(zs), seed=5)
    es = np.array([0, 0, 0, 5, 5, 9], np.int32)
    ed = np.array([1, 1, 1, 5, 6, 9], np.int32)
    ew = np.array([0.25, 1.5, 3.0, 2.0, 0.125, 7.0], np.float32)
    tw = oracle.AliasGraph(oracle.Graph().load_edges(np.concatenate([zs, es]), np.concatenate([zd, ed]), np.concatenate([w, ew])))
    want, _ = _twin_paths(oracle, tw, walk_length=25, num_walks=1, p=0.25, q=4.0, seed=21, fold=1)
    got, _ = _emu_wfold_walk(emu, oracle, tw, walk_length=25, p=0.25, q=4.0, seed=21)
    assert got == want


@pytest.mark.parametrize("walk_length", [0, 1, 2, 14, 15, 16, 17])
def test_emulated_weighted_and_classic_kernels_short_walks(emu, oracle, walk_length):
    """Stride boundaries of the staged stores and the first-order-only walk (walkLength 0) for the two kernels that share
    the flush code with the fold kernel."""
    s, d = synth.rmat_edges(8, 8, seed=42)
    w = synth.edge_weights(len(s), seed=43)
    tw = oracle.AliasGraph(oracle.Graph().load_edges(s, d, w))
    want, _ = _twin_paths(oracle, tw, walk_length=walk_length, num_walks=2, p=0.5, q=2.0, seed=5, fold=1)
    got, _ = _emu_wfold_walk(emu, oracle, tw, walk_length=walk_length, p=0.5, q=2.0, seed=5, rounds=2)
    assert got == want
    want, _ = _twin_paths(oracle, tw, walk_length=walk_length, num_walks=2, p=0.5, q=2.0, seed=5, fold=0)
    got, _ = _emu_alias_walk(emu, oracle, tw, walk_length=walk_length, p=0.5, q=2.0, seed=5, rounds=2)
    assert got == want
    td = oracle.AliasGraph(oracle.Graph().load_edges(s, d, None, directed=True), directed=True)     # dead ends: ragged paths
    want, _ = _twin_paths(oracle, td, walk_length=walk_length, num_walks=2, p=0.5, q=2.0, seed=5, fold=1)
    got, _ = _emu_alias_walk(emu, oracle, td, walk_length=walk_length, p=0.5, q=2.0, seed=5, rounds=2)
    assert got == want
    if walk_length > 0:
        assert len({len(x) for x in got}) > 1


# ---- id-space fold (SRW_FOLD_IDS): entries and hash sets carry original ids, the kernel emits ids ----
@pytest.mark.parametrize("p,q", [(0.5, 2.0), (0.25, 4.0)])
def test_emulated_fold_kernel_in_id_space(emu, oracle, p, q):
    s, d = synth.rmat_edges(10, 8, seed=42)
    s, d = (s * 7 + 3).astype(np.int32), (d * 7 + 3).astype(np.int32)          # sparse, non-dense ids: rank != id everywhere
    tw = oracle.AliasGraph(oracle.Graph().load_edges(s, d))
    v = tw.view()
    assert not np.array_equal(v["vids"], np.arange(tw.nv))
    off = np.ascontiguousarray(v["offsets"], np.int64)
    col = np.ascontiguousarray(v["col"], np.int32)
    vids = np.ascontiguousarray(v["vids"], np.int32)
    mult = _mult(col, off)
    for wl in (40, 13):
        want, _ = _twin_paths(oracle, tw, walk_length=wl, num_walks=2, p=p, q=q, seed=9, fold=1)
        n = 2 * tw.nv
        paths = np.full((n, wl + 2), -7, np.int32)
        lens = np.zeros(n, np.int32)
        rc = emu.emu_fold_walk_ids(tw.nv, off.ctypes.data, col.ctypes.data, mult.ctypes.data, vids.ctypes.data, p, q, 9, wl, 0, n,
                                   paths.ctypes.data, lens.ctypes.data, 0, 2)
        assert rc == 0
        assert [paths[i, :lens[i]].tolist() for i in range(n)] == want
