"""The warp-cooperative exact-sampler kernels (csrc/walk_exact.cuh) compiled for the host and run under a lockstep
32-lane warp emulator (tests/emu/warp_emu.h: one thread per lane, barrier-backed shuffles and ballots), against the
ORACLE -- the restated reference algorithm (RS:12-62, RW:51-133).  Bit-exact, as on the device; small graphs only
(the emulator spends microseconds per collective)."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from conftest import KARATE, ROOT

synth = importlib.import_module("stellar-random-walk_b200.synth")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libsrw_emu_exact.so")


@pytest.fixture(scope="module")
def emu():
    csrc = os.path.join(ROOT, "stellar-random-walk_b200", "csrc")
    srcs = [os.path.join(EMU_DIR, "emu_exact.cpp"), os.path.join(EMU_DIR, "warp_emu.h")] + [os.path.join(csrc, f) for f in ("walk_exact.cuh", "walk_conv.cuh", "layout.h", "philox.cuh")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++20", "-ffp-contract=off", "-Wno-unknown-pragmas", "-pthread", "-shared", "-fPIC", srcs[0], "-o", EMU_SO])
    lib = C.CDLL(EMU_SO)
    lib.emu_exact_walk.restype = C.c_int
    lib.emu_exact_walk.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_uint64, C.c_float,
                                   C.c_int32, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return lib


def _layout(og):
    """Appearance-order rows in rank space + the same rows sorted (what graph_build.cu produces with SRW_BUILD_ALL)."""
    vids = og.vertex_ids()
    rank = {int(v): i for i, v in enumerate(vids.tolist())}
    off = [0]
    col_app, w_app, col_sorted = [], [], []
    for v in vids.tolist():
        nb = og.neighbors(int(v)) or []
        r = [rank[d] for d, _ in nb]
        col_app += r
        w_app += [w for _, w in nb]
        col_sorted += sorted(r)
        off.append(len(col_app))
    return (vids, np.array(off, np.int64), np.array(col_app, np.int32), np.array(w_app, np.float32), np.array(col_sorted, np.int32))


def _emu_exact(emu, og, kernel, *, walk_length, p, q, seed, u_const=None, first=0, n=None, use_hash=True):
    vids, off, col_app, w_app, col_sorted = _layout(og)
    nv = len(vids)
    n = nv if n is None else n
    stride = walk_length + 2
    paths = np.full((n, stride), -7, np.int32)
    lens = np.zeros(n, np.int32)
    st = np.zeros(4, np.uint64)
    rc = emu.emu_exact_walk(kernel, nv, off.ctypes.data, col_app.ctypes.data, w_app.ctypes.data, col_sorted.ctypes.data, p, q, seed,
                            -1.0 if u_const is None else u_const, walk_length, first, n, paths.ctypes.data, lens.ctypes.data, int(use_hash), st.ctypes.data)
    assert rc == 0
    for i in range(n):
        assert (paths[i, lens[i]:] == -7).all()
    return [vids[paths[i, :lens[i]]].tolist() for i in range(n)], st


KERNELS = [0, 1, 2]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("directed,u", [(False, 0.1), (False, 0.9), (True, 0.9)])
def test_emulated_exact_kernels_constant_u_karate(emu, oracle, kernel, directed, u):
    """The reference's own walk scenarios (T-URW:181-291: constant generator) through the warp kernels."""
    og = oracle.Graph().load_file(KARATE, directed=directed)
    ids, offs = oracle.walk(og, walk_length=12, num_walks=1, u_const=u)
    got, _ = _emu_exact(emu, og, kernel, walk_length=12, p=1.0, q=1.0, seed=1, u_const=u)
    assert got == oracle.paths_as_lists(ids, offs)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("weighted,p,q", [(False, 0.5, 2.0), (True, 0.25, 4.0), (True, 2.0, 0.5)])
def test_emulated_exact_kernels_equal_oracle(emu, oracle, kernel, weighted, p, q):
    s, d = synth.rmat_edges(7, 8, seed=42)
    w = synth.edge_weights(len(s), seed=43) if weighted else None
    og = oracle.Graph().load_edges(s, d, w)
    ids, offs = oracle.walk(og, walk_length=8, num_walks=1, p=p, q=q, seed=7)
    want = oracle.paths_as_lists(ids, offs)
    n = 48
    got, _ = _emu_exact(emu, og, kernel, walk_length=8, p=p, q=q, seed=7, n=n)
    assert got == want[:n]
    if kernel == 2:
        got, _ = _emu_exact(emu, og, kernel, walk_length=8, p=p, q=q, seed=7, n=16, use_hash=False)
        assert got == want[:16]


def _hub_graph(oracle, n_leaves, weighted, seed=3):
    """Two hubs (ids 0, 1) and leaves: hub 0 - every leaf, hub 1 - every third leaf, some leaf-leaf edges, a hub-hub edge,
    a few parallel edges.  Rows of >= 2048 entries take cert2's two-level path; hub steps coming from a leaf build the
    common list."""
    rng = np.random.RandomState(seed)
    src, dst = [0], [1]
    for i in range(n_leaves):
        leaf = 2 + i
        src.append(0); dst.append(leaf)
        if i % 3 == 0:
            src.append(leaf); dst.append(1)
        if i % 5 == 0 and i + 1 < n_leaves:
            src.append(leaf); dst.append(leaf + 1)
        if i % 97 == 0:
            src.append(0); dst.append(leaf)          # parallel edge
    s, d = np.array(src, np.int32), np.array(dst, np.int32)
    w = (0.25 + rng.randint(0, 16, len(s)) / 4.0).astype(np.float32) if weighted else None
    return oracle.Graph().load_edges(s, d, w)


@pytest.mark.parametrize("weighted,p,q", [(False, 0.5, 2.0), (True, 0.25, 4.0), (True, 2.0, 0.5)])
def test_emulated_cert2_long_rows_equal_oracle(emu, oracle, weighted, p, q):
    og = _hub_graph(oracle, 2600, weighted)
    ids, offs = oracle.walk(og, walk_length=6, num_walks=1, p=p, q=q, seed=11, threads=0)
    want = oracle.paths_as_lists(ids, offs)
    n = 24                                            # start vertices: the two hubs and the first leaves
    for kernel in (1, 2):
        got, st = _emu_exact(emu, og, kernel, walk_length=6, p=p, q=q, seed=11, n=n)
        assert got == want[:n], kernel
    assert any(0 in pth[1:] for pth in got)           # the walks did pass through the long row


def test_emulated_cert2_group_boundary_replays(emu, oracle):
    """A star with 2048 unit-weight leaves: 32 groups of 64 entries; u = k/32 puts a group-end prefix exactly on u, inside
    the +-delta band, so the step is replayed in order and RS:20 `acc >= u` picks that very entry."""
    n_leaves = 2048
    s = np.zeros(n_leaves, np.int32)
    d = np.arange(1, n_leaves + 1, dtype=np.int32)
    og = oracle.Graph().load_edges(s, d)
    for u in (0.25, 0.5, 0.03125, 0.999):
        ids, offs = oracle.walk(og, walk_length=3, num_walks=1, u_const=u)
        want = oracle.paths_as_lists(ids, offs)
        got, st = _emu_exact(emu, og, 2, walk_length=3, p=1.0, q=1.0, seed=1, u_const=u, n=8)
        assert got == want[:8], u
        if u != 0.999:
            assert int(st[2]) > 0                     # in-order replays did happen
