"""The migrating-walker super-step kernel SOURCE (csrc/migrate.cuh, mig_step_kernel) compiled for the host with the lockstep
32-lane warp emulator (tests/emu/warp_emu.h) and run for 1..8 emulated GPUs against the CPU twin (oracle_alias_walk).  Checks,
before any GPU time is spent: tuple encode/decode, routing by owner, inbox regions / chunked slot claims / NOP padding, the
spill path (tiny regions), the replicated edge filter + exact symmetric test at owner(x) (tiny filters force the PENDING
round trip), the return-excess step without prev's extent (NEEDEXT), path delivery to the home rows, termination.
The parity tests proper are the `-m gpu` tests (tests/test_gpu_migrate.py), which run the same source on the device.
"""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from conftest import KARATE, ROOT
from test_kernel_emulation import _bounds, _mult, _rmat_twin, _twin_paths

synth = importlib.import_module("stellar-random-walk_b200.synth")
EMU_DIR = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libsrw_emu_migrate.so")
    csrc = os.path.join(ROOT, "stellar-random-walk_b200", "csrc")
    srcs = [os.path.join(EMU_DIR, "emu_migrate.cpp"), os.path.join(EMU_DIR, "warp_emu.h")]
    srcs += [os.path.join(csrc, f) for f in ("migrate.cuh", "walk_conv.cuh", "layout.h", "philox.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-std=c++20", "-DSRW_EMU_WARP",
                               "-pthread", srcs[0], "-o", so])
    lib = C.CDLL(so)
    lib.emu_migrate_walk.restype = C.c_int
    lib.emu_migrate_walk.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                     C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    return lib


def _emu_migrate(emu, oracle, tw, *, walk_length, p, q, seed, fold=True, shards=2, rounds=1, round_first=0, seg_cap=0, bloom_bits=16, blocks=2,
                 owner=None, hub_deg=0):
    v = tw.view()
    off = np.ascontiguousarray(v["offsets"], np.int64)
    col = np.ascontiguousarray(v["col"], np.int32)
    mult = _mult(col, off)
    nv = len(off) - 1
    stride = walk_length + 2
    paths = np.full((rounds * nv, stride), -9, np.int32)
    bounds = _bounds(off, shards)
    t_ret, t_common, t_far = oracle.alias_thresholds(p, q)
    st = np.zeros(8, np.uint64)
    rc = emu.emu_migrate_walk(nv, off.ctypes.data, col.ctypes.data, mult.ctypes.data, shards, bounds.ctypes.data, p, q, int(fold),
                              t_ret, t_common, t_far, seed, walk_length, round_first, rounds, seg_cap, bloom_bits, blocks,
                              paths.ctypes.data, st.ctypes.data, None if owner is None else np.ascontiguousarray(owner, np.uint8).ctypes.data, hub_deg)
    assert rc >= 0, rc
    assert int(st[6]) == 0, "device error flags %d" % int(st[6])
    assert (paths >= 0).all(), "a path slot was never written"
    vids = v["vids"]
    return [vids[paths[i]].tolist() for i in range(rounds * nv)], rc, {k: int(x) for k, x in zip(
        ("super_steps", "steps", "proposals", "tests", "exact_tests", "spills", "err", "tuples"), st)}


@pytest.mark.parametrize("shards", [1, 2, 3, 8])
@pytest.mark.parametrize("p,q", [(0.5, 2.0), (0.25, 4.0), (0.5, 0.5), (1.0, 1.0), (2.0, 0.5)])
def test_migrate_equals_twin_karate(emu, oracle, shards, p, q):
    tw = oracle.AliasGraph(oracle.Graph().load_file(KARATE))
    want, _ = _twin_paths(oracle, tw, walk_length=20, num_walks=2, p=p, q=q, seed=11, fold=1)
    got, folded, st = _emu_migrate(emu, oracle, tw, walk_length=20, p=p, q=q, seed=11, shards=shards, rounds=2)
    assert folded == (1 if 1.0 / p > max(1.0, 1.0 / q) else 0)
    assert got == want
    assert st["steps"] == sum(len(x) - 1 for x in want)


@pytest.mark.parametrize("shards,bloom_bits,seg_cap", [(2, 16, 0), (4, 16, 0), (8, 16, 0), (4, 1, 0), (3, 16, 64), (8, 2, 64)])
def test_migrate_equals_twin_rmat(emu, oracle, shards, bloom_bits, seg_cap):
    """RMAT-8 (multi-edges, self-loops, hubs).  bloom_bits = 1 or 2: a filter that says "maybe" most of the time, so the
    exact test at owner(x), its PENDING tuples and the bounce back run all the time.  seg_cap 64 / 96: regions of two or three
    chunks, so that tuples spill locally and are forwarded one super-step later."""
    tw = _rmat_twin(oracle, 8, 8)
    want, _ = _twin_paths(oracle, tw, walk_length=24, num_walks=2, p=0.5, q=2.0, seed=5, fold=1)
    got, _, st = _emu_migrate(emu, oracle, tw, walk_length=24, p=0.5, q=2.0, seed=5, shards=shards, rounds=2, bloom_bits=bloom_bits,
                              seg_cap=seg_cap, blocks=1)
    assert got == want
    if seg_cap and shards == 3:
        assert st["spills"] > 0
    if bloom_bits <= 2:
        assert st["exact_tests"] > st["tests"] // 4


def test_migrate_round_offset_and_classic_thresholds(emu, oracle):
    """round_first > 0 (walker ids continue across batches) and (p, q) where folding does not apply (classic thresholds, q < 1:
    an adjacent proposal is the one that may be REJECTED, so accepted-at-owner(x) and bounced-back both occur)."""
    tw = _rmat_twin(oracle, 7, 8)
    want, _ = _twin_paths(oracle, tw, walk_length=16, num_walks=5, p=2.0, q=0.5, seed=9, fold=1)
    nv = len(tw.view()["offsets"]) - 1
    got, folded, _ = _emu_migrate(emu, oracle, tw, walk_length=16, p=2.0, q=0.5, seed=9, shards=4, rounds=2, round_first=3, bloom_bits=2, blocks=1)
    assert folded == 0
    assert got == want[3 * nv:5 * nv]


@pytest.mark.parametrize("walk_length", [0, 1, 2])
def test_migrate_short_walks(emu, oracle, walk_length):
    tw = oracle.AliasGraph(oracle.Graph().load_file(KARATE))
    want, _ = _twin_paths(oracle, tw, walk_length=walk_length, num_walks=1, p=0.5, q=2.0, seed=3, fold=1)
    got, _, _ = _emu_migrate(emu, oracle, tw, walk_length=walk_length, p=0.5, q=2.0, seed=3, shards=3)
    assert got == want


# ---- VCut shard map: owner(v) = getPartition(v) mod world from a partition-id column instead of vertex ranges ----
@pytest.mark.parametrize("shards,bloom_bits,seg_cap", [(2, 16, 0), (4, 16, 0), (8, 2, 0), (3, 16, 64), (4, 1, 0)])
def test_migrate_vcut_owner_map_equals_twin(emu, oracle, shards, bloom_bits, seg_cap):
    """An arbitrary vertex -> shard map (interleaved, unbalanced, one shard possibly empty): the rows of a shard are no longer
    a contiguous rank range, extents come from the replicated table, seeds from the shard's vertex list.  Same paths."""
    tw = _rmat_twin(oracle, 8, 8)
    nv = len(tw.view()["offsets"]) - 1
    rng = np.random.default_rng(shards * 7 + bloom_bits)
    owner = rng.integers(0, shards, nv).astype(np.uint8)
    if shards == 4:
        owner[owner == 2] = 0                                     # an empty shard, an overloaded one
    want, _ = _twin_paths(oracle, tw, walk_length=24, num_walks=2, p=0.5, q=2.0, seed=5, fold=1)
    got, _, st = _emu_migrate(emu, oracle, tw, walk_length=24, p=0.5, q=2.0, seed=5, shards=shards, rounds=2, bloom_bits=bloom_bits,
                              seg_cap=seg_cap, blocks=1, owner=owner)
    assert got == want
    assert st["steps"] == sum(len(x) - 1 for x in want)


def test_migrate_vcut_classic_thresholds_karate(emu, oracle):
    tw = oracle.AliasGraph(oracle.Graph().load_file(KARATE))
    nv = len(tw.view()["offsets"]) - 1
    owner = (np.arange(nv) * 5 % 3).astype(np.uint8)
    for p, q in ((2.0, 0.5), (1.0, 1.0), (0.25, 4.0)):
        want, _ = _twin_paths(oracle, tw, walk_length=20, num_walks=2, p=p, q=q, seed=11, fold=1)
        got, _, _ = _emu_migrate(emu, oracle, tw, walk_length=20, p=p, q=q, seed=11, shards=3, rounds=2, owner=owner)
        assert got == want


# ---- replicated hub rows: the highest-degree rows are on every shard, a step onto a hub does not migrate ----
@pytest.mark.parametrize("shards,hub_deg,with_map,bloom_bits", [(2, 40, False, 16), (4, 20, False, 16), (8, 12, False, 2), (3, 30, True, 16), (4, 1, False, 16)])
def test_migrate_replicated_hub_rows_equal_twin(emu, oracle, shards, hub_deg, with_map, bloom_bits):
    """Same paths, fewer tuples.  hub_deg = 1: EVERY row is replicated (nothing migrates after the seeds)."""
    tw = _rmat_twin(oracle, 8, 8)
    nv = len(tw.view()["offsets"]) - 1
    owner = np.random.default_rng(shards).integers(0, shards, nv).astype(np.uint8) if with_map else None
    want, _ = _twin_paths(oracle, tw, walk_length=24, num_walks=2, p=0.5, q=2.0, seed=5, fold=1)
    base, _, st0 = _emu_migrate(emu, oracle, tw, walk_length=24, p=0.5, q=2.0, seed=5, shards=shards, rounds=2, bloom_bits=bloom_bits, blocks=1, owner=owner)
    got, _, st = _emu_migrate(emu, oracle, tw, walk_length=24, p=0.5, q=2.0, seed=5, shards=shards, rounds=2, bloom_bits=bloom_bits, blocks=1, owner=owner,
                              hub_deg=hub_deg)
    assert base == want and got == want
    assert st["steps"] == st0["steps"]
    assert st["tuples"] < st0["tuples"]
    if hub_deg == 1:
        assert st["tuples"] == 0
