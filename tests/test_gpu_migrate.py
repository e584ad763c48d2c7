"""SURVEY 8(e) gate for the migrating-walker sharded walk (csrc/migrate.cuh, srw_mig_*): the emitted paths are exactly the
CPU twin's / the single-GPU kernel's for any number of shards.  All W shards live on one device here (peer pointers are plain
pointers, shards run one after another inside a super-step); the same kernel over real NVLink peer memory with the NCCL
all-reduce barrier runs in test_two_ranks_nccl (self-launched torchrun, needs 2 GPUs) and in bench.py --gpus N."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
srw = importlib.import_module("stellar-random-walk_b200")
synth = importlib.import_module("stellar-random-walk_b200.synth")


def _shards(s, d, world):
    import torch
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    ds, dd = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
    return sh, [sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None, r, world, migrate=True) for r in range(world)]


def _assemble(shards, out, nv, rounds, stride):
    rows = np.full((rounds * nv, stride), -1, np.int32)
    for x, (paths, lens) in zip(shards, out):
        P = paths.cpu().numpy()
        assert (lens.cpu().numpy() == stride).all()
        for rnd in range(rounds):
            rows[rnd * nv + x.rank:(rnd + 1) * nv:x.world] = P[rnd * x.home_rows:(rnd + 1) * x.home_rows]
    return rows


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("sampler,p,q", [("fold", 0.5, 2.0), ("fold", 0.25, 4.0), ("alias", 0.5, 2.0), ("fold", 2.0, 0.5), ("alias", 1.0, 1.0)])
def test_migrate_equals_twin(oracle, world, sampler, p, q):
    sh, shards = _shards(*synth.rmat_edges(10, 8, seed=42), world)
    s, d = synth.rmat_edges(10, 8, seed=42)
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d, None))
    ids, offs, st = twin.walk(walk_length=30, num_walks=3, p=p, q=q, seed=9, fold=1 if sampler == "fold" else 0)
    mw = sh.MigrateWalker(shards, srw.Params(walkLength=30, numWalks=3, p=p, q=q, seed=9, sampler=sampler), 3, stats=True)
    out, stats = mw.run(0)
    rows = _assemble(shards, out, twin.nv, 3, 32)
    assert (np.diff(offs) == 32).all()
    assert (rows.reshape(-1) == ids).all()
    assert stats["steps"] == st.steps
    if world > 1:
        assert stats["tuples_sent_all_ranks"] > 0
    mw.free()


@pytest.mark.parametrize("world,seg_cap,bloom_bits", [(4, 256, 16), (8, 128, 16), (4, 0, 1), (8, 128, 2)])
def test_migrate_spill_and_weak_filter(oracle, monkeypatch, world, seg_cap, bloom_bits):
    """Regions of a few chunks (tuples spill locally and are forwarded a super-step later) and a 1-2 bit/edge filter (most
    tests go to the exact check at owner(x): PENDING tuples, bounces).  Same paths."""
    monkeypatch.setenv("SRW_BLOOM_BITS", str(bloom_bits))
    monkeypatch.setenv("SRW_MIG_BLOCKS", "8")
    s, d = synth.rmat_edges(11, 8, seed=5)
    sh, shards = _shards(s, d, world)
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d, None))
    ids, offs, st = twin.walk(walk_length=40, num_walks=2, p=0.5, q=2.0, seed=21, fold=1)
    mw = sh.MigrateWalker(shards, srw.Params(walkLength=40, numWalks=2, p=0.5, q=2.0, seed=21, sampler="fold"), 2, seg_cap=seg_cap, stats=True)
    for rep in range(2):                                   # the context is reusable
        out, stats = mw.run(0)
        rows = _assemble(shards, out, twin.nv, 2, 42)
        assert (rows.reshape(-1) == ids).all()
        assert stats["steps"] == st.steps
        if seg_cap:
            assert stats["spills"] > 0
        if bloom_bits <= 2:
            assert stats["exact_tests"] > stats["filter_probes"] // 4
    mw.free()


def test_migrate_matches_single_gpu_kernel_and_batches():
    """RMAT-14, 4 shards, rounds walked as two batches (round_first 0 and 2) == srw_walk_device on the unsharded graph."""
    import torch
    s, d = synth.rmat_edges(14, 8, seed=7)
    sh, shards = _shards(s, d, 4)
    ds, dd = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
    g = srw.Graph.from_device_edges(len(s), ds.data_ptr(), dd.data_ptr(), None, False, srw.BUILD_ALIAS)
    prm = srw.Params(walkLength=80, numWalks=4, p=0.5, q=2.0, seed=5, sampler="fold")
    ref_ids, ref_offs = g.walk(prm).arrays()
    nv = g.num_vertices
    mw = sh.MigrateWalker(shards, prm, 2)
    got = []
    for first in (0, 2):
        out, stats = mw.run(first)
        got.append(_assemble(shards, out, nv, 2, 82).copy())
        assert stats["steps"] == 2 * nv * 81
        assert stats["spills"] == 0            # the default regions hold the whole per-pair flow of a super-step
    assert (np.concatenate(got).reshape(-1) == ref_ids).all()
    mw.free()


def test_migrate_refuses_what_it_cannot_do():
    import torch
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    s, d = synth.rmat_edges(8, 4, seed=1)
    ds, dd = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
    plain = [sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None, r, 2) for r in range(2)]           # no edge filter
    with pytest.raises(srw.SrwError):
        sh.MigrateWalker(plain, srw.Params(walkLength=10, numWalks=1), 1)
    directed = [sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None, r, 2, directed=True, migrate=True) for r in range(2)]
    with pytest.raises(srw.SrwError):
        sh.MigrateWalker(directed, srw.Params(walkLength=10, numWalks=1), 1)


# ---- the VCut shard map -- owner(v) = getPartition(v) mod world from the partition-id column (VRW:23-26,121-134) ----
def _vcut_shards(s, d, pid, world):
    import torch
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    ds, dd, dp = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(pid).cuda()
    return sh, [sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None, r, world, migrate=True, d_pid=dp.data_ptr()) for r in range(world)]


@pytest.mark.parametrize("world,pid_kind", [(2, "random"), (4, "random"), (4, "skewed"), (3, "wide"), (8, "random")])
def test_vcut_shard_map_equals_twin(oracle, world, pid_kind):
    """Shards follow the partition-id column: every vertex lives on getPartition(v) mod world (the pid of the LAST input line in
    which it is a neighbour, GM:31), rows are not rank ranges, walkers are routed by that map -- and the paths are the twin's."""
    s, d = synth.rmat_edges(11, 8, seed=42)
    rng = np.random.default_rng(world)
    if pid_kind == "random":
        pid = rng.integers(0, world, len(s)).astype(np.int32)
    elif pid_kind == "skewed":                       # one partition gets most edges, one none
        pid = np.where(rng.random(len(s)) < 0.7, 0, rng.integers(2, world, len(s))).astype(np.int32)
    else:                                            # partition ids beyond the GPU count: HashPartitioner's pid mod numPartitions
        pid = rng.integers(0, 40, len(s)).astype(np.int32)
    og = oracle.Graph().load_edges(s, d, None, pid=pid)
    twin = oracle.AliasGraph(og)
    ids, offs, st = twin.walk(walk_length=30, num_walks=3, p=0.5, q=2.0, seed=9, fold=1)
    sh, shards = _vcut_shards(s, d, pid, world)
    # the shard map is getPartition: every vertex's row is on exactly that shard, complete
    vids = twin.view()["vids"]
    off = twin.view()["offsets"]
    probe = rng.choice(len(vids), 64, replace=False)
    L = srw.lib()
    for r in probe:
        v = int(vids[r])
        want_owner = og.partition(v) % world
        for x in shards:
            n = srw.C.c_int64()
            srw.check(L.srw_graph_neighbors(x.h, v, None, None, 0, srw.C.byref(n)))
            assert n.value == (int(off[r + 1] - off[r]) if x.rank == want_owner else -1), (v, x.rank, want_owner)
    assert sum(x.row_last - x.row_first for x in shards) == twin.nv
    assert sum(x.nnz_local for x in shards) == int(off[-1])
    mw = sh.MigrateWalker(shards, srw.Params(walkLength=30, numWalks=3, p=0.5, q=2.0, seed=9, sampler="fold"), 3, stats=True)
    out, stats = mw.run(0)
    rows = _assemble(shards, out, twin.nv, 3, 32)
    assert (rows.reshape(-1) == ids).all()
    assert stats["steps"] == st.steps
    mw.free()
    # the other sharded modes refuse such shards instead of walking them with range arithmetic
    with pytest.raises(srw.SrwError):
        shards[0].attach_local(shards)


# ---- replicated hub rows (vertex-cut): the highest-degree rows are kept by every shard, a step onto a hub does not migrate ----
@pytest.mark.parametrize("world,hub_fraction,with_pid", [(2, 0.5, False), (4, 0.5, False), (8, 0.75, False), (4, 0.25, True), (3, 0.95, False)])
def test_replicated_hub_rows_equal_twin(oracle, monkeypatch, world, hub_fraction, with_pid):
    import torch
    monkeypatch.setenv("SRW_MIG_BLOCKS", "8")            # few warps: the slot counts below are tuples, not the NOP padding of open chunks
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    s, d = synth.rmat_edges(12, 8, seed=42)
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d, None))
    ids, offs, st = twin.walk(walk_length=30, num_walks=3, p=0.5, q=2.0, seed=9, fold=1)
    ds, dd = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
    dp = torch.from_numpy(np.random.default_rng(1).integers(0, world, len(s)).astype(np.int32)).cuda() if with_pid else None
    prm = srw.Params(walkLength=30, numWalks=3, p=0.5, q=2.0, seed=9, sampler="fold")
    tuples = {}
    for hf in (0.0, hub_fraction):
        shards = [sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None, r, world, migrate=True, hub_fraction=hf,
                           d_pid=None if dp is None else dp.data_ptr()) for r in range(world)]
        if hf > 0:
            nnz = int(twin.view()["offsets"][-1])
            deg = np.diff(twin.view()["offsets"])
            x = shards[0]
            assert 0 < x.hub_entries <= hf * nnz
            assert x.hub_entries == int(deg[deg >= x.hub_min_degree].sum()) and x.hub_rows == int((deg >= x.hub_min_degree).sum())
            assert all((y.hub_rows, y.hub_entries, y.hub_min_degree) == (x.hub_rows, x.hub_entries, x.hub_min_degree) for y in shards)
            assert sum(y.nnz_local for y in shards) == nnz + (world - 1) * x.hub_entries      # every shard holds the hub rows
            assert sum(y.seed_rows for y in shards) == twin.nv
            # a hub's row answers on every shard (GM:109-120 through the replicated tables)
            hub_v = int(twin.view()["vids"][int(np.argmax(deg))])
            for y in shards:
                n = srw.C.c_int64()
                srw.check(srw.lib().srw_graph_neighbors(y.h, hub_v, None, None, 0, srw.C.byref(n)))
                assert n.value == int(deg.max())
        mw = sh.MigrateWalker(shards, prm, 3, stats=True)
        out, stats = mw.run(0)
        rows = _assemble(shards, out, twin.nv, 3, 32)
        assert (rows.reshape(-1) == ids).all()
        assert stats["steps"] == st.steps
        tuples[hf] = stats["tuples_sent_all_ranks"]
        mw.free()
        for y in shards:
            y.free()
    assert tuples[hub_fraction] < tuples[0.0]            # roughly (1 - f) of the migrations remain (plus chunk padding)


def test_two_ranks_nccl():
    """tests/dist_sharded_check.py under torchrun with one rank per GPU: the NCCL tuple exchange, peer-gather over symmetric
    memory and the migrating walk over real peer memory, each against the CPU twin.  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the driver's multi-GPU tier runs it)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29571", os.path.join(ROOT, "tests", "dist_sharded_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SHARDED_DIST_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


def test_one_process_two_gpus_through_the_abi(oracle, tmp_path):
    """`--gpus 2` behind the ordinary entry points (csrc/multi.cu): srw_graph_from_edges_multi / srw_graph_load build one shard
    per device in ONE handle, srw_walk and srw_walk_save (the CLI) walk it with the migrating-walker kernel over peer memory, CUDA
    events as the barrier.  Same paths and the same output files as one GPU.  Needs two GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the driver's multi-GPU tier runs it)")
    s, d = synth.rmat_edges(12, 8, seed=42)
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d, None))
    ids, offs, st = twin.walk(walk_length=30, num_walks=3, p=0.5, q=2.0, seed=9, fold=1)
    g = srw.Graph.from_edges_multi(s, d, 2)
    assert g.stats() == (twin.nv, int(twin.view()["offsets"][-1]))
    v0 = int(twin.view()["vids"][0])
    assert len(g.neighbors(v0)) == int(np.diff(twin.view()["offsets"])[0])
    got_ids, got_offs = g.walk(srw.Params(walkLength=30, numWalks=3, p=0.5, q=2.0, seed=9, sampler="fold", gpus=2)).arrays()
    assert (got_offs == offs).all() and (got_ids == ids).all()
    with pytest.raises(srw.SrwError):          # weighted / directed graphs are refused, not walked wrongly
        srw.Graph.from_edges_multi(s, d, 2, directed=True)
    g.free()
    # the CLI: same files with --gpus 2 as with one GPU
    inp = tmp_path / "edges.txt"
    inp.write_text(synth.edges_to_text(s[:20000], d[:20000]))
    outs = []
    for gpus in (1, 2):
        out = tmp_path / ("out%d" % gpus)
        rc = srw.Main.main(["--cmd", "randomwalk", "--input", str(inp), "--output", str(out), "--walkLength", "20", "--numWalks", "2",
                            "--p", "0.5", "--q", "2.0", "--weighted", "false", "--seed", "4", "--gpus", str(gpus)])
        assert rc == 0
        outs.append((out / "path" / "part-00000").read_bytes())
    assert outs[0] == outs[1] and len(outs[0]) > 0
    # VCutRandomWalk on two GPUs: `--partitioned true --gpus 2` shards by the partition-id column (owner(v) = getPartition(v) mod 2)
    pid = np.random.default_rng(3).integers(0, 7, 20000).astype(np.int32)
    inp2 = tmp_path / "edges_pid.txt"
    inp2.write_text("".join("%d %d %d\n" % (a, b, c) for a, b, c in zip(s[:20000], d[:20000], pid)))
    out = tmp_path / "out_vcut"
    rc = srw.Main.main(["--cmd", "randomwalk", "--input", str(inp2), "--output", str(out), "--walkLength", "20", "--numWalks", "2",
                        "--p", "0.5", "--q", "2.0", "--weighted", "false", "--seed", "4", "--gpus", "2", "--partitioned", "true"])
    assert rc == 0
    assert (out / "path" / "part-00000").read_bytes() == outs[0]
    gv = srw.Graph.from_edges_multi_vcut(s, d, np.random.default_rng(4).integers(0, 5, len(s)).astype(np.int32), 2)
    got_ids, got_offs = gv.walk(srw.Params(walkLength=30, numWalks=3, p=0.5, q=2.0, seed=9, sampler="fold", gpus=2)).arrays()
    assert (got_offs == offs).all() and (got_ids == ids).all()
    assert gv.partition(v0) is not None
    gv.free()
