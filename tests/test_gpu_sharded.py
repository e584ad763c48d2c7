"""SURVEY 8(e) gate: the vertex-range-sharded walk emits exactly the single-GPU / CPU-twin paths for
any number of shards.  All W shards are built on one device and exchanged in-process (LocalExchange);
the NCCL exchange itself is covered by tests/dist_sharded_check.py under torchrun (2 GPUs)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
srw = importlib.import_module("stellar-random-walk_b200")
synth = importlib.import_module("stellar-random-walk_b200.synth")


def _device_edges(s, d, w):
    import torch
    ds = torch.from_numpy(s).cuda()
    dd = torch.from_numpy(d).cuda()
    dw = torch.from_numpy(w).cuda() if w is not None else None
    return ds, dd, dw


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("weighted,directed,p,q", [(False, False, 0.5, 2.0), (True, False, 0.25, 4.0), (True, True, 2.0, 0.5)])
def test_sharded_equals_twin(oracle, world, weighted, directed, p, q):
    import torch
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    s, d = synth.rmat_edges(9, 8, seed=42)
    w = synth.edge_weights(len(s), seed=43) if weighted else None
    ds, dd, dw = _device_edges(s, d, w)
    shards = [sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None if dw is None else dw.data_ptr(), r, world, directed) for r in range(world)]
    # shard plan: contiguous, covering, edge-balanced, identical on every shard
    og = oracle.Graph().load_edges(s, d, w, directed=directed)
    twin = oracle.AliasGraph(og)
    tv = twin.view()
    assert all(x.bounds == shards[0].bounds for x in shards)
    assert shards[0].bounds == sh.plan_bounds(tv["offsets"], world)
    assert sum(x.nnz_local for x in shards) == int(tv["offsets"][-1])
    prm = srw.Params(walkLength=25, numWalks=3, p=p, q=q, seed=17, sampler="alias")
    out, stats = sh.run_sharded(shards, prm, 0, 3, rec_cap=1 << 12 if world == 3 else 1 << 20)   # small cap: exercises parking
    ids, offs, st = twin.walk(walk_length=25, num_walks=3, p=p, q=q, seed=17)
    want = oracle.paths_as_lists(ids, offs)
    nv = twin.nv
    got = [None] * (3 * nv)
    for x, (paths, lens) in zip(shards, out):
        P, Ln = paths.cpu().numpy(), lens.cpu().numpy()
        for rnd in range(3):
            for k, v in enumerate(x.home_vertices()):
                row = rnd * x.home_rows + k
                got[rnd * nv + v] = P[row, :Ln[row]].tolist()
    assert got == want
    assert stats["steps"] == st.steps
    if world > 1:
        assert stats["tuples_sent"] > 0 and stats["records_sent"] > 0


def test_sharded_matches_single_gpu_kernel():
    """Same graph, same seed: 4 shards == srw_walk_device on the unsharded graph (RMAT-12)."""
    import torch
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    s, d = synth.rmat_edges(12, 8, seed=7)
    ds, dd, _ = _device_edges(s, d, None)
    g = srw.Graph.from_device_edges(len(s), ds.data_ptr(), dd.data_ptr())
    prm = srw.Params(walkLength=40, numWalks=2, p=0.5, q=2.0, seed=5, sampler="alias")     # the tuple-exchange shards run the classic sampler
    ref_ids, ref_offs = g.walk(prm).arrays()
    shards = [sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None, r, 4) for r in range(4)]
    out, _ = sh.run_sharded(shards, prm, 0, 2)
    nv = g.num_vertices
    rows = np.full((2 * nv, 42), -1, np.int32)
    for x, (paths, lens) in zip(shards, out):
        P = paths.cpu().numpy()
        for rnd in range(2):
            rows[rnd * nv + x.rank: (rnd + 1) * nv: x.world] = P[rnd * x.home_rows:(rnd + 1) * x.home_rows]
    assert (np.diff(ref_offs) == 42).all()
    assert (rows.reshape(-1) == ref_ids).all()


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("sampler,p,q", [("fold", 0.5, 2.0), ("fold", 0.25, 4.0), ("alias", 0.5, 2.0), ("fold", 2.0, 0.5), ("alias", 1.0, 1.0)])
def test_peer_gather_equals_twin(oracle, world, sampler, p, q):
    """Peer-gather mode: W shards on one device, every shard attached to every other; each shard handle walks
    a slice of the walkers against the whole graph.  Output == CPU twin (== the unsharded kernel)."""
    import torch
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    s, d = synth.rmat_edges(10, 8, seed=42)
    ds, dd, _ = _device_edges(s, d, None)
    shards = [sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None, r, world) for r in range(world)]
    for x in shards:
        # handle-to-handle pointers, or (the path one process per GPU takes) rows relocated into per-shard blocks
        x.attach_blocks_local(shards) if (world == 3 or sampler == "alias") else x.attach_local(shards)
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d, None))
    ids, offs, st = twin.walk(walk_length=30, num_walks=2, p=p, q=q, seed=9, fold=1 if sampler == "fold" else 0)
    nv = twin.nv
    prm = srw.Params(walkLength=30, numWalks=2, p=p, q=q, seed=9, sampler=sampler)
    total = 2 * nv
    paths = torch.full((total, 32), -7, dtype=torch.int32, device="cuda")
    lens = torch.zeros(total, dtype=torch.int32, device="cuda")
    steps = 0
    for r, x in enumerate(shards):             # walker slices deliberately not aligned with the vertex ranges
        lo, hi = total * r // world, total * (r + 1) // world
        wi = x.walk_device(prm, lo, hi - lo, paths[lo:].data_ptr(), lens[lo:].data_ptr())
        steps += wi.steps
    P, Ln = paths.cpu().numpy(), lens.cpu().numpy()
    got = [P[i, :Ln[i]].tolist() for i in range(total)]
    assert got == oracle.paths_as_lists(ids, offs)
    assert steps == st.steps


def test_peer_gather_requires_every_peer():
    import torch
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    s, d = synth.rmat_edges(8, 4, seed=1)
    ds, dd, _ = _device_edges(s, d, None)
    shards = [sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None, r, 2) for r in range(2)]
    paths = torch.zeros((4, 12), dtype=torch.int32, device="cuda")
    lens = torch.zeros(4, dtype=torch.int32, device="cuda")
    with pytest.raises(srw.SrwError):
        shards[0].walk_device(srw.Params(walkLength=10, numWalks=1), 0, 4, paths.data_ptr(), lens.data_ptr())
    # weighted shards carry no neighbour entries: attaching is refused
    w = synth.edge_weights(len(s), seed=2)
    dw = torch.from_numpy(w).cuda()
    ws = [sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), dw.data_ptr(), r, 2) for r in range(2)]
    with pytest.raises(srw.SrwError):
        ws[0].attach_local(ws)
