"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

P0  reference KATs on the device functions (T-RS, T-GM, karate counts, constant-u walks)
P1  exact sampler  == oracle, bit for bit (Philox u), several (p, q, weighted, directed, seed)
P2  alias sampler  == CPU twin, bit for bit; graph layout (sorted rows, Vose slots) == twin's
"""
import importlib
import os

import numpy as np
import pytest

from conftest import KARATE, TESTGRAPH

pytestmark = pytest.mark.gpu
srw = importlib.import_module("stellar-random-walk_b200")
synth = importlib.import_module("stellar-random-walk_b200.synth")


# ---- P0: RandomSampleTest.scala on the device ------------------------------------------------
def test_kat_random_sample():
    edges = [(1, 1.0), (2, 1.0), (3, 1.0)]
    for u, e in ((0.1, edges[0]), (0.4, edges[1]), (0.7, edges[2])):
        assert srw.RandomSample(lambda: u).sample(edges) == e


def test_kat_second_order():
    w1 = 1.0
    e12, e21, e23, e24, e14, e15 = (2, w1), (1, w1), (3, w1), (4, w1), (4, w1), (5, w1)
    prev, prev_n, curr_n = 1, [e12, e14, e15], [e21, e23, e24]
    rs = srw.RandomSample()
    assert rs.computeSecondOrderWeights(1.0, 1.0, prev, prev_n, curr_n) == curr_n
    for u, e in ((0.1, e21), (0.4, e23), (0.7, e24)):
        assert srw.RandomSample(lambda: u).secondOrderSample(1.0, 1.0, prev, prev_n, curr_n) == e
    assert rs.computeSecondOrderWeights(2.0, 2.0, prev, [e12, e15], curr_n) == [(1, 0.5), (3, 0.5), (4, 0.5)]
    expect = [(1, 0.5), (3, 0.5), (4, 1.0)]
    assert rs.computeSecondOrderWeights(2.0, 2.0, prev, prev_n, curr_n) == expect
    for u, e in ((0.24, expect[0]), (0.26, expect[1]), (0.51, expect[2]), (0.99, expect[2])):
        assert srw.RandomSample(lambda: u).secondOrderSample(2.0, 2.0, prev, prev_n, curr_n) == e


def test_kat_philox_device(oracle):
    import ctypes as C
    for ctr, key in (([0] * 4, [0] * 2), ([0xffffffff] * 4, [0xffffffff] * 2),
                     ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])):
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        srw.check(srw.lib().srw_philox4x32_10(c, k, o))
        assert list(o) == oracle.philox(ctr, key).tolist()


# ---- P0: GraphMapTest.scala ------------------------------------------------------------------
def test_kat_graphmap():
    e1, e2, e3, e4 = [(2, 1.0)], [(3, 1.0)], [(3, 1.0)], [(1, 1.0)]
    gm = srw.GraphMap()
    gm.addVertex(1, e1)
    gm.addVertex(2)
    assert gm.getNumEdges == 1 and gm.getNumVertices == 2
    assert gm.getNeighbors(1) == e1
    gm.reset()
    gm.addVertex(1, e1 + e2)
    gm.addVertex(2)
    gm.addVertex(3)
    assert gm.getNeighbors(1) == e1 + e2
    gm.reset()
    gm.addVertex(2, e3 + e4)
    gm.addVertex(1, e1 + e2)
    gm.addVertex(3)
    assert gm.getNeighbors(1) == e1 + e2 and gm.getNeighbors(2) == e3 + e4
    assert gm.getNeighbors(3) == [] and gm.getNeighbors(99) is None
    gm.addVertex(1, e3)
    assert gm.getNeighbors(1) == e1 + e2


# ---- P0: load counts + first step (T-URW:33-86, T-VRW:32-85) ----------------------------------
@pytest.mark.parametrize("cls", [srw.UniformRandomWalk, srw.VCutRandomWalk])
def test_load_karate_and_first_step(cls):
    for directed, ne in ((False, 156), (True, 78)):
        rw = cls(srw.Params(input=KARATE, directed=directed))
        paths = rw.loadGraph()
        assert (rw.nEdges, rw.nVertices, len(paths)) == (ne, 34, 34)
    rw = cls(srw.Params(input=TESTGRAPH, directed=True))
    paths = rw.loadGraph()
    res = rw.initFirstStep(paths, lambda: 0.5)
    assert len(res) == len(paths)
    assert sorted(p for _, (p, _) in res) == [[1, 2], [2]]


def test_adjacency_matches_oracle(oracle):
    for directed in (False, True):
        og = oracle.Graph().load_file(KARATE, directed=directed)
        g = srw.UniformRandomWalk(srw.Params(input=KARATE, directed=directed))
        g.loadGraph()
        assert g.graph.vertex_ids().tolist() == og.vertex_ids().tolist()
        for v in og.vertex_ids():
            assert g.graph.neighbors(int(v)) == og.neighbors(int(v))
        assert g.graph.neighbors(1000) is None


# ---- P0: constant-u walk scenarios (T-URW:181-291, T-VRW:189-299) -----------------------------
SCENARIOS = [(False, 0.1, 1), (False, 0.1, 50), (False, 0.9, 50), (True, 0.9, 50), (True, 0.1, 50)]


@pytest.mark.parametrize("cls", [srw.UniformRandomWalk, srw.VCutRandomWalk])
@pytest.mark.parametrize("directed,u,wl", SCENARIOS)
def test_constant_u_walks(oracle, cls, directed, u, wl):
    cfg = srw.Params(input=KARATE, directed=directed, walkLength=wl, rddPartitions=8, numWalks=1)
    rw = cls(cfg)
    graph = rw.loadGraph()
    paths = rw.randomWalk(graph, lambda: u)
    assert paths.count() == rw.nVertices
    og = oracle.Graph().load_file(KARATE, directed=directed, partitioned=cls.partitioned)
    ids, offs = oracle.walk(og, walk_length=wl, num_walks=1, u_const=u)
    assert paths.collect() == oracle.paths_as_lists(ids, offs)


def test_derived_known_answers():
    rw = srw.UniformRandomWalk(srw.Params(input=KARATE, walkLength=10, numWalks=1, p=0.5, q=2.0))
    rw.loadGraph()
    P = {p[0]: p for p in rw.randomWalk(None, lambda: 0.37).collect()}
    assert P[1] == [1, 13] * 6
    rw = srw.UniformRandomWalk(srw.Params(input=KARATE, walkLength=50, numWalks=1))
    rw.loadGraph()
    P = {p[0]: p for p in rw.randomWalk(None, lambda: 0.9).collect()}
    assert P[1][:6] == [1, 3, 8, 4, 8, 4] and P[34][:5] == [34, 32, 33, 32, 33]


# ---- graphs for P1 / P2 -----------------------------------------------------------------------
def _rmat(scale, ef, weighted, seed=42):
    s, d = synth.rmat_edges(scale, ef, seed=seed)
    w = synth.edge_weights(len(s), seed=seed + 1) if weighted else None
    return s, d, w


CASES = [  # scale, ef, weighted, directed, p, q, seed
    (8, 8, False, False, 1.0, 1.0, 1),
    (8, 8, False, False, 0.5, 2.0, 2),
    (9, 8, True, False, 0.5, 2.0, 3),
    (9, 4, True, True, 0.25, 4.0, 4),
    (10, 8, False, True, 2.0, 0.5, 5),
    (10, 16, True, False, 4.0, 0.25, 6),
]


@pytest.mark.parametrize("scale,ef,weighted,directed,p,q,seed", CASES)
def test_p1_exact_sampler_bit_equal(oracle, scale, ef, weighted, directed, p, q, seed):
    s, d, w = _rmat(scale, ef, weighted)
    og = oracle.Graph().load_edges(s, d, w, directed=directed)
    g = srw.Graph.from_edges(s, d, w, directed=directed)
    assert g.stats() == (og.num_vertices, og.num_edges)
    ids, offs = oracle.walk(og, walk_length=20, num_walks=2, p=p, q=q, seed=seed)
    got_ids, got_offs = g.walk(srw.Params(walkLength=20, numWalks=2, p=p, q=q, seed=seed, sampler="exact")).arrays()
    assert (got_offs == offs).all()
    assert (got_ids == ids).all()


@pytest.mark.parametrize("scale,ef,weighted,directed,p,q,seed", CASES)
def test_p2_alias_sampler_bit_equal(oracle, scale, ef, weighted, directed, p, q, seed):
    s, d, w = _rmat(scale, ef, weighted)
    og = oracle.Graph().load_edges(s, d, w, directed=directed)
    twin = oracle.AliasGraph(og)
    g = srw.Graph.from_edges(s, d, w, directed=directed)
    # layout parity: offsets, sorted rows, Vose slots
    tv, lay = twin.view(), g.layout()
    assert (lay["offsets"] == tv["offsets"]).all() and (lay["col"] == tv["col"]).all()
    assert lay["has_alias"] == twin.has_alias == weighted
    if weighted:
        assert (lay["thr"] == tv["thr"]).all() and (lay["alias"] == tv["alias"]).all()
        assert (lay["own"].astype(np.int32) == tv["col"]).all()
    ids, offs, st = twin.walk(walk_length=30, num_walks=3, p=p, q=q, seed=seed)
    paths = g.walk(srw.Params(walkLength=30, numWalks=3, p=p, q=q, seed=seed, sampler="alias"))
    got_ids, got_offs = paths.arrays()
    assert (got_offs == offs).all()
    assert (got_ids == ids).all()
    assert paths.steps() == st.steps == len(ids) - (len(offs) - 1)


def test_alias_karate_text_output_matches_twin(oracle, tmp_path):
    """C1 through the native Main: 34 lines x 12 ids, same multiset of lines as the twin."""
    out = str(tmp_path / "out")
    rc = srw.Main.main(["--cmd", "randomwalk", "--input", KARATE, "--output", out, "--numWalks", "1", "--walkLength", "10", "--seed", "5"])
    assert rc == 0
    text = open(os.path.join(out, "path", "part-00000")).read()
    lines = text.split("\n")
    assert lines[-1] == "" and len(lines) == 35 and all(len(ln.split("\t")) == 12 for ln in lines[:-1])
    assert os.path.exists(os.path.join(out, "path", "_SUCCESS"))
    twin = oracle.AliasGraph(oracle.Graph().load_file(KARATE))
    ids, offs, _ = twin.walk(walk_length=10, num_walks=1, seed=5, fold=1)
    assert sorted(lines[:-1]) == sorted(oracle.format_paths(ids, offs).decode().split("\n")[:-1])
    # a second run into the same directory fails like Hadoop's saveAsTextFile
    assert srw.Main.main(["--cmd", "randomwalk", "--input", KARATE, "--output", out]) != 0


def test_ragged_and_edge_cases(oracle):
    # arbitrary (negative, sparse) ids, self loops, duplicates, dead ends
    txt = "-5 7\n7 7\n7 100000\n7 100000\n100000 -5\n3 -5\n"
    for directed in (False, True):
        og = oracle.Graph().load_text(txt, directed=directed)
        s, d, w, _ = srw.parse_edges(txt)
        g = srw.Graph.from_edges(s, d, w, directed=directed)
        assert g.vertex_ids().tolist() == [-5, 3, 7, 100000]
        for sampler, p, q in (("exact", 0.5, 2.0), ("alias", 0.5, 2.0)):
            got = g.walk(srw.Params(walkLength=6, numWalks=4, p=p, q=q, seed=9, sampler=sampler)).collect()
            if sampler == "exact":
                ids, offs = oracle.walk(og, walk_length=6, num_walks=4, p=p, q=q, seed=9)
            else:
                ids, offs, _ = oracle.AliasGraph(og).walk(walk_length=6, num_walks=4, p=p, q=q, seed=9)
            assert got == oracle.paths_as_lists(ids, offs)
    # empty graph, zero walks
    g = srw.Graph.from_edges(np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert g.stats() == (0, 0) and g.walk(srw.Params(walkLength=5, numWalks=2)).collect() == []
    g = srw.Graph.from_edges([1], [2])
    assert g.walk(srw.Params(walkLength=5, numWalks=0)).collect() == []
    assert g.walk(srw.Params(walkLength=0, numWalks=1)).collect() == [[1, 2], [2, 1]]


def test_device_rmat_generator_matches_numpy():
    import torch
    scale, ef = 10, 8
    n = ef << scale
    ds = torch.empty(n, dtype=torch.int32, device="cuda")
    dd = torch.empty(n, dtype=torch.int32, device="cuda")
    dw = torch.empty(n, dtype=torch.float32, device="cuda")
    srw.check(srw.lib().srw_synth_rmat_device(scale, ef, 42, 0, n, ds.data_ptr(), dd.data_ptr()))
    srw.check(srw.lib().srw_synth_weights_device(43, 0, n, dw.data_ptr()))
    s, d = synth.rmat_edges(scale, ef, seed=42)
    assert (ds.cpu().numpy() == s).all() and (dd.cpu().numpy() == d).all()
    assert (dw.cpu().numpy() == synth.edge_weights(n, seed=43)).all()


def test_full_size_properties():
    """RMAT-20 (BASELINE config C2 size): size-independent properties of the device walk."""
    import torch
    scale, ef = 20, 16
    n = ef << scale
    ds = torch.empty(n, dtype=torch.int32, device="cuda")
    dd = torch.empty(n, dtype=torch.int32, device="cuda")
    srw.check(srw.lib().srw_synth_rmat_device(scale, ef, 42, 0, n, ds.data_ptr(), dd.data_ptr()))
    g = srw.Graph.from_device_edges(n, ds.data_ptr(), dd.data_ptr())
    nv, nnz = g.stats()
    assert nnz == 2 * n and 600000 < nv < (1 << scale)
    lay_off = g.layout()["offsets"]
    assert lay_off[0] == 0 and lay_off[-1] == nnz and (np.diff(lay_off) > 0).all()   # undirected: no dead ends
    L = 80
    paths = torch.empty((nv, L + 2), dtype=torch.int32, device="cuda")
    lens = torch.empty(nv, dtype=torch.int32, device="cuda")
    cp = srw.Params(walkLength=L, numWalks=1, p=0.5, q=2.0, seed=1).to_c()
    import ctypes as C
    srw.check(srw.lib().srw_walk_device(g.h, C.byref(cp), 0, nv, paths.data_ptr(), lens.data_ptr(), None))
    assert bool((lens == L + 2).all())                       # every walk is full length
    vids = torch.from_numpy(g.vertex_ids()).cuda()
    assert bool((paths[:, 0] == vids).all())                 # one walker per vertex, in id order
    # every consecutive pair is an edge: check a sample of walkers against the sorted rows on the host
    lay = g.layout()
    rank = {int(v): i for i, v in enumerate(g.vertex_ids().tolist())} if nv < 2000000 else None
    hp = paths[:: max(1, nv // 2000)].cpu().numpy()
    for row in hp[:200]:
        for a, b in zip(row[:-1], row[1:]):
            ra, rb = rank[int(a)], rank[int(b)]
            seg = lay["col"][lay["offsets"][ra]:lay["offsets"][ra + 1]]
            k = np.searchsorted(seg, rb)
            assert k < len(seg) and seg[k] == rb
    # idempotence: same seed -> same paths; different launch split -> same paths
    paths2 = torch.empty_like(paths)
    half = nv // 2
    srw.check(srw.lib().srw_walk_device(g.h, C.byref(cp), 0, half, paths2.data_ptr(), lens.data_ptr(), None))
    srw.check(srw.lib().srw_walk_device(g.h, C.byref(cp), half, nv - half, paths2[half:].data_ptr(), lens[half:].data_ptr(), None))
    assert bool((paths2 == paths).all())
    wi = srw.last_walk_info()
    assert wi.steps == (nv - half) * (L + 1) and wi.kernel_ms > 0


# ---- BASELINE config C5 shape at test scale: Zipf hub graph, p=0.25 q=4 (membership stress) ----
@pytest.mark.parametrize("sampler", ["alias", "exact"])
def test_c5_zipf_hub_graph(oracle, sampler):
    s, d = synth.zipf_edges(2048, cap=1500, seed=7)
    assert np.bincount(s).max() == 1500                      # a hub with the capped stub count exists
    og = oracle.Graph().load_edges(s, d)
    g = srw.Graph.from_edges(s, d)
    kw = dict(walk_length=12 if sampler == "exact" else 40, num_walks=1 if sampler == "exact" else 3, p=0.25, q=4.0, seed=31)
    prm = srw.Params(walkLength=kw["walk_length"], numWalks=kw["num_walks"], p=0.25, q=4.0, seed=31, sampler=sampler)
    got_ids, got_offs = g.walk(prm).arrays()
    if sampler == "exact":
        ids, offs = oracle.walk(og, **kw)
    else:
        ids, offs, _ = oracle.AliasGraph(og).walk(**kw)
    assert (got_offs == offs).all() and (got_ids == ids).all()


# ---- BASELINE config C3 shape at test scale: weighted RMAT, p=0.5 q=2 (alias build + biased step) ----
def test_c3_weighted_rmat_layout_and_walk(oracle):
    s, d = synth.rmat_edges(13, 16, seed=42)
    w = synth.edge_weights(len(s), seed=43)
    og = oracle.Graph().load_edges(s, d, w)
    twin = oracle.AliasGraph(og)
    g = srw.Graph.from_edges(s, d, w, flags=srw.BUILD_ALIAS)
    tv, lay = twin.view(), g.layout()
    assert (lay["offsets"] == tv["offsets"]).all() and (lay["col"] == tv["col"]).all()
    assert (lay["thr"] == tv["thr"]).all() and (lay["alias"] == tv["alias"]).all()
    ids, offs, st = twin.walk(walk_length=80, num_walks=1, p=0.5, q=2.0, seed=1, fold=1)
    got_ids, got_offs = g.walk(srw.Params(walkLength=80, numWalks=1, p=0.5, q=2.0, seed=1)).arrays()
    assert (got_offs == offs).all() and (got_ids == ids).all()


# ---- the superseded kernel generations (profiles/museum/, a TEST-ONLY library) produce the bits of the product kernels ----
def _museum():
    """profiles/museum/libsrw_museum.so, built on demand with nvcc (present on the GPU box: same image)."""
    import ctypes as C, subprocess
    mdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "museum")
    so = os.path.join(mdir, "libsrw_museum.so")
    srcs = [os.path.join(mdir, f) for f in ("walk_museum.cu", "alias_generations.cuh", "exact_generations.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
        subprocess.check_call(["bash", os.path.join(mdir, "build.sh")])
    srw.lib()
    L = C.CDLL(so)
    L.srw_museum_walk.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p]
    L.srw_museum_walk.restype = C.c_int
    return L


def _museum_walk(L, g, prm, variant, n_walkers):
    """paths as vertex ids (the museum kernels emit ranks), lens"""
    import ctypes as C, torch
    stride = prm.walkLength + 2
    paths = torch.full((n_walkers, stride), -1, dtype=torch.int32, device="cuda")
    lens = torch.zeros(n_walkers, dtype=torch.int32, device="cuda")
    cp = prm.to_c()
    rc = L.srw_museum_walk(g.h, C.byref(cp), variant.encode(), 0, n_walkers, paths.data_ptr(), lens.data_ptr())
    assert rc == 0, (variant, rc)
    vids = np.asarray(g.vertex_ids())
    P, Ln = paths.cpu().numpy(), lens.cpu().numpy()
    return [vids[P[i, :Ln[i]]].tolist() for i in range(n_walkers)]


def _product_paths(g, prm):
    ids, offs = g.walk(prm).arrays()
    return [ids[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


def test_kernel_generations_agree():
    L = _museum()
    s, d = synth.rmat_edges(12, 16, seed=3)
    for w in (None, synth.edge_weights(len(s), seed=4)):
        g = srw.Graph.from_edges(s, d, w)
        prm = srw.Params(walkLength=60, numWalks=2, p=0.5, q=2.0, seed=9, sampler="alias")
        want = _product_paths(g, prm)
        for v in ("alias_v1", "alias_v2", "alias_v3"):
            assert _museum_walk(L, g, prm, v, len(want)) == want, v


# ---- VCut input format through the native Main (--partitioned true, 3rd column = partition id) ----
def test_vcut_cli_partitioned_input(oracle, tmp_path):
    rows = [ln.split() for ln in open(KARATE).read().split("\n") if ln]
    txt = "".join("%s %s %d %.2f\n" % (a, b, i % 4, 1.0 + (i % 7) / 4.0) for i, (a, b) in enumerate(rows))
    inp = tmp_path / "karate_vcut.txt"
    inp.write_text(txt)
    out = str(tmp_path / "o")
    rc = srw.Main.main(["--cmd", "randomwalk", "--input", str(inp), "--output", out, "--partitioned", "true", "--numWalks", "2",
                        "--walkLength", "15", "--p", "0.5", "--q", "2.0", "--seed", "8", "--singleOutput", "false", "--rddPartitions", "4"])
    assert rc == 0
    files = sorted(f for f in os.listdir(os.path.join(out, "path")) if f.startswith("part-"))
    assert files == ["part-0000%d" % k for k in range(4)]
    lines = []
    for f in files:
        lines += open(os.path.join(out, "path", f)).read().split("\n")[:-1]
    og = oracle.Graph().load_text(txt, weighted=True, partitioned=True)
    ids, offs, _ = oracle.AliasGraph(og).walk(walk_length=15, num_walks=2, p=0.5, q=2.0, seed=8, fold=1)
    assert sorted(lines) == sorted(oracle.format_paths(ids, offs).decode().split("\n")[:-1])
    # GraphMap.getPartition analogue
    rw = srw.VCutRandomWalk(srw.Params(input=str(inp), partitioned=True))
    rw.loadGraph()
    assert rw.graph.partition(2) == og.partition(2) and rw.graph.partition(12345) is None


# ---- SRW_SAMPLER_ALIAS_FOLD (kernel v4) == its CPU twin, bit for bit; falls back to alias when not applicable ----
@pytest.mark.parametrize("scale,ef,p,q,seed", [(8, 8, 0.5, 2.0, 1), (10, 16, 0.25, 4.0, 2), (11, 4, 0.1, 0.5, 3), (9, 8, 0.5, 1.0, 4)])
def test_fold_sampler_bit_equal(oracle, scale, ef, p, q, seed):
    s, d = synth.rmat_edges(scale, ef, seed=42)
    og = oracle.Graph().load_edges(s, d)
    twin = oracle.AliasGraph(og)
    g = srw.Graph.from_edges(s, d, None, flags=srw.BUILD_ALIAS)
    ids, offs, st = twin.walk(walk_length=60, num_walks=3, p=p, q=q, seed=seed, fold=1)
    paths = g.walk(srw.Params(walkLength=60, numWalks=3, p=p, q=q, seed=seed, sampler="fold"))
    got_ids, got_offs = paths.arrays()
    assert (got_offs == offs).all()
    assert (got_ids == ids).all()
    # odd stride (scalar path flush) and a text-loaded (weights all 1.0) graph
    ids, offs, _ = twin.walk(walk_length=7, num_walks=2, p=p, q=q, seed=seed, fold=1)
    w1 = np.ones(len(s), np.float32)
    g1 = srw.Graph.from_edges(s, d, w1, flags=srw.BUILD_ALIAS)
    got_ids, got_offs = g1.walk(srw.Params(walkLength=7, numWalks=2, p=p, q=q, seed=seed, sampler="fold")).arrays()
    assert (got_offs == offs).all() and (got_ids == ids).all()


# ---- SRW_BUILD_LEAN: only d_ent + d_hash_id stay in HBM; every alias-class walk gives the full build's (= the twin's) paths ----
@pytest.mark.parametrize("p,q,sampler,fold", [(0.5, 2.0, "fold", 1), (0.25, 4.0, "fold", 1), (2.0, 0.5, "fold", 0), (1.0, 1.0, "fold", 0), (0.5, 2.0, "alias", 0)])
def test_lean_build_same_paths(oracle, p, q, sampler, fold):
    s, d = synth.rmat_edges(10, 16, seed=42)
    og = oracle.Graph().load_edges(s, d)
    twin = oracle.AliasGraph(og)
    full = srw.Graph.from_edges(s, d, None, flags=srw.BUILD_ALIAS)
    lean = srw.Graph.from_edges(s, d, None, flags=srw.BUILD_ALIAS | srw.BUILD_LEAN)
    assert srw.lib().srw_graph_device_bytes(lean.h) < 0.75 * srw.lib().srw_graph_device_bytes(full.h)
    prm = srw.Params(walkLength=40, numWalks=2, p=p, q=q, seed=9, sampler=sampler)
    ids, offs, _ = twin.walk(walk_length=40, num_walks=2, p=p, q=q, seed=9, fold=fold)
    a = full.walk(prm).arrays()
    b = lean.walk(prm).arrays()
    assert srw.lib().srw_last_walk_kernel().decode().endswith(",1>")          # the id-space fold kernel
    assert (a[1] == offs).all() and (a[0] == ids).all()
    assert (b[1] == offs).all() and (b[0] == ids).all()
    # the layout query rebuilds the sorted column array from the neighbour entries
    nv, nnz = lean.stats()
    o1, c1 = np.zeros(nv + 1, np.int64), np.zeros(nnz, np.int32)
    o2, c2 = np.zeros(nv + 1, np.int64), np.zeros(nnz, np.int32)
    srw.check(srw.lib().srw_graph_layout(full.h, o1.ctypes.data, c1.ctypes.data, None, None))
    srw.check(srw.lib().srw_graph_layout(lean.h, o2.ctypes.data, c2.ctypes.data, None, None))
    assert (o1 == o2).all() and (c1 == c2).all()
    prof = __import__("json").loads(srw.lib().srw_graph_build_profile(lean.h).decode())
    assert "k_hash_insert_ids" in prof and "k_hash_insert" not in prof
    # what a lean handle cannot do fails loudly
    with pytest.raises(srw.SrwError):
        lean.walk(srw.Params(walkLength=5, numWalks=1, sampler="exact"))


def test_lean_flag_is_ignored_where_it_does_not_apply(oracle):
    s, d = synth.rmat_edges(8, 8, seed=42)
    w = synth.edge_weights(len(s), seed=43)
    for kw in (dict(w=w, directed=False), dict(w=None, directed=True)):
        g0 = srw.Graph.from_edges(s, d, kw["w"], directed=kw["directed"], flags=srw.BUILD_ALIAS)
        g1 = srw.Graph.from_edges(s, d, kw["w"], directed=kw["directed"], flags=srw.BUILD_ALIAS | srw.BUILD_LEAN)
        prm = srw.Params(walkLength=20, numWalks=2, p=0.5, q=2.0, seed=3, sampler="fold")
        a, b = g0.walk(prm).arrays(), g1.walk(prm).arrays()
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_fold_sampler_fallbacks(oracle):
    s, d = synth.rmat_edges(9, 8, seed=42)
    w = synth.edge_weights(len(s), seed=43)
    for kw, prm in ((dict(w=None, directed=True), (0.5, 2.0)), (dict(w=w, directed=True), (0.5, 2.0)), (dict(w=w, directed=False), (2.0, 0.5)),
                    (dict(w=None, directed=False), (2.0, 0.5))):
        g = srw.Graph.from_edges(s, d, kw["w"], directed=kw["directed"], flags=srw.BUILD_ALIAS)
        a = g.walk(srw.Params(walkLength=30, numWalks=2, p=prm[0], q=prm[1], seed=5, sampler="fold")).arrays()
        b = g.walk(srw.Params(walkLength=30, numWalks=2, p=prm[0], q=prm[1], seed=5, sampler="alias")).arrays()
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


# ---- exact sampler, certified parallel CDF search (walk_exact_cert_kernel): adversarial rows ----
def test_exact_cert_prefix_on_the_boundary(oracle):
    """Unit weights, power-of-two degree, constant u = k/deg: a prefix equals u exactly, so the +-delta band is
    hit and the step must be replayed in order (RS:20 `acc >= u` picks that very prefix)."""
    n = 9                                               # K9: every vertex has 8 unit-weight neighbours
    s, d = np.array([(i, j) for i in range(n) for j in range(i + 1, n)], dtype=np.int32).T
    og = oracle.Graph().load_edges(s, d)
    g = srw.Graph.from_edges(s, d)
    for u in (0.0, 0.125, 0.25, 0.5, 0.875, 0.99999994):
        for p, q in ((1.0, 1.0), (0.5, 2.0)):
            ids, offs = oracle.walk(og, walk_length=12, num_walks=1, p=p, q=q, u_const=u)
            got = g.walk(srw.Params(walkLength=12, numWalks=1, p=p, q=q, sampler="exact"), u_const=u).arrays()
            assert (got[1] == offs).all() and (got[0] == ids).all(), (u, p, q)


@pytest.mark.parametrize("kind", ["wide", "zeros", "tiny", "allzero"])
def test_exact_cert_adversarial_weights(oracle, kind):
    """Weights over 40 orders of magnitude, zero weights (RS:20 can pick a zero-weight edge when u == 0; an all-zero
    row makes sum == 0 and falls through to edges.head), denormal-sized weights."""
    rng = np.random.RandomState(5)
    s, d = synth.rmat_edges(9, 16, seed=11)
    m = len(s)
    if kind == "wide":
        w = (10.0 ** rng.uniform(-20, 20, m)).astype(np.float32)
    elif kind == "zeros":
        w = np.where(rng.rand(m) < 0.4, 0.0, rng.rand(m)).astype(np.float32)
    elif kind == "tiny":
        w = (rng.rand(m) * 1e-38).astype(np.float32)
    else:
        w = np.zeros(m, np.float32)
    og = oracle.Graph().load_edges(s, d, w)
    g = srw.Graph.from_edges(s, d, w)
    for p, q, seed in ((1.0, 1.0, 1), (0.5, 2.0, 2), (4.0, 0.25, 3)):
        ids, offs = oracle.walk(og, walk_length=15, num_walks=2, p=p, q=q, seed=seed)
        got = g.walk(srw.Params(walkLength=15, numWalks=2, p=p, q=q, seed=seed, sampler="exact")).arrays()
        assert (got[1] == offs).all() and (got[0] == ids).all(), (kind, p, q)


def test_exact_kernel_generations_agree():
    """thread (one walker per thread), warp (in-order fold through shuffles) and cert (first certified parallel search), kept in
    profiles/museum/, produce the bits of the product's walk_exact_cert2_kernel on weighted and hub graphs."""
    L = _museum()
    cases = []
    s, d = synth.rmat_edges(12, 16, seed=3)
    for w in (None, synth.edge_weights(len(s), seed=4)):
        cases.append((srw.Graph.from_edges(s, d, w), srw.Params(walkLength=20, numWalks=1, p=0.5, q=2.0, seed=9, sampler="exact")))
    s, d = synth.zipf_edges(4096, cap=3000, seed=7)
    cases.append((srw.Graph.from_edges(s, d), srw.Params(walkLength=20, numWalks=1, p=0.25, q=4.0, seed=2, sampler="exact")))
    for g, prm in cases:
        want = _product_paths(g, prm)
        for v in ("exact_thread", "exact_warp", "exact_cert"):
            assert _museum_walk(L, g, prm, v, len(want)) == want, v


# ---- the alias-fold kernel in rank space and in id space, and its pre-convergence generation (museum), produce the twin's bits ----
def test_fold_kernel_generations_agree(oracle, monkeypatch):
    L = _museum()
    s, d = synth.rmat_edges(13, 16, seed=3)
    monkeypatch.setenv("SRW_FOLD_IDS", "0")        # v4 walks rank-labelled entries only: build this handle in rank space
    g = srw.Graph.from_edges(s, d, None, flags=srw.BUILD_ALIAS)
    monkeypatch.delenv("SRW_FOLD_IDS")
    g_ids = srw.Graph.from_edges(s, d, None, flags=srw.BUILD_ALIAS)       # the default: id space
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d))
    for wl in (80, 13, 0):       # even and odd strides: both alignments of the staged path stores
        prm = srw.Params(walkLength=wl, numWalks=2, p=0.5, q=2.0, seed=9, sampler="fold")
        want_ids, want_offs, _ = twin.walk(walk_length=wl, num_walks=2, p=0.5, q=2.0, seed=9, fold=1)
        want = oracle.paths_as_lists(want_ids, want_offs)
        for name, h in (("rank-space", g), ("id-space", g_ids)):
            ids, offs = h.walk(prm).arrays()
            assert np.array_equal(ids, want_ids) and np.array_equal(offs, want_offs), (wl, name)
        assert _museum_walk(L, g, prm, "fold_v4", len(want)) == want, wl
        import ctypes as C
        cp = prm.to_c()
        assert L.srw_museum_walk(g_ids.h, C.byref(cp), b"fold_v4", 0, 1, None, None) != 0   # refuses id-labelled entries


# ---- SRW_SAMPLER_ALIAS_FOLD on WEIGHTED undirected graphs (walk_wfold_conv_kernel): device build (row weight sums, bundle
# weights, 32-byte slots) + kernel == the CPU twin, bit for bit ----
@pytest.mark.parametrize("scale,ef,p,q,seed", [(8, 8, 0.5, 2.0, 1), (10, 16, 0.25, 4.0, 2), (11, 4, 0.1, 0.5, 3), (9, 8, 0.5, 1.0, 4)])
def test_weighted_fold_sampler_bit_equal(oracle, scale, ef, p, q, seed):
    s, d = synth.rmat_edges(scale, ef, seed=42)
    w = synth.edge_weights(len(s), seed=43)
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d, w))
    g = srw.Graph.from_edges(s, d, w, flags=srw.BUILD_ALIAS)
    for wl, rounds in ((60, 3), (7, 2)):
        ids, offs, st = twin.walk(walk_length=wl, num_walks=rounds, p=p, q=q, seed=seed, fold=1)
        got_ids, got_offs = g.walk(srw.Params(walkLength=wl, numWalks=rounds, p=p, q=q, seed=seed, sampler="fold")).arrays()
        assert (got_offs == offs).all() and (got_ids == ids).all()
    # ... and it is not the classic sampler's output
    c_ids, _ = g.walk(srw.Params(walkLength=7, numWalks=2, p=p, q=q, seed=seed, sampler="alias")).arrays()
    assert not np.array_equal(c_ids, got_ids)


def test_weighted_fold_hubs_bundles_and_text_weights(oracle, tmp_path):
    """Zipf hubs, parallel edges with different weights (bundle weight != any single weight), self-loops; then karate
    with printed weights through the CLI (`--sampler fold --weighted true`)."""
    zs, zd = synth.zipf_edges(2048, seed=7, cap=600)
    w = synth.edge_weights(len(zs), seed=5)
    es, ed = np.array([0, 0, 0, 5, 5, 9], np.int32), np.array([1, 1, 1, 5, 6, 9], np.int32)
    ew = np.array([0.25, 1.5, 3.0, 2.0, 0.125, 7.0], np.float32)
    s, d, w = np.concatenate([zs, es]), np.concatenate([zd, ed]), np.concatenate([w, ew])
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d, w))
    g = srw.Graph.from_edges(s, d, w, flags=srw.BUILD_ALIAS)
    ids, offs, _ = twin.walk(walk_length=25, num_walks=2, p=0.25, q=4.0, seed=21, fold=1)
    got = g.walk(srw.Params(walkLength=25, numWalks=2, p=0.25, q=4.0, seed=21, sampler="fold")).arrays()
    assert (got[1] == offs).all() and (got[0] == ids).all()
    rows = [ln.split() for ln in open(KARATE).read().split("\n") if ln]
    txt = "".join("%s %s %.2f\n" % (a, b, 0.5 + (i % 7) / 4.0) for i, (a, b) in enumerate(rows))
    inp = tmp_path / "kw.txt"
    inp.write_text(txt)
    out = str(tmp_path / "o")
    assert srw.Main.main(["--cmd", "randomwalk", "--input", str(inp), "--output", out, "--numWalks", "3", "--walkLength", "20", "--p", "0.5",
                          "--q", "2.0", "--seed", "8", "--sampler", "fold"]) == 0
    og = oracle.Graph().load_text(txt, weighted=True)
    ids, offs, _ = oracle.AliasGraph(og).walk(walk_length=20, num_walks=3, p=0.5, q=2.0, seed=8, fold=1)
    assert open(os.path.join(out, "path", "part-00000"), "rb").read() == oracle.format_paths(ids, offs)


# ---- srw_walk_device_async / srw_walk_wait: rounds enqueued back to back, finalisation on the library's stream ----
def test_async_rounds_equal_blocking_rounds():
    import ctypes as C
    import torch
    s, d = synth.rmat_edges(14, 16, seed=42)
    g = srw.Graph.from_edges(s, d, None, flags=srw.BUILD_ALIAS)
    nv = g.num_vertices
    L = 80
    lib = srw.lib()
    cp = srw.Params(walkLength=L, numWalks=1, p=0.5, q=2.0, seed=6, sampler="fold").to_c()
    want = []
    p0 = torch.empty((nv, L + 2), dtype=torch.int32, device="cuda")
    l0 = torch.empty(nv, dtype=torch.int32, device="cuda")
    for r in range(5):
        srw.check(lib.srw_walk_device(g.h, C.byref(cp), r * nv, nv, p0.data_ptr(), l0.data_ptr(), None))
        want.append((p0.clone(), l0.clone(), srw.last_walk_info().steps))
    bufs = [(torch.empty_like(p0), torch.empty_like(l0)) for _ in range(2)]
    st = torch.cuda.Stream()
    tickets, got = [None, None], []

    def wait(b):
        if tickets[b] is not None:
            wi = srw.WalkInfo()
            srw.check(lib.srw_walk_wait(tickets[b], C.byref(wi)))
            tickets[b] = None
            got.append((bufs[b][0].clone(), bufs[b][1].clone(), wi.steps, wi.kernel_ms))

    for r in range(5):
        b = r & 1
        wait(b)
        tk = C.c_void_p()
        srw.check(lib.srw_walk_device_async(g.h, C.byref(cp), r * nv, nv, bufs[b][0].data_ptr(), bufs[b][1].data_ptr(), st.cuda_stream, C.byref(tk)))
        tickets[b] = tk
    wait(1)
    wait(0)
    assert len(got) == 5
    for (wp, wl, ws), (gp, gl, gs, ms) in zip(want, got):
        assert bool((wp == gp).all()) and bool((wl == gl).all()) and ws == gs and ms > 0
    # an empty launch still hands out a ticket that can be waited for
    tk = C.c_void_p()
    srw.check(lib.srw_walk_device_async(g.h, C.byref(cp), 0, 0, None, None, None, C.byref(tk)))
    wi = srw.WalkInfo()
    srw.check(lib.srw_walk_wait(tk, C.byref(wi)))
    assert wi.steps == 0


# ---- id-space fold (the default; SRW_FOLD_IDS=0 at build time opts out): the walk emits original ids, no rank -> id pass ----
def test_fold_in_id_space(oracle, monkeypatch):
    s, d = synth.rmat_edges(11, 8, seed=42)
    s, d = (s * 7 + 3).astype(np.int32), (d * 7 + 3).astype(np.int32)          # rank != id everywhere
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d))
    g = srw.Graph.from_edges(s, d, None, flags=srw.BUILD_ALL)          # id space is the default
    for wl in (60, 7):
        ids, offs, st = twin.walk(walk_length=wl, num_walks=2, p=0.5, q=2.0, seed=4, fold=1)
        got = g.walk(srw.Params(walkLength=wl, numWalks=2, p=0.5, q=2.0, seed=4, sampler="fold"))
        gi, go = got.arrays()
        assert (go == offs).all() and (gi == ids).all() and got.steps() == st.steps
    # the other samplers keep working on the same handle (they use the rank-labelled hash sets and columns)
    ids, offs, _ = twin.walk(walk_length=20, num_walks=1, p=2.0, q=0.5, seed=4, fold=1)            # not foldable: classic kernel
    gi, go = g.walk(srw.Params(walkLength=20, numWalks=1, p=2.0, q=0.5, seed=4, sampler="fold")).arrays()
    assert (go == offs).all() and (gi == ids).all()
    og = oracle.Graph().load_edges(s, d)
    ids, offs = oracle.walk(og, walk_length=10, num_walks=1, p=0.5, q=2.0, seed=4)
    gi, go = g.walk(srw.Params(walkLength=10, numWalks=1, p=0.5, q=2.0, seed=4, sampler="exact")).arrays()
    assert (go == offs).all() and (gi == ids).all()


# ---------------------------------------------------------------------------------------------
# Parity beyond the small cases (round-2 additions): samples of larger graphs against the checkers
# ---------------------------------------------------------------------------------------------
def _sampled(ids, offs, mod):
    """rows of a full walk whose walker id is a multiple of `mod`, as (ids, offsets)"""
    keep = np.arange(len(offs) - 1) % mod == 0
    lens = np.diff(offs)[keep]
    starts = offs[:-1][keep]
    out = np.concatenate([ids[s:s + n] for s, n in zip(starts, lens)]) if keep.any() else np.zeros(0, np.int32)
    return out, np.concatenate([[0], np.cumsum(lens)])


@pytest.mark.parametrize("scale,mod,wl", [(16, 13, 20), (20, 1300, 6)])
def test_exact_sampler_sample_of_larger_graphs(oracle, scale, mod, wl):
    """P1 at RMAT-16 and RMAT-20 (SURVEY 7 step 3 asks for RMAT-12..16; C2's graph is RMAT-20): the exact sampler's paths for
    every `mod`-th walker == the oracle (the literal reference algorithm, O(deg(curr) * deg(prev)) per step -- hence a sample)."""
    s, d = synth.rmat_edges(scale, 16, seed=42)
    og = oracle.Graph().load_edges(s, d, None)
    g = srw.Graph.from_edges(s, d, None)
    assert g.stats() == (og.num_vertices, og.num_edges)
    ids, offs = oracle.walk(og, walk_length=wl, num_walks=1, p=0.5, q=2.0, seed=7, sample_mod=mod)
    got_ids, got_offs = g.walk(srw.Params(walkLength=wl, numWalks=1, p=0.5, q=2.0, seed=7, sampler="exact")).arrays()
    s_ids, s_offs = _sampled(got_ids, got_offs, mod)
    assert len(offs) - 1 == len(s_offs) - 1 >= 256
    assert (s_offs == offs).all() and (s_ids == ids).all()


def test_c3_weighted_fold_rmat18_sample(oracle):
    """C3's sampler (weighted alias-fold: Vose slots with bundle weights) at RMAT-18 against the CPU twin, every 97th walker."""
    s, d = synth.rmat_edges(18, 16, seed=42)
    w = synth.edge_weights(len(s), seed=43)
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d, w))
    ids, offs, st = twin.walk(walk_length=80, num_walks=1, p=0.5, q=2.0, seed=3, fold=1, sample_mod=97)
    g = srw.Graph.from_edges(s, d, w, flags=srw.BUILD_ALIAS)
    got_ids, got_offs = g.walk(srw.Params(walkLength=80, numWalks=1, p=0.5, q=2.0, seed=3, sampler="fold")).arrays()
    s_ids, s_offs = _sampled(got_ids, got_offs, 97)
    assert len(offs) - 1 >= 1000
    assert (s_offs == offs).all() and (s_ids == ids).all()


def test_c5_zipf_hub_of_2e5_entries(oracle):
    """C5's shape with a hub row of 200 000 stubs (+ its in-edges): alias-fold at p = 0.25, q = 4 against the twin, every 31st walker.
    The hub's hash set spans ~1e5 buckets and most second-order tests are against it."""
    s, d = synth.zipf_edges(1 << 18, cap=200000, seed=7)
    twin = oracle.AliasGraph(oracle.Graph().load_edges(s, d, None))
    assert int(np.diff(twin.view()["offsets"]).max()) >= 200000
    ids, offs, st = twin.walk(walk_length=40, num_walks=1, p=0.25, q=4.0, seed=5, fold=1, sample_mod=31)
    g = srw.Graph.from_edges(s, d, None, flags=srw.BUILD_ALIAS)
    got_ids, got_offs = g.walk(srw.Params(walkLength=40, numWalks=1, p=0.25, q=4.0, seed=5, sampler="fold")).arrays()
    s_ids, s_offs = _sampled(got_ids, got_offs, 31)
    assert (s_offs == offs).all() and (s_ids == ids).all()


def test_offsets_beyond_2_31_hub_row_last(oracle):
    """More than 2^31 adjacency entries with ONE huge row placed LAST: 2^20 vertices on a ring, each with 1030 parallel edges to a
    hub whose id is the largest, so the hub's row starts beyond entry 1.08e9 and ends beyond 2^31 -- u32 row offsets, the 64-bit
    `off * 16` address math, hash placement at bucket ~5e8 and 24-bit multiplicities are all exercised.  Checker: the CPU twin
    over the analytically built CSR (oracle_fold_walk_csr_timed), every 251st walker."""
    import ctypes as C
    import torch
    if torch.cuda.mem_get_info()[1] < 150e9:
        pytest.skip("needs a 180 GB device")
    K, R = 1 << 20, 1030
    H = K
    ar = torch.arange(K, dtype=torch.int32, device="cuda")
    src = torch.cat([ar.repeat_interleave(R), ar])                      # spokes (i, H) x R, then the ring (i, i+1 mod K)
    dst = torch.cat([torch.full((K * R,), H, dtype=torch.int32, device="cuda"), (ar + 1) % K])
    g = srw.Graph.from_device_edges(src.numel(), src.data_ptr(), dst.data_ptr(), None, False, srw.BUILD_ALIAS)
    del src, dst
    torch.cuda.empty_cache()
    nv, nnz = g.stats()
    assert nv == K + 1 and nnz == 2 * (K * R + K) and nnz > (1 << 31)
    # the same CSR on the host, neighbour-sorted: row i = [i-1, i+1 (sorted)] + [H] * R, row H = each i repeated R times
    off = np.empty(nv + 1, np.int64)
    off[:K + 1] = np.arange(K + 1, dtype=np.int64) * (R + 2)
    off[K + 1] = nnz
    col = np.empty(nnz, np.int32)
    rows = col[:K * (R + 2)].reshape(K, R + 2)
    i = np.arange(K, dtype=np.int64)
    ring = np.sort(np.stack([(i - 1) % K, (i + 1) % K], 1), 1).astype(np.int32)
    rows[:, :2] = ring
    rows[:, 2:] = H
    col[K * (R + 2):] = np.repeat(np.arange(K, dtype=np.int32), R)
    L = oracle.lib()
    fn = L.oracle_fold_walk_csr_timed
    fn.restype = C.c_int64
    fn.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.POINTER(C.c_double),
                   C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.c_void_p]
    wl, mod = 40, 251
    n_s = (nv + mod - 1) // mod
    tw = np.full((n_s, wl + 2), -2, np.int32)
    cfg = oracle.make_cfg(walk_length=wl, num_walks=1, p=0.5, q=2.0, seed=11, threads=0, fold=1)
    el, dn, ck = C.c_double(), C.c_int64(), C.c_uint64()
    fn(nv, off.ctypes.data, col.ctypes.data, C.addressof(cfg), mod, 0, 600.0, C.byref(el), C.byref(dn), C.byref(ck), tw.ctypes.data)
    assert dn.value == n_s
    paths = torch.empty((nv, wl + 2), dtype=torch.int32, device="cuda")
    lens = torch.empty(nv, dtype=torch.int32, device="cuda")
    cp = srw.Params(walkLength=wl, numWalks=1, p=0.5, q=2.0, seed=11, sampler="fold").to_c()
    srw.check(srw.lib().srw_walk_device(g.h, C.byref(cp), 0, nv, paths.data_ptr(), lens.data_ptr(), None))
    got = paths[::mod].cpu().numpy()
    assert (lens.cpu().numpy() == wl + 2).all()
    assert (got == tw).all()            # ids == ranks here (ids 0..K are all present)
    assert (got == H).mean() > 0.3      # the walk really lives on the hub row
