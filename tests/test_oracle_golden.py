"""Pins the CPU oracle on every known-answer test the reference holds for the hot path.

Each test names the reference test it restates (paths relative to
/root/reference/randomwalk/src/test/scala/au/csiro/data61/randomwalk/algorithm/).
"""
import numpy as np
import pytest

from conftest import KARATE, TESTGRAPH


# ---- RandomSampleTest.scala:9-24 ---------------------------------------------------------
def test_random_sample_function(oracle):
    e1, e2, e3 = (1, 1.0), (2, 1.0), (3, 1.0)
    edges = [e1, e2, e3]
    assert oracle.sample(edges, 0.1) == e1
    assert oracle.sample(edges, 0.4) == e2
    assert oracle.sample(edges, 0.7) == e3


# ---- RandomSampleTest.scala:26-94 --------------------------------------------------------
def test_second_order_random_selection(oracle):
    w1 = 1.0
    e12, e21, e23, e24, e14, e15 = (2, w1), (1, w1), (3, w1), (4, w1), (4, w1), (5, w1)
    prev = 1
    prev_n = [e12, e14, e15]
    curr_n = [e21, e23, e24]
    p = q = 1.0
    assert oracle.second_order_weights(p, q, prev, prev_n, curr_n) == curr_n          # :42-44
    assert oracle.second_order_sample(p, q, prev, prev_n, curr_n, 0.1) == e21          # :46-48
    assert oracle.second_order_sample(p, q, prev, prev_n, curr_n, 0.4) == e23          # :50-52
    assert oracle.second_order_sample(p, q, prev, prev_n, curr_n, 0.7) == e24          # :54-56
    p = q = 2.0
    prev_n = [e12, e15]
    assert oracle.second_order_weights(p, q, prev, prev_n, curr_n) == [(1, w1 / p), (3, w1 / q), (4, w1 / q)]  # :58-66
    prev_n = [e12, e14, e15]
    expect = [(1, w1 / p), (3, w1 / q), (4, w1)]
    assert oracle.second_order_weights(p, q, prev, prev_n, curr_n) == expect          # :68-75
    assert oracle.second_order_sample(p, q, prev, prev_n, curr_n, 0.24) == expect[0]   # :76 (biased weight returned)
    assert oracle.second_order_sample(p, q, prev, prev_n, curr_n, 0.26) == expect[1]   # :78-80
    assert oracle.second_order_sample(p, q, prev, prev_n, curr_n, 0.51) == expect[2]   # :82-84
    assert oracle.second_order_sample(p, q, prev, prev_n, curr_n, 0.99) == expect[2]   # :86-88
    assert curr_n == [(1, 1.0), (3, 1.0), (4, 1.0)]                                     # :91-93 inputs unmodified


# ---- GraphMapTest.scala:7-33 -------------------------------------------------------------
def test_graphmap_data_structure(oracle):
    e1, e2, e3, e4 = [(2, 1.0)], [(3, 1.0)], [(3, 1.0)], [(1, 1.0)]
    g = oracle.Graph()
    g.add_vertex(1, e1)
    g.add_vertex(2)
    assert g.num_edges == 1 and g.num_vertices == 2
    assert g.neighbors(1) == e1
    g.reset()
    g.add_vertex(1, e1 + e2)
    g.add_vertex(2)
    g.add_vertex(3)
    assert g.neighbors(1) == e1 + e2
    g.reset()
    g.add_vertex(2, e3 + e4)
    g.add_vertex(1, e1 + e2)
    g.add_vertex(3)
    assert g.neighbors(1) == e1 + e2
    assert g.neighbors(2) == e3 + e4
    assert g.neighbors(3) == []          # GM:112-114 index -1 -> empty
    assert g.neighbors(99) is None       # GM:118 unknown -> null
    g.add_vertex(1, e3)                  # GM:42,54 first insert wins
    assert g.neighbors(1) == e1 + e2


# ---- UniformRandomWalkTest.scala:33-67 / VCutRandomWalkTest.scala:32-66 ------------------
@pytest.mark.parametrize("partitioned", [False, True])
def test_load_karate_counts(oracle, partitioned):
    g = oracle.Graph().load_file(KARATE, directed=False, partitioned=partitioned)
    assert (g.num_edges, g.num_vertices) == (156, 34)
    g = oracle.Graph().load_file(KARATE, directed=True, partitioned=partitioned)
    assert (g.num_edges, g.num_vertices) == (78, 34)


# ---- UniformRandomWalkTest.scala:69-86 / VCutRandomWalkTest.scala:68-85 ------------------
@pytest.mark.parametrize("partitioned", [False, True])
def test_first_step_on_testgraph(oracle, partitioned):
    g = oracle.Graph().load_file(TESTGRAPH, directed=True, partitioned=partitioned)
    assert g.num_vertices == 2
    ids, offs = oracle.walk(g, walk_length=0, num_walks=1, seed=7)
    assert oracle.paths_as_lists(ids, offs) == [[1, 2], [2]]


# ---- T-URW:293-321 restated independently in Python (the serial helper the tests compare to) --
def _py_sample(edges, u):
    s = 0.0
    for _, w in edges:
        s = s + float(np.float32(w))
    acc = 0.0
    for e in edges:
        acc += float(np.float32(e[1])) / s
        if acc >= float(np.float32(u)):
            return e
    return edges[0]


def _py_second_order(p, q, prev, pn, cn, u):
    p, q = np.float32(p), np.float32(q)
    pset = [d for d, _ in pn]
    nw = []
    for d, w in cn:
        w = np.float32(w)
        x = w / q
        if d == prev:
            x = w / p
        elif d in pset:
            x = w
        nw.append((d, float(x)))
    return _py_sample(nw, u)


def _py_walk(g, src, walk_length, u, p=1.0, q=1.0):
    path = [src]
    nb = g.neighbors(src)
    if not nb:
        return path
    path.append(_py_sample(nb, u)[0])
    for _ in range(walk_length):
        curr, prev = path[-1], path[-2]
        cn = g.neighbors(curr)
        if not cn:
            return path
        path.append(_py_second_order(p, q, prev, g.neighbors(prev), cn, u)[0])
    return path


SCENARIOS = [  # (directed, u, walkLength): T-URW:181-291, T-VRW:189-299
    (False, 0.1, 1), (False, 0.1, 50), (False, 0.9, 50), (False, 0.1, 50), (True, 0.9, 50), (True, 0.1, 50)]


@pytest.mark.parametrize("partitioned", [False, True])
@pytest.mark.parametrize("directed,u,wl", SCENARIOS)
def test_second_order_walk_scenarios(oracle, partitioned, directed, u, wl):
    g = oracle.Graph().load_file(KARATE, directed=directed, partitioned=partitioned)
    ids, offs = oracle.walk(g, walk_length=wl, num_walks=1, u_const=u)
    paths = oracle.paths_as_lists(ids, offs)
    assert len(paths) == g.num_vertices                     # "a path per vertex"
    for pth in paths:
        assert pth == _py_walk(g, pth[0], wl, u)


# ---- derived known answers --------------------------------------
def _by_start(paths):
    return {p[0]: p for p in paths}


def test_derived_known_answers(oracle):
    und = oracle.Graph().load_file(KARATE, directed=False)
    dire = oracle.Graph().load_file(KARATE, directed=True)
    P = _by_start(oracle.paths_as_lists(*oracle.walk(und, walk_length=1, num_walks=1, u_const=0.1)))
    assert all(len(p) == 3 for p in P.values())
    assert P[1] == [1, 22, 1] and P[34] == [34, 10, 3]
    P = _by_start(oracle.paths_as_lists(*oracle.walk(und, walk_length=50, num_walks=1, u_const=0.1)))
    assert all(len(p) == 52 for p in P.values())
    assert P[1][:6] == [1, 22, 1, 22, 1, 22] and P[34][:8] == [34, 10, 3, 2, 1, 22, 1, 22]
    P = _by_start(oracle.paths_as_lists(*oracle.walk(und, walk_length=50, num_walks=1, u_const=0.9)))
    assert P[1][:6] == [1, 3, 8, 4, 8, 4] and P[34][:5] == [34, 32, 33, 32, 33]
    P = _by_start(oracle.paths_as_lists(*oracle.walk(dire, walk_length=50, num_walks=1, u_const=0.1)))
    assert {len(p) for p in P.values()} <= {1, 2, 3} and P[1] == [1, 22] and P[34] == [34]
    P = _by_start(oracle.paths_as_lists(*oracle.walk(dire, walk_length=50, num_walks=1, u_const=0.9)))
    assert {len(p) for p in P.values()} <= {1, 2, 3, 4, 5} and P[1] == [1, 3, 4, 8] and P[34] == [34]
    P = _by_start(oracle.paths_as_lists(*oracle.walk(und, walk_length=10, num_walks=1, p=0.5, q=2.0, u_const=0.37)))
    assert P[1] == [1, 13] * 6


# ---- BASELINE config C1: karate, numWalks 1, walkLength 10 -> 34 lines x 12 ids ---------------
def test_c1_output_format(oracle):
    g = oracle.Graph().load_file(KARATE)
    ids, offs = oracle.walk(g, walk_length=10, num_walks=1, seed=1)
    text = oracle.format_paths(ids, offs).decode()
    lines = text.split("\n")
    assert lines[-1] == "" and len(lines) == 35
    for ln in lines[:-1]:
        toks = ln.split("\t")
        assert len(toks) == 12 and all(t.lstrip("-").isdigit() for t in toks)
    # every consecutive pair is an edge of the graph
    for ln in lines[:-1]:
        t = [int(x) for x in ln.split("\t")]
        for a, b in zip(t, t[1:]):
            assert b in [d for d, _ in g.neighbors(a)]


# ---- parse rules: URW:26-34, VRW:21-34 ----------------------------------------------------------
def test_parse_rules(oracle):
    g = oracle.Graph().load_text("1 2 0.5\n2\t3   2.5\n", weighted=True)
    assert g.neighbors(1) == [(2, 0.5)] and g.neighbors(2) == [(1, 0.5), (3, 2.5)]
    g = oracle.Graph().load_text("1 2 0.5\n", weighted=False)
    assert g.neighbors(1) == [(2, 1.0)]
    g = oracle.Graph().load_text("1 2 abc\n", weighted=True)              # Try(...).getOrElse(1.0f)
    assert g.neighbors(1) == [(2, 1.0)]
    # VRW: 3rd column is the partition id; weight only with > 3 columns
    g = oracle.Graph().load_text("1 2 7\n2 3 5 0.25\n", weighted=True, partitioned=True)
    assert g.neighbors(1) == [(2, 1.0)] and g.neighbors(3) == [(2, 0.25)]
    assert g.partition(2) in (5, 7) and g.partition(3) == 5
    # duplicates kept, self loop undirected = two entries, directed = one
    g = oracle.Graph().load_text("4 4\n4 5\n4 5\n")
    assert g.neighbors(4) == [(4, 1.0), (4, 1.0), (5, 1.0), (5, 1.0)]
    g = oracle.Graph().load_text("4 4\n", directed=True)
    assert g.neighbors(4) == [(4, 1.0)]
    # failures the reference throws on
    for bad in ["1 2\n\n3 4\n", " 1 2\n", "1\n", "1 x\n", "1 2147483648\n", "1.0 2\n"]:
        with pytest.raises(ValueError):
            oracle.Graph().load_text(bad)
    # CRLF and a missing final newline are fine; negative ids are legal ints
    g = oracle.Graph().load_text("-1 2\r\n2 +3")
    assert g.neighbors(2) == [(-1, 1.0), (3, 1.0)]


# ---- Philox4x32-10 known answers (Random123 kat_vectors) ----------------------------------
def test_philox_kat(oracle):
    assert oracle.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = oracle.u01(1, 2, 3)
    assert 0.0 <= u < 1.0 and (u * 2 ** 24) == int(u * 2 ** 24)


# ---- philox-driven walk equals the Python restatement fed the same u stream --------------------
def test_philox_walk_matches_python(oracle):
    g = oracle.Graph().load_file(KARATE)
    ids, offs = oracle.walk(g, walk_length=12, num_walks=2, p=0.5, q=2.0, seed=99)
    paths = oracle.paths_as_lists(ids, offs)
    vids = g.vertex_ids().tolist()
    nv = len(vids)
    assert len(paths) == 2 * nv
    for i, pth in enumerate(paths):
        assert pth[0] == vids[i % nv]
        walker = i
        path = [pth[0]]
        path.append(_py_sample(g.neighbors(path[0]), oracle.u01(99, walker, 0))[0])
        while len(path) != 14:
            curr, prev = path[-1], path[-2]
            path.append(_py_second_order(0.5, 2.0, prev, g.neighbors(prev), g.neighbors(curr),
                                         oracle.u01(99, walker, len(path) - 1))[0])
        assert pth == path
