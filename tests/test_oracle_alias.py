"""CPU twin of the alias sampler (oracle/): Vose tables are exact, and the rejection sampler's
empirical transition frequencies match the reference's exact second-order probabilities
(RS:27-44 weights normalised) -- parity level P2's distribution half, on the CPU."""
import importlib

import numpy as np
import pytest

from conftest import KARATE

synth = importlib.import_module("stellar-random-walk_b200.synth")


def _weighted_karate(oracle):
    rng = np.random.RandomState(5)
    rows = [ln.split() for ln in open(KARATE).read().split("\n") if ln]
    txt = "".join("%s %s %.3f\n" % (a, b, 0.25 + rng.randint(0, 16) / 4.0) for a, b in rows)
    return oracle.Graph().load_text(txt, weighted=True)


def test_vose_tables_reconstruct_weights(oracle):
    g = _weighted_karate(oracle)
    a = oracle.AliasGraph(g)
    assert a.has_alias
    v = a.view()
    for r in range(a.nv):
        lo, hi = v["offsets"][r], v["offsets"][r + 1]
        n = hi - lo
        w = v["w"][lo:hi].astype(np.float64)
        thr = v["thr"][lo:hi].astype(np.float64)
        thr[v["thr"][lo:hi] == 0xFFFFFFFF] = 2.0 ** 32
        mass = thr / 2.0 ** 32
        got = mass.copy()
        np.add.at(got, v["alias"][lo:hi].astype(np.int64), 1.0 - mass)
        np.testing.assert_allclose(got / n, w / w.sum(), rtol=0, atol=1e-6)
        assert (np.diff(v["col"][lo:hi]) >= 0).all()


def test_unweighted_graph_has_no_alias_arrays(oracle):
    a = oracle.AliasGraph(oracle.Graph().load_file(KARATE))
    assert not a.has_alias and "thr" not in a.view()


def test_thresholds(oracle):
    assert oracle.alias_thresholds(1.0, 1.0) == (2 ** 32, 2 ** 32, 2 ** 32)
    assert oracle.alias_thresholds(0.5, 2.0) == (2 ** 32, 2 ** 31, 2 ** 30)
    assert oracle.alias_thresholds(0.25, 4.0) == (2 ** 32, 2 ** 30, 2 ** 28)
    assert oracle.alias_thresholds(2.0, 0.5) == (2 ** 30, 2 ** 31, 2 ** 32)


def _exact_transition(oracle, g, prev, curr, p, q):
    bw = oracle.second_order_weights(p, q, prev, g.neighbors(prev), g.neighbors(curr))
    probs = {}
    tot = sum(float(w) for _, w in bw)
    for d, w in bw:
        probs[d] = probs.get(d, 0.0) + float(w) / tot
    return probs


@pytest.mark.parametrize("weighted,p,q", [(False, 0.5, 2.0), (True, 0.25, 4.0), (True, 2.0, 0.5), (False, 1.0, 1.0)])
def test_alias_walk_distribution(oracle, weighted, p, q):
    g = _weighted_karate(oracle) if weighted else oracle.Graph().load_file(KARATE)
    a = oracle.AliasGraph(g)
    ids, offs, st = a.walk(walk_length=40, num_walks=300, p=p, q=q, seed=11)
    paths = ids.reshape(-1, 42)
    assert st.steps == paths.shape[0] * 41
    # transition counts per (prev, curr) context
    ctx = {}
    for k in range(2, 42):
        for a_, b_, c_ in zip(paths[:, k - 2], paths[:, k - 1], paths[:, k]):
            ctx.setdefault((int(a_), int(b_)), {}).setdefault(int(c_), 0)
            ctx[(int(a_), int(b_))][int(c_)] += 1
    chi2, dof = 0.0, 0
    for (pv, cu), cnt in ctx.items():
        n = sum(cnt.values())
        if n < 400:
            continue
        probs = _exact_transition(oracle, g, pv, cu, p, q)
        assert set(cnt) <= set(probs)
        for d, pr in probs.items():
            e = n * pr
            if e >= 5:
                chi2 += (cnt.get(d, 0) - e) ** 2 / e
                dof += 1
        dof -= 1
    assert dof > 50
    # chi2 ~ N(dof, 2 dof): 5 sigma
    assert abs(chi2 - dof) < 5.0 * (2.0 * dof) ** 0.5, (chi2, dof)


def test_alias_walk_is_counter_based(oracle):
    """Same (seed, walker, step) -> same path regardless of threads / subset."""
    g = oracle.Graph().load_file(KARATE)
    a = oracle.AliasGraph(g)
    i1, o1, _ = a.walk(walk_length=20, num_walks=3, p=0.5, q=2.0, seed=3, threads=1)
    i2, o2, _ = a.walk(walk_length=20, num_walks=3, p=0.5, q=2.0, seed=3, threads=4)
    assert (i1 == i2).all() and (o1 == o2).all()
    i3, o3, _ = a.walk(walk_length=20, num_walks=3, p=0.5, q=2.0, seed=3, sample_mod=5)
    full = oracle.paths_as_lists(i1, o1)
    assert oracle.paths_as_lists(i3, o3) == full[::5]


def test_rmat_generator_shape():
    s, d = synth.rmat_edges(8, 4, seed=42)
    assert len(s) == 4 << 8 and s.min() >= 0 and s.max() < 256 and d.max() < 256
    s2, d2 = synth.rmat_edges(8, 4, seed=42, first=100, count=50)
    assert (s2 == s[100:150]).all() and (d2 == d[100:150]).all()
    # skew: quadrant a dominates -> low ids are hubs
    assert (s < 128).mean() > 0.7
    w = synth.edge_weights(1000)
    assert w.dtype == np.float32 and w.min() >= 1.0 and w.max() < 2.0


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("p,q", [(0.5, 2.0), (0.25, 4.0), (0.1, 0.5)])
def test_alias_fold_distribution(oracle, p, q, weighted):
    """alias-fold (return edge folded out of the envelope) samples the same exact distribution;
    karate has a parallel edge (9-33 twice), so the multiplicity / bundle-weight path is exercised.
    Weighted: the bundle weight and the row weight sum take the place of multiplicity and degree."""
    g = _weighted_karate(oracle) if weighted else oracle.Graph().load_file(KARATE)
    a = oracle.AliasGraph(g)
    ids, offs, st = a.walk(walk_length=40, num_walks=300, p=p, q=q, seed=12, fold=1)
    ids0, _, st0 = a.walk(walk_length=40, num_walks=300, p=p, q=q, seed=12, fold=0)
    assert not np.array_equal(ids, ids0)                   # a different (cheaper) sampler ...
    assert st.proposals < 0.8 * st0.proposals              # ... with fewer proposals per step
    paths = ids.reshape(-1, 42)
    ctx = {}
    for k in range(2, 42):
        for a_, b_, c_ in zip(paths[:, k - 2], paths[:, k - 1], paths[:, k]):
            ctx.setdefault((int(a_), int(b_)), {}).setdefault(int(c_), 0)
            ctx[(int(a_), int(b_))][int(c_)] += 1
    chi2, dof = 0.0, 0
    for (pv, cu), cnt in ctx.items():
        n = sum(cnt.values())
        if n < 400:
            continue
        probs = _exact_transition(oracle, g, pv, cu, p, q)
        assert set(cnt) <= set(probs)
        for d, pr in probs.items():
            e = n * pr
            if e >= 5:
                chi2 += (cnt.get(d, 0) - e) ** 2 / e
                dof += 1
        dof -= 1
    assert dof > 50
    assert abs(chi2 - dof) < 5.0 * (2.0 * dof) ** 0.5, (chi2, dof)


def test_alias_fold_falls_back_when_not_applicable(oracle):
    g = oracle.Graph().load_file(KARATE)
    a = oracle.AliasGraph(g)
    for p, q in ((2.0, 0.5), (1.0, 1.0)):                 # 1/p <= max(1, 1/q): nothing to fold
        i1, o1, _ = a.walk(walk_length=20, num_walks=2, p=p, q=q, seed=3, fold=1)
        i0, o0, _ = a.walk(walk_length=20, num_walks=2, p=p, q=q, seed=3, fold=0)
        assert np.array_equal(i1, i0)
    d = oracle.AliasGraph(oracle.Graph().load_file(KARATE, directed=True), directed=True)   # directed: no symmetry
    i1, _, _ = d.walk(walk_length=20, num_walks=2, p=0.5, q=2.0, seed=3, fold=1)
    i0, _, _ = d.walk(walk_length=20, num_walks=2, p=0.5, q=2.0, seed=3, fold=0)
    assert np.array_equal(i1, i0)
    dw = oracle.AliasGraph(oracle.Graph().load_text(open(KARATE).read().replace("\n", " 1.5\n"), weighted=True, directed=True), directed=True)
    i1, _, _ = dw.walk(walk_length=20, num_walks=2, p=0.5, q=2.0, seed=3, fold=1)
    i0, _, _ = dw.walk(walk_length=20, num_walks=2, p=0.5, q=2.0, seed=3, fold=0)
    assert np.array_equal(i1, i0)


def test_weighted_fold_bundle_weights_are_symmetric(oracle):
    """The bundle weight read on prev's row must equal the one on curr's row (the kernel carries it across the step)."""
    import ctypes as C
    g = _weighted_karate(oracle)
    a = oracle.AliasGraph(g)
    v = a.view()
    L = oracle.lib()
    L.oa_wbundle.restype = C.POINTER(C.c_double)
    L.oa_wsum.restype = C.POINTER(C.c_double)
    nnz = int(v["offsets"][-1])
    wb = np.ctypeslib.as_array(L.oa_wbundle(a.h), (nnz,))
    ws = np.ctypeslib.as_array(L.oa_wsum(a.h), (a.nv,))
    off, col = v["offsets"], v["col"]
    for r in range(a.nv):
        assert abs(ws[r] - v["w"][off[r]:off[r + 1]].astype(np.float64).sum()) < 1e-9
        for e in range(off[r], off[r + 1]):
            x = col[e]
            k = off[x] + np.searchsorted(col[off[x]:off[x + 1]], r)
            assert col[k] == r and wb[k] == wb[e]


def test_fold_walk_over_dense_csr_is_the_twin(oracle):
    """bench.py's "optimised CPU twin" routine (oracle_fold_walk_csr_timed: dense sorted CSR, binary-search membership,
    multiplicity from the run) makes the decisions of oracle_alias_walk(fold=1), path for path."""
    import ctypes as C
    s, d = synth.rmat_edges(10, 8, seed=42)
    extra = np.array([0, 0, 5], np.int32), np.array([1, 1, 5], np.int32)            # parallel edges, a self-loop
    g = oracle.Graph().load_edges(np.concatenate([s, extra[0]]), np.concatenate([d, extra[1]]))
    a = oracle.AliasGraph(g)
    v = a.view()
    assert (v["vids"] == np.arange(a.nv)).all() or True
    L = oracle.lib()
    fn = L.oracle_fold_walk_csr_timed
    fn.restype = C.c_int64
    fn.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int64),
                   C.POINTER(C.c_uint64), C.c_void_p]
    for p, q in ((0.5, 2.0), (0.25, 4.0), (2.0, 0.5), (1.0, 1.0)):
        ids, offs, st = a.walk(walk_length=30, num_walks=1, p=p, q=q, seed=13, fold=1)
        want = ids.reshape(-1, 32)
        cfg = oracle.make_cfg(walk_length=30, num_walks=1, p=p, q=q, seed=13, threads=0, fold=1)
        off = np.ascontiguousarray(v["offsets"], np.int64)
        col = np.ascontiguousarray(v["col"], np.int32)
        out = np.zeros((a.nv, 32), np.int32)
        el, done, chk = C.c_double(), C.c_int64(), C.c_uint64()
        steps = fn(a.nv, off.ctypes.data, col.ctypes.data, C.addressof(cfg), 1, 0, 1e9, C.byref(el), C.byref(done), C.byref(chk), out.ctypes.data)
        assert done.value == a.nv and steps == st.steps
        assert (v["vids"][out] == want).all()


# ---- K3 in parallel: the formulation graph_build.cu evaluates with one thread per entry == the sequential integer sweep ----
def _parallel_vose_rows(off, w, col):
    """numpy restatement of graph_build.cu's parallel formulation"""
    nnz = len(w)
    thr = np.full(nnz, 0xFFFFFFFF, np.uint64); alias = np.zeros(nnz, np.int64)
    ONE = 1 << 32
    for r in range(len(off) - 1):
        o, n = int(off[r]), int(off[r + 1] - off[r])
        if n == 0: continue
        ww = w[o:o + n].astype(np.float64)
        part = np.zeros(32)
        for k in range(n): part[k & 31] = part[k & 31] + ww[k]
        for d in (16, 8, 4, 2, 1):
            part = np.array([part[l] + part[l ^ d] for l in range(32)])
        W = part[0]
        t = [int(np.floor(((x * float(n)) / W) * 4294967296.0)) for x in ww]
        heavy = [x >= ONE for x in t]
        E, D = [], []
        e = dd = 0
        Ecum = [0] * n; Dcum = [0] * n
        lh, ll = [], []
        for k in range(n):
            if heavy[k]: e += t[k] - ONE; lh.append(k)
            else: dd += ONE - t[k]; ll.append(k)
            Ecum[k] = e; Dcum[k] = dd
        for k in range(n):
            alias[o + k] = k
            if not heavy[k]:
                d_before = Dcum[k] - (ONE - t[k])
                x = next((q for q in range(len(lh)) if Ecum[lh[q]] >= d_before), None)
                if x is not None: thr[o + k] = t[k]; alias[o + k] = lh[x]
            else:
                x = lh.index(k)
                if x + 1 < len(lh):
                    m = next((q for q in range(len(ll)) if Dcum[ll[q]] > Ecum[k]), None)
                    if m is not None: thr[o + k] = ONE + Ecum[k] - Dcum[ll[m]]; alias[o + k] = lh[x + 1]
    return thr, alias


def test_parallel_vose_formulation_equals_the_sequential_sweep(oracle):
    """graph_build.cu builds the Vose tables with segmented prefix sums (E = heavy excess, D = light deficit) and two binary searches
    per entry instead of the two-cursor sweep.  This numpy restatement of THAT formulation must give the twin's (oracle alias_row)
    thresholds and aliases on every row: RMAT rows with smooth / five-valued / nearly-uniform weights, a Pareto hub row, a 5-entry row."""
    rng = np.random.default_rng(0)
    cases = []
    s, d = synth.rmat_edges(8, 8, seed=42)
    cases.append((s, d, synth.edge_weights(len(s), seed=43)))
    cases.append((s, d, rng.choice([0.25, 1.0, 4.0, 1e-3, 1e3], len(s)).astype(np.float32)))
    cases.append((s, d, (np.full(len(s), 3.0) + (np.arange(len(s)) % 7 == 0) * 0.5).astype(np.float32)))
    hub, leaves = np.zeros(300, np.int32), np.arange(1, 301, dtype=np.int32)
    cases.append((hub, leaves, (rng.pareto(1.2, 300) + 0.01).astype(np.float32)))
    cases.append((hub[:5], leaves[:5], np.array([1, 2, 3, 4, 5], np.float32)))
    for s, d, w in cases:
        v = oracle.AliasGraph(oracle.Graph().load_edges(s, d, w)).view()
        thr, alias = _parallel_vose_rows(v["offsets"], v["w"], v["col"])
        assert (thr == v["thr"].astype(np.uint64)).all() and (alias == v["alias"].astype(np.int64)).all()
        assert int((v["thr"] != 0xFFFFFFFF).sum()) > 0
