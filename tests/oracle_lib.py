"""ctypes binding of the CPU oracle (oracle/libsrw_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ODIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ODIR, "libsrw_oracle.so")


def build(force=False):
    src = [os.path.join(_ODIR, f) for f in ("srw_oracle.c", "srw_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _ODIR, "-s"] + (["-B"] if force else []))
    return _SO


class WalkCfg(C.Structure):
    _fields_ = [("walk_length", C.c_int32), ("num_walks", C.c_int32), ("p", C.c_double), ("q", C.c_double),
                ("u_mode", C.c_int32), ("u_const", C.c_float), ("seed", C.c_uint64), ("threads", C.c_int32),
                ("sample_mod", C.c_int64), ("fold", C.c_int32)]


class AliasStats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("proposals", C.c_int64), ("probes_log2", C.c_int64),
                ("member_tests", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, i32p, i64p, f32p, u32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    L.oracle_philox4x32_10.argtypes = [u32p, u32p, u32p]
    L.oracle_u01.restype = C.c_float
    L.oracle_u01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    L.og_new.restype = vp
    L.og_free.argtypes = [vp]
    L.og_reset.argtypes = [vp]
    L.og_add_vertex.argtypes = [vp, C.c_int32, i32p, f32p, C.c_int64]
    L.og_add_vertex_pid.argtypes = [vp, C.c_int32, i32p, i32p, f32p, C.c_int64]
    L.og_add_vertex_empty.argtypes = [vp, C.c_int32]
    L.og_neighbors.restype = C.c_int64
    L.og_neighbors.argtypes = [vp, C.c_int32, C.POINTER(i32p), C.POINTER(f32p)]
    L.og_partition.argtypes = [vp, C.c_int32, i32p]
    L.og_num_vertices.restype = C.c_int64
    L.og_num_vertices.argtypes = [vp]
    L.og_num_edges.restype = C.c_int64
    L.og_num_edges.argtypes = [vp]
    L.og_vertex_ids.restype = C.c_int64
    L.og_vertex_ids.argtypes = [vp, i32p, C.c_int64]
    L.og_load_text.restype = C.c_int64
    L.og_load_text.argtypes = [vp, C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
    L.og_load_edges.argtypes = [vp, C.c_int64, i32p, i32p, f32p, i32p, C.c_int]
    L.oracle_sample.restype = C.c_int64
    L.oracle_sample.argtypes = [C.c_int64, f32p, C.c_float]
    L.oracle_second_order_weights.argtypes = [C.c_float, C.c_float, C.c_int32, C.c_int64, i32p, C.c_int64, i32p, f32p, f32p]
    L.oracle_second_order_sample.restype = C.c_int64
    L.oracle_second_order_sample.argtypes = [C.c_float, C.c_float, C.c_int32, C.c_int64, i32p, C.c_int64, i32p, f32p, C.c_float, f32p]
    L.oracle_walk.restype = C.c_int64
    L.oracle_walk.argtypes = [vp, C.POINTER(WalkCfg), i32p, C.c_int64, i64p]
    L.oracle_format_paths.restype = C.c_int64
    L.oracle_format_paths.argtypes = [C.c_int64, i32p, i64p, C.c_char_p, C.c_int64]
    L.oa_build.restype = vp
    L.oa_build.argtypes = [vp]
    L.oa_free.argtypes = [vp]
    L.oa_num_vertices.restype = C.c_int64
    L.oa_num_vertices.argtypes = [vp]
    L.oa_has_alias.argtypes = [vp]
    L.oa_view.argtypes = [vp, C.POINTER(i32p), C.POINTER(i64p), C.POINTER(i32p), C.POINTER(f32p), C.POINTER(u32p), C.POINTER(u32p)]
    L.oracle_alias_walk.restype = C.c_int64
    L.oracle_alias_walk.argtypes = [vp, C.POINTER(WalkCfg), i32p, C.c_int64, i64p, C.POINTER(AliasStats)]
    L.oracle_alias_thresholds.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(o, C.c_uint32))
    return o


def u01(seed, walker, step):
    return float(lib().oracle_u01(seed, walker, step))


class Graph:
    """GraphMap restatement handle."""

    def __init__(self):
        self.h = lib().og_new()

    def __del__(self):
        try:
            lib().og_free(self.h)
        except Exception:
            pass

    def reset(self):
        lib().og_reset(self.h)

    def add_vertex(self, vid, neighbors=None, pids=None):
        if not neighbors:
            if neighbors is None:
                lib().og_add_vertex_empty(self.h, vid)
            else:
                lib().og_add_vertex(self.h, vid, None, None, 0)
            return
        d = np.array([n[0] for n in neighbors], dtype=np.int32)
        w = np.array([n[-1] for n in neighbors], dtype=np.float32)
        if pids is None and len(neighbors[0]) == 2:
            lib().og_add_vertex(self.h, vid, _p(d, C.c_int32), _p(w, C.c_float), len(d))
        else:
            pp = np.array(pids if pids is not None else [n[1] for n in neighbors], dtype=np.int32)
            lib().og_add_vertex_pid(self.h, vid, _p(d, C.c_int32), _p(pp, C.c_int32), _p(w, C.c_float), len(d))

    def neighbors(self, vid):
        """None for an unknown vid (reference: null), else list of (dst, w)."""
        dp, wp = C.POINTER(C.c_int32)(), C.POINTER(C.c_float)()
        n = lib().og_neighbors(self.h, vid, C.byref(dp), C.byref(wp))
        if n < 0:
            return None
        return [(int(dp[i]), float(wp[i])) for i in range(n)]

    def neighbors_np(self, vid):
        dp, wp = C.POINTER(C.c_int32)(), C.POINTER(C.c_float)()
        n = lib().og_neighbors(self.h, vid, C.byref(dp), C.byref(wp))
        if n < 0:
            return None
        if n == 0:
            return np.zeros(0, np.int32), np.zeros(0, np.float32)
        return np.ctypeslib.as_array(dp, (n,)).copy(), np.ctypeslib.as_array(wp, (n,)).copy()

    def partition(self, vid):
        pid = C.c_int32(0)
        return int(pid.value) if lib().og_partition(self.h, vid, C.byref(pid)) else None

    @property
    def num_vertices(self):
        return int(lib().og_num_vertices(self.h))

    @property
    def num_edges(self):
        return int(lib().og_num_edges(self.h))

    def vertex_ids(self):
        n = self.num_vertices
        out = np.zeros(max(n, 1), dtype=np.int32)
        lib().og_vertex_ids(self.h, _p(out, C.c_int32), n)
        return out[:n]

    def load_text(self, data, weighted=True, directed=False, partitioned=False):
        if isinstance(data, str):
            data = data.encode()
        err = C.create_string_buffer(256)
        bad = lib().og_load_text(self.h, data, len(data), int(weighted), int(directed), int(partitioned), err, 256)
        if bad:
            raise ValueError(err.value.decode())
        return self

    def load_file(self, path, **kw):
        with open(path, "rb") as f:
            return self.load_text(f.read(), **kw)

    def load_edges(self, src, dst, w=None, pid=None, directed=False):
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
        pid = None if pid is None else np.ascontiguousarray(pid, dtype=np.int32)
        lib().og_load_edges(self.h, len(src), _p(src, C.c_int32), _p(dst, C.c_int32), _p(w, C.c_float),
                            _p(pid, C.c_int32), int(directed))
        return self


def sample(edges, u):
    """RS:12-25 on a list of (dst, w); returns the chosen (dst, w)."""
    w = np.array([e[1] for e in edges], dtype=np.float32)
    k = lib().oracle_sample(len(w), _p(w, C.c_float), C.c_float(u))
    return edges[k]


def second_order_weights(p, q, prev, prev_neighbors, curr_neighbors):
    pd = np.array([e[0] for e in prev_neighbors], dtype=np.int32)
    cd = np.array([e[0] for e in curr_neighbors], dtype=np.int32)
    cw = np.array([e[1] for e in curr_neighbors], dtype=np.float32)
    out = np.zeros(len(cd), dtype=np.float32)
    lib().oracle_second_order_weights(p, q, prev, len(pd), _p(pd, C.c_int32), len(cd), _p(cd, C.c_int32),
                                      _p(cw, C.c_float), _p(out, C.c_float))
    return [(int(cd[i]), float(out[i])) for i in range(len(cd))]


def second_order_sample(p, q, prev, prev_neighbors, curr_neighbors, u):
    pd = np.array([e[0] for e in prev_neighbors], dtype=np.int32)
    cd = np.array([e[0] for e in curr_neighbors], dtype=np.int32)
    cw = np.array([e[1] for e in curr_neighbors], dtype=np.float32)
    wo = C.c_float(0)
    k = lib().oracle_second_order_sample(p, q, prev, len(pd), _p(pd, C.c_int32), len(cd), _p(cd, C.c_int32),
                                         _p(cw, C.c_float), C.c_float(u), C.byref(wo))
    return (int(cd[k]), float(wo.value))


def make_cfg(walk_length=80, num_walks=10, p=1.0, q=1.0, u_const=None, seed=1, threads=0, sample_mod=0, fold=0):
    return WalkCfg(walk_length, num_walks, p, q, 0 if u_const is not None else 1,
                   0.0 if u_const is None else u_const, seed, threads, sample_mod, fold)


def _run_walk(fn, handle, cfg, n_vertices, extra=()):
    n_paths_max = cfg.num_walks * n_vertices
    cap = max(1, n_paths_max * (cfg.walk_length + 2))
    ids = np.zeros(cap, dtype=np.int32)
    offs = np.zeros(n_paths_max + 1, dtype=np.int64)
    n = fn(handle, C.byref(cfg), _p(ids, C.c_int32), cap, _p(offs, C.c_int64), *extra)
    assert n >= 0
    offs = offs[:n + 1]
    return ids[:offs[-1]], offs


def walk(graph, **kw):
    """Reference-algorithm walk.  Returns (ids, offsets) in (round, ascending vid) order."""
    cfg = make_cfg(**kw)
    return _run_walk(lib().oracle_walk, graph.h, cfg, graph.num_vertices)


def paths_as_lists(ids, offs):
    return [ids[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]


def format_paths(ids, offs):
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    n = len(offs) - 1
    need = lib().oracle_format_paths(n, _p(ids, C.c_int32), _p(offs, C.c_int64), None, 0)
    buf = C.create_string_buffer(int(need) + 1)
    lib().oracle_format_paths(n, _p(ids, C.c_int32), _p(offs, C.c_int64), buf, need)
    return buf.raw[:need]


class AliasGraph:
    """CPU twin of the product's alias-mode layout (sorted CSR + Vose tables)."""

    def __init__(self, graph, directed=False):
        self.h = lib().oa_build(graph.h)
        self.nv = int(lib().oa_num_vertices(self.h))
        lib().oa_set_directed.argtypes = [C.c_void_p, C.c_int]
        lib().oa_set_directed(self.h, int(directed))

    def __del__(self):
        try:
            lib().oa_free(self.h)
        except Exception:
            pass

    @property
    def has_alias(self):
        return bool(lib().oa_has_alias(self.h))

    def view(self):
        vids, offs, col = C.POINTER(C.c_int32)(), C.POINTER(C.c_int64)(), C.POINTER(C.c_int32)()
        w, thr, al = C.POINTER(C.c_float)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
        lib().oa_view(self.h, C.byref(vids), C.byref(offs), C.byref(col), C.byref(w), C.byref(thr), C.byref(al))
        nv = self.nv
        o = np.ctypeslib.as_array(offs, (nv + 1,)).copy()
        nnz = int(o[-1])
        d = {"vids": np.ctypeslib.as_array(vids, (max(nv, 1),))[:nv].copy(), "offsets": o}
        d["col"] = np.ctypeslib.as_array(col, (max(nnz, 1),))[:nnz].copy()
        d["w"] = np.ctypeslib.as_array(w, (max(nnz, 1),))[:nnz].copy()
        if self.has_alias:
            d["thr"] = np.ctypeslib.as_array(thr, (max(nnz, 1),))[:nnz].copy()
            d["alias"] = np.ctypeslib.as_array(al, (max(nnz, 1),))[:nnz].copy()
        return d

    def walk(self, **kw):
        cfg = make_cfg(**kw)
        st = AliasStats()
        ids, offs = _run_walk(lib().oracle_alias_walk, self.h, cfg, self.nv, (C.byref(st),))
        return ids, offs, st


def alias_thresholds(p, q):
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib().oracle_alias_thresholds(p, q, C.byref(a), C.byref(b), C.byref(c))
    return int(a.value), int(b.value), int(c.value)
