"""stellar-random-walk_b200 -- Python host mirror of the reference's walk-path interface over libsrw.

The product is the C-ABI library (include/srw.h, csrc/); this module is a thin ctypes binding that
keeps the reference's names so tests read like the reference's own:

    Params / CommandParser.parse        common/Params.scala:7-23, common/CommandParser.scala:32-109
    GraphMap                            algorithm/GraphMap.scala:11-121
    RandomSample                        algorithm/RandomSample.scala:5-63
    UniformRandomWalk / VCutRandomWalk  algorithm/{Uniform,VCut}RandomWalk.scala (loadGraph, randomWalk, execute, save)
    Main.main                           Main.scala:18-27

There is no CPU fallback: every graph / walk call goes to the CUDA library and raises SrwError
(SRW_ERR_NO_DEVICE) without a GPU; a missing libsrw.so raises at first use.
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsrw.so")

SRW_OK, SRW_ERR_ARG, SRW_ERR_USAGE, SRW_ERR_PARSE, SRW_ERR_IO, SRW_ERR_CUDA, SRW_ERR_NO_DEVICE, SRW_ERR_UNSUPPORTED = range(8)
TASK_NODE2VEC, TASK_RANDOMWALK, TASK_EMBEDDING = 0, 1, 2
SAMPLER_ALIAS, SAMPLER_EXACT, SAMPLER_ALIAS_FOLD = 0, 1, 2
U_PHILOX, U_CONST = 0, 1
BUILD_EXACT, BUILD_ALIAS, BUILD_ALL = 1, 2, 3
BUILD_MIGRATE, BUILD_LEAN = 4, 8      # srw.h

# every symbol include/srw.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "srw_last_error", "srw_version", "srw_device_count", "srw_params_default", "srw_params_parse_argv", "srw_usage",
    "srw_edges_parse_file", "srw_edges_parse_buffer", "srw_edges_view", "srw_edges_free",
    "srw_graph_from_edges", "srw_graph_from_device_edges", "srw_graph_load", "srw_graph_stats", "srw_graph_neighbors",
    "srw_graph_partition", "srw_graph_vertex_ids", "srw_graph_layout", "srw_graph_device_bytes", "srw_graph_build_profile", "srw_graph_free",
    "srw_graphmap_new", "srw_graphmap_add_vertex", "srw_graphmap_reset", "srw_graphmap_counts", "srw_graphmap_finalize",
    "srw_graphmap_free", "srw_sample", "srw_second_order_weights", "srw_second_order_sample", "srw_philox4x32_10",
    "srw_walk", "srw_walk_device", "srw_last_walk_info", "srw_walk_collect_stats", "srw_paths_view", "srw_paths_counts",
    "srw_save", "srw_paths_format", "srw_paths_free", "srw_main", "srw_synth_rmat_device", "srw_synth_weights_device",
    "srw_last_walk_kernel", "srw_edges_parse_buffer_device", "srw_paths_format_device", "srw_walk_save", "srw_walk_device_async", "srw_walk_wait",
    "srw_graph_from_device_edges_sharded", "srw_graph_shard_info", "srw_walker_msg_bytes", "srw_path_rec_bytes",
    "srw_shard_seed", "srw_shard_step", "srw_shard_apply", "srw_shard_finalize",
    "srw_shard_ipc_bytes", "srw_shard_ipc_export", "srw_shard_ipc_attach", "srw_shard_attach_local",
    "srw_shard_rows_info", "srw_shard_rows_relocate", "srw_shard_attach_block",
    "srw_mig_block_bytes", "srw_mig_create", "srw_mig_collect_stats", "srw_mig_begin", "srw_mig_superstep", "srw_mig_counters",
    "srw_mig_finish", "srw_mig_info", "srw_mig_free", "srw_graph_from_edges_multi",
    "srw_graph_from_device_edges_vcut", "srw_graph_from_edges_multi_vcut", "srw_graph_hub_info",
]


class SrwError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("srw status %d: %s" % (status, message))
        self.status = status


class CParams(C.Structure):
    """struct srw_params (include/srw.h)."""
    _fields_ = [("w2v_iter", C.c_int32), ("w2v_lr", C.c_double), ("w2v_partitions", C.c_int32), ("w2v_dim", C.c_int32),
                ("w2v_window", C.c_int32), ("walk_length", C.c_int32), ("num_walks", C.c_int32), ("p", C.c_double),
                ("q", C.c_double), ("weighted", C.c_int32), ("directed", C.c_int32), ("input", C.c_char * 1024),
                ("output", C.c_char * 1024), ("rdd_partitions", C.c_int32), ("single_output", C.c_int32),
                ("partitioned", C.c_int32), ("cmd", C.c_int32), ("seed", C.c_uint64), ("sampler", C.c_int32),
                ("u_mode", C.c_int32), ("u_const", C.c_float), ("num_gpus", C.c_int32)]


class WalkInfo(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("kernel_launches", C.c_int64), ("steps", C.c_int64), ("proposals", C.c_int64),
                ("member_tests", C.c_int64), ("probes_log2", C.c_int64)]


_lib = None


def lib():
    """Loads libsrw.so (fails loudly when the CUDA extension has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libsrw.so is missing: run `python stellar-random-walk_b200/build.py` (nvcc, sm_100a). "
                          "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    i32p, i64p, f32p, u32p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_uint32)
    L.srw_last_error.restype = C.c_char_p
    L.srw_version.restype = C.c_char_p
    L.srw_usage.restype = C.c_char_p
    L.srw_params_default.argtypes = [C.POINTER(CParams)]
    L.srw_params_parse_argv.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(CParams)]
    L.srw_edges_parse_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    L.srw_edges_parse_buffer.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(vp)]
    L.srw_edges_parse_buffer_device.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(vp)]
    L.srw_paths_format_device.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, C.c_int64, i64p, vp]
    L.srw_walk_save.argtypes = [vp, C.POINTER(CParams)]
    L.srw_edges_view.argtypes = [vp, i64p, C.POINTER(i32p), C.POINTER(i32p), C.POINTER(f32p), C.POINTER(i32p)]
    L.srw_edges_free.argtypes = [vp]
    L.srw_graph_from_edges.argtypes = [C.c_int64, vp, vp, vp, vp, C.c_int, C.c_uint, C.POINTER(vp)]
    L.srw_graph_from_device_edges.argtypes = [C.c_int64, vp, vp, vp, vp, C.c_int, C.c_uint, C.POINTER(vp)]
    L.srw_graph_load.argtypes = [C.POINTER(CParams), C.c_uint, C.POINTER(vp)]
    L.srw_graph_stats.argtypes = [vp, i64p, i64p]
    L.srw_graph_neighbors.argtypes = [vp, C.c_int32, vp, vp, C.c_int64, i64p]
    L.srw_graph_partition.argtypes = [vp, C.c_int32, i32p, C.POINTER(C.c_int)]
    L.srw_graph_vertex_ids.argtypes = [vp, vp, C.c_int64]
    L.srw_graph_layout.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int)]
    L.srw_graph_device_bytes.restype = C.c_int64
    L.srw_graph_device_bytes.argtypes = [vp]
    L.srw_graph_build_profile.restype = C.c_char_p
    L.srw_last_walk_kernel.restype = C.c_char_p
    L.srw_graph_build_profile.argtypes = [vp]
    L.srw_graph_free.argtypes = [vp]
    L.srw_graphmap_new.argtypes = [C.POINTER(vp)]
    L.srw_graphmap_add_vertex.argtypes = [vp, C.c_int32, C.c_int64, vp, vp, vp]
    L.srw_graphmap_reset.argtypes = [vp]
    L.srw_graphmap_counts.argtypes = [vp, i64p, i64p]
    L.srw_graphmap_finalize.argtypes = [vp, C.c_uint, C.POINTER(vp)]
    L.srw_graphmap_free.argtypes = [vp]
    L.srw_sample.argtypes = [C.c_int64, vp, vp, C.c_float, i32p, f32p]
    L.srw_second_order_weights.argtypes = [C.c_float, C.c_float, C.c_int32, C.c_int64, vp, C.c_int64, vp, vp, vp]
    L.srw_second_order_sample.argtypes = [C.c_float, C.c_float, C.c_int32, C.c_int64, vp, C.c_int64, vp, vp, C.c_float, i32p, f32p]
    L.srw_philox4x32_10.argtypes = [u32p, u32p, u32p]
    L.srw_walk.argtypes = [vp, C.POINTER(CParams), C.POINTER(vp)]
    L.srw_walk_device.argtypes = [vp, C.POINTER(CParams), C.c_uint64, C.c_int64, vp, vp, vp]
    L.srw_walk_device_async.argtypes = [vp, C.POINTER(CParams), C.c_uint64, C.c_int64, vp, vp, vp, C.POINTER(vp)]
    L.srw_walk_wait.argtypes = [vp, C.POINTER(WalkInfo)]
    L.srw_last_walk_info.argtypes = [C.POINTER(WalkInfo)]
    L.srw_walk_collect_stats.argtypes = [C.c_int]
    L.srw_paths_view.argtypes = [vp, i64p, C.POINTER(i32p), C.POINTER(i64p)]
    L.srw_paths_counts.argtypes = [vp, i64p, i64p]
    L.srw_save.argtypes = [vp, C.POINTER(CParams)]
    L.srw_paths_format.argtypes = [vp, C.c_char_p, C.c_int64, i64p]
    L.srw_paths_free.argtypes = [vp]
    L.srw_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.srw_synth_rmat_device.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int64, vp, vp]
    L.srw_synth_weights_device.argtypes = [C.c_uint64, C.c_int64, C.c_int64, vp]
    _lib = L
    return L


def check(status):
    if status != SRW_OK:
        raise SrwError(status, lib().srw_last_error().decode(errors="replace"))


def device_count():
    return int(lib().srw_device_count())


# ---------------------------------------------------------------------------------------------
# Params / CommandParser
# ---------------------------------------------------------------------------------------------
@dataclass
class Params:
    """common/Params.scala:7-23 (same names and defaults) + this build's additive options."""
    w2vIter: int = 10
    w2vLr: float = 0.025
    w2vPartitions: int = 1
    w2vDim: int = 128
    w2vWindow: int = 10
    walkLength: int = 80
    numWalks: int = 10
    p: float = 1.0
    q: float = 1.0
    weighted: bool = True
    directed: bool = False
    input: str = None
    output: str = None
    rddPartitions: int = 200
    singleOutput: bool = True
    partitioned: bool = False
    cmd: str = "node2vec"
    seed: int = 1
    sampler: str = "fold"
    gpus: int = 1

    _TASKS = ("node2vec", "randomwalk", "embedding")

    def to_c(self, u_const=None):
        c = CParams()
        lib().srw_params_default(C.byref(c))
        c.w2v_iter, c.w2v_lr, c.w2v_partitions, c.w2v_dim, c.w2v_window = self.w2vIter, self.w2vLr, self.w2vPartitions, self.w2vDim, self.w2vWindow
        c.walk_length, c.num_walks, c.p, c.q = self.walkLength, self.numWalks, self.p, self.q
        c.weighted, c.directed = int(self.weighted), int(self.directed)
        c.input = (self.input or "").encode()
        c.output = (self.output or "").encode()
        c.rdd_partitions, c.single_output, c.partitioned = self.rddPartitions, int(self.singleOutput), int(self.partitioned)
        c.cmd = self._TASKS.index(self.cmd)
        c.seed = self.seed
        c.sampler = {"exact": SAMPLER_EXACT, "fold": SAMPLER_ALIAS_FOLD}.get(self.sampler, SAMPLER_ALIAS)
        c.num_gpus = self.gpus
        if u_const is not None:
            c.u_mode, c.u_const, c.sampler = U_CONST, u_const, SAMPLER_EXACT
        return c

    @classmethod
    def from_c(cls, c):
        return cls(w2vIter=c.w2v_iter, w2vLr=c.w2v_lr, w2vPartitions=c.w2v_partitions, w2vDim=c.w2v_dim, w2vWindow=c.w2v_window,
                   walkLength=c.walk_length, numWalks=c.num_walks, p=c.p, q=c.q, weighted=bool(c.weighted),
                   directed=bool(c.directed), input=c.input.decode() or None, output=c.output.decode() or None,
                   rddPartitions=c.rdd_partitions, singleOutput=bool(c.single_output), partitioned=bool(c.partitioned),
                   cmd=cls._TASKS[c.cmd], seed=int(c.seed), sampler={SAMPLER_EXACT: "exact", SAMPLER_ALIAS_FOLD: "fold"}.get(c.sampler, "alias"),
                   gpus=c.num_gpus)


class CommandParser:
    """common/CommandParser.scala:107: parse(args) -> Some(Params) | None."""

    @staticmethod
    def parse(args):
        arr = (C.c_char_p * max(len(args), 1))(*[a.encode() for a in args])
        c = CParams()
        st = lib().srw_params_parse_argv(len(args), arr, C.byref(c))
        if st != SRW_OK:
            CommandParser.last_error = lib().srw_last_error().decode()
            return None
        return Params.from_c(c)

    @staticmethod
    def usage():
        return lib().srw_usage().decode()


# ---------------------------------------------------------------------------------------------
# edge lists, graphs, paths
# ---------------------------------------------------------------------------------------------
def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def parse_edges(data=None, path=None, weighted=True, partitioned=False, device=False):
    """URW:23-34 / VRW:19-34 line rules -> (src, dst, w, pid|None) numpy arrays.  device=False: the serial host
    parser (no GPU needed); device=True: the CUDA parser srw_graph_load uses (text_io.cu), same result."""
    h = C.c_void_p()
    if device:
        if path is not None:
            with open(path, "rb") as f:
                data = f.read()
        if isinstance(data, str):
            data = data.encode()
        check(lib().srw_edges_parse_buffer_device(data, len(data), int(weighted), int(partitioned), C.byref(h)))
    elif path is not None:
        check(lib().srw_edges_parse_file(path.encode(), int(weighted), int(partitioned), C.byref(h)))
    else:
        if isinstance(data, str):
            data = data.encode()
        check(lib().srw_edges_parse_buffer(data, len(data), int(weighted), int(partitioned), C.byref(h)))
    try:
        n = C.c_int64()
        s, d, p = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        w = C.POINTER(C.c_float)()
        check(lib().srw_edges_view(h, C.byref(n), C.byref(s), C.byref(d), C.byref(w), C.byref(p)))
        n = n.value
        if n == 0:
            return np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32), (np.zeros(0, np.int32) if partitioned else None)
        return (np.ctypeslib.as_array(s, (n,)).copy(), np.ctypeslib.as_array(d, (n,)).copy(),
                np.ctypeslib.as_array(w, (n,)).copy(), np.ctypeslib.as_array(p, (n,)).copy() if p else None)
    finally:
        lib().srw_edges_free(h)


class Graph:
    """Device-resident adjacency (srw_graph)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_edges(cls, src, dst, w=None, pid=None, directed=False, flags=BUILD_ALL):
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        w = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
        pid = None if pid is None else np.ascontiguousarray(pid, dtype=np.int32)
        h = C.c_void_p()
        check(lib().srw_graph_from_edges(len(src), _ptr(src), _ptr(dst), _ptr(w), _ptr(pid), int(directed), flags, C.byref(h)))
        return cls(h)

    @classmethod
    def from_edges_multi(cls, src, dst, num_gpus, directed=False):
        """One vertex-range shard per GPU of this process inside one handle (srw_graph_from_edges_multi): walk it with
        Params(gpus=num_gpus)."""
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        h = C.c_void_p()
        L = lib()
        L.srw_graph_from_edges_multi.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        check(L.srw_graph_from_edges_multi(len(src), _ptr(src), _ptr(dst), int(directed), num_gpus, C.byref(h)))
        return cls(h)

    @classmethod
    def from_edges_multi_vcut(cls, src, dst, pid, num_gpus, directed=False):
        """The same with the partition-id column as the shard map (VCutRandomWalk on N GPUs: owner(v) = getPartition(v) mod N,
        srw_graph_from_edges_multi_vcut)."""
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        pid = np.ascontiguousarray(pid, dtype=np.int32)
        h = C.c_void_p()
        L = lib()
        L.srw_graph_from_edges_multi_vcut.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        check(L.srw_graph_from_edges_multi_vcut(len(src), _ptr(src), _ptr(dst), _ptr(pid), int(directed), num_gpus, C.byref(h)))
        return cls(h)

    @classmethod
    def from_device_edges(cls, n, d_src, d_dst, d_w=None, directed=False, flags=BUILD_ALIAS):
        h = C.c_void_p()
        check(lib().srw_graph_from_device_edges(n, d_src, d_dst, d_w, None, int(directed), flags, C.byref(h)))
        return cls(h)

    @classmethod
    def load(cls, params, flags=BUILD_ALL):
        h = C.c_void_p()
        cp = params.to_c()
        check(lib().srw_graph_load(C.byref(cp), flags, C.byref(h)))
        return cls(h)

    def free(self):
        if self.h:
            lib().srw_graph_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def stats(self):
        nv, ne = C.c_int64(), C.c_int64()
        check(lib().srw_graph_stats(self.h, C.byref(nv), C.byref(ne)))
        return nv.value, ne.value

    @property
    def num_vertices(self):
        return self.stats()[0]

    @property
    def num_edges(self):
        return self.stats()[1]

    def neighbors(self, vid):
        """GM:109-120: None for an unknown vid, else [(dst, w)] in file-appearance order."""
        n = C.c_int64()
        check(lib().srw_graph_neighbors(self.h, vid, None, None, 0, C.byref(n)))
        if n.value < 0:
            return None
        d = np.zeros(max(n.value, 1), np.int32)
        w = np.zeros(max(n.value, 1), np.float32)
        check(lib().srw_graph_neighbors(self.h, vid, _ptr(d), _ptr(w), n.value, C.byref(n)))
        return [(int(d[i]), float(w[i])) for i in range(n.value)]

    def partition(self, vid):
        pid, found = C.c_int32(), C.c_int()
        check(lib().srw_graph_partition(self.h, vid, C.byref(pid), C.byref(found)))
        return pid.value if found.value else None

    def vertex_ids(self):
        nv = self.num_vertices
        out = np.zeros(max(nv, 1), np.int32)
        check(lib().srw_graph_vertex_ids(self.h, _ptr(out), nv))
        return out[:nv]

    def layout(self):
        nv, nnz = self.stats()
        has = C.c_int()
        check(lib().srw_graph_layout(self.h, None, None, None, C.byref(has)))
        off = np.zeros(nv + 1, np.int64)
        col = np.zeros(max(nnz, 1), np.int32)
        slots = np.zeros((max(nnz, 1), 4), np.uint32) if has.value else None
        check(lib().srw_graph_layout(self.h, _ptr(off), _ptr(col), _ptr(slots), C.byref(has)))
        out = {"offsets": off, "col": col[:nnz], "has_alias": bool(has.value)}
        if slots is not None:
            out["thr"], out["own"], out["alias_vertex"], out["alias"] = (slots[:nnz, k].copy() for k in range(4))
        return out

    def walk(self, params, u_const=None):
        """RW:75-176 for all rounds; returns (ids, offsets) ragged numpy arrays."""
        cp = params.to_c(u_const)
        h = C.c_void_p()
        check(lib().srw_walk(self.h, C.byref(cp), C.byref(h)))
        return Paths(h)

    def walk_save(self, params):
        """execute() + save() streamed through the device formatter (srw_walk_save): writes <output>/path/part-NNNNN."""
        cp = params.to_c()
        check(lib().srw_walk_save(self.h, C.byref(cp)))
        return last_walk_info()


def format_paths_device(d_paths, d_lens, n_paths, stride, d_text=None, cap=0, stream=None):
    """RW:234-241 on the device (srw_paths_format_device).  Pointers are raw device addresses (e.g. tensor.data_ptr()).
    Returns the number of text bytes; with d_text=None only sizes the buffer."""
    need = C.c_int64()
    check(lib().srw_paths_format_device(d_paths, d_lens, n_paths, stride, d_text, cap, C.byref(need), stream))
    return need.value


def last_walk_info():
    wi = WalkInfo()
    check(lib().srw_last_walk_info(C.byref(wi)))
    return wi


class Paths:
    """RDD[Array[Int]] replacement (srw_paths)."""

    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        try:
            if self.h:
                lib().srw_paths_free(self.h)
                self.h = None
        except Exception:
            pass

    def arrays(self):
        n = C.c_int64()
        ids, offs = C.POINTER(C.c_int32)(), C.POINTER(C.c_int64)()
        check(lib().srw_paths_view(self.h, C.byref(n), C.byref(ids), C.byref(offs)))
        o = np.ctypeslib.as_array(offs, (n.value + 1,)).copy()
        tot = int(o[-1])
        i = np.ctypeslib.as_array(ids, (tot,)).copy() if tot > 0 else np.zeros(0, np.int32)
        return i, o

    def collect(self):
        ids, offs = self.arrays()
        return [ids[offs[k]:offs[k + 1]].tolist() for k in range(len(offs) - 1)]

    def count(self):
        n, s = C.c_int64(), C.c_int64()
        check(lib().srw_paths_counts(self.h, C.byref(n), C.byref(s)))
        return n.value

    def steps(self):
        n, s = C.c_int64(), C.c_int64()
        check(lib().srw_paths_counts(self.h, C.byref(n), C.byref(s)))
        return s.value

    def format(self):
        need = C.c_int64()
        check(lib().srw_paths_format(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value + 1)
        check(lib().srw_paths_format(self.h, buf, need.value, C.byref(need)))
        return buf.raw[:need.value]

    def save(self, params):
        cp = params.to_c()
        check(lib().srw_save(self.h, C.byref(cp)))


# ---------------------------------------------------------------------------------------------
# reference-named classes
# ---------------------------------------------------------------------------------------------
class GraphMap:
    """algorithm/GraphMap.scala:11-121 over srw_graphmap (builder) + srw_graph (device rows)."""

    def __init__(self):
        self._m = C.c_void_p()
        check(lib().srw_graphmap_new(C.byref(self._m)))
        self._g = None
        self._added = set()

    def __del__(self):
        try:
            lib().srw_graphmap_free(self._m)
        except Exception:
            pass

    def addVertex(self, vId, neighbors=None):
        """neighbors: [(dst, w)] (GM:41), [(dst, pid, w)] (GM:23) or None/[] (GM:83)."""
        nb = list(neighbors or [])
        d = np.array([e[0] for e in nb], dtype=np.int32)
        w = np.array([e[-1] for e in nb], dtype=np.float32)
        pid = np.array([e[1] for e in nb], dtype=np.int32) if nb and len(nb[0]) == 3 else None
        check(lib().srw_graphmap_add_vertex(self._m, vId, len(nb), _ptr(d), _ptr(pid), _ptr(w)))
        self._added.add(vId)
        self._g = None

    def _graph(self):
        if self._g is None:
            h = C.c_void_p()
            check(lib().srw_graphmap_finalize(self._m, BUILD_ALL, C.byref(h)))
            self._g = Graph(h)
        return self._g

    def getNeighbors(self, vid):
        if vid not in self._added:
            return None                      # GM:118 case None => null
        return self._graph().neighbors(vid)

    def getPartition(self, vId):
        return self._graph().partition(vId)

    @property
    def getNumVertices(self):
        nv, ne = C.c_int64(), C.c_int64()
        check(lib().srw_graphmap_counts(self._m, C.byref(nv), C.byref(ne)))
        return nv.value

    @property
    def getNumEdges(self):
        nv, ne = C.c_int64(), C.c_int64()
        check(lib().srw_graphmap_counts(self._m, C.byref(nv), C.byref(ne)))
        return ne.value

    def reset(self):
        check(lib().srw_graphmap_reset(self._m))
        self._g = None
        self._added = set()


class RandomSample:
    """algorithm/RandomSample.scala:5-63; nextFloat is a callable returning the draw (RS:5)."""

    def __init__(self, nextFloat=None):
        self.nextFloat = nextFloat

    def _u(self):
        if self.nextFloat is None:
            raise SrwError(SRW_ERR_ARG, "RandomSample needs an injected nextFloat (the walk itself uses Philox on the device)")
        return float(self.nextFloat())

    def sample(self, edges):
        d = np.array([e[0] for e in edges], dtype=np.int32)
        w = np.array([e[1] for e in edges], dtype=np.float32)
        do, wo = C.c_int32(), C.c_float()
        check(lib().srw_sample(len(d), _ptr(d), _ptr(w), self._u(), C.byref(do), C.byref(wo)))
        return (do.value, wo.value)

    def computeSecondOrderWeights(self, p=1.0, q=1.0, prevId=0, prevNeighbors=(), currNeighbors=()):
        pd = np.array([e[0] for e in prevNeighbors], dtype=np.int32)
        cd = np.array([e[0] for e in currNeighbors], dtype=np.int32)
        cw = np.array([e[1] for e in currNeighbors], dtype=np.float32)
        out = np.zeros(max(len(cd), 1), np.float32)
        check(lib().srw_second_order_weights(p, q, prevId, len(pd), _ptr(pd), len(cd), _ptr(cd), _ptr(cw), _ptr(out)))
        return [(int(cd[i]), float(out[i])) for i in range(len(cd))]

    def secondOrderSample(self, p=1.0, q=1.0, prevId=0, prevNeighbors=(), currNeighbors=()):
        pd = np.array([e[0] for e in prevNeighbors], dtype=np.int32)
        cd = np.array([e[0] for e in currNeighbors], dtype=np.int32)
        cw = np.array([e[1] for e in currNeighbors], dtype=np.float32)
        do, wo = C.c_int32(), C.c_float()
        check(lib().srw_second_order_sample(p, q, prevId, len(pd), _ptr(pd), len(cd), _ptr(cd), _ptr(cw), self._u(),
                                            C.byref(do), C.byref(wo)))
        return (do.value, wo.value)


class _RandomWalk:
    """trait RandomWalk (RW:12-242).  `context` (SparkContext) has no counterpart."""
    partitioned = False

    def __init__(self, config, flags=BUILD_ALL):
        self.config = config
        self.flags = flags
        self.graph = None
        self.nVertices = 0
        self.nEdges = 0

    def loadGraph(self):
        """URW:17-88 / VRW:13-98: returns the (vid, [vid]) start paths like the reference."""
        cfg = self.config
        src, dst, w, pid = parse_edges(path=cfg.input, weighted=cfg.weighted, partitioned=self.partitioned)
        self.graph = Graph.from_edges(src, dst, w, pid, cfg.directed, self.flags)
        self.nVertices, self.nEdges = self.graph.stats()
        return [(int(v), [int(v)]) for v in self.graph.vertex_ids()]

    def _const_u(self, nextFloat):
        return None if nextFloat is None else float(nextFloat())

    def initFirstStep(self, paths=None, nextFloat=None):
        """RW:51-66: [v, first neighbour] or [v] on a dead end."""
        cfg = Params(**{**self.config.__dict__, "walkLength": 0, "numWalks": 1, "sampler": "exact"})
        res = self.graph.walk(cfg, self._const_u(nextFloat)).collect()
        return [(p[0], (p, len(p) == 1)) for p in res]

    def randomWalk(self, initPaths=None, nextFloat=None):
        """RW:75-176.  nextFloat: the constant generator of the reference's tests, or None for Philox."""
        return self.graph.walk(self.config, self._const_u(nextFloat))

    def execute(self):
        return self.randomWalk(self.loadGraph())

    def save(self, paths, partitions, output):
        cfg = Params(**{**self.config.__dict__, "output": output, "singleOutput": partitions == 1, "rddPartitions": partitions})
        paths.save(cfg)


class UniformRandomWalk(_RandomWalk):
    partitioned = False


class VCutRandomWalk(_RandomWalk):
    partitioned = True


class Main:
    @staticmethod
    def main(args):
        """Main.scala:18-27 through the native driver; returns the exit code."""
        arr = (C.c_char_p * max(len(args), 1))(*[a.encode() for a in args])
        return int(lib().srw_main(len(args), arr))

    @staticmethod
    def doRandomWalk(param):
        """Main.scala:53-62."""
        rw = VCutRandomWalk(param) if param.partitioned else UniformRandomWalk(param)
        paths = rw.execute()
        rw.save(paths, 1 if param.singleOutput else param.rddPartitions, param.output)
        return paths
