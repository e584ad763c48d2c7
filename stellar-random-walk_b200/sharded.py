"""Host side of the vertex-range-sharded walk: the super-step loop and the walker
exchange.  Replaces RandomWalk.randomWalk's do/while over `transferWalkersToTheirPartitions`
(RW:91-162, RW:186-192): each super-step every rank advances its resident walkers on the device
(csrc/shard.cu), then the 32-byte walker tuples and 16-byte path records are exchanged with an
all-to-all (NCCL over NVLink when ranks are processes) and the loop ends when no rank sent anything
(the reference's `remainingWalkers != 0`, RW:162).

Two deployments share this code:
  * one process per GPU (torchrun): `DistExchange` -- torch.distributed.all_to_all_single;
  * several shards inside one process on one device (tests): `LocalExchange` -- tensor slicing.
"""
import ctypes as C

import numpy as np
import torch

from . import BUILD_ALIAS, check, lib

MSG_BYTES = 32
REC_BYTES = 16


def _bind():
    L = lib()
    if getattr(L, "_shard_bound", False):
        return L
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    L.srw_graph_from_device_edges_sharded.argtypes = [C.c_int64, vp, vp, vp, C.c_int, C.c_uint, C.c_int, C.c_int, C.POINTER(vp)]
    L.srw_graph_shard_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), i64p, i64p, i64p, i64p]
    L.srw_shard_seed.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, i64p, vp, vp, vp]
    L.srw_shard_step.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, vp, vp, i64p, i64p, i64p, vp]
    L.srw_shard_apply.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp]
    L.srw_shard_finalize.argtypes = [vp, vp, C.c_int64, vp, vp, i64p, vp]
    assert L.srw_walker_msg_bytes() == MSG_BYTES and L.srw_path_rec_bytes() == REC_BYTES
    L._shard_bound = True
    return L


def plan_bounds(degree_prefix, world):
    """Edge-balanced contiguous vertex ranges: bounds[r] = first rank whose degree prefix reaches
    nnz * r / world (same rule as csrc/graph_build.cu; host twin for tests)."""
    nv = len(degree_prefix) - 1
    nnz = int(degree_prefix[-1])
    b = [0]
    for r in range(1, world):
        t = nnz * r // world
        k = int(np.searchsorted(degree_prefix, t, side="left"))
        b.append(max(min(k, nv), b[-1]))
    b.append(nv)
    return b


def owner_of(bounds, v):
    o = 0
    while o + 1 < len(bounds) - 1 and v >= bounds[o + 1]:
        o += 1
    return o


class Shard:
    """One rank's rows of the graph plus its walker pools (device buffers owned here)."""

    def __init__(self, n_edges, d_src, d_dst, d_w, rank, world, directed=False, device=None):
        L = _bind()
        self.h = C.c_void_p()
        check(L.srw_graph_from_device_edges_sharded(n_edges, d_src, d_dst, d_w, int(directed), BUILD_ALIAS, rank, world, C.byref(self.h)))
        r, w = C.c_int(), C.c_int()
        rf, rl, nl = C.c_int64(), C.c_int64(), C.c_int64()
        b = (C.c_int64 * (world + 1))()
        check(L.srw_graph_shard_info(self.h, C.byref(r), C.byref(w), C.byref(rf), C.byref(rl), b, C.byref(nl)))
        self.rank, self.world, self.row_first, self.row_last, self.nnz_local = r.value, w.value, rf.value, rl.value, nl.value
        self.bounds = list(b)
        nv, _ = C.c_int64(), C.c_int64()
        check(L.srw_graph_stats(self.h, C.byref(nv), None))
        self.nv = nv.value
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())

    def free(self):
        if self.h:
            lib().srw_graph_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def rows(self):
        return self.row_last - self.row_first


class LocalExchange:
    """All shards live in this process: route segments by slicing (tests, single device)."""

    def __init__(self, world):
        self.world = world

    def exchange(self, sends, counts, item_bytes):
        """sends[r]: uint8 tensor of rank r's send buffer (segments in destination order);
        counts[r][d]: items rank r sends to d.  Returns (recv tensors, recv totals) per rank."""
        out, tot = [], []
        for d in range(self.world):
            parts = []
            for r in range(self.world):
                first = sum(counts[r][:d]) * item_bytes
                parts.append(sends[r][first:first + counts[r][d] * item_bytes])
            out.append(torch.cat(parts) if parts else sends[d][:0])
            tot.append(sum(counts[r][d] for r in range(self.world)))
        return out, tot

    def total(self, values):
        return sum(values)


class DistExchange:
    """One shard per process: all-to-all of the counts, then of the payload (NCCL over NVLink on
    GPUs; the same calls run over gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)

    def exchange(self, sends, counts, item_bytes):
        dist = self.dist
        send, cnt = sends[0], counts[0]
        dev = send.device
        c_out = torch.tensor(cnt, dtype=torch.int64, device=dev)
        c_in = torch.empty_like(c_out)
        dist.all_to_all_single(c_in, c_out, group=self.group)
        c_in_l = [int(v) for v in c_in.tolist()]
        recv = torch.empty(sum(c_in_l) * item_bytes, dtype=torch.uint8, device=dev)
        dist.all_to_all_single(recv, send[:sum(cnt) * item_bytes], output_split_sizes=[c * item_bytes for c in c_in_l],
                               input_split_sizes=[c * item_bytes for c in cnt], group=self.group)
        return [recv], [sum(c_in_l)]

    def total(self, values):
        t = torch.tensor([sum(values)], dtype=torch.int64, device=self._dev)
        self.dist.all_reduce(t, group=self.group)
        return int(t.item())

    _dev = torch.device("cpu")


def run_sharded(shards, params, round_first=0, n_rounds=1, exchange=None, rec_cap=1 << 22, stream=None, inbox_cap=None):
    """Walks rounds [round_first, round_first + n_rounds) over `shards` (all W shards of the graph in
    this process with LocalExchange, or this process's single shard with DistExchange).

    Returns per local shard: (paths [rows * n_rounds, walkLength + 2] int32 vertex ids, lens int32),
    plus a stats dict (super-steps, tuples and records exchanged, sampled transitions)."""
    L = _bind()
    world = shards[0].world
    if exchange is None:
        exchange = LocalExchange(world) if len(shards) == world else DistExchange()
    if isinstance(exchange, DistExchange):
        exchange._dev = shards[0].device
    cp = params.to_c()
    stride = params.walkLength + 2
    st = 0 if stream is None else stream
    n_total = shards[0].nv * n_rounds
    state = []
    for s in shards:
        n_home = s.rows * n_rounds
        cap = inbox_cap or n_total
        state.append({
            "paths": torch.empty((max(n_home, 1), stride), dtype=torch.int32, device=s.device),
            "lens": torch.zeros(max(n_home, 1), dtype=torch.int32, device=s.device),
            "inbox": torch.empty(max(cap, 1) * MSG_BYTES, dtype=torch.uint8, device=s.device),
            "send_msgs": torch.empty(max(cap, 1) * MSG_BYTES, dtype=torch.uint8, device=s.device),
            "send_recs": torch.empty(rec_cap * REC_BYTES, dtype=torch.uint8, device=s.device),
            "n_in": 0, "cap": cap, "n_home": n_home,
        })
        n = C.c_int64()
        check(L.srw_shard_seed(s.h, C.byref(cp), round_first, n_rounds, state[-1]["inbox"].data_ptr(), cap, C.byref(n),
                               state[-1]["paths"].data_ptr(), state[-1]["lens"].data_ptr(), st))
        state[-1]["n_in"] = n.value
    stats = {"super_steps": 0, "tuples_sent": 0, "records_sent": 0, "steps": 0}
    while True:
        msg_counts, rec_counts = [], []
        for s, d in zip(shards, state):
            mc, rc = (C.c_int64 * world)(), (C.c_int64 * world)()
            steps = C.c_int64()
            check(L.srw_shard_step(s.h, C.byref(cp), round_first, n_rounds, d["inbox"].data_ptr(), d["n_in"], d["send_msgs"].data_ptr(),
                                   d["send_recs"].data_ptr(), rec_cap, d["paths"].data_ptr(), d["lens"].data_ptr(), mc, rc,
                                   C.byref(steps), st))
            msg_counts.append(list(mc))
            rec_counts.append(list(rc))
            stats["steps"] += steps.value
        stats["super_steps"] += 1
        sent = sum(sum(c) for c in msg_counts)
        stats["tuples_sent"] += sent
        stats["records_sent"] += sum(sum(c) for c in rec_counts)
        if exchange.total([sent]) == 0:           # RW:162 remainingWalkers == 0
            break
        recv_m, tot_m = exchange.exchange([d["send_msgs"] for d in state], msg_counts, MSG_BYTES)
        recv_r, tot_r = exchange.exchange([d["send_recs"] for d in state], rec_counts, REC_BYTES)
        for s, d, rm, nm, rr, nr in zip(shards, state, recv_m, tot_m, recv_r, tot_r):
            if nm > d["cap"]:
                raise RuntimeError("shard %d inbox overflow: %d tuples > capacity %d" % (s.rank, nm, d["cap"]))
            if nr:
                rr = rr.contiguous()
                check(L.srw_shard_apply(s.h, C.byref(cp), round_first, n_rounds, rr.data_ptr(), nr, d["paths"].data_ptr(), st))
            d["inbox"][:nm * MSG_BYTES].copy_(rm[:nm * MSG_BYTES])
            d["n_in"] = nm
        torch.cuda.synchronize()
    out = []
    for s, d in zip(shards, state):
        steps = C.c_int64()
        check(L.srw_shard_finalize(s.h, C.byref(cp), d["n_home"], d["paths"].data_ptr(), d["lens"].data_ptr(), C.byref(steps), st))
        out.append((d["paths"][:d["n_home"]], d["lens"][:d["n_home"]]))
    return out, stats
