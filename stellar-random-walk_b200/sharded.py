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
import os

import numpy as np
import torch

from . import BUILD_ALIAS, check, lib

BUILD_MIGRATE = 4   # srw.h SRW_BUILD_MIGRATE: the replicated edge filter of the migrating walk

MSG_BYTES = 32
REC_BYTES = 16


def _bind():
    L = lib()
    if getattr(L, "_shard_bound", False):
        return L
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    L.srw_graph_from_device_edges_sharded.argtypes = [C.c_int64, vp, vp, vp, C.c_int, C.c_uint, C.c_int, C.c_int, C.POINTER(vp)]
    L.srw_graph_from_device_edges_vcut.argtypes = [C.c_int64, vp, vp, vp, C.c_int, C.c_uint, C.c_int, C.c_int, C.c_double, C.POINTER(vp)]
    L.srw_graph_shard_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), i64p, i64p, i64p, i64p]
    L.srw_shard_seed.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, i64p, vp, vp, vp]
    L.srw_shard_step.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, vp, vp, i64p, i64p, i64p, vp]
    L.srw_shard_apply.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.c_int64, vp, vp]
    L.srw_shard_finalize.argtypes = [vp, vp, C.c_int64, vp, vp, i64p, vp]
    L.srw_shard_ipc_export.argtypes = [vp, vp]
    L.srw_shard_ipc_attach.argtypes = [vp, vp]
    L.srw_shard_attach_local.argtypes = [vp, vp]
    L.srw_shard_rows_info.argtypes = [vp, i64p, i64p, i64p, i64p]
    L.srw_shard_rows_relocate.argtypes = [vp, vp, C.c_int64]
    L.srw_shard_attach_block.argtypes = [vp, C.c_int, vp, C.c_int64, C.c_int64, C.c_int64]
    L.srw_mig_block_bytes.argtypes = [vp, vp, C.c_int64, C.c_int64, i64p]
    L.srw_mig_create.argtypes = [vp, vp, C.c_int64, C.c_int64, vp, C.POINTER(vp), C.POINTER(vp)]
    L.srw_mig_collect_stats.argtypes = [vp, C.c_int]
    L.srw_mig_begin.argtypes = [vp, C.c_int64, C.c_int64, vp]
    L.srw_mig_superstep.argtypes = [vp, C.c_int64, vp, vp]
    L.srw_mig_counters.argtypes = [vp, i64p, vp]
    L.srw_mig_finish.argtypes = [vp, C.POINTER(vp), vp, i64p, i64p, vp]
    L.srw_mig_info.argtypes = [vp, i64p, i64p, i64p, i64p]
    L.srw_mig_free.argtypes = [vp]
    L.srw_mig_free.restype = None
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


vp_t = C.c_void_p


class Shard:
    """One rank's rows of the graph plus its walker pools (device buffers owned here)."""

    def __init__(self, n_edges, d_src, d_dst, d_w, rank, world, directed=False, device=None, migrate=False, d_pid=None, hub_fraction=0.0):
        """d_pid (device pointer to the partition id of every input edge): the VCut shard map -- owner(v) = getPartition(v) mod
        world (VRW:121-134, GM:66-68) instead of an edge-balanced vertex range.  hub_fraction > 0: the highest-degree rows holding
        up to that share of the adjacency entries are replicated on every shard (a step onto a hub does not migrate).  Either one
        makes a table-mapped shard, which MigrateWalker walks (the other sharded modes refuse it)."""
        L = _bind()
        self.h = C.c_void_p()
        flags = BUILD_ALIAS | (BUILD_MIGRATE if migrate else 0)
        if d_pid is not None or hub_fraction > 0.0:
            check(L.srw_graph_from_device_edges_vcut(n_edges, d_src, d_dst, d_pid, int(directed), flags, rank, world, float(hub_fraction), C.byref(self.h)))
        else:
            check(L.srw_graph_from_device_edges_sharded(n_edges, d_src, d_dst, d_w, int(directed), flags, rank, world, C.byref(self.h)))
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
        self._block = self._symm = None
        hr, he, hd, sr = C.c_int64(), C.c_int64(), C.c_uint32(), C.c_int64()
        L.srw_graph_hub_info.argtypes = [vp_t, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)]
        check(L.srw_graph_hub_info(self.h, C.byref(hr), C.byref(he), C.byref(hd), C.byref(sr)))
        self.hub_rows, self.hub_entries, self.hub_min_degree, self.seed_rows = hr.value, he.value, hd.value, sr.value

    def free(self):
        if self.h:
            lib().srw_graph_free(self.h)
            self.h = None
        self._block = self._symm = None     # the rows lived there (srw_shard_rows_relocate): release after the handle

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # ---- peer-gather mode: every shard addressable from every GPU, no walker migration ----
    def attach_local(self, shards):
        """All shards live in this process (tests; one process driving several GPUs)."""
        for o in shards:
            if o is not self:
                check(lib().srw_shard_attach_local(self.h, o.h))
        return self

    def rows_info(self):
        r, n, hb, nb = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().srw_shard_rows_info(self.h, C.byref(r), C.byref(n), C.byref(hb), C.byref(nb)))
        return r.value, n.value, hb.value, nb.value

    def attach_blocks_local(self, shards):
        """Same-process variant of attach_dist (tests): every shard moves its rows into a block, then maps the others'."""
        for o in shards:
            if getattr(o, "_block", None) is None:
                nb = o.rows_info()[3]
                o._block = torch.empty(nb, dtype=torch.uint8, device=o.device)
                check(lib().srw_shard_rows_relocate(o.h, o._block.data_ptr(), nb))
        for o in shards:
            if o is not self:
                r, n, hb, _ = o.rows_info()
                check(lib().srw_shard_attach_block(self.h, o.rank, o._block.data_ptr(), r, n, hb))
        return self

    def attach_dist(self, group=None, mode=None):
        """One shard per process.  mode "symm" (default): the row arrays move into a torch symmetric-memory block
        (CUDA VMM allocation; torch exchanges the handles) and every peer's block is attached as an NVLink peer
        pointer.  mode "ipc": legacy cudaIpc handles of the cudaMalloc'ed arrays (kept for comparison: ~35x slower
        random loads on this platform)."""
        import torch.distributed as dist
        L = lib()
        mode = mode or os.environ.get("SRW_PEER_MAP", "symm")
        dev = self.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
        err = None
        if mode == "symm":
            import torch.distributed._symmetric_memory as symm
            mine = torch.tensor(self.rows_info(), dtype=torch.int64, device=dev)
            meta = torch.empty(self.world * 4, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(meta, mine, group=group)
            meta = meta.view(self.world, 4).tolist()
            size = max(m[3] for m in meta)
            try:
                blk = symm.empty(size, dtype=torch.uint8, device=self.device)
                hdl = symm.rendezvous(blk, group if group is not None else dist.group.WORLD)
                check(L.srw_shard_rows_relocate(self.h, blk.data_ptr(), size))
                self._block, self._symm = blk, hdl
            except Exception as e:   # noqa: BLE001
                err = e
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)     # also orders "everyone relocated" before any attach
            if int(ok.item()) == 0:
                raise RuntimeError("symmetric-memory block setup failed on at least one rank%s" % ("" if err is None else ": %s" % err))
            ptrs = list(self._symm.buffer_ptrs)
            for r in range(self.world):
                if r != self.rank:
                    check(L.srw_shard_attach_block(self.h, r, ptrs[r], meta[r][0], meta[r][1], meta[r][2]))
            torch.cuda.synchronize()
            dist.barrier(group=group)
            return self
        nb = L.srw_shard_ipc_bytes()
        blob = (C.c_uint8 * nb)()
        check(L.srw_shard_ipc_export(self.h, blob))
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).clone()
        mine = mine.to(dev)
        allb = torch.empty(self.world * nb, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allb, mine, group=group)
        allb = allb.cpu().numpy()
        try:
            for r in range(self.world):
                if r != self.rank:
                    buf = (C.c_uint8 * nb).from_buffer_copy(allb[r * nb:(r + 1) * nb].tobytes())
                    check(L.srw_shard_ipc_attach(self.h, buf))
        except Exception as e:   # noqa: BLE001  -- agree on the outcome before anyone walks (or waits in a barrier)
            err = e
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            raise RuntimeError("peer attach failed on at least one rank%s" % ("" if err is None else ": %s" % err))
        return self

    def walk_device(self, params, walker_first, n_walkers, d_paths, d_lens, stream=0):
        """srw_walk_device on this shard handle: walkers [walker_first, walker_first + n) against the whole
        (peer-attached) graph.  Returns the WalkInfo of the launch."""
        from . import last_walk_info
        cp = params.to_c()
        check(lib().srw_walk_device(self.h, C.byref(cp), walker_first, n_walkers, d_paths, d_lens, stream))
        return last_walk_info()

    @property
    def rows(self):
        """vertices whose adjacency rows this shard owns"""
        return self.row_last - self.row_first

    @property
    def home_rows(self):
        """path rows per round stored here: vertices v with v mod world == rank (round-robin homes)"""
        return (self.nv - self.rank + self.world - 1) // self.world

    def home_vertices(self):
        return range(self.rank, self.nv, self.world)


class LocalExchange:
    """All shards live in this process: route segments by slicing (tests, single device)."""

    def __init__(self, world):
        self.world = world

    def plan(self, counts):
        """counts[r] = what local rank r sends to every destination (any number of count kinds,
        concatenated).  Returns the full [world][len] matrix every rank would see."""
        return [list(c) for c in counts]

    def exchange(self, sends, counts, item_bytes, all_counts=None, outs=None):
        """sends[r]: uint8 tensor of rank r's send buffer (segments in destination order);
        counts[r][d]: items rank r sends to d.  Returns (recv tensors, recv totals) per rank."""
        out, tot = [], []
        for d in range(self.world):
            parts = []
            for r in range(self.world):
                first = sum(counts[r][:d]) * item_bytes
                parts.append(sends[r][first:first + counts[r][d] * item_bytes])
            n = sum(counts[r][d] for r in range(self.world))
            cat = torch.cat(parts) if parts else sends[d][:0]
            if outs is not None:
                # sends[d] may alias nothing in outs[d]; copy after the concatenation is materialised
                outs[d][:n * item_bytes].copy_(cat)
                cat = outs[d]
            out.append(cat)
            tot.append(n)
        return out, tot


class DistExchange:
    """One shard per process: ONE all-gather of every rank's count vector per super-step (gives the
    receive counts and the global termination test), then all-to-all of the payloads (NCCL over
    NVLink on GPUs; the same calls run over gloo in the CPU tests)."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._dev = device if device is not None else torch.device("cpu")

    def plan(self, counts):
        mine = torch.tensor(counts[0], dtype=torch.int64, device=self._dev)
        flat = torch.empty(self.world * mine.numel(), dtype=torch.int64, device=self._dev)
        self.dist.all_gather_into_tensor(flat, mine, group=self.group)
        return flat.view(self.world, mine.numel()).tolist()

    def exchange(self, sends, counts, item_bytes, all_counts=None, outs=None):
        dist = self.dist
        send, cnt = sends[0], counts[0]
        if all_counts is None:
            all_counts = self.plan([cnt])
        c_in = [int(all_counts[r][self.rank]) for r in range(self.world)]
        n_in = sum(c_in)
        recv = outs[0][:n_in * item_bytes] if outs is not None else torch.empty(n_in * item_bytes, dtype=torch.uint8, device=send.device)
        dist.all_to_all_single(recv, send[:sum(cnt) * item_bytes], output_split_sizes=[c * item_bytes for c in c_in],
                               input_split_sizes=[c * item_bytes for c in cnt], group=self.group)
        return [recv], [n_in]


class ShardedWalker:
    """Owns the device buffers of the local shard(s) and runs batches of rounds (RW:82 `for (_ <- 0 until
    numWalks)`; rounds are independent under the counter RNG, so a batch walks them concurrently)."""

    def __init__(self, shards, params, n_rounds=1, exchange=None, rec_cap=1 << 22, inbox_cap=None, stream=None):
        self.L = _bind()
        self.shards = shards
        self.params = params
        self.n_rounds = n_rounds
        self.world = shards[0].world
        if exchange is None:
            exchange = LocalExchange(self.world) if len(shards) == self.world else DistExchange()
        if isinstance(exchange, DistExchange) and shards[0].device.type == "cuda":
            exchange._dev = shards[0].device
        self.exchange = exchange
        self.cp = params.to_c()
        self.stride = params.walkLength + 2
        self.st = 0 if stream is None else stream
        self.rec_cap = rec_cap
        n_total = shards[0].nv * n_rounds
        self.state = []
        for s in shards:
            n_home = s.home_rows * n_rounds
            cap = inbox_cap or n_total
            self.state.append({
                "paths": torch.empty((max(n_home, 1), self.stride), dtype=torch.int32, device=s.device),
                "lens": torch.zeros(max(n_home, 1), dtype=torch.int32, device=s.device),
                "inbox": torch.empty(max(cap, 1) * MSG_BYTES, dtype=torch.uint8, device=s.device),
                "send_msgs": torch.empty(max(cap, 1) * MSG_BYTES, dtype=torch.uint8, device=s.device),
                "send_recs": torch.empty(rec_cap * REC_BYTES, dtype=torch.uint8, device=s.device),
                "n_in": 0, "cap": cap, "n_home": n_home,
            })

    def run(self, round_first=0):
        """Walks rounds [round_first, round_first + n_rounds).  Returns per local shard
        (paths [home_rows * n_rounds, walkLength + 2] int32 vertex ids, lens int32) and a stats dict
        (super-steps, tuples and records this process sent, sampled transitions it decided)."""
        L, cp, st, world, n_rounds = self.L, self.cp, self.st, self.world, self.n_rounds
        for s, d in zip(self.shards, self.state):
            n = C.c_int64()
            check(L.srw_shard_seed(s.h, C.byref(cp), round_first, n_rounds, d["inbox"].data_ptr(), d["cap"], C.byref(n),
                                   d["paths"].data_ptr(), d["lens"].data_ptr(), st))
            d["n_in"] = n.value
        stats = {"super_steps": 0, "tuples_sent": 0, "records_sent": 0, "steps": 0}
        while True:
            msg_counts, rec_counts = [], []
            for s, d in zip(self.shards, self.state):
                mc, rc = (C.c_int64 * world)(), (C.c_int64 * world)()
                steps = C.c_int64()
                check(L.srw_shard_step(s.h, C.byref(cp), round_first, n_rounds, d["inbox"].data_ptr(), d["n_in"],
                                       d["send_msgs"].data_ptr(), d["send_recs"].data_ptr(), self.rec_cap, d["paths"].data_ptr(),
                                       d["lens"].data_ptr(), mc, rc, C.byref(steps), st))
                msg_counts.append(list(mc))
                rec_counts.append(list(rc))
                stats["steps"] += steps.value
            stats["super_steps"] += 1
            stats["tuples_sent"] += sum(sum(c) for c in msg_counts)
            stats["records_sent"] += sum(sum(c) for c in rec_counts)
            # every rank learns every rank's counts: receive sizes + RW:162 `remainingWalkers != 0`
            allc = self.exchange.plan([m + r for m, r in zip(msg_counts, rec_counts)])
            if sum(sum(row[:world]) for row in allc) == 0:
                break
            all_m = [row[:world] for row in allc]
            all_r = [row[world:] for row in allc]
            recv_r, tot_r = self.exchange.exchange([d["send_recs"] for d in self.state], rec_counts, REC_BYTES, all_r)
            need = [sum(all_m[r][s.rank] for r in range(world)) for s in self.shards] if len(allc) == world else None
            for s, d, nm in zip(self.shards, self.state, need):
                if nm > d["cap"]:
                    raise RuntimeError("shard %d inbox overflow: %d tuples > capacity %d" % (s.rank, nm, d["cap"]))
            # the step kernel has consumed the inbox: receive the next tuples straight into it
            recv_m, tot_m = self.exchange.exchange([d["send_msgs"] for d in self.state], msg_counts, MSG_BYTES, all_m,
                                                   outs=[d["inbox"] for d in self.state])
            for s, d, nm, rr, nr in zip(self.shards, self.state, tot_m, recv_r, tot_r):
                if nr:
                    rr = rr.contiguous()
                    check(L.srw_shard_apply(s.h, C.byref(cp), round_first, n_rounds, rr.data_ptr(), nr, d["paths"].data_ptr(), st))
                d["n_in"] = nm
            torch.cuda.synchronize()
        out = []
        for s, d in zip(self.shards, self.state):
            steps = C.c_int64()
            check(L.srw_shard_finalize(s.h, C.byref(cp), d["n_home"], d["paths"].data_ptr(), d["lens"].data_ptr(), C.byref(steps), st))
            out.append((d["paths"][:d["n_home"]], d["lens"][:d["n_home"]]))
        return out, stats


def run_sharded(shards, params, round_first=0, n_rounds=1, exchange=None, rec_cap=1 << 22, stream=None, inbox_cap=None):
    """One-shot convenience wrapper around ShardedWalker."""
    return ShardedWalker(shards, params, n_rounds, exchange, rec_cap, inbox_cap, stream).run(round_first)


class MigrateWalker:
    """The migrating-walker sharded walk (csrc/migrate.cuh): ONE kernel per super-step and rank advances the resident walkers
    and stores every departing walker straight into the destination GPU's inbox over NVLink; the only collective is the
    all-reduce of the tuple count between super-steps (barrier + the RW:162 termination test).  Two deployments:
      * one shard per process (torchrun): blocks are torch symmetric-memory allocations, all-reduce over NCCL;
      * every shard in this process on one device (tests): blocks are plain tensors, shards run one after another.
    Shards must be built with migrate=True."""

    def __init__(self, shards, params, n_rounds=1, group=None, seg_cap=0, stats=False, check_every=4):
        self.L = _bind()
        self.shards = shards
        self.params = params
        self.n_rounds = n_rounds
        self.world = shards[0].world
        self.local = len(shards) == self.world
        self.group = group
        self.check_every = max(1, check_every)
        self.stride = params.walkLength + 2
        self.cp = params.to_c()
        L, cp = self.L, self.cp
        nb = C.c_int64()
        check(L.srw_mig_block_bytes(shards[0].h, C.byref(cp), n_rounds, seg_cap, C.byref(nb)))
        self.block_bytes = nb.value
        self.blocks, self.ctx, self._symm = [], [], None
        if self.local:
            for s in shards:
                self.blocks.append(torch.empty(nb.value, dtype=torch.uint8, device=s.device))
            ptrs = [b.data_ptr() for b in self.blocks]
        else:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            assert len(shards) == 1
            blk = symm.empty(nb.value, dtype=torch.uint8, device=shards[0].device)
            self._symm = symm.rendezvous(blk, group if group is not None else dist.group.WORLD)
            self.blocks.append(blk)
            ptrs = list(self._symm.buffer_ptrs)
        for i, s in enumerate(shards):
            arr = (C.c_void_p * self.world)(*[C.c_void_p(p) for p in ptrs])
            h = C.c_void_p()
            check(L.srw_mig_create(s.h, C.byref(cp), n_rounds, seg_cap, self.blocks[i].data_ptr(), arr, C.byref(h)))
            if stats:
                check(L.srw_mig_collect_stats(h, 1))
            self.ctx.append(h)
        dev = shards[0].device
        self.sent = torch.zeros(len(shards), dtype=torch.int64, device=dev)
        self.hist = torch.zeros(1 << 16, dtype=torch.int64, device=dev)
        self.lens = [torch.empty(max(1, s.home_rows * n_rounds), dtype=torch.int32, device=s.device) for s in shards]

    def free(self):
        for h in self.ctx:
            self.L.srw_mig_free(h)
        self.ctx = []
        self.blocks, self._symm = [], None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def run(self, round_first=0, n_rounds=None, max_super_steps=1 << 16):
        """Walks rounds [round_first, round_first + n_rounds).  Returns per local shard (paths [home_rows * n_rounds,
        walkLength + 2] int32 vertex ids -- a view into the shard's block, valid until the next run -- and lens), and a stats
        dict: super-steps, this process's steps / proposals / filter probes / exact tests / spills."""
        L, st = self.L, torch.cuda.current_stream().cuda_stream
        dist = None
        if not self.local:
            import torch.distributed as dist
        n_rounds = self.n_rounds if n_rounds is None else n_rounds
        for h in self.ctx:
            check(L.srw_mig_begin(h, round_first, n_rounds, st))
        if dist is not None:
            # peers store into this block from super-step 0 on: every rank must have (re)initialised its block first
            self.sent.zero_()
            dist.all_reduce(self.sent, group=self.group)
        s = 0
        while True:
            for i, h in enumerate(self.ctx):
                check(L.srw_mig_superstep(h, s, self.sent.data_ptr() + 8 * i, st))
            if dist is not None:
                dist.all_reduce(self.sent, group=self.group)       # the barrier between super-steps; sum == 0 <=> RW:162
                self.hist[s] = self.sent[0]
            else:
                self.hist[s] = self.sent.sum()
            s += 1
            if s >= 2 and (s % self.check_every == 0 or s >= max_super_steps):
                # super-step 1 still seeds walkers (the seeds are staggered over super-steps 0 and 1): only a zero from
                # super-step 1 on means that no walker is left
                h_hist = self.hist[max(1, s - self.check_every):s].tolist()
                if 0 in h_hist or s >= max_super_steps:
                    break
        h_all = self.hist[:s].tolist()
        h_all[0] = max(h_all[0], 1)
        super_steps = h_all.index(0) + 1 if 0 in h_all else s
        stats = {"super_steps": super_steps, "super_steps_launched": s, "tuples_sent_all_ranks": int(sum(h_all)), "steps": 0, "proposals": 0,
                 "filter_probes": 0, "exact_tests": 0, "exact_hits": 0, "spills": 0}
        out = []
        for i, (sh, h) in enumerate(zip(self.shards, self.ctx)):
            c8 = (C.c_int64 * 8)()
            check(L.srw_mig_counters(h, c8, st))
            p, n, steps = C.c_void_p(), C.c_int64(), C.c_int64()
            check(L.srw_mig_finish(h, C.byref(p), self.lens[i].data_ptr(), C.byref(n), C.byref(steps), st))
            off = p.value - self.blocks[i].data_ptr()
            paths = self.blocks[i][off:off + n.value * self.stride * 4].view(torch.int32).view(n.value, self.stride)
            out.append((paths, self.lens[i][:n.value]))
            assert n.value == sh.home_rows * n_rounds
            stats["steps"] += steps.value
            for k, j in (("proposals", 2), ("filter_probes", 3), ("exact_tests", 4), ("spills", 5), ("exact_hits", 7)):
                stats[k] += c8[j]
        if 0 not in h_all:
            raise RuntimeError("migrating walk did not terminate within %d super-steps" % s)
        return out, stats

    def profile(self, round_first=0, n_rounds=1, max_super_steps=4096):
        """One batch with a CUDA event around every kernel and every barrier: where a super-step's time goes on THIS rank.
        Returns sums over the batch's super-steps (ms): kernel, barrier (all-reduce incl. waiting for the slowest rank), total."""
        L, st = self.L, torch.cuda.current_stream().cuda_stream
        dist = None
        if not self.local:
            import torch.distributed as dist
        for h in self.ctx:
            check(L.srw_mig_begin(h, round_first, n_rounds, st))
        if dist is not None:
            self.sent.zero_()
            dist.all_reduce(self.sent, group=self.group)
        ev = []
        s = 0
        while True:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            for i, h in enumerate(self.ctx):
                check(L.srw_mig_superstep(h, s, self.sent.data_ptr() + 8 * i, st))
            e[1].record()
            if dist is not None:
                dist.all_reduce(self.sent, group=self.group)
                self.hist[s] = self.sent[0]
            else:
                self.hist[s] = self.sent.sum()
            e[2].record()
            ev.append(e)
            s += 1
            if s % 8 == 0 or s >= max_super_steps:
                if 0 in self.hist[max(1, s - 8):s].tolist() or s >= max_super_steps:
                    break
        torch.cuda.synchronize()
        h_all = self.hist[:s].tolist()
        h_all[0] = max(h_all[0], 1)
        n = h_all.index(0) + 1 if 0 in h_all else s
        k = sum(ev[i][0].elapsed_time(ev[i][1]) for i in range(n))
        b = sum(ev[i][1].elapsed_time(ev[i][2]) for i in range(n))
        for i, h in enumerate(self.ctx):
            p, nr, steps = C.c_void_p(), C.c_int64(), C.c_int64()
            check(L.srw_mig_finish(h, C.byref(p), self.lens[i].data_ptr(), C.byref(nr), C.byref(steps), st))
        return {"super_steps": n, "kernel_ms": k, "barrier_ms": b, "total_ms": ev[0][0].elapsed_time(ev[n - 1][2]),
                "kernel_ms_per_super_step": [ev[i][0].elapsed_time(ev[i][1]) for i in range(n)]}
