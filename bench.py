#!/usr/bin/env python
"""bench.py -- walk-steps/sec of the node2vec second-order walk hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scale S]

One "step" = one walk round: one walker per present vertex, walkLength 80 (81 sampled transitions
each) -- `--steps 10` is the reference's numWalks 10.  Workload: synthetic RMAT (Graph500
parameters, edge factor 16), undirected, p = 0.5, q = 2.0 (BASELINE config C4's graph and bias).
`value` is measured with the CSR resident in HBM and the paths left in HBM; `e2e` goes through the
host-buffer C ABI (edge list H2D + CSR build + rounds + paths D2H inside the timed region).

The oracle (oracle/) is used here only for the `cpu_baseline` object and the `--impl reference` arm.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "walk-steps/sec"
UNIT = "steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=int(os.environ.get("SRW_BENCH_SCALE", "26")))
    ap.add_argument("--edge-factor", type=int, default=16)
    ap.add_argument("--walk-length", type=int, default=80)
    ap.add_argument("--p", type=float, default=0.5)
    ap.add_argument("--q", type=float, default=2.0)
    ap.add_argument("--weighted", type=int, default=0)
    ap.add_argument("--mode", default="auto", choices=["auto", "sharded", "peer", "replicated"],
                    help="N > 1: auto = the vertex-range-sharded walk with migrating walkers (value) + the replicated-graph leg beside it "
                         "(parity check + extra key); sharded / peer = round 1's NCCL tuple exchange / peer-gather legs; replicated = replicas only")
    ap.add_argument("--batch-rounds", type=int, default=5, help="N > 1: rounds walked as one batch of the sharded walk")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-host-abi", action="store_true")
    ap.add_argument("--sampler", default="fold", choices=["fold", "alias"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--hub-fraction", type=float, default=0.5,
                    help="N > 1: the highest-degree rows holding up to this share of the adjacency entries are replicated on every shard "
                         "(a step onto a hub does not migrate); 0 = pure vertex-range shards")
    ap.add_argument("--build", choices=["lean", "full"], default="lean",
                    help="lean: SRW_BUILD_LEAN (only the arrays the alias/fold sampler reads stay in HBM); full: every array of SRW_BUILD_ALIAS")
    ap.add_argument("--e2e-reps", type=int, default=2)
    ap.add_argument("--async-rounds", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-exact", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--gen-seed", type=int, default=42)
    ap.add_argument("--seed", type=int, default=1)
    return ap.parse_args()


_T0 = time.time()
_OUT = None


def emit(obj):
    """The ONE JSON line of this run, on the real stdout (fd 1 is pointed at stderr for everything else: NCCL
    prints its version banner on stdout, libraries may print warnings)."""
    out = _OUT if _OUT is not None else sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def log(msg):
    if os.environ.get("RANK", "0") == "0":
        sys.stderr.write("[bench %7.1fs] %s\n" % (time.time() - _T0, msg))
        sys.stderr.flush()


def build_flags(a, srw):
    """SRW_BUILD_ALIAS, plus SRW_BUILD_LEAN unless --build full (the library ignores LEAN where it does not apply: weighted graphs)."""
    return srw.BUILD_ALIAS | (srw.BUILD_LEAN if a.build == "lean" else 0)


def workload_name(a):
    return "rmat-%d ef%d undirected %s p=%g q=%g walkLength=%d (one round = one walker per present vertex)" % (
        a.scale, a.edge_factor, "weighted" if a.weighted else "unweighted", a.p, a.q, a.walk_length)


def host_threads():
    """Every core this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms pass this
    count to the oracle explicitly instead of inheriting that."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def mem_available_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                return int(ln.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def gather_ceiling(table_mib=8192):
    """The random-gather ceiling beside the walk: profiles/probes/gather_sweep (built by __graft_entry__.build(); NOT part of the
    product library) on one table size.  Returns {probe name: 64-byte-fill gathers/s} or None."""
    exe = os.path.join(ROOT, "profiles", "probes", "gather_sweep")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe, str(max(1, table_mib >> 10)), "0.5", str(table_mib)], capture_output=True, text=True, timeout=120)
        out = {}
        for ln in r.stdout.splitlines():
            if ln.startswith("{"):
                d = json.loads(ln)
                out[d["probe"]] = d["gathers_per_s"]
        return out or None
    except Exception:
        return None


def path_checksum(torch, paths, v_first, v_step, chunk=1 << 20):
    """Order-independent checksum of a path matrix whose row i belongs to the walker that started at vertex rank
    v_first + i * v_step: sum_i w(v_i) * sum_k id[i][k] * (k + 1) in wrapping int64 -- equal for any distribution of the
    same (walker, path) pairs over ranks and rows."""
    rows, stride = paths.shape
    k = torch.arange(1, stride + 1, dtype=torch.int64, device=paths.device)
    acc = torch.zeros((), dtype=torch.int64, device=paths.device)
    for lo in range(0, rows, chunk):
        hi = min(rows, lo + chunk)
        v = torch.arange(lo, hi, dtype=torch.int64, device=paths.device) * v_step + v_first
        w = ((v * 2654435761) & 0xFFFFFFFF) | 1
        acc += ((paths[lo:hi].to(torch.int64) * k).sum(1) * w).sum()
    return int(acc.item())


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = None

    def launch(self):
        """Start the nvidia-smi loop early (its start-up takes driver locks); only samples taken
        after mark() -- i.e. during the timed region -- are reported."""
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=timestamp," + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "250"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        self.t0 = time.time()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        t1 = time.time()
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        import datetime
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if self.t0 is not None and not (self.t0 - 0.05 <= ts <= t1 + 0.05):
                    continue
                s_, m_ = float(c[2]), float(c[3])
            except ValueError:
                continue
            sm.append(s_)
            mx.append(m_)
            for nme, v in zip(names, c[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        try:
            os.unlink(self.f.name)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_walk_sample(oracle_lib, offsets, col, w, a, budget_s, phase=0):
    """Times the literal reference algorithm on a strided walker sample of the CSR."""
    L = oracle_lib.lib()
    nv = len(offsets) - 1
    cfg = oracle_lib.make_cfg(walk_length=a.walk_length, num_walks=1, p=a.p, q=a.q, seed=a.seed, threads=host_threads())
    elapsed, done, chk = C.c_double(), C.c_int64(), C.c_uint64()
    fn = L.oracle_walk_csr_timed
    fn.restype = C.c_int64
    fn.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double,
                   C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]
    L.oracle_max_threads.restype = C.c_int
    stride = max(1, nv // 4096)           # ~4096 sampled start vertices spread over the vertex ranks (as many as the budget allows are walked)
    steps = fn(nv, offsets.ctypes.data, col.ctypes.data, None if w is None else w.ctypes.data, C.addressof(cfg), stride, phase % stride,
               budget_s, C.byref(elapsed), C.byref(done), C.byref(chk))
    cores = host_threads()
    return {"steps": int(steps), "elapsed_s": elapsed.value, "walkers": int(done.value), "cores": cores,
            "value": steps / max(elapsed.value, 1e-9),
            "sample": "C port of RandomSample/RandomWalk (oracle/), not Spark: %d strided start vertices of %d, %.0f s budget, "
                      "%d OpenMP threads, row copies and hash lookups omitted" % (int(done.value), nv, budget_s, cores)}


def run_reference(a):
    """--impl reference: the reference's CPU algorithm (oracle port -- the reference itself is
    Scala/Spark and cannot run in this image) on the host cores, on the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib
    L = oracle_lib.lib()
    n_edges = a.edge_factor << a.scale
    need_gb = (n_edges * 8 + 2 * n_edges * 4 + (1 << a.scale) * 16) / 1e9
    if mem_available_gb() < need_gb * 1.5 + 8:
        emit({"impl": "reference", "unavailable": "host RAM too small for a CPU build of rmat-%d (%.0f GB needed)" % (a.scale, need_gb)})
        return
    log("reference arm: generating rmat-%d on the CPU" % a.scale)
    t0 = time.time()
    src = np.empty(n_edges, np.int32)
    dst = np.empty(n_edges, np.int32)
    L.oracle_rmat_edges.argtypes = [C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
    L.oracle_rmat_edges(a.scale, a.gen_seed, 0, n_edges, src.ctypes.data, dst.ctypes.data, 0)
    log("reference arm: building the CSR on the CPU")
    n_ids = 1 << a.scale
    offsets = np.empty(n_ids + 1, np.int64)
    col = np.empty(2 * n_edges, np.int32)
    L.oracle_csr_build.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.oracle_csr_build(n_ids, n_edges, src.ctypes.data, dst.ctypes.data, offsets.ctypes.data, col.ctypes.data, 0)
    del src, dst
    # the walkers of a round are the vertices PRESENT in the edge list (RW:23, URW:81-87): compact the nominal id space to
    # ranks, as the GPU arm does, so that the strided sample below strides over real start vertices
    deg = np.diff(offsets)
    present = deg > 0
    rank_of = np.cumsum(present, dtype=np.int64) - 1
    nv_present = int(present.sum())
    offsets = np.ascontiguousarray(np.concatenate([offsets[:-1][present], offsets[-1:]]))
    for lo in range(0, len(col), 1 << 27):
        col[lo:lo + (1 << 27)] = rank_of[col[lo:lo + (1 << 27)]]
    del deg, present, rank_of
    build_s = time.time() - t0
    log("reference arm: CPU build %.1f s (%d present vertices); walking on %d threads" % (build_s, nv_present, host_threads()))
    # each step: a bounded sample, sized so that steps+warmup end within a few minutes
    per_step = max(2.0, min(a.cpu_budget, 150.0 / max(1, a.steps + a.warmup)))
    for k in range(a.warmup):
        cpu_walk_sample(oracle_lib, offsets, col, None, a, per_step, phase=k)
    tot_steps, tot_s, tot_walkers, res = 0, 0.0, 0, None
    for k in range(a.steps):
        res = cpu_walk_sample(oracle_lib, offsets, col, None, a, per_step, phase=a.warmup + k)
        tot_steps += res["steps"]
        tot_s += res["elapsed_s"]
        tot_walkers += res["walkers"]
    v = tot_steps / max(tot_s, 1e-9)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * tot_s / max(1, a.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 weights / f64 cdf (reference arithmetic)", "data": "synthetic",
            "config": {"workload": workload_name(a), "vertices_present": nv_present, "adjacency_entries": int(offsets[-1]), "walkers_per_step": nv_present,
                       "cpu_graph_build_s": round(build_s, 1), "sampled_walkers_per_step": tot_walkers / max(1, a.steps),
                       "note": "a bounded sample of the round per step: the reference algorithm is O(deg(curr) * deg(prev)) per transition"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    srw = importlib.import_module("stellar-random-walk_b200")
    lib = srw.lib()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()
    n_edges = a.edge_factor << a.scale
    stride = a.walk_length + 2

    def gen_edges():
        s = torch.empty(n_edges, dtype=torch.int32, device=dev)
        d = torch.empty(n_edges, dtype=torch.int32, device=dev)
        srw.check(lib.srw_synth_rmat_device(a.scale, a.edge_factor, a.gen_seed, 0, n_edges, s.data_ptr(), d.data_ptr()))
        w = None
        if a.weighted:
            w = torch.empty(n_edges, dtype=torch.float32, device=dev)
            srw.check(lib.srw_synth_weights_device(a.gen_seed + 1, 0, n_edges, w.data_ptr()))
        return s, d, w

    def build(s, d, w):
        torch.cuda.synchronize()
        t = time.time()
        g = srw.Graph.from_device_edges(n_edges, s.data_ptr(), d.data_ptr(), None if w is None else w.data_ptr(), False, build_flags(a, srw))
        torch.cuda.synchronize()
        return g, time.time() - t

    log("generating %s" % workload_name(a))
    d_src, d_dst, d_w = gen_edges()
    want_e2e = not a.no_e2e
    h_edges = None
    if want_e2e and mem_available_gb() > world * ((n_edges * 8) / 1e9 * 2 + 2) + 16:
        h_edges = [t.cpu().pin_memory() for t in (d_src, d_dst)] + ([d_w.cpu().pin_memory()] if d_w is not None else [])
    if world > 1:      # all ranks or none (the e2e region holds collectives)
        fl = torch.tensor([1 if h_edges is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(fl, op=dist.ReduceOp.MIN)
        if int(fl.item()) == 0:
            h_edges = None
    log("host copy of the edge list done (e2e=%s); building CSR" % (h_edges is not None))
    g, build_s = build(d_src, d_dst, d_w)
    log("CSR built in %.2f s" % build_s)
    del d_src, d_dst, d_w
    torch.cuda.empty_cache()
    nv, nnz = g.stats()
    graph_bytes = int(lib.srw_graph_device_bytes(g.h))
    build_prof = json.loads(lib.srw_graph_build_profile(g.h).decode() or "{}")

    # walkers of one round are split across ranks (replicated graph) -- the sharded mode lives in shard.cu
    lo, hi = nv * rank // world, nv * (rank + 1) // world
    n_local = hi - lo
    paths = torch.empty((n_local, stride), dtype=torch.int32, device=dev)
    lens = torch.empty(n_local, dtype=torch.int32, device=dev)
    prm = srw.Params(walkLength=a.walk_length, numWalks=1, p=a.p, q=a.q, seed=a.seed, sampler=a.sampler)
    cp = prm.to_c()

    def one_round(r):
        srw.check(lib.srw_walk_device(g.h, C.byref(cp), r * nv + lo, n_local, paths.data_ptr(), lens.data_ptr(), stream.cuda_stream))
        return srw.last_walk_info()

    # ---- proposal / membership statistics (untimed, instrumented kernel) on a bounded sample ----
    lib.srw_walk_collect_stats(1)
    n_stat = min(n_local, 1 << 22)
    srw.check(lib.srw_walk_device(g.h, C.byref(cp), lo, n_stat, paths.data_ptr(), lens.data_ptr(), stream.cuda_stream))
    st = srw.last_walk_info()
    lib.srw_walk_collect_stats(0)
    T_bar = st.proposals / max(1, st.steps)
    probes_per_step = st.probes_log2 / max(1, st.steps)
    L_bar = st.probes_log2 / max(1, st.member_tests)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    clocks.launch()
    log("stats pass done: T=%.2f proposals/step; warm-up" % T_bar)
    for r in range(a.warmup):
        one_round(r)
    barrier()
    log("timed region: %d rounds" % a.steps)
    clocks.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    kernel_ms, steps, launches = 0.0, 0, 0
    if a.async_rounds:
        # A/B: srw_walk_device_async -- the rank -> id pass of round r on the library's stream under the walk of round r+1.
        # Measured (profiles/r1_bench_rmat26_v5_async_finalize.json): no gain, the two kernels compete for the same request slots.
        pbuf = [paths, torch.empty_like(paths)]
        lbuf = [lens, torch.empty_like(lens)]
        tickets = [None, None]

        def wait_ticket(b):
            nonlocal kernel_ms, steps, launches
            if tickets[b] is not None:
                wi = srw.WalkInfo()
                srw.check(lib.srw_walk_wait(tickets[b], C.byref(wi)))
                tickets[b] = None
                kernel_ms += wi.kernel_ms
                steps += wi.steps
                launches += wi.kernel_launches

        for k in range(a.steps):
            b = k & 1
            wait_ticket(b)
            tk = C.c_void_p()
            srw.check(lib.srw_walk_device_async(g.h, C.byref(cp), (a.warmup + k) * nv + lo, n_local, pbuf[b].data_ptr(), lbuf[b].data_ptr(),
                                                stream.cuda_stream, C.byref(tk)))
            tickets[b] = tk
        wait_ticket(a.steps & 1)
        wait_ticket((a.steps + 1) & 1)
        del pbuf, lbuf
    else:
        for k in range(a.steps):
            wi = one_round(a.warmup + k)
            kernel_ms += wi.kernel_ms
            steps += wi.steps
            launches += wi.kernel_launches
    e1.record(stream)
    barrier()
    clk = clocks.stop()
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms, float(steps)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        elapsed_ms, steps_all = float(tmax[0]), int(t[1])
    else:
        steps_all = steps
    value = steps_all / (elapsed_ms * 1e-3)
    log("timed region done: %.3e steps/s" % value)

    # ---- roofline of the dominant kernel ----
    # algorithmic bytes per sampled transition (DESIGN.md "bytes per step"): row extent 8 B +
    # T * (neighbour id 4 B [+ Vose slot 8 B when weighted]) + 4 B per binary-search probe + 4 B path write
    per_prop = 4 + (8 if a.weighted else 0) + (8 if (a.weighted and a.sampler == "fold") else 0)   # + bundle weight
    B = 8 + T_bar * per_prop + 4 * probes_per_step + 4
    B_survey = T_bar * (20 + 4 * (st.probes_log2 / max(1, st.proposals))) + 4      # SURVEY 8(d) formula, same measurements
    peak, peak_src = peaks()
    kernel_s = kernel_ms * 1e-3
    achieved = steps * B / kernel_s / 1e9
    traffic, l2_requests = None, None
    lib.srw_last_walk_kernel.restype = C.c_char_p
    kernel_full = (lib.srw_last_walk_kernel() or b"").decode()      # the variant the timed rounds launched, with its template arguments
    kernel_name = kernel_full.split("<")[0]
    tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic_note = None
    if os.path.exists(tj):
        try:
            tjd = json.load(open(tj))
            # only meaningful for the launch it was captured on: the SAME kernel instantiation, one GPU, a full round of the default workload
            if world == 1 and a.scale == 26 and not a.weighted and a.walk_length == 80 and tjd.get("kernel") == kernel_full:
                traffic = tjd.get("dram_bytes_per_launch")
                l2_requests = tjd.get("l2_read_requests_per_launch") or tjd.get("l2_requests_per_launch")
            else:
                traffic_note = "profiles/ncu_traffic.json describes %s on rmat-26: not this launch (%s), traffic left null" % (tjd.get("kernel"), kernel_full)
        except Exception:
            traffic = None
    steps_per_launch = steps / max(1, a.steps)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": kernel_full, "kernel_ms_per_launch": kernel_ms / max(1, a.steps), "peak_source": peak_src,
                "bytes_per_step": B, "bytes_per_step_survey_formula": B_survey, "proposals_per_step": T_bar,
                "member_tests_per_step": st.member_tests / max(1, st.steps), "mean_probes_per_test": L_bar,
                "kernel_share_of_step": kernel_ms / elapsed_ms,
                "note": "the kernel is a random gather over a footprint far beyond L2: what bounds it is DRAM's random-access rate (one row "
                        "activation per 16/32-byte gather; gather_ceiling below, profiles/README.md 'where the ceiling lives'), not streaming "
                        "bandwidth -- every missing gather fills 64 bytes (L2::64B), hence traffic / algorithmic bytes ~ 3"}
    if traffic_note:
        roofline["traffic_note"] = traffic_note
    if traffic and l2_requests:
        roofline["traffic_bytes_per_step_ncu"] = traffic / steps_per_launch
        roofline["l2_requests_per_step_ncu"] = l2_requests / steps_per_launch
    if rank == 0 and world == 1:
        gc = gather_ceiling()
        if gc:
            # DRAM-missing gathers the kernel cannot avoid, per step: one neighbour-entry / Vose-slot gather per proposal, one hash
            # bucket per membership test (rows shorter than 8 use a <= 3-probe search instead), 1/8 of a sector for the path id, and
            # (weighted / classic kernels, whose entries do not carry the neighbour's row extent) one row descriptor
            req = T_bar + st.member_tests / max(1, st.steps) + 0.125 + (0.0 if kernel_name == "walk_fold_conv_kernel" else 1.0)
            roofline["gather_ceiling"] = {"table": "8 GiB, 16-byte L2::64B gathers (profiles/probes/gather_sweep.cu, measured in this run)",
                                          "gathers_per_s": gc, "best": max(gc.values()),
                                          "note": "a LOWER bound of the ceiling: these are the probe's long plain points; its short points under ncu reach "
                                                  "43-45e9 (profiles/r2_gather_sweep_ncu.csv).  frac_of_gather_ceiling above 1 means the walk kernel sustains "
                                                  "more random accesses per second than the dedicated probe does over an equally long run"}
            roofline["requests_per_step_model"] = req
            roofline["frac_of_gather_ceiling"] = (steps / kernel_s) * req / max(gc.values())
        else:
            roofline["gather_ceiling"] = None

    log("roofline done")
    # ---- e2e: host edge list -> H2D -> CSR build -> K rounds, every round's paths read back ----
    e2e = None
    if want_e2e and h_edges is not None:
        g.free()
        torch.cuda.empty_cache()
        ring = [torch.empty(1 << 26, dtype=torch.int32).pin_memory() for _ in range(2)]   # 2 x 256 MiB staging ring
        # two device path buffers: the D2H of round r overlaps the walk of round r+1
        bufs = [paths, torch.empty_like(paths)]
        copy_stream = torch.cuda.Stream()
        runs = []
        g2, flat, dt, e_steps, d2h = None, None, None, 0, 0
        for rep in range(max(1, a.e2e_reps)):
            # the whole region is repeated: allocation of ~100 GB of device buffers and first-touch effects make single runs
            # on a fresh box swing by tens of percent; every run is listed, the best one is reported
            if g2 is not None:
                g2.free()
                g2 = None
                torch.cuda.empty_cache()
            copied = [None, None]
            barrier()
            t0 = time.time()
            dd = [t.to(dev, non_blocking=True) for t in h_edges]
            g2 = srw.Graph.from_device_edges(n_edges, dd[0].data_ptr(), dd[1].data_ptr(), dd[2].data_ptr() if len(dd) > 2 else None, False, build_flags(a, srw))
            del dd
            torch.cuda.synchronize()
            t_built = time.time() - t0
            r_steps, r_d2h = 0, 0
            for k in range(a.steps):
                b = k & 1
                if copied[b] is not None:
                    copied[b].synchronize()            # the buffer's previous contents have left the device
                flat = bufs[b].view(-1)
                srw.check(lib.srw_walk_device(g2.h, C.byref(cp), (a.warmup + k) * nv + lo, n_local, bufs[b].data_ptr(), lens.data_ptr(), stream.cuda_stream))
                r_steps += srw.last_walk_info().steps
                with torch.cuda.stream(copy_stream):
                    for i, off in enumerate(range(0, flat.numel(), ring[0].numel())):
                        n = min(ring[0].numel(), flat.numel() - off)
                        ring[i & 1][:n].copy_(flat[off:off + n], non_blocking=True)
                        r_d2h += n * 4
                    copied[b] = torch.cuda.Event()
                    copied[b].record(copy_stream)
            copy_stream.synchronize()
            torch.cuda.synchronize()
            r_dt = time.time() - t0
            runs.append({"seconds": round(r_dt, 3), "h2d_plus_build_s": round(t_built, 3)})
            if dt is None or r_dt < dt:
                dt, e_steps, d2h = r_dt, r_steps, r_d2h
        # the last chunk that reached the host really is the tail of the last round's paths
        last_off = ((flat.numel() - 1) // ring[0].numel()) * ring[0].numel()
        last_n = flat.numel() - last_off
        last_i = (flat.numel() - 1) // ring[0].numel()
        e2e_ok = bool(torch.equal(ring[last_i & 1][:last_n], flat[last_off:].cpu()))
        # stand-alone D2H rate of the same ring (reported, not part of the timed region)
        torch.cuda.synchronize()
        tb = time.time()
        for i, off in enumerate(range(0, min(flat.numel(), 16 * ring[0].numel()), ring[0].numel())):
            n = min(ring[0].numel(), flat.numel() - off)
            ring[i & 1][:n].copy_(flat[off:off + n], non_blocking=True)
        torch.cuda.synchronize()
        d2h_gbps = min(flat.numel(), 16 * ring[0].numel()) * 4 / (time.time() - tb) / 1e9
        h2d = sum(t.numel() * t.element_size() for t in h_edges)
        if world > 1:
            # whole job: every rank copied the edge list in, built its replica, walked and read back its slice
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            cc = torch.tensor([e_steps, h2d, d2h, 1 if e2e_ok else 0], dtype=torch.int64, device=dev)
            dist.all_reduce(cc)
            dt, e_steps, h2d, d2h, e2e_ok = float(tt[0]), int(cc[0]), int(cc[1]), int(cc[2]), int(cc[3]) == world
        e2e = {"value": e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d // max(1, a.steps), "d2h_bytes_per_step": d2h // max(1, a.steps),
               "seconds": dt, "runs": runs, "last_chunk_verified": e2e_ok, "d2h_GBps_standalone": d2h_gbps, "includes": "edge-list H2D + CSR build (once%s) + %d rounds + D2H of every round's paths through a pinned ring "
                           "(the copy of round r overlaps the walk of round r+1); wall clock, max over ranks; the best of the %d runs in `runs`" % (", on every rank" if world > 1 else "", a.steps, len(runs))}
        del bufs
        g = g2

    log("e2e done: %s" % (None if e2e is None else "%.3e steps/s" % e2e["value"]))
    # ---- CPU baseline beside it (rank 0, N=1): oracle port on the same CSR ----
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        try:
            import oracle_lib
            lay_off = np.empty(nv + 1, np.int64)
            lay_col = np.empty(nnz, np.int32)
            srw.check(lib.srw_graph_layout(g.h, lay_off.ctypes.data, lay_col.ctypes.data, None, None))
            r = cpu_walk_sample(oracle_lib, lay_off, lay_col, None, a, a.cpu_budget)
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                   "sample": r["sample"] + "; CSR = the device build copied to the host (sorted rows; timing only)"}
        except Exception as ex:   # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %s" % ex}
        # the GPU kernel's OWN algorithm (alias-fold, oracle twin) on the same host cores and the same CSR
        if cpu and cpu.get("value") and not a.weighted:
            try:
                L = oracle_lib.lib()
                fn = L.oracle_fold_walk_csr_timed
                fn.restype = C.c_int64
                fn.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.POINTER(C.c_double),
                               C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.c_void_p]
                cfg = oracle_lib.make_cfg(walk_length=a.walk_length, num_walks=1, p=a.p, q=a.q, seed=a.seed, threads=0, fold=1)
                el, dn, ck = C.c_double(), C.c_int64(), C.c_uint64()
                stride_t = max(1, nv // (1 << 21))
                st_t = fn(nv, lay_off.ctypes.data, lay_col.ctypes.data, C.addressof(cfg), stride_t, 0, 5.0, C.byref(el), C.byref(dn), C.byref(ck), None)
                cpu["alias_fold_twin"] = {"value": st_t / max(el.value, 1e-9), "unit": UNIT, "cores": cpu["cores"], "walkers": int(dn.value),
                                          "note": "the product's alias-fold sampler (oracle twin, not the reference algorithm) on the host cores: "
                                                  "same CSR, binary-search membership, 5 s budget"}
            except Exception as ex:   # noqa: BLE001
                cpu["alias_fold_twin"] = {"value": None, "error": str(ex)}

    log("cpu baseline done")
    # ---- parity AT THE BENCHMARKED SCALE: the CPU twin of the alias-fold sampler walks a strided sample of round 0 on the same
    # CSR; the GPU's paths for exactly those walkers must be identical, id for id (64-bit address math, u32 row offsets and the
    # hash placement at 2^31 entries are exercised here and nowhere smaller).  A mismatch fails the run. ----
    parity = None
    if rank == 0 and world == 1 and not a.no_parity and not a.weighted and a.sampler == "fold" and g is not None:
        import oracle_lib
        L = oracle_lib.lib()
        if "lay_off" not in locals():
            lay_off = np.empty(nv + 1, np.int64)
            lay_col = np.empty(nnz, np.int32)
            srw.check(lib.srw_graph_layout(g.h, lay_off.ctypes.data, lay_col.ctypes.data, None, None))
        fn = L.oracle_fold_walk_csr_timed
        fn.restype = C.c_int64
        fn.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.POINTER(C.c_double),
                       C.POINTER(C.c_int64), C.POINTER(C.c_uint64), C.c_void_p]
        stride_p = max(1, nv // (1 << 18))
        n_s = (nv + stride_p - 1) // stride_p
        tw = np.full((n_s, stride), -2, np.int32)
        cfg = oracle_lib.make_cfg(walk_length=a.walk_length, num_walks=1, p=a.p, q=a.q, seed=a.seed, threads=host_threads(), fold=1)
        el, dn, ck = C.c_double(), C.c_int64(), C.c_uint64()
        fn(nv, lay_off.ctypes.data, lay_col.ctypes.data, C.addressof(cfg), stride_p, 0, 60.0, C.byref(el), C.byref(dn), C.byref(ck), tw.ctypes.data)
        srw.check(lib.srw_walk_device(g.h, C.byref(cp), 0, nv, paths.data_ptr(), lens.data_ptr(), stream.cuda_stream))      # round 0, whole
        idx = torch.arange(0, nv, stride_p, device=dev)
        got = paths.index_select(0, idx).cpu().numpy()
        got_len = lens.index_select(0, idx).cpu().numpy()
        vids = g.vertex_ids()
        walked = tw[:, 0] != -2
        want = np.where(tw >= 0, vids[np.maximum(tw, 0)], -1)
        want_len = (tw >= 0).sum(1)
        col_ok = np.arange(stride)[None, :] < got_len[:, None]
        equal = bool((np.where(col_ok, got, -1)[walked] == want[walked]).all() and (got_len[walked] == want_len[walked]).all())
        parity = {"walkers": int(walked.sum()), "equal": equal, "round": 0, "sample": "every %d-th start vertex" % stride_p,
                  "checker": "oracle_fold_walk_csr_timed (CPU twin of the alias-fold sampler) on the device-built CSR copied to the host"}
        log("parity at scale: %s" % parity)
        if not equal:
            emit({"metric": METRIC, "error": "parity_at_scale failed: GPU paths differ from the CPU twin", "parity_at_scale": parity})
            raise SystemExit(3)
        del tw, got, want
    # ---- e2e through the HOST-BUFFER entry points the CLI / JNI shim call: srw_graph_from_edges (host arrays in, H2D + build
    # inside) + srw_walk (all rounds, ragged paths in host memory out) ----
    host_abi = None
    if rank == 0 and world == 1 and want_e2e and h_edges is not None and not a.no_host_abi and not a.weighted:
        try:
            k_rounds = 1
            if mem_available_gb() > k_rounds * nv * stride * 4 / 1e9 * 3 + 16:
                if g is not None:
                    g.free()
                    g = None
                torch.cuda.empty_cache()
                hs, hd = h_edges[0].numpy(), h_edges[1].numpy()
                t0 = time.time()
                gh = srw.Graph.from_edges(hs, hd, None, flags=build_flags(a, srw))
                t_b = time.time() - t0
                prm_h = srw.Params(walkLength=a.walk_length, numWalks=k_rounds, p=a.p, q=a.q, seed=a.seed, sampler=a.sampler)
                res_h = gh.walk(prm_h)
                n_paths, n_steps = res_h.count(), res_h.steps()
                dt_h = time.time() - t0
                host_abi = {"value": n_steps / dt_h, "unit": UNIT, "rounds": k_rounds, "seconds": round(dt_h, 3), "graph_from_edges_s": round(t_b, 3),
                            "paths": int(n_paths), "h2d_bytes": int(hs.nbytes + hd.nbytes), "d2h_bytes": int(n_paths) * stride * 4,
                            "includes": "srw_graph_from_edges (pageable host arrays -> H2D -> CSR build) + srw_walk (%d round(s): walk, D2H, "
                                        "ragged path array in host memory): the calls srw_main and the JNI shim make" % k_rounds}
                del res_h
                gh.free()
                g = None
        except Exception as ex:   # noqa: BLE001
            host_abi = {"value": None, "error": str(ex)}
        log("host-ABI e2e leg: %s" % host_abi)
    # ---- the bit-parity sampler on the same workload (N=1, bounded sample; not part of `value`) ----
    exact = None
    if rank == 0 and world == 1 and not a.no_exact:
        try:
            n_ex = min(nv, 1 << 16)
            if g is not None:
                g.free()
            del paths, lens
            torch.cuda.empty_cache()
            s_, d_, w_ = gen_edges()
            gx = srw.Graph.from_device_edges(n_edges, s_.data_ptr(), d_.data_ptr(), None if w_ is None else w_.data_ptr(), False, srw.BUILD_ALL)
            del s_, d_, w_
            px = torch.empty((n_ex, stride), dtype=torch.int32, device=dev)
            lx = torch.empty(n_ex, dtype=torch.int32, device=dev)
            cpx = srw.Params(walkLength=a.walk_length, numWalks=1, p=a.p, q=a.q, seed=a.seed, sampler="exact").to_c()
            first = (nv // 3) if nv > 3 * n_ex else 0
            srw.check(lib.srw_walk_device(gx.h, C.byref(cpx), first, min(n_ex, 4096), px.data_ptr(), lx.data_ptr(), stream.cuda_stream))   # warm-up
            srw.check(lib.srw_walk_device(gx.h, C.byref(cpx), first, n_ex, px.data_ptr(), lx.data_ptr(), stream.cuda_stream))
            wi = srw.last_walk_info()
            exact = {"value": wi.steps / (wi.kernel_ms * 1e-3), "unit": UNIT, "walkers": n_ex, "steps": int(wi.steps), "kernel_ms": wi.kernel_ms,
                     "kernel": "walk_exact_cert2_kernel",
                     "note": "SRW_SAMPLER_EXACT: RS:12-62 literally (float32 bias weights, in-order float64 CDF), bit-identical to the oracle "
                             "(tests/test_gpu_parity.py); %d walkers of the same graph starting at vertex rank %d, kernel time only" % (n_ex, first)}
            gx.free()
            g = None
        except Exception as ex:   # noqa: BLE001
            exact = {"value": None, "error": str(ex)}
        log("exact sampler sample done")
    sharded_line = None
    if False and world > 1 and a.mode == "auto":
        # also measure the vertex-range-sharded walk (BASELINE config C4) on the same ranks
        del paths, lens
        g.free()
        torch.cuda.empty_cache()
        sub = argparse.Namespace(**vars(a))
        sub.steps, sub.warmup = min(a.steps, 4), min(a.warmup, 1)
        sub.peer_steps, sub.peer_warmup = a.steps, a.warmup
        try:
            sharded_line = run_b200_sharded(sub, own_group=False)
        except Exception as ex:   # noqa: BLE001
            sharded_line = {"error": str(ex)}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": elapsed_ms / max(1, a.steps), "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "u32 (integer thresholds; f64 only in the Vose build)",
                "data": "synthetic",
                "config": {"workload": workload_name(a), "vertices_present": nv, "adjacency_entries": nnz, "walkers_per_step": nv,
                           "graph_bytes_hbm": graph_bytes, "build_s": round(build_s, 3), "build": a.build,
                           "build_ms_per_phase": build_prof,
                           "l2": "inputs larger than L2 (CSR %.1f GB >> 126 MB), no flush needed" % (nnz * 4 / 1e9),
                           "parallelism": "1 GPU" if world == 1 else
                           "walkers (the independent units) split %d ways, no data-path collective; every rank holds the whole CSR "
                           "(%.0f GB fits one B200: the ABI's 2^32-entry limit is ~120 GB).  The vertex-range-sharded walk with the NCCL "
                           "walker all-to-all is measured beside it in `sharded_c4`" % (world, graph_bytes / 1e9),
                           "sampler": "alias-fold (SRW_SAMPLER_ALIAS_FOLD; classic alias rejection when the graph is weighted/directed)" if a.sampler == "fold" else "alias"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clk}
        if parity is not None:
            line["parity_at_scale"] = parity
        if host_abi is not None:
            line["e2e_host_abi"] = host_abi
        if exact is not None:
            line["exact_sampler"] = exact
        if sharded_line is not None:
            pg = sharded_line.get("peer_gather")
            tup = {k: sharded_line.get(k) for k in ("value", "unit", "steps", "ms_per_step", "config", "error") if k in sharded_line}
            if pg and "value" in pg:
                # C4 as the north star cuts it (graph sharded by vertex range over the GPUs): the faster of the two
                # exchange mechanisms leads, the other is reported beside it
                line["sharded_c4"] = {k: pg.get(k) for k in ("value", "unit", "steps", "ms_per_step", "kernel_ms_per_step_max_rank", "config")}
                line["sharded_c4_tuple_exchange"] = tup
            else:
                line["sharded_c4"] = tup
                if pg:
                    line["sharded_c4_peer_gather_error"] = pg.get("error")
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_b200_multi(a):
    """N > 1, the default: BASELINE config C4 as the north star cuts it.  The graph is sharded by vertex range, one rank per
    GPU; walkers MIGRATE to the shard that owns their current vertex, and the step kernel itself stores the 32-byte walker
    tuples into the destination GPU's inbox over NVLink (csrc/migrate.cuh); an NCCL all-reduce of the tuple count is the
    barrier between super-steps.  `value` is that walk.  Beside it, on the same ranks and for the same rounds: the replicated
    fallback (whole CSR on every rank, walkers split) as `replicas`, and a full-round path checksum of both legs, which must
    be equal (`parity_at_scale`)."""
    import torch
    import torch.distributed as dist
    srw = importlib.import_module("stellar-random-walk_b200")
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    lib = srw.lib()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    n_edges = a.edge_factor << a.scale
    stride = a.walk_length + 2

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def gen_edges():
        s_ = torch.empty(n_edges, dtype=torch.int32, device=dev)
        d_ = torch.empty(n_edges, dtype=torch.int32, device=dev)
        srw.check(lib.srw_synth_rmat_device(a.scale, a.edge_factor, a.gen_seed, 0, n_edges, s_.data_ptr(), d_.data_ptr()))
        return s_, d_

    log("sharded: generating %s on every rank" % workload_name(a))
    d_src, d_dst = gen_edges()
    # e2e input: every rank holds 1/N of the edge list in pinned host memory
    want_e2e = not a.no_e2e and n_edges % world == 0
    e_lo, e_hi = n_edges * rank // world, n_edges * (rank + 1) // world
    h_slice = [t[e_lo:e_hi].cpu().pin_memory() for t in (d_src, d_dst)] if want_e2e else None
    torch.cuda.synchronize()
    t0 = time.time()
    shard = sh.Shard(n_edges, d_src.data_ptr(), d_dst.data_ptr(), None, rank, world, False, dev, migrate=True, hub_fraction=a.hub_fraction)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    shard_build_prof = json.loads(lib.srw_graph_build_profile(shard.h).decode() or "{}")
    hub = {"fraction_requested": a.hub_fraction, "rows": shard.hub_rows, "entries": shard.hub_entries, "min_degree": shard.hub_min_degree,
           "bytes_per_rank": shard.hub_entries * 24}
    del d_src, d_dst
    torch.cuda.empty_cache()
    nv = shard.nv
    _, nnz_g = C.c_int64(), C.c_int64()
    nnz_global = 2 * n_edges
    shard_bytes = int(lib.srw_graph_device_bytes(shard.h))
    batch = max(1, min(a.batch_rounds, max(a.steps, 1), ((1 << 32) - 1) // max(1, nv)))
    prm = srw.Params(walkLength=a.walk_length, numWalks=1, p=a.p, q=a.q, seed=a.seed, sampler=a.sampler)
    log("sharded: rank 0 owns ranks [%d, %d) of %d, %d entries, %.1f GB, built in %.2f s; batches of %d rounds" % (
        shard.row_first, shard.row_last, nv, shard.nnz_local, shard_bytes / 1e9, build_s, batch))
    mw = sh.MigrateWalker([shard], prm, batch, stats=True, check_every=8)

    # warm-up with the instrumented kernel (proposals / filter probes / exact tests per step), then the plain kernel
    w_steps, w_stats, _ = _walk_rounds_with(mw, batch, 0, max(1, a.warmup), None)
    for h in mw.ctx:
        srw.check(lib.srw_mig_collect_stats(h, 0))
    tot = torch.tensor([w_stats["proposals"], w_stats["filter_probes"], w_stats["exact_tests"], w_stats["steps"]], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    T_bar, probes_ps, exact_ps = (float(tot[i]) / max(1, int(tot[3])) for i in range(3))
    clocks = ClockSampler(local)
    clocks.launch()
    barrier()
    log("sharded: timed region, %d rounds" % a.steps)
    clocks.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    chk = {}

    e0.record(stream)
    steps, st_last, super_steps = _walk_rounds_with(mw, batch, a.warmup, a.steps, None)
    e1.record(stream)
    barrier()
    clk = clocks.stop()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t[0])
    tot = torch.tensor([steps, st_last["spills"]], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    steps_all = int(tot[0])
    value = steps_all / (elapsed_ms * 1e-3)
    tuples_per_step = st_last["tuples_total"] / max(1, steps_all)      # inbox slots (incl. the NOP padding of open chunks) per sampled transition
    log("sharded: %.3e steps/s, %d super-steps, %.1f ms" % (value, super_steps, elapsed_ms))
    # untimed: the checksum round again (same rounds => same paths), and per-super-step kernel / barrier shares from CUDA events
    last_first = a.warmup + a.steps - 1
    out, _ = mw.run(last_first, 1)
    chk["sharded"] = path_checksum(torch, out[0][0][:shard.home_rows], shard.rank, world)
    prof = mw.profile(last_first, 1)
    pk = torch.tensor([prof["kernel_ms"], prof["barrier_ms"], prof["total_ms"]], dtype=torch.float64, device=dev)
    pmax = pk.clone()
    dist.all_reduce(pmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(pk)
    # ---- e2e: host edge-list slice -> H2D -> all-gather over NVLink -> shard build -> K rounds, every batch's home paths read back ----
    e2e = None
    mw.free()
    shard.free()
    del mw, shard, out
    torch.cuda.empty_cache()
    if want_e2e:
        ring = [torch.empty(1 << 26, dtype=torch.int32).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream()
        barrier()
        t0 = time.time()
        parts = [t_.to(dev, non_blocking=True) for t_ in h_slice]
        full = [torch.empty(n_edges, dtype=torch.int32, device=dev) for _ in range(2)]
        for f_, p_ in zip(full, parts):
            dist.all_gather_into_tensor(f_, p_)
        sh2 = sh.Shard(n_edges, full[0].data_ptr(), full[1].data_ptr(), None, rank, world, False, dev, migrate=True, hub_fraction=a.hub_fraction)
        torch.cuda.synchronize()
        t_built = time.time() - t0
        del full, parts
        mw = sh.MigrateWalker([sh2], prm, batch, check_every=8)
        snap = torch.empty((sh2.home_rows * batch, stride), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        t_ctx = time.time() - t0 - t_built          # exchange blocks: symmetric-memory allocation + rendezvous of the inboxes / path rows
        state = {"ev": None, "d2h": 0, "steps": 0}

        def read_back(first, count, paths):
            if state["ev"] is not None:
                stream.wait_event(state["ev"])            # the previous batch's copy has left `snap`
            n = paths.numel()
            snap.view(-1)[:n].copy_(paths.reshape(-1))
            done = torch.cuda.Event()
            done.record(stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                flat = snap.view(-1)
                for i, off in enumerate(range(0, n, ring[0].numel())):
                    m = min(ring[0].numel(), n - off)
                    ring[i & 1][:m].copy_(flat[off:off + m], non_blocking=True)
                    state["d2h"] += m * 4
                state["ev"] = torch.cuda.Event()
                state["ev"].record(copy_stream)

        steps_e, _, _ = _walk_rounds_with(mw, batch, a.warmup, a.steps, read_back)
        copy_stream.synchronize()
        torch.cuda.synchronize()
        dt = time.time() - t0
        tt = torch.tensor([dt, t_built, t_ctx], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        cc = torch.tensor([steps_e, sum(x.numel() * 4 for x in h_slice), state["d2h"]], dtype=torch.int64, device=dev)
        dist.all_reduce(cc)
        e2e = {"value": int(cc[0]) / float(tt[0]), "unit": UNIT, "h2d_bytes_per_step": int(cc[1]) // max(1, a.steps), "d2h_bytes_per_step": int(cc[2]) // max(1, a.steps),
               "seconds": float(tt[0]), "h2d_allgather_build_s": float(tt[1]), "exchange_block_setup_s": float(tt[2]),
               "walk_and_d2h_s": float(tt[0]) - float(tt[1]) - float(tt[2]),
               "includes": "every rank copies 1/%d of the edge list from pinned host memory (H2D), NCCL all-gather of the edge list, shard build "
                           "(rows + hash sets + replicated edge filter), %d rounds of the sharded walk in batches of %d, D2H of every batch's home path "
                           "rows through a pinned ring (the copy of batch b overlaps the walk of batch b+1); wall clock, max over ranks" % (world, a.steps, batch)}
        mw.free()
        sh2.free()
        del mw, sh2, snap, ring
        torch.cuda.empty_cache()
        log("sharded: e2e %.3e steps/s (%.2f s, build part %.2f s)" % (e2e["value"], e2e["seconds"], e2e["h2d_allgather_build_s"]))
    # ---- replicas (SURVEY 8(e) fallback, a separate line): whole CSR on every rank, walkers split, no data-path collective ----
    replicas, parity = None, None
    if not a.no_parity:
        try:
            d_src, d_dst = gen_edges()
            g = srw.Graph.from_device_edges(n_edges, d_src.data_ptr(), d_dst.data_ptr(), None, False, build_flags(a, srw))
            del d_src, d_dst
            torch.cuda.empty_cache()
            lo, hi = nv * rank // world, nv * (rank + 1) // world
            paths = torch.empty((hi - lo, stride), dtype=torch.int32, device=dev)
            lens = torch.empty(hi - lo, dtype=torch.int32, device=dev)
            cp = prm.to_c()
            for r in range(min(2, a.warmup)):
                srw.check(lib.srw_walk_device(g.h, C.byref(cp), r * nv + lo, hi - lo, paths.data_ptr(), lens.data_ptr(), stream.cuda_stream))
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            rsteps = 0
            for k in range(a.steps):
                srw.check(lib.srw_walk_device(g.h, C.byref(cp), (a.warmup + k) * nv + lo, hi - lo, paths.data_ptr(), lens.data_ptr(), stream.cuda_stream))
                rsteps += srw.last_walk_info().steps
            r1.record(stream)
            barrier()
            chk["replicas"] = path_checksum(torch, paths, lo, 1)          # `paths` holds the last timed round
            tr = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
            cr = torch.tensor([rsteps, chk["sharded"], chk["replicas"]], dtype=torch.int64, device=dev)
            dist.all_reduce(cr)
            replicas = {"value": int(cr[0]) / (float(tr[0]) * 1e-3), "unit": UNIT, "ms_per_step": float(tr[0]) / max(1, a.steps),
                        "graph_bytes_hbm_per_rank": int(lib.srw_graph_device_bytes(g.h)),
                        "parallelism": "fallback of SURVEY 8(e): the whole CSR on every rank, the walkers of a round split %d ways, no data-path collective" % world}
            parity = {"round": last_first, "walkers": nv, "checksum_sharded": int(cr[1]), "checksum_replicas": int(cr[2]), "equal": int(cr[1]) == int(cr[2]),
                      "what": "order-free 64-bit checksum over (start vertex, position, vertex id) of every path of the last timed round: the sharded walk "
                              "(%d shards, migrating walkers) against the single-GPU kernel on a replicated graph" % world}
            g.free()
            del paths, lens
        except Exception as ex:   # noqa: BLE001
            replicas = {"value": None, "error": str(ex)}
        log("replicas: %s; parity: %s" % (None if not replicas else replicas.get("value"), parity))
    # the sharded walk's own full-round checksum, whether or not a replicated graph could be built beside it (a graph beyond 2^32 - 1
    # adjacency entries exists only sharded: its checksum is compared across DIFFERENT shardings, profiles/README.md)
    ck = torch.tensor([chk["sharded"]], dtype=torch.int64, device=dev)
    dist.all_reduce(ck)
    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes per step: the single-GPU figure + what the exchange adds per step
        B = 8 + T_bar * 4 + 4                                # row extent + neighbour ids + path write, as the N = 1 line
        B_mem = probes_ps * 8                                # one 8-byte filter word per test (the binary-search charge of the N=1 line does not apply)
        B_x = tuples_per_step * 96                           # tuple written to the peer inbox + read back from the own inbox, 48 bytes each
        per_gpu_steps = steps_all / world / (elapsed_ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": elapsed_ms / max(1, a.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32 (integer thresholds; f64 only in the return-component test)", "data": "synthetic",
                "config": {"workload": workload_name(a), "vertices_present": nv, "walkers_per_step": nv,
                           "hub_replication": dict(hub, entries_share=hub["entries"] / max(1, nnz_global),
                                                   note="vertex-cut: the rows of the highest-degree vertices are kept by every shard, a step onto one does not migrate "
                                                        "(VRW:43-54 replicates a vertex's adjacency into every partition it has an edge in); --hub-fraction 0 = pure ranges"),
                           "parallelism": "graph sharded into %d edge-balanced vertex ranges (one per GPU); walkers migrate to owner(curr): the step kernel "
                                          "stores 48-byte walker tuples straight into the destination GPU's inbox over NVLink (peer memory) and path entries into "
                                          "the home GPU's path rows; NCCL all-reduce of the tuple count = barrier + termination test between super-steps; "
                                          "membership test = replicated edge filter (1 byte per adjacency entry) + exact symmetric test at owner(x)" % world,
                           "shard_bytes_hbm_rank0": shard_bytes, "build_s": round(build_s, 3), "build_ms_per_phase_rank0": shard_build_prof, "batch_rounds": batch, "super_steps": super_steps,
                           "tuples_per_step": tuples_per_step, "spills": int(tot[1]),
                           "l2": "inputs larger than L2 (shard rows + 2 GB filter >> 126 MB), no flush needed", "sampler": "alias-fold" if a.sampler == "fold" else "alias"},
                "roofline": {"bound": "hbm", "achieved": per_gpu_steps * (B + B_mem + B_x) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": per_gpu_steps * (B + B_mem + B_x) / 1e9 / peak, "traffic": None, "kernel": "mig_step_kernel<0>", "peak_source": peak_src,
                             "per": "GPU (steps/s/GPU x algorithmic bytes per step)", "bytes_per_step": B + B_mem + B_x,
                             "proposals_per_step": T_bar, "filter_probes_per_step": probes_ps, "exact_tests_per_step": exact_ps,
                             "nvlink_GBps_per_gpu_out": per_gpu_steps * (tuples_per_step * 48 + 4 * (1 - 1.0 / world)) / 1e9,
                             "nvlink_peak_GBps": 770.0,
                             "super_step_profile": {"round": last_first, "super_steps": prof["super_steps"], "kernel_ms_mean_rank": float(pk[0]) / world,
                                                    "kernel_ms_max_rank": float(pmax[0]), "barrier_ms_mean_rank": float(pk[1]) / world,
                                                    "total_ms_max_rank": float(pmax[2]),
                                                    "kernel_share": float(pk[0]) / max(1e-9, float(pk[2])),
                                                    "note": "one extra untimed round with CUDA events around every kernel and every all-reduce; barrier = "
                                                            "all-reduce latency + waiting for the slowest rank of the super-step"}},
                "cpu_baseline": None, "e2e": e2e, "gpu_launches": super_steps, "clocks": clk,
                "replicas": replicas, "parity_at_scale": parity, "checksum_sharded_last_round": int(ck[0])}
        emit(line)
        if parity is not None and not parity["equal"]:
            dist.destroy_process_group()
            raise SystemExit(3)
    dist.destroy_process_group()


def _walk_rounds_with(mw, batch, first, count, on_batch):
    """rounds [first, first + count) in batches of `batch`; returns (steps decided for this rank's home rows, stats of the last
    batch + "tuples_total" over all batches and ranks, super-steps)"""
    steps, st_last, ss, tuples = 0, None, 0, 0
    r = first
    while r < first + count:
        b = min(batch, first + count - r)
        out, st_ = mw.run(r, b)
        steps += st_["steps"]
        ss += st_["super_steps"]
        tuples += st_["tuples_sent_all_ranks"]
        st_last = dict(st_, tuples_total=tuples)
        if on_batch is not None:
            on_batch(r, b, out[0][0])
        r += b
    return steps, st_last, ss


def run_b200_sharded(a, own_group=True):
    """The graph is sharded by vertex range, one rank per GPU, walkers exchanged by an NCCL
    all-to-all every super-step (BASELINE config C4).  The K timed rounds are walked as one batch
    (rounds are independent), so `steps` rounds of work are timed exactly once.  Returns the JSON
    line (rank 0) or None."""
    import torch
    import torch.distributed as dist
    srw = importlib.import_module("stellar-random-walk_b200")
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    lib = srw.lib()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if own_group:
        dist.init_process_group("nccl", device_id=dev)
    n_edges = a.edge_factor << a.scale
    log("sharded: generating %s on every rank" % workload_name(a))
    s = torch.empty(n_edges, dtype=torch.int32, device=dev)
    d = torch.empty(n_edges, dtype=torch.int32, device=dev)
    srw.check(lib.srw_synth_rmat_device(a.scale, a.edge_factor, a.gen_seed, 0, n_edges, s.data_ptr(), d.data_ptr()))
    w = None
    if a.weighted:
        w = torch.empty(n_edges, dtype=torch.float32, device=dev)
        srw.check(lib.srw_synth_weights_device(a.gen_seed + 1, 0, n_edges, w.data_ptr()))
    torch.cuda.synchronize()
    t0 = time.time()
    shard = sh.Shard(n_edges, s.data_ptr(), d.data_ptr(), None if w is None else w.data_ptr(), rank, world, False, dev)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    del s, d, w
    torch.cuda.empty_cache()
    log("sharded: rank 0 owns ranks [%d, %d) of %d, %d entries, built in %.2f s" % (shard.row_first, shard.row_last, shard.nv, shard.nnz_local, build_s))
    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # ---- (1) peer-gather: rows stay sharded, every GPU maps every shard (CUDA IPC) and the walk kernel loads
    # remote rows over NVLink; walkers are split evenly and never migrate ----
    peer = None
    if not a.weighted:
        try:
            peer = run_peer_gather(a, srw, sh, shard, rank, world, dev, barrier)
        except Exception as ex_:   # noqa: BLE001
            peer = {"error": str(ex_)}
            log("sharded: peer-gather failed: %s" % ex_)
    if a.mode == "peer":
        line = None
        if rank == 0:
            line = dict(peer)
        shard.free()
        torch.cuda.empty_cache()
        if own_group:
            if line is not None:
                emit(line)
            dist.destroy_process_group()
        return line

    # ---- (2) tuple exchange: walkers migrate to the rows, NCCL all-to-all every super-step ----
    prm = srw.Params(walkLength=a.walk_length, numWalks=1, p=a.p, q=a.q, seed=a.seed, sampler="alias")
    ex = sh.DistExchange(device=dev)
    # worst case every walker of the batch sits on one rank (after the first hop the owner of the
    # low-degree vertex range briefly holds most of them): size the inbox for all of them
    cap = shard.nv * max(a.steps, a.warmup, 1)

    if a.warmup > 0:
        wk = sh.ShardedWalker([shard], prm, a.warmup, ex, rec_cap=max(1 << 22, shard.nv * a.warmup), inbox_cap=cap)
        wk.run(0)
        del wk
        torch.cuda.empty_cache()
    # one record slot per walker of the batch: a super-step rarely decides more than one step per walker
    # on a non-home rank, and an overflow only parks the walker for the next super-step
    walker = sh.ShardedWalker([shard], prm, a.steps, ex, rec_cap=max(1 << 22, shard.nv * a.steps), inbox_cap=cap)
    clocks = ClockSampler(local)
    clocks.launch()
    time.sleep(1.0)
    barrier()
    log("sharded: timed region, %d rounds as one batch" % a.steps)
    clocks.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, stats = walker.run(a.warmup)
    e1.record()
    barrier()
    clk = clocks.stop()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t[0])
    tot = torch.tensor([stats["steps"], stats["tuples_sent"], stats["records_sent"], int((out[0][1].long() - 1).clamp(min=0).sum())],
                       dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    steps_all, tuples, recs, steps_check = (int(x) for x in tot.tolist())
    value = steps_all / (elapsed_ms * 1e-3)
    log("sharded: %.3e steps/s, %d super-steps" % (value, stats["super_steps"]))
    if rank == 0:
        peak, peak_src = peaks()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": elapsed_ms / max(1, a.steps), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32 (integer thresholds; f64 only in the Vose build)", "data": "synthetic",
                "config": {"workload": workload_name(a), "vertices_present": shard.nv, "walkers_per_step": shard.nv,
                           "parallelism": "graph sharded into %d edge-balanced vertex ranges, NCCL all-to-all of 32-byte walker tuples "
                                          "and 16-byte path records every super-step" % world,
                           "build_s": round(build_s, 3), "super_steps": stats["super_steps"], "tuples_exchanged": tuples,
                           "records_exchanged": recs, "nvlink_bytes": tuples * 32 + recs * 16, "steps_cross_check": steps_check == steps_all,
                           "l2": "inputs larger than L2, no flush needed", "sampler": "alias"},
                "roofline": {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                             "note": "single-GPU kernel roofline is reported by the N=1 line; the sharded step adds the exchange"},
                "cpu_baseline": None, "e2e": None, "gpu_launches": stats["super_steps"] * 3, "clocks": clk, "peer_gather": peer}
    else:
        line = None
    del walker, out
    shard.free()
    torch.cuda.empty_cache()
    if own_group:
        if line is not None:
            emit(line)
        dist.destroy_process_group()
    return line


def run_peer_gather(a, srw, sh, shard, rank, world, dev, barrier):
    """Vertex-range shards + NVLink peer loads (no collective on the data path).  Times K rounds, walkers of a
    round split evenly over the ranks, paths left in the rank's HBM.  Returns the JSON object (every rank)."""
    import torch
    import torch.distributed as dist
    t0 = time.time()
    shard.attach_dist()
    attach_s = time.time() - t0
    nv = shard.nv
    stride = a.walk_length + 2
    lo, hi = nv * rank // world, nv * (rank + 1) // world
    n_local = hi - lo
    paths = torch.empty((n_local, stride), dtype=torch.int32, device=dev)
    lens = torch.empty(n_local, dtype=torch.int32, device=dev)
    prm = srw.Params(walkLength=a.walk_length, numWalks=1, p=a.p, q=a.q, seed=a.seed, sampler=a.sampler)
    stream = torch.cuda.current_stream()
    a = argparse.Namespace(**vars(a))
    a.steps, a.warmup = getattr(a, "peer_steps", a.steps), getattr(a, "peer_warmup", a.warmup)
    # bounded probe first: the whole measurement must fit a time budget whatever a peer load costs on this box
    probe_n = min(n_local, 1 << 20)
    shard.walk_device(prm, lo, min(probe_n, 1 << 14), paths.data_ptr(), lens.data_ptr(), stream.cuda_stream)
    wi = shard.walk_device(prm, lo, probe_n, paths.data_ptr(), lens.data_ptr(), stream.cuda_stream)
    est = torch.tensor([wi.kernel_ms * 1e-3 * n_local / max(1, probe_n)], dtype=torch.float64, device=dev)
    dist.all_reduce(est, op=dist.ReduceOp.MAX)
    est_round_s = float(est[0])
    budget_s = float(os.environ.get("SRW_PEER_BUDGET_S", "60"))
    sampled = None
    if est_round_s * (a.steps + a.warmup) > budget_s:
        k = int(budget_s / max(est_round_s, 1e-9))
        if k >= 2:
            a.steps, a.warmup = k - 1, 1
        else:
            # even one full round does not fit: report the probe (a strided... no: the first probe_n walkers of this rank's slice)
            t = torch.tensor([wi.kernel_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot = torch.tensor([wi.steps], dtype=torch.int64, device=dev)
            dist.all_reduce(tot)
            value = int(tot[0]) / (float(t[0]) * 1e-3)
            log("sharded: peer-gather %.3e steps/s on a %d-walker probe per rank (a full round would take %.0f s)" % (value, probe_n, est_round_s))
            del paths, lens
            torch.cuda.empty_cache()
            return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": 0, "warmup": 0, "ms_per_step": est_round_s * 1e3,
                    "scaling": "strong", "sampled": "first %d walkers of every rank's slice, one launch; full rounds skipped (time budget %.0f s)" % (probe_n, budget_s),
                    "config": {"workload": workload_name(a), "vertices_present": nv, "sampler": a.sampler, "ipc_attach_s": round(attach_s, 3)}}
    for r in range(a.warmup):
        shard.walk_device(prm, r * nv + lo, n_local, paths.data_ptr(), lens.data_ptr(), stream.cuda_stream)
    barrier()
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    clocks.launch()
    clocks.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    steps, kernel_ms = 0, 0.0
    for k in range(a.steps):
        wi = shard.walk_device(prm, (a.warmup + k) * nv + lo, n_local, paths.data_ptr(), lens.data_ptr(), stream.cuda_stream)
        steps += wi.steps
        kernel_ms += wi.kernel_ms
    e1.record(stream)
    barrier()
    clk = clocks.stop()
    t = torch.tensor([e0.elapsed_time(e1), kernel_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([steps], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    # cross-check: a checksum of the last round's paths must not depend on how the graph is cut -- compare
    # the per-rank path checksum sum with nothing here (the parity tests do that); report it for the record
    chk = torch.tensor([int(paths.view(-1)[:: 97].to(torch.int64).sum())], dtype=torch.int64, device=dev)
    dist.all_reduce(chk)
    elapsed_ms = float(t[0])
    value = int(tot[0]) / (elapsed_ms * 1e-3)
    log("sharded: peer-gather %.3e steps/s (%d rounds, %.1f ms, IPC attach %.2f s)" % (value, a.steps, elapsed_ms, attach_s))
    del paths, lens
    torch.cuda.empty_cache()
    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": elapsed_ms / max(1, a.steps), "kernel_ms_per_step_max_rank": float(t[1]) / max(1, a.steps),
            "scaling": "strong", "path_checksum": int(chk[0]),
            "config": {"workload": workload_name(a), "vertices_present": nv, "walkers_per_step": nv,
                       "parallelism": "graph sharded into %d edge-balanced vertex ranges (one per GPU, %d adjacency entries on rank 0); "
                                      "every GPU maps every shard (symmetric-memory blocks, CUDA VMM) and walk_fold_conv_kernel<PEER> loads remote rows over "
                                      "NVLink; walkers split evenly, never migrate, no collective on the data path" % (world, shard.nnz_local),
                       "peer_mapping": os.environ.get("SRW_PEER_MAP", "symm"),
                       "sampler": a.sampler, "ipc_attach_s": round(attach_s, 3)},
            "clocks": clk}


if __name__ == "__main__":
    args = parse_args()
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.mode in ("sharded", "peer"):
        run_b200_sharded(args)
    elif int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.mode == "auto":
        run_b200_multi(args)
    else:
        run_b200(args)
