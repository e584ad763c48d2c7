"""Secondary BASELINE configs on one B200 (not bench lines): C2 (RMAT-20, p=q=1), C3 (RMAT-24 weighted,
p=0.5 q=2), C5 (Zipf hub graph, p=0.25 q=4) and the exact (bit-parity) sampler on RMAT-18.
    python profiles/run_configs.py > profiles/r1_configs.jsonl
Each line: kernel-only and wall steps/s of one round (walkLength 80), after one warm-up round."""
import ctypes as C
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

srw = importlib.import_module("stellar-random-walk_b200")
synth = importlib.import_module("stellar-random-walk_b200.synth")
lib = srw.lib()


def run(name, g, sampler, p, q, L=80, rounds=3, stats=True):
    nv, nnz = g.stats()
    paths = torch.empty((nv, L + 2), dtype=torch.int32, device="cuda")
    lens = torch.empty(nv, dtype=torch.int32, device="cuda")
    cp = srw.Params(walkLength=L, numWalks=1, p=p, q=q, seed=1, sampler=sampler).to_c()
    T = None
    if stats and sampler != "exact":
        lib.srw_walk_collect_stats(1)
        srw.check(lib.srw_walk_device(g.h, C.byref(cp), 0, min(nv, 1 << 21), paths.data_ptr(), lens.data_ptr(), None))
        wi = srw.last_walk_info()
        T = wi.proposals / max(1, wi.steps)
        lib.srw_walk_collect_stats(0)
    srw.check(lib.srw_walk_device(g.h, C.byref(cp), 0, nv, paths.data_ptr(), lens.data_ptr(), None))
    torch.cuda.synchronize()
    t0 = time.time()
    k_ms, steps = 0.0, 0
    for r in range(1, rounds + 1):
        srw.check(lib.srw_walk_device(g.h, C.byref(cp), r * nv, nv, paths.data_ptr(), lens.data_ptr(), None))
        wi = srw.last_walk_info()
        k_ms += wi.kernel_ms
        steps += wi.steps
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(json.dumps({"config": name, "sampler": sampler, "p": p, "q": q, "vertices": nv, "adjacency_entries": nnz, "rounds": rounds,
                      "steps_per_s_wall": steps / dt, "steps_per_s_kernel": steps / (k_ms * 1e-3), "proposals_per_step": T,
                      "graph_bytes_hbm": int(lib.srw_graph_device_bytes(g.h))}), flush=True)


def rmat(scale, weighted):
    n = 16 << scale
    s = torch.empty(n, dtype=torch.int32, device="cuda")
    d = torch.empty(n, dtype=torch.int32, device="cuda")
    srw.check(lib.srw_synth_rmat_device(scale, 16, 42, 0, n, s.data_ptr(), d.data_ptr()))
    w = None
    if weighted:
        w = torch.empty(n, dtype=torch.float32, device="cuda")
        srw.check(lib.srw_synth_weights_device(43, 0, n, w.data_ptr()))
    return n, s, d, w


n, s, d, w = rmat(20, False)
g = srw.Graph.from_device_edges(n, s.data_ptr(), d.data_ptr(), None, False, srw.BUILD_ALIAS)
run("C2 rmat-20 unweighted", g, "alias", 1.0, 1.0)
run("rmat-20 unweighted", g, "fold", 0.5, 2.0)
g.free()
n, s, d, w = rmat(18, False)
g = srw.Graph.from_device_edges(n, s.data_ptr(), d.data_ptr(), None, False, srw.BUILD_ALL)
run("rmat-18 unweighted (bit-parity sampler)", g, "exact", 0.5, 2.0, rounds=1, stats=False)
g.free()
n, s, d, w = rmat(24, True)
g = srw.Graph.from_device_edges(n, s.data_ptr(), d.data_ptr(), w.data_ptr(), False, srw.BUILD_ALIAS)
del s, d, w
run("C3 rmat-24 weighted", g, "alias", 0.5, 2.0)
run("C3 rmat-24 weighted", g, "fold", 0.5, 2.0)
g.free()
torch.cuda.empty_cache()
hs, hd = synth.zipf_edges(1 << 22, cap=1000000, seed=7)
g = srw.Graph.from_edges(hs, hd, None, flags=srw.BUILD_ALIAS)
run("C5 zipf 4M vertices, hub cap 1e6", g, "alias", 0.25, 4.0)
run("C5 zipf 4M vertices, hub cap 1e6", g, "fold", 0.25, 4.0)
g.free()
