"""Build-time breakdown (srw_graph_build_profile: ms per phase) of the weighted builds the verdict asked for: BASELINE config C3
(RMAT-24 weighted) and a weighted C5 (Zipf hub graph, rows of up to 1e6 entries), plus the weighted alias-fold walk on each.
The Vose tables (K3) are built by the parallel exact-integer formulation (graph_build.cu k_alias_*).
    python profiles/run_build_profile.py > profiles/r2_build_profiles.jsonl"""
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


def report(name, g, build_s, p, q):
    nv, nnz = g.stats()
    paths = torch.empty((nv, 82), dtype=torch.int32, device="cuda")
    lens = torch.empty(nv, dtype=torch.int32, device="cuda")
    cp = srw.Params(walkLength=80, numWalks=1, p=p, q=q, seed=1, sampler="fold").to_c()
    for r in range(2):
        srw.check(lib.srw_walk_device(g.h, C.byref(cp), r * nv, nv, paths.data_ptr(), lens.data_ptr(), None))
        wi = srw.last_walk_info()
    print(json.dumps({"config": name, "vertices": nv, "adjacency_entries": nnz, "build_s": round(build_s, 3),
                      "build_ms_per_phase": json.loads(lib.srw_graph_build_profile(g.h).decode()),
                      "graph_bytes_hbm": int(lib.srw_graph_device_bytes(g.h)), "walk_kernel": lib.srw_last_walk_kernel().decode(),
                      "steps_per_s_kernel": wi.steps / (wi.kernel_ms * 1e-3), "kernel_ms": wi.kernel_ms}), flush=True)
    del paths, lens


# C3: RMAT-24 weighted
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 16 << scale
s = torch.empty(n, dtype=torch.int32, device="cuda")
d = torch.empty(n, dtype=torch.int32, device="cuda")
w = torch.empty(n, dtype=torch.float32, device="cuda")
srw.check(lib.srw_synth_rmat_device(scale, 16, 42, 0, n, s.data_ptr(), d.data_ptr()))
srw.check(lib.srw_synth_weights_device(43, 0, n, w.data_ptr()))
for rep in range(2):                      # the second build is the warm one
    torch.cuda.synchronize()
    t0 = time.time()
    g = srw.Graph.from_device_edges(n, s.data_ptr(), d.data_ptr(), w.data_ptr(), False, srw.BUILD_ALIAS)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rep == 1:
        report("C3 rmat-%d weighted (p=0.5 q=2)" % scale, g, dt, 0.5, 2.0)
    g.free()
del s, d, w
torch.cuda.empty_cache()
# weighted C5: Zipf hub graph with weights
hs, hd = synth.zipf_edges(1 << 22, cap=1000000, seed=7)
hw = synth.edge_weights(len(hs), seed=43)
ds, dd, dw = torch.from_numpy(hs).cuda(), torch.from_numpy(hd).cuda(), torch.from_numpy(hw).cuda()
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    g = srw.Graph.from_device_edges(len(hs), ds.data_ptr(), dd.data_ptr(), dw.data_ptr(), False, srw.BUILD_ALIAS)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rep == 1:
        deg_max = int(np.bincount(np.concatenate([hs, hd])).max())
        report("C5 zipf 4M vertices weighted, max degree %d (p=0.25 q=4)" % deg_max, g, dt, 0.25, 4.0)
    g.free()
