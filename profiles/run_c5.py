"""One alias-fold round on BASELINE config C5's graph (Zipf alpha=2, 4 M vertices, hub rows capped at 1e6 entries; p=0.25, q=4)
for an ncu capture of walk_fold_conv_kernel on the graph the north star singles out for hub-row staging / membership.
    ncu --set full -k regex:walk_fold_conv -s 1 -c 1 -o gpurun_out/prof_c5 python profiles/run_c5.py
Prints the instrumented kernel's proposals / membership tests per step and the kernel-only rate."""
import ctypes as C
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

srw = importlib.import_module("stellar-random-walk_b200")
synth = importlib.import_module("stellar-random-walk_b200.synth")
lib = srw.lib()
hs, hd = synth.zipf_edges(1 << 22, cap=1000000, seed=7)
g = srw.Graph.from_edges(hs, hd, None, flags=srw.BUILD_ALIAS)
nv, nnz = g.stats()
L = 80
paths = torch.empty((nv, L + 2), dtype=torch.int32, device="cuda")
lens = torch.empty(nv, dtype=torch.int32, device="cuda")
cp = srw.Params(walkLength=L, numWalks=1, p=0.25, q=4.0, seed=1, sampler="fold").to_c()
lib.srw_walk_collect_stats(1)
srw.check(lib.srw_walk_device(g.h, C.byref(cp), 0, min(nv, 1 << 21), paths.data_ptr(), lens.data_ptr(), None))
st = srw.last_walk_info()
lib.srw_walk_collect_stats(0)
out = {"config": "C5 zipf 4M vertices, hub cap 1e6, p=0.25 q=4, alias-fold", "vertices": nv, "adjacency_entries": nnz,
       "proposals_per_step": st.proposals / max(1, st.steps), "member_tests_per_step": st.member_tests / max(1, st.steps),
       "mean_log2_deg_prev_per_test": st.probes_log2 / max(1, st.member_tests)}
for r in range(2):
    srw.check(lib.srw_walk_device(g.h, C.byref(cp), r * nv, nv, paths.data_ptr(), lens.data_ptr(), None))
    wi = srw.last_walk_info()
out.update(steps=int(wi.steps), kernel_ms=wi.kernel_ms, steps_per_s_kernel=wi.steps / (wi.kernel_ms * 1e-3))
print(json.dumps(out), flush=True)
