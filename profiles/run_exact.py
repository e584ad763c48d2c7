"""Exact (bit-parity, RS:12-62) sampler throughput per kernel generation on one B200 (not a bench line).
    python profiles/run_exact.py > profiles/r1_exact.jsonl
thread = one walker per thread; warp = one warp per walker, in-order float64 fold; cert = certified parallel
CDF search with in-order replay of ambiguous steps; cert2 = cert + common-neighbour list + two-level search on long rows.  The A/B switch SRW_EXACT is read once per process, so
every mode runs in its own subprocess.  131072 walkers of round 0, walkLength 80."""
import ctypes as C
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(mode, scale, p, q, n_walkers):
    import torch
    srw = importlib.import_module("stellar-random-walk_b200")
    lib = srw.lib()
    n = 16 << scale
    s = torch.empty(n, dtype=torch.int32, device="cuda")
    d = torch.empty(n, dtype=torch.int32, device="cuda")
    srw.check(lib.srw_synth_rmat_device(scale, 16, 42, 0, n, s.data_ptr(), d.data_ptr()))
    g = srw.Graph.from_device_edges(n, s.data_ptr(), d.data_ptr(), None, False, srw.BUILD_ALL)
    nv, nnz = g.stats()
    nw = min(n_walkers, nv)
    paths = torch.empty((nw, 82), dtype=torch.int32, device="cuda")
    lens = torch.empty(nw, dtype=torch.int32, device="cuda")
    cp = srw.Params(walkLength=80, numWalks=1, p=p, q=q, seed=1, sampler="exact").to_c()
    srw.check(lib.srw_walk_device(g.h, C.byref(cp), 0, min(nw, 4096), paths.data_ptr(), lens.data_ptr(), None))   # warm-up
    srw.check(lib.srw_walk_device(g.h, C.byref(cp), 0, nw, paths.data_ptr(), lens.data_ptr(), None))
    wi = srw.last_walk_info()
    chk = int(paths.to(torch.int64).sum())
    print(json.dumps({"config": "rmat-%d" % scale, "sampler": "exact", "kernel": mode, "p": p, "q": q, "walkers": nw, "steps": wi.steps,
                      "kernel_ms": wi.kernel_ms, "steps_per_s_kernel": wi.steps / (wi.kernel_ms * 1e-3),
                      "in_order_replays": wi.member_tests if mode.startswith("cert") else None, "path_checksum": chk}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), float(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6]))
    else:
        for scale, p, q, modes in ((18, 0.5, 2.0, ("thread", "warp", "cert", "cert2")), (20, 0.5, 2.0, ("cert", "cert2")),
                                   (20, 1.0, 1.0, ("cert2",)), (22, 0.5, 2.0, ("cert", "cert2")), (24, 0.5, 2.0, ("cert", "cert2"))):
            for mode in modes:
                env = dict(os.environ, SRW_EXACT=mode)
                subprocess.run([sys.executable, os.path.abspath(__file__), "--child", mode, str(scale), str(p), str(q), "131072"], env=env,
                               timeout=600)
