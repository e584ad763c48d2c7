"""A/B: resident blocks per SM the fold kernel is compiled for (register cap 128/64/48/40/32), RMAT-26, p=0.5 q=2.
    python profiles/run_occ.py > profiles/r1_fold_occupancy.jsonl        (one subprocess per variant: SRW_FOLD_OCC is read once)"""
import ctypes as C
import importlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(occ, scale):
    import torch
    srw = importlib.import_module("stellar-random-walk_b200")
    lib = srw.lib()
    n = 16 << scale
    s = torch.empty(n, dtype=torch.int32, device="cuda")
    d = torch.empty(n, dtype=torch.int32, device="cuda")
    srw.check(lib.srw_synth_rmat_device(scale, 16, 42, 0, n, s.data_ptr(), d.data_ptr()))
    g = srw.Graph.from_device_edges(n, s.data_ptr(), d.data_ptr(), None, False, srw.BUILD_ALIAS)
    del s, d
    nv, nnz = g.stats()
    paths = torch.empty((nv, 82), dtype=torch.int32, device="cuda")
    lens = torch.empty(nv, dtype=torch.int32, device="cuda")
    cp = srw.Params(walkLength=80, numWalks=1, p=0.5, q=2.0, seed=1, sampler="fold").to_c()
    ms, steps = [], 0
    for r in range(4):
        srw.check(lib.srw_walk_device(g.h, C.byref(cp), r * nv, nv, paths.data_ptr(), lens.data_ptr(), None))
        wi = srw.last_walk_info()
        if r:
            ms.append(wi.kernel_ms)
            steps = wi.steps
    print(json.dumps({"config": "rmat-%d" % scale, "fold_kernel_min_blocks_per_sm": occ or 4, "kernel_ms": ms,
                      "steps_per_s_kernel": steps / (min(ms) * 1e-3), "checksum": int(paths[::101].to(torch.int64).sum())}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]))
    else:
        for occ in (2, 4, 5, 6, 8):
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(occ), sys.argv[1] if len(sys.argv) > 1 else "26"],
                           env=dict(os.environ, SRW_FOLD_OCC=str(occ)), timeout=300)
