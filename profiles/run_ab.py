"""A/B of the alias-fold walk kernel variants on ONE graph in ONE process (the switches are read per launch):
    python profiles/run_ab.py [scale] > profiles/r1_fold_ab.jsonl
v4 = per-lane state machine (pre-convergence); v5 = warp-convergent phases; 64B = L2::64B gathers; occN = blocks per SM."""
import ctypes as C
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = [
    ("v4", {"SRW_FOLD": "v4"}),
    ("v5", {"SRW_FOLD_VAR": "0"}),
    ("v5-64B", {"SRW_FOLD_VAR": "1"}),
    ("v5-occ5", {"SRW_FOLD_VAR": "0", "SRW_FOLD_OCC": "5"}),
    ("v5-occ6", {"SRW_FOLD_VAR": "0", "SRW_FOLD_OCC": "6"}),
    ("v5-64B-occ5", {"SRW_FOLD_VAR": "1", "SRW_FOLD_OCC": "5"}),
    ("v5-64B-occ6", {"SRW_FOLD_VAR": "1", "SRW_FOLD_OCC": "6"}),
    ("v4-again", {"SRW_FOLD": "v4"}),
    ("v5-again", {"SRW_FOLD_VAR": "0"}),
    ("default", {}),
]


def main(scale):
    import torch
    srw = importlib.import_module("stellar-random-walk_b200")
    lib = srw.lib()
    n = 16 << scale
    s = torch.empty(n, dtype=torch.int32, device="cuda")
    d = torch.empty(n, dtype=torch.int32, device="cuda")
    srw.check(lib.srw_synth_rmat_device(scale, 16, 42, 0, n, s.data_ptr(), d.data_ptr()))
    os.environ["SRW_FOLD_IDS"] = "0"     # the v4 kernel walks rank-labelled entries: this A/B runs in rank space (id space: run_ids_ab.py)
    g = srw.Graph.from_device_edges(n, s.data_ptr(), d.data_ptr(), None, False, srw.BUILD_ALIAS)
    del s, d
    nv, nnz = g.stats()
    paths = torch.empty((nv, 82), dtype=torch.int32, device="cuda")
    lens = torch.empty(nv, dtype=torch.int32, device="cuda")
    cp = srw.Params(walkLength=80, numWalks=1, p=0.5, q=2.0, seed=1, sampler="fold").to_c()
    keys = ("SRW_FOLD", "SRW_FOLD_VAR", "SRW_FOLD_OCC")
    for name, env in VARIANTS:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        ms, steps = [], 0
        for r in range(4):
            srw.check(lib.srw_walk_device(g.h, C.byref(cp), r * nv, nv, paths.data_ptr(), lens.data_ptr(), None))
            wi = srw.last_walk_info()
            if r:
                ms.append(wi.kernel_ms)
                steps = wi.steps
        print(json.dumps({"config": "rmat-%d p=0.5 q=2 walkLength=80" % scale, "variant": name, "kernel_ms": [round(x, 2) for x in ms],
                          "steps_per_s_kernel": steps / (min(ms) * 1e-3), "checksum_round3": int(paths[::101].to(torch.int64).sum())}), flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 26)
