"""A/B: alias-fold in rank space (+ rank -> id pass) vs in id space (SRW_FOLD_IDS=1: no pass), RMAT-26, one process.
    python profiles/run_ids_ab.py [scale] > profiles/r1_fold_ids_ab.jsonl"""
import ctypes as C
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(scale):
    import torch
    srw = importlib.import_module("stellar-random-walk_b200")
    lib = srw.lib()
    n = 16 << scale
    s = torch.empty(n, dtype=torch.int32, device="cuda")
    d = torch.empty(n, dtype=torch.int32, device="cuda")
    srw.check(lib.srw_synth_rmat_device(scale, 16, 42, 0, n, s.data_ptr(), d.data_ptr()))
    cp = srw.Params(walkLength=80, numWalks=1, p=0.5, q=2.0, seed=1, sampler="fold").to_c()
    chk = {}
    for name, env in (("rank space + finalize", "0"), ("id space", "1")):
        os.environ["SRW_FOLD_IDS"] = env
        g = srw.Graph.from_device_edges(n, s.data_ptr(), d.data_ptr(), None, False, srw.BUILD_ALIAS)
        nv, nnz = g.stats()
        paths = torch.empty((nv, 82), dtype=torch.int32, device="cuda")
        lens = torch.empty(nv, dtype=torch.int32, device="cuda")
        wall, kms, steps = [], [], 0
        for r in range(4):
            torch.cuda.synchronize()
            t0 = time.time()
            srw.check(lib.srw_walk_device(g.h, C.byref(cp), r * nv, nv, paths.data_ptr(), lens.data_ptr(), None))
            torch.cuda.synchronize()
            if r:
                wall.append((time.time() - t0) * 1e3)
                kms.append(srw.last_walk_info().kernel_ms)
                steps = srw.last_walk_info().steps
        chk[name] = int(paths[::101].to(torch.int64).sum())
        print(json.dumps({"config": "rmat-%d p=0.5 q=2 walkLength=80" % scale, "variant": name, "round_ms_wall": [round(x, 2) for x in wall],
                          "walk_kernel_ms": [round(x, 2) for x in kms], "steps_per_s_round": steps / (min(wall) * 1e-3),
                          "graph_bytes_hbm": int(lib.srw_graph_device_bytes(g.h)), "checksum_round3": chk[name]}), flush=True)
        g.free()
        del paths, lens
        torch.cuda.empty_cache()
    print(json.dumps({"same_paths": len(set(chk.values())) == 1}), flush=True)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 26)
