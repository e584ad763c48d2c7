"""The migrating-walker kernel with every shard on ONE GPU (peer pointers = local pointers, shards run one after another):
what the super-step machinery itself costs next to the single-GPU kernel on the same graph -- inbox traffic, filter probes,
remote-style path stores, per-super-step launch overhead -- before NVLink enters.
    python profiles/run_migrate_local.py [scale] [rounds] > profiles/r2_migrate_local.jsonl"""
import ctypes as C
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

srw = importlib.import_module("stellar-random-walk_b200")
sh = importlib.import_module("stellar-random-walk_b200.sharded")
lib = srw.lib()
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = 16 << scale
s = torch.empty(n, dtype=torch.int32, device="cuda")
d = torch.empty(n, dtype=torch.int32, device="cuda")
srw.check(lib.srw_synth_rmat_device(scale, 16, 42, 0, n, s.data_ptr(), d.data_ptr()))
prm = srw.Params(walkLength=80, numWalks=1, p=0.5, q=2.0, seed=1, sampler="fold")
g = srw.Graph.from_device_edges(n, s.data_ptr(), d.data_ptr(), None, False, srw.BUILD_ALIAS)
nv, nnz = g.stats()
paths = torch.empty((nv, 82), dtype=torch.int32, device="cuda")
lens = torch.empty(nv, dtype=torch.int32, device="cuda")
cp = prm.to_c()
ref_sum = None
for r in range(rounds + 1):
    srw.check(lib.srw_walk_device(g.h, C.byref(cp), r * nv, nv, paths.data_ptr(), lens.data_ptr(), None))
    wi = srw.last_walk_info()
    if r == 1:
        ref_sum = int((paths.long() * torch.arange(1, 83, device="cuda")).sum())
print(json.dumps({"config": "rmat-%d single-GPU walk_fold_conv_kernel" % scale, "vertices": nv, "adjacency_entries": nnz,
                  "steps_per_s_kernel": wi.steps / (wi.kernel_ms * 1e-3), "kernel_ms": wi.kernel_ms}), flush=True)
g.free()
del paths, lens
torch.cuda.empty_cache()
only = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hub_fracs = [float(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0.0]
for world, hub in [(w, h) for w in ((only,) if only else (1, 2, 4, 8)) for h in hub_fracs]:
    shards = [sh.Shard(n, s.data_ptr(), d.data_ptr(), None, r, world, migrate=True, hub_fraction=hub if world > 1 else 0.0) for r in range(world)]
    for stats in (True, False):
        mw = sh.MigrateWalker(shards, prm, rounds, stats=stats, check_every=8)
        mw.run(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        out, st = mw.run(1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        chk = None
        if not stats:
            # round 1 is the first round of this batch: rows [0, home_rows) of every shard
            chk = sum(int((p[:x.home_rows].long() * torch.arange(1, 83, device="cuda")).sum()) for x, (p, _) in zip(shards, out))
        print(json.dumps({"config": "rmat-%d migrate, %d shards on one GPU" % (scale, world), "hub_fraction": hub, "hub_rows": shards[0].hub_rows,
                          "hub_entries_share": shards[0].hub_entries / max(1, nnz), "shard_bytes": [int(lib.srw_graph_device_bytes(x.h)) for x in shards],
                          "instrumented": stats, "rounds": rounds,
                          "steps": st["steps"], "ms": ms, "steps_per_s": st["steps"] / (ms * 1e-3), "super_steps": st["super_steps"],
                          "launched": st["super_steps_launched"], "tuples_per_step": st["tuples_sent_all_ranks"] / max(1, st["steps"]),
                          "proposals_per_step": st["proposals"] / max(1, st["steps"]), "filter_probes_per_step": st["filter_probes"] / max(1, st["steps"]),
                          "exact_tests_per_step": st["exact_tests"] / max(1, st["steps"]), "spills": st["spills"],
                          "checksum_equals_single_gpu": None if chk is None else chk == ref_sum, "block_gb": mw.block_bytes / 1e9}), flush=True)
        mw.free()
        del mw, out
        torch.cuda.empty_cache()
    for x in shards:
        x.free()
    del shards
    torch.cuda.empty_cache()
