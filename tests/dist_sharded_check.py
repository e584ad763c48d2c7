"""torchrun target: one shard per process.
   GPU (nccl):  torchrun --nproc-per-node 2 tests/dist_sharded_check.py            -> sharded walk == CPU twin
   CPU (gloo):  torchrun --nproc-per-node 2 tests/dist_sharded_check.py --exchange-only   -> exchange host logic only
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def exchange_only():
    dist.init_process_group("gloo")
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    rank, world = dist.get_rank(), dist.get_world_size()
    ex = sh.DistExchange()
    rng = np.random.RandomState(100 + rank)
    for it in range(5):
        counts = [int(c) for c in rng.randint(0, 50, size=world)]
        if it == 4:
            counts = [0] * world
        # item = 32 bytes: [src rank, dest rank, serial, ...]
        items = []
        for dst in range(world):
            for k in range(counts[dst]):
                rec = np.zeros(8, np.int32)
                rec[:3] = (rank, dst, k)
                items.append(rec)
        send = torch.from_numpy(np.concatenate(items).view(np.uint8) if items else np.zeros(0, np.uint8))
        recv, tot = ex.exchange([send], [counts], 32)
        got = recv[0].numpy().view(np.int32).reshape(-1, 8)
        assert tot[0] == len(got)
        assert (got[:, 1] == rank).all()
        # segments arrive grouped by source rank in rank order, each in send order
        pos = 0
        all_counts = [None] * world
        dist.all_gather_object(all_counts, counts)
        for src in range(world):
            n = all_counts[src][rank]
            seg = got[pos:pos + n]
            assert (seg[:, 0] == src).all() and (seg[:, 2] == np.arange(n)).all()
            pos += n
        allc = ex.plan([counts])
        assert allc == all_counts
        # receiving straight into a preallocated buffer
        buf = torch.zeros(64 * 32 * world, dtype=torch.uint8)
        recv2, tot2 = ex.exchange([send], [counts], 32, allc, outs=[buf])
        assert tot2 == tot and bytes(recv2[0].numpy()) == bytes(recv[0].numpy())
    # plan / owner helpers
    deg = rng.randint(0, 9, size=1000)
    pre = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    b = sh.plan_bounds(pre, 4)
    assert b[0] == 0 and b[-1] == 1000 and all(x <= y for x, y in zip(b, b[1:]))
    for v in (0, b[1], b[2] - 1 if b[2] > 0 else 0, 999):
        o = sh.owner_of(b, v)
        assert b[o] <= v < b[o + 1] or b[o] == b[o + 1]
    dist.barrier()
    if rank == 0:
        print("EXCHANGE_OK")
    dist.destroy_process_group()


def full():
    import oracle_lib
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    srw = importlib.import_module("stellar-random-walk_b200")
    sh = importlib.import_module("stellar-random-walk_b200.sharded")
    synth = importlib.import_module("stellar-random-walk_b200.synth")
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    seen = {"ok": True}

    def leg(name):
        """prints which leg turned the verdict (every rank, its own view)"""
        if seen["ok"] and not ok:
            print("LEG_FAIL rank %d: %s" % (rank, name), flush=True)
            seen["ok"] = False

    for weighted, p, q in ((False, 0.5, 2.0), (True, 0.25, 4.0)):
        s, d = synth.rmat_edges(11, 8, seed=42)
        w = synth.edge_weights(len(s), seed=43) if weighted else None
        ds, dd = torch.from_numpy(s).cuda(), torch.from_numpy(d).cuda()
        dw = torch.from_numpy(w).cuda() if weighted else None
        shard = sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None if dw is None else dw.data_ptr(), rank, world)
        prm = srw.Params(walkLength=30, numWalks=2, p=p, q=q, seed=23, sampler="alias")
        out, stats = sh.run_sharded([shard], prm, 0, 2)
        twin = oracle_lib.AliasGraph(oracle_lib.Graph().load_edges(s, d, w))
        ids, offs, st = twin.walk(walk_length=30, num_walks=2, p=p, q=q, seed=23)
        want = oracle_lib.paths_as_lists(ids, offs)
        P, Ln = out[0][0].cpu().numpy(), out[0][1].cpu().numpy()
        for rnd in range(2):
            for k, v in enumerate(shard.home_vertices()):
                row = rnd * shard.home_rows + k
                if P[row, :Ln[row]].tolist() != want[rnd * twin.nv + v]:
                    ok = False
        t = torch.tensor([stats["steps"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        ok = ok and int(t.item()) == st.steps
        leg("tuple exchange, weighted=%s" % weighted)
        if not weighted:
            # peer-gather mode over CUDA IPC: this rank walks its slice of the walkers against the whole graph
            shard.attach_dist()
            for sampler, fold in (("fold", 1), ("alias", 0)):
                ids, offs, st = twin.walk(walk_length=30, num_walks=2, p=p, q=q, seed=23, fold=fold)
                want = oracle_lib.paths_as_lists(ids, offs)
                total = 2 * twin.nv
                lo, hi = total * rank // world, total * (rank + 1) // world
                pp = torch.full((hi - lo, 32), -7, dtype=torch.int32, device="cuda")
                ll = torch.zeros(hi - lo, dtype=torch.int32, device="cuda")
                wi = shard.walk_device(srw.Params(walkLength=30, numWalks=2, p=p, q=q, seed=23, sampler=sampler), lo, hi - lo,
                                       pp.data_ptr(), ll.data_ptr())
                P, Ln = pp.cpu().numpy(), ll.cpu().numpy()
                for i in range(hi - lo):
                    if P[i, :Ln[i]].tolist() != want[lo + i]:
                        ok = False
                t = torch.tensor([wi.steps], dtype=torch.int64, device="cuda")
                dist.all_reduce(t)
                ok = ok and int(t.item()) == st.steps
                leg("peer-gather %s" % sampler)
            dist.barrier()
            # migrating walkers (csrc/migrate.cuh): the step kernel stores the tuples into the peer's inbox, NCCL all-reduce as barrier
            mshard = sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None, rank, world, migrate=True)
            for sampler, fold, seg_cap in (("fold", 1, 0), ("alias", 0, 0), ("fold", 1, 256)):
                ids, offs, st = twin.walk(walk_length=30, num_walks=2, p=p, q=q, seed=23, fold=fold)
                want = oracle_lib.paths_as_lists(ids, offs)
                mw = sh.MigrateWalker([mshard], srw.Params(walkLength=30, numWalks=2, p=p, q=q, seed=23, sampler=sampler), 2, seg_cap=seg_cap)
                for rep in range(2):                       # a context is reusable: second batch into the same block
                    mout, mstats = mw.run(0)
                    P, Ln = mout[0][0].cpu().numpy(), mout[0][1].cpu().numpy()
                    for rnd in range(2):
                        for k, v in enumerate(mshard.home_vertices()):
                            row = rnd * mshard.home_rows + k
                            if P[row, :Ln[row]].tolist() != want[rnd * twin.nv + v]:
                                ok = False
                    t = torch.tensor([mstats["steps"], mstats["spills"]], dtype=torch.int64, device="cuda")
                    dist.all_reduce(t)
                    ok = ok and int(t[0].item()) == st.steps
                    if seg_cap:
                        ok = ok and int(t[1].item()) > 0
                    leg("migrating walk %s seg_cap=%d rep=%d" % (sampler, seg_cap, rep))
                mw.free()
                dist.barrier()
            mshard.free()
            # the same walk over the VCut shard map (owner(v) = getPartition(v) mod world from a partition-id column)
            pid = np.random.default_rng(17).integers(0, 9, len(s)).astype(np.int32)
            dp = torch.from_numpy(pid).cuda()
            vshard = sh.Shard(len(s), ds.data_ptr(), dd.data_ptr(), None, rank, world, migrate=True, d_pid=dp.data_ptr())
            ids, offs, st = twin.walk(walk_length=30, num_walks=2, p=p, q=q, seed=23, fold=1)
            want = oracle_lib.paths_as_lists(ids, offs)
            mw = sh.MigrateWalker([vshard], srw.Params(walkLength=30, numWalks=2, p=p, q=q, seed=23, sampler="fold"), 2)
            mout, mstats = mw.run(0)
            P, Ln = mout[0][0].cpu().numpy(), mout[0][1].cpu().numpy()
            for rnd in range(2):
                for k, v in enumerate(vshard.home_vertices()):
                    row = rnd * vshard.home_rows + k
                    if P[row, :Ln[row]].tolist() != want[rnd * twin.nv + v]:
                        ok = False
            t = torch.tensor([mstats["steps"]], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            ok = ok and int(t[0].item()) == st.steps
            leg("VCut shard map")
            mw.free()
            dist.barrier()
            vshard.free()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_DIST_OK" if int(flag.item()) else "SHARDED_DIST_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    exchange_only() if "--exchange-only" in sys.argv else full()
