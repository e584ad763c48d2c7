"""Diagnostic (2 GPUs, one process): what does a peer load cost on this box?
  * topology / P2P attributes, device-to-device copy bandwidth
  * the peer-gather walk with both shards on device 0 (all local) vs shard 1 on device 1 (half the rows remote)"""
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

srw = importlib.import_module("stellar-random-walk_b200")
sh = importlib.import_module("stellar-random-walk_b200.sharded")
lib = srw.lib()
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
print(subprocess.run(["nvidia-smi", "nvlink", "--status", "-i", "0"], capture_output=True, text=True).stdout[:1500])
print("can_access_peer", torch.cuda.can_device_access_peer(0, 1), torch.cuda.can_device_access_peer(1, 0))
a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:0")
b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda:1")
for _ in range(2):
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    t = time.time()
    b.copy_(a)
    torch.cuda.synchronize(0); torch.cuda.synchronize(1)
    print("d2d copy 1 GiB: %.1f GB/s" % ((1 << 30) / (time.time() - t) / 1e9))
del a, b

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 16 << scale


def edges(dev):
    torch.cuda.set_device(dev)
    s = torch.empty(n, dtype=torch.int32, device="cuda:%d" % dev)
    d = torch.empty(n, dtype=torch.int32, device="cuda:%d" % dev)
    srw.check(lib.srw_synth_rmat_device(scale, 16, 42, 0, n, s.data_ptr(), d.data_ptr()))
    return s, d


def run(devs, label):
    shards = []
    for r, dev in enumerate(devs):
        s, d = edges(dev)
        shards.append(sh.Shard(n, s.data_ptr(), d.data_ptr(), None, r, len(devs), False, torch.device("cuda", dev)))
        del s, d
    for x in shards:
        x.attach_local(shards)
    torch.cuda.set_device(devs[0])
    nv = shards[0].nv
    nw = min(nv, 1 << 20)
    paths = torch.empty((nw, 82), dtype=torch.int32, device="cuda:%d" % devs[0])
    lens = torch.empty(nw, dtype=torch.int32, device="cuda:%d" % devs[0])
    prm = srw.Params(walkLength=80, numWalks=1, p=0.5, q=2.0, seed=1, sampler="fold")
    shards[0].walk_device(prm, 0, 1 << 14, paths.data_ptr(), lens.data_ptr())
    wi = shards[0].walk_device(prm, 0, nw, paths.data_ptr(), lens.data_ptr())
    print(json.dumps({"case": label, "scale": scale, "walkers": nw, "steps": wi.steps, "kernel_ms": wi.kernel_ms,
                      "steps_per_s": wi.steps / (wi.kernel_ms * 1e-3), "checksum": int(paths.to(torch.int64).sum())}), flush=True)
    for x in shards:
        x.free()


run([0, 0], "2 shards, both on device 0")
run([0, 1], "2 shards, shard 1 on device 1 (peer loads)")
