"""Device text path at size (csrc/text_io.cu), one GPU:
    python profiles/run_io.py [scale] > profiles/r1_text_io.jsonl
(1) RMAT-`scale` edge list -> text on the device (the path formatter with stride 2) -> host file -> srw_graph_load
    (mmap + CUDA parser + CSR build): parse(format(edges)) must reproduce the edge arrays, and both directions are timed;
(2) one round of walks -> RW:234-241 text on the device (formatter alone, text left in HBM), and the streamed
    srw_walk_save to a tmpfs directory (walk + format + D2H + write)."""
import ctypes as C
import importlib
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(scale):
    import torch
    srw = importlib.import_module("stellar-random-walk_b200")
    lib = srw.lib()
    os.environ["SRW_IO_TIMING"] = "1"
    n = 16 << scale
    ds = torch.empty(n, dtype=torch.int32, device="cuda")
    dd = torch.empty(n, dtype=torch.int32, device="cuda")
    srw.check(lib.srw_synth_rmat_device(scale, 16, 42, 0, n, ds.data_ptr(), dd.data_ptr()))
    e = torch.stack([ds, dd], dim=1).contiguous()
    two = torch.full((n,), 2, dtype=torch.int32, device="cuda")
    need = srw.format_paths_device(e.data_ptr(), two.data_ptr(), n, 2)
    text = torch.empty(need, dtype=torch.uint8, device="cuda")
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        srw.format_paths_device(e.data_ptr(), two.data_ptr(), n, 2, text.data_ptr(), need)
        torch.cuda.synchronize()
        best = min(best, time.time() - t0)
    print(json.dumps({"what": "format edge list (stride 2) on the device, text left in HBM", "config": "rmat-%d" % scale, "lines": n,
                      "text_bytes": need, "seconds": best, "GBps_read_plus_written": (n * 12 + need) / best / 1e9}), flush=True)
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        path = os.path.join(tmp, "edges.txt")
        host = text.cpu().numpy()
        with open(path, "wb") as f:
            f.write(host.data)
        del host, text, e, two
        torch.cuda.empty_cache()
        t0 = time.time()
        s, d, w, _ = srw.parse_edges(path=path, weighted=False, device=True)
        t_parse = time.time() - t0
        ok = bool((torch.from_numpy(s).cuda() == ds).all() and (torch.from_numpy(d).cuda() == dd).all())
        print(json.dumps({"what": "CUDA edge-list parser, host text -> device arrays -> host arrays (srw_edges_parse_buffer_device)", "lines": n,
                          "text_bytes": need, "seconds": t_parse, "lines_per_s": n / t_parse, "round_trip_equal": ok}), flush=True)
        del s, d, w
        t0 = time.time()
        g = srw.Graph.load(srw.Params(input=path, weighted=False), flags=srw.BUILD_ALIAS)
        torch.cuda.synchronize()
        t_load = time.time() - t0
        nv, nnz = g.stats()
        print(json.dumps({"what": "srw_graph_load: mmap + CUDA parser + CSR/alias build", "seconds": t_load, "text_GBps": need / t_load / 1e9,
                          "vertices": nv, "adjacency_entries": nnz}), flush=True)
        os.unlink(path)
        del ds, dd
        # ---- paths -> text ----
        stride = 82
        paths = torch.empty((nv, stride), dtype=torch.int32, device="cuda")
        lens = torch.empty(nv, dtype=torch.int32, device="cuda")
        cp = srw.Params(walkLength=80, numWalks=1, p=0.5, q=2.0, seed=1, sampler="fold").to_c()
        srw.check(lib.srw_walk_device(g.h, C.byref(cp), 0, nv, paths.data_ptr(), lens.data_ptr(), None))
        walk_ms = srw.last_walk_info().kernel_ms
        need = srw.format_paths_device(paths.data_ptr(), lens.data_ptr(), nv, stride)
        text = torch.empty(need, dtype=torch.uint8, device="cuda")
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            srw.format_paths_device(paths.data_ptr(), lens.data_ptr(), nv, stride, text.data_ptr(), need)
            torch.cuda.synchronize()
            best = min(best, time.time() - t0)
        print(json.dumps({"what": "format one round of paths on the device (size pass + scan + emit), text left in HBM", "paths": nv,
                          "ids": nv * stride, "text_bytes": need, "seconds": best, "GBps_read_plus_written": (2 * nv * stride * 4 + need) / best / 1e9,
                          "walk_kernel_ms_same_round": walk_ms}), flush=True)
        del text, paths, lens
        torch.cuda.empty_cache()
        out = os.path.join(tmp, "out")
        t0 = time.time()
        wi = g.walk_save(srw.Params(output=out, walkLength=80, numWalks=1, p=0.5, q=2.0, seed=1, sampler="fold"))
        t_save = time.time() - t0
        size = os.path.getsize(os.path.join(out, "path", "part-00000"))
        print(json.dumps({"what": "srw_walk_save: walk + device format + D2H + write() to tmpfs, one round", "seconds": t_save,
                          "file_bytes": size, "GBps_text": size / t_save / 1e9, "steps": wi.steps, "walk_kernel_ms": wi.kernel_ms}), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 22)
