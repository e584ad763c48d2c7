"""Builds libsrw.so (CUDA kernels + C ABI) and the native CLI `stellar-rw` in-tree for sm_100a.

    python stellar-random-walk_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsrw.so")
CLI = os.path.join(HERE, "stellar-rw")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-ccbin", "/usr/bin/g++",
         "--expt-relaxed-constexpr", "--extended-lambda", "-Xptxas", "-v"]
SOURCES = ["srw_host.cpp", "graph_build.cu", "walk.cu", "srw_abi.cu", "shard.cu", "migrate.cu", "multi.cu", "text_io.cu"]
HEADERS = ["srw_internal.h", "philox.cuh", "layout.h", "walk_conv.cuh", "migrate.cuh", "walk_exact.cuh", "text_io.cuh", os.path.join("..", "..", "include", "srw.h")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(OBJ, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + ARCH + FLAGS + ["-x", "cu", "-c", src, "-o", obj]
            log = open(obj + ".log", "w")
            procs.append((s, cmd, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), log, obj))
    for s, cmd, p, log, obj in procs:
        rc = p.wait()
        log.close()
        out = open(obj + ".log").read()
        if rc != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed for %s: %s" % (s, " ".join(cmd)))
        if verbose:
            sys.stderr.write(out)
    if force or procs or _stale(LIB, objs):
        subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-cudart", "static"])
    main_src = os.path.join(CSRC, "cli_main.cpp")
    if os.path.exists(main_src) and (force or _stale(CLI, [main_src, LIB])):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", main_src, "-o", CLI, "-L" + HERE, "-lsrw",
                               "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
