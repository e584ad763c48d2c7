"""CPU-only tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/srw.h
declares, the host mirror (Params / CommandParser / edge-list parser) behaves like the reference's,
the product fails loudly without a GPU, and the walker exchange works across 2 gloo ranks."""
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import KARATE, ROOT, TESTGRAPH

srw = importlib.import_module("stellar-random-walk_b200")
bld = importlib.import_module("stellar-random-walk_b200.build")


@pytest.fixture(scope="module", autouse=True)
def built():
    bld.build()


def test_library_exports_every_declared_symbol():
    L = srw.lib()
    hdr = open(os.path.join(ROOT, "include", "srw.h")).read()
    declared = set(re.findall(r"\b(srw_[a-z0-9_]+)\s*\(", hdr)) - {"srw_status"}
    assert declared == set(srw.ABI_SYMBOLS), declared ^ set(srw.ABI_SYMBOLS)
    for sym in declared:
        assert hasattr(L, sym), sym


def test_params_defaults_match_reference():
    p = srw.Params()          # Params.scala:7-23
    assert (p.w2vIter, p.w2vLr, p.w2vPartitions, p.w2vDim, p.w2vWindow) == (10, 0.025, 1, 128, 10)
    assert (p.walkLength, p.numWalks, p.p, p.q, p.weighted, p.directed) == (80, 10, 1.0, 1.0, True, False)
    assert (p.rddPartitions, p.singleOutput, p.partitioned, p.cmd) == (200, True, False, "node2vec")
    c = srw.CParams()
    srw.lib().srw_params_default(c)
    assert srw.Params.from_c(c) == srw.Params()


def test_command_parser():
    P = srw.CommandParser.parse
    ok = P(["--cmd", "randomwalk", "--input", "in.txt", "--output", "out", "--walkLength", "5", "--numWalks", "2", "--p", "0.25",
            "--q", "4", "--weighted", "false", "--directed", "true", "--partitioned", "yes", "--rddPartitions", "8",
            "--singleOutput", "0", "--lr", "0.1", "--iter", "3", "--dim", "16", "--window", "4", "--w2vPartitions", "2",
            "--seed", "77", "--sampler", "exact", "--gpus", "2"])
    assert ok == srw.Params(w2vIter=3, w2vLr=0.1, w2vPartitions=2, w2vDim=16, w2vWindow=4, walkLength=5, numWalks=2, p=0.25, q=4.0,
                            weighted=False, directed=True, input="in.txt", output="out", rddPartitions=8, singleOutput=False,
                            partitioned=True, cmd="randomwalk", seed=77, sampler="exact", gpus=2)
    assert P(["--cmd=randomwalk", "--input=a", "--output=b"]).cmd == "randomwalk"
    # CP:64-75 input/output/cmd are required; unknown options / bad values -> None (Main:25 exit 1)
    for bad in (["--cmd", "randomwalk", "--input", "a"], ["--cmd", "randomwalk", "--output", "a"], ["--input", "a", "--output", "b"],
                ["--cmd", "walk", "--input", "a", "--output", "b"], ["--cmd", "randomwalk", "--input", "a", "--output", "b", "--bogus", "1"],
                ["--cmd", "randomwalk", "--input", "a", "--output", "b", "--walkLength", "x"],
                ["--cmd", "randomwalk", "--input", "a", "--output", "b", "--weighted", "maybe"],
                ["--cmd", "randomwalk", "--input", "a", "--output", "b", "--numWalks"], ["stray"]):
        assert P(bad) is None
    assert "--walkLength" in srw.CommandParser.usage() and "--partitioned" in srw.CommandParser.usage()


def test_edge_list_parser_matches_oracle_rules(oracle):
    txt = "1 2 0.5\n2\t3   2.5\n-4 +5\r\n7 7 abc\n9 8 1e-1 3.5f"
    s, d, w, pid = srw.parse_edges(txt, weighted=True)
    assert s.tolist() == [1, 2, -4, 7, 9] and d.tolist() == [2, 3, 5, 7, 8] and pid is None
    assert w.tolist() == [0.5, 2.5, 1.0, 1.0, 3.5]
    s, d, w, pid = srw.parse_edges(txt, weighted=False)
    assert w.tolist() == [1.0] * 5
    # VRW:23-32: third column = partition id, weight only with > 3 columns
    s, d, w, pid = srw.parse_edges("1 2 7\n2 3 5 0.25\n4 5\n", weighted=True, partitioned=True)
    assert pid.tolist() == [7, 5, 0] and w.tolist() == [1.0, 0.25, 1.0]
    for bad in ["1 2\n\n3 4\n", " 1 2\n", "1\n", "1 x\n", "1 2147483648\n", "1.0 2\n"]:
        with pytest.raises(srw.SrwError) as e:
            srw.parse_edges(bad)
        assert e.value.status == srw.SRW_ERR_PARSE
        with pytest.raises(ValueError):
            oracle.Graph().load_text(bad)
    s, d, w, _ = srw.parse_edges(path=KARATE)
    assert len(s) == 78 and (w == 1.0).all()
    s, d, w, _ = srw.parse_edges(path=TESTGRAPH)
    assert (s.tolist(), d.tolist()) == ([1], [2])
    with pytest.raises(srw.SrwError) as e:
        srw.parse_edges(path="/nonexistent/file")
    assert e.value.status == srw.SRW_ERR_IO
    # the parser agrees with the oracle loader on a random file
    rng = np.random.RandomState(3)
    lines = ["%d %d %.3f" % (rng.randint(-50, 50), rng.randint(-50, 50), rng.rand()) for _ in range(200)]
    s, d, w, _ = srw.parse_edges("\n".join(lines))
    og = oracle.Graph().load_text("\n".join(lines))
    og2 = oracle.Graph().load_edges(s, d, w)
    for v in og.vertex_ids():
        assert og.neighbors(int(v)) == og2.neighbors(int(v))


def test_product_fails_loudly_without_gpu():
    if srw.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(srw.SrwError) as e:
        srw.Graph.from_edges([1], [2])
    assert e.value.status == srw.SRW_ERR_NO_DEVICE
    with pytest.raises(srw.SrwError):
        srw.RandomSample(lambda: 0.5).sample([(1, 1.0)])
    for call in (lambda: srw.parse_edges(data="1 2\n", device=True), lambda: srw.Graph.load(srw.Params(input=KARATE)),
                 lambda: srw.format_paths_device(None, None, 0, 2)):
        with pytest.raises(srw.SrwError) as e:
            call()
        assert e.value.status == srw.SRW_ERR_NO_DEVICE
    rc = srw.Main.main(["--cmd", "randomwalk", "--input", KARATE, "--output", "/tmp/srw_should_not_exist"])
    assert rc != 0 and not os.path.exists("/tmp/srw_should_not_exist")
    assert srw.Main.main(["--cmd", "randomwalk"]) == 1          # Main:25 sys.exit(1)


def test_native_cli_usage_exit_code():
    cli = os.path.join(ROOT, "stellar-random-walk_b200", "stellar-rw")
    r = subprocess.run([cli, "--cmd", "randomwalk"], capture_output=True, text=True)
    assert r.returncode == 1 and "Missing option --input" in r.stderr and "Usage" in r.stderr


def test_no_product_code_touches_the_oracle():
    pkg = os.path.join(ROOT, "stellar-random-walk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # comments may cite the oracle; code may not include, import, link or dlopen it
                assert not re.search(r"#include[^\n]*oracle|import\s+oracle|from\s+oracle|oracle_lib|libsrw_oracle|oracle/_ref", src), f


def test_walker_exchange_two_gloo_ranks():
    port = 29500 + os.getpid() % 2000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "dist_sharded_check.py"), "--exchange-only"],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "EXCHANGE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the B200 arm) at a toy scale: ONE JSON line on
    stdout with the contract's keys; rank > 0 prints nothing."""
    import json
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "10", "--steps", "1", "--warmup", "1", "--cpu-budget", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "walk-steps/sec" and d["unit"] == "steps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_abi_header_is_plain_c():
    """include/srw.h is the drop-in boundary: it must compile as C (no C++, no CUDA, no torch types) ..."""
    hdr = os.path.join(ROOT, "include", "srw.h")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # ... and every entry point says which reference interface it replaces
    src = open(hdr).read()
    assert len(re.findall(r"(RW|RS|GM|URW|VRW|CP|Main|Params)(\.scala)?:\d+", src)) >= 30
