"""Device text path (csrc/text_io.cu) through the C ABI: the CUDA edge-list parser == the serial host parser
(URW:23-34 / VRW:19-34 rules, pinned by the reference-KAT tests), the CUDA formatter == RW:234-241 text, the
streamed srw_walk_save writes the files srw_walk + srw_save write, and parse(format(edges)) == edges at
BASELINE config C2's size."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest

from conftest import KARATE, TESTGRAPH

import test_text_rules as rules

pytestmark = pytest.mark.gpu

srw = importlib.import_module("stellar-random-walk_b200")
synth = importlib.import_module("stellar-random-walk_b200.synth")


def _same_edges(a, b, partitioned):
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    if partitioned:
        assert np.array_equal(a[3], b[3])


def _param_cases():
    for mark in rules.test_lines_match_host_parser.pytestmark:
        if mark.name == "parametrize":
            return list(mark.args[1])
    raise AssertionError("no cases")


@pytest.mark.parametrize("text,weighted,partitioned", _param_cases())
def test_device_parser_equals_host_parser(text, weighted, partitioned):
    _same_edges(srw.parse_edges(data=text, weighted=weighted, partitioned=partitioned, device=True),
                srw.parse_edges(data=text, weighted=weighted, partitioned=partitioned), partitioned)


def test_device_parser_reference_fixtures():
    for path in (KARATE, TESTGRAPH):
        for weighted in (False, True):
            _same_edges(srw.parse_edges(path=path, weighted=weighted, device=True), srw.parse_edges(path=path, weighted=weighted), False)


@pytest.mark.parametrize("text", ["1 2\n\n3 4\n", " 1 2\n", "1\n", "a b\n", "2147483648 1\n", "1 2\n3 -2147483649\n", "1 2\n3 4\n5 6\nx 7\n8 y\n"])
def test_device_parser_reports_the_hosts_first_error(text):
    with pytest.raises(srw.SrwError) as host:
        srw.parse_edges(data=text)
    with pytest.raises(srw.SrwError) as dev:
        srw.parse_edges(data=text, device=True)
    assert str(dev.value) == str(host.value)


def test_device_parser_chunked_random_text(monkeypatch):
    """Chunks cut at line boundaries (never inside a \\r\\n), host float fallback inside chunks, line numbers of
    errors counted across chunks."""
    rng = np.random.RandomState(17)
    ints = ["0", "1", "34", "-7", "+9", "2147483647", "123456"]
    ws = ["1", "0.5", "1.250", "2e-2", ".125", "abc", "0x1.8p1", "3f", "-0.0", "12345678.9"]
    ends = ["\n", "\r\n", "\r"]
    lines = []
    for _ in range(20000):
        cols = [ints[rng.randint(len(ints))], ints[rng.randint(len(ints))], str(rng.randint(0, 8)), ws[rng.randint(len(ws))]]
        lines.append(" ".join(cols[:rng.randint(2, 5)]) + ends[rng.randint(len(ends))])
    text = "".join(lines) + "1 2\n"
    for chunk in ("64", "1000", "65536"):
        monkeypatch.setenv("SRW_PARSE_CHUNK_BYTES", chunk)
        for partitioned in (False, True):
            _same_edges(srw.parse_edges(data=text, weighted=True, partitioned=partitioned, device=True),
                        srw.parse_edges(data=text, weighted=True, partitioned=partitioned), partitioned)
    bad = "".join(lines[:15000]) + "oops 1\n" + "".join(lines[15000:])
    monkeypatch.setenv("SRW_PARSE_CHUNK_BYTES", "4096")
    with pytest.raises(srw.SrwError) as dev:
        srw.parse_edges(data=bad, device=True)
    assert "line 15001:" in str(dev.value)


def test_graph_load_uses_the_device_parser(oracle, tmp_path):
    """srw_graph_load (device parse + device build) == the oracle's adjacency, karate with weights and CRLF."""
    rows = [ln.split() for ln in open(KARATE).read().split("\n") if ln]
    txt = "".join("%s\t%s %.2f\r\n" % (a, b, 0.5 + (i % 5) / 4.0) for i, (a, b) in enumerate(rows))
    inp = tmp_path / "k.txt"
    inp.write_bytes(txt.encode())
    g = srw.Graph.load(srw.Params(input=str(inp), weighted=True))
    og = oracle.Graph().load_text(txt, weighted=True)
    assert g.stats() == (og.num_vertices, og.num_edges)
    for v in og.vertex_ids():
        assert g.neighbors(int(v)) == og.neighbors(int(v))


def test_graph_load_accepts_a_directory_of_part_files(oracle, tmp_path):
    """SparkContext.textFile on a directory (URW:23): part files in name order, _SUCCESS / hidden files skipped, a last
    line without a newline still ends its file."""
    rows = open(KARATE).read().split("\n")
    rows = [r for r in rows if r]
    d = tmp_path / "in"
    d.mkdir()
    (d / "part-00000").write_text("\n".join(rows[:30]))              # no trailing newline
    (d / "part-00001").write_text("\n".join(rows[30:]) + "\n")
    (d / "_SUCCESS").write_text("")
    (d / ".part-00000.crc").write_text("garbage that must not be parsed")
    g = srw.Graph.load(srw.Params(input=str(d)))
    og = oracle.Graph().load_file(KARATE)
    assert g.stats() == (og.num_vertices, og.num_edges)
    for v in og.vertex_ids():
        assert g.neighbors(int(v)) == og.neighbors(int(v))


def _format_on_device(paths, lens):
    import torch
    n, stride = paths.shape
    dp = torch.from_numpy(paths).cuda()
    dl = torch.from_numpy(lens).cuda()
    need = srw.format_paths_device(dp.data_ptr(), dl.data_ptr(), n, stride)
    buf = torch.empty(need + 64, dtype=torch.uint8, device="cuda")
    base = buf.data_ptr()
    outs = []
    for shift in (0, 1, 7):                       # destination phase relative to 16 bytes: head/body/tail split of every line
        buf.fill_(0x55)
        got = srw.format_paths_device(dp.data_ptr(), dl.data_ptr(), n, stride, base + shift, need)
        assert got == need
        host = buf.cpu().numpy()
        assert (host[:shift] == 0x55).all() and (host[shift + need:] == 0x55).all()
        outs.append(host[shift:shift + need].tobytes())
    assert outs[0] == outs[1] == outs[2]
    return outs[0]


@pytest.mark.parametrize("stride", [1, 2, 12, 33, 82, 200])
def test_device_formatter_equals_join(stride):
    rng = np.random.RandomState(stride)
    n = 3000
    vals = np.concatenate([np.array([0, 1, 9, 10, 99, 100, 2147483647, -2147483648, -1, -10, 1000000000, 999999999], np.int32),
                           rng.randint(-2**31, 2**31 - 1, 300).astype(np.int32), rng.randint(0, 40000000, 3000).astype(np.int32)])
    paths = vals[rng.randint(0, len(vals), (n, stride))].astype(np.int32)
    lens = rng.randint(1, stride + 1, n).astype(np.int32)
    lens[::7] = stride
    want = "".join("\t".join(str(int(v)) for v in paths[i, :lens[i]]) + "\n" for i in range(n)).encode()
    assert _format_on_device(paths, lens) == want


def test_device_formatter_long_lines():
    """walkLength 5000: one line (60 KB of staging) per block with opt-in shared memory."""
    rng = np.random.RandomState(1)
    paths = rng.randint(-2**31, 2**31 - 1, (40, 5002)).astype(np.int32)
    lens = np.full(40, 5002, np.int32)
    lens[3] = 17
    want = "".join("\t".join(str(int(v)) for v in paths[i, :lens[i]]) + "\n" for i in range(40)).encode()
    assert _format_on_device(paths, lens) == want


@pytest.mark.parametrize("single,parts,chunk", [(True, 200, None), (False, 7, "4000"), (False, 200, "100000"), (False, 3, None)])
def test_walk_save_writes_the_files_of_walk_plus_save(tmp_path, monkeypatch, single, parts, chunk):
    s, d = synth.rmat_edges(9, 8, seed=42)
    g = srw.Graph.from_edges(s, d, None, flags=srw.BUILD_ALL)
    if chunk:
        monkeypatch.setenv("SRW_SAVE_CHUNK_BYTES", chunk)
    for sampler, directed in (("fold", False), ("exact", False)):
        a, b = str(tmp_path / ("a_%s" % sampler)), str(tmp_path / ("b_%s" % sampler))
        prm = dict(walkLength=20, numWalks=3, p=0.5, q=2.0, seed=5, sampler=sampler, singleOutput=single, rddPartitions=parts)
        g.walk(srw.Params(output=a, **prm)).save(srw.Params(output=a, **prm))
        wi = g.walk_save(srw.Params(output=b, **prm))
        fa, fb = sorted(os.listdir(os.path.join(a, "path"))), sorted(os.listdir(os.path.join(b, "path")))
        assert fa == fb and "_SUCCESS" in fb and len([f for f in fb if f.startswith("part-")]) == (1 if single else parts)
        for f in fa:
            assert open(os.path.join(a, "path", f), "rb").read() == open(os.path.join(b, "path", f), "rb").read(), f
        assert wi.steps == 3 * g.stats()[0] * 21
    with pytest.raises(srw.SrwError):          # Hadoop refuses an existing output directory
        g.walk_save(srw.Params(output=b, **prm))


def test_walk_save_ragged_paths_directed(tmp_path, oracle):
    """Dead ends give short lines (RW:115-119): the streamed writer == the oracle's text, line for line."""
    out = str(tmp_path / "o")
    g = srw.Graph.load(srw.Params(input=KARATE, directed=True), flags=srw.BUILD_ALL)
    g.walk_save(srw.Params(output=out, directed=True, walkLength=10, numWalks=2, p=0.5, q=2.0, seed=3, sampler="exact"))
    og = oracle.Graph().load_file(KARATE, directed=True)
    ids, offs = oracle.walk(og, walk_length=10, num_walks=2, p=0.5, q=2.0, seed=3)
    assert open(os.path.join(out, "path", "part-00000"), "rb").read() == oracle.format_paths(ids, offs)


def test_round_trip_at_c2_size():
    """parse(format(edges)) == edges on RMAT-20 (16.8 M lines, ~230 MB of text): a size-independent property of
    the two device text kernels together, on BASELINE config C2's edge list."""
    import torch
    scale, ef = 20, 16
    n = ef << scale
    e = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    ds = torch.empty(n, dtype=torch.int32, device="cuda")
    dd = torch.empty(n, dtype=torch.int32, device="cuda")
    srw.check(srw.lib().srw_synth_rmat_device(scale, ef, 42, 0, n, ds.data_ptr(), dd.data_ptr()))
    e[:, 0], e[:, 1] = ds, dd
    lens = torch.full((n,), 2, dtype=torch.int32, device="cuda")
    need = srw.format_paths_device(e.data_ptr(), lens.data_ptr(), n, 2)
    text = torch.empty(need, dtype=torch.uint8, device="cuda")
    assert srw.format_paths_device(e.data_ptr(), lens.data_ptr(), n, 2, text.data_ptr(), need) == need
    host = text.cpu().numpy().tobytes()
    assert host[:200].decode().split("\n")[0] == "%d\t%d" % (int(ds[0]), int(dd[0]))
    ps, pd, pw, _ = srw.parse_edges(data=host, weighted=False, device=True)
    assert np.array_equal(ps, ds.cpu().numpy()) and np.array_equal(pd, dd.cpu().numpy()) and (pw == 1.0).all()
