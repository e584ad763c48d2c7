"""Text rules of the device parser / formatter (csrc/text_io.cuh) compiled for the host (tests/emu/emu_text.cpp)
and checked against (i) the serial host parser of libsrw, which the reference-KAT tests pin on URW:23-34 /
VRW:19-34, and (ii) the oracle's formatter (RW:234-241).  The `-m gpu` tests run the same rules in the kernels."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from conftest import KARATE, ROOT, TESTGRAPH

srw = importlib.import_module("stellar-random-walk_b200")
bld = importlib.import_module("stellar-random-walk_b200.build")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libsrw_emu_text.so")


@pytest.fixture(scope="module", autouse=True)
def built():
    bld.build()          # the serial host parser these tests compare against lives in libsrw.so (nvcc cross-compiles without a GPU)


@pytest.fixture(scope="module")
def emu():
    srcs = [os.path.join(EMU_DIR, "emu_text.cpp"), os.path.join(ROOT, "stellar-random-walk_b200", "csrc", "text_io.cuh")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(s) > os.path.getmtime(EMU_SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", srcs[0], "-o", EMU_SO])
    lib = C.CDLL(EMU_SO)
    lib.emu_parse_buffer.restype = C.c_int64
    lib.emu_parse_buffer.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    lib.emu_float_fast.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_float)]
    lib.emu_format_paths.restype = C.c_int64
    lib.emu_format_paths.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64]
    return lib


def emu_parse(emu, data, weighted, partitioned):
    if isinstance(data, str):
        data = data.encode()
    cap = data.count(b"\n") + data.count(b"\r") + 2
    s, d, p = (np.zeros(cap, np.int32) for _ in range(3))
    w = np.zeros(cap, np.float32)
    f = np.zeros(cap, np.uint8)
    n = emu.emu_parse_buffer(data, len(data), int(weighted), int(partitioned), s.ctypes.data, d.ctypes.data, w.ctypes.data,
                             p.ctypes.data, f.ctypes.data, cap)
    return s[:n], d[:n], w[:n], p[:n], f[:n]


def host_parse_line_weight(line, weighted, partitioned):
    _, _, w, _ = srw.parse_edges(data=line, weighted=weighted, partitioned=partitioned)
    return w[0]


def check_same_as_host(emu, text, weighted, partitioned):
    """Device rules == host parser on `text`: same edges; a line flagged for the host float path gets the host's
    weight by construction, so only its other fields are compared; an error must be the host's first error line."""
    s, d, w, p, f = emu_parse(emu, text, weighted, partitioned)
    bad = np.nonzero(f == 2)[0]
    try:
        hs, hd, hw, hp = srw.parse_edges(data=text, weighted=weighted, partitioned=partitioned)
    except srw.SrwError as ex:
        assert len(bad) > 0, "host parser failed (%s) but the device rules accepted every line" % ex
        assert ("line %d:" % (bad[0] + 1)) in str(ex)
        return "error"
    assert len(bad) == 0
    assert np.array_equal(s, hs) and np.array_equal(d, hd)
    if partitioned:
        assert np.array_equal(p, hp)
    fast = f == 0
    assert np.array_equal(w[fast].view(np.uint32), hw[fast].view(np.uint32))     # bit-exact, -0.0 included
    return int((f == 1).sum())


def test_reference_fixtures(emu):
    for path in (KARATE, TESTGRAPH):
        data = open(path, "rb").read()
        for weighted in (False, True):
            assert check_same_as_host(emu, data, weighted, False) == 0


@pytest.mark.parametrize("text,weighted,partitioned", [
    ("1 2\n2 3\n", True, False),
    ("1   2", True, False),                                # testgraph.txt: three spaces, no newline (T-URW:69-86)
    ("1\t2\t0.5\r\n3 4 1.25\r5 6 2\n", True, False),       # \r\n, bare \r, tabs
    ("1 2 7 0.5\n3 4 8\n5 6\n", True, True),               # VRW: pid column, weight only with > 3 columns
    ("1 2 x 0.5\n", True, True),                           # unparsable pid -> 0
    ("1 2 0.5 extra 0.25\n", True, False),                 # weight = LAST column
    ("-5 +7 1e-3\n2147483647 -2147483648 3\n", True, False),
    ("1 2 3\n", False, False),                             # unweighted: third column ignored
    ("1 2 \t \n", True, False),                            # trailing whitespace
    ("1 2 abc\n3 4 1.5.2\n5 6 --1\n7 8 1e\n9 10 .\n", True, False),     # not floats -> 1.0f (URW:31)
    ("1 2 0x1p3\n3 4 NaN\n5 6 Infinity\n7 8 -Infinity\n9 10 1.5f\n11 12 2D\n", True, False),    # host float path
    ("1 2 1.\n3 4 .5\n5 6 -0\n7 8 +0.0\n9 10 0e99\n11 12 00012.5000\n", True, False),
    ("1 2 123456789\n3 4 16777217\n5 6 0.1234567891\n7 8 1e11\n9 10 1e-11\n11 12 3.4e38\n13 14 1e-46\n", True, False),
    ("", True, False),
])
def test_lines_match_host_parser(emu, text, weighted, partitioned):
    assert check_same_as_host(emu, text, weighted, partitioned) != "error"


@pytest.mark.parametrize("text", ["1 2\n\n3 4\n", " 1 2\n", "1\n", "a b\n", "1 b\n", "2147483648 1\n", "1 2\n3 -2147483649\n",
                                  "1 2\r\n\r\n", "1.0 2\n", "\n", "1 2\n \n"])
def test_malformed_lines_are_the_hosts_errors(emu, text):
    assert check_same_as_host(emu, text, True, False) == "error"


def test_float_fast_path_is_correctly_rounded(emu):
    """Every token the fast path accepts equals strtof (glibc: correctly rounded = Float.parseFloat); the weights of
    BASELINE config C3 ("1.xyz") and ordinary decimals must all take the fast path."""
    rng = np.random.RandomState(3)
    toks = ["1.%03d" % k for k in range(0, 1000, 7)]
    for _ in range(4000):
        m = int(rng.randint(0, 1 << 24))
        frac = int(rng.randint(0, 9))
        sm = str(m)
        if frac:
            sm = sm.rjust(frac + 1, "0")
            sm = sm[:-frac] + "." + sm[-frac:]
        e = int(rng.randint(-6, 7))
        tok = ("-" if rng.rand() < 0.3 else "") + sm + (("e%d" % e) if rng.rand() < 0.4 else "")
        toks.append(tok)
    toks += ["0.1", "0.2", "0.3", "1e10", "16777215e-10", "1.0000000", "5e-10", "9999999e3", "0.000001", "123.456"]
    out = C.c_float()
    n_fast = 0
    for t in toks:
        r = emu.emu_float_fast(t.encode(), len(t), C.byref(out))
        want = host_parse_line_weight("1 2 %s\n" % t, True, False)
        if r == 1:
            n_fast += 1
            assert np.float32(out.value).view(np.uint32) == np.float32(want).view(np.uint32), t
    assert n_fast > 0.8 * len(toks)
    for t in ["1.%03d" % k for k in range(1000)] + ["0.5", "2", "10.25", "1e-3"]:
        assert emu.emu_float_fast(t.encode(), len(t), C.byref(out)) == 1, t


def test_random_edge_lists_match_host_parser(emu):
    rng = np.random.RandomState(11)
    ints = ["0", "1", "34", "-7", "+9", "2147483647", "-2147483648", "123456"]
    ws = ["1", "0.5", "1.250", "2e-2", "7.", ".125", "1e400", "abc", "0x1.8p1", "3f", "1_0", "1e+2", "-0.0", "12345678.9"]
    seps = [" ", "\t", "  ", " \t "]
    ends = ["\n", "\r\n", "\r"]
    for weighted in (True, False):
        for partitioned in (True, False):
            lines = []
            for _ in range(3000):
                cols = [ints[rng.randint(len(ints))], ints[rng.randint(len(ints))]]
                for _k in range(rng.randint(0, 4)):
                    cols.append((ints + ws)[rng.randint(len(ints) + len(ws))])
                sep = seps[rng.randint(len(seps))]
                lines.append(sep.join(cols) + ("" if rng.rand() < 0.9 else sep) + ends[rng.randint(len(ends))])
            text = "".join(lines)
            if text.endswith("\r"):
                text += "\n"
            r = check_same_as_host(emu, text, weighted, partitioned)
            assert r != "error"
            if weighted:
                assert r > 0          # some tokens did go to the host float path


def test_format_rules_match_oracle(emu, oracle):
    rng = np.random.RandomState(5)
    n, stride = 500, 12
    vals = np.concatenate([np.array([0, 1, 9, 10, 99, 100, 2147483647, -2147483648, -1, -10, 1000000000, 999999999], np.int32),
                           (10 ** rng.randint(0, 10, 200) * rng.randint(1, 10, 200) - rng.randint(0, 2, 200)).astype(np.int64).clip(-2**31, 2**31 - 1).astype(np.int32),
                           rng.randint(-2**31, 2**31 - 1, 400).astype(np.int32)])
    paths = vals[rng.randint(0, len(vals), (n, stride))].astype(np.int32)
    lens = rng.randint(1, stride + 1, n).astype(np.int32)
    need = emu.emu_format_paths(paths.ctypes.data, lens.ctypes.data, n, stride, None, 0)
    buf = C.create_string_buffer(int(need))
    assert emu.emu_format_paths(paths.ctypes.data, lens.ctypes.data, n, stride, buf, need) == need
    ids = np.concatenate([paths[i, :lens[i]] for i in range(n)])
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    assert buf.raw[:need] == oracle.format_paths(ids, offs)
    assert buf.raw[:need] == "".join("\t".join(str(int(v)) for v in paths[i, :lens[i]]) + "\n" for i in range(n)).encode()
