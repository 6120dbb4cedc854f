"""The device OBJ parser's per-line functions (csrc/rt_obj_core.h) compiled for the host (tests/emul): its strtof against
glibc's strtof on a few hundred thousand literals, and its face-corner parser on the loader's syntax and quirks.
The whole parser is compared with the host loader on the GPU (tests/test_gpu_ingest.py)."""
import ctypes as C

import numpy as np

import orclib


def sweep(literals):
    lib = orclib.emul()
    buf = ("\n".join(literals) + "\n").encode()
    out = (C.c_longlong * 7)()
    lib.emu_obj_real_sweep(buf, C.c_uint64(len(buf)), out)
    return dict(ok=out[0], hard=out[1], refused=out[2], none=out[3], bad_value=out[4], bad_end=out[5], first_bad=out[6])


def test_device_strtof_equals_glibc_strtof():
    rng = np.random.default_rng(7)
    lits = []
    x = rng.normal(size=60000) * 10.0 ** rng.integers(-6, 7, 60000)
    lits += ["%.6f" % v for v in x[:20000]]                                   # what OBJ writers emit
    lits += ["%.9g" % v for v in x[20000:40000]]
    lits += [repr(float(np.float32(v))) for v in x[40000:50000]]              # shortest round-trip forms
    lits += ["%.3e" % v for v in x[50000:55000]] + ["%+.5E" % v for v in x[55000:60000]]
    ints = rng.integers(-2**40, 2**40, 5000)
    lits += [str(int(v)) for v in ints] + ["%d." % v for v in ints[:500]] + [".%d" % abs(v) for v in ints[:500]]
    n_typical = len(lits)
    r = sweep(lits)
    assert r["bad_value"] == 0 and r["bad_end"] == 0, (r, lits[r["first_bad"]])
    assert r["ok"] + r["hard"] == n_typical and r["refused"] == 0 and r["none"] == 0
    assert r["hard"] <= 2e-3 * n_typical                                        # almost everything converts on the device
    # the literals fp64 cannot decide, and the ones past its exact range: handed to the host, never converted wrongly
    hard = ["16777217", "16777217.0", "0.1234567890123456789", "1e23", "1e-30", "123456789012345678901234567890", "1e-45", "3.4028235e38",
            "1.17549435e-38", "1e39", "1e-60", "4.9e-324", "9007199254740993", "0.000000000000000000000000000000000000011754942"]
    r = sweep(hard)
    assert r["bad_value"] == 0 and r["bad_end"] == 0 and r["hard"] >= 6 and r["ok"] + r["hard"] == len(hard), r
    # floats on and next to every kind of rounding boundary: the midpoints of consecutive floats written out exactly
    f = np.abs(rng.normal(size=4000).astype(np.float32)) + np.float32(0.5)
    nxt = np.nextafter(f, np.float32(np.inf))
    mids = (f.astype(np.float64) + nxt.astype(np.float64)) / 2
    lits = ["%.30g" % m for m in mids] + ["%.17g" % np.nextafter(m, 0.0) for m in mids] + ["%.17g" % np.nextafter(m, np.inf) for m in mids]
    r = sweep(lits)
    assert r["bad_value"] == 0 and r["bad_end"] == 0, (r, lits[r["first_bad"]])
    assert r["hard"] >= 4000                                                    # the exact midpoints at least
    # syntax: what ends a literal, what is no literal at all, what is refused
    odd = ["1.5x", "1e", "1e+", "1e+x", "1.2.3", "--1", "+-1", "-", "+", ".", "e5", "abc", "", " 7", "\t-0", "-0.0", "+0", "0e0", "00012.500", "1E5", "1e05",
           "5e-1\r", "1,5", "1 2", "0x", "0xg", "1f"]
    r = sweep(odd)
    assert r["bad_value"] == 0 and r["bad_end"] == 0, (r, odd[r["first_bad"]])
    assert r["none"] == sum(1 for s in odd if s.strip(" \t") in ("--1", "+-1", "-", "+", ".", "e5", "abc", ""))
    r = sweep(["inf", "-Infinity", "nan", "NAN(1)", "0x1p3", "0X.8", "-0x10"])
    assert r["refused"] == 7 and r["bad_value"] == 0


def face(line, nv=10, nt=5, nn=4):
    lib = orclib.emul()
    b = line.encode()
    out = (C.c_int * 12)()
    n = lib.emu_obj_face(b, len(b), nv, nt, nn, out)
    return [tuple(out[3 * j:3 * j + 3]) for j in range(n)]


def test_face_corner_parser_follows_the_loader():
    assert face(" 1 2 3\n") == [(0, -1, -1), (1, -1, -1), (2, -1, -1)]
    assert face(" 1/2 3/4 5/1\n") == [(0, 1, -1), (2, 3, -1), (4, 0, -1)]
    assert face(" 1//2 3//4 5//1\r\n") == [(0, -1, 1), (2, -1, 3), (4, -1, 0)]
    assert face("\t1/2/3\t4/5/1 2/2/2  7/1/4") == [(0, 1, 2), (3, 4, 0), (1, 1, 1), (6, 0, 3)]
    assert face(" -1 -2 -10\n") == [(9, -1, -1), (8, -1, -1), (0, -1, -1)]                # relative to the counts so far
    assert face(" -1/-1/-1 -2/-2/-2 -3/-3/-3\n") == [(9, 4, 3), (8, 3, 2), (7, 2, 1)]
    assert face(" 1 2 3 4 5 6\n") == [(0, -1, -1), (1, -1, -1), (2, -1, -1), (3, -1, -1)]  # at most four corners
    assert face(" 1 2 x 3\n") == [(0, -1, -1), (1, -1, -1)]                                # stops at what is not a corner
    assert face(" 1 2 3 # c\n") == [(0, -1, -1), (1, -1, -1), (2, -1, -1)]
    assert face(" 1// 2 3\n") == [(0, -1, 1), (2, -1, -1)]                                 # integer() skips blanks: "1// 2" is v 1, vn 2
    assert face(" 5/ 6/ 7/\n") == [(4, 5, 6)]                                              # ... and "5/ 6/ 7/" is ONE corner
    assert face(" 1/2/ 3 4\n") == [(0, 1, 2), (3, -1, -1)]
    assert face(" 0 1 2\n") == [(-1, -1, -1), (0, -1, -1), (1, -1, -1)]                    # index 0 resolves to -1 (refused later)
    assert face("\n") == [] and face("") == []
