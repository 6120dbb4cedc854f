#!/usr/bin/env python
"""Builds experiment variants of librt_b200.so side by side (build/lib_<name>.so): the objects that do not change
are compiled once, rt_trace.cu once per variant with its -D switches, all in parallel.

  python tools/build_variants.py name1:-DRT_X_FOO=0,-DRT_X_BAR=1 name2:...
Select one at run time with RT_B200_LIB=build/lib_<name>.so (raytracinginonesemester_b200/api.py)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracinginonesemester_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "build", "variants")
os.makedirs(OUT, exist_ok=True)
FLAGS = [f for f in B.NVCC_FLAGS if f != "-shared"]


def cc(src, obj, extra=()):
    cmd = [B.nvcc_path()] + FLAGS + list(extra) + ["-c", os.path.join(B.CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout, r.stderr)
        raise SystemExit("compile failed: " + src)
    return obj


def main():
    variants = []
    for a in sys.argv[1:]:
        name, _, defs = a.partition(":")
        variants.append((name, [d for d in defs.split(",") if d]))
    common = [s for s in B.SOURCES if s != "rt_trace.cu"]
    jobs = []
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        cobjs = [ex.submit(cc, s, os.path.join(OUT, os.path.basename(s) + ".o")) for s in common]
        for name, defs in variants:
            jobs.append((name, ex.submit(cc, "rt_trace.cu", os.path.join(OUT, "rt_trace_%s.o" % name), defs + ["-Xptxas", "-v"])))
        cobjs = [f.result() for f in cobjs]
        for name, f in jobs:
            obj = f.result()
            lib = os.path.join(ROOT, "build", "lib_%s.so" % name)
            subprocess.run([B.nvcc_path(), "-shared", "-o", lib, obj] + cobjs + ["-ldl", "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
            print(lib)


if __name__ == "__main__":
    main()
