"""In-tree build of the sm_100a shared library (librt_b200.so) with nvcc.

The library is the product: hand-written CUDA kernels + the C ABI of include/rt_api.h.
Built in-tree so the .so travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "librt_b200.so")
SOURCES = ["rt_api.cu", "rt_build.cu", "rt_trace.cu", "rt_obj_device.cu", "rt_host.cpp", "rt_mesh_api.cpp", os.path.join("..", "host", "mesh_ingest.cpp")]
HEADERS = ["rt_math.h", "rt_core.h", "rt_kernels.h", "rt_params.h", "rt_trace_core.h", "rt_build_core.h", "rt_obj_core.h",
           os.path.join("..", "host", "mesh_ingest.hpp"), os.path.join("..", "..", "include", "rt_api.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, extra=()):
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building librt_b200.so")
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose=True, extra=["-Xptxas", "-v"] if "-v" in sys.argv else ()))
