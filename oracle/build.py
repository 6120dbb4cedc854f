"""Build recipe for the test oracle and the in-place reference shims.

TEST INFRASTRUCTURE ONLY.  Produces
  oracle/librt_oracle.so          our CPU restatement (oracle/rt_oracle.c)
  oracle/_ref/libref_hw1.so       reference HW1 sources compiled where they lie
  oracle/_ref/libref_hw2.so       reference HW2/GPUandCPU sources (CPU build) compiled where they lie
  oracle/_ref/libref_ppm.so       reference ppm_p6_lib compiled where it lies
  oracle/_ref/libref_hw2_main.so  the reference's whole bvh_viz program (main renamed), for end-to-end fixtures
  oracle/_ref/libref_cpuonly.so   reference HW2/CPUOnly sources (TraceRay, camera, transform, OBJ loader) compiled where they lie
  oracle/_ref/libref_hw2_cuda.so  reference HW2/GPUandCPU sources compiled as CUDA (its ENABLE_GPU build) for sm_100: the
                                  reference's own GPU renderer, timed next to the product by bench.py ("reference_cuda")
The _ref outputs need /root/reference (present in the authoring container only); on the GPU
box the prebuilt files travel with the snapshot.  No reference source is copied into the repo.
Flags: -O2 -ffp-contract=off, no -march=native (hit ids depend on no-FMA rounding).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RT_REFERENCE_ROOT", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")
FP = ["-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread"]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in srcs)


def build_oracle(force=False):
    out = os.path.join(HERE, "librt_oracle.so")
    srcs = [os.path.join(HERE, "rt_oracle.c"), os.path.join(HERE, "..", "include", "rt_api.h")]
    if force or _stale(out, srcs):
        _run(["gcc", "-std=gnu11"] + FP + ["-o", out, srcs[0], "-lm"])
    return out


def have_reference():
    return os.path.isdir(os.path.join(REF, "HW1", "include"))


def build_ref(force=False):
    """Compile the reference's own sources in place.  Returns dict name -> path (existing files)."""
    outs = {n: os.path.join(REF_OUT, "lib%s.so" % n) for n in ("ref_hw1", "ref_hw2", "ref_ppm", "ref_hw2_main", "ref_cpuonly", "ref_hw2_cuda")}
    if have_reference():
        os.makedirs(REF_OUT, exist_ok=True)
        hw1, g = os.path.join(REF, "HW1"), os.path.join(REF, "HW2", "HW2", "GPUandCPU")
        s = os.path.join(HERE, "ref_shim_hw1.cpp")
        if force or _stale(outs["ref_hw1"], [s]):
            _run(["g++", "-std=c++17", "-w"] + FP + ["-I", os.path.join(hw1, "include"), "-o", outs["ref_hw1"],
                  s, os.path.join(hw1, "src", "MeshOBJ.cpp")])
        s = os.path.join(HERE, "ref_shim_hw2.cpp")
        if force or _stale(outs["ref_hw2"], [s]):
            inc = ["-I", os.path.join(g, "third_party", "glm"), "-I", os.path.join(g, "include"), "-I", os.path.join(g, "src")]
            _run(["g++", "-x", "c++", "-std=c++14", "-w", "-D__device__="] + FP + inc + ["-o", outs["ref_hw2"],
                  s, os.path.join(g, "include", "bvh.cu"), os.path.join(g, "include", "query.cu")])
        s = os.path.join(HERE, "ref_shim_hw2_main.cpp")
        if force or _stale(outs["ref_hw2_main"], [s]):
            inc = ["-I", os.path.join(g, "third_party", "glm"), "-I", os.path.join(g, "include"), "-I", os.path.join(g, "src")]
            _run(["g++", "-x", "c++", "-std=c++14", "-w", "-D__device__="] + FP + inc + ["-o", outs["ref_hw2_main"],
                  s, os.path.join(g, "include", "bvh.cu"), os.path.join(g, "include", "query.cu")])
        s = os.path.join(HERE, "ref_shim_cpuonly.cpp")
        if force or _stale(outs["ref_cpuonly"], [s]):
            c = os.path.join(REF, "HW2", "HW2", "CPUOnly")
            _run(["g++", "-std=c++17", "-w"] + FP + ["-I", os.path.join(c, "include"), "-o", outs["ref_cpuonly"],
                  s, os.path.join(c, "src", "MeshOBJ.cpp")])
        s = os.path.join(HERE, "ref_shim_ppm.cpp")
        if force or _stale(outs["ref_ppm"], [s]):
            p = os.path.join(hw1, "ppm_p6_lib")
            _run(["g++", "-std=c++17", "-w"] + FP + ["-I", os.path.join(p, "include"), "-o", outs["ref_ppm"],
                  s, os.path.join(p, "src", "ppm_p6.cpp")])
        s = os.path.join(HERE, "ref_shim_hw2_cuda.cu")
        if force or _stale(outs["ref_hw2_cuda"], [s]):
            # the reference's own CUDA flags (GPUandCPU/CMakeLists.txt:27); it names no architecture, sm_100 is ours
            inc = ["-I", os.path.join(g, "third_party", "glm"), "-I", os.path.join(g, "include"), "-I", os.path.join(g, "src")]
            try:        # timing aid only (bench.py "reference_cuda"): a failure here must not fail the build of the checker
                _run(["nvcc", "-std=c++17", "-O3", "-w", "--extended-lambda", "--expt-relaxed-constexpr", "--use_fast_math",
                      "-gencode", "arch=compute_100,code=sm_100", "-Xcompiler", "-fPIC", "-shared"] + inc +
                     ["-o", outs["ref_hw2_cuda"], s, os.path.join(g, "include", "bvh.cu"), os.path.join(g, "include", "query.cu")])
            except (RuntimeError, OSError) as e:
                sys.stderr.write("oracle/build.py: reference CUDA build skipped: %s\n" % str(e)[:400])
    return {k: v for k, v in outs.items() if os.path.exists(v)}


if __name__ == "__main__":
    print(build_oracle(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv))
