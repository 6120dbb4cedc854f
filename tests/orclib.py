"""ctypes bindings of the TEST oracle (oracle/), the in-place reference shims (oracle/_ref/) and
the host emulation of the device functions (tests/emul/).  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from raytracinginonesemester_b200 import _abi as A  # noqa: E402

import build as oracle_build  # noqa: E402  (oracle/build.py)


class orc_counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays_primary", "rays_shadow", "node_pops", "box_tests", "tri_tests", "max_stack")]


_cache = {}


def oracle():
    if "orc" not in _cache:
        lib = C.CDLL(oracle_build.build_oracle())
        lib.orc_camera_init.argtypes = [C.POINTER(A.rt_camera), A.f32p, A.f32p, A.f32p, C.c_double, C.c_double, C.c_int, C.c_int]
        lib.orc_jitter_table.argtypes = [A.f32p, C.c_int, C.c_uint32, C.c_int]
        lib.orc_bvh_build.restype = C.c_void_p
        lib.orc_bvh_build.argtypes = [C.POINTER(A.rt_scene)]
        lib.orc_bvh_free.argtypes = [C.c_void_p]
        lib.orc_bvh_export.argtypes = [C.c_void_p, A.u32p, A.f32p]
        lib.orc_render.argtypes = [C.POINTER(A.rt_scene), C.POINTER(A.rt_frame), C.c_void_p, C.POINTER(A.rt_image),
                                   C.c_int, C.c_int, C.c_int, C.POINTER(orc_counters)]
        lib.orc_ray_triangle.argtypes = [C.c_int, A.f32p, A.f32p, C.c_int, A.f32p, A.f32p, A.f32p, A.f32p]
        lib.orc_quantise.restype = C.c_uint8
        lib.orc_quantise.argtypes = [C.c_float, C.c_int]
        _cache["orc"] = lib
    return _cache["orc"]


def ref_libs():
    """dict name -> CDLL for the reference shims that exist (built here when /root/reference is present)."""
    if "ref" not in _cache:
        paths = oracle_build.build_ref()
        _cache["ref"] = {k: C.CDLL(v) for k, v in paths.items()}
    return _cache["ref"]


def emul():
    if "emu" not in _cache:
        here = os.path.join(ROOT, "tests", "emul")
        out = os.path.join(here, "libemul_host.so")
        src = os.path.join(here, "emul_host.cpp")
        csrc = os.path.join(ROOT, "raytracinginonesemester_b200", "csrc")
        deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
        if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
            subprocess.run(["g++", "-x", "c++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src], check=True)
        lib = C.CDLL(out)
        lib.emu_build.restype = C.c_void_p
        lib.emu_build.argtypes = [C.POINTER(A.rt_scene), C.c_uint32]
        lib.emu_adopt.restype = C.c_void_p
        lib.emu_adopt.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(A.rt_scene)]
        lib.emu_free.argtypes = [C.c_void_p]
        lib.emu_num_nodes.restype = C.c_uint32
        lib.emu_num_nodes.argtypes = [C.c_void_p]
        lib.emu_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.emu_validate.argtypes = [C.c_void_p]
        lib.emu_render.argtypes = [C.c_void_p, C.POINTER(A.rt_frame), C.POINTER(A.rt_image), C.POINTER(C.c_uint64)]
        _cache["emu"] = lib
    return _cache["emu"]


def _planes(fr, want):
    H, W = fr.height, fr.width
    img = A.rt_image()
    out = {}
    if "rgb" in want:
        out["rgb"] = np.zeros((H, W, 3), np.float32); img.rgb = out["rgb"].ctypes.data_as(A.f32p)
    if "rgb8" in want:
        out["rgb8"] = np.zeros((H, W, 3), np.uint8); img.rgb8 = out["rgb8"].ctypes.data_as(A.u8p)
    if "tri_id" in want:
        out["tri_id"] = np.full((H, W), -2, np.int32); img.tri_id = out["tri_id"].ctypes.data_as(A.i32p)
    if "t" in want:
        out["t"] = np.full((H, W), -2, np.float32); img.t = out["t"].ctypes.data_as(A.f32p)
    return img, out


def oracle_render(scene, frame, bvh=None, threads=8, row_begin=0, row_step=1, want=("rgb", "rgb8", "tri_id", "t")):
    """scene/frame: raytracinginonesemester_b200.api.Scene / Frame.  bvh: handle from oracle_bvh()
    (reference-exact LBVH + SearchBVH) or None (canonical brute force)."""
    lib = oracle()
    s, f = scene.c_struct(), frame.c_struct()
    img, out = _planes(frame, want)
    cnt = orc_counters()
    rc = lib.orc_render(C.byref(s), C.byref(f), bvh, C.byref(img), threads, row_begin, row_step, C.byref(cnt))
    assert rc == 0, rc
    out["counters"] = {n: getattr(cnt, n) for n, _ in orc_counters._fields_}
    return out


def oracle_bvh(scene):
    s = scene.c_struct()
    return oracle().orc_bvh_build(C.byref(s))


def emul_build(scene, leaf_max=4):
    s = scene.c_struct()
    return emul().emu_build(C.byref(s), leaf_max)


def emul_render(handle, frame, want=("rgb", "rgb8", "tri_id", "t")):
    f = frame.c_struct()
    img, out = _planes(frame, want)
    stats = (C.c_uint64 * 5)()
    rc = emul().emu_render(handle, C.byref(f), C.byref(img), stats)
    assert rc == 0
    out["stats"] = dict(rays_primary=stats[0], rays_shadow=stats[1], nodes=stats[2], tris=stats[3], max_stack=stats[4])
    return out


def reference_render(scene, frame, row_begin=0, row_step=1, threads=None, want=("rgb", "tri_id", "t")):
    """The frame through the REFERENCE ITSELF (oracle/_ref/libref_hw2.so: the unmodified GPUandCPU sources compiled in place,
    Camera::get_ray + SearchBVH + TraceRayIterative per pixel over its own LBVH), threaded over rows; falls back to the
    oracle's restatement of the same path (reference LBVH + SearchBVH) when the shim is not on this box.
    Returns (planes dict, "reference" | "port").  HW2-BVH frames with a camera made by api.camera_init only."""
    threads = threads or os.cpu_count() or 1
    # the reference's pixel loop always jitters with jittered_samples(spp, 42u) (query.cu:142-148): the frame must do the same
    want_j = np.zeros((frame.spp, 2), np.float32)
    oracle().orc_jitter_table(want_j.ctypes.data_as(A.f32p), frame.spp, 42, 1)
    assert frame.jitter is not None and np.array_equal(np.asarray(frame.jitter, np.float32).reshape(-1, 2), want_j), \
        "reference_render: the frame must use the reference's jitter table (api.jitter_table(spp, 42, True))"
    libs = ref_libs()
    if "ref_hw2" not in libs:
        out = oracle_render(scene, frame, bvh=oracle_bvh(scene), threads=threads, row_begin=row_begin, row_step=row_step,
                            want=tuple(w for w in want if w != "rgb8") + (("rgb8",) if "rgb8" in want else ()))
        return out, "port"
    lib = libs["ref_hw2"]
    lib.ref_hw2_world.restype = C.c_void_p
    lib.ref_hw2_build.restype = C.c_double
    f32p, u32p, i32p = A.f32p, A.u32p, A.i32p
    nrm = scene.normals.ctypes.data_as(f32p) if scene.normals is not None else None
    obj = scene.tri_obj_ids.ctypes.data_as(i32p) if scene.tri_obj_ids is not None else None
    h = lib.ref_hw2_world(scene.positions.ctypes.data_as(f32p), nrm, C.c_uint64(scene.positions.shape[0]),
                          scene.indices.ctypes.data_as(u32p), C.c_uint64(scene.indices.shape[0]), obj)
    lib.ref_hw2_build(C.c_void_p(h))
    W, H = frame.width, frame.height
    cp = frame.cam.params
    f3 = lambda v: np.array(v, np.float32)
    cpos, look, up, ms = f3(cp["pos"]), f3(cp["look_at"]), f3(cp["up"]), f3(frame.miss_color)
    out = {}
    rgb = tid = tt = None
    if "rgb" in want:
        out["rgb"] = np.zeros((H, W, 3), np.float32); rgb = out["rgb"].ctypes.data_as(f32p)
    if "tri_id" in want:
        out["tri_id"] = np.full((H, W), -2, np.int32); tid = out["tri_id"].ctypes.data_as(i32p)
    if "t" in want:
        out["t"] = np.full((H, W), -2, np.float32); tt = out["t"].ctypes.data_as(f32p)
    mats = scene.materials or []
    marr = (A.rt_material * max(1, len(mats)))(*mats)
    larr = (A.rt_light * max(1, len(frame.lights)))(*frame.lights)
    lib.ref_hw2_render_rows(C.c_void_p(h), cpos.ctypes.data_as(f32p), look.ctypes.data_as(f32p), up.ctypes.data_as(f32p),
                            C.c_double(cp["focal_mm"]), C.c_double(cp["sensor_mm"]), W, H, ms.ctypes.data_as(f32p), int(frame.max_depth), int(frame.spp),
                            marr, len(mats), larr, len(frame.lights), int(frame.diffuse_bounce), row_begin, row_step, threads, rgb, tid, tt)
    lib.ref_hw2_free(C.c_void_p(h))
    return out, "reference"
