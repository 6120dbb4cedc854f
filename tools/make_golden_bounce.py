"""Fixtures for TraceRayIterative's bounce loop, produced by running the unmodified reference
compiled in place (oracle/_ref/libref_hw2.so): camera inside cornellbox.obj, mirror / mixed materials,
max_depth > 1, with and without diffuse_bounce, 1-3 samples per pixel.

  tests/golden/ref_hw2_bounce_cornell.npz   rgb / tri_id / t per case + the case table

Authoring container only (needs /root/reference).  Data only; no reference source is copied.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orclib  # noqa: E402
from raytracinginonesemester_b200 import _abi as A, scenes  # noqa: E402

CASES = [  # name, W, H, spp, max_depth, diffuse_bounce
    ("mirror_d4", 64, 48, 1, 4, 0),
    ("diffuse_d5_spp2", 64, 48, 2, 5, 1),
    ("diffuse_d8_spp3", 48, 36, 3, 8, 1),
]


def main():
    h2 = orclib.ref_libs()["ref_hw2"]
    h2.ref_hw2_world.restype = C.c_void_p
    sc, cam_args, lights, miss = scenes.cornell_bounce_scene(os.path.join(ROOT, "tests", "golden", "cornell_mesh.npz"))
    fp = lambda a: a.ctypes.data_as(A.f32p)
    w = h2.ref_hw2_world(fp(sc.positions), fp(sc.normals) if sc.normals is not None else None, C.c_uint64(sc.positions.shape[0]),
                         sc.indices.ctypes.data_as(A.u32p), C.c_uint64(sc.indices.shape[0]), sc.tri_obj_ids.ctypes.data_as(A.i32p))
    h2.ref_hw2_build(C.c_void_p(w))
    marr = (A.rt_material * len(sc.materials))(*sc.materials)
    larr = (A.rt_light * len(lights))(*lights)
    cp, lk, up = (np.array(v, np.float32) for v in cam_args[:3])
    ms = np.array(miss, np.float32)
    out = {}
    for name, W, H, spp, depth, diffuse in CASES:
        rgb = np.zeros((H, W, 3), np.float32); rgb2 = np.zeros((H, W, 3), np.float32)
        tid = np.zeros((H, W), np.int32); tt = np.zeros((H, W), np.float32)
        # the reference's own render() and the row driver of the same per-pixel body must agree
        h2.ref_hw2_render(C.c_void_p(w), fp(cp), fp(lk), fp(up), C.c_double(cam_args[3]), C.c_double(cam_args[4]), W, H, fp(ms), depth, spp,
                          marr, len(sc.materials), larr, len(lights), diffuse, fp(rgb))
        h2.ref_hw2_render_rows(C.c_void_p(w), fp(cp), fp(lk), fp(up), C.c_double(cam_args[3]), C.c_double(cam_args[4]), W, H, fp(ms), depth, spp,
                               marr, len(sc.materials), larr, len(lights), diffuse, 0, 1, 8, fp(rgb2), tid.ctypes.data_as(A.i32p), fp(tt))
        assert np.array_equal(rgb, rgb2), name
        out[name + "_rgb"], out[name + "_tri_id"], out[name + "_t"] = rgb, tid, tt
        print(name, "hit px", int((tid >= 0).sum()), "of", tid.size, "mean rgb", rgb.mean((0, 1)))
    h2.ref_hw2_free(C.c_void_p(w))
    out["cases"] = np.array([(n, W, H, s, d, df) for n, W, H, s, d, df in CASES], dtype="U32")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_hw2_bounce_cornell.npz"), **out)


if __name__ == "__main__":
    main()
