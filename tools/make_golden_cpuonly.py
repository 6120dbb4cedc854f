"""Fixtures for RT_MODE_HW2_CPU, produced by the UNMODIFIED HW2/HW2/CPUOnly reference sources compiled
in place (oracle/_ref/libref_cpuonly.so): its OBJ loader + ApplyTransformToMeshSOA, camera, TraceRay.

  tests/golden/cpuonly_scenes.npz
    sphere_point_*  config/sphere_point.json (962 triangles, hard shadow, BRDF, sky): the baked mesh, materials, camera and
                    light, the reference's committed golden output/sphere_point_output.png (360x240, bit-exact with HEAD) and
                    a 120x80 render (rgb f32, tri_id, t)
    sphere_*        config/sphere.json (4 802 triangles, mirror spheres, max_bounces 4, diffuse_bounce false) at 120x80
    cornell_*       cornellbox.obj (no vertex normals -> per-triangle face normals, render.cpp:88-96), mirror floor, 96x72

Authoring container only (needs /root/reference).  Data only; no reference source is copied.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orclib  # noqa: E402
from raytracinginonesemester_b200 import _abi as A, api  # noqa: E402

R = "/root/reference/HW2/HW2/CPUOnly"
lib = orclib.ref_libs()["ref_cpuonly"]
lib.ref_cpu_load_obj.restype = C.c_void_p
fp = lambda a: a.ctypes.data_as(A.f32p)
f3 = lambda v: np.array(v, np.float32)


class L(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("color", C.c_float * 3), ("intensity", C.c_float)]


def mat13(m):
    return np.array(list(m.albedo) + [m.kd] + list(m.specular_color) + [m.ks, m.shininess, m.kr] + list(m.emission), np.float32)


def load_scene(cfg):
    j = json.load(open(os.path.join(R, "config", cfg)))
    P, N, I, O, M, off = [], [], [], [], [], 0
    for k, node in enumerate(j["scene"]):
        h = lib.ref_cpu_load_obj(os.path.join(R, node["path"]).encode())
        assert h
        t = node.get("transform", {})
        lib.ref_cpu_mesh_transform(C.c_void_p(h), fp(f3(t.get("position", [0, 0, 0]))), fp(f3(t.get("rotation", [0, 0, 0]))), fp(f3(t.get("scale", [1, 1, 1]))))
        nv, nn, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.ref_cpu_mesh_counts(C.c_void_p(h), C.byref(nv), C.byref(nn), C.byref(nt))
        pos = np.zeros((nv.value, 3), np.float32); nrm = np.zeros((nn.value, 3), np.float32); idx = np.zeros((nt.value, 3), np.uint32)
        lib.ref_cpu_mesh_copy(C.c_void_p(h), fp(pos), fp(nrm), idx.ctypes.data_as(A.u32p))
        lib.ref_cpu_mesh_free(C.c_void_p(h))
        assert nn.value == nv.value
        P.append(pos); N.append(nrm); I.append(idx + off); O.append(np.full(nt.value, k, np.int32)); off += nv.value
        m = node.get("material", {})
        M.append(api.make_material(albedo=m.get("albedo", (0.8, 0.8, 0.8)), kd=m.get("kd", 1.0), ks=m.get("ks", 0.0), shininess=m.get("shininess", 32.0),
                                   specular_color=m.get("specular_color", (0.04, 0.04, 0.04)), kr=m.get("kr", 0.0), emission=m.get("emission", (0, 0, 0))))
    return j, np.concatenate(P), np.concatenate(N), np.concatenate(I), np.concatenate(O), M


def render(pos, nrm, idx, obj, M, cam, light, W, H, depth):
    """cam = (pos, look, up, focal, sensor_h, sensor_w); light = (pos, color, intensity)."""
    l = L(); l.position[:] = light[0]; l.color[:] = light[1]; l.intensity = light[2]
    rgb = np.zeros((H, W, 3), np.float32); tid = np.zeros((H, W), np.int32); tt = np.zeros((H, W), np.float32)
    marr = (A.rt_material * len(M))(*M)
    lib.ref_cpu_render_rows(fp(pos), fp(nrm) if nrm is not None else None, C.c_uint64(len(pos)), idx.ctypes.data_as(A.u32p), C.c_uint64(len(idx)),
                            obj.ctypes.data_as(A.i32p), marr, len(M), fp(f3(cam[0])), fp(f3(cam[1])), fp(f3(cam[2])), C.c_double(cam[3]), C.c_double(cam[4]),
                            C.c_double(cam[5]), W, H, C.byref(l), 1, depth, 0, 1, fp(rgb), tid.ctypes.data_as(A.i32p), fp(tt))
    return rgb, tid, tt


def pack(out, name, pos, nrm, idx, obj, M, cam, light, W, H, depth, rgb, tid, tt):
    out[name + "_positions"], out[name + "_indices"], out[name + "_tri_obj_ids"] = pos, idx, obj
    out[name + "_normals"] = nrm if nrm is not None else np.zeros((0, 3), np.float32)
    out[name + "_materials"] = np.stack([mat13(m) for m in M])
    out[name + "_camera"] = np.array(list(cam[0]) + list(cam[1]) + list(cam[2]) + [cam[3], cam[4], cam[5]], np.float64)
    out[name + "_light"] = np.array(list(light[0]) + list(light[1]) + [light[2]], np.float64)
    out[name + "_frame"] = np.array([W, H, depth])
    out[name + "_rgb"], out[name + "_tri_id"], out[name + "_t"] = rgb, tid, tt


def main():
    from PIL import Image
    out = {}
    for cfg, name in (("sphere_point.json", "sphere_point"), ("sphere.json", "sphere")):
        j, pos, nrm, idx, obj, M = load_scene(cfg)
        c, li = j["camera"], j["light"]
        cam = (c["position"], c["look_at"], c["up"], c["focal_length_mm"], c.get("sensor_height_mm", 24.0), c.get("sensor_width_mm", 36.0))
        light = (li["position"], li["color"], li["intensity"])
        depth = j["settings"]["max_bounces"]
        W, H = 120, 80
        rgb, tid, tt = render(pos, nrm, idx, obj, M, cam, light, W, H, depth)
        pack(out, name, pos, nrm, idx, obj, M, cam, light, W, H, depth, rgb, tid, tt)
        print(name, len(idx), "tris, hit", (tid >= 0).mean())
        if name == "sphere_point":       # the reference's committed golden image, reproduced bit-exactly by HEAD
            full, _, _ = render(pos, nrm, idx, obj, M, cam, light, c["pixel_width"], c["pixel_height"], depth)
            q = np.zeros(full.shape, np.uint8)
            lib.ref_cpu_quantise(fp(full), C.c_uint64(full.shape[0] * full.shape[1]), q.ctypes.data_as(A.u8p))
            png = np.array(Image.open(os.path.join(R, "output", "sphere_point_output.png")).convert("RGB"))
            assert np.array_equal(q, png), "reference HEAD no longer reproduces its committed golden"
            out["sphere_point_golden_png"] = png
    d = np.load(os.path.join(ROOT, "tests", "golden", "cornell_mesh.npz"))
    nobj = int(d["tri_obj_ids"].max()) + 1
    M = [api.make_material(albedo=(0.7, 0.7, 0.7), kd=0.9, ks=0.2, kr=0.0) for _ in range(nobj)]
    M[0] = api.make_material(albedo=(0.6, 0.6, 0.7), kd=0.5, ks=0.3, kr=0.6, specular_color=(0.9, 0.9, 0.9), shininess=64.0)
    M[2] = api.make_material(albedo=(0.8, 0.1, 0.1), kd=0.8, kr=0.3, specular_color=(0.8, 0.6, 0.6), emission=(0.02, 0.0, 0.0))
    cam = ((278.0, 273.0, -800.0), (278.0, 273.0, 0.0), (0.0, 1.0, 0.0), 35.0, 24.0, 32.0)
    light = ((278.0, 500.0, 279.5), (1.0, 0.95, 0.9), 1.5)
    W, H, depth = 96, 72, 5
    rgb, tid, tt = render(d["positions"], None, d["indices"], d["tri_obj_ids"], M, cam, light, W, H, depth)
    pack(out, "cornell", d["positions"], None, d["indices"], d["tri_obj_ids"], M, cam, light, W, H, depth, rgb, tid, tt)
    print("cornell hit", (tid >= 0).mean(), "mean", rgb.mean((0, 1)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cpuonly_scenes.npz"), **out)


if __name__ == "__main__":
    main()
