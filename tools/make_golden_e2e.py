"""End-to-end fixture: the reference's whole bvh_viz program (HW2/HW2/GPUandCPU/src/main.cu, CPU build, compiled in
place as oracle/_ref/libref_hw2_main.so) run on a scene JSON + OBJ files written by this script.

  tests/golden/e2e_scene.npz: the input texts (scene JSON, OBJ files), the 8-bit image the reference wrote
  (render.png, decoded), and the meshes as the reference loaders + applyObjectTransform produced them.

The GPU test writes the same files, runs rt_render_cli on them and compares the PPM with the reference's image;
the CPU test checks the C library's OBJ ingest (rt_mesh_*) against the reference-loaded arrays.
Authoring container only.  Data only; no reference source is copied.
"""
import ctypes as C
import json
import math
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orclib  # noqa: E402
from raytracinginonesemester_b200 import _abi as A  # noqa: E402


def sphere_obj(nu=14, nv=9):
    """UV sphere with vertex normals, texture coordinates, quads on the body (split by the loader), triangles at the
    poles, negative (relative) indices on one face and two `o` groups."""
    L = ["# synthetic test mesh", "o upper"]
    verts = []
    for j in range(nv + 1):
        th = math.pi * j / nv
        for i in range(nu):
            ph = 2 * math.pi * i / nu
            verts.append((math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)))
    for v in verts:
        L.append("v %.6f %.6f %.6f" % v)
    for j in range(nv + 1):
        for i in range(nu):
            L.append("vt %.4f %.4f" % (i / nu, j / nv))
    for v in verts:
        L.append("vn %.6f %.6f %.6f" % v)
    idx = lambda j, i: j * nu + (i % nu) + 1
    for j in range(nv):
        if j == nv // 2:
            L.append("o lower")
        for i in range(nu):
            a, b, c, d = idx(j, i), idx(j + 1, i), idx(j + 1, i + 1), idx(j, i + 1)
            f = lambda k: "%d/%d/%d" % (k, k, k)
            if j == 0:
                L.append("f %s %s %s" % (f(a), f(b), f(c)))
            elif j == nv - 1:
                L.append("f %s %s %s" % (f(a), f(b), f(d)))
            else:
                L.append("f %s %s %s %s" % (f(a), f(b), f(c), f(d)))
    n = len(verts)
    L.append("f %d//%d %d//%d %d//%d" % (-n, -n, -n + 1, -n + 1, -n + 2, -n + 2))   # relative indices, v//vn form
    return "\n".join(L) + "\n"


PLANE = "o ground\nv -2.5 -2.5 0\nv 2.5 -2.5 0\nv 2.5 2.5 0\nv -2.5 2.5 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\n"

SCENE = {
    "settings": {"max_bounces": 3, "spp": 2, "diffuse_bounce": False},
    "miss_color": [0.5, 0.7, 1.0],
    "camera": {"focal_length_mm": 28.0, "sensor_height_mm": 24.0, "pixel_width": 96, "pixel_height": 64,
               "position": [0.0, -2.5, 1.2], "look_at": [0.0, 0.0, 0.4], "up": [0.0, 0.0, 1.0]},
    "lights": [{"position": [-2.0, -1.0, 1.5], "color": [1.0, 1.0, 1.0], "intensity": 5.0},
               {"position": [1.5, -1.5, 2.0], "color": [0.2, 0.3, 1.0], "intensity": 2.0}],
    "scene": [
        {"name": "ball", "type": "mesh", "path": "./ball.obj",
         "transform": {"position": [0.5, 0.0, 0.5], "rotation": [10.0, 25.0, 40.0], "scale": [0.5, 0.45, 0.5]},
         "material": {"albedo": [0.8, 0.2, 0.2], "kd": 1, "ks": 0.5, "specular_color": [0.04, 0.04, 0.04], "shininess": 64, "kr": 0}},
        {"name": "mirror", "type": "mesh", "path": "./ball.obj",
         "transform": {"position": [-0.6, 0.2, 0.35], "rotation": [0.0, 0.0, 0.0], "scale": [0.35, 0.35, 0.35]},
         "material": {"albedo": [1, 1, 1], "kd": 0, "ks": 1, "specular_color": [0.9, 0.9, 0.9], "shininess": 500, "kr": 0.9}},
        {"name": "ground", "type": "mesh", "path": "./ground.obj",
         "transform": {"position": [0.0, 0.0, 0.0], "rotation": [0.0, 0.0, 0.0], "scale": [1.0, 1.0, 1.0]},
         "material": {"albedo": [0.6, 0.55, 0.5], "kd": 1, "ks": 0, "shininess": 1, "kr": 0.1, "specular_color": [0.5, 0.5, 0.5]}},
    ],
}


def main():
    from PIL import Image
    libs = orclib.ref_libs()
    hm, h2 = libs["ref_hw2_main"], libs["ref_hw2"]
    h2.ref_hw2_load_obj.restype = C.c_void_p
    ball, out = sphere_obj(), {}
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "ball.obj"), "w").write(ball)
        open(os.path.join(td, "ground.obj"), "w").write(PLANE)
        cwd = os.getcwd()
        os.chdir(td)
        try:
            for name, diffuse in (("mirror", False), ("diffuse", True)):
                sc = json.loads(json.dumps(SCENE))
                sc["settings"]["diffuse_bounce"] = diffuse
                txt = json.dumps(sc, indent=1)
                open("scene.json", "w").write(txt)
                argv = (C.c_char_p * 2)(b"bvh_viz", b"scene.json")
                rc = hm.ref_hw2_main(2, argv)
                assert rc == 0, rc
                img = np.array(Image.open("render.png").convert("RGB"))
                out["image_" + name], out["json_" + name] = img, np.array(txt)
                print(name, img.shape, "mean", img.mean((0, 1)))
            # meshes as the reference loads and transforms them (LoadOBJ_ToMesh + applyObjectTransform)
            nid = C.c_int(0)
            for k, o in enumerate(SCENE["scene"]):
                first = nid.value
                w = h2.ref_hw2_load_obj(o["path"][2:].encode(), C.byref(nid))
                nv, nn, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
                h2.ref_hw2_mesh_counts(C.c_void_p(w), C.byref(nv), C.byref(nn), C.byref(nt))
                pos = np.zeros((nv.value, 3), np.float32); nrm = np.zeros((nn.value, 3), np.float32)
                idx = np.zeros((nt.value, 3), np.uint32); obj = np.zeros(nt.value, np.int32)
                h2.ref_hw2_mesh_copy(C.c_void_p(w), pos.ctypes.data_as(A.f32p), nrm.ctypes.data_as(A.f32p), idx.ctypes.data_as(A.u32p), obj.ctypes.data_as(A.i32p))
                h2.ref_hw2_free(C.c_void_p(w))
                out["obj%d_raw_positions" % k] = pos.copy()
                t = o["transform"]
                f3 = lambda v: np.array(v, np.float32)
                hm.ref_hw2_transform(pos.ctypes.data_as(A.f32p), nrm.ctypes.data_as(A.f32p) if nn.value else None, C.c_uint64(nv.value),
                                     f3(t["position"]).ctypes.data_as(A.f32p), f3(t["rotation"]).ctypes.data_as(A.f32p), f3(t["scale"]).ctypes.data_as(A.f32p))
                out["obj%d_positions" % k], out["obj%d_normals" % k], out["obj%d_indices" % k], out["obj%d_ids" % k] = pos, nrm, idx, obj
                out["obj%d_first_next" % k] = np.array([first, nid.value])
                print(o["name"], pos.shape, idx.shape, "ids", first, "->", nid.value)
        finally:
            os.chdir(cwd)
    out["ball_obj"], out["ground_obj"] = np.array(ball), np.array(PLANE)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "e2e_scene.npz"), **out)


if __name__ == "__main__":
    main()
