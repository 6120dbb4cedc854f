"""Generates the committed fixtures under tests/golden/ from the reference itself.

Runs in the authoring container only (needs /root/reference and the shims built by oracle/build.py).
Everything written here is DATA produced by running the unmodified reference (or read from its
committed golden images); no reference source is copied.

  frog_mesh.npz / sphere_mesh.npz / cornell_mesh.npz   meshes as parsed by the reference loaders
  hw1_frog_output.npz      HW1/frog_output.png (the reference's committed 320x180 golden), as uint8
  ref_hw1_frog_96x54.npz   reference HW1 loop (ray_intersection + shade): rgb f32, rgb8, tri_id, t
  ref_hw2_frog_160x90.npz  reference render() + SearchBVH on frog.json's camera (depth 1): rgb, id, t
  ref_hw2_frogfill_160x90.npz  frame-filling frog view
  ref_hw2_terrain_128x72.npz   terrain(60,30) through the reference BVH path (primary + shadow)
  ref_hw2_cornell_96x96.npz    cornellbox.obj (no normals, 9 objects, axis-aligned walls) multi-material
  ref_vectors.npz          cameras, jitter tables, ray/triangle unit vectors, LBVH of a small mesh
  ppm_gradient.npz         ppm_p6_lib example images (8/16-bit) md5 + header bytes
"""
import ctypes as C
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orclib  # noqa: E402
from raytracinginonesemester_b200 import _abi as A, scenes  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
R = orclib.ref_libs()
h1, h2, hp = R["ref_hw1"], R["ref_hw2"], R["ref_ppm"]
h1.ref_hw1_load_obj.restype = C.c_void_p
h2.ref_hw2_load_obj.restype = C.c_void_p
h2.ref_hw2_world.restype = C.c_void_p
h2.ref_hw2_build.restype = C.c_double
h1.ref_hw1_render.restype = C.c_uint64


def fp(a):
    return a.ctypes.data_as(A.f32p)


def load_hw2(path):
    nid = C.c_int(0)
    w = h2.ref_hw2_load_obj(path.encode(), C.byref(nid))
    assert w, path
    nv, nn, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
    h2.ref_hw2_mesh_counts(C.c_void_p(w), C.byref(nv), C.byref(nn), C.byref(nt))
    pos = np.zeros((nv.value, 3), np.float32); nrm = np.zeros((nn.value, 3), np.float32)
    idx = np.zeros((nt.value, 3), np.uint32); obj = np.zeros(nt.value, np.int32)
    h2.ref_hw2_mesh_copy(C.c_void_p(w), fp(pos), fp(nrm), idx.ctypes.data_as(A.u32p), obj.ctypes.data_as(A.i32p))
    h2.ref_hw2_free(C.c_void_p(w))
    return pos, nrm, idx, obj, nid.value + 1


def load_hw1(path):
    m = h1.ref_hw1_load_obj(path.encode())
    assert m, path
    nv, nn, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
    h1.ref_hw1_mesh_counts(C.c_void_p(m), C.byref(nv), C.byref(nn), C.byref(nt))
    pos = np.zeros((nv.value, 3), np.float32); nrm = np.zeros((nn.value, 3), np.float32); idx = np.zeros((nt.value, 3), np.uint32)
    h1.ref_hw1_mesh_copy(C.c_void_p(m), fp(pos), fp(nrm), idx.ctypes.data_as(A.u32p))
    h1.ref_hw1_mesh_free(C.c_void_p(m))
    return pos, nrm, idx


def hw2_render(pos, nrm, idx, obj, mats, cam, W, H, miss, lights, spp=1, max_depth=1):
    """cam = (pos, look, up, focal, sensor).  Returns rgb (reference render()), tri_id, t (SearchBVH), nodes, aabbs."""
    w = h2.ref_hw2_world(fp(pos), fp(nrm) if nrm is not None and nrm.size else None, C.c_uint64(pos.shape[0]),
                         idx.ctypes.data_as(A.u32p), C.c_uint64(idx.shape[0]), obj.ctypes.data_as(A.i32p) if obj is not None else None)
    h2.ref_hw2_build(C.c_void_p(w))
    P = idx.shape[0]
    nodes = np.zeros((2 * P - 1, 4), np.uint32); aabbs = np.zeros((2 * P - 1, 6), np.float32)
    h2.ref_hw2_bvh_export(C.c_void_p(w), nodes.ctypes.data_as(A.u32p), fp(aabbs))
    cp, lk, up = (np.array(v, np.float32) for v in cam[:3])
    ms = np.array(miss, np.float32)
    marr = (A.rt_material * len(mats))(*mats)
    larr = (A.rt_light * len(lights))(*lights)
    rgb = np.zeros((H, W, 3), np.float32); rgb2 = np.zeros((H, W, 3), np.float32)
    tid = np.zeros((H, W), np.int32); tt = np.zeros((H, W), np.float32)
    h2.ref_hw2_render(C.c_void_p(w), fp(cp), fp(lk), fp(up), C.c_double(cam[3]), C.c_double(cam[4]), W, H, fp(ms), max_depth, spp,
                      marr, len(mats), larr, len(lights), 1, fp(rgb))
    h2.ref_hw2_render_rows(C.c_void_p(w), fp(cp), fp(lk), fp(up), C.c_double(cam[3]), C.c_double(cam[4]), W, H, fp(ms), max_depth, spp,
                           marr, len(mats), larr, len(lights), 1, 0, 1, 8, fp(rgb2), tid.ctypes.data_as(A.i32p), fp(tt))
    assert np.array_equal(rgb, rgb2), "row driver must equal the reference's own render()"
    h2.ref_hw2_free(C.c_void_p(w))
    return rgb, tid, tt, nodes, aabbs


def main():
    from PIL import Image
    os.makedirs(OUT, exist_ok=True)
    G = os.path.join(REF, "HW2", "HW2", "GPUandCPU")
    # ---- meshes ----
    pos, nrm, idx = load_hw1(os.path.join(REF, "HW1/assets/meshes/frog.obj"))
    pos2, nrm2, idx2, obj2, nobj = load_hw2(os.path.join(G, "assets/meshes/frog.obj"))
    assert np.array_equal(pos, pos2) and np.array_equal(nrm, nrm2) and np.array_equal(idx, idx2)
    np.savez_compressed(os.path.join(OUT, "frog_mesh.npz"), positions=pos, normals=nrm, indices=idx, tri_obj_ids=obj2)
    spos, snrm, sidx = load_hw1(os.path.join(REF, "HW1/assets/meshes/sphere.obj"))
    np.savez_compressed(os.path.join(OUT, "sphere_mesh.npz"), positions=spos, normals=snrm, indices=sidx)
    cpos, cnrm, cidx, cobj, cn = load_hw2(os.path.join(G, "assets/meshes/cornellbox.obj"))
    np.savez_compressed(os.path.join(OUT, "cornell_mesh.npz"), positions=cpos, normals=cnrm, indices=cidx, tri_obj_ids=cobj)
    print("frog", pos.shape, idx.shape, "sphere", sidx.shape, "cornell", cidx.shape, "objects", cn, "normals", cnrm.shape)

    # ---- committed golden image of the reference ----
    img = np.array(Image.open(os.path.join(REF, "HW1/frog_output.png")).convert("RGB"))
    np.savez_compressed(os.path.join(OUT, "hw1_frog_output.npz"), rgb8=img)

    # ---- HW1 loop through the reference functions ----
    W, H = 96, 54
    cam = (np.array([0, -1, 1], np.float32), np.array([0, 0.15, 0], np.float32), np.array([0, 0, 1], np.float32))
    lp = np.array([-3, 0, 1], np.float32)
    for name, lc in (("white", np.array([1, 1, 1], np.float32)), ("magenta", np.array([1, 0, 1], np.float32))):
        rgb = np.zeros((H, W, 3), np.float32); rgb8 = np.zeros((H, W, 3), np.uint8)
        tid = np.zeros((H, W), np.int32); tt = np.zeros((H, W), np.float32)
        h1.ref_hw1_render(fp(pos), fp(nrm), idx.ctypes.data_as(A.u32p), C.c_uint64(idx.shape[0]), fp(cam[0]), fp(cam[1]), fp(cam[2]),
                          C.c_double(255.0), C.c_double(24.0), W, H, fp(lp), fp(lc), 1, 42, 0, 1, 8,
                          fp(rgb), rgb8.ctypes.data_as(A.u8p), tid.ctypes.data_as(A.i32p), fp(tt))
        np.savez_compressed(os.path.join(OUT, "ref_hw1_frog_96x54_%s.npz" % name), rgb=rgb, rgb8=rgb8, tri_id=tid, t=tt)
    # C1 "as the repo runs it today": sphere.obj with the frog camera at 64x36
    rgb8 = np.zeros((36, 64, 3), np.uint8); tid = np.zeros((36, 64), np.int32)
    lc = np.array([1, 0, 1], np.float32)
    h1.ref_hw1_render(fp(spos), fp(snrm), sidx.ctypes.data_as(A.u32p), C.c_uint64(sidx.shape[0]), fp(cam[0]), fp(cam[1]), fp(cam[2]),
                      C.c_double(255.0), C.c_double(24.0), 64, 36, fp(lp), fp(lc), 1, 42, 0, 1, 8,
                      None, rgb8.ctypes.data_as(A.u8p), tid.ctypes.data_as(A.i32p), None)
    np.savez_compressed(os.path.join(OUT, "ref_hw1_sphere_64x36.npz"), rgb8=rgb8, tri_id=tid)

    # ---- HW2 BVH path ----
    from raytracinginonesemester_b200.api import make_light, make_material
    frog_mat = [make_material(**scenes.FROG_MATERIAL)]
    frog_light = [make_light((-3.0, 0.0, 1.0), (1.0, 1.0, 0.0), 5)]
    for name, c in (("frog", ((0.0, -0.2, 0.2), (0.0, 0.1, 0.0), (0, 0, 1), 45.0, 24.0)),
                    ("frogfill", ((0.0, -0.2, 0.2), (0.0, 0.095, 0.03), (0, 0, 1), 170.0, 24.0))):
        rgb, tid, tt, nodes, aabbs = hw2_render(pos, nrm, idx, obj2, frog_mat, c, 160, 90, (0, 0, 0), frog_light)
        np.savez_compressed(os.path.join(OUT, "ref_hw2_%s_160x90.npz" % name), rgb=rgb, tri_id=tid, t=tt)
        print(name, "hit px", (tid >= 0).sum())
    tp, ti = scenes.terrain(60, 30, 42)
    tobj = np.zeros(ti.shape[0], np.int32)
    tmat = [make_material(**scenes.TERRAIN_MATERIAL)]
    tl = [make_light((-2.0, -1.0, 1.5), (1, 1, 1), 5)]
    rgb, tid, tt, nodes, aabbs = hw2_render(tp, None, ti, tobj, tmat, ((0, 0, 1), (0, 0, 0), (0, 1, 0), 24.0, 24.0), 128, 72, (0.5, 0.7, 1.0), tl)
    np.savez_compressed(os.path.join(OUT, "ref_hw2_terrain_128x72.npz"), rgb=rgb, tri_id=tid, t=tt, nodes=nodes, aabbs=aabbs)
    # multi-sample (4 spp, reference jitter table) on the same terrain
    rgb4, tid4, tt4, _, _ = hw2_render(tp, None, ti, tobj, tmat, ((0, 0, 1), (0, 0, 0), (0, 1, 0), 24.0, 24.0), 64, 36, (0.5, 0.7, 1.0), tl, spp=4)
    np.savez_compressed(os.path.join(OUT, "ref_hw2_terrain_64x36_spp4.npz"), rgb=rgb4, tri_id=tid4, t=tt4)
    # cornell box: 9 objects, 3 materials cycling, two lights, zero normals
    mats = [make_material(albedo=(0.7, 0.7, 0.7)), make_material(albedo=(0.8, 0.1, 0.1), ks=0.4, shininess=16.0),
            make_material(albedo=(0.1, 0.8, 0.1), kd=0.5, ks=0.5, specular_color=(0.9, 0.9, 0.9), shininess=64.0, emission=(0.05, 0.0, 0.0))]
    cmats = [mats[i % 3] for i in range(cn)]
    lo, hi = cpos.min(0), cpos.max(0)
    ctr = (lo + hi) / 2
    ext = (hi - lo).max()
    ccam = (tuple(ctr + np.array([0.0, -1.6 * ext, 0.1 * ext])), tuple(ctr), (0, 0, 1), 35.0, 24.0)
    cl = [make_light(tuple(ctr + np.array([0.2 * ext, -0.3 * ext, 0.4 * ext])), (1, 1, 1), 2), make_light(tuple(ctr + np.array([-0.3 * ext, -1.0 * ext, 0.3 * ext])), (0.4, 0.4, 1.0), 1)]
    rgb, tid, tt, _, _ = hw2_render(cpos, cnrm if cnrm.size else None, cidx, cobj, cmats, ccam, 96, 96, (0.1, 0.2, 0.3), cl)
    np.savez_compressed(os.path.join(OUT, "ref_hw2_cornell_96x96.npz"), rgb=rgb, tri_id=tid, t=tt,
                        cam=np.array(list(ccam[0]) + list(ccam[1]), np.float64), lights=np.array([list(l.position) for l in cl], np.float32))
    print("cornell hit px", (tid >= 0).sum(), "of", tid.size)

    # ---- small vectors ----
    cams = []
    for args in (((0, -1, 1), (0, 0.15, 0), (0, 0, 1), 255.0, 24.0, 320, 180), ((0, 0, 1), (0, 0, 0), (0, 1, 0), 24.0, 24.0, 3840, 2160),
                 ((0.0, -0.2, 0.2), (0.0, 0.1, 0.0), (0, 0, 1), 45.0, 24.0, 1920, 1080), ((1, 2, 3), (1, 2, 3), (0, 0, 1), 50.0, 24.0, 7, 5),
                 ((0, 0, 0), (0, 0, 5), (0, 0, 1), 35.0, 36.0, 1, 1)):
        o1 = np.zeros(12, np.float32); o2 = np.zeros(12, np.float32)
        a = [np.array(v, np.float32) for v in args[:3]]
        h1.ref_hw1_camera(fp(a[0]), fp(a[1]), fp(a[2]), C.c_double(args[3]), C.c_double(args[4]), args[5], args[6], fp(o1))
        h2.ref_hw2_camera(fp(a[0]), fp(a[1]), fp(a[2]), C.c_double(args[3]), C.c_double(args[4]), args[5], args[6], fp(o2))
        assert np.array_equal(o1, o2)
        cams.append(np.concatenate([np.array(list(args[0]) + list(args[1]) + list(args[2]) + [args[3], args[4], args[5], args[6]], np.float64), o1.astype(np.float64)]))
    j16 = np.zeros((16, 2), np.float32); h2.ref_hw2_jitter(16, 42, fp(j16))
    j1h = np.zeros((4, 2), np.float32); h1.ref_hw1_jitter(4, 42, fp(j1h))
    j7 = np.zeros((700, 2), np.float32); h2.ref_hw2_jitter(700, 12345, fp(j7))
    # ray/triangle vectors of HW1/test_ray_tri_inter_STANDALONE (8 directed + 57-point sweep) + random probes
    v0, v1, v2 = (np.array(v, np.float32) for v in ((-5, -5, -10), (0, 5, -10), (5, -5, -10)))
    origin = np.zeros(3, np.float32)
    dirs = [(-5, -5, -10), (0, 0, -10), (6, 0, -10), (0, -5, -10), (1, 0, 0), (0, 0, 10), (-2.5 + 0.001, 0, -10), (-2.5 - 0.001, 0, -10)]
    for ai in range(11):
        for bi in range(11 - ai):
            al, be = 0.1 * ai, 0.1 * bi
            ga = 1.0 - al - be
            if ga < -1e-9:
                continue
            p = al * v0.astype(np.float64) + be * v1.astype(np.float64) + ga * v2.astype(np.float64)
            dirs.append(tuple(np.float32(p)))
    rng = np.random.default_rng(7)
    for _ in range(400):
        dirs.append(tuple(np.float32(rng.uniform(-7, 7, 2)).tolist() + [np.float32(-10.0)]))
    dirs = np.array(dirs, np.float32)
    rt = np.zeros((len(dirs), 4), np.float32)
    for k, d in enumerate(dirs):
        t1 = C.c_float(); t2 = C.c_float()
        hit1 = h1.ref_hw1_ray_triangle(fp(origin), fp(d), fp(v0), fp(v1), fp(v2), C.byref(t1))
        du = (d / np.float32(np.sqrt(np.float32(np.float32(d[0] * d[0]) + np.float32(d[1] * d[1])) + np.float32(d[2] * d[2])))).astype(np.float32)
        hit2 = h2.ref_hw2_ray_triangle(fp(origin), fp(du), fp(v0), fp(v1), fp(v2), C.byref(t2))
        rt[k] = (hit1, t1.value, hit2, t2.value)
    print("directed:", rt[:8, 0], "sweep hits:", int(rt[8:8 + 66, 0].sum()), "of", 66)
    np.savez_compressed(os.path.join(OUT, "ref_vectors.npz"), cameras=np.array(cams), jitter16_seed42=j16, jitter_hw1_4_seed42=j1h,
                        jitter700_seed12345=j7, tri=np.stack([v0, v1, v2]), ray_dirs=dirs, ray_results=rt)

    # ---- ppm_p6_lib example (gradient, 8 and 16 bit) ----
    Wg = Hg = 256
    xs = np.arange(Wg, dtype=np.float64) / (Wg - 1)
    ys = np.arange(Hg, dtype=np.float64) / (Hg - 1)
    g = np.zeros((Hg, Wg, 3), np.float32)
    g[..., 0] = xs[None, :]; g[..., 1] = ys[:, None]; g[..., 2] = 0.25
    res = {}
    for name, maxval, gamma in (("g8", 255, 0), ("g8gamma", 255, 1), ("g16", 65535, 0)):
        path = "/tmp/_golden_%s.ppm" % name
        rc = hp.ref_ppm_write_rgbf(path.encode(), Wg, Hg, fp(g), maxval, 1, gamma, 0, None, 0)
        assert rc == 0
        data = open(path, "rb").read()
        res[name + "_md5"] = hashlib.md5(data).hexdigest()
        res[name + "_size"] = len(data)
        res[name + "_head"] = np.frombuffer(data[:64], np.uint8)
        res[name + "_tail"] = np.frombuffer(data[-64:], np.uint8)
    np.savez_compressed(os.path.join(OUT, "ppm_gradient.npz"), **res)
    print({k: v for k, v in res.items() if "md5" in k or "size" in k})


if __name__ == "__main__":
    main()
