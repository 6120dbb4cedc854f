"""Pins the oracle (oracle/rt_oracle.c) against the reference: the committed golden image
HW1/frog_output.png, fixtures produced by running the unmodified reference in place
(tools/make_golden.py -> tests/golden/), the reference's own unit vectors, and — when the in-place
shims exist (oracle/_ref/, authoring container) — the reference itself, live."""
import ctypes as C
import hashlib

import os

import numpy as np
import pytest

import orclib
from raytracinginonesemester_b200 import _abi as A, api, scenes

ALL = A.RT_OUT_RGB_F32 | A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T


def test_hw1_frog_golden_png_bit_exact(frog_scene, golden):
    """HW1/frog_output.png (320x180, white light): 0 of 57 600 pixels differ."""
    fr = scenes.hw1_frame(320, 180, light_color=(1, 1, 1), outputs=ALL)
    o = orclib.oracle_render(frog_scene, fr)
    ref = golden("hw1_frog_output.npz")["rgb8"]
    assert np.array_equal(o["rgb8"], ref)
    assert (o["tri_id"] >= 0).sum() == 3795


@pytest.mark.parametrize("name,col", [("white", (1, 1, 1)), ("magenta", (1, 0, 1))])
def test_hw1_loop_matches_reference_functions(frog_scene, golden, name, col):
    ref = golden("ref_hw1_frog_96x54_%s.npz" % name)
    o = orclib.oracle_render(frog_scene, scenes.hw1_frame(96, 54, light_color=col, outputs=ALL))
    for k in ("tri_id", "t", "rgb", "rgb8"):
        assert np.array_equal(o[k], ref[k]), k


def test_hw1_sphere_as_the_repo_runs_it(golden):
    d = golden("sphere_mesh.npz")
    sc = api.Scene(d["positions"], d["indices"], normals=d["normals"])
    ref = golden("ref_hw1_sphere_64x36.npz")
    o = orclib.oracle_render(sc, scenes.hw1_frame(64, 36, outputs=ALL))
    assert np.array_equal(o["rgb8"], ref["rgb8"]) and np.array_equal(o["tri_id"], ref["tri_id"])
    assert np.all(o["rgb8"] == np.array([20, 5, 5], np.uint8))


@pytest.mark.parametrize("name,filling", [("frog", False), ("frogfill", True)])
def test_hw2_bvh_path_matches_reference(frog_scene, golden, name, filling):
    """Reference LBVH + SearchBVH + ShadeDirect restated: ids, t and float rgb bit-exact."""
    ref = golden("ref_hw2_%s_160x90.npz" % name)
    bvh = orclib.oracle_bvh(frog_scene)
    o = orclib.oracle_render(frog_scene, scenes.frog_frame(160, 90, filling=filling, outputs=ALL), bvh=bvh)
    for k in ("tri_id", "t", "rgb"):
        assert np.array_equal(o[k], ref[k]), k
    orclib.oracle().orc_bvh_free(bvh)


def test_hw2_terrain_lbvh_and_render_match_reference(golden):
    sc = scenes.terrain_scene(60, 30)
    ref = golden("ref_hw2_terrain_128x72.npz")
    bvh = orclib.oracle_bvh(sc)
    P = sc.indices.shape[0]
    nodes = np.zeros((2 * P - 1, 4), np.uint32); aabbs = np.zeros((2 * P - 1, 6), np.float32)
    orclib.oracle().orc_bvh_export(bvh, nodes.ctypes.data_as(A.u32p), aabbs.ctypes.data_as(A.f32p))
    # topology (left/right/object) and boxes of the reference's buildBVH, node for node
    assert np.array_equal(nodes[:, 1:], ref["nodes"][:, 1:])
    assert np.array_equal(nodes[1:, 0], ref["nodes"][1:, 0])          # parent (root's parent is never written)
    assert np.array_equal(aabbs, ref["aabbs"])
    o = orclib.oracle_render(sc, scenes.terrain_frame(128, 72, outputs=ALL), bvh=bvh)
    for k in ("tri_id", "t", "rgb"):
        assert np.array_equal(o[k], ref[k]), k
    ref4 = golden("ref_hw2_terrain_64x36_spp4.npz")
    o4 = orclib.oracle_render(sc, scenes.terrain_frame(64, 36, spp=4, outputs=ALL), bvh=bvh)
    for k in ("tri_id", "t", "rgb"):
        assert np.array_equal(o4[k], ref4[k]), k


def test_hw2_cornell_multi_material(golden):
    d = golden("cornell_mesh.npz")
    mats = [api.make_material(albedo=(0.7, 0.7, 0.7)), api.make_material(albedo=(0.8, 0.1, 0.1), ks=0.4, shininess=16.0),
            api.make_material(albedo=(0.1, 0.8, 0.1), kd=0.5, ks=0.5, specular_color=(0.9, 0.9, 0.9), shininess=64.0, emission=(0.05, 0.0, 0.0))]
    nobj = int(d["tri_obj_ids"].max()) + 2
    sc = api.Scene(d["positions"], d["indices"], normals=None, tri_obj_ids=d["tri_obj_ids"], materials=[mats[i % 3] for i in range(nobj)])
    ref = golden("ref_hw2_cornell_96x96.npz")
    cam = api.camera_init(ref["cam"][:3], ref["cam"][3:6], (0, 0, 1), 35.0, 24.0, 96, 96)
    lp = ref["lights"]
    fr = api.Frame(cam, 96, 96, lights=[api.make_light(lp[0], (1, 1, 1), 2), api.make_light(lp[1], (0.4, 0.4, 1.0), 1)],
                   miss_color=(0.1, 0.2, 0.3), jitter=api.jitter_table(1, 42, True), outputs=ALL)
    bvh = orclib.oracle_bvh(sc)
    o = orclib.oracle_render(sc, fr, bvh=bvh)
    for k in ("tri_id", "t", "rgb"):
        assert np.array_equal(o[k], ref[k]), k


def _bounce_cases(golden):
    g = golden("ref_hw2_bounce_cornell.npz")
    sc, cam, lights, miss = scenes.cornell_bounce_scene(os.path.join(os.path.dirname(__file__), "golden", "cornell_mesh.npz"))
    for name, W, H, spp, depth, diffuse in g["cases"]:
        fr = scenes.cornell_bounce_frame(cam, lights, miss, int(W), int(H), int(spp), int(depth), int(diffuse), outputs=ALL)
        yield str(name), sc, fr, {k: g["%s_%s" % (name, k)] for k in ("rgb", "tri_id", "t")}


def test_hw2_bounce_loop_matches_reference(golden):
    """TraceRayIterative with max_depth > 1 (mirror and hash-RNG diffuse bounces, query.h:32-70,193-216): the oracle
    over the reference LBVH, and the canonical brute force, reproduce the reference's frames bit for bit."""
    for name, sc, fr, ref in _bounce_cases(golden):
        bvh = orclib.oracle_bvh(sc)
        o = orclib.oracle_render(sc, fr, bvh=bvh)
        for k in ("tri_id", "t", "rgb"):
            assert np.array_equal(o[k], ref[k]), (name, k)
        assert o["counters"]["rays_primary"] > fr.width * fr.height * fr.spp      # bounce rays were traced
        b = orclib.oracle_render(sc, fr, bvh=None)
        for k in ("tri_id", "t", "rgb"):
            assert np.array_equal(b[k], ref[k]), (name, "canonical", k)


@pytest.mark.parametrize("name", ["sphere_point", "sphere", "cornell"])
def test_cpuonly_mode_matches_reference(golden, name):
    """RT_MODE_HW2_CPU: the oracle's restatement of the CPUOnly renderer (ray_intersection, IntersectScene,
    ShadeDirect/ShadowVisibility, EvaluateBRDF, mirror TraceRay, sky) against frames rendered by the reference itself."""
    g = golden("cpuonly_scenes.npz")
    sc, fr = scenes.cpuonly_case(g, name, outputs=ALL)
    o = orclib.oracle_render(sc, fr)
    for k in ("tri_id", "t", "rgb"):
        assert np.array_equal(o[k], g["%s_%s" % (name, k)]), (name, k)
    if name != "sphere_point":
        assert o["counters"]["rays_primary"] > fr.width * fr.height          # mirror bounces were traced


def test_cpuonly_committed_golden_png_bit_exact(golden):
    """CPUOnly/output/sphere_point_output.png (360x240), the one HW2 golden the reference still reproduces bit for bit."""
    g = golden("cpuonly_scenes.npz")
    png = g["sphere_point_golden_png"]
    sc, fr = scenes.cpuonly_case(g, "sphere_point", width=png.shape[1], height=png.shape[0], outputs=A.RT_OUT_RGB8)
    o = orclib.oracle_render(sc, fr, want=("rgb8",))
    assert np.array_equal(o["rgb8"], png)


def test_canonical_brute_force_equals_reference_bvh(frog_scene, golden):
    """SURVEY §8c last row: 'min t, then min id over intersectTriangle on all triangles' reproduces the
    reference BVH result (differences only at exact-t ties, far below the 99.99 % bar)."""
    ref = golden("ref_hw2_frogfill_160x90.npz")
    fr = scenes.frog_frame(160, 90, filling=True, outputs=ALL, accel=A.RT_ACCEL_BRUTE)
    o = orclib.oracle_render(frog_scene, fr)
    mism = o["tri_id"] != ref["tri_id"]
    assert mism.sum() <= 1e-4 * mism.size
    assert np.array_equal(o["t"][~mism], ref["t"][~mism])
    assert np.array_equal(o["t"][mism], ref["t"][mism])               # mismatches are exact-t ties


def test_reference_vectors(golden):
    v = golden("ref_vectors.npz")
    lib = orclib.oracle()
    for row in v["cameras"]:
        cam = A.rt_camera()
        p, l, u = (np.array(row[i:i + 3], np.float32) for i in (0, 3, 6))
        rc = lib.orc_camera_init(C.byref(cam), p.ctypes.data_as(A.f32p), l.ctypes.data_as(A.f32p), u.ctypes.data_as(A.f32p),
                                 row[9], row[10], int(row[11]), int(row[12]))
        assert rc == 0
        got = np.array(list(cam.center) + list(cam.pixel00_loc) + list(cam.pixel_delta_u) + list(cam.pixel_delta_v), np.float32)
        assert np.array_equal(got, row[13:].astype(np.float32))
    cam = A.rt_camera()
    z = np.zeros(3, np.float32)
    assert lib.orc_camera_init(C.byref(cam), z.ctypes.data_as(A.f32p), z.ctypes.data_as(A.f32p), z.ctypes.data_as(A.f32p), 50.0, 24.0, 0, 10) != 0
    for key, n, seed, centered in (("jitter16_seed42", 16, 42, 1), ("jitter_hw1_4_seed42", 4, 42, 0), ("jitter700_seed12345", 700, 12345, 1)):
        out = np.zeros((n, 2), np.float32)
        lib.orc_jitter_table(out.ctypes.data_as(A.f32p), n, seed, centered)
        assert np.array_equal(out, v[key]), key
    assert abs(float(v["jitter16_seed42"][0, 0]) - (-0.12545988)) < 1e-8 and abs(float(v["jitter16_seed42"][0, 1]) - 0.296543002) < 1e-8
    # ray/triangle unit vectors of HW1/test_ray_tri_inter_STANDALONE + random probes, both contracts
    tri, dirs, res = v["tri"], v["ray_dirs"], v["ray_results"]
    o = np.zeros(3, np.float32)
    for d, r in zip(dirs, res):
        t = C.c_float()
        hit = lib.orc_ray_triangle(0, o.ctypes.data_as(A.f32p), d.ctypes.data_as(A.f32p), 1, *(x.ctypes.data_as(A.f32p) for x in tri), C.byref(t))
        assert hit == int(r[0]) and (not hit or t.value == r[1])
        hit = lib.orc_ray_triangle(1, o.ctypes.data_as(A.f32p), d.ctypes.data_as(A.f32p), 1, *(x.ctypes.data_as(A.f32p) for x in tri), C.byref(t))
        assert hit == int(r[2]) and (not hit or t.value == r[3])
    assert list(res[:8, 0]) == [1, 1, 0, 1, 0, 0, 1, 0]     # the 8 directed REQUIREs of the reference test
    assert int(res[8:74, 0].sum()) == 65                      # barycentric sweep: exactly one miss without FMA


def test_ppm_quantiser_matches_reference_writer(golden):
    """float_to_sample of ppm_p6.cpp:137-155: the oracle's quantiser rebuilds the bytes of the reference's
    gradient example (md5 bb750b71... as recorded in SURVEY §4)."""
    g = golden("ppm_gradient.npz")
    W = H = 256
    xs = np.arange(W, dtype=np.float64) / (W - 1)
    ys = np.arange(H, dtype=np.float64) / (H - 1)
    img = np.zeros((H, W, 3), np.float32)
    img[..., 0] = xs[None, :]; img[..., 1] = ys[:, None]; img[..., 2] = 0.25
    lib = orclib.oracle()
    for name, q in (("g8", A.RT_QUANT_PPM_LROUND), ("g8gamma", A.RT_QUANT_PPM_GAMMA2)):
        body = bytes(lib.orc_quantise(float(c), q) for c in img.ravel())
        data = b"P6\n256 256\n255\n" + body
        assert len(data) == int(g[name + "_size"])
        assert hashlib.md5(data).hexdigest() == str(g[name + "_md5"])
    assert str(g["g8_md5"]).startswith("bb750b71")


def test_live_reference_when_available():
    """Random small scene through the reference compiled in place vs the oracle (authoring container)."""
    libs = orclib.ref_libs()
    if "ref_hw2" not in libs:
        pytest.skip("oracle/_ref not built here")
    h2 = libs["ref_hw2"]
    h2.ref_hw2_world.restype = C.c_void_p
    rng = np.random.default_rng(3)
    pos = rng.uniform(-1, 1, (90, 3)).astype(np.float32)
    idx = rng.integers(0, 90, (200, 3)).astype(np.uint32)
    nrm = rng.normal(size=(90, 3)).astype(np.float32)
    obj = rng.integers(0, 3, 200).astype(np.int32)
    mats = [api.make_material(albedo=(0.9, 0.3, 0.2), ks=0.2), api.make_material(albedo=(0.2, 0.9, 0.2), kd=0.7, ks=0.6, shininess=8.0),
            api.make_material(emission=(0.1, 0.1, 0.2))]
    sc = api.Scene(pos, idx, normals=nrm, tri_obj_ids=obj, materials=mats)
    W, H = 72, 48
    cpos, look, up = np.array([0.2, -3, 0.4], np.float32), np.zeros(3, np.float32), np.array([0, 0, 1], np.float32)
    lights = [api.make_light((2, -2, 3), (1, 0.9, 0.8), 3), api.make_light((-2, -1, -2), (0.3, 0.3, 1), 2)]
    w = h2.ref_hw2_world(pos.ctypes.data_as(A.f32p), nrm.ctypes.data_as(A.f32p), C.c_uint64(90), idx.ctypes.data_as(A.u32p), C.c_uint64(200), obj.ctypes.data_as(A.i32p))
    h2.ref_hw2_build(C.c_void_p(w))
    rgb = np.zeros((H, W, 3), np.float32); tid = np.zeros((H, W), np.int32); tt = np.zeros((H, W), np.float32)
    ms = np.array([0.1, 0.1, 0.1], np.float32)
    marr = (A.rt_material * 3)(*mats); larr = (A.rt_light * 2)(*lights)
    h2.ref_hw2_render_rows(C.c_void_p(w), cpos.ctypes.data_as(A.f32p), look.ctypes.data_as(A.f32p), up.ctypes.data_as(A.f32p), C.c_double(30.0), C.c_double(24.0),
                           W, H, ms.ctypes.data_as(A.f32p), 1, 2, marr, 3, larr, 2, 1, 0, 1, 2, rgb.ctypes.data_as(A.f32p), tid.ctypes.data_as(A.i32p), tt.ctypes.data_as(A.f32p))
    h2.ref_hw2_free(C.c_void_p(w))
    cam = api.camera_init(cpos, look, up, 30.0, 24.0, W, H)
    fr = api.Frame(cam, W, H, lights=lights, miss_color=(0.1, 0.1, 0.1), spp=2, jitter=api.jitter_table(2, 42, True), outputs=ALL)
    bvh = orclib.oracle_bvh(sc)
    o = orclib.oracle_render(sc, fr, bvh=bvh)
    assert (tid >= 0).sum() > 100
    for k, a in (("tri_id", tid), ("t", tt), ("rgb", rgb)):
        assert np.array_equal(o[k], a), k
