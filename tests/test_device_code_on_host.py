"""The product's per-element device functions (csrc/rt_build_core.h, rt_trace_core.h), compiled for
the host by tests/emul/, against the oracle: BVH build + flattening + traversal + shading logic is
checked here without a GPU; the -m gpu tests then check the kernels that wrap the same functions."""
import ctypes as C

import numpy as np
import pytest

import orclib
from raytracinginonesemester_b200 import _abi as A, api, scenes

ALL = A.RT_OUT_RGB_F32 | A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T


def check(scene, frame, leaf_max=4, exact_rgb=True):
    h = orclib.emul_build(scene, leaf_max)
    assert h
    assert orclib.emul().emu_validate(h) == 0
    e = orclib.emul_render(h, frame)
    frame.accel = A.RT_ACCEL_BRUTE
    o = orclib.oracle_render(scene, frame)
    frame.accel = A.RT_ACCEL_BVH
    assert np.array_equal(e["tri_id"], o["tri_id"])
    assert np.array_equal(e["t"], o["t"])
    if exact_rgb:
        assert np.array_equal(e["rgb"], o["rgb"]) and np.array_equal(e["rgb8"], o["rgb8"])
    assert e["stats"]["max_stack"] < 32
    orclib.emul().emu_free(h)
    return e, o


@pytest.mark.parametrize("leaf_max", [1, 2, 4, 8])
def test_terrain_all_leaf_sizes(leaf_max):
    e, o = check(scenes.terrain_scene(48, 24), scenes.terrain_frame(128, 72, spp=2, outputs=ALL), leaf_max)
    assert e["stats"]["rays_shadow"] == o["counters"]["rays_shadow"] > 0


def test_frog_hw2_and_hw1(frog_scene):
    check(frog_scene, scenes.frog_frame(96, 54, filling=True, outputs=ALL))
    check(frog_scene, scenes.hw1_frame(96, 54, light_color=(1, 1, 1), outputs=ALL, accel=A.RT_ACCEL_BVH))


def test_matches_reference_fixture_bar(frog_scene, golden):
    """Emulated device code vs the reference BVH path: the north-star bar (99.99 % ids, t 1e-5)."""
    ref = golden("ref_hw2_frogfill_160x90.npz")
    h = orclib.emul_build(frog_scene, 4)
    e = orclib.emul_render(h, scenes.frog_frame(160, 90, filling=True, outputs=ALL))
    mism = e["tri_id"] != ref["tri_id"]
    assert mism.sum() <= 1e-4 * mism.size
    assert np.array_equal(e["t"][mism], ref["t"][mism])
    hit = (ref["tri_id"] >= 0) & ~mism
    assert np.array_equal(e["t"][hit], ref["t"][hit])
    assert np.array_equal(e["rgb"][~mism], ref["rgb"][~mism])


def test_tiny_degenerate_and_tied_scenes():
    pos = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0], [0, 0, 0.5], [0, 0, 0.5], [0, 0, 0.5]], np.float32)
    for idx in ([[0, 1, 2]], [[0, 1, 2], [0, 2, 1]], [[0, 1, 2], [3, 4, 5]], [[0, 1, 2]] * 7 + [[3, 4, 5]] * 3):
        sc = api.Scene(pos, np.array(idx, np.uint32))
        cam = api.camera_init((0.1, 0.05, 3), (0, 0, 0), (0, 1, 0), 50.0, 24.0, 40, 30)
        fr = api.Frame(cam, 40, 30, lights=[api.make_light((1, 1, 2), (1, 1, 1), 3)], miss_color=(0.2, 0.3, 0.4), outputs=ALL)
        for lm in (1, 4):
            e, o = check(sc, fr, lm)
            assert e["tri_id"].max() == 0 and (e["tri_id"] < 0).any()


def test_axis_parallel_rays_and_flat_boxes():
    g = np.linspace(-1, 1, 9, dtype=np.float32)
    xx, yy = np.meshgrid(g, g)
    pos = np.stack([xx.ravel(), yy.ravel(), np.zeros(81, np.float32)], 1)
    idx = []
    for j in range(8):
        for i in range(8):
            a = j * 9 + i
            idx += [[a, a + 1, a + 10], [a, a + 10, a + 9]]
    sc = api.Scene(pos, np.array(idx, np.uint32))
    cam = api.camera_init((0, 0, 2), (0, 0, 0), (0, 1, 0), 30.0, 24.0, 33, 33)
    fr = api.Frame(cam, 33, 33, lights=[api.make_light((0, 0, 5), (1, 1, 1), 2)], outputs=ALL)
    e, o = check(sc, fr, 2)
    assert e["tri_id"][16, 16] >= 0


def test_random_soup_with_materials_and_two_lights():
    rng = np.random.default_rng(11)
    pos = rng.uniform(-1, 1, (300, 3)).astype(np.float32)
    idx = rng.integers(0, 300, (700, 3)).astype(np.uint32)
    nrm = rng.normal(size=(300, 3)).astype(np.float32)
    obj = rng.integers(-1, 4, 700).astype(np.int32)          # includes out-of-range ids -> default material
    mats = [api.make_material(albedo=(0.9, 0.3, 0.2), ks=0.2), api.make_material(albedo=(0.2, 0.9, 0.2), kd=0.7, ks=0.6, shininess=8.0),
            api.make_material(emission=(0.1, 0.1, 0.2))]
    sc = api.Scene(pos, idx, normals=nrm, tri_obj_ids=obj, materials=mats)
    cam = api.camera_init((0.2, -3, 0.4), (0, 0, 0), (0, 0, 1), 30.0, 24.0, 80, 60)
    fr = api.Frame(cam, 80, 60, lights=[api.make_light((2, -2, 3), (1, 0.9, 0.8), 3), api.make_light((-2, -1, -2), (0.3, 0.3, 1), 2)],
                   miss_color=(0.1, 0.1, 0.1), spp=3, jitter=api.jitter_table(3, 42, True), outputs=ALL)
    check(sc, fr, 4)
    fr.shadows = False
    check(sc, fr, 4)
    fr.max_depth = 0                                           # TraceRayIterative returns black
    e, o = check(sc, fr, 4)
    assert not e["rgb"].any()


def test_bounce_loop_matches_reference_fixture(golden):
    """The product's per-ray sample function (rt_sample_bvh: closest hit, shading, shadow rays, mirror / diffuse
    bounces) compiled for the host, over the product's own BVH, against frames rendered by the reference."""
    import os
    g = golden("ref_hw2_bounce_cornell.npz")
    sc, cam, lights, miss = scenes.cornell_bounce_scene(os.path.join(os.path.dirname(__file__), "golden", "cornell_mesh.npz"))
    for leaf in (1, 4):
        h = orclib.emul_build(sc, leaf)
        for name, W, H, spp, depth, diffuse in g["cases"]:
            fr = scenes.cornell_bounce_frame(cam, lights, miss, int(W), int(H), int(spp), int(depth), int(diffuse), outputs=ALL)
            e = orclib.emul_render(h, fr)
            for k in ("tri_id", "t", "rgb"):
                assert np.array_equal(e[k], g["%s_%s" % (name, k)]), (name, leaf, k)
            assert e["stats"]["rays_primary"] > int(W) * int(H) * int(spp)


@pytest.mark.parametrize("name", ["sphere_point", "sphere", "cornell"])
def test_cpuonly_mode_matches_reference_fixture(golden, name):
    """RT_MODE_HW2_CPU through the product's per-ray sample function and BVH (host build of the device code)."""
    g = golden("cpuonly_scenes.npz")
    sc, fr = scenes.cpuonly_case(g, name, outputs=ALL)
    for leaf in (1, 4):
        e = orclib.emul_render(orclib.emul_build(sc, leaf), fr)
        for k in ("tri_id", "t", "rgb"):
            assert np.array_equal(e[k], g["%s_%s" % (name, k)]), (name, leaf, k)


def test_tile_sharding_covers_the_frame_once():
    """rt_map_pixel / rt_unpack_index (used by the kernels): every pixel is owned by exactly one rank and
    pack -> unpack is the identity, for sizes that are not multiples of the 16x8 tile."""
    sc = scenes.terrain_scene(16, 8)
    h = orclib.emul_build(sc, 4)
    lib = orclib.emul()
    from raytracinginonesemester_b200 import parallel
    lib.emu_render_rank.argtypes = [C.c_void_p, C.POINTER(A.rt_frame), C.POINTER(A.rt_image), C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_int]
    lib.emu_unpack_rgb8.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, A.u8p, A.u8p]
    for (W, H) in ((50, 21), (64, 32), (7, 3), (130, 70)):
        fr = scenes.terrain_frame(W, H, outputs=A.RT_OUT_RGB8)
        full = orclib.emul_render(h, fr, want=("rgb8",))["rgb8"]
        for world, cpr in ((2, 0), (3, 1), (8, 0), (2, 7), (8, 64)):
            img = np.full((H, W, 3), 255, np.uint8)
            total_prim = 0
            owner = parallel.tile_owner_map(W, H, world, cpr)
            for rank in range(world):
                ntl = parallel.tiles_of_rank(W, H, rank, world, cpr)
                packed = np.zeros((max(ntl, 1) * 128, 3), np.uint8)
                im = A.rt_image(); im.rgb8 = packed.ctypes.data_as(A.u8p)
                f = fr.c_struct()
                got = lib.emu_render_rank(h, C.byref(f), C.byref(im), None, rank, world, cpr)
                assert got == ntl
                assert im.rays_primary == int((owner == rank).sum())        # python mirror of the ownership map
                total_prim += im.rays_primary
                lib.emu_unpack_rgb8(W, H, world, cpr, rank, packed.ctypes.data_as(A.u8p), img.ctypes.data_as(A.u8p))
            assert total_prim == W * H
            assert np.array_equal(img, full)


def test_lazy_triangle_test_never_disagrees_with_the_reference_order_test(golden):
    """rt_moller_trumbore_lazy (division-free front end, what the packet kernels run) vs rt_moller_trumbore (the
    reference's operation order): the reference's own 8 directed cases and 57-point barycentric sweep
    (HW1/test_ray_tri_inter_STANDALONE/test_ray_triangle_inter.cpp:17-126) under both contracts with the approximate
    reciprocal at its worst, then 2e7 random probes aimed at edges, vertices and the t limits (1e8 were run once for
    DESIGN.md: 0 disagreements)."""
    lib = orclib.emul()
    lib.emu_mt_lazy.argtypes = [C.c_int, A.f32p, A.f32p, C.c_int, A.f32p, A.f32p, A.f32p, C.c_float, A.f32p]
    lib.emu_mt_lazy_sweep.restype = C.c_uint64
    lib.emu_mt_lazy_sweep.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    v = golden("ref_vectors.npz")
    tri, dirs, res = v["tri"], v["ray_dirs"], v["ray_results"]
    o = np.zeros(3, np.float32)
    fp = lambda a: a.ctypes.data_as(A.f32p)
    for scale in (1.0 - 2.0 ** -22, 1.0, 1.0 + 2.0 ** -22):
        for d, r in zip(dirs, res):
            t = C.c_float()
            for mode in (0, 1):
                assert lib.emu_mt_lazy(mode, fp(o), fp(d), 1, *(fp(x) for x in tri), scale, C.byref(t)) == int(r[2 * mode])
                assert not r[2 * mode] or t.value == r[2 * mode + 1]
    stats = (C.c_uint64 * 3)()
    bad = lib.emu_mt_lazy_sweep(20_000_000, 42, stats)
    assert bad == 0
    accepted, early, late = (int(x) for x in stats)
    assert accepted > 2_000_000                      # the probes do hit
    assert early > 5 * late                          # ... and most rejects leave before the IEEE divide, even with every probe aimed at a boundary
