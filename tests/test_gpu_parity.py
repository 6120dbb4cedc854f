"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle (oracle/), the
committed reference fixtures (tests/golden/) and size-independent properties at full size.

Bars (BASELINE.json north_star): closest-hit triangle id bit-exact on >= 99.99 % of pixels against
the reference CPU path (mismatches only at exact-t ties), hit t within 1e-5 relative, 8-bit image
within 1 LSB.  Against the canonical oracle (min t, then min id over the reference's
intersectTriangle on every triangle) the bar is bit-exact ids, t and float rgb."""
import ctypes as C

import numpy as np
import pytest

import orclib
from raytracinginonesemester_b200 import _abi as A, api, scenes

pytestmark = pytest.mark.gpu
ALL = A.RT_OUT_RGB_F32 | A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T


def run(renderer, frame):
    renderer.render(frame)
    return renderer.download()


def assert_reference_bar(got, ref_id, ref_t, ref_rgb=None, quant=A.RT_QUANT_HW2_TRUNC, what="", coplanar_overlaps=False):
    """coplanar_overlaps: the scene stacks distinct triangles in one plane (cornellbox.obj: block
    footprints on the floor), so whole regions are t-ties decided by traversal order in the reference
    (last visited wins, query.h:105) and by min id here; the id bar then applies to non-tied pixels."""
    n = ref_id.size
    mism = got["tri_id"] != ref_id
    if not coplanar_overlaps:
        assert mism.sum() <= 1e-4 * n, "%s: %d of %d ids differ" % (what, mism.sum(), n)
    # every id mismatch must be a t-tie (both hit, same t to 1e-6 relative) — the "stated epsilon ties"
    if mism.any():
        assert np.all((got["tri_id"][mism] >= 0) & (ref_id[mism] >= 0)), what + ": hit/miss flip"
        tie = np.abs(got["t"][mism] - ref_t[mism]) <= 1e-6 * np.abs(ref_t[mism])
        assert tie.all(), what + ": id mismatch that is not a t-tie"
        if not coplanar_overlaps:
            assert np.array_equal(got["t"][mism], ref_t[mism]), what + ": id mismatch that is not an exact-t tie"
    hit = (ref_id >= 0) & ~mism
    rel = np.abs(got["t"][hit] - ref_t[hit]) / np.maximum(np.abs(ref_t[hit]), 1e-30)
    assert rel.size == 0 or rel.max() <= 1e-5, "%s: t rel err %g" % (what, rel.max())
    if ref_rgb is not None:
        q = np.vectorize(lambda c: orclib.oracle().orc_quantise(float(c), quant), otypes=[np.uint8])
        ref8 = q(ref_rgb)
        d = np.abs(got["rgb8"].astype(int) - ref8.astype(int))
        ok = ~np.repeat(mism[..., None], 3, -1)
        assert d[ok].max() <= 1, "%s: 8-bit image differs by %d LSB" % (what, d[ok].max())


def _strided_bar(a, orc, rows, what):
    n = orc["tri_id"][rows].size
    mism = a["tri_id"][rows] != orc["tri_id"][rows]
    assert mism.sum() <= 1e-4 * n, "%s: %d of %d ids differ" % (what, mism.sum(), n)
    hit = (orc["tri_id"][rows] >= 0) & ~mism
    if hit.any():
        rel = np.abs(a["t"][rows][hit] - orc["t"][rows][hit]) / orc["t"][rows][hit]
        assert rel.max() <= 1e-5, what
    assert np.abs(a["rgb8"][rows].astype(int) - orc["rgb8"][rows].astype(int))[~mism].max() <= 1, what


# ------------------------------------------------------------------------------ HW1 ----
def test_hw1_frog_brute_matches_reference_fixture(renderer, frog_scene, golden):
    renderer.upload_scene(frog_scene)
    for name, col in (("white", (1, 1, 1)), ("magenta", (1, 0, 1))):
        ref = golden("ref_hw1_frog_96x54_%s.npz" % name)
        got = run(renderer, scenes.hw1_frame(96, 54, light_color=col, outputs=ALL))
        assert np.array_equal(got["tri_id"], ref["tri_id"])
        assert np.array_equal(got["t"], ref["t"])
        assert np.array_equal(got["rgb8"], ref["rgb8"])
        d = np.abs(got["rgb"] - ref["rgb"])
        assert d.max() <= 2e-7, d.max()      # powf: fp64 pow narrowed vs glibc powf


def test_hw1_frog_golden_png_bit_exact(renderer, frog_scene, golden):
    """HW1/frog_output.png, the reference's committed golden (rendered with a white light)."""
    renderer.upload_scene(frog_scene)
    got = run(renderer, scenes.hw1_frame(320, 180, light_color=(1, 1, 1), outputs=ALL))
    ref = golden("hw1_frog_output.npz")["rgb8"]
    assert np.array_equal(got["rgb8"], ref), "%d px differ" % (got["rgb8"] != ref).any(-1).sum()
    assert (got["tri_id"] >= 0).sum() == 3795           # hit mask of 3 795 px


def test_hw1_bvh_accel_equals_brute(renderer, frog_scene):
    renderer.upload_scene(frog_scene)
    a = run(renderer, scenes.hw1_frame(160, 90, light_color=(1, 1, 1), outputs=ALL, accel=A.RT_ACCEL_BRUTE))
    b = run(renderer, scenes.hw1_frame(160, 90, light_color=(1, 1, 1), outputs=ALL, accel=A.RT_ACCEL_BVH))
    for k in ("tri_id", "t", "rgb", "rgb8"):
        assert np.array_equal(a[k], b[k]), k


def test_hw1_sphere_as_the_repo_runs_it(renderer, golden):
    """C1: sphere.obj under the frog camera — every pixel is (20,5,5)."""
    d = golden("sphere_mesh.npz")
    renderer.upload_scene(api.Scene(d["positions"], d["indices"], normals=d["normals"], build_flags=A.RT_BUILD_NO_BVH))
    ref = golden("ref_hw1_sphere_64x36.npz")
    got = run(renderer, scenes.hw1_frame(64, 36, outputs=ALL))
    assert np.array_equal(got["rgb8"], ref["rgb8"]) and np.array_equal(got["tri_id"], ref["tri_id"])
    assert np.all(got["rgb8"] == np.array([20, 5, 5], np.uint8))


# ---------------------------------------------------------------------------- HW2 BVH ----
@pytest.mark.parametrize("name,filling", [("frog", False), ("frogfill", True)])
def test_hw2_frog_vs_reference_and_oracle(renderer, frog_scene, golden, name, filling):
    renderer.upload_scene(frog_scene)
    fr = scenes.frog_frame(160, 90, filling=filling, outputs=ALL)
    got = run(renderer, fr)
    ref = golden("ref_hw2_%s_160x90.npz" % name)
    assert_reference_bar(got, ref["tri_id"], ref["t"], ref["rgb"], what=name)
    fr.accel = A.RT_ACCEL_BRUTE
    orc = orclib.oracle_render(frog_scene, fr)        # canonical brute force on the CPU
    assert np.array_equal(got["tri_id"], orc["tri_id"]) and np.array_equal(got["t"], orc["t"])
    assert np.abs(got["rgb"] - orc["rgb"]).max() <= 2e-7
    assert np.abs(got["rgb8"].astype(int) - orc["rgb8"].astype(int)).max() <= 1


def test_hw2_terrain_vs_reference(renderer, golden):
    sc = scenes.terrain_scene(60, 30)
    renderer.upload_scene(sc)
    ref = golden("ref_hw2_terrain_128x72.npz")
    got = run(renderer, scenes.terrain_frame(128, 72, outputs=ALL, quantiser=A.RT_QUANT_HW2_TRUNC))
    assert_reference_bar(got, ref["tri_id"], ref["t"], ref["rgb"], what="terrain")
    assert got["rays_primary"] == 128 * 72 and got["rays_shadow"] > 0
    ref4 = golden("ref_hw2_terrain_64x36_spp4.npz")
    got4 = run(renderer, scenes.terrain_frame(64, 36, spp=4, outputs=ALL, quantiser=A.RT_QUANT_HW2_TRUNC))
    assert_reference_bar(got4, ref4["tri_id"], ref4["t"], ref4["rgb"], what="terrain spp4")
    assert np.abs(got4["rgb"] - ref4["rgb"]).max() <= 1e-6


def test_hw2_brute_accel_equals_bvh(renderer):
    sc = scenes.terrain_scene(24, 12)
    renderer.upload_scene(sc)
    a = run(renderer, scenes.terrain_frame(80, 45, spp=2, outputs=ALL, accel=A.RT_ACCEL_BVH))
    b = run(renderer, scenes.terrain_frame(80, 45, spp=2, outputs=ALL, accel=A.RT_ACCEL_BRUTE))
    for k in ("tri_id", "t", "rgb", "rgb8"):
        assert np.array_equal(a[k], b[k]), k
    assert a["rays_shadow"] == b["rays_shadow"]


def test_hw2_cornell_multi_material(renderer, golden):
    d = golden("cornell_mesh.npz")
    mats = [api.make_material(albedo=(0.7, 0.7, 0.7)), api.make_material(albedo=(0.8, 0.1, 0.1), ks=0.4, shininess=16.0),
            api.make_material(albedo=(0.1, 0.8, 0.1), kd=0.5, ks=0.5, specular_color=(0.9, 0.9, 0.9), shininess=64.0, emission=(0.05, 0.0, 0.0))]
    nobj = int(d["tri_obj_ids"].max()) + 2
    sc = api.Scene(d["positions"], d["indices"], normals=None, tri_obj_ids=d["tri_obj_ids"], materials=[mats[i % 3] for i in range(nobj)])
    renderer.upload_scene(sc)
    ref = golden("ref_hw2_cornell_96x96.npz")
    cam = ref["cam"]
    c = api.camera_init(cam[:3], cam[3:6], (0, 0, 1), 35.0, 24.0, 96, 96)
    lp = ref["lights"]
    fr = api.Frame(c, 96, 96, lights=[api.make_light(lp[0], (1, 1, 1), 2), api.make_light(lp[1], (0.4, 0.4, 1.0), 1)],
                   miss_color=(0.1, 0.2, 0.3), jitter=api.jitter_table(1, 42, True), outputs=ALL, quantiser=A.RT_QUANT_HW2_TRUNC)
    got = run(renderer, fr)
    assert_reference_bar(got, ref["tri_id"], ref["t"], ref["rgb"], what="cornell", coplanar_overlaps=True)
    assert (got["tri_id"] != ref["tri_id"]).mean() < 0.08


@pytest.mark.parametrize("leaf_max", [1, 2, 8])
def test_leaf_size_does_not_change_results(renderer, frog_scene, leaf_max):
    renderer.upload_scene(frog_scene)
    base = run(renderer, scenes.frog_frame(128, 72, filling=True, outputs=ALL))
    sc = api.Scene(frog_scene.positions, frog_scene.indices, normals=frog_scene.normals, tri_obj_ids=frog_scene.tri_obj_ids,
                   materials=frog_scene.materials, build_flags=A.RT_BUILD_LEAF_MAX(leaf_max))
    renderer.upload_scene(sc)
    got = run(renderer, scenes.frog_frame(128, 72, filling=True, outputs=ALL))
    for k in ("tri_id", "t", "rgb"):
        assert np.array_equal(base[k], got[k]), k


def test_device_bvh_is_sound_and_host_walk_agrees(renderer):
    """Download the BVH the device built, check it structurally and walk it on the host with the same
    per-ray functions: ids/t must equal what the kernel produced."""
    sc = scenes.terrain_scene(50, 25)
    info = renderer.upload_scene(sc)
    nodes, geom, ids = renderer.download_bvh()
    assert sorted(ids.tolist()) == list(range(sc.indices.shape[0]))
    s = sc.c_struct()
    import ctypes as C
    h = orclib.emul().emu_adopt(nodes.ctypes.data, int(info.num_nodes), geom.ctypes.data, int(info.num_triangles), C.byref(s))
    assert orclib.emul().emu_validate(h) == 0
    fr = scenes.terrain_frame(96, 54, outputs=ALL)
    got = run(renderer, fr)
    emu = orclib.emul_render(h, fr)
    for k in ("tri_id", "t", "rgb8"):
        assert np.array_equal(got[k], emu[k]), k
    fr.kernel_variant = A.RT_VARIANT_PER_RAY_STATS          # per-ray kernel: the same walk as the host emulation
    st = run(renderer, fr)
    nv, nt, nl, nb = renderer.frame_stats()
    assert (nv, nt) == (emu["stats"]["nodes"], emu["stats"]["tris"]) and (nl, nb) == (nv, nt)
    assert np.array_equal(st["tri_id"], got["tri_id"])
    fr.kernel_variant = A.RT_VARIANT_PACKET_STATS           # per-lane packet traversal: a lane tests a superset, results identical
    pk = run(renderer, fr)
    nvp, ntp, nlp, nbp = renderer.frame_stats()
    assert nvp >= nv and ntp >= nt
    assert nlp * 32 >= nvp and nbp * 32 >= ntp and nlp < nvp          # one line per warp visit, shared by its lanes
    for k in ("tri_id", "t", "rgb8"):
        assert np.array_equal(pk[k], got[k]), k
    fr.kernel_variant = A.RT_VARIANT_STATS                  # default = frustum traversal: a lane still tests a superset of its own triangles
    fk = run(renderer, fr)
    _, ntf, nlf, nbf = renderer.frame_stats()
    assert ntf >= nt and nbf * 32 >= ntf and nlf > 0
    for k in ("tri_id", "t", "rgb8"):
        assert np.array_equal(fk[k], got[k]), k


@pytest.mark.parametrize("variant", [A.RT_VARIANT_DEFAULT, A.RT_VARIANT_PACKET_OCC6, A.RT_VARIANT_PACKET_OCC10, A.RT_VARIANT_PACKET_EXACT_SLAB, A.RT_VARIANT_PER_RAY, A.RT_VARIANT_PACKET, A.RT_VARIANT_FRUSTUM])
def test_all_kernel_variants_agree(renderer, frog_scene, variant):
    renderer.upload_scene(frog_scene)
    base = run(renderer, scenes.frog_frame(200, 120, filling=True, outputs=ALL, accel=A.RT_ACCEL_BRUTE))
    fr = scenes.frog_frame(200, 120, filling=True, outputs=ALL)
    fr.kernel_variant = variant
    got = run(renderer, fr)
    for k in ("tri_id", "t", "rgb", "rgb8"):
        assert np.array_equal(base[k], got[k]), k
    assert got["rays_shadow"] == base["rays_shadow"]


def test_tiny_and_degenerate_scenes(renderer):
    # one triangle, two triangles, a degenerate (zero-area) triangle and duplicated triangles (exact ties)
    pos = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0], [0, 0, 0.5], [0, 0, 0.5], [0, 0, 0.5]], np.float32)
    for idx in ([[0, 1, 2]], [[0, 1, 2], [0, 2, 1]], [[0, 1, 2], [3, 4, 5]], [[0, 1, 2]] * 7 + [[3, 4, 5]] * 3):
        sc = api.Scene(pos, np.array(idx, np.uint32))
        renderer.upload_scene(sc)
        cam = api.camera_init((0.1, 0.05, 3), (0, 0, 0), (0, 1, 0), 50.0, 24.0, 40, 30)
        fr = api.Frame(cam, 40, 30, lights=[api.make_light((1, 1, 2), (1, 1, 1), 3)], miss_color=(0.2, 0.3, 0.4), outputs=ALL)
        got = run(renderer, fr)
        fr.accel = A.RT_ACCEL_BRUTE
        orc = orclib.oracle_render(sc, fr)
        assert np.array_equal(got["tri_id"], orc["tri_id"]) and np.array_equal(got["t"], orc["t"])
        assert np.abs(got["rgb"] - orc["rgb"]).max() <= 2e-7
        assert (got["tri_id"] >= 0).any() and (got["tri_id"] < 0).any()
        assert got["tri_id"].max() == 0                  # ties resolve to the lowest triangle id


def test_axis_parallel_rays_and_flat_boxes(renderer):
    """Camera looking straight down an axis at an axis-aligned plane: direction components that are
    exactly zero and zero-thickness boxes (the reference's origin-in-slab branch, bvh.h:90-91)."""
    g = np.linspace(-1, 1, 9, dtype=np.float32)
    xx, yy = np.meshgrid(g, g)
    pos = np.stack([xx.ravel(), yy.ravel(), np.zeros(81, np.float32)], 1)
    idx = []
    for j in range(8):
        for i in range(8):
            a = j * 9 + i
            idx += [[a, a + 1, a + 10], [a, a + 10, a + 9]]
    sc = api.Scene(pos, np.array(idx, np.uint32))
    renderer.upload_scene(sc)
    cam = api.camera_init((0, 0, 2), (0, 0, 0), (0, 1, 0), 30.0, 24.0, 33, 33)   # centre pixel ray = (0,0,-1) exactly
    fr = api.Frame(cam, 33, 33, lights=[api.make_light((0, 0, 5), (1, 1, 1), 2)], outputs=ALL)
    got = run(renderer, fr)
    fr.accel = A.RT_ACCEL_BRUTE
    orc = orclib.oracle_render(sc, fr)
    assert np.array_equal(got["tri_id"], orc["tri_id"]) and np.array_equal(got["t"], orc["t"])
    assert got["tri_id"][16, 16] >= 0
    # ... and against the reference's SearchBVH over its own LBVH, whose |d| < 1e-8 branch NARROWS acceptance to
    # "origin inside the slab" (bvh.h:90-91,103-104,116-117): a hit the reference culls there would show as a hit/miss flip
    fr.accel = A.RT_ACCEL_BVH
    ref = orclib.oracle_render(sc, fr, bvh=orclib.oracle_bvh(sc))
    assert np.array_equal(got["tri_id"] >= 0, ref["tri_id"] >= 0)
    # (pixel rays pass exactly through grid lines of the mesh: exact-t ties between the triangles that share an edge or a vertex are
    # the rule here, so the id bar applies to non-tied pixels only — every id difference must be a t-tie, and hit/miss must agree)
    assert_reference_bar(got, ref["tri_id"], ref["t"], what="axis-parallel rays vs reference BVH", coplanar_overlaps=True)
    # (the reference compiled in place cannot render this frame: its pixel loop always applies jittered_samples(spp, 42),
    # query.cu:142-148, and a jittered centre ray is no longer axis-parallel — the restated SearchBVH above is the check)
    # the same with the camera rays exactly parallel to x and y (flat boxes seen edge-on: every ray misses, in both)
    for pos_, up_ in (((3, 0, 0), (0, 0, 1)), ((0, 3, 0), (0, 0, 1))):
        fr2 = api.Frame(api.camera_init(pos_, (0, 0, 0), up_, 30.0, 24.0, 33, 33), 33, 33, lights=[api.make_light((0, 0, 5), (1, 1, 1), 2)], outputs=ALL)
        g2 = run(renderer, fr2)
        r2 = orclib.oracle_render(sc, fr2, bvh=orclib.oracle_bvh(sc))
        assert np.array_equal(g2["tri_id"] >= 0, r2["tri_id"] >= 0)
        assert_reference_bar(g2, r2["tri_id"], r2["t"], what="edge-on rays vs reference BVH", coplanar_overlaps=True)


# ------------------------------------------------------------------------ bounce loop ----
def test_hw2_bounce_loop_vs_reference_fixture(renderer, golden):
    """max_depth > 1: mirror and hash-RNG diffuse bounces of TraceRayIterative on the device against
    frames rendered by the reference (bit-exact ids, t and float rgb on this scene) and against the oracle."""
    import os
    g = golden("ref_hw2_bounce_cornell.npz")
    sc, cam, lights, miss = scenes.cornell_bounce_scene(os.path.join(os.path.dirname(__file__), "golden", "cornell_mesh.npz"))
    renderer.upload_scene(sc)
    for name, W, H, spp, depth, diffuse in g["cases"]:
        fr = scenes.cornell_bounce_frame(cam, lights, miss, int(W), int(H), int(spp), int(depth), int(diffuse), outputs=ALL)
        a = run(renderer, fr)
        assert a["rays_primary"] > int(W) * int(H) * int(spp)
        for k in ("tri_id", "t"):
            assert np.array_equal(a[k], g["%s_%s" % (name, k)]), (name, k)
        assert np.abs(a["rgb"] - g["%s_rgb" % name]).max() <= 2e-6, name          # powf: fp64 pow narrowed vs glibc powf
        o = orclib.oracle_render(sc, fr)
        assert np.array_equal(a["rgb8"], o["rgb8"]), name
        assert a["rays_primary"] == o["counters"]["rays_primary"] and a["rays_shadow"] == o["counters"]["rays_shadow"]


def test_bounce_frame_at_size_terrain_mirror(renderer):
    """A mirror-like terrain at 1920x1080, depth 4: determinism, more closest-hit queries than pixels, strided rows
    against the reference-exact oracle."""
    pos, idx = scenes.terrain(300, 150)
    sc = api.Scene(pos, idx, tri_obj_ids=np.zeros(idx.shape[0], np.int32),
                   materials=[api.make_material(albedo=(0.5, 0.5, 0.6), kd=0.5, ks=0.2, kr=0.5, specular_color=(0.9, 0.9, 0.9))])
    renderer.upload_scene(sc)
    W, H = 1920, 1080
    cam = api.camera_init((0.3, -1.4, 0.6), (0, 0, 0), (0, 0, 1), 30.0, 24.0, W, H)
    mk = lambda outs: api.Frame(cam, W, H, lights=[api.make_light((-2.0, -1.0, 1.5), (1, 1, 1), 5)], miss_color=(0.5, 0.7, 1.0), spp=1,
                                jitter=api.jitter_table(1, 42, True), max_depth=4, diffuse_bounce=True, outputs=outs, quantiser=A.RT_QUANT_HW2_TRUNC)
    a = run(renderer, mk(A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T))
    b = run(renderer, mk(A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T))
    for k in ("tri_id", "t", "rgb8"):
        assert np.array_equal(a[k], b[k]), k
    assert a["rays_primary"] > 1.2 * (a["tri_id"] >= 0).sum()
    bvh = orclib.oracle_bvh(sc)
    step = 120
    orc = orclib.oracle_render(sc, mk(ALL), bvh=bvh, row_begin=9, row_step=step, want=("rgb8", "tri_id", "t"))
    _strided_bar(a, orc, slice(9, H, step), "bounce terrain")


# --------------------------------------------------------------- CPUOnly renderer (N1) ----
@pytest.mark.parametrize("name", ["sphere_point", "sphere", "cornell"])
def test_cpuonly_mode_vs_reference_fixture(renderer, golden, name):
    """RT_MODE_HW2_CPU on the device against frames rendered by the CPUOnly reference: ids and t bit-exact, float rgb to
    powf rounding, 8-bit image equal to the oracle's."""
    g = golden("cpuonly_scenes.npz")
    sc, fr = scenes.cpuonly_case(g, name, outputs=ALL)
    renderer.upload_scene(sc)
    a = run(renderer, fr)
    for k in ("tri_id", "t"):
        assert np.array_equal(a[k], g["%s_%s" % (name, k)]), (name, k)
    assert np.abs(a["rgb"] - g["%s_rgb" % name]).max() <= 2e-6, name
    o = orclib.oracle_render(sc, fr)
    assert np.abs(a["rgb8"].astype(int) - o["rgb8"].astype(int)).max() <= 1
    assert a["rays_primary"] == o["counters"]["rays_primary"] and a["rays_shadow"] == o["counters"]["rays_shadow"]


def test_cpuonly_committed_golden_png(renderer, golden):
    """The reference's committed CPUOnly/output/sphere_point_output.png (360x240) from the device: within 1 LSB, and
    identical where powf rounding does not sit on a quantiser boundary."""
    g = golden("cpuonly_scenes.npz")
    png = g["sphere_point_golden_png"]
    sc, fr = scenes.cpuonly_case(g, "sphere_point", width=png.shape[1], height=png.shape[0], outputs=A.RT_OUT_RGB8)
    renderer.upload_scene(sc)
    a = run(renderer, fr)
    d = np.abs(a["rgb8"].astype(int) - png.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3, (d.max(), (d > 0).sum())


def test_cpuonly_mode_rejects_what_has_no_deterministic_reference(renderer, golden):
    g = golden("cpuonly_scenes.npz")
    sc, fr = scenes.cpuonly_case(g, "sphere_point")
    renderer.upload_scene(sc)
    fr.diffuse_bounce = True
    with pytest.raises(api.RtError):
        renderer.render(fr)
    fr.diffuse_bounce = False
    fr.accel = A.RT_ACCEL_BRUTE
    with pytest.raises(api.RtError):
        renderer.render(fr)


# --------------------------------------------------------------- full-size properties ----
def test_c4_full_size_properties(renderer):
    """BASELINE config C4 (1M-triangle terrain, 3840x2160, primary + shadow): properties that do not
    need a full CPU render + a strided comparison with the reference-exact oracle."""
    sc = scenes.terrain_scene(1000, 500)
    info = renderer.upload_scene(sc)
    assert info.num_triangles == 1000000
    W, H = 3840, 2160
    fr = scenes.terrain_frame(W, H, outputs=A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T)
    a = run(renderer, fr)
    assert a["rays_primary"] == W * H
    assert (a["tri_id"] >= 0).mean() > 0.9999          # straight-down camera: every ray hits (bar cracks)
    b = run(renderer, fr)
    for k in ("tri_id", "t", "rgb8"):                   # idempotence / determinism
        assert np.array_equal(a[k], b[k]), k
    # triangles of one grid cell are 2c, 2c+1: hit ids must walk the grid monotonically along a row
    cell = a["tri_id"][H // 2] // 2
    assert np.all(np.diff(cell[cell >= 0] % 1000) >= 0)
    # strided rows against the reference-exact CPU oracle (reference LBVH + SearchBVH restated)
    bvh = orclib.oracle_bvh(sc)
    step = 135
    orc = orclib.oracle_render(sc, scenes.terrain_frame(W, H, outputs=ALL), bvh=bvh, row_begin=7, row_step=step, want=("rgb8", "tri_id", "t"))
    rows = slice(7, H, step)
    n = orc["tri_id"][rows].size
    mism = a["tri_id"][rows] != orc["tri_id"][rows]
    assert mism.sum() <= 1e-4 * n, mism.sum()
    hit = (orc["tri_id"][rows] >= 0) & ~mism
    rel = np.abs(a["t"][rows][hit] - orc["t"][rows][hit]) / orc["t"][rows][hit]
    assert rel.max() <= 1e-5
    assert np.abs(a["rgb8"][rows].astype(int) - orc["rgb8"][rows].astype(int))[~mism].max() <= 1
    # brute force on the device at a size it finishes quickly: BVH result == every-triangle result
    small = scenes.terrain_frame(240, 135, outputs=ALL)
    x = run(renderer, small)
    small.accel = A.RT_ACCEL_BRUTE
    y = run(renderer, small)
    for k in ("tri_id", "t", "rgb8"):
        assert np.array_equal(x[k], y[k]), k


def test_c2_full_size_hw1_brute_force(renderer, frog_scene):
    """BASELINE config C2 (HW1 frog, 1920x1080, every ray against every triangle): determinism, the HW1 BVH
    frame must be the same image, strided rows against the oracle's HW1 loop."""
    renderer.upload_scene(frog_scene)
    W, H = 1920, 1080
    fr = scenes.hw1_frame(W, H, accel=A.RT_ACCEL_BRUTE, outputs=A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T)
    a = run(renderer, fr)
    assert a["rays_primary"] == W * H and a["rays_shadow"] == 0
    assert 0.05 < (a["tri_id"] >= 0).mean() < 0.9
    fr.accel = A.RT_ACCEL_BVH
    b = run(renderer, fr)
    for k in ("tri_id", "t", "rgb8"):
        assert np.array_equal(a[k], b[k]), k
    step = 90
    fo = scenes.hw1_frame(W, H, accel=A.RT_ACCEL_BRUTE, outputs=ALL)
    orc = orclib.oracle_render(frog_scene, fo, row_begin=11, row_step=step, want=("rgb8", "tri_id", "t"))
    rows = slice(11, H, step)
    for k in ("tri_id", "t", "rgb8"):                   # canonical oracle == HW1 loop: bit-exact
        assert np.array_equal(a[k][rows], orc[k][rows]), k


@pytest.mark.parametrize("filling", [False, True])
def test_c3_full_size_hw2_frog_4k(renderer, frog_scene, filling):
    """BASELINE config C3 (HW2-BVH frog at 3840x2160), stock frog.json view and the frame-filling view:
    determinism, coverage, and strided rows against the reference-exact oracle (reference LBVH + SearchBVH)."""
    renderer.upload_scene(frog_scene)
    W, H = 3840, 2160
    outs = A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T
    fr = scenes.frog_frame(W, H, filling=filling, outputs=outs)
    a = run(renderer, fr)
    b = run(renderer, fr)
    for k in ("tri_id", "t", "rgb8"):
        assert np.array_equal(a[k], b[k]), k
    cover = (a["tri_id"] >= 0).mean()
    assert (cover > 0.4) if filling else (0.02 < cover < 0.05)      # the stock view hits on ~3.2 % of pixels
    assert a["rays_primary"] == W * H and 0 < a["rays_shadow"] <= (a["tri_id"] >= 0).sum()
    bvh = orclib.oracle_bvh(frog_scene)
    step = 180
    orc = orclib.oracle_render(frog_scene, scenes.frog_frame(W, H, filling=filling, outputs=ALL), bvh=bvh, row_begin=3, row_step=step,
                               want=("rgb8", "tri_id", "t"))
    _strided_bar(a, orc, slice(3, H, step), "c3 filling=%s" % filling)


def test_c5_scene_at_reduced_frame(renderer):
    """BASELINE config C5's scene (10M-triangle terrain, 16 spp jitter table) at a frame the oracle can check:
    device BVH over 10M triangles, 16-sample accumulation, strided rows against the reference-exact oracle."""
    sc = scenes.terrain_scene(2500, 2000)
    info = renderer.upload_scene(sc)
    assert info.num_triangles == 10000000
    W, H = 768, 432
    fr = scenes.terrain_frame(W, H, spp=16, outputs=A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T)
    a = run(renderer, fr)
    assert a["rays_primary"] == W * H * 16
    assert (a["tri_id"] >= 0).mean() > 0.9999
    bvh = orclib.oracle_bvh(sc)
    step = 54
    orc = orclib.oracle_render(sc, scenes.terrain_frame(W, H, spp=16, outputs=ALL), bvh=bvh, row_begin=5, row_step=step, want=("rgb8", "tri_id", "t"))
    _strided_bar(a, orc, slice(5, H, step), "c5 scene")


def test_device_transform_bake_matches_reference(renderer, golden, tmp_path):
    """rt_scene.transforms (applyObjectTransform on the device, main.cu:57-96): the baked vertices, read back from the
    triangle blocks, equal the positions the reference's own loader + applyObjectTransform produced (fixture
    e2e_scene.npz) bit for bit, and the frame (normals included) equals the one rendered from the host-baked mesh
    (host ingest is itself pinned to the reference by tests/test_host_ingest.py)."""
    import json
    import os
    g = golden("e2e_scene.npz")
    open(tmp_path / "ball.obj", "w").write(str(g["ball_obj"]))
    open(tmp_path / "ground.obj", "w").write(str(g["ground_obj"]))
    scene = json.loads(str(g["json_mirror"]))
    raw_p, raw_n, host_p, host_n, idx, obj, xf = [], [], [], [], [], [], []
    off = nid = 0
    for k, o in enumerate(scene["scene"]):
        path = os.path.join(str(tmp_path), o["path"][2:])
        t = o["transform"]
        p0, n0, i0, o0, nxt = api.load_obj(path, nid)
        p1, n1, _, _, _ = api.load_obj(path, nid, position=t["position"], rotation=t["rotation"], scale=t["scale"])
        assert np.array_equal(p1, g["obj%d_positions" % k])
        xf.append((off, p0.shape[0], t["position"], t["rotation"], t["scale"]))
        raw_p.append(p0); raw_n.append(n0); host_p.append(p1); host_n.append(n1); idx.append(i0 + off); obj.append(o0)
        off += p0.shape[0]; nid = nxt
    idx, obj, host_p = np.concatenate(idx), np.concatenate(obj), np.concatenate(host_p)
    mats = [api.make_material(albedo=(0.7, 0.6, 0.5), ks=0.2) for _ in range(nid)]
    dev = api.Scene(np.concatenate(raw_p), idx, normals=np.concatenate(raw_n), tri_obj_ids=obj, materials=mats, transforms=xf)
    info = renderer.upload_scene(dev)
    _, blocks, ids = renderer.download_bvh()
    v0 = np.zeros((info.num_triangles, 3), np.float32)
    v0[ids] = blocks[:, 0:3]
    assert np.array_equal(v0, host_p[idx[:, 0]]), "device-baked vertices differ from applyObjectTransform"
    cam = api.camera_init((0.0, -2.5, 1.2), (0.0, 0.0, 0.4), (0, 0, 1), 28.0, 24.0, 160, 96)
    fr = api.Frame(cam, 160, 96, lights=[api.make_light((-2.0, -1.0, 1.5), (1, 1, 1), 5)], miss_color=(0.5, 0.7, 1.0), spp=1,
                   jitter=api.jitter_table(1, 42, True), outputs=ALL)
    a = run(renderer, fr)
    renderer.upload_scene(api.Scene(host_p, idx, normals=np.concatenate(host_n), tri_obj_ids=obj, materials=mats))
    b = run(renderer, fr)
    for k in ("tri_id", "t", "rgb"):
        assert np.array_equal(a[k], b[k]), k
    bad = api.Scene(np.concatenate(raw_p), idx, tri_obj_ids=obj, materials=mats, transforms=[(10, 10 ** 9, (0, 0, 0), (0, 0, 0), (1, 1, 1))])
    with pytest.raises(api.RtError):
        renderer.upload_scene(bad)


@pytest.mark.parametrize("spp", [2, 3, 4, 6, 8, 16, 32, 64])
def test_sample_major_packets_equal_pixel_major(renderer, frog_scene, spp):
    """spp > 1: the default kernel traces G samples of one pixel side by side (G = largest power of two dividing spp, <= 32)
    and sums them per pixel in sample order; it must equal the pixel-major kernel and the per-ray kernel bit for bit
    (float accumulation order included), on frames that are not multiples of the tile."""
    renderer.upload_scene(frog_scene)
    out = {}
    for variant in (A.RT_VARIANT_DEFAULT, A.RT_VARIANT_PACKET_PIXEL_MAJOR, A.RT_VARIANT_PER_RAY, A.RT_VARIANT_PACKET, A.RT_VARIANT_FRUSTUM):
        fr = scenes.frog_frame(203, 77, filling=True, outputs=ALL)
        fr.spp, fr.jitter, fr.kernel_variant = spp, api.jitter_table(spp, 42, True), variant
        out[variant] = run(renderer, fr)
        assert out[variant]["rays_primary"] == 203 * 77 * spp
    for variant in (A.RT_VARIANT_PACKET_PIXEL_MAJOR, A.RT_VARIANT_PER_RAY, A.RT_VARIANT_PACKET, A.RT_VARIANT_FRUSTUM):
        for k in ("rgb", "rgb8", "tri_id", "t"):
            assert np.array_equal(out[A.RT_VARIANT_DEFAULT][k], out[variant][k]), (spp, variant, k)
        assert out[A.RT_VARIANT_DEFAULT]["rays_shadow"] == out[variant]["rays_shadow"]


def _both_traversals(renderer, frame):
    out = {}
    for variant in (A.RT_VARIANT_PACKET, A.RT_VARIANT_FRUSTUM):
        frame.kernel_variant = variant
        out[variant] = run(renderer, frame)
    a, b = out[A.RT_VARIANT_PACKET], out[A.RT_VARIANT_FRUSTUM]
    for k in ("rgb", "rgb8", "tri_id", "t"):
        assert np.array_equal(a[k], b[k]), k
    assert a["rays_primary"] == b["rays_primary"] and a["rays_shadow"] == b["rays_shadow"]
    return a


def test_frustum_traversal_equals_per_lane_traversal(renderer, frog_scene, golden):
    """The frustum-culled wide traversal (one lane = one box; inner nodes culled against the packet's bounding planes)
    may only make a lane test MORE triangles than the per-lane traversal: every plane must be identical, bit for bit —
    on terrain (primary + shadow packets, 1 and 4 spp, frame not a multiple of the tile), on the frog (mostly-miss stock
    view, frame-filling view), inside the Cornell box (wide frusta, axis-aligned walls, lights behind the camera), on
    tiny / degenerate / tied scenes, with two lights, and with the camera far outside the scene (exact-slab path)."""
    renderer.upload_scene(scenes.terrain_scene(200, 100))
    for W, H, spp in ((640, 360, 1), (333, 187, 4), (64, 40, 16)):
        got = _both_traversals(renderer, scenes.terrain_frame(W, H, spp=spp, outputs=ALL))
        assert (got["tri_id"] >= 0).mean() > 0.99 and got["rays_shadow"] > 0
    fr = scenes.terrain_frame(320, 180, outputs=ALL)        # grazing view: long thin frusta, large depth range
    fr.cam = api.camera_init((-1.4, -0.2, 0.12), (0.5, 0.1, 0.0), (0, 0, 1), 18.0, 24.0, 320, 180)
    fr.lights = [api.make_light((-2.0, -1.0, 1.5), (1, 1, 1), 5), api.make_light((1.5, 0.8, 0.3), (1, 0.5, 0.2), 3)]
    got = _both_traversals(renderer, fr)
    assert (got["tri_id"] >= 0).any() and (got["tri_id"] < 0).any()
    fr.cam = api.camera_init((0.0, 0.0, 400.0), (0, 0, 0), (0, 1, 0), 4000.0, 24.0, 320, 180)   # > 8 scene extents away
    _both_traversals(renderer, fr)
    renderer.upload_scene(frog_scene)
    for filling in (False, True):
        _both_traversals(renderer, scenes.frog_frame(320, 200, filling=filling, outputs=ALL))
    d = golden("cornell_mesh.npz")
    sc = api.Scene(d["positions"], d["indices"], normals=d["normals"] if d["normals"].size else None, tri_obj_ids=d["tri_obj_ids"],
                   materials=[api.make_material(albedo=(0.7, 0.6, 0.5), kd=0.9, ks=0.2) for _ in range(int(d["tri_obj_ids"].max()) + 1)])
    renderer.upload_scene(sc)
    lo, hi = d["positions"].min(0), d["positions"].max(0)
    mid = 0.5 * (lo + hi)
    for pos, look, focal in (((mid[0], mid[1], lo[2] + 0.05 * (hi[2] - lo[2])), tuple(mid), 12.0), (tuple(mid), (hi[0], hi[1], mid[2]), 8.0),
                             ((mid[0], mid[1], lo[2] - 1.5 * (hi[2] - lo[2])), tuple(mid), 35.0)):
        cam = api.camera_init(pos, look, (0, 1, 0), focal, 24.0, 200, 160)
        fr = api.Frame(cam, 200, 160, lights=[api.make_light((mid[0], hi[1] - 0.05 * (hi[1] - lo[1]), mid[2]), (1, 1, 1), 60000),
                                              api.make_light(tuple(lo - 0.3 * (hi - lo)), (0.3, 0.4, 1.0), 90000)],
                       miss_color=(0.1, 0.1, 0.2), outputs=ALL)
        _both_traversals(renderer, fr)
    pos = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0], [0, 0, 0.5], [0, 0, 0.5], [0, 0, 0.5]], np.float32)
    for idx in ([[0, 1, 2]], [[0, 1, 2], [0, 2, 1]], [[0, 1, 2], [3, 4, 5]], [[0, 1, 2]] * 7 + [[3, 4, 5]] * 3):
        renderer.upload_scene(api.Scene(pos, np.array(idx, np.uint32)))
        cam = api.camera_init((0.1, 0.05, 3), (0, 0, 0), (0, 1, 0), 50.0, 24.0, 40, 30)
        _both_traversals(renderer, api.Frame(cam, 40, 30, lights=[api.make_light((1, 1, 2), (1, 1, 1), 3)], miss_color=(0.2, 0.3, 0.4), outputs=ALL))


def test_render_into_equals_render_plus_download(renderer, frog_scene):
    """rt_render_into (band-pipelined render + copy) must deliver exactly what rt_render + rt_download_image deliver,
    for every plane, odd sizes, frames smaller than a band, all modes and multi-sample frames."""
    renderer.upload_scene(frog_scene)
    frames = [scenes.frog_frame(333, 201, filling=True, outputs=ALL), scenes.frog_frame(64, 8, filling=True, outputs=ALL),
              scenes.frog_frame(17, 100, filling=True, outputs=A.RT_OUT_RGB8 | A.RT_OUT_T), scenes.hw1_frame(200, 120, accel=A.RT_ACCEL_BVH, outputs=ALL),
              scenes.frog_frame(1920, 1080, filling=True, outputs=A.RT_OUT_RGB8)]
    frames[0].spp = 3; frames[0].jitter = api.jitter_table(3, 42, True)
    for fr in frames:
        a = run(renderer, fr)
        b = renderer.render_into(fr)
        assert (a["rays_primary"], a["rays_shadow"]) == (b["rays_primary"], b["rays_shadow"])
        for k in ("rgb", "rgb8", "tri_id", "t"):
            if k in a:
                assert np.array_equal(a[k], b[k]), (fr.width, fr.height, k)
    fr = scenes.frog_frame(64, 64, filling=True, outputs=A.RT_OUT_RGB8)
    img = A.rt_image(); buf = np.zeros((64, 64), np.float32); img.t = buf.ctypes.data_as(A.f32p)
    f = fr.c_struct()
    assert renderer.lib.rt_render_into(renderer.ctx, C.byref(f), C.byref(img)) == A.RT_ERR_STATE      # plane not requested


# ----------------------------------------------------------------------- error behaviour ----
def test_error_behaviour():
    r = api.Renderer(0)
    cam = api.camera_init((0, 0, 1), (0, 0, 0), (0, 1, 0), 24, 24, 8, 8)
    with pytest.raises(api.RtError) as e:
        r.render(api.Frame(cam, 8, 8))
    assert e.value.code == A.RT_ERR_STATE
    with pytest.raises(api.RtError):
        r.upload_scene(api.Scene(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32)))
    with pytest.raises(api.RtError):
        r.upload_scene(api.Scene(np.zeros((3, 3), np.float32), np.array([[0, 1, 7]], np.uint32)))   # index out of range
    r.upload_scene(scenes.terrain_scene(4, 4, build_flags=A.RT_BUILD_NO_BVH))
    with pytest.raises(api.RtError) as e:
        r.render(api.Frame(cam, 8, 8, accel=A.RT_ACCEL_BVH))
    assert e.value.code == A.RT_ERR_STATE
    with pytest.raises(api.RtError) as e:
        r.render(api.Frame(cam, 0, 8, accel=A.RT_ACCEL_BRUTE))
    assert e.value.code == A.RT_ERR_ARG
    with pytest.raises(api.RtError):
        r.render(api.Frame(cam, 8, 8, mode=A.RT_MODE_HW1, accel=A.RT_ACCEL_BRUTE))   # HW1 needs a light
    with pytest.raises(api.RtError):
        api.camera_init((0, 0, 1), (0, 0, 0), (0, 1, 0), 24, 24, 0, 8)              # HW1 camera throws on W < 1
    fr = api.Frame(cam, 8, 8, accel=A.RT_ACCEL_BRUTE, outputs=A.RT_OUT_RGB8)
    r.render(fr)
    fr.outputs = A.RT_OUT_RGB_F32
    with pytest.raises(api.RtError):
        r.download()                                                                 # plane not requested
    r.close()


def test_ppm_bytes_match_reference_writer(renderer, frog_scene, tmp_path):
    """The u8 plane quantised on the device equals the bytes the reference's ppm_p6 writer emits for
    the float image (ppm_p6.cpp:137-155, 257-301)."""
    libs = orclib.ref_libs()
    if "ref_ppm" not in libs:
        pytest.skip("oracle/_ref/libref_ppm.so not built (needs /root/reference at build time)")
    import ctypes as C
    renderer.upload_scene(frog_scene)
    for quant, gamma in ((A.RT_QUANT_PPM_LROUND, 0), (A.RT_QUANT_PPM_GAMMA2, 1)):
        got = run(renderer, scenes.frog_frame(160, 90, filling=True, outputs=ALL, quantiser=quant))
        path = str(tmp_path / ("q%d.ppm" % quant))
        rc = libs["ref_ppm"].ref_ppm_write_rgbf(path.encode(), 160, 90, got["rgb"].ctypes.data_as(A.f32p), 255, 1, gamma, 0, None, 0)
        assert rc == 0
        data = open(path, "rb").read()
        head = b"P6\n160 90\n255\n"
        assert data[:len(head)] == head
        assert data[len(head):] == got["rgb8"].tobytes()


def test_compact_wide_view_phases_render_the_same_image(frog_scene, monkeypatch):
    """rt_build_wide keeps WideNodes only for every third BVH2 depth; forcing each phase (RT_B200_WIDE_PHASE; -1 = the
    smallest) changes the arena size and the traversal order, never the image."""
    from raytracinginonesemester_b200 import Renderer
    fr = scenes.frog_frame(320, 180, filling=True, outputs=ALL)
    base, sizes = None, {}
    for phase in ("", "0", "1", "2", "-1"):
        monkeypatch.setenv("RT_B200_WIDE_PHASE", phase)
        r = Renderer(0)
        try:
            info = r.upload_scene(frog_scene)
            sizes[phase] = int(info.arena_bytes)
            got = run(r, fr)
        finally:
            r.close()
        if base is None:
            base = got
        for k in ("tri_id", "t", "rgb", "rgb8"):
            assert np.array_equal(base[k], got[k]), (phase, k)
    assert sizes[""] == sizes["0"] and sizes["-1"] == min(sizes[p] for p in "012")
    assert sizes[""] < 64 * info.num_nodes + 96 * info.num_triangles + 256 * info.num_nodes // 2      # well under one WideNode per node
