"""Full-size parity on the BASELINE configurations, against the reference itself.

C4: the WHOLE 3840x2160 frame (8.3 M pixels) through the unmodified reference sources compiled in place
(oracle/_ref/libref_hw2.so, all host threads, ~2-4 s) vs the device: closest-hit ids, t, the crack pixels the reference
leaves (query.h:105: a ray through a shared edge can miss both triangles), and the 8-bit image.
C5: the real 7680x4320 x 16 spp frame of the 10M-triangle scene on the device, a strided row set through the reference.
Bars (BASELINE.json north_star): ids bit-exact on >= 99.99 % of pixels, mismatches only at epsilon ties, t within 1e-5
relative, 8-bit image within 1 LSB.  The epsilon, stated: where the ids differ both sides hit, the device's t is the
smaller one and the two differ by at most 1e-6 relative (observed on C4: <= 7.3e-7, i.e. 6 ulp).  Such a pixel looks
through the shared edge of two facets; both triangles are accepted by the (bit-identical) Moeller-Trumbore test with t a
few ulp apart (t = (e2 . qvec) / det carries the cancellation error of two cross products); the device
keeps the minimum over every accepted triangle (canonical rule), the reference keeps whichever its traversal reaches:
once it holds the farther hit, its fp64 slab test against the neighbour's UNPADDED box uses that t as the far bound
and can cull the neighbour by a few ulp (bvh.h:81-129, query.h:255-267).  At those pixels the two facets'
normals differ, so the colour may too; they are excluded from the 1-LSB image bar and counted."""
import numpy as np
import pytest

import orclib
from raytracinginonesemester_b200 import _abi as A, api, scenes

pytestmark = pytest.mark.gpu
ALL = A.RT_OUT_RGB_F32 | A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T


def run(renderer, frame):
    renderer.render(frame)
    return renderer.download()


def quantise(rgb, q):
    lib = orclib.oracle()
    flat = np.ascontiguousarray(rgb, np.float32).ravel()
    return np.fromiter((lib.orc_quantise(float(c), q) for c in flat), np.uint8, flat.size).reshape(rgb.shape)


def bar(got, ref, rows, what, quant, spp=1):
    """got: device planes (full frame); ref: reference planes (valid on `rows`)."""
    gid, rid = got["tri_id"][rows], ref["tri_id"][rows]
    gt, rt = got["t"][rows], ref["t"][rows]
    n = rid.size
    mism = gid != rid
    assert mism.sum() <= 1e-4 * n, "%s: %d of %d ids differ" % (what, mism.sum(), n)
    if mism.any():       # canonical rule (min t, then min id) vs the reference's traversal-order-dependent choice among epsilon ties
        assert np.all((gid[mism] >= 0) & (rid[mism] >= 0)), what + ": hit/miss flip"
        assert np.all(gt[mism] <= rt[mism]), what + ": the device's hit is farther than the reference's"
        assert np.all(rt[mism] - gt[mism] <= 1e-6 * rt[mism]), what + ": id mismatch that is not an epsilon tie (max %g relative)" % ((rt[mism] - gt[mism]) / rt[mism]).max()
    assert np.array_equal(gid < 0, rid < 0), what + ": crack / miss pixels differ"
    hit = rid >= 0
    rel = np.abs(gt[hit] - rt[hit]) / np.maximum(np.abs(rt[hit]), 1e-30)
    assert rel.size == 0 or rel.max() <= 1e-5, "%s: t rel err %g" % (what, rel.max())
    if "rgb" in ref:
        # quantise the reference's float image with the quantiser the frame asked for (a vectorised restatement of
        # ppm_p6.cpp:137-155 for RT_QUANT_PPM_LROUND; the oracle's quantiser is pinned to the reference writer elsewhere)
        assert quant == A.RT_QUANT_PPM_LROUND
        r8 = np.floor(np.clip(ref["rgb"][rows].astype(np.float64), 0.0, 1.0) * 255.0 + 0.5).astype(np.int32)
        d = np.abs(got["rgb8"][rows].astype(np.int32) - r8)
        ok = ~mism
        if spp == 1:
            assert d[ok].max() <= 1, "%s: 8-bit image differs by %d LSB" % (what, d[ok].max())
        else:
            # the id plane describes sample 0 only; any of the spp samples of a pixel can be an epsilon tie (a different facet, a
            # different colour for that sample).  The id budget of 1e-4 per sample therefore allows spp x 1e-4 of the pixels here.
            far = (d > 1).any(-1) & ok
            assert far.sum() <= spp * 1e-4 * n, "%s: %d pixels differ by more than 1 LSB" % (what, far.sum())
        return int(mism.sum()), int((rid < 0).sum()), float((d[ok] != 0).mean())
    return int(mism.sum()), int((rid < 0).sum()), None


def test_c4_whole_frame_against_the_reference_itself(renderer):
    sc = scenes.terrain_scene(1000, 500)
    info = renderer.upload_scene(sc)
    assert info.num_triangles == 1000000 and info.num_leaves == info.num_nodes + 1
    # compact 8-wide view (one WideNode per BVH2 node at depth 0, 3, 6, ...): nodes 64 B + 2 x 48 B per triangle + ~1/7 x 256 B per node
    assert info.arena_bytes <= 175e6, info.arena_bytes
    W, H = 3840, 2160
    fr = scenes.terrain_frame(W, H, outputs=A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T)
    got = run(renderer, fr)
    ref, kind = orclib.reference_render(sc, scenes.terrain_frame(W, H, outputs=ALL))
    mism, cracks, frac8 = bar(got, ref, slice(0, H), "c4 whole frame vs %s" % kind, fr.quantiser)
    print("c4 whole frame vs %s: %d id mismatches (all epsilon ties: device t <= reference t <= device t (1 + 1e-6)), %d crack pixels (identical set), %.4f %% of channels off by 1 LSB"
          % (kind, mism, cracks, 100.0 * (frac8 or 0.0)))
    # the same frame from the round-1 block-per-tile kernel and the per-lane traversal: identical planes
    for variant in (A.RT_VARIANT_FRUSTUM, A.RT_VARIANT_PACKET, A.RT_VARIANT_PERSIST_EXACT_MT):
        fr.kernel_variant = variant
        other = run(renderer, fr)
        for k in ("tri_id", "t", "rgb8"):
            assert np.array_equal(got[k], other[k]), (variant, k)
        assert other["rays_shadow"] == got["rays_shadow"]


def test_c5_real_frame_strided_rows_against_the_reference(renderer):
    sc = scenes.terrain_scene(2500, 2000)
    info = renderer.upload_scene(sc)
    assert info.num_triangles == 10000000
    W, H, spp = 7680, 4320, 16
    fr = scenes.terrain_frame(W, H, spp=spp, outputs=A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T)
    got = run(renderer, fr)
    assert got["rays_primary"] == W * H * spp
    again = run(renderer, fr)
    for k in ("tri_id", "t", "rgb8"):
        assert np.array_equal(got[k], again[k]), k
    step, first = 270, 11                               # 16 rows x 7680 px x 16 spp = 2.0 M reference rays
    ref, kind = orclib.reference_render(sc, scenes.terrain_frame(W, H, spp=spp, outputs=ALL), row_begin=first, row_step=step)
    mism, cracks, frac8 = bar(got, ref, slice(first, H, step), "c5 real frame vs %s" % kind, fr.quantiser, spp=spp)
    print("c5 8K x 16 spp, rows %d::%d vs %s: %d id mismatches, %d miss pixels, %.4f %% of channels off by 1 LSB" % (first, step, kind, mism, cracks, 100.0 * (frac8 or 0.0)))


def test_scene_bounds_with_negative_zero_coordinates(renderer):
    """float min/max through integer atomics must not let -0.0f (bit pattern INT_MIN) win a signed atomicMin
    or lose a signed atomicMax.  > 256 triangles per reduction block, -0.0 in one block, negative / positive values in
    others; several uploads (the outcome used to depend on block completion order)."""
    rng = np.random.default_rng(5)
    n = 4096
    pos = np.zeros((3 * n, 3), np.float32)
    pos[:, 0] = rng.uniform(-5.0, -1.0, 3 * n)          # x: every value negative, except ...
    pos[600:1200, 0] = -0.0                             # ... one block's worth of -0.0: the true maximum of x is -0.0
    pos[:, 1] = rng.uniform(1.0, 5.0, 3 * n)            # y: every value positive, except ...
    pos[1500:2100, 1] = -0.0                            # ... -0.0: the true minimum
    pos[:, 2] = rng.uniform(-3.0, 3.0, 3 * n)
    pos[2400:3000, 2] = -0.0
    idx = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    sc = api.Scene(pos, idx)
    for _ in range(8):
        info = renderer.upload_scene(sc)
        lo, hi = np.array(list(info.scene_min)), np.array(list(info.scene_max))
        assert np.array_equal(lo, pos.min(0)) and np.array_equal(hi, pos.max(0)), (lo, hi, pos.min(0), pos.max(0))


def test_frog_json_as_shipped_bounces_whole_frame_against_the_reference(renderer):
    """assets/json_files/frog.json AS SHIPPED — max_bounces 8, hash-RNG diffuse bounces (scene.h:18 default) — at its own
    1920x1080: the whole frame through the reference's TraceRayIterative (oracle/_ref, all host threads) vs the device.
    Primary ids / t to the same bars as C4; the image (a sum over up to eight path segments whose directions come from
    the deterministic per-pixel hash RNG, query.h:32-70) within 1 LSB on all but the epsilon-tie budget of pixels — a tie
    on any segment sends the rest of that path elsewhere."""
    import bench
    wl = bench.make_workload("c3bounce")
    sc = wl["scene"]()
    fr = wl["frame"]
    fr.outputs = ALL
    renderer.upload_scene(sc)
    got = run(renderer, fr)
    assert got["rays_primary"] > 1.05 * (got["tri_id"] >= 0).sum()               # bounce segments were traced
    ref, kind = orclib.reference_render(sc, fr, want=("rgb", "tri_id", "t"))
    rows = slice(0, fr.height)
    gid, rid = got["tri_id"], ref["tri_id"]
    n = rid.size
    mism = gid != rid
    assert mism.sum() <= 1e-4 * n, "%d of %d primary ids differ" % (mism.sum(), n)
    assert np.array_equal(gid < 0, rid < 0)
    hit = (rid >= 0) & ~mism
    rel = np.abs(got["t"][hit] - ref["t"][hit]) / np.maximum(np.abs(ref["t"][hit]), 1e-30)
    assert rel.max() <= 1e-5
    r8 = np.floor(np.clip(ref["rgb"].astype(np.float64), 0.0, 1.0) * 255.0 + 0.5).astype(np.int32)
    d = np.abs(got["rgb8"].astype(np.int32) - r8)
    far = (d > 1).any(-1)
    assert far.sum() <= 8 * 1e-4 * n, "%d of %d pixels differ by more than 1 LSB (%s)" % (far.sum(), n, kind)
    print("frog.json as shipped vs %s: %d id mismatches, %d pixels beyond 1 LSB of %d, %.2f %% of the bytes differ at all"
          % (kind, mism.sum(), far.sum(), n, 100.0 * (d != 0).mean()))
