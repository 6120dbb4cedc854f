"""Soft shadows of the CPUOnly renderer (CPUOnly/include/raytracer.h:37-46, 76-93, 121-168): disk light,
shadow_samples rays per lit hit, visibility = unoccluded / samples.

The reference draws its disk samples from a process-wide std::mt19937 seeded by std::random_device: it has no
reproducible output, so parity with it is STATISTICAL — tests/golden/cpuonly_area.npz holds the per-pixel mean and
standard deviation of 64 runs of the unmodified reference on its own config/sphere_area.json (tools/make_golden_area.py).
The product and the oracle take the samples from the reference's other generator (the per-pixel hash RNG of
GPUandCPU/include/query.h:32-48) and must agree with EACH OTHER bit for bit."""
import numpy as np
import pytest

import orclib
from raytracinginonesemester_b200 import _abi as A, scenes

ALL = A.RT_OUT_RGB_F32 | A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T


def against_reference_statistics(rgb, g, what):
    mean, std = g["mean"].astype(np.float64), g["std"].astype(np.float64)
    runs = int(g["frame"][3])
    # outside the penumbra all 64 reference runs agree: so must we (to float rounding of the 64-term mean)
    flat = (std <= 1e-7).all(-1)
    off = np.abs(rgb[flat] - mean[flat]).max(-1) > 2e-6
    assert off.mean() <= 2e-3, "%s: %d of %d pixels outside the reference's penumbra differ from it" % (what, off.sum(), flat.sum())
    # inside: one of our frames is one more draw from the same per-pixel distribution (visibility k/S, k ~ Binomial):
    # z = (ours - reference mean) / reference std.  The channels move together; use the one with the largest spread.
    pen = ~flat
    ch = std.reshape(-1, 3).max(0).argmax()
    z = (rgb[..., ch][pen] - mean[..., ch][pen]) / np.maximum(std[..., ch][pen], 1e-9)
    # pixels whose 64 runs saw only one or two occluded samples have a tiny std estimate: exclude std below one sample's weight / 20
    solid = std[..., ch][pen] > 0.05 * std[..., ch][pen].max()
    z = z[solid]
    assert z.size > 300, z.size
    se = np.sqrt(1.0 / z.size + 1.0 / runs)               # our draw + the error of the 64-run mean
    assert abs(z.mean()) <= 4.0 * se, "%s: biased against the reference: mean z = %.3f (n = %d)" % (what, z.mean(), z.size)
    assert 0.6 <= z.std() <= 1.4, "%s: spread differs from the reference's: std z = %.3f" % (what, z.std())
    assert (np.abs(z) <= 3.5).mean() >= 0.985, "%s: %.2f %% of the penumbra pixels beyond 3.5 sigma" % (what, 100 * (np.abs(z) > 3.5).mean())
    return z


def test_oracle_soft_shadows_match_the_reference_statistically(golden):
    g = golden("cpuonly_area.npz")
    zs = []
    for seed in (0, 1, 12345):
        sc, fr = scenes.cpuonly_area_case(g, outputs=ALL, rng_seed=seed)
        o = orclib.oracle_render(sc, fr)
        zs.append(against_reference_statistics(o["rgb"].astype(np.float64), g, "oracle, seed %d" % seed))
        assert o["counters"]["rays_shadow"] > 8 * 0.3 * fr.width * fr.height     # 8 shadow rays per lit hit
    assert not np.array_equal(zs[0], zs[1])                # the seed matters
    # radius 0 or one sample: the point-light frame, bit for bit
    sc, fr = scenes.cpuonly_area_case(g, outputs=ALL)
    from raytracinginonesemester_b200 import api
    point = orclib.oracle_render(sc, api.Frame(fr.cam, fr.width, fr.height, mode=fr.mode, accel=fr.accel, lights=fr.lights, spp=1, jitter=fr.jitter,
                                               max_depth=fr.max_depth, outputs=ALL, quantiser=fr.quantiser))
    fr.light_radius[:] = 0.0
    zero = orclib.oracle_render(sc, fr)
    for k in ("rgb", "tri_id", "t"):
        assert np.array_equal(point[k], zero[k]), k


def test_device_code_on_host_soft_shadows_equal_the_oracle(golden):
    """The product's per-element device functions (compiled for the host, tests/emul) vs the oracle: same hash RNG, same frame."""
    g = golden("cpuonly_area.npz")
    sc, fr = scenes.cpuonly_area_case(g, outputs=ALL, rng_seed=7)
    h = orclib.emul_build(sc, 2)
    e = orclib.emul_render(h, fr)
    o = orclib.oracle_render(sc, fr)
    orclib.emul().emu_free(h)
    assert np.array_equal(e["tri_id"], o["tri_id"]) and np.array_equal(e["t"], o["t"])
    assert np.array_equal(e["rgb"], o["rgb"])
    assert e["stats"]["rays_shadow"] == o["counters"]["rays_shadow"]


@pytest.mark.gpu
def test_device_soft_shadows(renderer, golden):
    g = golden("cpuonly_area.npz")
    for seed in (0, 99):
        sc, fr = scenes.cpuonly_area_case(g, outputs=ALL, rng_seed=seed)
        renderer.upload_scene(sc)
        renderer.render(fr)
        a = renderer.download()
        o = orclib.oracle_render(sc, fr)
        assert np.array_equal(a["tri_id"], o["tri_id"]) and np.array_equal(a["t"], o["t"])
        assert np.abs(a["rgb"] - o["rgb"]).max() <= 2e-6          # powf: fp64 pow narrowed vs glibc powf
        assert np.abs(a["rgb8"].astype(int) - o["rgb8"].astype(int)).max() <= 1
        assert a["rays_shadow"] == o["counters"]["rays_shadow"] and a["rays_primary"] == o["counters"]["rays_primary"]
        against_reference_statistics(a["rgb"].astype(np.float64), g, "device, seed %d" % seed)
    # disk lights belong to RT_MODE_HW2_CPU
    from raytracinginonesemester_b200 import api
    sc, fr = scenes.cpuonly_area_case(g, outputs=ALL)
    fr.mode = A.RT_MODE_HW2_BVH
    fr.lights = [api.make_light((-1, -1, 1), (1, 1, 1), 5)]
    fr._light_arr = (A.rt_light * 1)(*fr.lights)
    with pytest.raises(api.RtError):
        renderer.render(fr)
