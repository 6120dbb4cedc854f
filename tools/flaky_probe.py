#!/usr/bin/env python
"""Repeats the frustum-vs-per-lane comparison of tests/test_gpu_parity.py and reports every mismatch (diagnostic)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracinginonesemester_b200 import _abi as A, api, scenes  # noqa: E402

ALL = A.RT_OUT_RGB_F32 | A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [A.RT_VARIANT_PACKET, A.RT_VARIANT_FRUSTUM, A.RT_VARIANT_DEFAULT]
r = api.Renderer(0)


def run(fr, v):
    fr.kernel_variant = v
    r.render(fr)
    return r.download()


def cases():
    d = np.load(os.path.join(ROOT, "tests", "golden", "frog_mesh.npz"))
    frog = api.Scene(d["positions"], d["indices"], normals=d["normals"], tri_obj_ids=d["tri_obj_ids"], materials=[api.make_material(**scenes.FROG_MATERIAL)])
    terr = scenes.terrain_scene(200, 100)
    for W, H, spp in ((640, 360, 1), (333, 187, 4), (64, 40, 16)):
        yield "terrain %dx%d spp%d" % (W, H, spp), terr, scenes.terrain_frame(W, H, spp=spp, outputs=ALL)
    fr = scenes.terrain_frame(320, 180, outputs=ALL)
    fr.cam = api.camera_init((-1.4, -0.2, 0.12), (0.5, 0.1, 0.0), (0, 0, 1), 18.0, 24.0, 320, 180)
    fr.lights = [api.make_light((-2.0, -1.0, 1.5), (1, 1, 1), 5), api.make_light((1.5, 0.8, 0.3), (1, 0.5, 0.2), 3)]
    fr._light_arr = (A.rt_light * 2)(*fr.lights)
    yield "terrain grazing 2 lights", terr, fr
    for filling in (False, True):
        yield "frog filling=%s" % filling, frog, scenes.frog_frame(320, 200, filling=filling, outputs=ALL)
    yield "terrain big 1920x1080", scenes.terrain_scene(400, 200), scenes.terrain_frame(1920, 1080, outputs=ALL)
    fr = scenes.terrain_frame(320, 180, outputs=ALL)
    fr.lights = [api.make_light((-2.0, -1.0, 1.5), (1, 1, 1), 5), api.make_light((1.5, 0.8, 0.3), (1, 0.5, 0.2), 3)]
    fr._light_arr = (A.rt_light * 2)(*fr.lights)
    fr.cam = api.camera_init((0.0, 0.0, 400.0), (0, 0, 0), (0, 1, 0), 4000.0, 24.0, 320, 180)
    yield "terrain far camera", terr, fr
    d = np.load(os.path.join(ROOT, "tests", "golden", "cornell_mesh.npz"))
    sc = api.Scene(d["positions"], d["indices"], normals=d["normals"] if d["normals"].size else None, tri_obj_ids=d["tri_obj_ids"],
                   materials=[api.make_material(albedo=(0.7, 0.6, 0.5), kd=0.9, ks=0.2) for _ in range(int(d["tri_obj_ids"].max()) + 1)])
    lo, hi = d["positions"].min(0), d["positions"].max(0)
    mid = 0.5 * (lo + hi)
    for n, (pos, look, focal) in enumerate((((mid[0], mid[1], lo[2] + 0.05 * (hi[2] - lo[2])), tuple(mid), 12.0), (tuple(mid), (hi[0], hi[1], mid[2]), 8.0),
                                            ((mid[0], mid[1], lo[2] - 1.5 * (hi[2] - lo[2])), tuple(mid), 35.0))):
        cam = api.camera_init(pos, look, (0, 1, 0), focal, 24.0, 200, 160)
        yield "cornell %d" % n, sc, api.Frame(cam, 200, 160, lights=[api.make_light((mid[0], hi[1] - 0.05 * (hi[1] - lo[1]), mid[2]), (1, 1, 1), 60000),
                                                                  api.make_light(tuple(lo - 0.3 * (hi - lo)), (0.3, 0.4, 1.0), 90000)],
                                              miss_color=(0.1, 0.1, 0.2), outputs=ALL)
    pos = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0], [0, 0, 0.5], [0, 0, 0.5], [0, 0, 0.5]], np.float32)
    for n, idx in enumerate(([[0, 1, 2]], [[0, 1, 2], [0, 2, 1]], [[0, 1, 2], [3, 4, 5]], [[0, 1, 2]] * 7 + [[3, 4, 5]] * 3)):
        cam = api.camera_init((0.1, 0.05, 3), (0, 0, 0), (0, 1, 0), 50.0, 24.0, 40, 30)
        yield "tiny %d" % n, api.Scene(pos, np.array(idx, np.uint32)), api.Frame(cam, 40, 30, lights=[api.make_light((1, 1, 2), (1, 1, 1), 3)], miss_color=(0.2, 0.3, 0.4), outputs=ALL)


bad = 0
cur = None
for name, sc, fr in cases():
    if sc is not cur or name.startswith("tiny"):
        r.upload_scene(sc)                            # (tiny scenes: re-upload every time, the BVH build is part of the probe)
        cur = sc
    base = None
    for rep in range(reps):
        if name.startswith("tiny") or name.startswith("cornell"):
            r.upload_scene(sc)                        # the BVH build is part of the probe
        for v in variants:
            got = run(fr, v)
            if base is None:
                base = got
                continue
            for k in ("tri_id", "t", "rgb", "rgb8"):
                if not np.array_equal(base[k], got[k]):
                    m = base[k] != got[k]
                    if m.ndim == 3:
                        m = m.any(-1)
                    ys, xs = np.nonzero(m)
                    bad += 1
                    print("MISMATCH %s rep %d variant %d plane %s: %d px, first (%d,%d) base %s got %s  base id %d got id %d" %
                          (name, rep, v, k, m.sum(), xs[0], ys[0], base[k][ys[0], xs[0]], got[k][ys[0], xs[0]], base["tri_id"][ys[0], xs[0]], got["tri_id"][ys[0], xs[0]]), flush=True)
print("flaky_probe: %d mismatching planes" % bad)
