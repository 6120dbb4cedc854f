#!/usr/bin/env python
"""Arena size, build time and C4 frame time of the compact 8-wide view in each phase (RT_B200_WIDE_PHASE; rt_build_wide)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401
from raytracinginonesemester_b200 import Renderer, scenes  # noqa: E402

for ntri, nx, ny in ((1_000_000, 1000, 500), (10_000_000, 2500, 2000)):
    sc = scenes.terrain_scene(nx, ny)
    for phase in ("", "0", "1", "2", "-1"):
        os.environ["RT_B200_WIDE_PHASE"] = phase
        r = Renderer(0)
        r.upload_scene(sc)
        info = r.upload_scene(sc)          # second build: steady-state time
        line = "%d tris phase %-4s nodes %d arena MB %.3f build ms %.2f" % (sc.indices.shape[0], phase or "auto", info.num_nodes, info.arena_bytes / 1e6, info.build_ms)
        if ntri == 1_000_000:
            fr = scenes.terrain_frame(3840, 2160)
            for _ in range(3):
                r.render(fr); r.sync()
            ms = []
            for _ in range(10):
                r.render(fr)
                ms.append(r.sync())
            line += "  frame ms %.4f" % (sum(ms) / len(ms))
        print(line, flush=True)
        r.close()
