#!/usr/bin/env python
"""Wall-clock vs device time of the two end-to-end paths on the C4 frame: rt_render + rt_download_image and rt_render_into."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raytracinginonesemester_b200 import _abi as A, api, scenes
r = api.Renderer(0)
r.upload_scene(scenes.terrain_scene(1000, 500))
W, H = 3840, 2160
fr = scenes.terrain_frame(W, H, outputs=A.RT_OUT_RGB8)
pinned = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True).numpy()
pageable = np.empty((H, W, 3), np.uint8)
for name, buf in (("pinned", pinned), ("pageable", pageable)):
    for mode in ("render+download", "render_into"):
        wall, dev = [], []
        for i in range(25):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if mode == "render_into":
                o = r.render_into(fr, into={"rgb8": buf})
            else:
                r.render(fr); o = r.download(into={"rgb8": buf})
            wall.append(time.perf_counter() - t0); dev.append(o["gpu_ms"])
        print("%-9s %-16s wall %.3f ms  device(ev0..ev1) %.3f ms  kernel-only %.3f ms" % (name, mode, 1e3 * np.median(wall[5:]), np.median(dev[5:]), r.frame_times()[1]))
r.close()
