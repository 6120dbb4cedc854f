#!/usr/bin/env python
"""torchrun worker: wall-clock of rt_render + rt_download_image vs rt_render_into on the C4 frame, N GPUs."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raytracinginonesemester_b200 import _abi as A, api, parallel, scenes
dist, rank, world, local_rank = parallel.init_process_group("nccl")
r = parallel.make_renderer(dist, rank, world, local_rank)
r.upload_scene(scenes.terrain_scene(1000, 500) if rank == 0 else None)
W, H = 3840, 2160
fr = scenes.terrain_frame(W, H, outputs=A.RT_OUT_RGB8)
pinned = torch.empty((H, W, 3), dtype=torch.uint8, pin_memory=True).numpy() if rank == 0 else None
for gm in (A.RT_GATHER_PEER, A.RT_GATHER_NCCL):
    r.set_gather(gm)
    for mode in ("render+download", "render_into", "render+sync"):
        wall, dev, kern = [], [], []
        for i in range(25):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            into = {"rgb8": pinned} if rank == 0 else None
            if mode == "render_into":
                o = r.render_into(fr, into=into)
            elif mode == "render+download":
                r.render(fr); o = r.download(into=into)
            else:
                r.render(fr); o = {"gpu_ms": r.sync()}
            wall.append(time.perf_counter() - t0); dev.append(o["gpu_ms"]); kern.append(r.frame_times()[1])
        print("rank %d gather %d %-16s wall %.3f ms  device(ev0..ev1) %.3f ms  kernel-only %.3f ms" % (rank, gm, mode, 1e3 * np.median(wall[5:]), np.median(dev[5:]), np.median(kern[5:])), flush=True)
dist.barrier()
r.close()
dist.destroy_process_group()
