#!/usr/bin/env python
"""Per-rank load balance and kernel tail of the tile sharding, measured on ONE GPU: renders each rank's share of
the C4 frame separately (rt_debug_set_shard) and prints the kernel times next to frame/world.

  python tools/shard_probe.py [--world 8] [--chunks 0,1,16,64] [--workload c4]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from raytracinginonesemester_b200 import _abi as A, api, scenes  # noqa: E402

WORK = {"c4": (1000, 500, 3840, 2160, 1), "c5": (2500, 2000, 7680, 4320, 16)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", default="8")
    ap.add_argument("--chunks", default="0")
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-flush", action="store_true", help="leave L2 warm between launches")
    ap.add_argument("--no-shadows", action="store_true")
    ap.add_argument("--sizes", default="", help="also time whole frames at WxH,WxH,... (tail/ramp cost vs pixel count)")
    a = ap.parse_args()
    nx, ny, W, H, spp = WORK[a.workload]
    r = api.Renderer(0)
    r.upload_scene(scenes.terrain_scene(nx, ny))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def t(frame, n=5):
        out = []
        for _ in range(n):
            if not a.no_flush:
                flush.zero_()
            torch.cuda.synchronize()
            r.render(frame)
            out.append(r.frame_times()[1])
        return float(np.median(out))

    fr = scenes.terrain_frame(W, H, spp=spp, outputs=A.RT_OUT_RGB8, kernel_variant=a.variant, shadows=not a.no_shadows)
    r.lib.rt_debug_set_shard(r.ctx, 0, 0)
    t(fr, 3)
    full = t(fr)
    res = {"workload": a.workload, "full_frame_kernel_ms": full, "shards": []}
    for world in [int(x) for x in a.world.split(",")]:
        for cpr in [int(x) for x in a.chunks.split(",")]:
            r.set_sharding(cpr)
            ks = []
            for rank in range(world):
                r.lib.rt_debug_set_shard(r.ctx, rank, world)
                ks.append(t(fr))
            res["shards"].append({"world": world, "chunks_per_rank": cpr, "kernel_ms": ks, "max": max(ks), "mean": float(np.mean(ks)),
                                  "ideal": full / world, "max_over_ideal": max(ks) / (full / world)})
    r.lib.rt_debug_set_shard(r.ctx, 0, 0)
    for wh in [x for x in a.sizes.split(",") if x]:
        w, h = [int(v) for v in wh.split("x")]
        f2 = scenes.terrain_frame(w, h, spp=spp, outputs=A.RT_OUT_RGB8, kernel_variant=a.variant)
        k = t(f2)
        res.setdefault("sizes", []).append({"size": wh, "kernel_ms": k, "ns_per_pixel": 1e6 * k / (w * h)})
    print(json.dumps(res))
    r.close()


if __name__ == "__main__":
    main()
