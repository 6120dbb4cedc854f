"""Worker for tests/test_multi_gpu.py (launched with torch.distributed.run, one rank per GPU):
renders tile-sharded frames through the C ABI (scene + BVH broadcast from rank 0 over NCCL, tile
gather to rank 0) and compares the gathered image with a single-GPU render of the same frame."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracinginonesemester_b200 import _abi as A, api, parallel, scenes  # noqa: E402

ALL = A.RT_OUT_RGB_F32 | A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T


def main():
    dist, rank, world, local_rank = parallel.init_process_group("nccl")
    r = parallel.make_renderer(dist, rank, world, local_rank)
    sc = scenes.terrain_scene(120, 60)
    info = r.upload_scene(sc if rank == 0 else None)          # ranks != 0 receive the arena by broadcast
    assert info.num_triangles == 120 * 60 * 2 and info.num_nodes > 0
    ok = True
    want_mode = {"auto": A.RT_GATHER_AUTO, "nccl": A.RT_GATHER_NCCL, "peer": A.RT_GATHER_PEER}
    cases = [("auto", 333, 201, 1, ALL), ("nccl", 333, 201, 1, ALL), ("nccl", 256, 128, 2, ALL), ("nccl", 40, 9, 1, ALL),
             ("peer", 333, 201, 1, ALL), ("peer", 64, 32, 1, A.RT_OUT_RGB8), ("peer", 256, 128, 2, ALL), ("peer", 40, 9, 1, ALL),
             ("peer", 700, 400, 1, ALL), ("nccl", 700, 400, 1, ALL), ("peer", 333, 201, 1, ALL)]
    for n, (gm, W, H, spp, outs) in enumerate(cases):
        cpr = (0, 1, 3, 64)[n % 4]
        r.set_sharding(cpr)
        mode = r.set_gather(want_mode[gm])
        assert gm == "auto" or mode == want_mode[gm], (gm, mode)
        if rank == 0:
            print("gather", gm, "->", {A.RT_GATHER_NCCL: "nccl", A.RT_GATHER_PEER: "peer"}[mode], W, H, spp, flush=True)
        fr = scenes.terrain_frame(W, H, spp=spp, outputs=outs)
        r.render(fr)
        got = r.download()
        import torch
        cnt = torch.tensor([float(got["rays_primary"]), float(got["rays_shadow"])], device="cuda")
        dist.all_reduce(cnt)
        assert int(cnt[0].item()) == W * H * spp
        assert got["rays_primary"] == int((parallel.tile_owner_map(W, H, world, cpr) == rank).sum()) * spp
        if rank == 0:
            solo = api.Renderer(local_rank)
            solo.upload_scene(sc)
            solo.render(fr)
            ref = solo.download()
            solo.close()
            for k in ("tri_id", "t", "rgb", "rgb8"):
                if k not in got or got[k] is None:
                    continue
                if not np.array_equal(got[k], ref[k]):
                    print("MISMATCH", W, H, spp, k, int((got[k] != ref[k]).sum()))
                    ok = False
            assert int(cnt[1].item()) == ref["rays_shadow"]
    # rt_render_into: chunk-pipelined delivery + download (peer gather), plain gather + download (NCCL gather)
    for gm in ("peer", "nccl"):
        for cpr, (W, H), outs in ((0, (517, 301), ALL), (3, (640, 360), A.RT_OUT_RGB8), (16, (333, 201), ALL), (64, (333, 201), ALL), (1, (40, 9), ALL)):
            r.set_sharding(cpr)
            r.set_gather(want_mode[gm])
            fr = scenes.terrain_frame(W, H, outputs=outs)
            for rep in range(3):
                got = r.render_into(fr)
            if rank == 0:
                solo = api.Renderer(local_rank)
                solo.upload_scene(sc)
                ref = solo.render_into(fr)
                solo.close()
                for k in ("tri_id", "t", "rgb", "rgb8"):
                    if k in ref and not np.array_equal(got[k], ref[k]):
                        print("MISMATCH render_into", gm, cpr, W, H, k, int((got[k] != ref[k]).sum()))
                        ok = False
    # shared host image: every rank copies its own bands to one host buffer (rt_host_image_create + rt_render_into on every rank)
    for cpr, (W, H), outs in ((0, (517, 301), ALL), (3, (640, 360), A.RT_OUT_RGB8), (16, (333, 201), ALL), (1, (40, 9), ALL), (0, (1920, 1080), A.RT_OUT_RGB8)):
        r.set_sharding(cpr)
        r.set_gather(A.RT_GATHER_AUTO)
        fr = scenes.terrain_frame(W, H, outputs=outs)
        n = W * H
        buf = r.host_image(23 * n + 64)
        into, off = {}, 0
        for name, bit, bpp, dt, shape in (("rgb", A.RT_OUT_RGB_F32, 12, np.float32, (H, W, 3)), ("tri_id", A.RT_OUT_TRI_ID, 4, np.int32, (H, W)),
                                          ("t", A.RT_OUT_T, 4, np.float32, (H, W)), ("rgb8", A.RT_OUT_RGB8, 3, np.uint8, (H, W, 3))):
            if outs & bit:
                into[name] = buf[off:off + bpp * n].view(dt).reshape(shape)
                off += bpp * n
        for rep in range(3):
            if rank == 0:
                for a in into.values():
                    a[...] = 0
            dist.barrier()
            got = r.render_into(fr, into=into)
        if rank == 0:
            solo = api.Renderer(local_rank)
            solo.upload_scene(sc)
            ref = solo.render_into(fr)
            solo.close()
            for k in into:
                if not np.array_equal(into[k], ref[k]):
                    print("MISMATCH shared host image", cpr, W, H, k, int((into[k] != ref[k]).sum()))
                    ok = False
        dist.barrier()
    # back-to-back frames without a download in between (what bench.py's timed loop does), then one download
    r.set_sharding(0)
    r.set_gather(A.RT_GATHER_AUTO)
    fr = scenes.terrain_frame(512, 256, outputs=A.RT_OUT_RGB8)
    for _ in range(5):
        r.render(fr)
        r.sync()
    got = r.download()
    if rank == 0:
        solo = api.Renderer(local_rank)
        solo.upload_scene(sc)
        solo.render(fr)
        if not np.array_equal(got["rgb8"], solo.download()["rgb8"]):
            print("MISMATCH back-to-back")
            ok = False
        solo.close()
    dist.barrier()
    r.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK" if ok else "MULTI_GPU_FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
