"""Tile-sharded rendering on every visible GPU of the box (needs >= 2): one process per GPU under
torch.distributed.run, NCCL broadcast of the scene arena, NCCL tile gather to rank 0."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_sharded_frame_equals_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (covered on the CPU by tests/test_host_gloo.py)")
    n = min(n, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_single_process_group_equals_single_gpu():
    """rt_create_multi: ONE process and ONE host thread over every GPU of the box (the layout an unmodified single-process main()
    of the reference can adopt): scene copied to the other GPUs over NVLink, fused peer-store gather, rt_render_into with every
    GPU copying its own bands into the caller's buffers — all planes equal a single-GPU render bit for bit."""
    import numpy as np
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 8)
    sys.path.insert(0, ROOT)
    from raytracinginonesemester_b200 import _abi as A, api, scenes
    ALL = A.RT_OUT_RGB_F32 | A.RT_OUT_RGB8 | A.RT_OUT_TRI_ID | A.RT_OUT_T
    group = api.Renderer(devices=list(range(n)))
    solo = api.Renderer(0)
    for sc in (scenes.terrain_scene(120, 60), scenes.terrain_scene(300, 150)):
        info = group.upload_scene(sc)
        solo.upload_scene(sc)
        assert info.num_triangles == sc.indices.shape[0]
        for cpr, (W, H), spp, outs in ((0, (333, 201), 1, ALL), (3, (640, 360), 2, ALL), (16, (517, 301), 1, A.RT_OUT_RGB8 | A.RT_OUT_T), (1, (40, 9), 1, ALL), (0, (1920, 1080), 1, A.RT_OUT_RGB8)):
            group.set_sharding(cpr)
            fr = scenes.terrain_frame(W, H, spp=spp, outputs=outs)
            solo.render(fr)
            ref = solo.download()
            for rep in range(2):
                group.render(fr)
                got = group.download()
                for k in ("rgb", "rgb8", "tri_id", "t"):
                    if k in ref:
                        assert np.array_equal(got[k], ref[k]), (cpr, W, H, k, "render+download")
                assert got["rays_primary"] == W * H * spp and got["rays_shadow"] == ref["rays_shadow"]
                into = group.render_into(fr)
                for k in ("rgb", "rgb8", "tri_id", "t"):
                    if k in ref:
                        assert np.array_equal(into[k], ref[k]), (cpr, W, H, k, "render_into")
                assert into["rays_primary"] == W * H * spp and into["rays_shadow"] == ref["rays_shadow"]
    group.close()
    solo.close()
