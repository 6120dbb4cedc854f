"""N > 1 host path on the CPU: two gloo processes exchange the communicator id, render their own
tiles (host emulation of the device functions, tile-packed exactly like the kernels), gather the
packed planes to rank 0 and unpack — the result must equal the single-rank frame."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, H, out_path):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import orclib
    from raytracinginonesemester_b200 import _abi as A, parallel, scenes
    dist, r, w, _ = parallel.init_process_group("gloo")
    assert (r, w) == (rank, world)
    ident = parallel.broadcast_bytes(dist, bytes(range(128)) if rank == 0 else b"", 128, 0)
    assert ident == bytes(range(128))                       # every rank holds rank 0's 128-byte id
    sc = scenes.terrain_scene(24, 12)
    h = orclib.emul_build(sc, 2)
    lib = orclib.emul()
    lib.emu_render_rank.argtypes = [C.c_void_p, C.POINTER(A.rt_frame), C.POINTER(A.rt_image), C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_int]
    lib.emu_unpack_rgb8.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, A.u8p, A.u8p]
    fr = scenes.terrain_frame(W, H, outputs=A.RT_OUT_RGB8)
    ntl = parallel.tiles_of_rank(W, H, rank, world)
    cap = max(parallel.tiles_of_rank(W, H, q, world) for q in range(world))
    packed = np.zeros((cap * 128, 3), np.uint8)
    im = A.rt_image(); im.rgb8 = packed.ctypes.data_as(A.u8p)
    f = fr.c_struct()
    assert lib.emu_render_rank(h, C.byref(f), C.byref(im), None, rank, world, 0) == ntl
    owner = parallel.tile_owner_map(W, H, world)
    assert im.rays_primary == int((owner == rank).sum())    # each rank traced exactly the pixels it owns
    mine = torch.from_numpy(packed)
    bufs = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, bufs, dst=0)
    tot = torch.tensor([float(im.rays_primary)])
    dist.all_reduce(tot)
    assert int(tot.item()) == W * H
    if rank == 0:
        img = np.full((H, W, 3), 7, np.uint8)
        for q in range(world):
            b = np.ascontiguousarray(bufs[q].numpy())
            lib.emu_unpack_rgb8(W, H, world, 0, q, b.ctypes.data_as(A.u8p), img.ctypes.data_as(A.u8p))
        np.save(out_path, img)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("W,H", [(70, 37), (64, 32)])
def test_two_rank_tile_sharding_equals_single_rank(tmp_path, W, H):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orclib
    from raytracinginonesemester_b200 import _abi as A, scenes
    out = str(tmp_path / "gathered.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, W, H, out), nprocs=2, join=True)
    got = np.load(out)
    sc = scenes.terrain_scene(24, 12)
    h = orclib.emul_build(sc, 2)
    full = orclib.emul_render(h, scenes.terrain_frame(W, H, outputs=A.RT_OUT_RGB8), want=("rgb8",))["rgb8"]
    assert np.array_equal(got, full)
