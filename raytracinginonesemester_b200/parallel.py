"""One process per GPU: torch.distributed is the plumbing (rendezvous, barriers, max-over-ranks
timing, handing the NCCL unique id to every rank); the data path — scene/BVH broadcast from rank 0
and the per-frame tile gather to rank 0 — runs inside the C library on its own NCCL communicator
(rt_comm_init, rt_upload_scene, rt_render).  Screen-space tiles (16x8 px) are owned round-robin:
tile k of the frame belongs to rank k % world (csrc/rt_params.h, rt_map_pixel)."""
import os

import numpy as np


def env_ranks():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """Initialises torch.distributed from the torchrun environment (MASTER_ADDR/PORT, RANK, WORLD_SIZE).
    backend defaults to nccl when CUDA is available, else gloo.  Returns (dist or None, rank, world, local_rank)."""
    rank, world, local_rank = env_ranks()
    if world == 1:
        return None, rank, world, local_rank
    import torch
    import torch.distributed as dist
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        kw["device_id"] = torch.device("cuda", local_rank)
    dist.init_process_group(backend, **kw)
    return dist, rank, world, local_rank


def broadcast_bytes(dist, payload, nbytes, src=0):
    """Broadcasts a fixed-size byte string (e.g. the 128-byte NCCL unique id) from rank `src`."""
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def make_renderer(dist, rank, world, local_rank):
    """Renderer bound to this rank's GPU, with the library's own NCCL communicator when world > 1."""
    from . import api
    nccl_id = None
    if world > 1:
        nccl_id = broadcast_bytes(dist, api.Renderer.nccl_unique_id() if rank == 0 else b"", 128, 0)
    return api.Renderer(local_rank, rank, world, nccl_id)


def tiles_of_rank(width, height, rank, world, tile_w=16, tile_h=8):
    total = ((width + tile_w - 1) // tile_w) * ((height + tile_h - 1) // tile_h)
    return (total - rank + world - 1) // world if total > rank else 0


def tile_owner_map(width, height, world, tile_w=16, tile_h=8):
    """(H, W) array: which rank renders each pixel."""
    tx = (width + tile_w - 1) // tile_w
    ys, xs = np.mgrid[0:height, 0:width]
    return ((ys // tile_h) * tx + (xs // tile_w)) % world
