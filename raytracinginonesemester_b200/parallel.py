"""One process per GPU: torch.distributed is the plumbing (rendezvous, barriers, max-over-ranks
timing, handing the NCCL unique id to every rank); the data path — scene/BVH broadcast from rank 0
and the per-frame tile gather to rank 0 — runs inside the C library on its own NCCL communicator
(rt_comm_init, rt_upload_scene, rt_render).  Screen-space tiles (16x8 px) are owned in contiguous chunks
(horizontal bands, several per rank): chunk c of the row-major tile order belongs to rank c % world
(csrc/rt_params.h, rt_global_tile)."""
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


DEFAULT_CHUNKS_PER_RANK = 4


def chunk_tiles(width, height, world, chunks_per_rank=0, tile_w=16, tile_h=8):
    """Tiles per ownership chunk (csrc/rt_params.h, rt_chunk_tiles)."""
    total = ((width + tile_w - 1) // tile_w) * ((height + tile_h - 1) // tile_h)
    c = world * (chunks_per_rank or DEFAULT_CHUNKS_PER_RANK)
    return max(1, (total + c - 1) // c)


def tiles_of_rank(width, height, rank, world, chunks_per_rank=0, tile_w=16, tile_h=8):
    """Local tile slots of a rank (padding included; the same for every rank)."""
    if world <= 1:
        return ((width + tile_w - 1) // tile_w) * ((height + tile_h - 1) // tile_h)
    return (chunks_per_rank or DEFAULT_CHUNKS_PER_RANK) * chunk_tiles(width, height, world, chunks_per_rank, tile_w, tile_h)


def tile_owner_map(width, height, world, chunks_per_rank=0, tile_w=16, tile_h=8):
    """(H, W) array: which rank renders each pixel (contiguous chunks of row-major tiles, chunk c -> rank c % world)."""
    tx = (width + tile_w - 1) // tile_w
    ys, xs = np.mgrid[0:height, 0:width]
    g = (ys // tile_h) * tx + (xs // tile_w)
    return (g // chunk_tiles(width, height, world, chunks_per_rank, tile_w, tile_h)) % world
