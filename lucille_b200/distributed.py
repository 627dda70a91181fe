"""Multi-GPU frame rendering: one process per GPU, scene replicated, framebuffer tiles sharded, ONE gather at frame end.

The reference's only data-parallel axis is image tiles over a shared read-only scene: 32x32 buckets popped by pthread
workers (src/render/render.c:582-710, 1043-1105), and its compiled-out MPI layer has rank 0 own the display
(src/base/parallel.c:45-232, render.c:468-514).  Here bucket ``b`` of the spiral order belongs to rank ``b % world``;
every rank renders its pixels with no inter-GPU traffic and writes them PACKED, in visiting order, straight into its NCCL
send buffer (the resolve kernel does it: no staging copy); rank 0 gathers the slabs over NVLink and scatters them into the
framebuffer.  torch.distributed is plumbing only (process group + the collective).
"""
from __future__ import annotations

import copy

import numpy as np

from . import accel as _accel


def shard_counts(frame: "_accel.Frame", world: int):
    """Pixels per rank for ``world`` ranks (host logic; every rank computes the same table)."""
    out = []
    for r in range(world):
        f = copy.copy(frame)
        f.rank, f.world = r, world
        out.append(len(_accel.frame_pixels(f)))
    return out


def scatter_tiles(width: int, height: int, pixel_lists, slabs) -> np.ndarray:
    """Assemble the framebuffer (row H-1-y, render.c:962-964) from per-rank pixel lists and packed RGB slabs."""
    rgb = np.zeros((height, width, 3), dtype=np.float32)
    for pix, slab in zip(pixel_lists, slabs):
        x = (pix & 0xFFFF).astype(np.int64)
        y = (pix >> 16).astype(np.int64)
        rgb[height - 1 - y, x] = np.asarray(slab, dtype=np.float32).reshape(-1, 3)[: len(pix)]
    return rgb


def gather_frame(local_slab, frame: "_accel.Frame", rank: int, world: int, group=None):
    """Gather the packed per-rank slabs on rank 0 and return the assembled framebuffer there (None elsewhere).

    ``local_slab``: torch tensor [npix_rank, 3] float32 (CUDA with the nccl backend, CPU with gloo)."""
    import torch
    import torch.distributed as dist

    counts = shard_counts(frame, world)
    pad = max(counts)
    send = torch.zeros((pad, 3), dtype=torch.float32, device=local_slab.device)
    send[: counts[rank]] = local_slab[: counts[rank]]
    if world == 1:
        bufs = [send]
    else:
        bufs = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
        dist.gather(send, bufs, dst=0, group=group)
    if rank != 0:
        return None
    lists = []
    for r in range(world):
        f = copy.copy(frame)
        f.rank, f.world = r, world
        lists.append(_accel.frame_pixels(f))
    return scatter_tiles(frame.width, frame.height, lists, [b.cpu().numpy() for b in bufs])


def render_ao_distributed(acc: "_accel.Accel", frame: "_accel.Frame", rank: int, world: int, group=None, stream=None):
    """Render this rank's tiles on its GPU, gather on rank 0.  Returns (framebuffer or None, FrameStats of this rank)."""
    import torch

    f = copy.copy(frame)
    f.rank, f.world = rank, world
    npix = len(_accel.frame_pixels(f))
    slab = torch.empty((max(npix, 1), 3), dtype=torch.float32, device="cuda")
    stats = acc.render_ao_tiles_dev(f, slab, stream)
    torch.cuda.synchronize()
    return gather_frame(slab[:npix], f, rank, world, group), stats
