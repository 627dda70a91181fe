"""Multi-GPU frame rendering: one process per GPU, scene replicated, framebuffer tiles sharded, ONE gather at frame end.

The reference's only data-parallel axis is image tiles over a shared read-only scene: 32x32 buckets popped by pthread
workers (src/render/render.c:582-710, 1043-1105), and its compiled-out MPI layer has rank 0 own the display
(src/base/parallel.c:45-232, render.c:468-514).  Here bucket ``b`` of the spiral order belongs to rank ``b % world``;
every rank renders its pixels with no inter-GPU traffic and writes them PACKED, in visiting order, straight into its NCCL
send buffer (the resolve kernel does it: no staging copy); rank 0 gathers the slabs over NVLink and scatters them into the
framebuffer -- or, fused (PeerFramebuffer / render_ao_distributed_peer), every rank's resolve kernel stores its tiles straight into
rank 0's framebuffer over NVLink peer memory and no data collective runs at all.  torch.distributed is plumbing only (process
group, the 64-byte IPC handle, barriers, and the per-bucket hit counts of the shared MT19937 stream mode).
"""
from __future__ import annotations

import copy

import numpy as np

from . import accel as _accel


def shard_counts(frame: "_accel.Frame", world: int):
    """Pixels per rank for ``world`` ranks (host logic; every rank computes the same table; cached per frame geometry)."""
    key = (frame.width, frame.height, frame.bucket_size, world)
    if key not in _COUNTS_CACHE:
        out = []
        for r in range(world):
            f = copy.copy(frame)
            f.rank, f.world = r, world
            out.append(len(_accel.frame_pixels(f)))
        _COUNTS_CACHE.clear()
        _COUNTS_CACHE[key] = out
    return list(_COUNTS_CACHE[key])


_COUNTS_CACHE = {}


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
    if bufs[0].is_cuda:
        # assemble on rank 0's GPU (one indexed store per slab, plumbing), then ONE device-to-host copy of the framebuffer
        key = (frame.width, frame.height, frame.bucket_size, world)
        idx = _SCATTER_CACHE.get(key)
        if idx is None:
            idx = []
            for r in range(world):
                f = copy.copy(frame)
                f.rank, f.world = r, world
                pix = _accel.frame_pixels(f).astype(np.int64)
                lin = (frame.height - 1 - (pix >> 16)) * frame.width + (pix & 0xFFFF)          # row H-1-y, render.c:962-964
                idx.append(torch.from_numpy(lin).to(bufs[0].device))
            _SCATTER_CACHE.clear()
            _SCATTER_CACHE[key] = idx
        fbuf = torch.zeros((frame.height * frame.width, 3), dtype=torch.float32, device=bufs[0].device)
        for r in range(world):
            fbuf[idx[r]] = bufs[r][: counts[r]]
        host = _PINNED.get(key)
        if host is None or host.shape != fbuf.shape:
            _PINNED.clear()
            host = _PINNED[key] = torch.empty(fbuf.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(fbuf, non_blocking=True)
        torch.cuda.synchronize()
        return host.numpy().reshape(frame.height, frame.width, 3)          # a view of a cached pinned buffer: valid until the next gather
    lists = []
    for r in range(world):
        f = copy.copy(frame)
        f.rank, f.world = r, world
        lists.append(_accel.frame_pixels(f))
    return scatter_tiles(frame.width, frame.height, lists, [b.cpu().numpy() for b in bufs])


_SCATTER_CACHE = {}
_PINNED = {}


def bucket_bases(all_hits, world: int):
    """all_hits[r][i] = hit samples of bucket r + i*world (rank r's i-th bucket; ranks with one bucket fewer are padded with 0).
    Returns bases[r][i] = hit samples in all buckets that precede bucket r + i*world in spiral order, and the frame's total."""
    all_hits = np.asarray(all_hits, dtype=np.int64)              # [world][per_rank]
    order = all_hits.T.reshape(-1)                               # bucket b = i*world + r
    ex = np.cumsum(order) - order
    return ex.reshape(-1, world).T.astype(np.uint64), int(order.sum())


def hit_exchange(rank: int, world: int, group=None, device=None):
    """The ri_b200_set_hit_exchange callback over torch.distributed: one all-gather of the per-bucket hit counts."""
    import torch
    import torch.distributed as dist

    def fn(hits):
        # hits is None when this rank failed before it had counts (the library's guard call): it still takes part, with the failure
        # flag raised, and every rank leaves the exchange with an error instead of waiting for it
        failed = hits is None
        head = torch.tensor([0 if failed else len(hits), 1 if failed else 0], dtype=torch.int64, device=device)
        dist.all_reduce(head, op=dist.ReduceOp.MAX, group=group)
        if int(head[1].item()):
            raise _accel.B200Error("hit-count exchange: a rank of the frame failed before the exchange")
        mine = torch.zeros(int(head[0].item()), dtype=torch.int64, device=device)
        mine[:len(hits)] = torch.from_numpy(hits.astype(np.int64))
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        bases, total = bucket_bases(torch.stack(parts).cpu().numpy(), world)
        return bases[rank][:len(hits)], total
    return fn


def _frame_call_all_ranks(call, world: int, group=None, device=None):
    """Run this rank's frame call; with several ranks, agree on the outcome before anybody enters the next collective: one all-reduce
    of a failure flag (which also is the "every rank's stores have landed" barrier of the peer path).  A failure on any rank raises
    on every rank -- nobody is left waiting in a gather or a barrier for a rank that has gone."""
    import torch
    import torch.distributed as dist

    err = None
    try:
        out = call()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
    except Exception as e:                         # noqa: BLE001 -- re-raised below, after the other ranks know
        out, err = None, e
    if world > 1:
        flag = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if err is None and int(flag.item()):
            raise _accel.B200Error("the frame failed on another rank")
    if err is not None:
        raise err
    return out


class PeerFramebuffer:
    """One framebuffer on rank 0's GPU, mapped by every rank of the node (CUDA IPC, NVLink / NVSwitch peer memory).  Made once and
    reused frame after frame: mapping a handle costs far more than a frame.  torch.distributed carries the 64-byte handle."""

    def __init__(self, width: int, height: int, rank: int, world: int, device: int, group=None):
        import torch.distributed as dist

        self.shape, self.rank, self.world, self.device, self.group = (height, width, 3), rank, world, device, group
        box = [None]
        if rank == 0:
            self.ptr, box[0] = _accel.peer_alloc(width * height * 3 * 4, device)
        if world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        if rank != 0:
            self.ptr = _accel.peer_open(box[0], device)
        # rank 0 reads every frame into ONE pinned host buffer it keeps (a fresh pageable array per frame costs more than the kernels
        # of a small frame: 201 MB at 4096^2 is ~40 ms of page faults and staged copies against ~8 ms into pinned memory)
        self._host_ptr, self._host = None, None
        if rank == 0:
            nbytes = width * height * 3 * 4
            self._host_ptr = _accel.load_library().ri_b200_host_alloc(nbytes)
            if self._host_ptr:
                import ctypes
                self._host = np.ctypeslib.as_array(ctypes.cast(self._host_ptr, ctypes.POINTER(ctypes.c_float)), shape=(height, width, 3))

    def read(self) -> np.ndarray:
        """Rank 0: the frame, as a view of this object's pinned host buffer -- valid until the next read() or close(); copy it to keep it."""
        if self._host is None:
            return _accel.peer_read(self.ptr, self.shape, self.device)
        _accel._check(_accel.load_library().ri_b200_peer_read(_accel.C.c_void_p(self.ptr), _accel.C.c_void_p(self._host_ptr), self._host.nbytes,
                                                               self.device))
        return self._host

    def close(self):
        import torch.distributed as dist

        if self.ptr is None:
            return
        if self.world > 1:
            dist.barrier(group=self.group)             # nobody frees or unmaps while another rank may still store
        (_accel.peer_free if self.rank == 0 else _accel.peer_close)(self.ptr, self.device)
        self.ptr = None
        if self._host_ptr:
            self._host = None
            _accel.load_library().ri_b200_host_free(_accel.C.c_void_p(self._host_ptr))
            self._host_ptr = None


def render_ao_distributed_peer(acc: "_accel.Accel", frame: "_accel.Frame", fb: PeerFramebuffer, stream=None):
    """The same frame with the gather FUSED into the resolve kernels: every rank's resolve kernel stores its tiles straight into
    rank 0's framebuffer (ri_b200_render_ao_peer_dev); one barrier says the stores have landed, then rank 0 reads the frame.
    Returns (framebuffer on rank 0 or None, FrameStats of this rank); the framebuffer is a VIEW of fb's pinned host buffer -- copy it
    if it must outlive the next frame rendered into fb or fb.close()."""
    import torch
    import torch.distributed as dist

    assert (frame.height, frame.width, 3) == fb.shape
    f = copy.copy(frame)
    f.rank, f.world = fb.rank, fb.world
    if f.rng_mode == 0 and fb.world > 1:
        acc.set_hit_exchange(hit_exchange(fb.rank, fb.world, fb.group, device="cuda"))
    if fb.world > 1:
        dist.barrier(group=fb.group)                   # rank 0 has read the previous frame out of the buffer
    stats = _frame_call_all_ranks(lambda: acc.render_ao_peer_dev(f, fb.ptr, stream), fb.world, fb.group, device="cuda")
    rgb = fb.read() if fb.rank == 0 else None          # a view of fb's pinned buffer: valid until the next frame / fb.close()
    return rgb, stats


def render_ao_distributed(acc: "_accel.Accel", frame: "_accel.Frame", rank: int, world: int, group=None, stream=None):
    """Render this rank's tiles on its GPU, gather on rank 0.  Returns (framebuffer or None, FrameStats of this rank)."""
    import torch

    f = copy.copy(frame)
    f.rank, f.world = rank, world
    if f.rng_mode == 0 and world > 1:              # the reference's single MT19937 stream, shared by the ranks (bit-exact frames)
        acc.set_hit_exchange(hit_exchange(rank, world, group, device="cuda"))
    npix = len(_accel.frame_pixels(f))
    slab = torch.empty((max(npix, 1), 3), dtype=torch.float32, device="cuda")
    stats = _frame_call_all_ranks(lambda: acc.render_ao_tiles_dev(f, slab, stream), world, group, device="cuda")
    return gather_frame(slab[:npix], f, rank, world, group), stats
