"""CPU: the multi-GPU sharding and gather logic with world_size 2 over gloo (no GPU needed): tile ownership is a
partition of the image in the reference's visiting order, and gathering packed slabs reassembles the frame."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from lucille_b200 import accel, distributed


def _frame(w, h, rank=0, world=1):
    return accel.make_frame(np.eye(4).reshape(16), 2.0, True, w, h, 2, 2, 16, rng_mode=1, rank=rank, world=world)


@pytest.mark.parametrize("w,h,world", [(640, 480, 1), (640, 480, 8), (97, 61, 3), (33, 200, 2), (31, 17, 4)])
def test_tiles_partition_the_image(w, h, world):
    seen = np.zeros((h, w), dtype=np.int32)
    for r in range(world):
        pix = accel.frame_pixels(_frame(w, h, r, world))
        seen[(pix >> 16).astype(np.int64), (pix & 0xFFFF).astype(np.int64)] += 1
    assert np.all(seen == 1)
    assert sum(distributed.shard_counts(_frame(w, h), world)) == w * h


def test_single_rank_order_is_the_reference_bucket_order(oracle):
    """world == 1: spiral buckets, row-major pixels inside a bucket -- checked against the oracle's bucket list
    (itself bit-pinned to the reference renderer through the C1 frames)."""
    w, h = 640, 480
    pix = accel.frame_pixels(_frame(w, h))
    want = []
    for bx, by, bw, bh in oracle.bucket_list(w, h, 32):
        ys, xs = np.mgrid[by:by + bh, bx:bx + bw]
        want.append((xs.ravel().astype(np.uint32) | (ys.ravel().astype(np.uint32) << 16)))
    assert np.array_equal(pix, np.concatenate(want))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, w, h, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = _frame(w, h, rank, world)
    pix = accel.frame_pixels(f)
    x, y = (pix & 0xFFFF).astype(np.float32), (pix >> 16).astype(np.float32)
    slab = torch.from_numpy(np.stack([x, y, x * 1000 + y], axis=1))         # a fake renderer: colour = f(x, y)
    out = distributed.gather_frame(slab, f, rank, world)
    if rank == 0:
        q.put(out)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_over_gloo_world2():
    w, h, world = 100, 70, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    rgb = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ys, xs = np.mgrid[0:h, 0:w]
    want = np.stack([xs, ys, xs * 1000 + ys], axis=-1).astype(np.float32)[::-1]   # row H-1-y
    assert np.array_equal(rgb, want)


def test_bucket_bases_is_the_spiral_prefix():
    """distributed.bucket_bases: with buckets dealt round-robin (b % world == rank), the base of a bucket is the number of hit
    samples in all buckets before it in spiral order -- the position of its first gather ray in the single MT19937 stream."""
    rng = np.random.default_rng(3)
    for world, nb in ((1, 7), (2, 7), (3, 300), (8, 300)):
        hits = rng.integers(0, 9217, size=nb)
        per = (nb + world - 1) // world
        padded = np.zeros((world, per), dtype=np.int64)
        for b in range(nb):
            padded[b % world, b // world] = hits[b]
        bases, total = distributed.bucket_bases(padded, world)
        want = np.cumsum(hits) - hits
        assert total == hits.sum()
        for b in range(nb):
            assert int(bases[b % world, b // world]) == int(want[b])


def _exchange_worker(rank, world, port, nb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hits = (np.arange(nb, dtype=np.uint32) * 37 + 5) % 1000            # the frame's per-bucket counts, known to the test
    mine = hits[rank::world]
    base, total = distributed.hit_exchange(rank, world)(mine)
    q.put((rank, np.asarray(base), total))
    dist.barrier()
    dist.destroy_process_group()


def test_hit_exchange_over_gloo_world2():
    """The callback a multi-rank rng_mode-0 frame calls between its eye pass and its gather pass: ranks with different bucket
    counts (7 buckets over 2 ranks) get the frame-wide prefix of their own buckets."""
    nb, world = 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((r, (b, t)) for r, b, t in (q.get(timeout=120) for _ in range(world)))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    hits = (np.arange(nb, dtype=np.int64) * 37 + 5) % 1000
    want = np.cumsum(hits) - hits
    for r in range(world):
        assert got[r][1] == hits.sum()
        assert np.array_equal(got[r][0].astype(np.int64), want[r::world])


def _failure_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lucille_b200 import accel
    out = []
    # (1) rank 1 failed before it had counts: the library's guard calls the exchange with NULL -> fn(None); rank 0 is in its real call
    try:
        distributed.hit_exchange(rank, world)(None if rank == 1 else np.arange(4, dtype=np.uint32))
        out.append("returned")
    except accel.B200Error as e:
        out.append("raised: " + str(e))
    # (2) rank 1's frame call fails after the exchange: the outcome all-reduce raises on rank 0 too, before it would enter the gather

    def call():
        if rank == 1:
            raise accel.B200Error("boom on rank 1")
        return "stats"
    try:
        distributed._frame_call_all_ranks(call, world)
        out.append("returned")
    except accel.B200Error as e:
        out.append("raised: " + str(e))
    # (3) and a frame that succeeds everywhere goes through
    out.append(distributed._frame_call_all_ranks(lambda: "stats", world))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_rank_failures_fail_the_frame_on_every_rank_gloo_world2():
    """A rank that fails around the hit-count exchange or in its frame call must not leave the others waiting in a collective
    (900 s pytest timeout): the failure travels through the exchange / the outcome all-reduce and every rank raises."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_failure_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0][0].startswith("raised: hit-count exchange") and got[1][0].startswith("raised: hit-count exchange")
    assert got[0][1] == "raised: the frame failed on another rank" and got[1][1] == "raised: boom on rank 1"
    assert got[0][2] == "stats" and got[1][2] == "stats"
