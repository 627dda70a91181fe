#!/usr/bin/env python
"""K6 experiment: what does ray ORDER buy the pooled occlusion kernel on the C3 batch?

The batch is permuted on the host (so the sort itself costs nothing here) and the resident-ray kernel is timed per order:
  batch      : the bench order -- point-major, 64 hemisphere rays per point
  oct_point  : inside every point, rays grouped by direction octant
  oct_global : octant-major, then point (a chunk of 128 rays = one octant, ~10 neighbouring points)
  morton_oct : 3-D Morton cell of the origin (2^-6 grid), then octant, then point
  random     : a random permutation (the incoherent floor)
An upper bound on what an on-device sort could win: if no order beats `batch` by more than the ~1 ms a 16 Mi-key radix sort costs,
K6 does not pay on this workload.  Prints one line per order; results equal the unsorted ones after un-permuting (checked)."""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from lucille_b200 import accel, scenes  # noqa: E402


def part1by2(x):
    x = x.astype(np.uint64) & np.uint64(0x3FF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x30000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x300F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x9249249)
    return x


def main():
    npoints = int(os.environ.get("POINTS", bench.NPOINTS))
    ntris = int(os.environ.get("NTRIS", bench.NTRIS))                # NTRIS=10000000 SEED=c5: the beyond-L2 scene of configs[4]
    seed = scenes.SEED_C5 if os.environ.get("SEED", "c3") == "c5" else scenes.SEED_C3
    tris = scenes.triangle_soup(ntris, seed)
    a = accel.Accel.bind(accel.RI_ACCEL_B200).build(tris, accel.PREC_F32)
    P, n = bench.primary_points(a.intersect, tris[a.triorder()])
    rays = scenes.ao_rays(P[:npoints], n[:npoints], bench.NTHETA, bench.NPHI, seed)
    nr = len(rays)
    octant = ((rays[:, 4] < 0).astype(np.uint64) | ((rays[:, 5] < 0).astype(np.uint64) << np.uint64(1)) |
              ((rays[:, 6] < 0).astype(np.uint64) << np.uint64(2)))
    point = (np.arange(nr, dtype=np.uint64) // np.uint64(64))
    cell = np.clip((rays[:, 0:3] * 64.0).astype(np.int64), 0, 63)
    morton = part1by2(cell[:, 0]) | (part1by2(cell[:, 1]) << np.uint64(1)) | (part1by2(cell[:, 2]) << np.uint64(2))
    rng = np.random.default_rng(1)
    orders = {
        "batch": np.arange(nr),
        "oct_point": np.argsort(point * np.uint64(8) + octant, kind="stable"),
        "oct_global": np.argsort(octant * np.uint64(1 << 32) + point, kind="stable"),
        "oct_8pts": np.argsort((point // np.uint64(8)) * np.uint64(8) + octant, kind="stable"),     # 8 neighbouring points' rays grouped by octant
        "oct_32pts": np.argsort((point // np.uint64(32)) * np.uint64(8) + octant, kind="stable"),
        "morton_oct": np.argsort((morton << np.uint64(35)) | (octant << np.uint64(32)) | point, kind="stable"),
        "random": rng.permutation(nr),
    }
    if os.environ.get("ORDERS"):
        orders = {k: v for k, v in orders.items() if k in os.environ["ORDERS"].split(",")}
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    d_occ = torch.empty((nr,), dtype=torch.uint8, device="cuda")
    ref = None
    for name, perm in orders.items():
        d_rays = torch.from_numpy(np.ascontiguousarray(rays[perm])).cuda()
        for _ in range(3):
            a.occluded_dev(d_rays, nr, d_occ, stream.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        steps = 8
        for _ in range(steps):
            a.occluded_dev(d_rays, nr, d_occ, stream.cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        occ = np.empty(nr, dtype=np.uint8)
        occ[perm] = d_occ.cpu().numpy()
        if ref is None:
            ref = occ
        ok = np.array_equal(occ, ref)
        print(f"order {name:11s} {ms:7.3f} ms  {nr / ms / 1e3:8.1f} Mrays/s  same_answers={ok}", flush=True)
        del d_rays


if __name__ == "__main__":
    t0 = time.time()
    main()
    print(f"exp_sort done in {time.time() - t0:.1f}s")
