#!/usr/bin/env python
"""bench.py -- Mrays/s of the ray-intersection hot path on BASELINE.json configs[2]:
synthetic 1 M-triangle soup, 16 M incoherent ambient-occlusion hemisphere rays, per GPU.

    python bench.py --gpus N --steps K --warmup W          (torchrun launches it for N > 1, one rank per GPU)
    python bench.py --impl reference ...                   the reference's own CPU traversal on the host cores

A "step" is one pass of the hot path over one batch: one occlusion (any-hit) traversal of the whole 16 Mi-ray batch
on the reference-identical BVH.  `value` is whole-job Mrays/s with rays resident in HBM; `e2e` is the same metric
through the reference-facing C-ABI call with HOST buffers: ri_b200_occlusion_points_f32 = calculate_occlusion
(ambientocclusion.c:42-151) for the batch's 262 144 shading points -- pinned host points in (H2D), rays generated and
traced on the device, occluded-ray counts out (D2H), all inside the timed region; the older per-ray host path
(ri_b200_occluded_batch_f32: 32 B per ray up, 1 B per ray down) is reported beside it.  Multi-GPU: the scene is
replicated, every rank traces its own 16 Mi-ray batch (weak scaling), no data-path collective (rays never interact);
timing is barrier + CUDA events, max over ranks.

`config.frames` is the other multi-GPU shape BASELINE.json names (configs[4] and configs[0]): ONE frame strong-scaled over
the N ranks -- 32x32 buckets of the spiral order dealt round-robin to the ranks, scene replicated, the framebuffer gathered
on rank 0 either by the resolve kernels themselves storing into rank 0's buffer over NVLink peer memory (fused) or by one
NCCL gather of packed slabs.  Reported per frame: device time per rank, wall time until rank 0 holds the host framebuffer,
and the frame's sha256 (equal to the single-rank frame; for ambient_occlusion.rib equal to the reference's own framebuffer).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from lucille_b200 import scenes  # noqa: E402

NTRIS = 1_000_000
NPOINTS = 262_144
NTHETA = NPHI = 8
NRAYS = NPOINTS * NTHETA * NPHI          # 16 777 216
WORKLOAD = "configs[2]: synthetic 1M-triangle soup, 16Mi incoherent AO hemisphere rays (8x8 per point, 262144 primary-hit points)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class quiet_stdout:
    """The compiled reference logs with printf ("[lucille] info ..."): send fd 1 to stderr while it runs so that this
    process prints exactly ONE line on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def primary_points(intersect_f32, tris_post):
    """First NPOINTS primary hits of the 1024x1024 pinhole camera -> (P, Ns) in float64."""
    rays = scenes.pinhole_rays(1024, 1024)
    hits = intersect_f32(rays)
    idx = np.flatnonzero(hits["prim"] != 0xFFFFFFFF)
    if len(idx) < NPOINTS:
        raise RuntimeError(f"only {len(idx)} primary hits")
    idx = idx[:NPOINTS]
    t = hits["t"][idx].astype(np.float64)
    P = rays[idx, 0:3].astype(np.float64) + rays[idx, 4:7].astype(np.float64) * t[:, None]
    tri = tris_post[hits["prim"][idx]]
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return P, n


def algorithmic_bytes_per_ray(cnt, anyhit=True):
    """B_ray = 32 (ray in) + 1|16 (occlusion byte | hit record out) + 64*I + 48*T (SURVEY 8d / BASELINE.md section 4)."""
    I = cnt["ninner"] / cnt["nrays_total"]
    T = cnt["ntris"] / cnt["nrays_total"]
    return 32.0 + (1.0 if anyhit else 16.0) + 64.0 * I + 48.0 * T, I, T


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU traversal (compiled, unmodified: oracle/_ref) on the host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    cores = os.cpu_count() or 1
    tris = scenes.triangle_soup(NTRIS, scenes.SEED_C3)
    orc = ol.Oracle().build(tris)
    order = orc.triorder()
    P, n = primary_points(orc.intersect_f32, tris[order])
    sample_pts = max(1024, min(NPOINTS, (200_000 * cores) // (NTHETA * NPHI)))
    rays8 = scenes.ao_rays(P[:sample_pts], n[:sample_pts], NTHETA, NPHI, scenes.SEED_C3)
    rays6 = scenes.rays_f32_to_f64(rays8)
    if ol.reference_available():
        with quiet_stdout():
            kind, scene = "reference", ol.Reference().build(tris)
        step = lambda: scene.intersect(rays6, nthreads=cores, want_hits=False)[1]   # noqa: E731
        threads = cores
    else:
        kind, threads = "port", 1
        def step():
            t0 = time.perf_counter()
            orc.occluded_f64(rays6)
            return time.perf_counter() - t0
    for _ in range(args.warmup):
        step()
    secs = [step() for _ in range(args.steps)]
    total = sum(secs)
    value = len(rays6) * args.steps / total / 1e6
    sample = f"{len(rays6)} of the {NRAYS} rays per step (first {sample_pts} points), closest-hit ri_bvh_intersect incl. state build"
    single = None
    if kind == "reference":                       # the one-thread figure beside the all-threads one (best of 3, bounded sample)
        one = rays6[: max(50_000, len(rays6) // (4 * cores))]
        sec1 = min(scene.intersect(one, nthreads=1, want_hits=False)[1] for _ in range(3))
        single = {"value": len(one) / sec1 / 1e6, "unit": "Mrays/s", "cores": 1, "sample": f"first {len(one)} rays of the same sample, best of 3"}
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "kernel": "occlusion (any-hit) traversal, reference-identical binary BVH, leaf<=16",
                   "rays_per_gpu_per_step": NRAYS, "ntris": NTRIS, "sampled": True,
                   "sample": "each step traces a bounded sample of the batch (see cpu_baseline.sample): the rate is per ray"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample, "single_thread": single,
                         "baseline_kernel": BASELINE_KERNEL[kind]},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# what the CPU arm runs per ray, next to what the GPU arm runs (the reference's AO loop calls the closest-hit query and builds the
# hit state for every occlusion ray, ambientocclusion.c:123; the GPU answers the same question with an any-hit query)
BASELINE_KERNEL = {
    "reference": {"function": "ri_bvh_intersect (bvh.c:430-542) via ri_raytrace", "query": "closest hit + ri_intersection_state_build",
                  "precision": "f64", "gpu_query": "any-hit (occlusion), fp32 records"},
    "port": {"function": "orc_occluded_f64 (oracle/lucille_oracle.c, restatement of bvh_traverse)", "query": "any-hit",
             "precision": "f64", "gpu_query": "any-hit (occlusion), fp32 records"},
}


def parity_sample(tris, rays8, gpu_occ, ref_scene=None):
    """Part of the cpu_baseline leg (rank 0, N=1, outside every timed region): the occlusion bytes the GPU produced for the first rays
    of the batch against the oracle's fp32 instantiation (the bit-exact contract of the fp32 records) and against the compiled
    reference's double closest-hit flags (the fraction of rays where fp32 records and the double reference disagree)."""
    try:
        import oracle_lib as ol
        n = min(len(rays8), 200_000)
        want = ol.Oracle().build(tris).occluded_f32(rays8[:n])
        got = np.asarray(gpu_occ[:n])
        res = {"rays": int(n), "mismatches_vs_oracle_f32": int(np.count_nonzero((got != 0) != (want != 0)))}
        if ref_scene is not None:
            hits, _ = ref_scene.intersect(scenes.rays_f32_to_f64(rays8[:n]), nthreads=os.cpu_count() or 1, want_hits=True)
            res["disagree_with_reference_f64"] = float(np.mean((got != 0) != (hits["hit"] == 1)))
        return res
    except Exception as e:      # pragma: no cover -- the bench line must survive a broken checker
        return {"error": repr(e)}


def cpu_baseline(tris, rays8_sample, gpu_occ=None):
    """Reported baseline (rank 0, N=1): the compiled reference on the box's host cores, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import oracle_lib as ol
    except Exception as e:      # pragma: no cover
        return {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "unavailable", "sample": str(e)}
    cores = os.cpu_count() or 1
    rays6 = scenes.rays_f32_to_f64(rays8_sample)
    if ol.reference_available():
        with quiet_stdout():
            scene = ol.Reference().build(tris)
        scene.intersect(rays6[:20000], nthreads=cores, want_hits=False)
        sec = min(scene.intersect(rays6, nthreads=cores, want_hits=False)[1] for _ in range(2))
        kind, threads = "reference", cores
        ref_scene = scene
    else:
        ref_scene = None
        t = ol.Oracle().build(tris)
        t0 = time.perf_counter()
        t.occluded_f64(rays6[:200000])
        sec, kind, threads = (time.perf_counter() - t0) * len(rays6) / 200000, "port", 1
    out = {"value": len(rays6) / sec / 1e6, "unit": "Mrays/s", "cores": threads, "kind": kind,
           "sample": f"first {len(rays6)} rays of the batch, closest-hit ri_bvh_intersect incl. state build, {threads} threads",
           "baseline_kernel": BASELINE_KERNEL[kind]}
    if gpu_occ is not None:
        out["parity"] = parity_sample(tris, rays8_sample, gpu_occ, ref_scene)
    return out


def _wall_max(fn, world, dist, torch):
    """barrier | fn() | barrier, wall seconds, max over ranks; returns (seconds, fn's result)."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt, out


def frame_leg(args, rank, world, local_rank, torch, dist):
    """config.frames: the tiled multi-GPU frames of BASELINE configs[4] (10 M-triangle soup) and configs[0] (ambient_occlusion.rib)."""
    import hashlib
    import math
    from lucille_b200 import accel, distributed

    def gather_stats(stats):
        t = torch.tensor([stats.ms_total, float(stats.nrays)], device="cuda", dtype=torch.float64)
        parts = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(parts, t)
        else:
            parts = [t]
        return [float(x[0]) for x in parts], sum(float(x[1]) for x in parts)

    out = {}
    # ---- configs[4]: synthetic 10 M-triangle soup, 4096 x 4096, AO; strong scaling over the ranks ------------------------------
    ntris, res = args.frame_tris, args.frame_res
    t0 = time.time()
    tris = scenes.triangle_soup(ntris, scenes.SEED_C5)
    a = accel.Accel.bind(accel.RI_ACCEL_B200).build(tris, accel.PREC_F32, device=local_rank)
    info = a.info()
    del tris
    log(f"[rank {rank}] frame leg: {ntris} triangles built in {info.build_seconds:.2f}s (setup {time.time() - t0:.1f}s), "
        f"{info.device_bytes / 1e6:.0f} MB of records")
    c2w = np.eye(4)
    c2w[3, :3] = (0.5, 0.5, -2.0)
    fr = accel.make_frame(c2w.reshape(16), 1.0 / math.tan(math.radians(40.0) / 2), False, res, res, 1, 1, 64, rng_mode=1, seed=5,
                          precision=accel.PREC_F32)
    fb = distributed.PeerFramebuffer(res, res, rank, world, local_rank)
    distributed.render_ao_distributed_peer(a, fr, fb)                                           # warm-up (allocations, first launches)
    wall_p, (rgb_p, st_p) = _wall_max(lambda: distributed.render_ao_distributed_peer(a, fr, fb), world, dist, torch)
    dev_p, nrays = gather_stats(st_p)
    if rank == 0:                                    # rgb_p / rgb_n are views of buffers the next frame reuses: hash them now
        sha_p, mean_p = hashlib.sha256(np.ascontiguousarray(rgb_p).tobytes()).hexdigest(), float(rgb_p.mean())
    distributed.render_ao_distributed(a, fr, rank, world)                                       # warm-up of this arm too (pixel lists, NCCL buffers)
    wall_n, (rgb_n, st_n) = _wall_max(lambda: distributed.render_ao_distributed(a, fr, rank, world), world, dist, torch)
    dev_n, _ = gather_stats(st_n)
    if rank == 0:
        sha_n = hashlib.sha256(np.ascontiguousarray(rgb_n).tobytes()).hexdigest()
    fb.close()
    c5 = {"scene": f"synthetic {ntris}-triangle soup (seed C5), {info.device_bytes / 1e6:.0f} MB of records per GPU (beyond L2)",
          "frame": f"{res}x{res}, 1x1 pixel samples, 8x8 AO rays per hit, fp32 records, counter RNG; 32x32 buckets b % {world} == rank",
          "rays": nrays, "scaling": "strong"}
    if rank == 0:
        c5["fused_peer_store"] = {"wall_ms_to_rank0_host_framebuffer": wall_p * 1e3, "device_ms_per_rank": dev_p,
                                  "mrays_s": nrays / wall_p / 1e6, "sha256": sha_p}
        c5["nccl_gather"] = {"wall_ms_to_rank0_host_framebuffer": wall_n * 1e3, "device_ms_per_rank": dev_n,
                             "mrays_s": nrays / wall_n / 1e6, "sha256": sha_n}
        if world > 1:                                                                           # the same frame rendered by rank 0 alone
            one, _ = a.render_ao(fr)
            sha_1 = hashlib.sha256(np.ascontiguousarray(one).tobytes()).hexdigest()
            c5["equals_one_rank_frame"] = bool(sha_1 == sha_p and sha_1 == sha_n)
        else:
            c5["equals_one_rank_frame"] = bool(sha_p == sha_n)
        c5["mean"] = mean_p
    out["configs[4]"] = c5
    a.free()

    # ---- configs[0]: examples/ambient_occlusion/ambient_occlusion.rib, 640 x 480, 3 x 3, 64 AO rays, the reference's MT19937 stream --
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_scene.npz"))
    dig = np.load(os.path.join(ROOT, "tests", "golden", "c1_frame_640x480_digest.npz"))
    cam = g["cam"]
    c1a = accel.Accel.bind(accel.RI_ACCEL_B200).build(g["tris"], accel.PREC_F64, device=local_rank)
    fr1 = accel.make_frame(cam[:16], cam[16], bool(cam[17]), 640, 480, 3, 3, gather_nsamples=64)
    fb1 = distributed.PeerFramebuffer(640, 480, rank, world, local_rank)
    distributed.render_ao_distributed_peer(c1a, fr1, fb1)
    wall_1, (rgb_1, st_1) = _wall_max(lambda: distributed.render_ao_distributed_peer(c1a, fr1, fb1), world, dist, torch)
    dev_1, nrays_1 = gather_stats(st_1)
    rgb_1 = None if rgb_1 is None else np.array(rgb_1, dtype=np.float32)          # rgb_1 is a view of fb1's pinned buffer: copy before close()
    fb1.close()
    c1 = {"scene": "ambient_occlusion.rib (322 triangles), 640x480, PixelSamples 3 3, 64 AO rays per hit, fp64 records, the reference's "
                   "one MT19937 stream shared by the ranks (per-bucket hit counts all-gathered between eye pass and gather pass)",
          "rays": nrays_1, "scaling": "strong"}
    if rank == 0:
        sha = hashlib.sha256(np.ascontiguousarray(rgb_1, dtype=np.float32).tobytes()).hexdigest()
        c1.update({"wall_ms_to_rank0_host_framebuffer": wall_1 * 1e3, "device_ms_per_rank": dev_1, "mrays_s": nrays_1 / wall_1 / 1e6,
                   "sha256": sha, "equals_reference_framebuffer": bool(sha == str(dig["sha256"])),
                   "reference_rays": int(dig["nrays"]), "reference_seconds_1thread": float(dig["seconds_1thread"])})
    out["configs[0]"] = c1
    c1a.free()
    return out


def kernel_metrics():
    """ncu counters of ONE launch of the timed kernel on this batch (profiles/r02_kernel_metrics.json, written by
    scripts/make_profiles.py from a committed `ncu --set full` capture): what physically limits the kernel."""
    p = os.path.join(ROOT, "profiles", "r02_kernel_metrics.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-frames", action="store_true", help="skip config.frames (the tiled multi-GPU frame leg)")
    ap.add_argument("--frame-tris", type=int, default=10_000_000)
    ap.add_argument("--frame-res", type=int, default=4096)
    ap.add_argument("--points", type=int, default=NPOINTS, help="debug: fewer AO points (invalidates the number)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from lucille_b200 import accel

    if not torch.cuda.is_available() or accel.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: lucille_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    npoints = args.points
    nrays = npoints * NTHETA * NPHI
    t0 = time.time()
    tris = scenes.triangle_soup(NTRIS, scenes.SEED_C3)
    # both record sets: fp32 for the headline batch, double for the double-exact leg (hybrid.cuh reads both)
    a = accel.Accel.bind(accel.RI_ACCEL_B200).build(tris, accel.PREC_F32 | accel.PREC_F64, device=local_rank)
    info = a.info()
    order = a.triorder()
    P, n = primary_points(a.intersect, tris[order])
    # every rank traces its own batch: same points, its own hemisphere samples (weak scaling).  The batch is the one the point entry
    # generates on the device (calculate_occlusion's ray set-up; oracle restatement orc_ao_point_rays_f32, tests/test_gpu_fullsize.py)
    seed = scenes.SEED_C3 + 1000 * rank
    pts_np = np.ascontiguousarray(np.concatenate([P[:npoints], n[:npoints]], axis=1))
    h_rays = torch.empty((nrays, 8), dtype=torch.float32, pin_memory=True)
    accel._check(a.lib.ri_b200_ao_point_rays_f32(a.data, accel.C.byref(accel.AoPoints(NTHETA, NPHI, seed, 1.0e-6)), accel._ptr(pts_np), npoints,
                                                  accel._ptr(h_rays.numpy())))
    rays_np = h_rays.numpy()
    log(f"[rank {rank}] setup {time.time() - t0:.1f}s: build {info.build_seconds:.2f}s, {info.ninner} inner nodes, depth {info.max_depth}, "
        f"{info.device_bytes / 1e6:.1f} MB on device")

    h_occ = torch.empty((nrays,), dtype=torch.uint8, pin_memory=True)
    h_pts = torch.empty((npoints, 6), dtype=torch.float64, pin_memory=True)
    h_pts.numpy()[:] = pts_np
    h_cnt = torch.empty((npoints,), dtype=torch.int32, pin_memory=True)
    d_rays = h_rays.cuda(non_blocking=True)
    d_occ = torch.empty((nrays,), dtype=torch.uint8, device="cuda")
    d_hits = torch.empty((nrays, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    # a real (non-default) stream: the C ABI treats a NULL stream as "the accelerator's own stream", and CUDA events
    # only see work on the stream they are recorded on
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    # counters of the reference traversal order on this exact batch (outside the timed region)
    cnt = a.count(rays_np, anyhit=True)
    cnt["nrays_total"] = nrays
    cnt_closest = a.count(rays_np, anyhit=False)
    cnt_closest["nrays_total"] = nrays
    b_ray, I, T = algorithmic_bytes_per_ray(cnt, anyhit=True)
    b_ray_c, Ic, Tc = algorithmic_bytes_per_ray(cnt_closest, anyhit=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def wall(fn, steps):
        barrier()
        t1 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        ms = (time.perf_counter() - t1) * 1e3
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    step_any = lambda: a.occluded_dev(d_rays, nrays, d_occ, stream)      # noqa: E731
    step_closest = lambda: a.intersect_dev(d_rays, nrays, d_hits, stream)  # noqa: E731

    for _ in range(args.warmup):
        step_any()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = accel.launch_count()
    ms = timed(step_any, args.steps)
    launches = accel.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * nrays * args.steps / (ms * 1e-3) / 1e6

    # closest-hit on the same batch (reported, not the headline)
    csteps = max(3, args.steps // 2)
    for _ in range(2):
        step_closest()
    ms_c = timed(step_closest, csteps)
    closest = world * nrays * csteps / (ms_c * 1e-3) / 1e6

    # the same batch as DOUBLE rays with the double reference's answer for every ray (what real, RIB-scale scenes need): through the
    # fp32 records with certified decisions (csrc/hybrid.cuh) and, for comparison, through the double kernel alone
    d_rays64 = torch.from_numpy(np.ascontiguousarray(rays_np[:, [0, 1, 2, 4, 5, 6]].astype(np.float64))).cuda()
    d_occ_h = torch.empty((nrays,), dtype=torch.uint8, device="cuda")
    d_occ_p = torch.empty((nrays,), dtype=torch.uint8, device="cuda")
    step_hyb = lambda: a.occluded_dev(d_rays64, nrays, d_occ_h, stream, f64=True)      # noqa: E731
    step_f64 = lambda: a.occluded_dev(d_rays64, nrays, d_occ_p, stream, f64=True)      # noqa: E731
    dsteps = 3
    step_hyb()
    ms_h = timed(step_hyb, dsteps)
    os.environ["B200_HYBRID"] = "0"
    step_f64()
    ms_p = timed(step_f64, dsteps)
    os.environ.pop("B200_HYBRID", None)
    d_h64a = torch.empty((nrays, 4), dtype=torch.float64, device="cuda")
    d_h64b = torch.empty((nrays, 4), dtype=torch.float64, device="cuda")
    step_ch = lambda: a.intersect_dev(d_rays64, nrays, d_h64a, stream, f64=True)       # noqa: E731
    step_cp = lambda: a.intersect_dev(d_rays64, nrays, d_h64b, stream, f64=True)       # noqa: E731
    step_ch()
    ms_ch = timed(step_ch, dsteps)
    os.environ["B200_HYBRID_CLOSEST"] = "0"
    step_cp()
    ms_cp = timed(step_cp, dsteps)
    os.environ.pop("B200_HYBRID_CLOSEST", None)
    closest_identical = bool(torch.equal(d_h64a.view(torch.int64), d_h64b.view(torch.int64)))
    del d_h64a, d_h64b
    h_cnt64 = torch.empty((npoints,), dtype=torch.int32, pin_memory=True)

    def e2e64_step():
        accel._check(a.lib.ri_b200_occlusion_points_f64(a.data, accel.C.byref(accel.AoPoints(NTHETA, NPHI, seed, 1.0e-6)), accel._ptr(h_pts.numpy()),
                                                        npoints, accel._ptr(h_cnt64.numpy())))
    e2e64_step()
    e2e64_ms = wall(e2e64_step, dsteps)
    double_exact = {
        "what": "the same batch as double rays, the DOUBLE reference's verdict for every ray",
        "hybrid_mrays_s": world * nrays * dsteps / (ms_h * 1e-3) / 1e6,
        "double_kernel_mrays_s": world * nrays * dsteps / (ms_p * 1e-3) / 1e6,
        "identical_verdicts": bool(torch.equal(d_occ_h, d_occ_p)),
        "closest_hit_hybrid_mrays_s": world * nrays * dsteps / (ms_ch * 1e-3) / 1e6,
        "closest_hit_double_kernel_mrays_s": world * nrays * dsteps / (ms_cp * 1e-3) / 1e6,
        "closest_hit_identical_records": closest_identical,
        "e2e_points_f64_mrays_s": world * nrays * dsteps / (e2e64_ms * 1e-3) / 1e6,
        "api": "ri_b200_occluded_dev_f64 / ri_b200_occlusion_points_f64 (host points in, double rays generated + traced on the device, counts out)",
    }
    del d_rays64, d_occ_h, d_occ_p

    # end to end through the reference-facing host call: calculate_occlusion for the batch's shading points.  Pinned host points in,
    # occluded-ray counts out; the rays are generated and traced on the device inside the call.
    par = accel.AoPoints(NTHETA, NPHI, seed, 1.0e-6)
    pts_host, cnt_host = h_pts.numpy(), h_cnt.numpy()

    def e2e_step():
        accel._check(a.lib.ri_b200_occlusion_points_f32(a.data, accel.C.byref(par), accel._ptr(pts_host), npoints, accel._ptr(cnt_host)))

    esteps = max(3, args.steps // 2)
    for _ in range(2):
        e2e_step()
    e2e_ms = wall(e2e_step, esteps)
    e2e_value = world * nrays * esteps / (e2e_ms * 1e-3) / 1e6

    # the per-ray host path (rays up, occlusion bytes down), as in round 1
    rays_host, occ_host = h_rays.numpy(), h_occ.numpy()

    def e2e_rays_step():
        accel._check(a.lib.ri_b200_occluded_batch_f32(a.data, accel._ptr(rays_host), nrays, accel._ptr(occ_host)))

    for _ in range(2):
        e2e_rays_step()
    e2e_rays_ms = wall(e2e_rays_step, esteps)
    e2e_rays_value = world * nrays * esteps / (e2e_rays_ms * 1e-3) / 1e6
    assert np.array_equal(occ_host, d_occ.cpu().numpy()), "host-buffer path and device path disagree"
    assert np.array_equal(cnt_host.astype(np.int64), occ_host.reshape(npoints, NTHETA * NPHI).sum(axis=1, dtype=np.int64)), \
        "point entry's counts and the per-ray occlusion bytes disagree"

    frames = None
    if not args.no_frames:
        a_keep = a                      # the C3 accelerator stays alive for the cpu_baseline parity sample below
        frames = frame_leg(args, rank, world, local_rank, torch, dist)
        a = a_keep

    if rank == 0:
        peak, peak_src = peaks()
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes of ONE launch of this kernel on this batch (ncu --set full)
        if os.path.exists(tp) and npoints == NPOINTS:
            with open(tp) as f:
                traffic = json.load(f).get("occluded_f32_c3_bytes_per_launch")
        per_gpu_rays_s = nrays * args.steps / (ms * 1e-3)
        achieved = per_gpu_rays_s * b_ray / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "achieved_kind": "modelled: measured rays/s x ALGORITHMIC bytes per ray (counters of the reference "
                "traversal order on this batch); the measured DRAM bytes are `traffic`, their rate over the peak is `dram_frac`",
                "bytes_per_ray": b_ray, "inner_visits_per_ray": I, "tri_tests_per_ray": T,
                "note": "contract roofline: ALGORITHMIC bytes of the reference traversal (SURVEY 8d) over the measured HBM copy bandwidth; the "
                        "records are L2-resident, so the physical limiter is NOT HBM -- see `limiter` (ncu, one launch of this kernel)",
                "closest_hit": {"bytes_per_ray": b_ray_c, "inner_visits_per_ray": Ic, "tri_tests_per_ray": Tc,
                                "achieved": nrays * csteps / (ms_c * 1e-3) * b_ray_c / 1e9,
                                "frac": nrays * csteps / (ms_c * 1e-3) * b_ray_c / 1e9 / peak}}
        km = kernel_metrics()
        if km and traffic:
            roof["dram_frac"] = traffic / (ms / args.steps * 1e-3) / 1e9 / peak
        if km:
            roof["limiter"] = km
            # `bound` above is the contract's label for the algorithmic-bytes roofline; what ncu shows as the limiting unit:
            roof["physical_bound"] = ("l1_data_pipe_lsu_wavefronts" if km.get("lsu_wavefront_frac", 0) >= km.get("issue_frac", 0) else "warp_instruction_issue")
            roof["physical_bound_frac"] = max(km.get("lsu_wavefront_frac", 0), km.get("issue_frac", 0))
        out = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "kernel": "occlusion (any-hit) traversal, reference-identical binary BVH, leaf<=16",
                       "rays_per_gpu_per_step": nrays, "ntris": NTRIS, "inner_nodes": int(info.ninner), "depth": int(info.max_depth),
                       "l2": "ray batch (512 MiB) exceeds L2 every step; scene records (54 MB) are L2-resident by nature of the workload",
                       "closest_hit_mrays_s": closest, "double_exact": double_exact, "occluded_fraction": float(occ_host.mean()),
                       "frames": frames},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": world * npoints * 48, "d2h_bytes_per_step": world * npoints * 4,
                    "ms_per_step": e2e_ms / esteps,
                    "api": "ri_b200_occlusion_points_f32 (calculate_occlusion for 262144 shading points: pinned host points in, rays generated + "
                           "traced on the device, occluded-ray counts out)",
                    "ray_batch": {"value": e2e_rays_value, "unit": "Mrays/s", "h2d_bytes_per_step": world * nrays * 32,
                                  "d2h_bytes_per_step": world * nrays, "ms_per_step": e2e_rays_ms / esteps,
                                  "api": "ri_b200_occluded_batch_f32 (pinned host ray records in, occlusion bytes out)"}},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            n_sample = min(nrays, 150_000 * (os.cpu_count() or 1))
            out["cpu_baseline"] = cpu_baseline(tris, rays_np[:n_sample], occ_host[:n_sample])
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
