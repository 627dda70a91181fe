#!/usr/bin/env python
"""bench.py -- Mrays/s of the ray-intersection hot path on BASELINE.json configs[2]:
synthetic 1 M-triangle soup, 16 M incoherent ambient-occlusion hemisphere rays, per GPU.

    python bench.py --gpus N --steps K --warmup W          (torchrun launches it for N > 1, one rank per GPU)
    python bench.py --impl reference ...                   the reference's own CPU traversal on the host cores

A "step" is one pass of the hot path over one batch: one occlusion (any-hit) traversal of the whole 16 Mi-ray batch
on the reference-identical BVH.  `value` is whole-job Mrays/s with rays resident in HBM; `e2e` is the same metric
through the C-ABI call with pinned HOST buffers (H2D of the rays and D2H of the occlusion bytes inside the timed
region).  Multi-GPU: the scene is replicated, every rank traces its own 16 Mi-ray batch (weak scaling), no data-path
collective (rays never interact); timing is barrier + CUDA events, max over ranks.
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
        "config": {"workload": WORKLOAD, "sampled": True},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": kind, "sample": sample, "single_thread": single},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


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
           "sample": f"first {len(rays6)} rays of the batch, closest-hit ri_bvh_intersect incl. state build, {threads} threads"}
    if gpu_occ is not None:
        out["parity"] = parity_sample(tris, rays8_sample, gpu_occ, ref_scene)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    a = accel.Accel.bind(accel.RI_ACCEL_B200).build(tris, accel.PREC_F32, device=local_rank)
    info = a.info()
    order = a.triorder()
    P, n = primary_points(a.intersect, tris[order])
    # every rank traces its own batch: same points, its own hemisphere samples (weak scaling)
    rays_np = scenes.ao_rays(P[:npoints], n[:npoints], NTHETA, NPHI, scenes.SEED_C3 + 1000 * rank)
    log(f"[rank {rank}] setup {time.time() - t0:.1f}s: build {info.build_seconds:.2f}s, {info.ninner} inner nodes, depth {info.max_depth}, "
        f"{info.device_bytes / 1e6:.1f} MB on device")

    h_rays = torch.empty((nrays, 8), dtype=torch.float32, pin_memory=True)
    h_rays.numpy()[:] = rays_np
    h_occ = torch.empty((nrays,), dtype=torch.uint8, pin_memory=True)
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
    for _ in range(2):
        step_closest()
    ms_c = timed(step_closest, max(3, args.steps // 2))
    closest = world * nrays * max(3, args.steps // 2) / (ms_c * 1e-3) / 1e6

    # end to end through the host-buffer C-ABI call: pinned host rays in, occlusion bytes out
    rays_host, occ_host = h_rays.numpy(), h_occ.numpy()

    def e2e_step():
        accel._check(a.lib.ri_b200_occluded_batch_f32(a.data, accel._ptr(rays_host), nrays, accel._ptr(occ_host)))

    for _ in range(2):
        e2e_step()
    barrier()
    t1 = time.perf_counter()
    esteps = max(3, args.steps // 2)
    for _ in range(esteps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t1) * 1e3
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * nrays * esteps / (e2e_ms * 1e-3) / 1e6
    assert np.array_equal(occ_host, d_occ.cpu().numpy()), "host-buffer path and device path disagree"

    if rank == 0:
        peak, peak_src = peaks()
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes of ONE launch of this kernel on this batch (ncu --set full)
        if os.path.exists(tp) and npoints == NPOINTS:
            with open(tp) as f:
                traffic = json.load(f).get("occluded_f32_c3_bytes_per_launch")
        per_gpu_rays_s = nrays * args.steps / (ms * 1e-3)
        achieved = per_gpu_rays_s * b_ray / 1e9
        out = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "kernel": "occlusion (any-hit) traversal, reference-identical binary BVH, leaf<=16",
                       "rays_per_gpu_per_step": nrays, "ntris": NTRIS, "inner_nodes": int(info.ninner), "depth": int(info.max_depth),
                       "l2": "ray batch (512 MiB) exceeds L2 every step; scene records (54 MB) are L2-resident by nature of the workload",
                       "closest_hit_mrays_s": closest, "occluded_fraction": float(occ_host.mean())},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": world * nrays * 32, "d2h_bytes_per_step": world * nrays,
                    "ms_per_step": e2e_ms / esteps, "api": "ri_b200_occluded_batch_f32 (pinned host buffers)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "bytes_per_ray": b_ray, "inner_visits_per_ray": I, "tri_tests_per_ray": T,
                         "closest_hit": {"bytes_per_ray": b_ray_c, "inner_visits_per_ray": Ic, "tri_tests_per_ray": Tc,
                                         "achieved": nrays * max(3, args.steps // 2) / (ms_c * 1e-3) * b_ray_c / 1e9,
                                         "frac": nrays * max(3, args.steps // 2) / (ms_c * 1e-3) * b_ray_c / 1e9 / peak}},
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
