"""Diagnostics on the GPU box: copy bandwidths, kernel times per variant on the C3 batch."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucille_b200 import accel, scenes
sys.path.insert(0, "."); import bench

npoints = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tris = scenes.triangle_soup(1_000_000, scenes.SEED_C3)
a = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.PREC_F64)
P, n = bench.primary_points(a.intersect, tris[a.triorder()])
rays = scenes.ao_rays(P[:npoints], n[:npoints], 8, 8, scenes.SEED_C3)
nr = len(rays)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
h = torch.empty((nr, 8), dtype=torch.float32, pin_memory=True); h.numpy()[:] = rays
d = h.cuda(); occ = torch.empty(nr, dtype=torch.uint8, device="cuda"); hits = torch.empty((nr, 4), dtype=torch.float32, device="cuda")
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = ev(lambda: d.copy_(h, non_blocking=True)); print(f"H2D pinned {nr*32/ms/1e6:.1f} GB/s")
ho = torch.empty(nr, dtype=torch.uint8, pin_memory=True)
ms = ev(lambda: ho.copy_(occ, non_blocking=True)); print(f"D2H pinned {nr/ms/1e6:.1f} GB/s")
ms = ev(lambda: a.occluded_dev(d, nr, occ, st.cuda_stream)); print(f"anyhit  f32 {ms:.3f} ms  {nr/ms/1e3:.1f} Mrays/s")
ms = ev(lambda: a.intersect_dev(d, nr, hits, st.cuda_stream)); print(f"closest f32 {ms:.3f} ms  {nr/ms/1e3:.1f} Mrays/s")
r6 = torch.from_numpy(scenes.rays_f32_to_f64(rays)).cuda(); h64 = torch.empty((nr, 4), dtype=torch.float64, device="cuda")
ms = ev(lambda: a.occluded_dev(r6, nr, occ, st.cuda_stream, f64=True)); print(f"anyhit  f64 {ms:.3f} ms  {nr/ms/1e3:.1f} Mrays/s")
ms = ev(lambda: a.intersect_dev(r6, nr, h64, st.cuda_stream, f64=True)); print(f"closest f64 {ms:.3f} ms  {nr/ms/1e3:.1f} Mrays/s")
# coherent primaries (C2-like on the 1M soup)
pr = torch.from_numpy(scenes.pinhole_rays(1024, 1024)).cuda(); ph = torch.empty((1 << 20, 4), dtype=torch.float32, device="cuda")
ms = ev(lambda: a.intersect_dev(pr, 1 << 20, ph, st.cuda_stream)); print(f"primary closest f32 (1M tris) {ms:.3f} ms  {(1<<20)/ms/1e3:.1f} Mrays/s")
print("counters anyhit", a.count(rays, anyhit=True)); print("counters closest", a.count(rays, anyhit=False))
# C1 frame
g = np.load("tests/golden/c1_scene.npz"); cam = g["cam"]
c1 = accel.Accel.bind().build(g["tris"], accel.PREC_F64 | accel.PREC_F32)
for prec, mode in ((accel.PREC_F64, 0), (accel.PREC_F64, 1), (accel.PREC_F32, 1)):
    fr = accel.make_frame(cam[:16], cam[16], bool(cam[17]), 640, 480, 3, 3, 64, rng_mode=mode, precision=prec)
    c1.render_ao(fr)
    rgb, s = c1.render_ao(fr)
    print(f"C1 prec={prec} rng={mode}: total {s.ms_total:.2f} ms (primary {s.ms_primary:.2f}, rng {s.ms_rng:.2f}, ao {s.ms_ao:.2f}, resolve {s.ms_resolve:.2f}) rays {s.nrays} -> {s.nrays/s.ms_total/1e3:.1f} Mrays/s")
# synthetic frame (C5-like, scaled): 1M-triangle soup, 1024x1024, 2x2 sub-samples, 8x8 AO rays, counter RNG, fp32
c2w = np.eye(4); c2w[3, :3] = (0.5, 0.5, -2.0)
import math, os
for label, env in (("wavefront", "0"), ("fused", "1")):
    os.environ["B200_FUSED_AO"] = env
for prec in (accel.PREC_F32, accel.PREC_F64):
    fr = accel.make_frame(c2w.reshape(16), 1.0 / math.tan(math.radians(40.0) / 2), False, 1024, 1024, 2, 2, 64, rng_mode=1, seed=5, precision=prec)
    a.render_ao(fr)
    rgb, s = a.render_ao(fr)
    print(f"soup frame prec={prec}: total {s.ms_total:.2f} ms (primary {s.ms_primary:.2f}, ao {s.ms_ao:.2f}) rays {s.nrays} hits {s.nhits_primary} -> {s.nrays/s.ms_total/1e3:.1f} Mrays/s  mean {rgb.mean():.4f}")
# C2: 100K-triangle soup, 1024x1024 coherent primary rays, closest hit (BASELINE configs[1])
t2 = scenes.triangle_soup(100_000, scenes.SEED_C2)
a2 = accel.Accel.bind().build(t2, accel.PREC_F32 | accel.PREC_F64)
pr2 = torch.from_numpy(scenes.pinhole_rays(1024, 1024)).cuda(); ph2 = torch.empty((1 << 20, 4), dtype=torch.float32, device="cuda")
ms = ev(lambda: a2.intersect_dev(pr2, 1 << 20, ph2, st.cuda_stream), reps=20); print(f"C2 closest f32 (100K tris, 1M primary rays) {ms:.3f} ms  {(1<<20)/ms/1e3:.1f} Mrays/s")
c2cnt = a2.count(scenes.pinhole_rays(1024, 1024)); print("C2 counters", c2cnt, "bytes/ray", 32 + 16 + 64 * c2cnt["ninner"] / c2cnt["nrays"] + 48 * c2cnt["ntris"] / c2cnt["nrays"])
pr2d = torch.from_numpy(scenes.rays_f32_to_f64(scenes.pinhole_rays(1024, 1024))).cuda(); ph2d = torch.empty((1 << 20, 4), dtype=torch.float64, device="cuda")
ms = ev(lambda: a2.intersect_dev(pr2d, 1 << 20, ph2d, st.cuda_stream, f64=True), reps=20); print(f"C2 closest f64 {ms:.3f} ms  {(1<<20)/ms/1e3:.1f} Mrays/s")
# C4: plane_sphere path trace, 256x256, 256 spp
g4 = np.load("tests/golden/c4_scene.npz"); cam4 = g4["cam"]
a4 = accel.Accel.bind().build(g4["tris"], accel.PREC_F64)
pf = accel.make_path_frame(cam4[:16], cam4[16], bool(cam4[17]), 256, 256, spp=256, max_vertices=10, seed=11, kd=1.0, Le=1.0)
a4.render_pathtrace(pf); rgb4, s4 = a4.render_pathtrace(pf)
print(f"C4 pathtrace 256x256x256spp: {s4.ms_total:.2f} ms, {s4.nrays} rays -> {s4.nrays/s4.ms_total/1e3:.1f} Mrays/s mean {rgb4.mean():.4f}")
