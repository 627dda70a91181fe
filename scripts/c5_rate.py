"""Occlusion / closest-hit rates on the 10 M-triangle soup of BASELINE configs[4] (one GPU, 4 Mi AO rays from primary hits)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucille_b200 import accel, scenes
import bench
tris = scenes.triangle_soup(10_000_000, 0xB2000005)
a = accel.Accel.bind().build(tris, accel.PREC_F32)
i = a.info()
print(f"build {i.build_seconds:.2f}s, {i.ninner} inner nodes, depth {i.max_depth}, {i.device_bytes/1e6:.0f} MB on device", flush=True)
bench.NPOINTS = 65536
P, n = bench.primary_points(a.intersect, tris[a.triorder()])
rays = scenes.ao_rays(P[:65536], n[:65536], 8, 8, 0xB2000005)
nr = len(rays)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
d = torch.from_numpy(rays).cuda(); occ = torch.empty(nr, dtype=torch.uint8, device="cuda"); hits = torch.empty((nr, 4), dtype=torch.float32, device="cuda")
def ev(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = ev(lambda: a.occluded_dev(d, nr, occ, st.cuda_stream)); print(f"C5 soup anyhit  f32 {ms:.3f} ms  {nr/ms/1e3:.1f} Mrays/s  occluded {float(occ.float().mean()):.3f}")
ms = ev(lambda: a.intersect_dev(d, nr, hits, st.cuda_stream)); print(f"C5 soup closest f32 {ms:.3f} ms  {nr/ms/1e3:.1f} Mrays/s")
pr = torch.from_numpy(scenes.pinhole_rays(1024, 1024)).cuda(); ph = torch.empty((1 << 20, 4), dtype=torch.float32, device="cuda")
ms = ev(lambda: a.intersect_dev(pr, 1 << 20, ph, st.cuda_stream)); print(f"C5 soup primary closest f32 {ms:.3f} ms  {(1<<20)/ms/1e3:.1f} Mrays/s")
print("counters anyhit", a.count(rays[:1 << 20], anyhit=True))
