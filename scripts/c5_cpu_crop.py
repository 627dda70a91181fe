"""CPU side of BASELINE configs[4]: the COMPILED REFERENCE on the 10 M-triangle soup, timed on the central 256x256
crop of the 4096x4096 camera (one primary ray per pixel + 64 AO rays per hit, closest-hit ri_bvh_intersect incl. state build, all host
threads) and extrapolated to the full frame (16 sub-samples x (1 + 64 per hit) rays per pixel).  CPU only:  python scripts/c5_cpu_crop.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from lucille_b200 import scenes

ntris = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cores = os.cpu_count() or 1
tris = scenes.triangle_soup(ntris, scenes.SEED_C5)
t0 = time.perf_counter()
ref = ol.Reference().build(tris)
print(f"reference ri_bvh_build of {ntris} triangles: {time.perf_counter() - t0:.1f} s (one thread)", flush=True)
full = scenes.pinhole_rays(4096, 4096).reshape(4096, 4096, 8)
crop = np.ascontiguousarray(full[1920:2176, 1920:2176].reshape(-1, 8))
del full
rays6 = scenes.rays_f32_to_f64(crop)
hits, sec_p = ref.intersect(rays6, nthreads=cores, want_hits=True)
m = hits["hit"] == 1
print(f"primary: {len(rays6)} rays in {sec_p:.2f} s = {len(rays6)/sec_p/1e6:.2f} Mrays/s on {cores} threads, hit fraction {m.mean():.3f}", flush=True)
orc = ol.Oracle().build(tris)
st = orc.state_build(rays6, orc.intersect_f64(rays6))
P, N = st["P"][m][:, :3], st["Ns"][m][:, :3]
ao = scenes.rays_f32_to_f64(scenes.ao_rays(P, N, 8, 8, scenes.SEED_C5))
_, sec_a = ref.intersect(ao, nthreads=cores, want_hits=False)
print(f"AO: {len(ao)} rays in {sec_a:.2f} s = {len(ao)/sec_a/1e6:.2f} Mrays/s", flush=True)
crop_rays, crop_sec = len(rays6) + len(ao), sec_p + sec_a
frame_rays = 4096 * 4096 * 16 * (1 + 64 * m.mean())
print(f"crop: {crop_rays} rays in {crop_sec:.2f} s = {crop_rays/crop_sec/1e6:.2f} Mrays/s; full frame = {frame_rays/1e9:.2f} G rays "
      f"-> {frame_rays/(crop_rays/crop_sec)/3600:.2f} h extrapolated on {cores} threads", flush=True)
