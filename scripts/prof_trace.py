"""Small driver for ncu: builds the C3 scene, generates a 4Mi-ray AO batch, launches any-hit and closest-hit a few times."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucille_b200 import accel, scenes
import bench
npoints = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tris = scenes.triangle_soup(1_000_000, scenes.SEED_C3)
a = accel.Accel.bind().build(tris, accel.PREC_F32)
P, n = bench.primary_points(a.intersect, tris[a.triorder()])
rays = scenes.ao_rays(P[:npoints], n[:npoints], 8, 8, scenes.SEED_C3)
nr = len(rays)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
d = torch.from_numpy(rays).cuda(); occ = torch.empty(nr, dtype=torch.uint8, device="cuda"); hits = torch.empty((nr, 4), dtype=torch.float32, device="cuda")
for _ in range(3):
    a.occluded_dev(d, nr, occ, st.cuda_stream)
    a.intersect_dev(d, nr, hits, st.cuda_stream)
torch.cuda.synchronize()
print("done", nr)
