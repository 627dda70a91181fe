"""Times the pooled occlusion kernel on the full C3 batch; knobs come from the environment (one process per setting)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucille_b200 import accel, scenes
import bench
npoints = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
tris = scenes.triangle_soup(1_000_000, scenes.SEED_C3)
a = accel.Accel.bind().build(tris, accel.PREC_F32)
P, n = bench.primary_points(a.intersect, tris[a.triorder()])
rays = scenes.ao_rays(P[:npoints], n[:npoints], 8, 8, scenes.SEED_C3)
nr = len(rays)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
d = torch.from_numpy(rays).cuda(); occ = torch.empty(nr, dtype=torch.uint8, device="cuda")
for _ in range(2): a.occluded_dev(d, nr, occ, st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): a.occluded_dev(d, nr, occ, st.cuda_stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print({k: v for k, v in os.environ.items() if k.startswith("B200_")}, f"{ms:.3f} ms {nr/ms/1e3:.1f} Mrays/s", int(occ.sum().item()))
