"""Times the pooled closest-hit kernel on the C3 batch (argument: number of AO points, default 65536 = 4 Mi rays)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucille_b200 import accel, scenes
import bench
npoints = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tris = scenes.triangle_soup(1_000_000, scenes.SEED_C3)
a = accel.Accel.bind().build(tris, accel.PREC_F32)
def dev_intersect(r8):            # ONE launch (the host-buffer call would ramp through several chunk sizes: awkward under ncu -s)
    dr = torch.from_numpy(r8).cuda(); dh = torch.empty((len(r8), 4), dtype=torch.float32, device="cuda")
    a.intersect_dev(dr, len(r8), dh); torch.cuda.synchronize()
    return dh.cpu().numpy().view(accel.HIT32_DTYPE).reshape(-1)
P, n = bench.primary_points(dev_intersect, tris[a.triorder()])
rays = scenes.ao_rays(P[:npoints], n[:npoints], 8, 8, scenes.SEED_C3)
nr = len(rays)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
d = torch.from_numpy(rays).cuda(); hits = torch.empty((nr, 4), dtype=torch.float32, device="cuda")
for _ in range(2): a.intersect_dev(d, nr, hits, st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): a.intersect_dev(d, nr, hits, st.cuda_stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print({k: v for k, v in os.environ.items() if k.startswith("B200_")}, f"{ms:.3f} ms {nr/ms/1e3:.1f} Mrays/s", float(hits[:, 0].clamp(max=10).sum().item()))
