"""Driver for the round-2 ncu captures: python scripts/prof_r02.py <what> [npoints]
  c3      1 M-triangle soup (L2-resident records): fp32 occlusion + closest, fp64 occlusion + closest on the first `npoints` AO points
  c5      10 M-triangle soup (1.2 GB of records, beyond L2): fp32 occlusion + closest
Each kernel is launched three times, the third inside cudaProfilerStart/Stop: ncu --profile-from-start off -k regex:<name> -c 1."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucille_b200 import accel, scenes
import bench

what = sys.argv[1] if len(sys.argv) > 1 else "c3"
npoints = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
ntris, seed = (1_000_000, scenes.SEED_C3) if what == "c3" else (10_000_000, scenes.SEED_C5)
prec = accel.PREC_F32 | (accel.PREC_F64 if what == "c3" else 0)
tris = scenes.triangle_soup(ntris, seed)
a = accel.Accel.bind().build(tris, prec)
bench.NPOINTS = npoints
P, n = bench.primary_points(a.intersect, tris[a.triorder()])
rays = a.ao_point_rays(np.concatenate([P[:npoints], n[:npoints]], axis=1), 8, 8, seed)
nr = len(rays)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
d32 = torch.from_numpy(rays).cuda()
occ = torch.empty(nr, dtype=torch.uint8, device="cuda")
h32 = torch.empty((nr, 4), dtype=torch.float32, device="cuda")
if what == "c3":
    d64 = torch.from_numpy(scenes.rays_f32_to_f64(rays)).cuda()
    h64 = torch.empty((nr, 4), dtype=torch.float64, device="cuda")
# two warm launches of every kernel, then ONE launch of each inside the profiler range (ncu --profile-from-start off): the set-up above
# launches the same kernels on other batches (primary_points), which a launch-count filter would pick up instead
for it in range(3):
    if it == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    a.occluded_dev(d32, nr, occ, st.cuda_stream)
    a.intersect_dev(d32, nr, h32, st.cuda_stream)
    if what == "c3":
        a.occluded_dev(d64, nr, occ, st.cuda_stream, f64=True)            # both record sets resident: hybrid.cuh
        os.environ["B200_HYBRID"] = "0"
        a.occluded_dev(d64, nr, occ, st.cuda_stream, f64=True)            # the plain double kernel
        os.environ.pop("B200_HYBRID")
        a.intersect_dev(d64, nr, h64, st.cuda_stream, f64=True)           # hybrid closest hit
        os.environ["B200_HYBRID_CLOSEST"] = "0"
        a.intersect_dev(d64, nr, h64, st.cuda_stream, f64=True)           # the double kernel
        os.environ.pop("B200_HYBRID_CLOSEST")
torch.cuda.synchronize()
torch.cuda.profiler.stop()
cnt = a.count(rays[: 1 << 20], anyhit=True)
print("done", what, nr, "rays; reference-order counters of the first Mi rays:", cnt)
