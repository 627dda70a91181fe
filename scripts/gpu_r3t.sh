mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r3t_tests.txt
python - <<'PY' 2>&1 | grep -v "lucille\]" | tee gpurun_out/r3t_k6.txt
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from lucille_b200 import accel, scenes
import bench
tris = scenes.triangle_soup(10_000_000, scenes.SEED_C5)
a = accel.Accel.bind().build(tris, accel.PREC_F32)
bench.NPOINTS = 262144
P, n = bench.primary_points(a.intersect, tris[a.triorder()])
pts = np.ascontiguousarray(np.concatenate([P[:262144], n[:262144]], axis=1))
d_pts = torch.from_numpy(pts).cuda(); d_cnt = torch.empty(len(pts), dtype=torch.int32, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
res = {}
for k6 in ("0", "1", "auto"):
    if k6 == "auto": os.environ.pop("B200_K6", None)
    else: os.environ["B200_K6"] = k6
    for _ in range(2): a.occlusion_points_dev(d_pts, len(pts), d_cnt, 8, 8, 5, stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): a.occlusion_points_dev(d_pts, len(pts), d_cnt, 8, 8, 5, stream=st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    res[k6] = d_cnt.cpu().numpy().copy()
    print(f"C5 point entry (16 Mi rays, generation + K6 + traversal) B200_K6={k6}: {ms:.2f} ms = {len(pts) * 64 / ms / 1e3:.1f} Mrays/s", flush=True)
print("identical counts:", bool(np.array_equal(res["0"], res["1"]) and np.array_equal(res["0"], res["auto"])))
c2w = np.eye(4); c2w[3, :3] = (0.5, 0.5, -2.0)
import math
fr = accel.make_frame(c2w.reshape(16), 1.0 / math.tan(math.radians(40.0) / 2), False, 2048, 2048, 1, 1, 64, rng_mode=1, seed=5, precision=accel.PREC_F32)
for k6 in ("0", "auto"):
    if k6 == "auto": os.environ.pop("B200_K6", None)
    else: os.environ["B200_K6"] = k6
    a.render_ao(fr); rgb, s = a.render_ao(fr)
    print(f"C5 frame 2048^2 B200_K6={k6}: {s.ms_total:.1f} ms, {s.nrays} rays = {s.nrays / s.ms_total / 1e3:.1f} Mrays/s, mean {rgb.mean():.5f}", flush=True)
PY
