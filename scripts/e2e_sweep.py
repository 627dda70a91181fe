"""e2e throughput of ri_b200_occluded_batch_f32 for the chunk size given by B200_CHUNK (host pinned buffers)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucille_b200 import accel, scenes
import bench
tris = scenes.triangle_soup(1_000_000, scenes.SEED_C3)
a = accel.Accel.bind().build(tris, accel.PREC_F32)
P, n = bench.primary_points(a.intersect, tris[a.triorder()])
rays = scenes.ao_rays(P, n, 8, 8, scenes.SEED_C3)
h = torch.empty((len(rays), 8), dtype=torch.float32, pin_memory=True); h.numpy()[:] = rays
o = torch.empty(len(rays), dtype=torch.uint8, pin_memory=True)
rh, oh = h.numpy(), o.numpy()
import ctypes
for _ in range(2): a.lib.ri_b200_occluded_batch_f32(a.data, accel._ptr(rh), len(rays), accel._ptr(oh))
if os.environ.get('E2E_SKIP'): ctypes.CDLL(None).setenv(b'B200_E2E_SKIP', os.environ['E2E_SKIP'].encode(), 1)
if os.environ.get('E2E_TRACE'):
    ctypes.CDLL(None).setenv(b'B200_E2E_TRACE', b'1', 1)
    a.lib.ri_b200_occluded_batch_f32(a.data, accel._ptr(rh), len(rays), accel._ptr(oh))
    ctypes.CDLL(None).unsetenv(b'B200_E2E_TRACE')
t0 = time.perf_counter()
for _ in range(5): a.lib.ri_b200_occluded_batch_f32(a.data, accel._ptr(rh), len(rays), accel._ptr(oh))
dt = (time.perf_counter() - t0) / 5
print({k: v for k, v in os.environ.items() if k.startswith("B200_") or k.startswith("E2E")}, f" {dt*1e3:.2f} ms  {len(rays)/dt/1e6:.1f} Mrays/s e2e")
