"""Rates of every frame transport and of the per-point gathers on the 1 M-triangle soup of BASELINE configs[2] (one GPU):
python scripts/transport_rates.py [ntris] [res].  Device time from the frame statistics (CUDA events inside the call)."""
import math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lucille_b200 import accel, scenes
import oracle_lib as ol

ntris = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
tris = scenes.triangle_soup(ntris, scenes.SEED_C3)
a = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.PREC_F64)
c2w = np.eye(4); c2w[3, :3] = (0.5, 0.5, -2.0)
flen = 1.0 / math.tan(math.radians(40.0) / 2)
g = np.load(os.path.join(ROOT, "tests", "golden", "sunsky.npz"))
blk = ol.sunsky_block(g["frame_block"], g)
import ctypes
sky = accel.Sunsky()
ctypes.memmove(ctypes.byref(sky), ctypes.byref(blk), ctypes.sizeof(blk))      # orc_sunsky_t and ri_b200_sunsky_t share one layout
env = ol.test_texture(256, 256, 5)


def frame(prec, spp=2, gather=64, rng=1):
    return accel.make_frame(c2w.reshape(16), flen, False, res, res, spp, spp, gather, rng_mode=rng, seed=5, precision=prec)


def run(label, fn):
    fn()
    rgb, st = fn()
    print(f"{label:34s} {st.nrays/1e6:9.1f} Mrays {st.ms_total:9.2f} ms  {st.nrays/st.ms_total/1e3:8.1f} Mrays/s   "
          f"(eye {st.ms_primary:.2f} ms, gather {st.ms_ao:.2f} ms)", flush=True)


for prec, pn in ((accel.PREC_F32, "f32"), (accel.PREC_F64, "f64")):
    run(f"AO 8x8 {pn}", lambda: a.render_ao(frame(prec)))
    run(f"sun-sky 8x8 + 1 sun {pn}", lambda: a.render_sunsky(frame(prec), sky))
    run(f"dirt map 4x4 {pn}", lambda: a.render_dirtmap(frame(prec)))
run("whitted (f64)", lambda: a.render_whitted(frame(accel.PREC_F64), env))
run("hit mask (f64)", lambda: a.render_sample(frame(accel.PREC_F64)))
run("AO 8x8 f64, MT19937 stream", lambda: a.render_ao(frame(accel.PREC_F64, rng=0)))
rays = scenes.rays_f32_to_f64(scenes.pinhole_rays(512, 512))
h = a.intersect(rays); s = a.state(rays, h); m = h["hit"] == 1
pts = np.concatenate([s["P"][m][:, :3], s["Ns"][m][:, :3]], axis=1)
for kind, name in ((accel.GATHER_OCCLUSION, "occlusion() shadeop"), (accel.GATHER_IBL, "ibl cosweight"), (accel.GATHER_DOME, "dome light")):
    a.mt_prepare(2 * 48 * len(pts))                                          # the stream's state table: once per (accelerator, seed)
    a.gather_points(kind, 48, pts[:1000], env)
    t0 = time.perf_counter(); out, n = a.gather_points(kind, 48, pts, env); dt = time.perf_counter() - t0
    print(f"gather {name:26s} {len(pts)} points x 48: {n/1e6:7.1f} Mrays {dt*1e3:9.2f} ms  {n/dt/1e6:8.1f} Mrays/s (host call incl. copies)", flush=True)
