"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck): tiny soup, few rays, tiny frames."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from lucille_b200 import accel, scenes
tris = scenes.triangle_soup(3000, 5)
rays8 = scenes.pinhole_rays(48, 48)
rays6 = scenes.rays_f32_to_f64(rays8)
for flag in (accel.BUILD_HOST, accel.BUILD_DEVICE):
    a = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.PREC_F64 | flag)
    h32 = a.intersect(rays8); h64 = a.intersect(rays6)
    o32 = a.occluded(rays8); o64 = a.occluded(rays6)
    st = a.state(rays6, h64)
    print(flag, int((h32["prim"] != 0xFFFFFFFF).sum()), int(h64["hit"].sum()), int(o32.sum()), int(o64.sum()))
g = np.load(os.path.join("tests", "golden", "c1_scene.npz")); cam = g["cam"]
c1 = accel.Accel.bind().build(g["tris"], accel.PREC_F64 | accel.PREC_F32)
for prec in (accel.PREC_F64, accel.PREC_F32):
    for env in ("0", "1"):
        os.environ["B200_FUSED_AO_TEST"] = env
        fr = accel.make_frame(cam[:16], cam[16], bool(cam[17]), 32, 24, 2, 2, gather_nsamples=16, precision=prec)
        rgb, s = c1.render_ao(fr)
        print("frame", prec, env, float(rgb.mean()), s.nrays)
print("done")
