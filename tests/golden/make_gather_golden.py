"""Golden vectors for the per-point hemisphere gathers: the occlusion() shadeop (shader.c:680-768),
ri_ibl_sample_cosweight (ibl.c:53-228) and ri_domelight_sample (ibl.c:231-389) of the COMPILED REFERENCE (oracle/_ref), called once per
shading point through oracle/ref/ref_shim.c: lref_point_gather.  Build container only:   python tests/golden/make_gather_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as ol  # noqa: E402
from lucille_b200 import scenes  # noqa: E402

NTRIS, SEED, NPOINTS = 20000, 9, 800
COL, INTENSITY = (0.9, 0.5, 0.25), 2.5
CASES = ((0, 48), (0, 7), (1, 48), (2, 27))                  # (kind, nsamples)

tris = scenes.triangle_soup(NTRIS, SEED)
ref, orc = ol.Reference(), ol.Oracle()
rs, ot = ref.build(tris), orc.build(tris)
env = ol.test_texture(32, 32, 5)
rs.set_envmap(env)
rays = scenes.rays_f32_to_f64(scenes.pinhole_rays(64, 64))
hits = ot.intersect_f64(rays)
st = ot.state_build(rays, hits)
m = hits["hit"] == 1
pts = np.concatenate([st["P"][m][:, :3], st["Ns"][m][:, :3]], axis=1)[:NPOINTS]
out = dict(ntris=NTRIS, seed=SEED, points=pts, env=env, col=np.array(COL), intensity=INTENSITY, cases=np.array(CASES))
for kind, ns in CASES:
    out[f"k{kind}_n{ns}"] = rs.point_gather(kind, ns, pts, COL, INTENSITY)
# the quasi-Monte Carlo branches (Option "use_qmc"): per-point instance numbers (inray->i), dimension inray->d
QMC_CASES = ((1, 48, 0), (2, 31, 3))                          # (kind, nsamples, dim)
inst = (np.arange(len(pts)) * 7919 % 5000).astype(np.int32)
out["qmc_cases"], out["qmc_instance"] = np.array(QMC_CASES), inst
for kind, ns, dim in QMC_CASES:
    out[f"q{kind}_n{ns}_d{dim}"] = rs.point_gather_qmc(kind, ns, pts, inst, dim, COL, INTENSITY)
np.savez_compressed(os.path.join(HERE, "point_gathers.npz"), **out)
print("point_gathers.npz", len(pts), {k: float(v.mean()) for k, v in out.items() if k.startswith("k")})
