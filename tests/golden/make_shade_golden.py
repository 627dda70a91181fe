"""Golden vectors for the two shading-language callers of ri_raytrace: the trace() shadeop (shader.c:895-976) and
the light samples of next_lightsource() (shader.c:1116-1186, 1236-1310) of the COMPILED REFERENCE (oracle/_ref), called through
oracle/ref/ref_shim.c (lref_shade_trace with a capturing shader procedure on every geom, lref_light_samples).
Build container only:   python tests/golden/make_shade_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as ol  # noqa: E402
from lucille_b200 import scenes  # noqa: E402

SIZES, SEED_T, ATTR_SEED, NPAIRS = [1500, 900, 1300, 1100, 1200], 21, 4, 1500
NTRIS_L, SEED_L, NPOINTS = 20000, 9, 250
LIGHT_CASES = ((48, 1.2), (27, 1.5707963267948966), (5, 0.6))          # (narealight_rays, cone angle)

ref, orc = ol.Reference(), ol.Oracle()
env = ol.test_texture(32, 32, 5)
out = dict(env=env, sizes=np.array(SIZES), seed_t=SEED_T, attr_seed=ATTR_SEED, ntris_l=NTRIS_L, seed_l=SEED_L,
           light_cases=np.array(LIGHT_CASES))

# trace(): five geoms with colours / shared and unshared texture coordinates / two-sided geometry, IBL light with the map
tris = scenes.triangle_soup(sum(SIZES), SEED_T)
col, st, geom_flags, has_col, has_st, inside = ol.attribute_case(len(tris), SIZES, ATTR_SEED)
rs = ref.build(tris, geom_sizes=SIZES)
rs.set_attributes(col, st, geom_flags)
rs.set_envmap(env)
rng = np.random.default_rng(4)
P = rng.uniform(-0.3, 1.3, (NPAIRS, 3))
R = rng.uniform(0.0, 1.0, (NPAIRS, 3)) - P
R *= rng.uniform(0.3, 2.5, (NPAIRS, 1))                                 # trace() does not normalise R
pr = np.concatenate([P, R], axis=1)
want = rs.shade_trace(pr)
out["pr"] = pr
for f in ("Cs", "P", "N", "Ng", "dPdu", "dPdv", "I", "dst", "s", "t", "called"):
    out["trace_" + f] = want[f]

# next_lightsource(): shading points on a 20 K-triangle soup
tris = scenes.triangle_soup(NTRIS_L, SEED_L)
rs, ot = ref.build(tris), orc.build(tris)
rs.set_envmap(env)
rays = scenes.rays_f32_to_f64(scenes.pinhole_rays(64, 64))
hits = ot.intersect_f64(rays)
stt = ot.state_build(rays, hits)
m = hits["hit"] == 1
pts = np.concatenate([stt["P"][m][:, :3], stt["Ns"][m][:, :3]], axis=1)[:NPOINTS]
out["points"] = pts
for i, (ns, angle) in enumerate(LIGHT_CASES):
    L, Cl, cnt = rs.light_samples(int(ns), float(angle), pts, maxm=64)
    k = int(cnt.max())
    out[f"light{i}_L"], out[f"light{i}_Cl"], out[f"light{i}_count"] = L[:, :k], Cl[:, :k], cnt
np.savez_compressed(os.path.join(HERE, "shade_callers.npz"), **out)
print("shade_callers.npz", int(want["called"].sum()), "of", NPAIRS, "pairs hit;", {i: int(out[f"light{i}_count"].sum()) for i in range(3)})
