"""CPU analysis (no GPU): traversal steps of the C3 occlusion rays when two levels of the reference's binary tree are walked per step.
python scripts/study/wide_node_study.py [nrays]   -- builds scripts/study/_wide.so with gcc, uses the oracle's tree."""
import ctypes as C, os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from lucille_b200 import scenes
import bench

so = os.path.join(HERE, "_wide.so")
subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", os.path.join(HERE, "wide_node_study.c"), "-o", so, "-lm"])
lib = C.CDLL(so)
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
tris = scenes.triangle_soup(1_000_000, scenes.SEED_C3)
t = ol.Oracle().build(tris)
order = t.triorder()
P, n = bench.primary_points(t.intersect_f32, tris[order])
rays8 = scenes.ao_rays(P, n, 8, 8, scenes.SEED_C3)
sel = np.random.default_rng(1).choice(len(rays8), nr, replace=False)
rays6 = np.ascontiguousarray(scenes.rays_f32_to_f64(rays8[np.sort(sel)]))
nodes = np.ascontiguousarray(t.nodes())
post = np.ascontiguousarray(tris[order].reshape(-1, 9).astype(np.float32).astype(np.float64))
out = np.zeros(6)
lib.wide_node_study(nodes.ctypes.data_as(C.c_void_p), post.ctypes.data_as(C.c_void_p), rays6.ctypes.data_as(C.c_void_p), C.c_uint64(nr),
                    out.ctypes.data_as(C.c_void_p))
_, cnt = t.occluded_f64(rays6, counters=True)
print(f"oracle counters on the same rays: {cnt['ninner']/nr:.2f} inner, {cnt['nleaf']/nr:.2f} leaf, {cnt['ntris']/nr:.2f} triangle tests per ray")
print(f"{nr} of the C3 rays, occluded fraction {out[5]:.4f}")
print(f"binary walk (reference order): {out[0]:.2f} inner visits, {out[1]:.2f} leaf visits, {out[2]:.2f} triangle tests per ray")
print(f"two levels per step, same leaves: {out[3]:.2f} steps, {out[4]:.2f} child boxes tested per ray "
      f"({out[3]/out[0]:.2f} of the steps, {out[4]/(2*out[0]):.2f} of the box tests)")
