"""Golden frame for the material-texture clause of the AO transport (ambientocclusion.c:393-401), rendered by the COMPILED
REFERENCE (oracle/_ref) from tests/scenes/textured_quads.rib with a seeded float RGBA image attached as the material texture of every
geom (oracle/ref/ref_shim.c: lref_set_frame_texture).  Build container only:   python tests/golden/make_texture_golden.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as ol  # noqa: E402

tex = ol.test_texture()
rib = os.path.join(ROOT, "tests", "scenes", "textured_quads.rib")
with tempfile.TemporaryDirectory() as tmp:
    rgb, _, nrays = ol.run_oracle_rib(rib, os.path.join(tmp, "f.bin"), scene=os.path.join(tmp, "s.bin"), width=120, height=90, gather=16,
                                      texture=tex, attr=os.path.join(tmp, "a.bin"))
    tris, geom, cam, _ = ol.read_scene(os.path.join(tmp, "s.bin"))
    st, has_st = ol.read_attr(os.path.join(tmp, "a.bin"), len(tris))
np.savez_compressed(os.path.join(HERE, "textured_quads.npz"), rgb=rgb, nrays=np.uint64(nrays), tris=tris, cam=cam, st=st, has_st=has_st, tex=tex)
print("textured_quads.npz", rgb.shape, nrays, len(tris), has_st)
