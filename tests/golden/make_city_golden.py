"""Golden frame of a RIB-scale architectural scene (C6, this repo's addition to the parity cases): scenes.box_city(36, 36) -- 9072
triangles, shared vertices, coplanar faces, zero-extent boxes -- written as a RIB at world scale (24 units across, camera of
ambient_occlusion.rib), parsed and transformed by the reference's own Ri layer (vertices leave it as doubles that are NOT fp32
numbers), and rendered by the COMPILED REFERENCE with its AO transport on one thread.  The scene dump and the float framebuffer are
what tests/test_gpu_parity.py::test_city_frame_through_the_hybrid_path compares the device frame with (gather rays through
csrc/hybrid.cuh, the filter's own records).  Build container only:   python tests/golden/make_city_golden.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as ol  # noqa: E402
from lucille_b200 import scenes  # noqa: E402

W, H, PS, GATHER = 200, 150, 2, 16
tris = scenes.box_city(36, 36, 17)
tris = (tris - np.array([0.5, 0.0, 0.5])) * np.array([24.0, 9.0, 24.0])            # 24 x 24 units on the ground, towers up to 6 high
cam = ("ConcatTransform [0.994530 0.008385 -0.104111 0.000000 0.052799 0.819679 0.570385 0.000000 0.090120 -0.572762 0.814753 0.000000 "
       "-0.000009 -0.000015 -15.529361 1.000000 ]")
with tempfile.TemporaryDirectory() as tmp:
    rib = os.path.join(tmp, "city.rib")
    with open(rib, "w") as f:
        f.write(f'Display "city.hdr" "file" "rgb"\nPixelSamples {PS} {PS}\nProjection "perspective" "fov" [45.0]\nOrientation "rh"\n{cam}\nWorldBegin\n')
        f.write("Translate 1.3 0.0 -0.7\nRotate 11.0 0 1 0\nScale 1.01 1.01 1.01\n")   # world coordinates = doubles that are not fp32 numbers
        for g0 in range(0, len(tris), 1512):                                         # the geoms
            part = tris[g0:g0 + 1512]
            n = len(part)
            f.write("AttributeBegin\nPointsPolygons [" + " ".join(["3"] * n) + "] [" + " ".join(str(i) for i in range(3 * n)) + '] "P" [' +
                    " ".join(repr(float(x)) for x in part.reshape(-1)) + "]\nAttributeEnd\n")
        f.write("WorldEnd\n")
    rgb, sec, nrays = ol.run_oracle_rib(rib, os.path.join(tmp, "f.bin"), scene=os.path.join(tmp, "s.bin"), width=W, height=H, pixelsamples=PS, gather=GATHER)
    t, geom, camrec, nrm = ol.read_scene(os.path.join(tmp, "s.bin"))
assert len(t) == len(tris)
frac32 = float(np.mean(t.astype(np.float32).astype(np.float64) == t))
np.savez_compressed(os.path.join(HERE, "c6_city.npz"), tris=t, cam=camrec, rgb=rgb, nrays=np.uint64(nrays), width=W, height=H, ps=PS, gather=GATHER)
print("c6_city.npz:", len(t), "triangles,", nrays, "rays,", f"{sec:.2f} s on one reference thread, mean {rgb.mean():.4f}, "
      f"{frac32:.3f} of the world-space coordinates are fp32 numbers, extent", t.reshape(-1, 3).min(axis=0), t.reshape(-1, 3).max(axis=0))
