"""Golden digests of the .hdr files the COMPILED REFERENCE's display driver (display/hdrdrv.c + imageio/rgbe.c) writes for the
committed frames and the seeded test framebuffers.  Build container only:   python tests/golden/make_hdr_golden.py"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as ol  # noqa: E402

ref = ol.Reference()
frames = dict(ol.hdr_cases())
frames["c1"] = np.load(os.path.join(HERE, "c1_frame_160x120.npz"))["rgb"]
frames["sunsky"] = np.load(os.path.join(HERE, "sunsky.npz"))["frame_rgb"]
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for name, rgb in frames.items():
        data = ref.hdr_file(rgb, os.path.join(tmp, name + ".hdr"))
        out[name + "_size"] = len(data)
        out[name + "_sha256"] = hashlib.sha256(data).hexdigest()
np.savez_compressed(os.path.join(HERE, "hdr.npz"), **out)
print(out)
