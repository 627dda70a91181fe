"""Golden digests of the byte stream lucille's socket display driver (display/sockdrv.c) sends for committed frames: captured from the
COMPILED REFERENCE's sock_dd_open / sock_dd_write / sock_dd_close talking to a listener inside the process (oracle/ref/ref_shim.c:
lref_sockdrv_stream).  Build container only:   python tests/golden/make_sockdrv_golden.py"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as ol  # noqa: E402
from lucille_b200 import accel  # noqa: E402

ref = ol.Reference()
frames = dict(ol.hdr_cases())
frames["c1"] = np.load(os.path.join(HERE, "c1_frame_160x120.npz"))["rgb"]
frames["sunsky"] = np.load(os.path.join(HERE, "sunsky.npz"))["frame_rgb"]
out = {}
for name, rgb in frames.items():
    h, w = rgb.shape[:2]
    pix = accel.frame_pixels(accel.make_frame(np.eye(4).reshape(16), 1.0, False, w, h, 1, 1))
    disp = (pix & 0xFFFF) | ((np.uint32(h - 1) - (pix >> 16)) << 16)
    data = ref.sockdrv_stream(rgb, disp)
    out[name + "_size"], out[name + "_sha256"] = len(data), hashlib.sha256(data).hexdigest()
    print(name, rgb.shape, len(data))
np.savez_compressed(os.path.join(HERE, "sockdrv.npz"), **out)
