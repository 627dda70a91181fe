"""Golden data for the sun-sky gather (row a12), produced by RUNNING THE COMPILED REFERENCE (oracle/_ref) in the build
container:

    python tests/golden/make_sunsky_golden.py        ->  tests/golden/sunsky.npz

* tables the reference's sky lookup reads: S0/S1/S2Amplitudes (data symbols of the compiled sunsky.c), the 81x3 CIE
  colour-matching table (a function-local static of specrend.c:387-414, read from the source text as DATA) and the
  chromaticities of specrend.h's CIEsystem;
* ri_sunsky_init() parameter records and ri_sunsky_get_sky_rgb() outputs for seeded direction batches;
* one frame of ambient_occlusion.rib with `AreaLightSource "sunsky"` added, rendered by the reference with one thread,
  with the light block the transport saw.
"""
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as ol  # noqa: E402

REF_SRC = os.environ.get("LUCILLE_REF", "/root/reference")

SKY_CASES = [dict(latitude=35.39, longitude=139.44, sm=9.0, jd=20, tod=10.5, turbidity=2.0),      # lightsource.c:287-300 defaults
             dict(latitude=48.85, longitude=2.35, sm=1.0, jd=172, tod=15.25, turbidity=4.5),
             dict(latitude=-33.87, longitude=151.2, sm=10.0, jd=300, tod=7.0, turbidity=7.0)]


def cie_table():
    src = open(os.path.join(REF_SRC, "src", "render", "specrend.c")).read()
    body = src[src.rindex("static float cie_colour_match[81][3]"):]
    body = body[: body.index("};")]
    nums = re.findall(r"\{\s*([0-9.]+)\s*,\s*([0-9.]+)\s*,\s*([0-9.]+)\s*\}", body)
    t = np.array(nums, dtype=np.float32)
    assert t.shape == (81, 3)
    return t


def cie_system():
    hdr = open(os.path.join(REF_SRC, "src", "render", "specrend.h")).read()
    m = re.search(r"CIEsystem\s*=\s*\{\s*\"CIE\"\s*,([^}]*)\}", hdr)
    vals = [v.strip() for v in m.group(1).split(",")]
    e = re.search(r"#define\s+IlluminantE\s+([0-9.]+)\s*,\s*([0-9.]+)", hdr)
    out = [float(v) for v in vals[:6]] + [float(e.group(1)), float(e.group(2))]
    assert vals[6] == "IlluminantE"
    return np.array(out, dtype=np.float32)


def main():
    ref = ol.Reference()
    out = dict(S0=ref.table("S0Amplitudes", 41), S1=ref.table("S1Amplitudes", 41), S2=ref.table("S2Amplitudes", 41),
               cie=cie_table(), cs=cie_system())
    for k, case in enumerate(SKY_CASES):
        dirs = ol.sky_dirs(4096, 100 + k)
        rgb, rec = ref.sunsky_eval(dirs, **case)
        out[f"sky{k}_case"] = np.array([case[x] for x in ("latitude", "longitude", "sm", "jd", "tod", "turbidity")], dtype=np.float64)
        out[f"sky{k}_rec"] = rec
        out[f"sky{k}_rgb"] = rgb
    out["nsky"] = len(SKY_CASES)

    rib = os.path.join(ol.REF_DIR, "scenes", "ambient_occlusion.rib")
    text = open(rib).read()
    assert "WorldBegin" in text
    # an area light opens an "arealight block" that swallows the geometry up to the next AttributeEnd (lightsource.c:254-260,
    # attribute.c:159, polygon.c:1070-1083): keep it inside its own, empty attribute block
    text = text.replace("WorldBegin", 'WorldBegin\nAttributeBegin\nAreaLightSource "sunsky" 1 "turbidity" [3.0] "time_of_day" [14.0]\nAttributeEnd\n', 1)
    with tempfile.TemporaryDirectory(dir=os.path.dirname(rib)) as tmp:
        path = os.path.join(tmp, "sunsky_ao.rib")
        open(path, "w").write(text)
        rgb, _, nrays = ol.run_oracle_rib(path, os.path.join(tmp, "f.bin"), scene=os.path.join(tmp, "s.bin"),
                                          sunsky=os.path.join(tmp, "k.bin"), width=120, height=90, pixelsamples=2)
        tris, geom, cam, _ = ol.read_scene(os.path.join(tmp, "s.bin"))
        blk = np.fromfile(os.path.join(tmp, "k.bin"), dtype=np.float64)
    assert len(blk) == 45
    out.update(frame_rgb=rgb, frame_nrays=np.uint64(nrays), frame_cam=cam, frame_block=blk, frame_tris=tris)
    np.savez_compressed(os.path.join(HERE, "sunsky.npz"), **out)
    print("sunsky.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
