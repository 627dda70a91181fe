"""Generates the golden vectors under tests/golden/ by RUNNING THE COMPILED REFERENCE (oracle/_ref, built by
oracle/build_ref.sh from the unmodified sources under /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

The reference's own tests pin no numbers on this path, so these outputs of the reference itself
are the anchor: tests/test_oracle_golden.py checks the oracle restatement against them on any machine, including the
GPU box where /root/reference does not exist.
"""
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
from lucille_b200 import scenes  # noqa: E402


def tree_digest(nodes: np.ndarray, triorder: np.ndarray) -> str:
    h = hashlib.sha256()
    for name in ol.NODE_DTYPE.names:
        h.update(np.ascontiguousarray(nodes[name]).tobytes())
    h.update(np.ascontiguousarray(triorder).tobytes())
    return h.hexdigest()


def main():
    ref = ol.Reference()
    refs = ol.Reference(stats=True)

    # 1. per-ray vectors on a small seeded soup (closest hit through ri_bvh_intersect + ri_intersection_state_build)
    tris = scenes.triangle_soup(2000, 0xB2000001)
    rays8 = scenes.pinhole_rays(64, 64)
    rays6 = scenes.rays_f32_to_f64(rays8)
    sc = ref.build(tris)
    hits, _ = sc.intersect(rays6)
    scs = refs.build(tris)
    refs.stats_reset()
    scs.intersect(rays6)
    st = refs.stats_get()
    np.savez_compressed(os.path.join(HERE, "soup2k_rays.npz"),
                        seed=np.uint64(0xB2000001), ntris=2000, width=64, height=64,
                        hit=hits["hit"].astype(np.uint8), t=hits["t"], u=hits["u"], v=hits["v"],
                        tri_index=(hits["index"] // 3).astype(np.uint32),
                        P=hits["P"], Ng=hits["Ng"], Ns=hits["Ns"], tangent=hits["tangent"], binormal=hits["binormal"],
                        counters=np.array([st["nrays"], st["ninner"], st["nleaf"], st["ntris"], st["nhit_tris"]], dtype=np.uint64),
                        tree_digest=tree_digest(sc.nodes(), sc.triorder()), nnodes=len(sc.nodes()), max_depth=sc.max_depth())

    # 2. tree digests at larger sizes / awkward inputs
    digests = {}
    for name, t in [("soup100k_c2", scenes.triangle_soup(100000, scenes.SEED_C2)),
                    ("soup17", scenes.triangle_soup(17, 7)),
                    ("soup16", scenes.triangle_soup(16, 7)),
                    ("soup1", scenes.triangle_soup(1, 7)),
                    ("dup300", np.repeat(scenes.triangle_soup(3, 9), 100, axis=0)),       # identical boxes -> median fallback
                    ("flat500", scenes.triangle_soup(500, 11) * np.array([1.0, 1.0, 0.0]))]:  # zero-extent axis
        s = ref.build(t)
        digests[name] = tree_digest(s.nodes(), s.triorder())
        digests[name + "_nnodes"] = len(s.nodes())
    np.savez_compressed(os.path.join(HERE, "tree_digests.npz"), **digests)

    # 3. C1: ambient_occlusion.rib through the reference renderer, 1 thread, reduced resolution
    rib = os.path.join(ol.REF_DIR, "scenes", "ambient_occlusion.rib")
    with tempfile.TemporaryDirectory() as tmp:
        rgb, sec, nrays = ol.run_oracle_rib(rib, os.path.join(tmp, "f.bin"), scene=os.path.join(tmp, "s.bin"), width=160, height=120)
        tris_c1, geom, cam, _ = ol.read_scene(os.path.join(tmp, "s.bin"))
        np.savez_compressed(os.path.join(HERE, "c1_scene.npz"), tris=tris_c1, geom=geom, cam=cam)
        np.savez_compressed(os.path.join(HERE, "c1_frame_160x120.npz"), rgb=rgb, nrays=np.uint64(nrays))
        rgb2, sec2, nrays2 = ol.run_oracle_rib(rib, os.path.join(tmp, "g.bin"), width=97, height=61, pixelsamples=2, gather=16)
        np.savez_compressed(os.path.join(HERE, "c1_frame_97x61_ps2_g16.npz"), rgb=rgb2, nrays=np.uint64(nrays2))
        # full-size frame: only its digest and ray count are committed (3.7 MB of floats otherwise)
        rgb3, sec3, nrays3 = ol.run_oracle_rib(rib, os.path.join(tmp, "h.bin"))
        np.savez_compressed(os.path.join(HERE, "c1_frame_640x480_digest.npz"),
                            sha256=hashlib.sha256(rgb3.tobytes()).hexdigest(), nrays=np.uint64(nrays3),
                            mean=np.float64(rgb3.astype(np.float64).mean()), seconds_1thread=np.float64(sec3),
                            rgb_half=rgb3[::2, ::2, 0].copy())

    # 3b. C4 scene: examples/plane_sphere (1986 triangles) -- scene dump + the reference's AO render of it at reduced size
    #     (the shipped reference renders every RIB with the AO transport, render.c:800-804: "C4'" in SURVEY 6.2)
    rib4 = os.path.join(ol.REF_DIR, "scenes", "plane_sphere", "Scene_DEFAULT_Set0.rib")
    with tempfile.TemporaryDirectory() as tmp:
        rgb4, _, nrays4 = ol.run_oracle_rib(rib4, os.path.join(tmp, "f.bin"), scene=os.path.join(tmp, "s.bin"), width=96, height=96,
                                            pixelsamples=2, gather=16)
        tris4, geom4, cam4, nrm4 = ol.read_scene(os.path.join(tmp, "s.bin"))
        # cam4[25] == 1: polygon.c gave this geometry vertex normals, so the AO transport shades with interpolated Ns
        np.savez_compressed(os.path.join(HERE, "c4_scene.npz"), tris=tris4, geom=geom4, cam=cam4, normals=nrm4)
        np.savez_compressed(os.path.join(HERE, "c4_ao_frame_96x96_ps2_g16.npz"), rgb=rgb4, nrays=np.uint64(nrays4))

    # 3c. beam visibility codes from ri_beam_set + ri_bvh_intersect_beam_visibility (all four outcome classes)
    cfg = [(20000, scenes.SEED_C2, 5, 0.35, 0.02), (300, 3, 6, 0.35, 0.002), (20000, scenes.SEED_C2, 7, 0.001, 0.05)]
    beams_out = dict(ntris=np.array([c[0] for c in cfg]), soup_seed=np.array([c[1] for c in cfg], dtype=np.uint64),
                     beam_seed=np.array([c[2] for c in cfg]), spread=np.array([c[3] for c in cfg]), width=np.array([c[4] for c in cfg]))
    for k, (nt, ss, bs, spread, width) in enumerate(cfg):
        rs = ref.build(scenes.triangle_soup(nt, ss))
        bb = scenes.random_beams(1500, bs, spread=spread, width=width)
        beams_out[f"codes{k}"] = np.array([rs.beam_visibility(b[:3], b[3:]) for b in bb], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "beams.npz"), **beams_out)

    # 4. MT19937: first outputs of randomMT2() as consumed by the reference (checked indirectly by the frames above;
    #    committed as integers for the device generator test)
    orc = ol.Oracle()
    np.savez_compressed(os.path.join(HERE, "mt19937_seed4357.npz"), first=orc.mt_stream_u32(2000),
                        at_1e6=orc.mt_stream_u32(1000008)[-8:])
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
