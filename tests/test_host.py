"""CPU: host-side logic of the product -- the C-ABI library loads and exports every declared symbol, the host BVH
builder reproduces the reference tree (via the oracle), flat device records decode back to the canonical tree, and
every compute entry fails loudly without a device (no CPU fallback)."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest

import oracle_lib as ol
from lucille_b200 import accel, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lucille_b200.h")).read()
    declared = set(re.findall(r"\b(ri_b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(accel.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/lucille_b200.h but not exported"
    assert declared == {n for n, _, _ in accel.ABI}, "python binding table out of sync with the header"


def test_bind_rejects_other_methods():
    with pytest.raises(accel.B200Error):
        accel.Accel.bind(accel.RI_ACCEL_BVH)          # ri_accel_bind -> -1 for methods this library does not provide
    assert accel.Accel.bind(accel.RI_ACCEL_B200).data is None


@pytest.mark.parametrize("maker", [
    lambda: scenes.triangle_soup(100000, scenes.SEED_C2),
    lambda: scenes.triangle_soup(70000, 21),            # above the parallel threshold: threaded build must not change the tree
    lambda: scenes.triangle_soup(600000, 22),           # above 2^19: the top ranges are binned/partitioned by all threads
    lambda: scenes.triangle_soup(17, 7),
    lambda: scenes.triangle_soup(16, 7),
    lambda: scenes.triangle_soup(1, 7),
    lambda: np.repeat(scenes.triangle_soup(3, 9), 100, axis=0),
    lambda: scenes.triangle_soup(500, 11) * np.array([1.0, 1.0, 0.0]),
    lambda: scenes.box_city(40, 40, 3),                 # axis-aligned architecture: shared vertices, coplanar faces, exact ties everywhere
    lambda: scenes.box_city(150, 120, 5),               # 216 K triangles of the same, above the parallel threshold
])
def test_host_builder_matches_oracle_tree(oracle, maker):
    tris = maker()
    a = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.PREC_F64 | accel.HOST_ONLY)
    t = oracle.build(tris)
    pn, on = a.nodes(), t.nodes()
    assert len(pn) == len(on)
    for f in ol.NODE_DTYPE.names:
        assert np.array_equal(pn[f], on[f]), f
    assert np.array_equal(a.triorder(), t.triorder())
    info = a.info()
    assert info.max_depth == t.max_depth() and info.ninner + info.nleaf == len(on)
    assert np.array_equal(np.array(info.bmin), t.bbox()[0]) and np.array_equal(np.array(info.bmax), t.bbox()[1])


def _decode(flat, prec):
    """Walk the flat device records from root_word and rebuild the canonical DFS-preorder node list."""
    nodes = flat["nodes32"] if prec == 32 else flat["nodes64"]
    out = []

    def rec(word):
        me = len(out)
        out.append(None)
        if word & 0x80000000:
            out[me] = dict(is_leaf=1, tri_start=word & ((1 << 27) - 1), ntris=((word >> 27) & 15) + 1)
            return me
        n = nodes[word]
        d = dict(is_leaf=0, axis=int(n["axis"]),
                 lbox=[n["x"][0], n["y"][0], n["z"][0], n["x"][1], n["y"][1], n["z"][1]],
                 rbox=[n["x"][2], n["y"][2], n["z"][2], n["x"][3], n["y"][3], n["z"][3]])
        out[me] = d
        d["child0"] = rec(int(n["c0"]))
        d["child1"] = rec(int(n["c1"]))
        return me

    rec(flat["root_word"])
    return out


@pytest.mark.parametrize("n", [5000, 12])
def test_flat_records_decode_to_canonical_tree(n):
    import sys
    sys.setrecursionlimit(10000)
    tris = scenes.triangle_soup(n, 3)
    a = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.PREC_F64 | accel.HOST_ONLY)
    canon, flat = a.nodes(), a.flat()
    for prec in (32, 64):
        dec = _decode(flat, prec)
        assert len(dec) == len(canon)
        for d, c in zip(dec, canon):
            assert d["is_leaf"] == c["is_leaf"]
            if d["is_leaf"]:
                assert d["ntris"] == c["ntris"]          # the word's start is a SLOT index, checked below
            else:
                assert d["axis"] == c["axis"] and d["child0"] == c["child0"] and d["child1"] == c["child1"]
                lb, rb = np.array(d["lbox"], dtype=np.float64), np.array(d["rbox"], dtype=np.float64)
                if prec == 64:
                    assert np.array_equal(lb, c["lbox"]) and np.array_equal(rb, c["rbox"])
                else:                                    # fp32 boxes are rounded outward, by at most one ulp
                    for got, want in ((lb, c["lbox"]), (rb, c["rbox"])):
                        assert np.all(got[:3] <= want[:3]) and np.all(got[3:] >= want[3:])
                        assert np.all(np.abs(got - want) <= np.spacing(np.abs(want).astype(np.float32)).astype(np.float64))
    # triangle slots: every leaf starts at a multiple-of-four slot, owns round_up(ntris, 4) slots, carries prim ids in post-build
    # order; v0, e1 = v1 - v0, e2 = v2 - v0; the filler slots are zero-area triangles with prim = MISS
    order = a.triorder()
    src = tris[order]
    leaves = canon[canon["is_leaf"] == 1]
    t64, t32 = flat["tris64"], flat["tris32"]
    assert flat["nslots"] == int(((leaves["ntris"] + 3) // 4 * 4).sum()) == len(t32) == len(t64)
    leaf_words = [d for d in _decode(flat, 32) if d["is_leaf"]]
    for d, c in zip(leaf_words, leaves):
        s0, cnt, p0 = d["tri_start"], int(c["ntris"]), int(c["tri_start"])
        assert s0 % 4 == 0
        assert np.array_equal(t32["prim"][s0:s0 + cnt], np.arange(p0, p0 + cnt))
        assert np.array_equal(t64["prim"][s0:s0 + cnt], np.arange(p0, p0 + cnt))
        assert np.array_equal(t64["v0"][s0:s0 + cnt], src[p0:p0 + cnt, 0])
        assert np.array_equal(t64["e1"][s0:s0 + cnt], src[p0:p0 + cnt, 1] - src[p0:p0 + cnt, 0])
        assert np.array_equal(t64["e2"][s0:s0 + cnt], src[p0:p0 + cnt, 2] - src[p0:p0 + cnt, 0])
        s32 = src[p0:p0 + cnt].astype(np.float32)
        assert np.array_equal(t32["v0"][s0:s0 + cnt], s32[:, 0]) and np.array_equal(t32["e1"][s0:s0 + cnt], s32[:, 1] - s32[:, 0])
        assert np.array_equal(t32["e2"][s0:s0 + cnt], s32[:, 2] - s32[:, 0])
        for f in range(s0 + cnt, s0 + (cnt + 3) // 4 * 4):
            assert t32["prim"][f] == 0xFFFFFFFF and not t32["e1"][f].any() and not t32["e2"][f].any()
    assert flat["top_count"] == min(1024, flat["ninner"])


def test_leaf_transposed_copies_are_a_permutation_of_the_slots():
    """The pooled kernels read the leaf-TRANSPOSED copies: chunk k (32 bytes) of item j of a leaf whose rows hold m items sits at
    slot0 * sizeof(slot) + (k * m + j) * 32 (fp32: item = pair of slots, m = slots/2; fp64: item = slot, m = slots; slots =
    round_up(ntris, 4)).  An fp32 item holds its two triangles (A, B) INTERLEAVED word by word -- word 2f+h = field f of half h,
    f = v0.xyz e1.xyz e2.xyz, words 18/19 = prim of A/B -- so that every field arrives as an aligned register pair for the packed
    FFMA2 arithmetic (csrc/packed.cuh); rows 0 and 1 hold words 0..7 / 8..15 (32 B per item), row 2 words 16..19 (16 B per item, at
    byte 64 m of the leaf's block); the masked half of an odd last pair gets unit edges."""
    tris = scenes.triangle_soup(3000, 3)
    a = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.PREC_F64 | accel.HOST_ONLY)
    flat = a.flat()
    t32 = flat["tris32"]
    t64 = flat["tris64"].view(np.uint8).reshape(-1, 96)
    t32t, t64t = flat["tris32t"], flat["tris64t"]
    leaves = [n for n in a.nodes() if n["is_leaf"]]
    slot0, odd_leaves = 0, 0
    for leaf in leaves:
        ntris = int(leaf["ntris"])
        ns = (ntris + 3) // 4 * 4
        assert np.array_equal(t32["prim"][slot0:slot0 + ntris], np.arange(leaf["tri_start"], leaf["tri_start"] + ntris))
        m = ns // 2
        for j in range(m):
            want = np.zeros(24, dtype=np.uint32)
            for h in range(2):
                s = t32[slot0 + 2 * j + h]
                f = np.concatenate([s["v0"], s["e1"], s["e2"]]).astype(np.float32)
                want[h:18:2] = f.view(np.uint32)
                want[18 + h] = s["prim"]
            if (ntris & 1) and j == (ntris + 1) // 2 - 1:          # masked half of the last used pair: unit edges
                one = np.float32(1.0).view(np.uint32)
                want[2 * 3 + 1], want[2 * 7 + 1] = one, one
                odd_leaves += 1
            for k in range(2):                                      # rows 0 and 1: 32 bytes per item
                got = t32t[slot0 * 48 + (k * m + j) * 32: slot0 * 48 + (k * m + j) * 32 + 32]
                assert np.array_equal(got.view(np.uint32), want[8 * k: 8 * k + 8])
            got = t32t[slot0 * 48 + m * 64 + j * 16: slot0 * 48 + m * 64 + j * 16 + 16]      # row 2: 16 bytes per item
            assert np.array_equal(got.view(np.uint32), want[16:20])
        for j in range(ns):
            for k in range(3):
                got = t64t[slot0 * 96 + (k * ns + j) * 32: slot0 * 96 + (k * ns + j) * 32 + 32]
                assert np.array_equal(got, t64[slot0 + j, 32 * k: 32 * k + 32])
        slot0 += ns
    assert slot0 == flat["nslots"] and odd_leaves > 10


def test_empty_scene_builds_valid_accelerator():
    a = accel.Accel.bind().build(np.zeros((0, 3, 3)), accel.PREC_F32 | accel.HOST_ONLY)
    info = a.info()
    assert info.empty == 1 and info.ntris == 0 and len(a.nodes()) == 0
    assert a.flat()["root_word"] == 0x7FFFFFFF


def test_no_cpu_fallback():
    """Without a device the product refuses to compute (it must never route through the oracle)."""
    tris = scenes.triangle_soup(100, 1)
    a = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.HOST_ONLY)
    with pytest.raises(accel.B200Error, match="no CPU fallback"):
        a.intersect(scenes.pinhole_rays(4, 4))
    with pytest.raises(accel.B200Error, match="no CPU fallback"):
        a.occluded(scenes.pinhole_rays(4, 4))
    if accel.device_count() == 0:
        with pytest.raises(accel.B200Error, match="no CPU fallback"):
            accel.Accel.bind().build(tris)
    src = open(os.path.join(ROOT, "lucille_b200", "accel.py")).read() + open(os.path.join(ROOT, "lucille_b200", "scenes.py")).read()
    assert "oracle" not in src.replace("the oracle", "").replace("oracle,", "") or "import oracle" not in src


def test_synthetic_scenes_are_deterministic():
    a = scenes.triangle_soup(1000, scenes.SEED_C2)
    b = scenes.triangle_soup(1000, scenes.SEED_C2, chunk=128)
    assert np.array_equal(a, b)
    assert hashlib.sha256(a.tobytes()).hexdigest() == hashlib.sha256(scenes.triangle_soup(1000, scenes.SEED_C2).tobytes()).hexdigest()
    assert np.array_equal(a, a.astype(np.float32).astype(np.float64))          # fp32-representable
    r = scenes.pinhole_rays(64, 32)
    assert r.shape == (2048, 8) and np.all(r[:, 4:7] != 0.0)
    o = ol.Oracle()
    assert int(scenes.splitmix64(np.array([12345], dtype=np.uint64))[0]) == o.lib.orc_splitmix64(12345)


def test_ao_ray_generator_is_cosine_hemisphere():
    pts = np.array([[0.5, 0.5, 0.5]] * 50)
    nrm = np.tile(np.array([[0.0, 0.0, 1.0]]), (50, 1))
    rays = scenes.ao_rays(pts, nrm, 8, 8, 99)
    assert rays.shape == (50 * 64, 8)
    d = rays[:, 4:7].astype(np.float64)
    assert np.all(d[:, 2] >= 0.0) and np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-6)
    assert abs(d[:, 2].mean() - 2.0 / 3.0) < 0.02                                # E[cos] under a cosine-weighted pdf
    assert np.allclose(rays[:, 2], 0.5 + 1e-6, atol=1e-7)


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes / numpy mirrors in lucille_b200/accel.py against the C header: a probe compiled from include/lucille_b200.h prints
    sizeof (and the offset of the last member) of every struct that crosses the boundary."""
    import ctypes
    import subprocess
    from lucille_b200 import accel
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    probe = tmp_path / "probe.c"
    probe.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "lucille_b200.h"
#define S(t, last) printf(#t " %zu %zu\\n", sizeof(t), offsetof(t, last))
int main(void) {
    S(ri_b200_hit_f32, prim); S(ri_b200_hit_f64, hit); S(ri_b200_state_f64, binormal); S(ri_b200_state_ext_f64, hit);
    S(ri_b200_info_t, build_seconds); S(ri_b200_counters_t, nhit_tris); S(ri_b200_frame_t, precision);
    S(ri_b200_frame_stats_t, ms_resolve); S(ri_b200_gather_t, qmc_instance); S(ri_b200_sunsky_t, nsun); S(ri_b200_path_frame_t, bucket_size);
    S(ri_b200_trace_rec_f64, hit); S(ri_b200_light_t, env_height); S(ri_b200_ao_points_t, eps);
    return 0;
}
''')
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(probe), "-o", str(exe)])
    got = {}
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        name, size, off = line.split()
        got[name] = (int(size), int(off))

    def ct(cls, last):
        return ctypes.sizeof(cls), getattr(cls, last).offset

    def nd(dt, last):
        return dt.itemsize, dt.fields[last][1]

    want = {
        "ri_b200_hit_f32": nd(accel.HIT32_DTYPE, "prim"), "ri_b200_hit_f64": nd(accel.HIT64_DTYPE, "hit"),
        "ri_b200_state_f64": nd(accel.STATE_DTYPE, "binormal"), "ri_b200_state_ext_f64": nd(accel.STATE_EXT_DTYPE, "hit"),
        "ri_b200_info_t": ct(accel.Info, "build_seconds"), "ri_b200_counters_t": ct(accel.Counters, "nhit_tris"),
        "ri_b200_frame_t": ct(accel.Frame, "precision"), "ri_b200_frame_stats_t": ct(accel.FrameStats, "ms_resolve"),
        "ri_b200_gather_t": ct(accel.Gather, "qmc_instance"), "ri_b200_sunsky_t": ct(accel.Sunsky, "nsun"),
        "ri_b200_path_frame_t": ct(accel.PathFrame, "bucket_size"),
        "ri_b200_trace_rec_f64": nd(accel.TRACE_REC_DTYPE, "hit"), "ri_b200_light_t": ct(accel.Light, "env_height"),
        "ri_b200_ao_points_t": ct(accel.AoPoints, "eps"),
    }
    assert got == want


def test_light_sample_count_is_init_lightsource_s():
    """ri_b200_light_samples_count = the number of samples init_lightsource() sets up for Option narealight_rays (shader.c:1263-1266):
    ntheta = (int)sqrt((int)(n / 3.0)), at least 1; m = ntheta * 3 ntheta.  Pure host logic (no device needed)."""
    import math
    lib = accel.load_library()
    for n in (1, 2, 3, 5, 11, 12, 26, 27, 47, 48, 75, 108, 300, 1000):
        nt = max(1, int(math.sqrt(int(n / 3.0))))
        assert lib.ri_b200_light_samples_count(n) == 3 * nt * nt, n
