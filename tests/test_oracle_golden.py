"""CPU: the oracle restatement against golden vectors produced by the compiled reference (tests/golden/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as ol
from lucille_b200 import scenes


def tree_digest(nodes, triorder):
    h = hashlib.sha256()
    for name in ol.NODE_DTYPE.names:
        h.update(np.ascontiguousarray(nodes[name]).tobytes())
    h.update(np.ascontiguousarray(triorder).tobytes())
    return h.hexdigest()


def test_per_ray_vectors(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "soup2k_rays.npz"))
    tris = scenes.triangle_soup(int(g["ntris"]), int(g["seed"]))
    rays8 = scenes.pinhole_rays(int(g["width"]), int(g["height"]))
    rays6 = scenes.rays_f32_to_f64(rays8)
    t = oracle.build(tris)
    assert tree_digest(t.nodes(), t.triorder()) == str(g["tree_digest"])
    assert len(t.nodes()) == int(g["nnodes"]) and t.max_depth() == int(g["max_depth"])
    hits, cnt = t.intersect_f64(rays6, counters=True)
    m = g["hit"] == 1
    assert np.array_equal(hits["hit"] == 1, m) and m.sum() > 500
    for f in ("t", "u", "v"):                       # bit-exact: same arithmetic, same order
        assert np.array_equal(hits[f][m], g[f][m]), f
    assert np.array_equal(t.triorder()[hits["prim"][m]], g["tri_index"][m])
    st = t.state_build(rays6, hits)
    for f in ("P", "Ng", "Ns", "tangent", "binormal"):
        assert np.array_equal(st[f][m], g[f][m]), f
    assert [int(x) for x in cnt] == [int(x) for x in g["counters"]]
    # occlusion boolean == closest-hit boolean (the AO transport only uses the flag, ambientocclusion.c:123-129)
    assert np.array_equal(t.occluded_f64(rays6) == 1, m)


@pytest.mark.parametrize("name,maker", [
    ("soup100k_c2", lambda: scenes.triangle_soup(100000, scenes.SEED_C2)),
    ("soup17", lambda: scenes.triangle_soup(17, 7)),
    ("soup16", lambda: scenes.triangle_soup(16, 7)),
    ("soup1", lambda: scenes.triangle_soup(1, 7)),
    ("dup300", lambda: np.repeat(scenes.triangle_soup(3, 9), 100, axis=0)),
    ("flat500", lambda: scenes.triangle_soup(500, 11) * np.array([1.0, 1.0, 0.0])),
])
def test_tree_digests(oracle, golden_dir, name, maker):
    g = np.load(os.path.join(golden_dir, "tree_digests.npz"))
    t = oracle.build(maker())
    assert len(t.nodes()) == int(g[name + "_nnodes"])
    assert tree_digest(t.nodes(), t.triorder()) == str(g[name])


def test_empty_scene(oracle):
    t = oracle.build(np.zeros((0, 3, 3)))
    assert t.empty and len(t.nodes()) == 0
    rays = scenes.rays_f32_to_f64(scenes.pinhole_rays(4, 4))
    hits, cnt = t.intersect_f64(rays, counters=True)
    assert hits["hit"].sum() == 0 and int(cnt["nrays"]) == 0      # bvh.c:446: returns before the counter


def test_mt19937(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "mt19937_seed4357.npz"))
    assert np.array_equal(oracle.mt_stream_u32(2000), g["first"])
    assert np.array_equal(oracle.mt_stream_u32(1000008)[-8:], g["at_1e6"])
    d = oracle.mt_stream(10)
    assert np.array_equal(d, g["first"][:10].astype(np.float64) * 2.3283064365386963e-10)


@pytest.mark.parametrize("fname,kw", [
    ("c1_frame_160x120.npz", dict(width=160, height=120)),
    ("c1_frame_97x61_ps2_g16.npz", dict(width=97, height=61, xsamples=2, ysamples=2, gather=16)),
])
def test_c1_frame_bit_identical(oracle, golden_dir, fname, kw):
    """ambient_occlusion.rib (BASELINE configs[0]) at reduced size: the oracle's frame equals the reference's float
    framebuffer bit for bit, with the same number of rays (single MT stream, spiral bucket order)."""
    sc = np.load(os.path.join(golden_dir, "c1_scene.npz"))
    g = np.load(os.path.join(golden_dir, fname))
    t = oracle.build(sc["tris"])
    w, h = kw.pop("width"), kw.pop("height")
    fp = ol.frame_params(sc["cam"], w, h, **kw)
    rgb, nrays = t.render_ao(fp)
    assert nrays == int(g["nrays"])
    assert np.array_equal(rgb, g["rgb"])


def test_f32_restatement_tracks_double(oracle):
    """The fp32 instantiation (the bit-exact target of the fp32 kernels) against the double one: same hits except
    near-degenerate cases, |dt|/t within a few fp32 ulps."""
    tris = scenes.triangle_soup(20000, scenes.SEED_C2)
    rays8 = scenes.pinhole_rays(128, 128)
    t = oracle.build(tris)
    h64 = t.intersect_f64(scenes.rays_f32_to_f64(rays8))
    h32 = t.intersect_f32(rays8)
    m64, m32 = h64["hit"] == 1, h32["prim"] != ol.MISS_PRIM
    assert (m64 == m32).mean() > 0.9995
    both = m64 & m32
    same = h64["prim"][both] == h32["prim"][both]
    assert same.mean() > 0.999
    rel = np.abs(h32["t"][both][same].astype(np.float64) - h64["t"][both][same]) / h64["t"][both][same]
    assert rel.max() < 1e-5
    assert np.array_equal(t.occluded_f32(rays8) == 1, m32)


def test_plane_sphere_ao_frame_with_vertex_normals(oracle, golden_dir):
    """examples/plane_sphere (BASELINE configs[3] scene; the shipped reference renders it with the AO transport): the sphere
    carries vertex normals, so calculate_occlusion samples about the interpolated, un-normalised Ns -- bit-identical frame."""
    sc = np.load(os.path.join(golden_dir, "c4_scene.npz"))
    g = np.load(os.path.join(golden_dir, "c4_ao_frame_96x96_ps2_g16.npz"))
    t = oracle.build(sc["tris"])
    t.set_normals(sc["normals"])
    rgb, nrays = t.render_ao(ol.frame_params(sc["cam"], 96, 96, xsamples=2, ysamples=2, gather=16))
    assert nrays == int(g["nrays"]) and np.array_equal(rgb, g["rgb"])
    t.set_normals(None)
    flat, _ = t.render_ao(ol.frame_params(sc["cam"], 96, 96, xsamples=2, ysamples=2, gather=16))
    assert not np.array_equal(flat, g["rgb"])          # the normals matter


def test_city_frame_golden(oracle, golden_dir):
    """C6 (tests/golden/make_city_golden.py): a 15 552-triangle architectural scene at RIB scale whose world-space vertices are not fp32
    numbers, rendered by the compiled reference -- the restatement's frame is bit-identical, with the reference's ray count."""
    g = np.load(os.path.join(golden_dir, "c6_city.npz"))
    t = oracle.build(g["tris"])
    rgb, nrays = t.render_ao(ol.frame_params(g["cam"], int(g["width"]), int(g["height"]), xsamples=int(g["ps"]), ysamples=int(g["ps"]), gather=int(g["gather"])))
    assert nrays == int(g["nrays"]) and np.array_equal(rgb, g["rgb"])


def test_beam_visibility_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "beams.npz"))
    for k in range(3):
        tris = scenes.triangle_soup(int(g["ntris"][k]), int(g["soup_seed"][k]))
        beams = scenes.random_beams(1500, int(g["beam_seed"][k]), spread=float(g["spread"][k]), width=float(g["width"][k]))
        assert np.array_equal(oracle.build(tris).beam_visibility(beams), g[f"codes{k}"])
    assert set(np.unique(np.concatenate([g["codes0"], g["codes1"], g["codes2"]]))) == {-1, 0, 1, 2}


def test_sunsky_sky_lookup_golden(oracle, golden_dir):
    """Row a12: ri_sunsky_get_sky_rgb (sunsky.c:322-408) -- the restatement reproduces the compiled reference's float RGB
    bit for bit on three sites/times/turbidities x 4096 directions (grazing and below-horizon ones included)."""
    g = np.load(os.path.join(golden_dir, "sunsky.npz"))
    for k in range(int(g["nsky"])):
        blk = ol.sunsky_block(g[f"sky{k}_rec"], g)
        got = oracle.sunsky_sky_rgb(blk, ol.sky_dirs(4096, 100 + k))
        assert np.array_equal(got, g[f"sky{k}_rgb"])
        assert (got != 0).any(axis=1).sum() > 1500 and (got == 0).all(axis=1).sum() > 1500


def test_sunsky_frame_golden(oracle, golden_dir):
    """Row a12: gather_sunsky + contribution_from_sunlight (ambientocclusion.c:153-324) through the pixel loop:
    ambient_occlusion.rib with an AreaLightSource "sunsky", 120x90, 2x2 pixel samples, reference at one thread."""
    g = np.load(os.path.join(golden_dir, "sunsky.npz"))
    blk = ol.sunsky_block(g["frame_block"], g)
    assert blk.nsun == 1
    t = oracle.build(g["frame_tris"])
    rgb, nrays = t.render_sunsky(ol.frame_params(g["frame_cam"], 120, 90, xsamples=2, ysamples=2), blk)
    assert nrays == int(g["frame_nrays"])
    assert np.array_equal(rgb, g["frame_rgb"])
    assert rgb.max() > 1000.0 and (rgb[..., 2] > rgb[..., 0]).mean() > 0.3      # a blue-ish sky lit the scene


def test_hdr_output_golden(oracle, golden_dir):
    """The .hdr files of the reference's display driver for the committed frames (sha256 + size in hdr.npz, written by
    tests/golden/make_hdr_golden.py from the compiled reference)."""
    import hashlib
    g = np.load(os.path.join(golden_dir, "hdr.npz"))
    frames = dict(ol.hdr_cases())
    frames["c1"] = np.load(os.path.join(golden_dir, "c1_frame_160x120.npz"))["rgb"]
    frames["sunsky"] = np.load(os.path.join(golden_dir, "sunsky.npz"))["frame_rgb"]
    for name, rgb in frames.items():
        data = oracle.hdr_encode(rgb)
        assert len(data) == int(g[name + "_size"]) and hashlib.sha256(data).hexdigest() == str(g[name + "_sha256"]), name


def test_textured_ao_frame_golden(oracle, golden_dir):
    """Row a11, the material-texture clause (ambientocclusion.c:393-401): st lerp + ri_texture_fetch + per-channel multiply through the
    pixel loop -- the reference's framebuffer of tests/scenes/textured_quads.rib, bit for bit (st spanning several periods, negative
    st, a quad without texture coordinates)."""
    g = np.load(os.path.join(golden_dir, "textured_quads.npz"))
    t = oracle.build(g["tris"])
    t.set_attributes(None, None, g["st"], g["has_st"], None)
    rgb, nrays = t.render_ao_textured(ol.frame_params(g["cam"], 120, 90, gather=16), g["tex"])
    assert nrays == int(g["nrays"]) and np.array_equal(rgb, g["rgb"])
    assert (np.ptp(rgb, axis=2) > 1e-3).mean() > 0.2          # the texture coloured the image


def test_point_gathers_golden(oracle, golden_dir):
    """SURVEY 8f rank 2, per-point hemisphere gathers: the compiled reference's occlusion() shadeop, ri_ibl_sample_cosweight and
    ri_domelight_sample at 800 shading points (tests/golden/make_gather_golden.py) -- the restatement reproduces them bit for bit."""
    g = np.load(os.path.join(golden_dir, "point_gathers.npz"))
    t = oracle.build(scenes.triangle_soup(int(g["ntris"]), int(g["seed"])))
    for kind, ns in g["cases"]:
        got, _ = t.point_gather(int(kind), int(ns), g["points"], g["env"] if kind == 1 else None, g["col"], float(g["intensity"]))
        assert np.array_equal(got, g[f"k{kind}_n{ns}"]), (kind, ns)
    for kind, ns, dim in g["qmc_cases"]:                       # Option "use_qmc" branches
        got, _ = t.point_gather_qmc(int(kind), int(ns), g["points"], g["qmc_instance"], int(dim), g["env"] if kind == 1 else None,
                                    g["col"], float(g["intensity"]))
        assert np.array_equal(got, g[f"q{kind}_n{ns}_d{dim}"]), (kind, ns, dim)


def _shade_scene(g):
    sizes = [int(x) for x in g["sizes"]]
    tris = scenes.triangle_soup(sum(sizes), int(g["seed_t"]))
    col, st, _flags, has_col, has_st, inside = ol.attribute_case(len(tris), sizes, int(g["attr_seed"]))
    return tris, (col, has_col, st, has_st, inside)


def test_shade_callers_golden(oracle, golden_dir):
    """SURVEY 8f rank 2, the two shading-language callers of ri_raytrace: the compiled reference's trace() shadeop at 1500 (P, R)
    pairs and its next_lightsource() loop at 250 shading points (tests/golden/make_shade_golden.py) -- the restatement reproduces the
    shader input blocks, the miss colours and the returned light samples bit for bit."""
    g = np.load(os.path.join(golden_dir, "shade_callers.npz"))
    tris, attrs = _shade_scene(g)
    t = oracle.build(tris)
    t.set_attributes(*attrs)
    got = t.shade_trace(g["pr"], g["env"])
    hit = got["hits"]["hit"] == 1
    assert np.array_equal(g["trace_called"] == 1, hit) and 0.2 < hit.mean() < 0.9
    assert np.array_equal(g["trace_dst"][~hit], got["miss_rgb"][~hit])
    for name, field in (("Cs", got["exts"]["color"]), ("P", got["states"]["P"]), ("N", got["states"]["Ns"]), ("Ng", got["states"]["Ng"]),
                        ("dPdu", got["states"]["tangent"]), ("dPdv", got["states"]["binormal"]), ("I", got["eye"])):
        assert np.array_equal(g["trace_" + name][hit], field[hit]), name
    assert np.array_equal(g["trace_s"][hit], got["hits"]["u"][hit].astype(np.float32))
    assert np.array_equal(g["trace_t"][hit], got["hits"]["v"][hit].astype(np.float32))
    lt = oracle.build(scenes.triangle_soup(int(g["ntris_l"]), int(g["seed_l"])))
    for i, (ns, angle) in enumerate(g["light_cases"]):
        L, Cl, vis, nrays = lt.light_samples(int(ns), float(angle), g["points"], g["env"])
        cnt = g[f"light{i}_count"]
        assert np.array_equal(vis.sum(axis=1), cnt) and nrays >= vis.sum()
        for p in range(len(cnt)):
            k = int(cnt[p])
            assert np.array_equal(L[p][vis[p] == 1], g[f"light{i}_L"][p, :k]) and np.array_equal(Cl[p][vis[p] == 1], g[f"light{i}_Cl"][p, :k])


def test_socket_display_stream_golden(oracle, golden_dir):
    """The socket display driver's byte stream for the committed frames (sha256 + size in sockdrv.npz, captured from the compiled
    reference by tests/golden/make_sockdrv_golden.py)."""
    import hashlib
    g = np.load(os.path.join(golden_dir, "sockdrv.npz"))
    frames = dict(ol.hdr_cases())
    frames["c1"] = np.load(os.path.join(golden_dir, "c1_frame_160x120.npz"))["rgb"]
    frames["sunsky"] = np.load(os.path.join(golden_dir, "sunsky.npz"))["frame_rgb"]
    for name, rgb in frames.items():
        data = oracle.sockdrv_encode(rgb)
        assert len(data) == int(g[name + "_size"]) and hashlib.sha256(data).hexdigest() == str(g[name + "_sha256"]), name


def test_point_ao_ray_setup_matches_the_numpy_generator():
    """orc_ao_point_rays_f32 (calculate_occlusion's ray set-up, ambientocclusion.c:56-117, with the counter RNG and the deterministic
    sin/cos) against the independent numpy statement of the same loop (scenes.ao_rays: libm sin/cos): identical fp32 records except
    where the last place of sin/cos moves a float rounding, never by more than one ulp; windows of a batch are position-keyed."""
    rng = np.random.default_rng(3)
    P = rng.random((2000, 3))
    n = rng.normal(size=(2000, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n[:3] = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, -1.0]]
    pts = np.concatenate([P, n], axis=1)
    for ntheta, nphi in ((8, 8), (3, 5)):
        got = ol.Oracle().ao_point_rays(pts, ntheta, nphi, scenes.SEED_C3)
        want = scenes.ao_rays(P, n, ntheta, nphi, scenes.SEED_C3)
        assert got.shape == want.shape
        same = got.view(np.uint32) == want.view(np.uint32)
        assert same.mean() > 0.9999
        assert np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64)).max() <= 1
        tail = ol.Oracle().ao_point_rays(pts[1500:], ntheta, nphi, scenes.SEED_C3, first_point=1500)
        assert np.array_equal(tail.view(np.uint32), got[1500 * ntheta * nphi:].view(np.uint32))
