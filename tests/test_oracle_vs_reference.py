"""CPU, build container only: the oracle restatement against the compiled, unmodified reference (oracle/_ref).
Skipped where oracle/_ref has not been built (it needs /root/reference); the golden-vector tests cover that case."""
import os

import numpy as np
import pytest

import oracle_lib as ol
from lucille_b200 import scenes

pytestmark = pytest.mark.skipif(not ol.reference_available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def ref():
    return ol.Reference()


@pytest.mark.parametrize("n,seed", [(5000, 1), (30000, 2), (33, 3)])
def test_tree_and_rays(oracle, ref, n, seed):
    tris = scenes.triangle_soup(n, seed)
    ot, rs = oracle.build(tris), ref.build(tris)
    on, rn = ot.nodes(), rs.nodes()
    assert len(on) == len(rn)
    for f in ol.NODE_DTYPE.names:
        assert np.array_equal(on[f], rn[f]), f
    assert np.array_equal(ot.triorder(), rs.triorder())
    assert all(np.array_equal(a, b) for a, b in zip(ot.bbox(), rs.bbox()))

    rng = np.random.default_rng(seed)
    org = rng.uniform(-0.5, 1.5, (20000, 3))
    tgt = rng.uniform(0.0, 1.0, (20000, 3))
    rays6 = np.concatenate([org, tgt - org], axis=1)          # un-normalised directions are legal (simplerender.cpp:202)
    oh = ot.intersect_f64(rays6)
    rh, _ = rs.intersect(rays6)
    m = rh["hit"] != 0
    assert np.array_equal(oh["hit"] == 1, m)
    for f in ("t", "u", "v"):
        assert np.array_equal(oh[f][m], rh[f][m]), f
    assert np.array_equal(ot.triorder()[oh["prim"][m]], rh["index"][m] // 3)
    st = ot.state_build(rays6, oh)
    for f in ("P", "Ng", "Ns", "tangent", "binormal"):
        assert np.array_equal(st[f][m], rh[f][m]), f


def test_multi_geom_flattening(oracle, ref):
    """create_triangle_list (bvh.c:1736-1826) concatenates geoms in list order: 3 geoms == 1 flattened soup."""
    tris = scenes.triangle_soup(900, 5)
    a = ref.build(tris, geom_sizes=[100, 500, 300])
    b = oracle.build(tris)
    for f in ol.NODE_DTYPE.names:
        assert np.array_equal(a.nodes()[f], b.nodes()[f]), f
    assert np.array_equal(a.triorder(), b.triorder())


def test_counters_match_reference_statistics(oracle):
    refs = ol.Reference(stats=True)
    tris = scenes.triangle_soup(20000, scenes.SEED_C2)
    rays6 = scenes.rays_f32_to_f64(scenes.pinhole_rays(96, 96))
    rs = refs.build(tris)
    refs.stats_reset()
    rs.intersect(rays6)
    want = refs.stats_get()
    _, cnt = oracle.build(tris).intersect_f64(rays6, counters=True)
    assert {k: int(cnt[k]) for k in want} == want


def test_beam_visibility(oracle, ref):
    """Row a10: frustum visibility query (ri_beam_set + ri_bvh_intersect_beam_visibility) -- every outcome class."""
    for ntris, seed, kw in [(20000, 5, {}), (300, 6, dict(width=0.002)), (20000, 7, dict(spread=0.001, width=0.05))]:
        tris = scenes.triangle_soup(ntris, scenes.SEED_C2 if ntris > 1000 else 3)
        beams = scenes.random_beams(1500, seed, **kw)
        got = oracle.build(tris).beam_visibility(beams)
        rs = ref.build(tris)
        want = np.array([rs.beam_visibility(b[:3], b[3:]) for b in beams], dtype=np.int32)
        assert np.array_equal(got, want)


def test_sunsky_sky_lookup(oracle, ref, golden_dir):
    """Row a12: the sky lookup against the compiled ri_sunsky_init + ri_sunsky_get_sky_rgb on a site/time the goldens do not hold."""
    import os
    tables = np.load(os.path.join(golden_dir, "sunsky.npz"))
    for name, n in (("S0", 41), ("S1", 41), ("S2", 41)):
        assert np.array_equal(ref.table(name + "Amplitudes", n), tables[name])
    dirs = ol.sky_dirs(20000, 77)
    want, rec = ref.sunsky_eval(dirs, latitude=60.17, longitude=24.94, sm=2.0, jd=200, tod=18.5, turbidity=3.3)
    got = oracle.sunsky_sky_rgb(ol.sunsky_block(rec, tables), dirs)
    assert np.array_equal(got, want)


def test_hdr_output_step(oracle, ref, golden_dir, tmp_path):
    """the .hdr display driver (hdr_dd_open/write/close -> RGBE_WriteHeader + RGBE_WritePixels_RLE):
    the restatement writes the reference's file byte for byte."""
    import os
    cases = dict(ol.hdr_cases())
    cases["c1"] = np.load(os.path.join(golden_dir, "c1_frame_160x120.npz"))["rgb"]
    cases["sunsky"] = np.load(os.path.join(golden_dir, "sunsky.npz"))["frame_rgb"]
    for name, rgb in cases.items():
        assert oracle.hdr_encode(rgb) == ref.hdr_file(rgb, str(tmp_path / (name + ".hdr"))), name


def test_hit_state_colours_texcoords_inside(oracle, ref):
    """Row a8, the rest of ri_intersection_state_build (intersection_state.c:123-133, 192-246): E, I, vertex-colour lerp or (1,1,1),
    st lerp (shared and unshared texture coordinates) or 0, the two-sided back-side flag -- bit-identical to the reference."""
    sizes = [300, 200, 250, 150, 100]
    tris = scenes.triangle_soup(sum(sizes), 21)
    colors, st, flags, has_color, has_st, inside = ol.attribute_case(len(tris), sizes, 4)
    rs = ref.build(tris, geom_sizes=sizes)
    rs.set_attributes(colors, st, flags)
    ot = oracle.build(tris)
    ot.set_attributes(colors, has_color, st, has_st, inside)
    rng = np.random.default_rng(8)
    org = rng.uniform(-0.5, 1.5, (30000, 3))
    rays6 = np.concatenate([org, rng.uniform(0.0, 1.0, (30000, 3)) - org], axis=1)
    want = rs.intersect_ext(rays6)
    hits = ot.intersect_f64(rays6)
    got = ot.state_ext(rays6, hits)
    m = want["hit"] != 0
    assert m.sum() > 3000 and np.array_equal(got["hit"] != 0, m)
    for f in ("E", "I", "color", "st", "t", "inside"):
        assert np.array_equal(got[f][m], want[f][m]), f
    # every class occurs: coloured and default colour, st and none, both sides
    assert (got["color"][m] != 1.0).any() and (got["color"][m] == 1.0).all(axis=1).any()
    assert (got["st"][m] != 0.0).any() and (got["st"][m] == 0.0).all(axis=1).any() and set(np.unique(got["inside"][m])) == {0, 1}


def test_texture_fetch(oracle, ref):
    """ri_texture_fetch (render/texture.c:86-236) on 50 000 coordinates incl. integers and values just below them."""
    tex = ol.test_texture()
    rng = np.random.default_rng(5)
    uv = rng.uniform(-3, 4, (50000, 2))
    uv[:100] = np.round(uv[:100])
    uv[100:200, 0] = np.nextafter(np.round(uv[100:200, 0]), -np.inf)
    assert np.array_equal(oracle.texture_fetch(tex, uv), ref.texture_fetch(tex, uv))


def test_transports_called_directly(oracle, ref):
    """ri_transport_dirtmap (compiled into the reference, never called by its pixel loop) and, for symmetry,
    ri_transport_ambientocclusion, called per eye ray with the thread's MT19937 stream re-seeded: the restatement returns the same
    radiance bit for bit.  The soup is scaled by 3 so that the dirt map's 0.1 / 0.5 distance window is exercised on both sides."""
    tris = scenes.triangle_soup(3000, 9) * 3.0
    rs, ot = ref.build(tris), oracle.build(tris)
    rng = np.random.default_rng(2)
    org = rng.uniform(-1, 4, (3000, 3))
    rays6 = np.concatenate([org, rng.uniform(0, 3, (3000, 3)) - org], axis=1)
    for which in (1, 0, 3):                                  # 3 = ri_transport_sample (transport.c:50-173): white on a hit
        want, got = rs.transport_batch(which, rays6), ot.transport_batch(which, rays6)
        assert np.array_equal(got, want) and (want[:, 0] > 0).sum() > 1000
    dirt = ot.transport_batch(1, rays6)[:, 0]
    assert np.unique(np.round(dirt * 16, 6)).size > 100          # distances inside the window: more than the 17 AO levels


def _glass_stack(n=12):
    """n parallel quads (two triangles each) one behind the other: a chain of refractions that outlives MAX_TRACE_DEPTH."""
    tris = []
    for i in range(n):
        z = 0.2 * i
        tris += [[[-1, -1, z], [1, -1, z], [1, 1, z]], [[-1, -1, z], [1, 1, z], [-1, 1, z]]]
    return np.array(tris, dtype=np.float64)


def test_whitted_transport_called_directly(oracle, ref):
    """ri_transport_whitted of the compiled reference per eye ray (refraction chains, total internal reflection,
    angular-map environment lookup) against the restatement: bit-identical radiance; the glass stack cuts chains at depth 8."""
    env = ol.test_texture(31, 29, 7)
    rng = np.random.default_rng(2)
    for tris, lo, hi in ((scenes.triangle_soup(20000, 9), 0.0, 1.0), (_glass_stack(), -0.6, 0.6)):
        rs, ot = ref.build(tris), oracle.build(tris)
        rs.set_envmap(env)
        org = rng.uniform(-0.3, 1.3, (4000, 3))
        org[:, 2] = -1.0 - rng.uniform(0, 1, 4000)
        tgt = np.concatenate([rng.uniform(lo, hi, (4000, 2)), rng.uniform(0.0, 1.0, (4000, 1))], axis=1)
        rays6 = np.concatenate([org, tgt - org], axis=1)
        want = rs.transport_batch(2, rays6)
        got, nrays = ot.transport_whitted(rays6, env)
        assert np.array_equal(got, want)
        assert nrays > 1.5 * len(rays6)
    assert (got.sum(axis=1) == 0).sum() > 100          # chains cut at MAX_TRACE_DEPTH leave zero radiance


def _shading_points(oracle_tree, n_side=64):
    """(P, N) of the primary hits of a small camera on the tree's scene: realistic shading points for the per-point gathers."""
    rays = scenes.rays_f32_to_f64(scenes.pinhole_rays(n_side, n_side))
    hits = oracle_tree.intersect_f64(rays)
    st = oracle_tree.state_build(rays, hits)
    m = hits["hit"] == 1
    return np.concatenate([st["P"][m][:, :3], st["Ns"][m][:, :3]], axis=1)


@pytest.mark.parametrize("kind,nsamples", [(0, 48), (0, 16), (0, 2), (1, 48), (1, 30), (2, 27), (2, 100)])
def test_point_gathers_called_directly(oracle, ref, kind, nsamples):
    """SURVEY 8f rank 2, the per-point hemisphere gathers: the occlusion() shadeop (shader.c:680-768), ri_ibl_sample_cosweight
    (ibl.c:53-228) and ri_domelight_sample (ibl.c:231-389) of the compiled reference, called once per shading point, against the
    restatement: bit-identical results (sample counts that are and are not 3 * ntheta^2)."""
    tris = scenes.triangle_soup(20000, 9)
    rs, ot = ref.build(tris), oracle.build(tris)
    env = ol.test_texture(32, 32, 5)
    rs.set_envmap(env)
    pts = _shading_points(ot)[:1500]
    assert len(pts) > 500
    col, inten = (0.9, 0.5, 0.25), 2.5
    want = rs.point_gather(kind, nsamples, pts, col, inten)
    got, nrays = ot.point_gather(kind, nsamples, pts, env if kind == 1 else None, col, inten)
    nth = max(1, int(np.sqrt(int(nsamples / 3.0))))
    assert nrays == len(pts) * 3 * nth * nth
    assert np.array_equal(got, want)
    assert np.ptp(got[:, 0]) > 0.05                     # neither all open nor all blocked


@pytest.mark.parametrize("kind,nsamples,dim", [(1, 48, 0), (1, 7, 0), (2, 48, 0), (2, 31, 3), (2, 16, 24)])
def test_point_gathers_qmc_called_directly(oracle, ref, kind, nsamples, dim):
    """The quasi-Monte Carlo branches of the same gathers (Option "use_qmc"): scrambled Halton / Hammersley points from lucille's
    Faure permutation table, per-point instance numbers (inray->i) and dimension (inray->d) -- bit-identical to the compiled
    reference."""
    tris = scenes.triangle_soup(20000, 9)
    rs, ot = ref.build(tris), oracle.build(tris)
    env = ol.test_texture(32, 32, 5)
    rs.set_envmap(env)
    pts = _shading_points(ot)[:1200]
    inst = (np.arange(len(pts)) * 7919 % 5000).astype(np.int32)
    col, inten = (0.9, 0.5, 0.25), 2.5
    for instance in (None, inst):
        want = rs.point_gather_qmc(kind, nsamples, pts, instance, dim, col, inten)
        got, nrays = ot.point_gather_qmc(kind, nsamples, pts, instance, dim, env if kind == 1 else None, col, inten)
        assert nrays == len(pts) * nsamples
        assert np.array_equal(got, want)
    assert np.ptp(got[:, 0]) > 0.05


def test_trace_shadeop_called_directly(oracle, ref):
    """SURVEY 8f rank 2, the trace() shadeop (shader.c:895-976): the compiled function called per (P, R) pair, a capturing shader
    procedure on every geom of a five-geom scene with colours / texture coordinates / two-sided geometry, an IBL light carrying the
    environment map.  The restatement reproduces, bit for bit, what the reference hands the hit surface's shader (Cs, P, N, Ng, dPdu,
    dPdv, I, s, t) and what it returns on a miss (the environment colour along the UNnormalised direction)."""
    sizes = [1500, 900, 1300, 1100, 1200]
    tris = scenes.triangle_soup(sum(sizes), 21)
    col, st, geom_flags, has_col, has_st, inside = ol.attribute_case(len(tris), sizes, 4)
    rs = ref.build(tris, geom_sizes=sizes)
    rs.set_attributes(col, st, geom_flags)
    ot = oracle.build(tris)
    ot.set_attributes(col, has_col, st, has_st, inside)
    env = ol.test_texture(32, 32, 5)
    rs.set_envmap(env)
    rng = np.random.default_rng(4)
    n = 4000
    P = rng.uniform(-0.3, 1.3, (n, 3))
    R = rng.uniform(0.0, 1.0, (n, 3)) - P
    R *= rng.uniform(0.3, 2.5, (n, 1))                   # trace() does not normalise R
    pr = np.concatenate([P, R], axis=1)
    want = rs.shade_trace(pr)
    got = ot.shade_trace(pr, env)
    hit = got["hits"]["hit"] == 1
    assert 0.2 < hit.mean() < 0.9
    assert np.array_equal(want["called"] == 1, hit)
    assert np.array_equal(want["dst"][~hit], got["miss_rgb"][~hit]) and np.ptp(got["miss_rgb"][~hit]) > 0.1
    assert np.all(want["dst"][hit] == [0.25, 0.5, 0.75])         # the capturing shader's Ci
    for name, field in (("Cs", got["exts"]["color"]), ("P", got["states"]["P"]), ("N", got["states"]["Ns"]), ("Ng", got["states"]["Ng"]),
                        ("dPdu", got["states"]["tangent"]), ("dPdv", got["states"]["binormal"]), ("I", got["eye"])):
        assert np.array_equal(want[name][hit], field[hit]), name
    assert np.array_equal(want["s"][hit], got["hits"]["u"][hit].astype(np.float32))
    assert np.array_equal(want["t"][hit], got["hits"]["v"][hit].astype(np.float32))
    assert np.all(want["ray_depth"][hit] == 1)


@pytest.mark.parametrize("nsamples,angle", [(48, 1.2), (27, 1.5707963267948966), (5, 0.6)])
def test_next_lightsource_called_directly(oracle, ref, nsamples, angle):
    """The illuminance loop's light samples (next_lightsource + init_lightsource, shader.c:1116-1186, 1236-1310): per shading point the
    compiled reference returns, in order, the environment-map samples inside the cone that no triangle occludes -- never the last
    sample of the set (shader.c:1170-1177).  The restatement's visible samples are the same list: L and Cl bit for bit."""
    tris = scenes.triangle_soup(20000, 9)
    rs, ot = ref.build(tris), oracle.build(tris)
    env = ol.test_texture(32, 32, 5)
    rs.set_envmap(env)
    pts = _shading_points(ot)[:800]
    Lw, Clw, cw = rs.light_samples(nsamples, angle, pts)
    L, Cl, vis, nrays = ot.light_samples(nsamples, angle, pts, env)
    assert np.array_equal(vis.sum(axis=1), cw)
    assert 0 < vis.sum() < vis.size and nrays > vis.sum()
    assert not vis[:, -1].any()
    for p in range(len(pts)):
        k = int(cw[p])
        assert np.array_equal(L[p][vis[p] == 1], Lw[p, :k]) and np.array_equal(Cl[p][vis[p] == 1], Clw[p, :k]), p


@pytest.mark.parametrize("w,h", [(64, 64), (160, 120), (97, 61), (33, 31)])
def test_socket_display_stream(oracle, ref, w, h):
    """SURVEY 8f rank 4, second half: what the reference's socket display driver (display/sockdrv.c) puts on the wire for a frame --
    captured from sock_dd_open/write/close talking to a listener in this process -- against the restatement, byte for byte: header,
    one message per 1024 pixels in bucket order, the finish command, and the remainder below 1024 pixels that is never sent."""
    from lucille_b200 import accel
    rgb = np.random.default_rng(w * h).uniform(-1, 5, (h, w, 3)).astype(np.float32)
    fr = accel.make_frame(np.eye(4).reshape(16), 1.0, False, w, h, 1, 1)
    pix = accel.frame_pixels(fr)                                     # visiting order (x | y << 16), y counted from the bucket origin
    disp = (pix & 0xFFFF) | ((np.uint32(h - 1) - (pix >> 16)) << 16)  # bucket_write: row H-1-y
    try:
        want = ref.sockdrv_stream(rgb, disp)
    except RuntimeError as e:                                         # the driver's fixed port 12346 is taken on this machine
        pytest.skip(str(e))
    got = oracle.sockdrv_encode(rgb)
    assert len(want) == 16 + (w * h // 1024) * (8 + 24 * 1024) + 4
    assert got == want


def test_ibl_bruteforce_returns_zero_power(oracle, ref):
    """Why ri_ibl_sample_bruteforce (ibl.c:395-518) has no device form: after the first ray that escapes, the function overwrites the
    ray ORIGIN with the direction and measures the distance between the two (ibl.c:497-505) -- zero -- so every term of its sum is
    multiplied by invdist = 0.  The compiled reference agrees: exactly zero power at every shading point, bright environment or not."""
    tris = scenes.triangle_soup(2000, 9)
    rs, ot = ref.build(tris), oracle.build(tris)
    rs.set_envmap(ol.test_texture(16, 16, 5) + 3.0)
    pts = _shading_points(ot, 24)[:40]
    assert len(pts) >= 20
    out = rs.point_gather(3, 1, pts)
    assert not out.any()


@pytest.mark.parametrize("mesh", ["bunny.obj", "sponza_cooked.obj", "sphere.obj", "cornellbox.obj"])
def test_real_meshes_against_the_reference(oracle, ref, mesh):
    """the reference's own testbed meshes, read where they lie (skipped where /root/reference is absent): shared
    vertices, thousands of axis-aligned walls and zero-extent boxes (sponza, cornell box), sliver triangles.  Restatement against the
    compiled reference: same tree (every node, box and the triangle order), same closest hits bit for bit for camera rays from three
    sides and for rays between random points inside the bounding box, same traversal counters."""
    path = os.path.join(ol.REFERENCE_MESH_DIR, mesh)
    if not os.path.exists(path):
        pytest.skip("reference testbed meshes not present")
    tris = ol.load_obj_triangles(path)
    assert len(tris) > 10
    rs, ot = ref.build(tris), oracle.build(tris)
    rn, on = rs.nodes(), ot.nodes()
    assert len(rn) == len(on)
    for f in ol.NODE_DTYPE.names:
        assert np.array_equal(rn[f], on[f]), f
    assert np.array_equal(rs.triorder(), ot.triorder())
    lo, hi = tris.reshape(-1, 3).min(axis=0), tris.reshape(-1, 3).max(axis=0)
    c, ext = 0.5 * (lo + hi), float((hi - lo).max())
    rng = np.random.default_rng(len(tris))
    rays = []
    for axis in range(3):                                            # a camera on each side, looking at the centre
        eye = c.copy()
        eye[axis] += 1.7 * ext
        tgt = c + rng.uniform(-0.55, 0.55, (4000, 3)) * ext
        rays.append(np.concatenate([np.tile(eye, (4000, 1)), tgt - eye], axis=1))
    a, b = rng.uniform(lo, hi, (6000, 3)), rng.uniform(lo, hi, (6000, 3))
    rays.append(np.concatenate([a, b - a], axis=1))                  # incoherent rays that start inside the scene
    rays6 = np.concatenate(rays)
    want, _ = rs.intersect(rays6, nthreads=1, want_hits=True)
    got, cnt = ot.intersect_f64(rays6, counters=True)
    m = want["hit"] == 1
    assert np.array_equal(got["hit"] == 1, m) and m.sum() > 1000
    for f in ("t", "u", "v"):
        assert np.array_equal(got[f][m], want[f][m]), f
    assert np.array_equal(ot.triorder()[got["prim"][m]], want["index"][m] // 3)        # state.index = 3 * triangle number (bvh.c:1813)


def test_box_city_against_the_reference(oracle, ref):
    """The procedural stand-in for the real meshes that CAN travel to the GPU box (scenes.box_city: axis-aligned boxes on a tiled
    ground, lattice coordinates -- equal box bounds, zero-extent boxes, coplanar faces, rays grazing edges): restatement against the
    compiled reference, tree and closest hits bit for bit."""
    tris = scenes.box_city(40, 40, 3)
    rs, ot = ref.build(tris), oracle.build(tris)
    rn, on = rs.nodes(), ot.nodes()
    assert len(rn) == len(on)
    for f in ol.NODE_DTYPE.names:
        assert np.array_equal(rn[f], on[f]), f
    assert np.array_equal(rs.triorder(), ot.triorder())
    rays6 = scenes.rays_f32_to_f64(np.concatenate([scenes.pinhole_rays(96, 96, eye=(0.5, 0.9, -1.2)), _city_rays(6000, 2)]))
    want, _ = rs.intersect(rays6, nthreads=1, want_hits=True)
    got = ot.intersect_f64(rays6)
    m = want["hit"] == 1
    assert np.array_equal(got["hit"] == 1, m) and 2000 < m.sum() < len(m)
    for f in ("t", "u", "v"):
        assert np.array_equal(got[f][m], want[f][m]), f
    assert np.array_equal(ot.triorder()[got["prim"][m]], want["index"][m] // 3)


def _city_rays(n, seed):
    """Rays between random points above the city: many run along walls and roofs."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(0, 1, (n, 3)) * np.array([1.0, 0.7, 1.0])
    b = rng.uniform(0, 1, (n, 3)) * np.array([1.0, 0.7, 1.0])
    r = np.zeros((n, 8), dtype=np.float32)
    r[:, 0:3], r[:, 4:7], r[:, 7] = a, b - a, 1.0e38
    return r
