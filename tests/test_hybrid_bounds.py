"""CPU (-m "not gpu"): the error bounds of csrc/hybrid.cuh, checked as mathematics.

The hybrid kernels decide box and triangle tests in fp32 when a certified bound allows it and recompute the rest in double.  Their
bit-exactness rests on one claim: a decision declared CERTAIN in fp32 is the decision the double reference takes.  The GPU suite checks
the kernels' answers on millions of rays; this file restates the fp32 classification in numpy (float32 arrays: every product and sum
rounded separately, as the kernel's packed FFMA2 arithmetic does) and throws adversarial inputs at it -- rays aimed within 1e-12..1e-4
of edges and vertices, origins 1e-9..1e-3 off the triangle's plane on either side, slivers, near-parallel rays, scenes at scales
1e-3..1e3 and offsets up to 1e4 -- asserting that no certain decision ever contradicts the double expression tree of bvh.c:730-791 /
869-936 (restated here in float64 exactly as oracle/oracle_trav.inc does).
The test has teeth: with every constant of the bounds cut to 1/8 of the kernel's (12 u, 9 u, 1 eta, E = 1 u ...) five of the fourteen
cases find contradictions, at 1/16 eleven do; at 1/2 and 1/4 none does -- the kernel's constants carry a factor >= 4 over what these
adversarial inputs can provoke, on top of the third they already have over the derivation."""
import numpy as np
import pytest

U = np.float32(2.0 ** -24)
F = np.float32


def _f32_down(x):
    f = x.astype(np.float32)
    return np.where(f.astype(np.float64) > x, np.nextafter(f, F(-np.inf)), f)


def _f32_up(x):
    f = x.astype(np.float32)
    return np.where(f.astype(np.float64) < x, np.nextafter(f, F(np.inf)), f)


def _tri_ref(O, D, v0, e1, e2):
    """triangle_isect in double (bvh.c:730-791) from t_leaf = 1e38: accepted with t < 1e38."""
    with np.errstate(all="ignore"):
        p = np.stack([D[:, 1] * e2[:, 2] - D[:, 2] * e2[:, 1], D[:, 2] * e2[:, 0] - D[:, 0] * e2[:, 2], D[:, 0] * e2[:, 1] - D[:, 1] * e2[:, 0]], axis=1)
        a = e1[:, 0] * p[:, 0] + e1[:, 1] * p[:, 1] + e1[:, 2] * p[:, 2]
        ok = np.abs(a) > 1.0e-14
        inva = 1.0 / a
        s = O - v0
        q = np.stack([s[:, 1] * e1[:, 2] - s[:, 2] * e1[:, 1], s[:, 2] * e1[:, 0] - s[:, 0] * e1[:, 2], s[:, 0] * e1[:, 1] - s[:, 1] * e1[:, 0]], axis=1)
        u = (s[:, 0] * p[:, 0] + s[:, 1] * p[:, 1] + s[:, 2] * p[:, 2]) * inva
        v = (q[:, 0] * D[:, 0] + q[:, 1] * D[:, 1] + q[:, 2] * D[:, 2]) * inva
        t = (e2[:, 0] * q[:, 0] + e2[:, 1] * q[:, 1] + e2[:, 2] * q[:, 2]) * inva
        ok &= ~((u < 0.0) | (u > 1.0)) & ~((v < 0.0) | ((u + v) > 1.0)) & ~((t < 0.0) | (t > 1.0e38)) & (t < 1.0e38)
    return ok


def _cross32(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)


def _dot32(a, b):
    return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]


def _tri_class32(O, D, v0, e1, e2, c, own, bmax):
    """hyb_tri_class of csrc/hybrid.cuh: 1 certainly accepted, 0 certainly rejected, 2 undecided.  own = the filter's own records
    (coordinates minus c, edges rounded once from the doubles); else the shared records (edges subtracted in fp32 from fp32 vertices)."""
    with np.errstate(all="ignore"):
        oc = O - c
        oh = oc.astype(np.float32)
        ol = (oc - oh.astype(np.float64)).astype(np.float32)
        d = D.astype(np.float32)
        if own:
            v0f, e1f, e2f = (v0 - c).astype(np.float32), e1.astype(np.float32), e2.astype(np.float32)
            exact = False
            eta0, de = U * F(bmax), F(0.0)
        else:
            v0f = v0.astype(np.float32)
            v1f, v2f = (v0 + e1).astype(np.float32), (v0 + e2).astype(np.float32)
            e1f, e2f = v1f - v0f, v2f - v0f
            exact = bool(np.array_equal(v0f.astype(np.float64), v0) and np.array_equal(v1f.astype(np.float64), v0 + e1) and np.array_equal(v2f.astype(np.float64), v0 + e2))
            eta0 = F(0.0) if exact else U * F(bmax)
            de = F(2.0) * eta0
        p = _cross32(d, e2f)
        a = _dot32(e1f, p)
        s = (oh - v0f) + ol
        q = _cross32(s, e1f)
        Uu, Vv, Tt = _dot32(s, p), _dot32(q, d), _dot32(e2f, q)
        Md, Me1, Me2, Ms = (np.abs(x).max(axis=1) for x in (d, e1f, e2f, s))
        Mo = np.abs(oh).max(axis=1) + F(np.abs(c).max())
        G0 = F(8.0) * (eta0 + (F(4.0) * U * U) * Mo) + F(1.0e-30)
        G = F(72.0) * U * Ms + G0
        P12, MdE1, MdE2 = Me1 * Me2, Md * Me1, Md * Me2
        ea, eU, eV, eT = F(96.0) * U * (Md * P12), MdE2 * G, MdE1 * G, P12 * G
        if not exact:
            Ms1, Es, d8 = Ms + G, (Me1 + Me2) + de, F(8.0) * de
            ea, eU, eV, eT = ea + d8 * Md * Es, eU + d8 * Md * Ms1, eV + d8 * Md * Ms1, eT + d8 * Ms1 * Es
        sg = np.where(np.signbit(a), F(-1.0), F(1.0))
        A, Us, Vs, Ts = np.abs(a), Uu * sg, Vv * sg, Tt * sg
        Alo, Ahi = A - ea, A + ea
        Ulo, Uhi, Vlo, Vhi, Tlo, Thi = Us - eU, Us + eU, Vs - eV, Vs + eV, Ts - eT, Ts + eT
        acc = (Alo > F(1.0001e-14)) & (Ulo >= 0) & (Vlo >= 0) & (Tlo >= 0) & ((Alo - Uhi) - Vhi >= 0) & (Thi <= F(1.0e30) * Alo)
        rej = (Ahi <= F(0.9999e-14)) | (Uhi < 0) | (Vhi < 0) | (Thi < 0) | (Ulo > Ahi) | (Ulo + Vlo > Ahi)
    return np.where(acc, 1, np.where(rej, 0, 2))


def _adversarial_triangles(rng, n, scale, offset, snap32):
    """Triangles and rays that sit on every threshold of the window: rays through points within tiny distances of edges and vertices,
    origins a hair off the plane on both sides (the AO pattern), slivers, grazing directions."""
    c0 = offset + rng.uniform(-1.0, 1.0, (n, 3)) * scale
    e1 = rng.normal(size=(n, 3)) * scale * 10.0 ** rng.uniform(-3.0, 0.0, (n, 1))
    e2 = rng.normal(size=(n, 3)) * scale * 10.0 ** rng.uniform(-3.0, 0.0, (n, 1))
    sliver = rng.random(n) < 0.15
    e2[sliver] = e1[sliver] * rng.uniform(-2.0, 2.0, (sliver.sum(), 1)) + rng.normal(size=(sliver.sum(), 3)) * scale * 10.0 ** rng.uniform(-9.0, -4.0, (sliver.sum(), 1))
    v0 = c0
    if snap32:                                   # fp32-representable vertices: the shared records are exact
        v0 = v0.astype(np.float32).astype(np.float64)
        v1, v2 = (v0 + e1).astype(np.float32).astype(np.float64), (v0 + e2).astype(np.float32).astype(np.float64)
        e1, e2 = v1 - v0, v2 - v0
    # target point in barycentric coordinates, often ON the window's boundary (u = 0, v = 0, u + v = 1, vertices) give or take a hair
    kind = rng.integers(0, 6, n)
    bu, bv = rng.random(n), rng.random(n)
    bu = np.where(kind == 0, 0.0, np.where(kind == 3, 1.0, bu))
    bv = np.where(kind == 1, 0.0, np.where(kind == 2, 1.0 - bu, np.where(kind == 3, 0.0, bv)))
    hair = 10.0 ** rng.uniform(-12.0, -3.0, n) * rng.choice([-1.0, 1.0], n)
    bu, bv = bu + hair * (kind < 4), bv + hair * rng.choice([-1.0, 0.0, 1.0], n) * (kind < 4)
    tgt = v0 + bu[:, None] * e1 + bv[:, None] * e2
    nrm = np.cross(e1, e2)
    nn = np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm = np.where(nn > 0, nrm / np.where(nn > 0, nn, 1.0), [0.0, 0.0, 1.0])
    d = rng.normal(size=(n, 3))
    graze = rng.random(n) < 0.2                  # nearly in the plane
    d[graze] -= nrm[graze] * ((d[graze] * nrm[graze]).sum(axis=1, keepdims=True)) * (1.0 - 10.0 ** rng.uniform(-8.0, -2.0, (graze.sum(), 1)))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= 10.0 ** rng.uniform(-1.0, 1.0, (n, 1)) * (rng.random((n, 1)) < 0.3) + (rng.random((n, 1)) >= 0.0) * 1.0   # some rays are not normalised
    dist = np.where(rng.random(n) < 0.5, scale * 10.0 ** rng.uniform(-9.0, -3.0, n) * rng.choice([-1.0, 1.0], n), scale * rng.uniform(0.01, 3.0, n))
    O = tgt - d * dist[:, None]
    return O, d, v0, e1, e2


@pytest.mark.parametrize("scale,offset,snap32,own", [
    (1.0, 0.0, True, False), (1.0, 0.0, False, False), (1.0, 0.0, False, True),
    (1.0e-3, 0.0, False, True), (1.0e3, 0.0, False, True), (4.0, 1.0e3, False, True), (0.05, 1.0e4, False, True),
    (30.0, 5.0, True, False), (30.0, 5.0, False, False),
])
def test_certain_triangle_decisions_are_the_double_decisions(scale, offset, snap32, own):
    rng = np.random.default_rng(int(scale * 1000) + int(offset) + 2 * snap32 + own)
    n = 400_000
    off = np.array([offset, -0.7 * offset, 0.3 * offset])
    O, D, v0, e1, e2 = _adversarial_triangles(rng, n, scale, off, snap32)
    c = off.astype(np.float32).astype(np.float64) if own else np.zeros(3)
    bmax = float(np.abs(np.concatenate([v0, v0 + e1, v0 + e2]) - c).max())           # the (translated) scene box's largest |coordinate|
    ref = _tri_ref(O, D, v0, e1, e2)
    cls = _tri_class32(O, D, v0, e1, e2, c, own, bmax)
    assert not (ref[cls == 0]).any(), int(ref[cls == 0].sum())                      # certainly rejected, yet the reference accepts
    assert (ref[cls == 1]).all(), int((~ref[cls == 1]).sum())                       # certainly accepted, yet the reference rejects
    assert 0.05 < ref.mean() < 0.6                                                  # the inputs do straddle the window
    assert (cls == 2).mean() < 0.9 and (cls == 1).sum() > 1000 and (cls == 0).sum() > 1000    # most inputs sit ON a threshold by construction


def _slab64(lo, hi, O, inv, sign):
    near = np.where(sign, hi, lo)
    far = np.where(sign, lo, hi)
    with np.errstate(all="ignore"):
        tn, tf = (near - O) * inv, (far - O) * inv
    tmin = np.where(tn[:, 0] > tn[:, 1], tn[:, 0], tn[:, 1])
    tmax = np.where(tf[:, 0] < tf[:, 1], tf[:, 0], tf[:, 1])
    tmin = np.where(tmin > tn[:, 2], tmin, tn[:, 2])
    tmax = np.where(tmax < tf[:, 2], tmax, tf[:, 2])
    return (tmax > 0.0) & (tmin <= tmax) & (tmin < 1.0e38)


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1.0e-3, 0.0), (1.0e3, 0.0), (4.0, 1.0e3), (0.05, 1.0e4)])
def test_certain_box_decisions_are_the_double_decisions(scale, offset):
    """The child-box test: fp32 slab values from outward-rounded (translated) boxes with the bound E of hybrid.cuh; rays that graze
    box faces, edges and corners, start on faces, or run along an axis."""
    rng = np.random.default_rng(int(scale * 1000) + int(offset))
    n = 400_000
    off = np.array([offset, -0.7 * offset, 0.3 * offset])
    c = off.astype(np.float32).astype(np.float64)
    ctr = off + rng.uniform(-1.0, 1.0, (n, 3)) * scale
    half = scale * 10.0 ** rng.uniform(-4.0, -0.3, (n, 3))
    half[rng.random((n, 3)) < 0.05] = 0.0                                           # zero-extent boxes (axis-aligned geometry)
    lo, hi = ctr - half - 1.0e-14, ctr + half + 1.0e-14                             # bbox_add_margin
    # aim at a point on the box's surface region, give or take a hair
    tgt = ctr + half * rng.choice([-1.0, 1.0, 0.0], (n, 3), p=[0.35, 0.35, 0.3]) * (1.0 + 10.0 ** rng.uniform(-12.0, -3.0, (n, 3)) * rng.choice([-1.0, 1.0], (n, 3)))
    D = rng.normal(size=(n, 3))
    axisp = rng.random(n) < 0.1
    D[axisp] *= rng.choice([0.0, 1.0e-20, 1.0], (axisp.sum(), 3), p=[0.3, 0.2, 0.5])
    D[np.abs(D).sum(axis=1) == 0] = [0.0, 0.0, 1.0]
    O = tgt - D * (scale * rng.uniform(-0.5, 3.0, (n, 1)))
    sign = D < 0.0
    with np.errstate(all="ignore"):
        inv = np.where(np.abs(D) > 1.0e-14, 1.0 / D, np.where(D < 0.0, -np.finfo(np.float64).max, np.finfo(np.float64).max))
        ref = _slab64(lo, hi, O, inv, sign)
        lo32, hi32 = _f32_down(lo - c), _f32_up(hi - c)
        bmaxs = np.maximum(np.abs(lo - c), np.abs(hi - c)).max(axis=0)                # per axis, over the "scene"
        bmax32 = np.nextafter(bmaxs.astype(np.float32), F(np.inf))
        oh = (O - c).astype(np.float32)
        inv32 = inv.astype(np.float32)
        E = F(8.0) * U * (np.abs(inv32) * (bmax32 + np.abs(oh))).max(axis=1)
        E = np.where(E < F(1.0e30), E, F(np.inf))
        near = np.where(sign, hi32, lo32)
        far = np.where(sign, lo32, hi32)
        tn, tf = (near - oh) * inv32, (far - oh) * inv32
        tmin, tmax = tn.max(axis=1), tf.min(axis=1)
        E2 = E + E
        pass_c = (tmax - E > 0) & (tmax - tmin >= E2) & (tmin < F(1.0e37))
        fail_c = (tmax + E <= 0) | (tmin - tmax > E2)
    assert ref[pass_c].all(), int((~ref[pass_c]).sum())
    assert not ref[fail_c & ~pass_c].any(), int(ref[fail_c & ~pass_c].sum())
    und = ~pass_c & ~fail_c
    assert 0.1 < ref.mean() < 0.9 and pass_c.sum() > 1000 and fail_c.sum() > 1000 and und.mean() < 0.9


@pytest.mark.parametrize("scale,offset,own", [(1.0, 0.0, False), (1.0, 0.0, True), (4.0, 1.0e3, True), (1.0e3, 0.0, True)])
def test_beyond_best_t_filter_of_the_closest_hit_kernel(scale, offset, own):
    """closest_hybrid_kernel skips a triangle when T_lo > round_up(best_t) * |a|_hi (hyb_tri_maybe): then the reference's t for that
    triangle, if it accepts it at all, is strictly greater than best_t -- with best_t placed within hairs of the true t on both sides."""
    rng = np.random.default_rng(int(scale) + int(offset) + own)
    n = 400_000
    off = np.array([offset, -0.7 * offset, 0.3 * offset])
    O, D, v0, e1, e2 = _adversarial_triangles(rng, n, scale, off, snap32=not own)
    c = off.astype(np.float32).astype(np.float64) if own else np.zeros(3)
    bmax = float(np.abs(np.concatenate([v0, v0 + e1, v0 + e2]) - c).max())
    with np.errstate(all="ignore"):
        p = np.cross(D, e2)
        a64 = (e1 * p).sum(axis=1)
        t64 = (e2 * np.cross(O - v0, e1)).sum(axis=1) / a64
    ref = _tri_ref(O, D, v0, e1, e2)
    best_t = np.abs(t64) * (1.0 + 10.0 ** rng.uniform(-15.0, -2.0, n) * rng.choice([-1.0, 1.0], n))
    best_t = np.where(np.isfinite(best_t) & (best_t > 0), best_t, 1.0)
    best_hi = _f32_up(best_t)
    # the fp32 quantities of _tri_class32, inlined for T_lo and |a|_hi
    with np.errstate(all="ignore"):
        oc = O - c
        oh = oc.astype(np.float32); ol = (oc - oh.astype(np.float64)).astype(np.float32); d = D.astype(np.float32)
        if own:
            v0f, e1f, e2f, eta0, de, exact = (v0 - c).astype(np.float32), e1.astype(np.float32), e2.astype(np.float32), U * F(bmax), F(0.0), False
        else:
            v0f = v0.astype(np.float32); e1f = (v0 + e1).astype(np.float32) - v0f; e2f = (v0 + e2).astype(np.float32) - v0f
            eta0, de, exact = F(0.0), F(0.0), True
        pf = _cross32(d, e2f); a = _dot32(e1f, pf); s = (oh - v0f) + ol; q = _cross32(s, e1f); Tt = _dot32(e2f, q)
        Md, Me1, Me2, Ms = (np.abs(x).max(axis=1) for x in (d, e1f, e2f, s))
        Mo = np.abs(oh).max(axis=1) + F(np.abs(c).max())
        G = F(72.0) * U * Ms + (F(8.0) * (eta0 + (F(4.0) * U * U) * Mo) + F(1.0e-30))
        ea, eT = F(96.0) * U * (Md * (Me1 * Me2)), (Me1 * Me2) * G
        Ts = Tt * np.where(np.signbit(a), F(-1.0), F(1.0))
        skip = (Ts - eT) > best_hi * (np.abs(a) + ea) * F(1.000001)
    wrong = skip & ref & ~(t64 > best_t)
    assert not wrong.any(), int(wrong.sum())
    assert skip.sum() > 1000 and (ref & ~skip).sum() > 1000
