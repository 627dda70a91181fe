"""GPU (-m gpu): parity at BASELINE.json's FULL sizes, and the point-based AO entry.

  * configs[2] -- 1 M-triangle soup, the 16 Mi-ray AO batch of bench.py: >= 1 Mi rays sampled ACROSS the whole batch against the oracle's
    fp32 instantiation, occlusion and closest hit, bit for bit; the per-point counts of ri_b200_occlusion_points_f32 over the WHOLE
    batch against the occlusion bytes of the same rays (a checksum of checksums: 262 144 counts = 16 777 216 verdicts).
  * configs[4] -- 10 M-triangle soup: the device-built tree equals the host builder's node for node; rays sampled over a 4096^2-style
    camera against the oracle (whose tree is the compiled reference's, tests/test_oracle_vs_reference.py).
  * calculate_occlusion as a batch: generated rays equal the oracle's restatement bit for bit, counts equal the oracle's verdicts."""
import numpy as np
import pytest

import oracle_lib as ol
from lucille_b200 import accel, scenes

pytestmark = pytest.mark.gpu


def _need_gpu():
    if accel.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the B200 box (no CPU fallback exists)")


def _primary_points(a, tris_post, npoints, w=1024, h=1024):
    rays = scenes.pinhole_rays(w, h)
    hits = a.intersect(rays)
    idx = np.flatnonzero(hits["prim"] != accel.MISS_PRIM)[:npoints]
    t = hits["t"][idx].astype(np.float64)
    P = rays[idx, 0:3].astype(np.float64) + rays[idx, 4:7].astype(np.float64) * t[:, None]
    tri = tris_post[hits["prim"][idx]]
    n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return np.ascontiguousarray(np.concatenate([P, n], axis=1))


@pytest.mark.parametrize("ntheta,nphi,npts", [(8, 8, 3000), (3, 5, 1777), (1, 1, 100), (16, 16, 300)])
def test_point_ao_rays_equal_the_restatement(ntheta, nphi, npts):
    """ri_b200_ao_point_rays_f32 == orc_ao_point_rays_f32 (calculate_occlusion's ray set-up, ambientocclusion.c:56-117) bit for bit,
    and ri_b200_occlusion_points_f32 == per-point sums of the oracle's fp32 occlusion verdicts on those rays."""
    _need_gpu()
    tris = scenes.triangle_soup(20000, scenes.SEED_C2)
    a = accel.Accel.bind().build(tris, accel.PREC_F32)
    orc = ol.Oracle().build(tris)
    pts = _primary_points(a, tris[a.triorder()], npts, 256, 256)
    assert len(pts) == npts
    # also points whose normal picks each branch of ri_ortho_basis (|n_k| >= 0.6 on the first axes)
    pts[:3, 3:6] = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.6, 0.6, 0.5291502622129182]]
    seed = 0xB2000003 + ntheta
    got = a.ao_point_rays(pts, ntheta, nphi, seed)
    want = ol.Oracle().ao_point_rays(pts, ntheta, nphi, seed)
    assert got.shape == want.shape == (npts * ntheta * nphi, 8)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    counts = a.occlusion_points(pts, ntheta, nphi, seed)
    verdicts = orc.occluded_f32(want)
    assert np.array_equal(counts, verdicts.reshape(npts, ntheta * nphi).sum(axis=1).astype(np.uint32))
    assert 0 < counts.sum() < len(want)
    assert len(a.occlusion_points(pts[:0], ntheta, nphi, seed)) == 0            # empty batch


def test_point_ao_empty_scene():
    _need_gpu()
    a = accel.Accel.bind().build(np.zeros((0, 3, 3)), accel.PREC_F32)
    pts = np.array([[0.5, 0.5, 0.5, 0.0, 0.0, 1.0]] * 5)
    assert not a.occlusion_points(pts, 4, 4, 1).any()


def test_config2_full_size_batch_against_the_oracle():
    """BASELINE configs[2] at full size: 1 M triangles, the bench's 16 Mi-ray batch (262 144 points x 8 x 8)."""
    _need_gpu()
    import torch
    NT, NP, N = 1_000_000, 262_144, 64
    tris = scenes.triangle_soup(NT, scenes.SEED_C3)
    a = accel.Accel.bind().build(tris, accel.PREC_F32)                       # device-built tree (>= 32 Ki triangles)
    orc = ol.Oracle().build(tris)
    assert np.array_equal(a.triorder(), orc.triorder())
    pts = _primary_points(a, tris[a.triorder()], NP)
    assert len(pts) == NP
    rays = a.ao_point_rays(pts, 8, 8, scenes.SEED_C3)
    assert len(rays) == NP * N == 16_777_216
    # (1) the generated batch is the restatement's: three windows of 4096 points (head, middle, tail of the batch)
    for p0 in (0, NP // 2 - 2048, NP - 4096):
        want_rays = ol.Oracle().ao_point_rays(pts[p0:p0 + 4096], 8, 8, scenes.SEED_C3, first_point=p0)
        assert np.array_equal(rays[p0 * N:(p0 + 4096) * N].view(np.uint32), want_rays.view(np.uint32)), p0
    # (2) whole-batch verdicts on the device, resident rays
    d_rays = torch.from_numpy(rays).cuda()
    d_occ = torch.empty(len(rays), dtype=torch.uint8, device="cuda")
    a.occluded_dev(d_rays, len(rays), d_occ)
    torch.cuda.synchronize()
    occ = d_occ.cpu().numpy()
    # (3) >= 1 Mi rays sampled across the WHOLE batch, bit-exact against the oracle's fp32 instantiation: occlusion and closest hit
    sample = np.sort(np.random.default_rng(7).choice(len(rays), 1 << 20, replace=False))
    sample = np.unique(np.r_[sample, 0:4096, len(rays) - 4096:len(rays)])
    assert len(sample) >= (1 << 20)
    want_occ = orc.occluded_f32(rays[sample])
    assert np.array_equal(occ[sample] != 0, want_occ != 0)
    d_hits = torch.empty((len(sample), 4), dtype=torch.float32, device="cuda")
    d_sr = torch.from_numpy(np.ascontiguousarray(rays[sample])).cuda()
    a.intersect_dev(d_sr, len(sample), d_hits)
    torch.cuda.synchronize()
    hits = d_hits.cpu().numpy()
    want = orc.intersect_f32(rays[sample])
    assert np.array_equal(hits[:, 0].view(np.uint32), want["t"].view(np.uint32))
    assert np.array_equal(hits[:, 1].view(np.uint32), want["u"].view(np.uint32))
    assert np.array_equal(hits[:, 2].view(np.uint32), want["v"].view(np.uint32))
    assert np.array_equal(hits[:, 3].view(np.uint32), want["prim"])
    assert np.array_equal(want["prim"] != accel.MISS_PRIM, want_occ != 0)    # the oracle's own consistency
    # (4) checksum of checksums over ALL 16 Mi verdicts: the point entry's counts == per-point sums of the occlusion bytes
    counts = a.occlusion_points(pts, 8, 8, scenes.SEED_C3)
    assert np.array_equal(counts, occ.reshape(NP, N).sum(axis=1, dtype=np.uint32))
    assert 0.6 < occ.mean() < 0.75


def test_config4_ten_million_triangles():
    """BASELINE configs[4] scene: 10 M-triangle soup.  Device-built tree == host-built tree (nodes, boxes, leaf order); closest hits and
    occlusion of rays sampled over the camera against the oracle, whose tree is the compiled reference's."""
    _need_gpu()
    NT = 10_000_000
    tris = scenes.triangle_soup(NT, scenes.SEED_C5)
    dev = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.BUILD_DEVICE)
    host = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.HOST_ONLY)
    assert np.array_equal(dev.triorder(), host.triorder())
    nd, nh = dev.nodes(), host.nodes()
    assert len(nd) == len(nh)
    for f in nd.dtype.names:
        assert np.array_equal(nd[f], nh[f]), f
    del host, nh
    orc = ol.Oracle().build(tris)
    assert np.array_equal(dev.triorder(), orc.triorder())
    rays = scenes.pinhole_rays(1024, 1024)
    pick = np.sort(np.random.default_rng(11).choice(len(rays), 200_000, replace=False))
    hits = dev.intersect(rays[pick])
    want = orc.intersect_f32(rays[pick])
    for f in ("t", "u", "v", "prim"):
        assert np.array_equal(hits[f], want[f]), f
    assert np.array_equal(dev.occluded(rays[pick]) != 0, want["prim"] != accel.MISS_PRIM)
    pts = _primary_points(dev, tris[dev.triorder()], 4096)
    ao = dev.ao_point_rays(pts, 8, 8, scenes.SEED_C5)
    counts = dev.occlusion_points(pts, 8, 8, scenes.SEED_C5)
    assert np.array_equal(counts, orc.occluded_f32(ao).reshape(len(pts), 64).sum(axis=1).astype(np.uint32))


def test_config2_hybrid_double_exact_occlusion_full_size():
    """csrc/hybrid.cuh at BASELINE configs[2] size: 1 M triangles, 4 Mi double AO rays (origins P + 1e-6 N in full double).  The hybrid
    kernel's verdicts equal the plain double kernel's on EVERY ray and the oracle's f64 instantiation on 300 000 rays sampled across the
    batch; the double point entry's counts are the per-point sums of those verdicts."""
    _need_gpu()
    import os
    import torch
    NT, NP, N = 1_000_000, 65_536, 64
    tris = scenes.triangle_soup(NT, scenes.SEED_C3)
    a = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.PREC_F64)
    orc = ol.Oracle().build(tris)
    pts = _primary_points(a, tris[a.triorder()], NP)
    rays = ol.Oracle().ao_point_rays_f64(pts, 8, 8, scenes.SEED_C3)
    assert len(rays) == NP * N
    d_rays = torch.from_numpy(rays).cuda()
    d_h = torch.empty(len(rays), dtype=torch.uint8, device="cuda")
    d_p = torch.empty(len(rays), dtype=torch.uint8, device="cuda")
    a.occluded_dev(d_rays, len(rays), d_h, f64=True)                        # hybrid (both record sets resident)
    os.environ["B200_HYBRID"] = "0"
    try:
        a.occluded_dev(d_rays, len(rays), d_p, f64=True)                    # the plain double kernel
    finally:
        os.environ.pop("B200_HYBRID", None)
    torch.cuda.synchronize()
    assert torch.equal(d_h, d_p)
    occ = d_h.cpu().numpy()
    sample = np.sort(np.random.default_rng(5).choice(len(rays), 300_000, replace=False))
    assert np.array_equal(occ[sample] != 0, orc.occluded_f64(rays[sample]) != 0)
    counts = a.occlusion_points(pts, 8, 8, scenes.SEED_C3, f64=True)
    assert np.array_equal(counts, occ.reshape(NP, N).sum(axis=1, dtype=np.uint32))
    assert 0.5 < occ.mean() < 0.75
    # closest hit on the same 4 Mi double rays: closest_hybrid_kernel's records == the double kernel's, bit for bit, and == the oracle's
    h_h = torch.empty((len(rays), 4), dtype=torch.float64, device="cuda")
    h_p = torch.empty((len(rays), 4), dtype=torch.float64, device="cuda")
    a.intersect_dev(d_rays, len(rays), h_h, f64=True)
    os.environ["B200_HYBRID_CLOSEST"] = "0"
    try:
        a.intersect_dev(d_rays, len(rays), h_p, f64=True)
    finally:
        os.environ.pop("B200_HYBRID_CLOSEST", None)
    torch.cuda.synchronize()
    assert torch.equal(h_h.view(torch.int64), h_p.view(torch.int64))
    pick = sample[:100_000]
    got = h_h.cpu().numpy()[pick]
    want = orc.intersect_f64(rays[pick])
    assert np.array_equal(got[:, 0], want["t"]) and np.array_equal(got[:, 1], want["u"]) and np.array_equal(got[:, 2], want["v"])
    assert np.array_equal(got[:, 3].view(np.uint64) & 0xFFFFFFFF, want["prim"].astype(np.uint64))
    assert np.array_equal(got[:, 3].view(np.uint64) >> 32, want["hit"].astype(np.uint64))
