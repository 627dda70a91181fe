"""Synthetic workloads of BASELINE.json: seeded triangle soups and ray batches.

Host-side numpy only; shared by tests, bench.py and the golden-vector generator so that the
oracle, the compiled reference and the CUDA path always see byte-identical inputs.

All geometry and rays are generated in float64, rounded to float32 and widened again, so the
double-precision reference and the fp32 kernels consume exactly the same numbers.
"""
from __future__ import annotations

import math

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
SEED_C2 = 0xB2000002
SEED_C3 = 0xB2000003
SEED_C5 = 0xB2000005


def splitmix64(x: np.ndarray) -> np.ndarray:
    """SplitMix64 finaliser applied to ``x + GOLDEN`` (uint64, wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, start: int, count: int) -> np.ndarray:
    """``count`` doubles in [0,1): element k is splitmix64(seed + (start+k)*GOLDEN) >> 11, scaled by 2^-53."""
    with np.errstate(over="ignore"):
        k = np.arange(start, start + count, dtype=np.uint64)
        z = splitmix64(np.uint64(seed) + k * GOLDEN)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def triangle_soup(ntris: int, seed: int, chunk: int = 1 << 20) -> np.ndarray:
    """Random small triangles in the unit cube, shape [ntris, 3, 3] float64 (fp32-representable).

    centre c ~ U[0,1)^3, edges e1,e2 ~ U[-s,s]^3 with s = 0.5 * ntris^(-1/3); vertices c, c+e1, c+e2.
    """
    s = 0.5 * ntris ** (-1.0 / 3.0)
    out = np.empty((ntris, 3, 3), dtype=np.float64)
    for lo in range(0, ntris, chunk):
        n = min(chunk, ntris - lo)
        u = uniform01(seed, 9 * lo, 9 * n).reshape(n, 9)
        c = u[:, 0:3]
        e1 = (2.0 * u[:, 3:6] - 1.0) * s
        e2 = (2.0 * u[:, 6:9] - 1.0) * s
        out[lo:lo + n, 0] = c
        out[lo:lo + n, 1] = c + e1
        out[lo:lo + n, 2] = c + e2
    return out.astype(np.float32).astype(np.float64)


def box_city(nx: int, nz: int, seed: int) -> np.ndarray:
    """An architectural test scene in the unit cube, [ntris, 3, 3] float64 (fp32-representable): an nx x nz grid of axis-aligned boxes
    on a ground of axis-aligned quads, coordinates snapped to a 1/64 lattice.  Unlike the soup it is full of what real meshes have:
    shared vertices, coplanar faces, triangles whose boxes have zero extent on one axis, many triangles with EQUAL box bounds
    (so the builder's bins, SAH costs and partition predicate see exact ties everywhere) and rays that graze edges."""
    u = uniform01(seed, 0, 4 * nx * nz).reshape(nx * nz, 4)
    tris = []

    def quad(a, b, c, d):
        tris.append((a, b, c))
        tris.append((a, c, d))

    for i in range(nx):
        for k in range(nz):
            x0, x1, z0, z1 = i / nx, (i + 1) / nx, k / nz, (k + 1) / nz
            quad((x0, 0, z0), (x1, 0, z0), (x1, 0, z1), (x0, 0, z1))                       # ground tile
            r = u[i * nz + k]
            w = np.floor(r[0] * 24 + 8) / 64.0 / max(nx, nz) * 2.0                         # footprint, lattice-snapped
            h = np.floor(r[1] * 40 + 4) / 64.0                                              # height
            cx = np.floor((x0 + (x1 - x0) * (0.25 + 0.5 * r[2])) * 64) / 64.0
            cz = np.floor((z0 + (z1 - z0) * (0.25 + 0.5 * r[3])) * 64) / 64.0
            a, b, c, d = (cx, 0, cz), (cx + w, 0, cz), (cx + w, 0, cz + w), (cx, 0, cz + w)
            e, f, g, hh = (cx, h, cz), (cx + w, h, cz), (cx + w, h, cz + w), (cx, h, cz + w)
            quad(e, f, g, hh)                                                              # roof
            quad(a, b, f, e); quad(b, c, g, f); quad(c, d, hh, g); quad(d, a, e, hh)       # walls
    return np.asarray(tris, dtype=np.float64).astype(np.float32).astype(np.float64)


def pinhole_rays(width: int, height: int, eye=(0.5, 0.5, -2.0), fov_deg: float = 40.0) -> np.ndarray:
    """Row-major pixel-centre rays looking down +z, shape [h*w, 8] float32: ox,oy,oz,tmin,dx,dy,dz,tmax.

    Pixel centres sit at half-integers, so no direction component is exactly zero (bug 1).
    """
    tanh = math.tan(math.radians(fov_deg) * 0.5)
    xs = ((np.arange(width, dtype=np.float64) + 0.5) - 0.5 * width) / (0.5 * width) * tanh
    ys = ((np.arange(height, dtype=np.float64) + 0.5) - 0.5 * height) / (0.5 * height) * tanh
    dx, dy = np.meshgrid(xs, ys)
    d = np.stack([dx, dy, np.ones_like(dx)], axis=-1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((width * height, 8), dtype=np.float32)
    rays[:, 0:3] = np.asarray(eye, dtype=np.float32)
    rays[:, 3] = 0.0
    rays[:, 4:7] = d.astype(np.float32)
    rays[:, 7] = 1.0e38
    return rays


def ortho_basis(n: np.ndarray):
    """Vectorised ri_ortho_basis (reference: src/render/reflection.c:312-333), float64."""
    n = np.asarray(n, dtype=np.float64)
    small = (n < 0.6) & (n > -0.6)
    idx = np.where(small.any(axis=1), small.argmax(axis=1), 0)
    b1 = np.zeros_like(n)
    b1[np.arange(len(n)), idx] = 1.0
    b0 = np.cross(b1, n)
    l0 = np.einsum("ij,ij->i", b0, b0)
    b0 = np.where((l0 > 1e-17)[:, None], b0 / np.sqrt(np.where(l0 > 0, l0, 1.0))[:, None], b0)
    b1 = np.cross(n, b0)
    l1 = np.einsum("ij,ij->i", b1, b1)
    b1 = np.where((l1 > 1e-17)[:, None], b1 / np.sqrt(np.where(l1 > 0, l1, 1.0))[:, None], b1)
    return b0, b1, n


def ao_rays(points: np.ndarray, normals: np.ndarray, ntheta: int, nphi: int, seed: int,
            eps: float = 1.0e-6) -> np.ndarray:
    """Stratified cosine-weighted hemisphere rays about each (P, Ns) pair, shape [npts*nphi*ntheta, 8] float32.

    Sample loop of ambientocclusion.c:83-117 (outer j over phi, inner i over theta, draw z0 then z1) with the
    counter-based uniform ``u(point, j, i, k) = uniform01(seed)[ (point*nphi*ntheta + j*ntheta + i)*2 + k ]``
    replacing randomMT2.  Origin = P + eps*Ns.
    """
    points = np.asarray(points, dtype=np.float64)
    normals = np.asarray(normals, dtype=np.float64)
    npts = len(points)
    per = ntheta * nphi
    b0, b1, b2 = ortho_basis(normals)
    out = np.zeros((npts * per, 8), dtype=np.float32)
    org = points + normals * eps
    chunk = max(1, (1 << 22) // per)
    for lo in range(0, npts, chunk):
        n = min(chunk, npts - lo)
        u = uniform01(seed, 2 * per * lo, 2 * per * n).reshape(n, nphi, ntheta, 2)
        i = np.arange(ntheta, dtype=np.float64)[None, None, :]
        j = np.arange(nphi, dtype=np.float64)[None, :, None]
        z0 = (i + u[..., 0]) / float(ntheta)
        z1 = (j + u[..., 1]) / float(nphi)
        ct = np.sqrt(z0)
        phi = 2.0 * math.pi * z1
        lx = np.cos(phi) * ct
        ly = np.sin(phi) * ct
        lz = np.sqrt(1.0 - ct * ct)
        d = (lx[..., None] * b0[lo:lo + n, None, None, :]
             + ly[..., None] * b1[lo:lo + n, None, None, :]
             + lz[..., None] * b2[lo:lo + n, None, None, :])
        sl = slice(lo * per, (lo + n) * per)
        out[sl, 0:3] = np.repeat(org[lo:lo + n], per, axis=0).astype(np.float32)
        out[sl, 4:7] = d.reshape(-1, 3).astype(np.float32)
    out[:, 3] = 0.0
    out[:, 7] = 1.0e38
    return out


def rays_f32_to_f64(rays: np.ndarray) -> np.ndarray:
    """[n,8] float32 ray records -> [n,6] float64 (org, dir) with identical values."""
    rays = np.asarray(rays, dtype=np.float32)
    return np.ascontiguousarray(np.concatenate([rays[:, 0:3], rays[:, 4:7]], axis=1).astype(np.float64))


def random_beams(n: int, seed: int, eye=(0.5, 0.5, -2.0), spread: float = 0.35, width: float = 0.02) -> np.ndarray:
    """Small square frusta from ``eye`` (the testbed's use of beams: one beam per pixel block, simplerender.cpp:540-580),
    shape [n, 15] float64 = org.xyz + 4 corner directions in the winding ri_beam_set expects (consecutive corners)."""
    u = uniform01(seed, 0, 3 * n).reshape(n, 3)
    cx = (2.0 * u[:, 0] - 1.0) * spread
    cy = (2.0 * u[:, 1] - 1.0) * spread
    w = width * (0.25 + u[:, 2])
    out = np.zeros((n, 15), dtype=np.float64)
    out[:, 0:3] = np.asarray(eye, dtype=np.float64)
    corners = [(-1, -1), (1, -1), (1, 1), (-1, 1)]
    for j, (sx, sy) in enumerate(corners):
        out[:, 3 + 3 * j + 0] = cx + sx * w
        out[:, 3 + 3 * j + 1] = cy + sy * w
        out[:, 3 + 3 * j + 2] = 1.0
    return out
