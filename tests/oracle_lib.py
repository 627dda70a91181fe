"""ctypes bindings for the two checkers (TEST INFRASTRUCTURE):

* ``Oracle``    -- oracle/liblucille_oracle.so, the CPU restatement (oracle/lucille_oracle.c)
* ``Reference`` -- oracle/_ref/libluciref{,_stat}.so, the compiled unmodified reference (oracle/build_ref.sh)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liblucille_oracle.so")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_SO = os.path.join(REF_DIR, "libluciref.so")
REF_STAT_SO = os.path.join(REF_DIR, "libluciref_stat.so")
ORACLE_RIB = os.path.join(REF_DIR, "oracle_rib")

NODE_DTYPE = np.dtype([("is_leaf", "<i4"), ("axis", "<i4"), ("child0", "<i8"), ("child1", "<i8"),
                       ("tri_start", "<i8"), ("ntris", "<i8"), ("lbox", "<f8", 6), ("rbox", "<f8", 6)])
HIT64_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"), ("prim", "<u4"), ("hit", "<u4")])
HIT32_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])
STATE_DTYPE = np.dtype([("P", "<f8", 3), ("Ng", "<f8", 3), ("Ns", "<f8", 3), ("tangent", "<f8", 3), ("binormal", "<f8", 3)])
COUNTERS_DTYPE = np.dtype([("nrays", "<u8"), ("ninner", "<u8"), ("nleaf", "<u8"), ("ntris", "<u8"), ("nhit_tris", "<u8")])
REFHIT_DTYPE = np.dtype([("hit", "<i4"), ("index", "<u4"), ("geom_id", "<u4"), ("pad", "<u4"),
                         ("t", "<f8"), ("u", "<f8"), ("v", "<f8"),
                         ("P", "<f8", 3), ("Ng", "<f8", 3), ("Ns", "<f8", 3), ("tangent", "<f8", 3), ("binormal", "<f8", 3)])
TRACE_REC_DTYPE = np.dtype([("Cs", "<f8", 3), ("P", "<f8", 3), ("N", "<f8", 3), ("Ng", "<f8", 3), ("dPdu", "<f8", 3), ("dPdv", "<f8", 3),
                            ("I", "<f8", 3), ("dst", "<f8", 3), ("s", "<f4"), ("t", "<f4"), ("called", "<i4"), ("ray_depth", "<i4")])
STATE_EXT_DTYPE = np.dtype([("E", "<f8", 3), ("I", "<f8", 3), ("color", "<f8", 3), ("st", "<f8", 2), ("t", "<f8"), ("inside", "<i4"), ("hit", "<i4")])
MISS_PRIM = 0xFFFFFFFF


class FrameParams(C.Structure):
    _fields_ = [("c2w", C.c_double * 16), ("flength", C.c_double), ("is_rh", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("xsamples", C.c_int32), ("ysamples", C.c_int32),
                ("ntheta", C.c_int32), ("nphi", C.c_int32), ("bucket_size", C.c_int32)]


class PathFrameParams(C.Structure):
    _fields_ = [("c2w", C.c_double * 16), ("flength", C.c_double), ("is_rh", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("spp", C.c_int32), ("max_vertices", C.c_int32), ("seed", C.c_uint32), ("kd", C.c_double), ("Le", C.c_double),
                ("rank", C.c_int32), ("world", C.c_int32), ("bucket_size", C.c_int32)]


class SunskyBlock(C.Structure):
    """orc_sunsky_t == ri_b200_sunsky_t (include/lucille_b200.h)."""
    _fields_ = [("sun_theta", C.c_float), ("sun_phi", C.c_float),
                ("perez_x", C.c_float * 5), ("perez_y", C.c_float * 5), ("perez_Y", C.c_float * 5),
                ("zenith_x", C.c_float), ("zenith_y", C.c_float), ("zenith_Y", C.c_float),
                ("S0", C.c_float * 41), ("S1", C.c_float * 41), ("S2", C.c_float * 41),
                ("cie", C.c_float * 243), ("cs", C.c_float * 8),
                ("nsun", C.c_int32), ("pad", C.c_int32),
                ("sun_dir", C.c_double * 12), ("sun_col", C.c_double * 12)]


def attribute_case(ntris: int, geom_sizes, seed: int):
    """Seeded per-corner colours / st and per-geom flags (1 Cs, 2 shared st, 4 unshared st, 8 two-sided) for the hit-state tests,
    with the per-triangle presence / back-side arrays the oracle and the product take."""
    rng = np.random.default_rng(seed)
    colors = rng.uniform(0.0, 1.0, (ntris, 3, 3))
    st = rng.uniform(-2.0, 3.0, (ntris, 3, 2))
    flags = np.array([(1 | 2), 0, (4 | 8), (1 | 8), 2][: len(geom_sizes)], dtype=np.uint8)
    has_color, has_st, inside = np.zeros(ntris, np.uint8), np.zeros(ntris, np.uint8), np.zeros(ntris, np.uint8)
    off = 0
    for n, f in zip(geom_sizes, flags):
        has_color[off:off + n] = 1 if f & 1 else 0
        has_st[off:off + n] = 1 if f & 6 else 0
        if f & 8:
            inside[off + (3 * n // 2 + 2) // 3: off + n] = 1        # index >= nindices / 2, index = 3 * triangle
        off += n
    return colors, st, flags, has_color, has_st, inside


def test_texture(w: int = 23, h: int = 17, seed: int = 3) -> np.ndarray:
    """A seeded float RGBA image (not a power of two, not square) for the material-texture tests."""
    rng = np.random.default_rng(seed)
    return rng.uniform(0.0, 2.0, (h, w, 4)).astype(np.float32)


test_texture.__test__ = False


def read_attr(path: str, ntris: int):
    raw = np.fromfile(path, dtype=np.uint8)
    st = raw[: 48 * ntris].view("<f8").reshape(ntris, 3, 2).copy()
    has_st = raw[48 * ntris: 49 * ntris].copy()
    return st, has_st


def hdr_cases():
    """Float framebuffers for the .hdr output tests: noise with negative components, long runs, a width below 8 (flat pixels),
    values under the 1e-32 cut-off, exactly 8 columns, a constant image."""
    rng = np.random.default_rng(1)
    return {"noise": rng.uniform(-0.2, 3, (37, 53, 3)).astype(np.float32),
            "runs": np.repeat(rng.uniform(0, 1, (16, 9, 3)), 17, axis=1).astype(np.float32),
            "narrow": rng.uniform(0, 1, (5, 7, 3)).astype(np.float32),
            "tiny": (rng.uniform(0, 1, (9, 40, 3)) * 1e-30).astype(np.float32),
            "w8": rng.uniform(0, 1e4, (3, 8, 3)).astype(np.float32),
            "const": np.full((10, 300, 3), 0.25, np.float32)}


def sky_dirs(n: int, seed: int) -> np.ndarray:
    """Seeded unit directions for the sky-lookup tests, with grazing ones (the t[2] < 0.001 branch of sunsky.c:349-355),
    below-horizon ones and the axes."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[: n // 8, 1] = np.abs(d[: n // 8, 1]) * 1e-3
    d[n // 8: n // 8 + 4] = [[0, 1, 0], [1, 0, 0], [0, -1, 0], [0, 0, 1]]
    return d.astype(np.float32)


def sunsky_block(params45: np.ndarray, tables) -> SunskyBlock:
    """params45: the lref_frame_sunsky()/lref_sunsky_eval() record; tables: mapping with S0 S1 S2 cie cs
    (tests/golden/sunsky.npz, extracted from the reference by tests/golden/make_sunsky_golden.py)."""
    b = SunskyBlock()
    p = np.asarray(params45, dtype=np.float64)
    b.sun_theta, b.sun_phi = float(p[0]), float(p[1])
    for i in range(5):
        b.perez_x[i], b.perez_y[i], b.perez_Y[i] = float(p[2 + i]), float(p[7 + i]), float(p[12 + i])
    b.zenith_x, b.zenith_y, b.zenith_Y = float(p[17]), float(p[18]), float(p[19])
    for name in ("S0", "S1", "S2"):
        arr = np.asarray(tables[name], dtype=np.float32)
        for i in range(41):
            getattr(b, name)[i] = float(arr[i])
    cie = np.asarray(tables["cie"], dtype=np.float32).reshape(243)
    for i in range(243):
        b.cie[i] = float(cie[i])
    cs = np.asarray(tables["cs"], dtype=np.float32)
    for i in range(8):
        b.cs[i] = float(cs[i])
    n = int(p[20]) if len(p) > 20 else 0
    b.nsun = n
    for l in range(n):
        for k in range(3):
            b.sun_dir[3 * l + k] = float(p[21 + 6 * l + k])
            b.sun_col[3 * l + k] = float(p[21 + 6 * l + 3 + k])
    return b


REFERENCE_MESH_DIR = "/root/reference/src/testbed"           # read where it lies; never copied (absent on the GPU box -> tests skip)


def load_obj_triangles(path: str) -> np.ndarray:
    """Triangles [n,3,3] (float64) of a Wavefront OBJ: `v x y z` and `f a b c ...` with a/b/c, a//c or a forms, polygons fanned
    from their first corner.  For the real-mesh parity tests (the reference's testbed meshes: shared vertices, axis-aligned walls,
    degenerate faces)."""
    verts, faces = [], []
    with open(path, "r", errors="replace") as f:
        for line in f:
            if line.startswith("v "):
                verts.append([float(x) for x in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    faces.append((idx[0], idx[k], idx[k + 1]))
    v = np.asarray(verts, dtype=np.float64)
    return v[np.asarray(faces, dtype=np.int64)]


def oracle_available() -> bool:
    return os.path.exists(ORACLE_SO)


def reference_available() -> bool:
    return os.path.exists(REF_SO)


class _quiet:
    """The reference printf()s from its setup code: keep fd 1 clean while it runs."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *exc):
        C.CDLL(None).fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleTree:
    def __init__(self, lib, handle, ntris):
        self.lib, self.h, self.ntris = lib, handle, ntris

    def __del__(self):
        try:
            self.lib.orc_free(self.h)
        except Exception:
            pass

    @property
    def empty(self):
        return bool(self.lib.orc_is_empty(self.h))

    def nodes(self) -> np.ndarray:
        n = self.lib.orc_num_nodes(self.h)
        out = np.zeros(n, dtype=NODE_DTYPE)
        if n:
            self.lib.orc_get_nodes(self.h, _ptr(out))
        return out

    def triorder(self) -> np.ndarray:
        out = np.zeros(self.ntris, dtype=np.uint32)
        self.lib.orc_get_triorder(self.h, _ptr(out))
        return out

    def bbox(self):
        a, b = np.zeros(3), np.zeros(3)
        self.lib.orc_scene_bbox(self.h, _ptr(a), _ptr(b))
        return a, b

    def max_depth(self):
        return self.lib.orc_max_depth(self.h)

    def set_normals(self, tri_normals):
        if tri_normals is None:
            self.lib.orc_set_normals(self.h, None)
        else:
            n = np.ascontiguousarray(tri_normals, dtype=np.float64).reshape(-1, 9)
            assert len(n) == self.ntris
            self.lib.orc_set_normals(self.h, _ptr(n))

    def intersect_f64(self, rays6: np.ndarray, counters: bool = False):
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64)
        out = np.zeros(len(rays6), dtype=HIT64_DTYPE)
        cnt = np.zeros(1, dtype=COUNTERS_DTYPE)
        self.lib.orc_intersect_f64(self.h, _ptr(rays6), C.c_uint64(len(rays6)), _ptr(out), _ptr(cnt) if counters else None)
        return (out, cnt[0]) if counters else out

    def intersect_f32(self, rays8: np.ndarray, counters: bool = False):
        rays8 = np.ascontiguousarray(rays8, dtype=np.float32)
        out = np.zeros(len(rays8), dtype=HIT32_DTYPE)
        cnt = np.zeros(1, dtype=COUNTERS_DTYPE)
        self.lib.orc_intersect_f32(self.h, _ptr(rays8), C.c_uint64(len(rays8)), _ptr(out), _ptr(cnt) if counters else None)
        return (out, cnt[0]) if counters else out

    def occluded_f64(self, rays6: np.ndarray, counters: bool = False):
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64)
        out = np.zeros(len(rays6), dtype=np.uint8)
        cnt = np.zeros(1, dtype=COUNTERS_DTYPE)
        self.lib.orc_occluded_f64(self.h, _ptr(rays6), C.c_uint64(len(rays6)), _ptr(out), _ptr(cnt) if counters else None)
        return (out, cnt[0]) if counters else out

    def occluded_f32(self, rays8: np.ndarray, counters: bool = False):
        rays8 = np.ascontiguousarray(rays8, dtype=np.float32)
        out = np.zeros(len(rays8), dtype=np.uint8)
        cnt = np.zeros(1, dtype=COUNTERS_DTYPE)
        self.lib.orc_occluded_f32(self.h, _ptr(rays8), C.c_uint64(len(rays8)), _ptr(out), _ptr(cnt) if counters else None)
        return (out, cnt[0]) if counters else out

    def state_build(self, rays6: np.ndarray, hits: np.ndarray) -> np.ndarray:
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64)
        out = np.zeros(len(rays6), dtype=STATE_DTYPE)
        self.lib.orc_state_build_f64(self.h, _ptr(rays6), _ptr(hits), C.c_uint64(len(rays6)), _ptr(out))
        return out

    def set_attributes(self, colors=None, has_color=None, st=None, has_st=None, inside=None):
        """Per-corner colours [n,3,3] / texture coordinates [n,3,2] with per-triangle presence flags, and the back-side flag."""
        def arr(a, dt, shape):
            return None if a is None else np.ascontiguousarray(a, dtype=dt).reshape(shape)
        self._attr = [arr(colors, np.float64, (-1, 9)), arr(has_color, np.uint8, -1), arr(st, np.float64, (-1, 6)), arr(has_st, np.uint8, -1),
                      arr(inside, np.uint8, -1)]
        self.lib.orc_set_attributes(self.h, *[None if a is None else _ptr(a) for a in self._attr])

    def state_ext(self, rays6: np.ndarray, hits: np.ndarray) -> np.ndarray:
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(len(rays6), dtype=STATE_EXT_DTYPE)
        self.lib.orc_state_ext_build_f64(self.h, _ptr(rays6), _ptr(np.ascontiguousarray(hits)), C.c_uint64(len(rays6)), _ptr(out))
        return out

    def beam_visibility(self, beams15: np.ndarray) -> np.ndarray:
        beams15 = np.ascontiguousarray(beams15, dtype=np.float64).reshape(-1, 15)
        out = np.zeros(len(beams15), dtype=np.int32)
        self.lib.orc_beam_visibility(self.h, _ptr(beams15), C.c_uint64(len(beams15)), _ptr(out))
        return out

    def render_pathtrace(self, frame: "PathFrameParams"):
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        nrays = C.c_uint64(0)
        self.lib.orc_render_pathtrace(self.h, C.byref(frame), _ptr(rgb), C.byref(nrays))
        return rgb, nrays.value

    def render_ao(self, frame: "FrameParams"):
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        nrays = C.c_uint64(0)
        self.lib.orc_render_ao(self.h, C.byref(frame), _ptr(rgb), C.byref(nrays))
        return rgb, nrays.value

    def transport_batch(self, which: int, rays6: np.ndarray, ntheta: int = 8, nphi: int = 8) -> np.ndarray:
        """Radiance per eye ray of a transport (0 ambient occlusion, 1 dirt map), MT19937 stream seeded 4357 and consumed in ray order."""
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros((len(rays6), 3), dtype=np.float64)
        self.lib.orc_transport_batch(self.h, which, ntheta, nphi, _ptr(rays6), C.c_uint64(len(rays6)), _ptr(out))
        return out

    def point_gather(self, kind: int, nsamples: int, points6: np.ndarray, env=None, col=(1.0, 1.0, 1.0), intensity=1.0, seed=4357):
        """Hemisphere gather at shading points (P, N): 0 occlusion() shadeop, 1 ri_ibl_sample_cosweight, 2 ri_domelight_sample."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        env = None if env is None else np.ascontiguousarray(env, dtype=np.float32)
        c = np.ascontiguousarray(col, dtype=np.float64)
        out = np.zeros((len(pts), 3), dtype=np.float64)
        nrays = C.c_uint64(0)
        self.lib.orc_point_gather(self.h, kind, nsamples, seed, _ptr(pts), C.c_uint64(len(pts)), None if env is None else _ptr(env),
                                  0 if env is None else env.shape[1], 0 if env is None else env.shape[0], _ptr(c), C.c_double(intensity),
                                  _ptr(out), C.byref(nrays))
        return out, nrays.value

    def shade_trace(self, pr6: np.ndarray, env=None, use_env: bool = True):
        """trace() shadeop up to the shader call (shader.c:895-976) for (P, R) pairs: dict(rays, hits, states, exts, eye, miss_rgb)."""
        pr = np.ascontiguousarray(pr6, dtype=np.float64).reshape(-1, 6)
        env = None if env is None else np.ascontiguousarray(env, dtype=np.float32)
        n = len(pr)
        rays = np.zeros((n, 6)); hits = np.zeros(n, dtype=HIT64_DTYPE); states = np.zeros(n, dtype=STATE_DTYPE)
        exts = np.zeros(n, dtype=STATE_EXT_DTYPE); eye = np.zeros((n, 3)); miss = np.zeros((n, 3))
        self.lib.orc_shade_trace(self.h, _ptr(pr), C.c_uint64(n), None if env is None else _ptr(env), 0 if env is None else env.shape[1],
                                 0 if env is None else env.shape[0], int(use_env), _ptr(rays), _ptr(hits), _ptr(states), _ptr(exts),
                                 _ptr(eye), _ptr(miss))
        return dict(rays=rays, hits=hits, states=states, exts=exts, eye=eye, miss_rgb=miss)

    def light_samples(self, nsamples: int, angle: float, points6: np.ndarray, env, seed: int = 4357):
        """Light samples of next_lightsource() (shader.c:1116-1310) at points (P, N): (L [n,m,3], Cl [n,m,3], visible [n,m] u8, rays)."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        env = np.ascontiguousarray(env, dtype=np.float32)
        nt = max(1, int(math.sqrt(int(nsamples / 3.0))))
        m = nt * 3 * nt
        L = np.zeros((len(pts), m, 3)); Cl = np.zeros((len(pts), m, 3)); vis = np.zeros((len(pts), m), dtype=np.uint8)
        nrays = C.c_uint64(0)
        got = self.lib.orc_light_samples(self.h, nsamples, C.c_double(angle), seed, _ptr(pts), C.c_uint64(len(pts)), _ptr(env),
                                         env.shape[1], env.shape[0], _ptr(L), _ptr(Cl), _ptr(vis), C.byref(nrays))
        assert got == m
        return L, Cl, vis, nrays.value

    def point_gather_qmc(self, kind: int, nsamples: int, points6: np.ndarray, instance=None, dim: int = 0, env=None,
                         col=(1.0, 1.0, 1.0), intensity=1.0):
        """The quasi-Monte Carlo branches (Option "use_qmc") of the IBL (1) and dome-light (2) gathers; instance = inray->i per point."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        env = None if env is None else np.ascontiguousarray(env, dtype=np.float32)
        inst = None if instance is None else np.ascontiguousarray(instance, dtype=np.int32)
        c = np.ascontiguousarray(col, dtype=np.float64)
        out = np.zeros((len(pts), 3), dtype=np.float64)
        nrays = C.c_uint64(0)
        self.lib.orc_point_gather_qmc(self.h, kind, nsamples, _ptr(pts), C.c_uint64(len(pts)), None if inst is None else _ptr(inst), dim,
                                      None if env is None else _ptr(env), 0 if env is None else env.shape[1],
                                      0 if env is None else env.shape[0], _ptr(c), C.c_double(intensity), _ptr(out), C.byref(nrays))
        return out, nrays.value

    def transport_whitted(self, rays6: np.ndarray, env):
        """Radiance per eye ray of the Whitted refraction tracer with the angular-map environment ``env`` ([h,w,4] float32 or None)."""
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64).reshape(-1, 6)
        env = None if env is None else np.ascontiguousarray(env, dtype=np.float32)
        out = np.zeros((len(rays6), 3), dtype=np.float64)
        nrays = C.c_uint64(0)
        self.lib.orc_transport_whitted(self.h, None if env is None else _ptr(env), 0 if env is None else env.shape[1],
                                       0 if env is None else env.shape[0], _ptr(rays6), C.c_uint64(len(rays6)), _ptr(out), C.byref(nrays))
        return out, nrays.value

    def render_whitted(self, frame: "FrameParams", env):
        env = None if env is None else np.ascontiguousarray(env, dtype=np.float32)
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        nrays = C.c_uint64(0)
        self.lib.orc_render_whitted(self.h, C.byref(frame), None if env is None else _ptr(env), 0 if env is None else env.shape[1],
                                    0 if env is None else env.shape[0], _ptr(rgb), C.byref(nrays))
        return rgb, nrays.value

    def render_hitmask(self, frame: "FrameParams"):
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        nrays = C.c_uint64(0)
        self.lib.orc_render_hitmask(self.h, C.byref(frame), _ptr(rgb), C.byref(nrays))
        return rgb, nrays.value

    def render_dirtmap(self, frame: "FrameParams"):
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        nrays = C.c_uint64(0)
        self.lib.orc_render_dirtmap(self.h, C.byref(frame), _ptr(rgb), C.byref(nrays))
        return rgb, nrays.value

    def render_ao_textured(self, frame: "FrameParams", rgba: np.ndarray):
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        nrays = C.c_uint64(0)
        self.lib.orc_render_ao_textured(self.h, C.byref(frame), _ptr(rgba), rgba.shape[1], rgba.shape[0], _ptr(rgb), C.byref(nrays))
        return rgb, nrays.value

    def render_sunsky(self, frame: "FrameParams", block: "SunskyBlock"):
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        nrays = C.c_uint64(0)
        self.lib.orc_render_sunsky(self.h, C.byref(frame), C.byref(block), _ptr(rgb), C.byref(nrays))
        return rgb, nrays.value


class Oracle:
    def __init__(self):
        if not oracle_available():
            raise RuntimeError(f"{ORACLE_SO} missing: run `make -C oracle liblucille_oracle.so` or __graft_entry__.build()")
        lib = C.CDLL(ORACLE_SO)
        lib.orc_build.restype = C.c_void_p
        lib.orc_build.argtypes = [C.c_void_p, C.c_uint64]
        lib.orc_free.argtypes = [C.c_void_p]
        lib.orc_is_empty.argtypes = [C.c_void_p]
        lib.orc_num_nodes.restype = C.c_int64
        lib.orc_num_nodes.argtypes = [C.c_void_p]
        lib.orc_get_nodes.restype = C.c_int64
        lib.orc_get_nodes.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_get_triorder.argtypes = [C.c_void_p, C.c_void_p]
        lib.orc_scene_bbox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_max_depth.argtypes = [C.c_void_p]
        lib.orc_set_normals.argtypes = [C.c_void_p, C.c_void_p]
        for name in ("orc_intersect_f64", "orc_intersect_f32", "orc_occluded_f64", "orc_occluded_f32"):
            getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
            getattr(lib, name).restype = None
        lib.orc_state_build_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.orc_mt_stream.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p]
        lib.orc_mt_stream_u32.argtypes = [C.c_uint32, C.c_uint64, C.c_void_p]
        lib.orc_bucket_list.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        lib.orc_subpixel_jitter.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.orc_camera_ray.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        lib.orc_render_ao.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_beam_visibility.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.orc_render_pathtrace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_splitmix64.restype = C.c_uint64
        lib.orc_splitmix64.argtypes = [C.c_uint64]
        lib.orc_set_attributes.argtypes = [C.c_void_p] * 6
        lib.orc_state_ext_build_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.orc_texture_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.orc_render_ao_textured.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.orc_transport_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.orc_point_gather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        lib.orc_point_gather_qmc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                             C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        lib.orc_ao_point_rays_f32.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_void_p]
        lib.orc_ao_point_rays_f64.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_double, C.c_void_p]
        lib.orc_shade_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
        lib.orc_light_samples.restype = C.c_int
        lib.orc_light_samples.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_render_dirtmap.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_transport_whitted.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        lib.orc_render_whitted.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.orc_render_hitmask.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_sockdrv_encode.restype = C.c_uint64
        lib.orc_sockdrv_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
        lib.orc_hdr_encode.restype = C.c_uint64
        lib.orc_hdr_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
        lib.orc_sunsky_sky_rgb.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.orc_render_sunsky.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib = lib

    def ao_point_rays(self, points6: np.ndarray, ntheta: int, nphi: int, seed: int, eps: float = 1.0e-6, first_point: int = 0) -> np.ndarray:
        """Ray batch of the point-based AO call (calculate_occlusion's ray set-up, counter RNG + deterministic sin/cos), [n*N, 8] f32.
        ``first_point``: position of points6[0] in the whole batch (the counter RNG is keyed by it)."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros((len(pts) * ntheta * nphi, 8), dtype=np.float32)
        self.lib.orc_ao_point_rays_f32(_ptr(pts), C.c_uint64(len(pts)), C.c_uint64(first_point), ntheta, nphi, C.c_uint64(seed),
                                       C.c_double(eps), _ptr(out))
        return out

    def ao_point_rays_f64(self, points6: np.ndarray, ntheta: int, nphi: int, seed: int, eps: float = 1.0e-6, first_point: int = 0) -> np.ndarray:
        """The same batch without the rounding to fp32 records: [n*N, 6] doubles (org, dir)."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros((len(pts) * ntheta * nphi, 6), dtype=np.float64)
        self.lib.orc_ao_point_rays_f64(_ptr(pts), C.c_uint64(len(pts)), C.c_uint64(first_point), ntheta, nphi, C.c_uint64(seed),
                                       C.c_double(eps), _ptr(out))
        return out

    def sockdrv_encode(self, rgb: np.ndarray, bucket_size: int = 32) -> bytes:
        """Byte stream of the reference's socket display driver for the finished frame ``rgb`` [h,w,3] (display order)."""
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        h, w = rgb.shape[:2]
        need = self.lib.orc_sockdrv_encode(_ptr(rgb), w, h, bucket_size, None, 0)
        out = np.zeros(need, dtype=np.uint8)
        n = self.lib.orc_sockdrv_encode(_ptr(rgb), w, h, bucket_size, _ptr(out), need)
        return out[:n].tobytes()

    def texture_fetch(self, rgba: np.ndarray, uv: np.ndarray) -> np.ndarray:
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
        out = np.zeros((len(uv), 4), dtype=np.float64)
        self.lib.orc_texture_fetch(_ptr(rgba), rgba.shape[1], rgba.shape[0], _ptr(uv), C.c_uint64(len(uv)), _ptr(out))
        return out

    def hdr_encode(self, rgb: np.ndarray) -> bytes:
        """The .hdr file the reference's display driver writes for this float framebuffer ([h][w][3], display order)."""
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        h, w = rgb.shape[:2]
        n = self.lib.orc_hdr_encode(_ptr(rgb), w, h, None, 0)
        out = np.zeros(n, dtype=np.uint8)
        assert self.lib.orc_hdr_encode(_ptr(rgb), w, h, _ptr(out), n) == n
        return out.tobytes()

    def sunsky_sky_rgb(self, block: SunskyBlock, dirs: np.ndarray) -> np.ndarray:
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        out = np.zeros_like(dirs)
        self.lib.orc_sunsky_sky_rgb(C.byref(block), _ptr(dirs), C.c_uint64(len(dirs)), _ptr(out))
        return out

    def build(self, tris: np.ndarray) -> OracleTree:
        tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 9)
        return OracleTree(self.lib, self.lib.orc_build(_ptr(tris), C.c_uint64(len(tris))), len(tris))

    def mt_stream(self, n: int, seed: int = 4357) -> np.ndarray:
        out = np.zeros(n, dtype=np.float64)
        self.lib.orc_mt_stream(seed, n, _ptr(out))
        return out

    def mt_stream_u32(self, n: int, seed: int = 4357) -> np.ndarray:
        out = np.zeros(n, dtype=np.uint32)
        self.lib.orc_mt_stream_u32(seed, n, _ptr(out))
        return out

    def bucket_list(self, w: int, h: int, bucket: int = 32) -> np.ndarray:
        nb = ((w + bucket - 1) // bucket) * ((h + bucket - 1) // bucket)
        out = np.zeros((nb, 4), dtype=np.int32)
        n = self.lib.orc_bucket_list(w, h, bucket, _ptr(out), nb)
        assert n == nb
        return out

    def subpixel_jitter(self, xs, ys, xsamples, ysamples):
        a, b = C.c_double(), C.c_double()
        self.lib.orc_subpixel_jitter(xs, ys, xsamples, ysamples, C.byref(a), C.byref(b))
        return a.value, b.value

    def camera_ray(self, frame: FrameParams, x: float, y: float):
        o, d = np.zeros(3), np.zeros(3)
        self.lib.orc_camera_ray(C.byref(frame), x, y, _ptr(o), _ptr(d))
        return o, d


class ReferenceScene:
    def __init__(self, lib, handle, ntris):
        self.lib, self.h, self.ntris = lib, handle, ntris

    @property
    def empty(self):
        return bool(self.lib.lref_is_empty(self.h))

    def build_seconds(self):
        return self.lib.lref_build_seconds(self.h)

    def nodes(self) -> np.ndarray:
        ninner, nleaf, depth = C.c_int64(), C.c_int64(), C.c_int()
        self.lib.lref_tree_count(self.h, C.byref(ninner), C.byref(nleaf), C.byref(depth))
        out = np.zeros(ninner.value + nleaf.value, dtype=NODE_DTYPE)
        if len(out):
            self.lib.lref_tree_dump(self.h, _ptr(out))
        return out

    def max_depth(self):
        ninner, nleaf, depth = C.c_int64(), C.c_int64(), C.c_int()
        self.lib.lref_tree_count(self.h, C.byref(ninner), C.byref(nleaf), C.byref(depth))
        return depth.value

    def triorder(self) -> np.ndarray:
        out = np.zeros(self.ntris, dtype=np.uint32)
        if not self.empty:
            self.lib.lref_tree_triorder(self.h, _ptr(out))
        return out

    def bbox(self):
        a, b = np.zeros(3), np.zeros(3)
        self.lib.lref_scene_bbox(self.h, _ptr(a), _ptr(b))
        return a, b

    def intersect(self, rays6: np.ndarray, nthreads: int = 1, want_hits: bool = True):
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64)
        out = np.zeros(len(rays6), dtype=REFHIT_DTYPE) if want_hits else None
        sec = self.lib.lref_intersect(self.h, _ptr(rays6), C.c_uint64(len(rays6)), nthreads,
                                      _ptr(out) if want_hits else None)
        return out, sec

    def set_envmap(self, rgba: np.ndarray):
        """scene->envmap_light with this angular-map texture ([h,w,4] float32)."""
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        self.lib.lref_set_envmap(self.h, _ptr(rgba), rgba.shape[1], rgba.shape[0])

    def transport_batch(self, which: int, rays6: np.ndarray) -> np.ndarray:
        """ri_transport_ambientocclusion (0) / ri_transport_dirtmap (1) / ri_transport_whitted (2) of the compiled reference, one call
        per eye ray."""
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros((len(rays6), 3), dtype=np.float64)
        with _quiet():
            self.lib.lref_transport_batch(self.h, which, _ptr(rays6), C.c_uint64(len(rays6)), _ptr(out))
        return out

    def point_gather(self, kind: int, nsamples: int, points6: np.ndarray, col=(1.0, 1.0, 1.0), intensity=1.0) -> np.ndarray:
        """The compiled reference's occlusion() shadeop (0), ri_ibl_sample_cosweight (1, after set_envmap) or ri_domelight_sample (2),
        one call per shading point (P, N), generators reseeded to 4357 first."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        c = np.ascontiguousarray(col, dtype=np.float64)
        out = np.zeros((len(pts), 3), dtype=np.float64)
        with _quiet():
            self.lib.lref_point_gather(self.h, kind, nsamples, _ptr(pts), C.c_uint64(len(pts)), _ptr(c), C.c_double(intensity), _ptr(out))
        return out

    def shade_trace(self, pr6: np.ndarray) -> np.ndarray:
        """The compiled reference's trace() shadeop (shader.c:895-976) per (P, R) pair with a capturing shader procedure on every geom:
        records of the input block it hands the shader (called = 1) and of dst (the environment colour or zero on a miss)."""
        pr = np.ascontiguousarray(pr6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(len(pr), dtype=TRACE_REC_DTYPE)
        with _quiet():
            self.lib.lref_shade_trace(self.h, _ptr(pr), C.c_uint64(len(pr)), _ptr(out))
        return out

    def light_samples(self, nsamples: int, angle: float, points6: np.ndarray, maxm: int = 1024):
        """What an illuminance loop receives from the compiled next_lightsource() (shader.c:1116-1186) at each point (P, N):
        (L [n,maxm,3], Cl [n,maxm,3], count [n]), generators reseeded to 4357 first; needs set_envmap."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        L = np.zeros((len(pts), maxm, 3)); Cl = np.zeros((len(pts), maxm, 3)); cnt = np.zeros(len(pts), dtype=np.int32)
        with _quiet():
            self.lib.lref_light_samples(self.h, nsamples, C.c_double(angle), _ptr(pts), C.c_uint64(len(pts)), maxm, _ptr(L), _ptr(Cl), _ptr(cnt))
        return L, Cl, cnt

    def point_gather_qmc(self, kind: int, nsamples: int, points6: np.ndarray, instance=None, dim: int = 0, col=(1.0, 1.0, 1.0),
                         intensity=1.0) -> np.ndarray:
        """The same reference functions with Option "use_qmc" on (ibl.c:107-151, 266-320); instance[p] -> inray->i, dim -> inray->d."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        inst = None if instance is None else np.ascontiguousarray(instance, dtype=np.int32)
        c = np.ascontiguousarray(col, dtype=np.float64)
        out = np.zeros((len(pts), 3), dtype=np.float64)
        with _quiet():
            self.lib.lref_point_gather_ex(self.h, kind, nsamples, _ptr(pts), C.c_uint64(len(pts)), _ptr(c), C.c_double(intensity), 1,
                                          None if inst is None else _ptr(inst), dim, _ptr(out))
        return out

    def set_attributes(self, colors, st, geom_flags):
        """colors [n,3,3], st [n,3,2] (either may be None); geom_flags per geom: 1 Cs, 2 shared st, 4 unshared st, 8 two-sided."""
        c = None if colors is None else np.ascontiguousarray(colors, dtype=np.float64)
        t = None if st is None else np.ascontiguousarray(st, dtype=np.float64)
        f = np.ascontiguousarray(geom_flags, dtype=np.uint8)
        self.lib.lref_scene_set_attr(self.h, None if c is None else _ptr(c), None if t is None else _ptr(t), _ptr(f))

    def intersect_ext(self, rays6: np.ndarray) -> np.ndarray:
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(len(rays6), dtype=STATE_EXT_DTYPE)
        self.lib.lref_intersect_ext(self.h, _ptr(rays6), C.c_uint64(len(rays6)), _ptr(out))
        return out

    def beam_visibility(self, org, dirs) -> int:
        org = np.ascontiguousarray(org, dtype=np.float64)
        dirs = np.ascontiguousarray(dirs, dtype=np.float64)
        return self.lib.lref_beam_visibility(self.h, _ptr(org), _ptr(dirs))


class Reference:
    """The compiled reference.  ``stats=True`` loads the -DRI_BVH_TRACE_STATISTICS build (single-threaded counters)."""

    def __init__(self, stats: bool = False):
        path = REF_STAT_SO if stats else REF_SO
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run oracle/build_ref.sh where /root/reference exists")
        lib = C.CDLL(path)
        lib.lref_scene_build.restype = C.c_void_p
        lib.lref_scene_build.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        lib.lref_build_seconds.restype = C.c_double
        lib.lref_build_seconds.argtypes = [C.c_void_p]
        lib.lref_is_empty.argtypes = [C.c_void_p]
        lib.lref_scene_bbox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.lref_tree_count.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.lref_tree_dump.restype = C.c_int64
        lib.lref_tree_dump.argtypes = [C.c_void_p, C.c_void_p]
        lib.lref_tree_triorder.argtypes = [C.c_void_p, C.c_void_p]
        lib.lref_intersect.restype = C.c_double
        lib.lref_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        lib.lref_stats_get.argtypes = [C.c_void_p]
        lib.lref_beam_visibility.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.lref_set_envmap.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.lref_transport_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.lref_shade_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.lref_light_samples.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.lref_point_gather.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_double, C.c_void_p]
        lib.lref_point_gather_ex.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_double, C.c_int,
                                             C.c_void_p, C.c_int, C.c_void_p]
        lib.lref_scene_set_attr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.lref_intersect_ext.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.lref_sunsky_eval.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_uint64,
                                         C.c_void_p, C.c_void_p]
        lib.lref_sunsky_eval.restype = None
        lib.lref_sockdrv_stream.restype = C.c_int64
        lib.lref_sockdrv_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        lib.lref_hdr_file.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        lib.lref_texture_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        self.lib = lib
        self.stats = stats

    def sockdrv_stream(self, rgb: np.ndarray, pixels_xy: np.ndarray) -> bytes:
        """What sock_dd_open / sock_dd_write / sock_dd_close of the compiled reference put on the wire for this frame: a listener in
        this process plays the viewer.  pixels_xy = display coordinates (x | y << 16) in the order bucket_write writes them."""
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        pix = np.ascontiguousarray(pixels_xy, dtype=np.uint32)
        h, w = rgb.shape[:2]
        out = np.zeros(64 + len(pix) * 25, dtype=np.uint8)
        with _quiet():
            n = self.lib.lref_sockdrv_stream(_ptr(rgb), w, h, _ptr(pix), C.c_uint64(len(pix)), _ptr(out), C.c_uint64(len(out)))
        if n < 0:
            raise RuntimeError("socket capture failed (port 12346 busy?)")
        return out[:n].tobytes()

    def sunsky_eval(self, dirs: np.ndarray, latitude=35.39, longitude=139.44, sm=9.0, jd=20, tod=10.5, turbidity=2.0):
        """ri_sunsky_init + ri_sunsky_get_sky_rgb of the compiled reference: (rgb [n][3] float32, 21-double parameter record)."""
        dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        rgb = np.zeros_like(dirs)
        rec = np.zeros(45, dtype=np.float64)
        with _quiet():
            self.lib.lref_sunsky_eval(latitude, longitude, sm, jd, tod, turbidity, _ptr(dirs), C.c_uint64(len(dirs)), _ptr(rgb), _ptr(rec))
        return rgb, rec

    def texture_fetch(self, rgba: np.ndarray, uv: np.ndarray) -> np.ndarray:
        """ri_texture_fetch of the compiled reference on a caller-supplied float RGBA image."""
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
        out = np.zeros((len(uv), 4), dtype=np.float64)
        self.lib.lref_texture_fetch(_ptr(rgba), rgba.shape[1], rgba.shape[0], _ptr(uv), C.c_uint64(len(uv)), _ptr(out))
        return out

    def hdr_file(self, rgb: np.ndarray, path: str) -> bytes:
        """hdr_dd_open / hdr_dd_write per pixel / hdr_dd_close of the compiled reference (display/hdrdrv.c): the file's bytes."""
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        h, w = rgb.shape[:2]
        with _quiet():
            rc = self.lib.lref_hdr_file(_ptr(rgb), w, h, path.encode())
        assert rc == 0
        with open(path, "rb") as f:
            return f.read()

    def table(self, name: str, n: int) -> np.ndarray:
        """A float table the reference exports as a data symbol (sunsky.dat: S0Amplitudes, S1Amplitudes, S2Amplitudes)."""
        return np.array((C.c_float * n).in_dll(self.lib, name), dtype=np.float32)

    def build(self, tris: np.ndarray, geom_sizes=None) -> ReferenceScene:
        tris = np.ascontiguousarray(tris, dtype=np.float64).reshape(-1, 9)
        if geom_sizes is None:
            h = self.lib.lref_scene_build(_ptr(tris), C.c_uint64(len(tris)), None, 1)
        else:
            gs = np.ascontiguousarray(geom_sizes, dtype=np.uint64)
            h = self.lib.lref_scene_build(_ptr(tris), C.c_uint64(len(tris)), _ptr(gs), len(gs))
        return ReferenceScene(self.lib, h, len(tris))

    def stats_reset(self):
        self.lib.lref_stats_reset()

    def stats_get(self):
        out = np.zeros(6, dtype=np.uint64)
        self.lib.lref_stats_get(_ptr(out))
        return dict(nrays=int(out[0]), ninner=int(out[1]), nleaf=int(out[2]), ntris=int(out[3]), nhit_tris=int(out[4]))


def read_frame(path: str):
    with open(path, "rb") as f:
        assert f.read(4) == b"LFRM"
        w, h, _ = np.frombuffer(f.read(12), dtype="<u4")
        sec = np.frombuffer(f.read(8), dtype="<f8")[0]
        nrays = int(np.frombuffer(f.read(8), dtype="<u8")[0])
        rgb = np.frombuffer(f.read(), dtype="<f4").reshape(int(h), int(w), 3).copy()
    return rgb, float(sec), nrays


def read_scene(path: str):
    with open(path, "rb") as f:
        assert f.read(4) == b"LSCN"
        f.read(4)
        n = int(np.frombuffer(f.read(8), dtype="<u8")[0])
        cam = np.frombuffer(f.read(8 * 27), dtype="<f8").copy()
        tris = np.frombuffer(f.read(8 * 9 * n), dtype="<f8").reshape(n, 3, 3).copy()
        geom = np.frombuffer(f.read(4 * n), dtype="<u4").copy()
        rest = f.read(8 * 9 * n)
        normals = np.frombuffer(rest, dtype="<f8").reshape(n, 3, 3).copy() if len(rest) == 8 * 9 * n else np.zeros((n, 3, 3))
    return tris, geom, cam, normals


def run_oracle_rib(rib: str, out: str, scene: str | None = None, nthreads: int = 1, width: int = 0, height: int = 0,
                   pixelsamples: int = 0, gather: int = 0, timeout: int = 3600, sunsky: str | None = None,
                   texture: np.ndarray | None = None, attr: str | None = None):
    """Run the compiled reference renderer on a RIB in a subprocess (one frame per process)."""
    cmd = [ORACLE_RIB, rib, "--nthreads", str(nthreads), "--out", out]
    if scene:
        cmd += ["--scene", scene]
    if width and height:
        cmd += ["--width", str(width), "--height", str(height)]
    if pixelsamples:
        cmd += ["--pixelsamples", str(pixelsamples)]
    if gather:
        cmd += ["--gather", str(gather)]
    if sunsky:
        cmd += ["--sunsky", sunsky]
    if texture is not None:
        tex = np.ascontiguousarray(texture, dtype=np.float32)
        tpath = out + ".tex"
        with open(tpath, "wb") as f:
            f.write(b"LTEX")
            f.write(np.array([tex.shape[1], tex.shape[0]], dtype="<u4").tobytes())
            f.write(tex.tobytes())
        cmd += ["--texture", tpath]
    if attr:
        cmd += ["--attr", attr]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout)
    return read_frame(out)


def frame_params(cam: np.ndarray, width: int, height: int, xsamples: int | None = None, ysamples: int | None = None,
                 gather: int | None = None, bucket_size: int | None = None) -> FrameParams:
    """Build FrameParams from the 27-double camera record written by oracle_rib --scene."""
    import math
    fp = FrameParams()
    for i in range(16):
        fp.c2w[i] = cam[i]
    fp.flength = cam[16]
    fp.is_rh = int(cam[17])
    fp.width, fp.height = width, height
    fp.xsamples = int(cam[20]) if xsamples is None else xsamples
    fp.ysamples = int(cam[21]) if ysamples is None else ysamples
    g = int(cam[22]) if gather is None else gather
    n = int(math.sqrt(float(g)))
    fp.ntheta = fp.nphi = n
    fp.bucket_size = int(cam[23]) if bucket_size is None else bucket_size
    return fp
