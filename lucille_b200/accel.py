"""Host-side mirror of lucille's accelerator interface over the C ABI of ``include/lucille_b200.h``.

The reference's plugin boundary is ``ri_accel_t {build, free, intersect}`` selected by ``ri_accel_bind(accel, method)``
(src/render/accel.h:44-84, accel.c:72-109), and ``ri_raytrace()`` is the one entry every transport uses
(src/render/raytrace.c:31-69).  ``Accel`` keeps those names and meanings:

    accel = Accel.bind(RI_ACCEL_B200)         # ri_accel_bind: unknown method -> error (the reference returns -1)
    accel.build(triangles)                     # accel->build(scene): triangle soup -> opaque device structure
    hit = accel.intersect(rays)                # accel->intersect, batched: (t, u, v, prim) per ray
    occ = accel.occluded(rays)                 # the boolean the AO transport uses (ambientocclusion.c:123-129)
    accel.free()                               # accel->free

All compute happens in ``libb200accel.so`` (hand-written CUDA for sm_100a).  There is no CPU fallback: if the
library or a CUDA device is missing, calls raise ``B200Error``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

RI_ACCEL_UGRID = 0      # accel.h:20
RI_ACCEL_BVH = 1        # accel.h:21
RI_ACCEL_B200 = 2       # the id this backend registers

PREC_F32 = 1
PREC_F64 = 2
HOST_ONLY = 0x100
BUILD_DEVICE = 0x200    # build the tree on the device (same tree, bit for bit)
BUILD_HOST = 0x400      # build it with the threaded host builder (neither flag: device from 32 Ki triangles up)
MISS_PRIM = 0xFFFFFFFF
RI_INFINITY = 1.0e38

# B200_LIB: an alternative build of the same library (A/B experiments: scripts/build_variants.sh); default = the in-tree product
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200accel.so")

HIT32_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("prim", "<u4")])
HIT64_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"), ("prim", "<u4"), ("hit", "<u4")])
STATE_EXT_DTYPE = np.dtype([("E", "<f8", 3), ("I", "<f8", 3), ("color", "<f8", 3), ("st", "<f8", 2), ("t", "<f8"), ("inside", "<i4"), ("hit", "<i4")])
TRACE_REC_DTYPE = np.dtype([("Cs", "<f8", 3), ("P", "<f8", 3), ("N", "<f8", 3), ("Ng", "<f8", 3), ("dPdu", "<f8", 3), ("dPdv", "<f8", 3),
                            ("I", "<f8", 3), ("Ci", "<f8", 3), ("t", "<f8"), ("s", "<f4"), ("tt", "<f4"), ("prim", "<u4"), ("hit", "<i4")])
STATE_DTYPE = np.dtype([("P", "<f8", 3), ("Ng", "<f8", 3), ("Ns", "<f8", 3), ("tangent", "<f8", 3), ("binormal", "<f8", 3)])
NODE_DTYPE = np.dtype([("is_leaf", "<i4"), ("axis", "<i4"), ("child0", "<i8"), ("child1", "<i8"),
                       ("tri_start", "<i8"), ("ntris", "<i8"), ("lbox", "<f8", 6), ("rbox", "<f8", 6)])
NODE32_DTYPE = np.dtype([("x", "<f4", 4), ("y", "<f4", 4), ("z", "<f4", 4), ("c0", "<u4"), ("c1", "<u4"), ("axis", "<u4"), ("pad", "<u4")])
NODE64_DTYPE = np.dtype([("x", "<f8", 4), ("y", "<f8", 4), ("z", "<f8", 4), ("c0", "<u4"), ("c1", "<u4"), ("axis", "<u4"), ("pad", "<u4", 5)])
TRI32_DTYPE = np.dtype([("v0", "<f4", 3), ("prim", "<u4"), ("e1", "<f4", 3), ("pad1", "<u4"), ("e2", "<f4", 3), ("pad2", "<u4")])
TRI64_DTYPE = np.dtype([("v0", "<f8", 3), ("prim", "<u8"), ("e1", "<f8", 3), ("pad1", "<u8"), ("e2", "<f8", 3), ("pad2", "<u8")])


class B200Error(RuntimeError):
    pass


class Info(C.Structure):
    _fields_ = [("ntris", C.c_uint64), ("ninner", C.c_int64), ("nleaf", C.c_int64), ("max_depth", C.c_int32),
                ("empty", C.c_int32), ("precisions", C.c_uint32), ("device", C.c_int32),
                ("bmin", C.c_double * 3), ("bmax", C.c_double * 3), ("build_seconds", C.c_double),
                ("upload_seconds", C.c_double), ("device_bytes", C.c_uint64)]


class Counters(C.Structure):
    _fields_ = [("nrays", C.c_uint64), ("ninner", C.c_uint64), ("nleaf", C.c_uint64), ("ntris", C.c_uint64),
                ("nhit_tris", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Frame(C.Structure):
    """ri_b200_frame_t: camera + sampling description of one ambient-occlusion frame."""
    _fields_ = [("c2w", C.c_double * 16), ("flength", C.c_double), ("is_rh", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32), ("xsamples", C.c_int32), ("ysamples", C.c_int32),
                ("ntheta", C.c_int32), ("nphi", C.c_int32), ("bucket_size", C.c_int32), ("rng_mode", C.c_int32),
                ("seed", C.c_uint32), ("rank", C.c_int32), ("world", C.c_int32), ("precision", C.c_int32)]


class PathFrame(C.Structure):
    """ri_b200_path_frame_t: one path-traced frame (row P of the scope table)."""
    _fields_ = [("c2w", C.c_double * 16), ("flength", C.c_double), ("is_rh", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("spp", C.c_int32), ("max_vertices", C.c_int32), ("seed", C.c_uint32), ("kd", C.c_double), ("Le", C.c_double),
                ("rank", C.c_int32), ("world", C.c_int32), ("bucket_size", C.c_int32)]


class Sunsky(C.Structure):
    """ri_b200_sunsky_t: what the host owns after ri_sunsky_init() (sunsky.c:176-295) plus the scene's sun lights."""
    _fields_ = [("sun_theta", C.c_float), ("sun_phi", C.c_float),
                ("perez_x", C.c_float * 5), ("perez_y", C.c_float * 5), ("perez_Y", C.c_float * 5),
                ("zenith_x", C.c_float), ("zenith_y", C.c_float), ("zenith_Y", C.c_float),
                ("S0", C.c_float * 41), ("S1", C.c_float * 41), ("S2", C.c_float * 41),
                ("cie", C.c_float * 243), ("cs", C.c_float * 8),
                ("nsun", C.c_int32), ("pad", C.c_int32),
                ("sun_dir", C.c_double * 12), ("sun_col", C.c_double * 12)]


class Light(C.Structure):
    """ri_b200_light_t: the importance-sampled environment light of next_lightsource() (shader.c:1236-1310)."""
    _fields_ = [("nsamples", C.c_int32), ("seed", C.c_uint32), ("stream_offset", C.c_uint64), ("angle", C.c_double),
                ("env_rgba", C.c_void_p), ("env_width", C.c_int32), ("env_height", C.c_int32)]


class FrameStats(C.Structure):
    _fields_ = [("nrays_primary", C.c_uint64), ("nrays_ao", C.c_uint64), ("nhits_primary", C.c_uint64),
                ("ms_total", C.c_double), ("ms_primary", C.c_double), ("ms_rng", C.c_double), ("ms_ao", C.c_double),
                ("ms_resolve", C.c_double)]

    @property
    def nrays(self):
        return int(self.nrays_primary + self.nrays_ao)


# every symbol include/lucille_b200.h declares: (name, restype, argtypes)
_P, _U64, _U32, _I = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
ABI = [
    ("ri_b200_last_error", C.c_char_p, []),
    ("ri_b200_device_count", _I, []),
    ("ri_b200_build", _P, [_P, _U64, _U32, _I]),
    ("ri_b200_free", None, [_P]),
    ("ri_b200_set_normals", _I, [_P, _P]),
    ("ri_b200_info", _I, [_P, _P]),
    ("ri_b200_export_nodes", C.c_int64, [_P, _P, C.c_int64]),
    ("ri_b200_triorder", _I, [_P, _P]),
    ("ri_b200_export_flat", C.c_int64, [_P, _P, _P, _P, _P, _P]),
    ("ri_b200_export_flat_transposed", _I, [_P, _P, _P]),
    ("ri_b200_intersect1", _I, [_P, _P, _P, _P, _P]),
    ("ri_b200_intersect_batch_f32", _I, [_P, _P, _U64, _P]),
    ("ri_b200_occluded_batch_f32", _I, [_P, _P, _U64, _P]),
    ("ri_b200_intersect_batch_f64", _I, [_P, _P, _U64, _P]),
    ("ri_b200_occluded_batch_f64", _I, [_P, _P, _U64, _P]),
    ("ri_b200_state_batch_f64", _I, [_P, _P, _P, _U64, _P]),
    ("ri_b200_set_attributes", _I, [_P, _P, _P, _P, _P, _P]),
    ("ri_b200_state_ext_batch_f64", _I, [_P, _P, _P, _U64, _P]),
    ("ri_b200_set_texture", _I, [_P, _P, _I, _I, _P]),
    ("ri_b200_intersect_dev_f32", _I, [_P, _P, _U64, _P, _P]),
    ("ri_b200_occluded_dev_f32", _I, [_P, _P, _U64, _P, _P]),
    ("ri_b200_intersect_dev_f64", _I, [_P, _P, _U64, _P, _P]),
    ("ri_b200_occluded_dev_f64", _I, [_P, _P, _U64, _P, _P]),
    ("ri_b200_count_batch", _I, [_P, _P, _U64, _U32, _I, _P]),
    ("ri_b200_launch_count", _U64, []),
    ("ri_b200_host_alloc", _P, [_U64]),
    ("ri_b200_host_free", None, [_P]),
    ("ri_b200_render_ao", _I, [_P, _P, _P, _P]),
    ("ri_b200_render_ao_dev", _I, [_P, _P, _P, _P, _P]),
    ("ri_b200_render_ao_tiles_dev", _I, [_P, _P, _P, _P, _P]),
    ("ri_b200_frame_pixels", C.c_int64, [_P, _P, C.c_int64]),
    ("ri_b200_render_ao_peer_dev", _I, [_P, _P, _P, _P, _P]),
    ("ri_b200_set_hit_exchange", _I, [_P, _P, _P]),
    ("ri_b200_gather_points_f64", _I, [_P, _P, _P, _U64, _P, _P]),
    ("ri_b200_shade_trace_f64", _I, [_P, _P, _U64, _P, _I, _I, _P]),
    ("ri_b200_light_samples_count", _I, [_I]),
    ("ri_b200_light_samples_f64", _I, [_P, _P, _P, _U64, _P, _P, _P, _P]),
    ("ri_b200_occlusion_points_f32", _I, [_P, _P, _P, _U64, _P]),
    ("ri_b200_occlusion_points_dev_f32", _I, [_P, _P, _P, _U64, _P, _P]),
    ("ri_b200_ao_point_rays_f32", _I, [_P, _P, _P, _U64, _P]),
    ("ri_b200_occlusion_points_f64", _I, [_P, _P, _P, _U64, _P]),
    ("ri_b200_occlusion_points_dev_f64", _I, [_P, _P, _P, _U64, _P, _P]),
    ("ri_b200_peer_alloc", _P, [_U64, _I, _P]),
    ("ri_b200_peer_open", _P, [_P, _I]),
    ("ri_b200_peer_close", _I, [_P, _I]),
    ("ri_b200_peer_free", _I, [_P, _I]),
    ("ri_b200_peer_read", _I, [_P, _P, _U64, _I]),
    ("ri_b200_render_sunsky", _I, [_P, _P, _P, _P, _P]),
    ("ri_b200_render_sunsky_tiles_dev", _I, [_P, _P, _P, _P, _P, _P]),
    ("ri_b200_sunsky_rgb", _I, [_P, _P, _U64, _P, _I]),
    ("ri_b200_render_dirtmap", _I, [_P, _P, _P, _P]),
    ("ri_b200_render_whitted", _I, [_P, _P, _P, _I, _I, _P, _P]),
    ("ri_b200_render_sample", _I, [_P, _P, _P, _P]),
    ("ri_b200_hdr_encode", C.c_int64, [_P, _I, _I, _P, _U64, _I, _I]),
    ("ri_b200_sockdrv_encode", C.c_int64, [_P, _P, _P, _U64, _I, _I]),
    ("ri_b200_beam_visibility_batch", _I, [_P, _P, _U64, _P]),
    ("ri_b200_render_pathtrace", _I, [_P, _P, _P, _P]),
    ("ri_b200_render_pathtrace_tiles_dev", _I, [_P, _P, _P, _P, _P]),
    ("ri_b200_mt_stream", _I, [_P, _U32, _U64, _P, _I]),
    ("ri_b200_mt_prepare", _I, [_P, _U32, _U64]),
]

_lib = None


def load_library():
    """dlopen the in-tree ``libb200accel.so``; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a). lucille_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, restype, argtypes in ABI:
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def last_error() -> str:
    return load_library().ri_b200_last_error().decode()


def _check(rc: int):
    if rc < 0:
        raise B200Error(last_error())
    return rc


def device_count() -> int:
    return load_library().ri_b200_device_count()


def launch_count() -> int:
    return int(load_library().ri_b200_launch_count())


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.data_ptr())        # torch tensor


def mt_stream(n: int, seed: int = 4357, device: int = 0, accel: "Accel" = None) -> np.ndarray:
    """First ``n`` 32-bit outputs of the reference's randomMT2() stream, generated on the device (sequentially by one
    CTA, or -- with ``accel`` -- by the frame path: jump-ahead states + one CTA per segment)."""
    out = np.zeros(n, dtype=np.uint32)
    _check(load_library().ri_b200_mt_stream(accel.data if accel is not None else None, seed, n, _ptr(out), device))
    return out


class Accel:
    """``ri_accel_t`` for the B200 backend."""

    def __init__(self, method: int = RI_ACCEL_B200):
        if method != RI_ACCEL_B200:
            # ri_accel_bind returns -1 for unknown methods (accel.c:102-106); UGRID/BVH live in the CPU reference
            raise B200Error(f"ri_accel_bind: method {method} is not provided by lucille_b200 (only RI_ACCEL_B200={RI_ACCEL_B200})")
        self.lib = load_library()
        self.data = None           # ri_accel_t.data
        self.ntris = 0

    @classmethod
    def bind(cls, method: int = RI_ACCEL_B200) -> "Accel":
        return cls(method)

    # -- accel->build ----------------------------------------------------------------------------
    def build(self, triangles: np.ndarray, precisions: int = PREC_F32 | PREC_F64, device: int = 0) -> "Accel":
        """triangles: [ntris,3,3] (or [ntris,9]) float64 in the order the scene's geoms are flattened."""
        tris = np.ascontiguousarray(triangles, dtype=np.float64).reshape(-1, 9)
        if self.data is not None:
            self.free()
        h = self.lib.ri_b200_build(_ptr(tris), len(tris), precisions, device)
        if not h:
            raise B200Error(last_error())
        self.data = h
        self.ntris = len(tris)
        return self

    def set_normals(self, tri_normals) -> "Accel":
        """Per-corner vertex normals [ntris,3,3] in input order (ri_geom_t.normals); None removes them."""
        if tri_normals is None:
            _check(self.lib.ri_b200_set_normals(self._h(), None))
        else:
            n = np.ascontiguousarray(tri_normals, dtype=np.float64).reshape(-1, 9)
            assert len(n) == self.ntris
            _check(self.lib.ri_b200_set_normals(self._h(), _ptr(n)))
        return self

    # -- accel->free -----------------------------------------------------------------------------
    def free(self):
        if self.data is not None:
            self.lib.ri_b200_free(self.data)
            self.data = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def _h(self):
        if self.data is None:
            raise B200Error("accelerator has not been built")
        return self.data

    # -- introspection ---------------------------------------------------------------------------
    def info(self) -> Info:
        out = Info()
        _check(self.lib.ri_b200_info(self._h(), C.byref(out)))
        return out

    def nodes(self) -> np.ndarray:
        n = _check(self.lib.ri_b200_export_nodes(self._h(), None, 0))
        out = np.zeros(n, dtype=NODE_DTYPE)
        if n:
            _check(self.lib.ri_b200_export_nodes(self._h(), _ptr(out), n))
        return out

    def triorder(self) -> np.ndarray:
        out = np.zeros(self.ntris, dtype=np.uint32)
        if self.ntris:
            _check(self.lib.ri_b200_triorder(self._h(), _ptr(out)))
        return out

    def flat(self):
        """Flat device records of a HOST_ONLY accelerator: dict(nodes32, nodes64, tris32, tris64, root_word, top_count)."""
        hdr = np.zeros(4, dtype=np.uint32)
        ninner = _check(self.lib.ri_b200_export_flat(self._h(), None, None, None, None, _ptr(hdr)))
        info = self.info()
        n32 = np.zeros(ninner if info.precisions & PREC_F32 else 0, dtype=NODE32_DTYPE)
        n64 = np.zeros(ninner if info.precisions & PREC_F64 else 0, dtype=NODE64_DTYPE)
        nslots = int(hdr[3])
        t32 = np.zeros(nslots if info.precisions & PREC_F32 else 0, dtype=TRI32_DTYPE)
        t64 = np.zeros(nslots if info.precisions & PREC_F64 else 0, dtype=TRI64_DTYPE)
        _check(self.lib.ri_b200_export_flat(self._h(), _ptr(n32) if len(n32) else None, _ptr(n64) if len(n64) else None,
                                            _ptr(t32) if len(t32) else None, _ptr(t64) if len(t64) else None, _ptr(hdr)))
        t32t = np.zeros(len(t32) * 48, dtype=np.uint8)       # leaf-transposed copies, raw bytes (layout: csrc/bvh_build.h)
        t64t = np.zeros(len(t64) * 96, dtype=np.uint8)
        _check(self.lib.ri_b200_export_flat_transposed(self._h(), _ptr(t32t) if len(t32t) else None, _ptr(t64t) if len(t64t) else None))
        return dict(nodes32=n32, nodes64=n64, tris32=t32, tris64=t64, tris32t=t32t, tris64t=t64t, root_word=int(hdr[0]),
                    ninner=int(hdr[1]), top_count=int(hdr[2]), nslots=nslots)

    # -- accel->intersect, batched (host buffers) ---------------------------------------------------
    def intersect(self, rays: np.ndarray) -> np.ndarray:
        """Closest hit per ray.  float32 [n,8] rays -> HIT32 records; float64 [n,6] rays -> HIT64 records."""
        rays = np.ascontiguousarray(rays)
        if rays.dtype == np.float32:
            assert rays.ndim == 2 and rays.shape[1] == 8, "fp32 rays are [n,8]: ox,oy,oz,tmin,dx,dy,dz,tmax"
            out = np.zeros(len(rays), dtype=HIT32_DTYPE)
            _check(self.lib.ri_b200_intersect_batch_f32(self._h(), _ptr(rays), len(rays), _ptr(out)))
        elif rays.dtype == np.float64:
            assert rays.ndim == 2 and rays.shape[1] == 6, "fp64 rays are [n,6]: org.xyz, dir.xyz"
            out = np.zeros(len(rays), dtype=HIT64_DTYPE)
            _check(self.lib.ri_b200_intersect_batch_f64(self._h(), _ptr(rays), len(rays), _ptr(out)))
        else:
            raise B200Error(f"unsupported ray dtype {rays.dtype}")
        return out

    def occluded(self, rays: np.ndarray) -> np.ndarray:
        """1 where ri_raytrace() would report a hit (what calculate_occlusion counts), else 0."""
        rays = np.ascontiguousarray(rays)
        out = np.zeros(len(rays), dtype=np.uint8)
        if rays.dtype == np.float32:
            assert rays.ndim == 2 and rays.shape[1] == 8
            _check(self.lib.ri_b200_occluded_batch_f32(self._h(), _ptr(rays), len(rays), _ptr(out)))
        elif rays.dtype == np.float64:
            assert rays.ndim == 2 and rays.shape[1] == 6
            _check(self.lib.ri_b200_occluded_batch_f64(self._h(), _ptr(rays), len(rays), _ptr(out)))
        else:
            raise B200Error(f"unsupported ray dtype {rays.dtype}")
        return out

    def raytrace(self, org, dir):
        """``ri_raytrace()`` for one ray (raytrace.c:31-69): returns (hit, HIT64 record, STATE record)."""
        o = np.ascontiguousarray(org, dtype=np.float64)
        d = np.ascontiguousarray(dir, dtype=np.float64)
        hit = np.zeros(1, dtype=HIT64_DTYPE)
        state = np.zeros(1, dtype=STATE_DTYPE)
        rc = _check(self.lib.ri_b200_intersect1(self._h(), _ptr(o), _ptr(d), _ptr(hit), _ptr(state)))
        return bool(rc), hit[0], state[0]

    def state(self, rays6: np.ndarray, hits: np.ndarray) -> np.ndarray:
        """``ri_intersection_state_build`` for a batch (intersection_state.c:99-248)."""
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64)
        hits = np.ascontiguousarray(hits)
        assert hits.dtype == HIT64_DTYPE
        out = np.zeros(len(rays6), dtype=STATE_DTYPE)
        _check(self.lib.ri_b200_state_batch_f64(self._h(), _ptr(rays6), _ptr(hits), len(rays6), _ptr(out)))
        return out

    def set_attributes(self, colors=None, has_color=None, st=None, has_st=None, inside=None) -> "Accel":
        """Per-corner vertex colours [n,3,3] / texture coordinates [n,3,2] with per-triangle presence flags, and the back-side flag
        of two-sided geometry (intersection_state.c:192-246)."""
        def arr(x, dt, shape):
            return None if x is None else np.ascontiguousarray(x, dtype=dt).reshape(shape)
        args = [arr(colors, np.float64, (-1, 9)), arr(has_color, np.uint8, -1), arr(st, np.float64, (-1, 6)), arr(has_st, np.uint8, -1),
                arr(inside, np.uint8, -1)]
        _check(self.lib.ri_b200_set_attributes(self._h(), *[_ptr(x) for x in args]))
        return self

    def set_texture(self, rgba, tri_textured=None) -> "Accel":
        """Material texture of the AO transport ([h,w,4] float32; None removes it); call after set_attributes."""
        if rgba is None:
            _check(self.lib.ri_b200_set_texture(self._h(), None, 0, 0, None))
            return self
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        mask = None if tri_textured is None else np.ascontiguousarray(tri_textured, dtype=np.uint8)
        _check(self.lib.ri_b200_set_texture(self._h(), _ptr(rgba), rgba.shape[1], rgba.shape[0], _ptr(mask)))
        return self

    def state_ext(self, rays6: np.ndarray, hits: np.ndarray) -> np.ndarray:
        """E, I, colour, st, t, inside of ``ri_intersection_state_build`` for a batch of fp64 hits."""
        rays6 = np.ascontiguousarray(rays6, dtype=np.float64).reshape(-1, 6)
        hits = np.ascontiguousarray(hits)
        out = np.zeros(len(rays6), dtype=STATE_EXT_DTYPE)
        _check(self.lib.ri_b200_state_ext_batch_f64(self._h(), _ptr(rays6), _ptr(hits), len(rays6), _ptr(out)))
        return out

    def count(self, rays: np.ndarray, anyhit: bool = False) -> dict:
        """The reference's RI_BVH_TRACE_STATISTICS counters for this batch (bvh.c:682-706)."""
        rays = np.ascontiguousarray(rays)
        prec = PREC_F32 if rays.dtype == np.float32 else PREC_F64
        out = Counters()
        _check(self.lib.ri_b200_count_batch(self._h(), _ptr(rays), len(rays), prec, int(anyhit), C.byref(out)))
        return out.as_dict()

    # -- device-resident batches (torch tensors or raw device pointers), asynchronous on `stream` ------
    def intersect_dev(self, d_rays, n: int, d_out, stream: Optional[int] = None, f64: bool = False):
        fn = self.lib.ri_b200_intersect_dev_f64 if f64 else self.lib.ri_b200_intersect_dev_f32
        _check(fn(self._h(), _ptr(d_rays), n, _ptr(d_out), C.c_void_p(stream) if stream else None))

    def occluded_dev(self, d_rays, n: int, d_out, stream: Optional[int] = None, f64: bool = False):
        fn = self.lib.ri_b200_occluded_dev_f64 if f64 else self.lib.ri_b200_occluded_dev_f32
        _check(fn(self._h(), _ptr(d_rays), n, _ptr(d_out), C.c_void_p(stream) if stream else None))

    # -- calculate_occlusion as a batch: shading points in, occluded-ray counts out (rays generated on the device) ------
    def occlusion_points(self, points6: np.ndarray, ntheta: int, nphi: int, seed: int, eps: float = 1.0e-6, f64: bool = False) -> np.ndarray:
        """calculate_occlusion for a batch of shading points (P, Ns): occluded-ray counts.  f64: double rays against the double records
        (the double reference's answer for every ray) instead of fp32 ray records."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(len(pts), dtype=np.uint32)
        par = AoPoints(ntheta, nphi, seed, eps)
        fn = self.lib.ri_b200_occlusion_points_f64 if f64 else self.lib.ri_b200_occlusion_points_f32
        _check(fn(self._h(), C.byref(par), _ptr(pts), len(pts), _ptr(out)))
        return out

    def occlusion_points_dev(self, d_points, n: int, d_out, ntheta: int, nphi: int, seed: int, eps: float = 1.0e-6,
                             stream: Optional[int] = None):
        par = AoPoints(ntheta, nphi, seed, eps)
        _check(self.lib.ri_b200_occlusion_points_dev_f32(self._h(), C.byref(par), _ptr(d_points), n, _ptr(d_out),
                                                          C.c_void_p(stream) if stream else None))

    def ao_point_rays(self, points6: np.ndarray, ntheta: int, nphi: int, seed: int, eps: float = 1.0e-6) -> np.ndarray:
        """The ray batch ri_b200_occlusion_points_f32 traces, [n*ntheta*nphi, 8] float32."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros((len(pts) * ntheta * nphi, 8), dtype=np.float32)
        par = AoPoints(ntheta, nphi, seed, eps)
        _check(self.lib.ri_b200_ao_point_rays_f32(self._h(), C.byref(par), _ptr(pts), len(pts), _ptr(out)))
        return out

    # -- the frame-level transport ------------------------------------------------------------------
    def render_ao(self, frame: Frame):
        """One ambient-occlusion frame -> (rgb [h,w,3] float32 on the host, FrameStats)."""
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        stats = FrameStats()
        _check(self.lib.ri_b200_render_ao(self._h(), C.byref(frame), _ptr(rgb), C.byref(stats)))
        return rgb, stats

    def render_whitted(self, frame: Frame, env=None):
        """One frame with the Whitted refraction tracer (transport/whitted.c) and the angular-map environment ``env`` ([h,w,4] float32
        or None) -> (rgb [h,w,3] float32 on the host, FrameStats)."""
        env = None if env is None else np.ascontiguousarray(env, dtype=np.float32)
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        stats = FrameStats()
        _check(self.lib.ri_b200_render_whitted(self._h(), C.byref(frame), _ptr(env), 0 if env is None else env.shape[1],
                                               0 if env is None else env.shape[0], _ptr(rgb), C.byref(stats)))
        return rgb, stats

    def render_sample(self, frame: Frame):
        """One frame with ``ri_transport_sample`` (transport/transport.c): white where the eye ray hits."""
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        stats = FrameStats()
        _check(self.lib.ri_b200_render_sample(self._h(), C.byref(frame), _ptr(rgb), C.byref(stats)))
        return rgb, stats

    def render_dirtmap(self, frame: Frame):
        """One frame with the dirt-map transport (transport/dirtmap.c) -> (rgb [h,w,3] float32 on the host, FrameStats)."""
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        stats = FrameStats()
        _check(self.lib.ri_b200_render_dirtmap(self._h(), C.byref(frame), _ptr(rgb), C.byref(stats)))
        return rgb, stats

    def render_sunsky(self, frame: Frame, sky: "Sunsky"):
        """One frame with the sun-sky gather (ambientocclusion.c:206-324) -> (rgb [h,w,3] float32 on the host, FrameStats)."""
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        stats = FrameStats()
        _check(self.lib.ri_b200_render_sunsky(self._h(), C.byref(frame), C.byref(sky), _ptr(rgb), C.byref(stats)))
        return rgb, stats

    def beam_visibility(self, beams15: np.ndarray) -> np.ndarray:
        """``ri_bvh_intersect_beam_visibility`` for a batch: [n,15] float64 beams -> int32 codes (0 miss, 1 full hit, 2 partial, -1 invalid)."""
        beams15 = np.ascontiguousarray(beams15, dtype=np.float64).reshape(-1, 15)
        out = np.zeros(len(beams15), dtype=np.int32)
        _check(self.lib.ri_b200_beam_visibility_batch(self._h(), _ptr(beams15), len(beams15), _ptr(out)))
        return out

    def render_pathtrace(self, frame: "PathFrame"):
        """One path-traced frame -> (rgb [h,w,3] float32 on the host, FrameStats with the ray count)."""
        rgb = np.zeros((frame.height, frame.width, 3), dtype=np.float32)
        stats = FrameStats()
        _check(self.lib.ri_b200_render_pathtrace(self._h(), C.byref(frame), _ptr(rgb), C.byref(stats)))
        return rgb, stats

    def render_ao_tiles_dev(self, frame: Frame, d_packed, stream: Optional[int] = None, want_stats: bool = True):
        """This rank's pixels only, packed [npixels,3] in visiting order, left in device memory."""
        stats = FrameStats()
        _check(self.lib.ri_b200_render_ao_tiles_dev(self._h(), C.byref(frame), _ptr(d_packed),
                                                    C.c_void_p(stream) if stream else None, C.byref(stats) if want_stats else None))
        return stats

    def gather_points(self, kind: int, nsamples: int, points6, env=None, col=(1.0, 1.0, 1.0), intensity: float = 1.0,
                      seed: int = 4357, stream_offset: int = 0, qmc: bool = False, qmc_instance=None, qmc_dim: int = 0):
        """Hemisphere gathers at shading points (P, N) (ri_b200_gather_points_f64): GATHER_OCCLUSION = the occlusion() shadeop,
        GATHER_IBL = ri_ibl_sample_cosweight with the angular map ``env`` [h,w,4], GATHER_DOME = ri_domelight_sample.
        Returns ([n,3] float64, rays traced)."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        g = Gather()
        g.kind, g.nsamples, g.seed, g.stream_offset = int(kind), int(nsamples), int(seed), int(stream_offset)
        if env is not None:
            env = np.ascontiguousarray(env, dtype=np.float32)
            g.env_rgba, g.env_width, g.env_height = env.ctypes.data, env.shape[1], env.shape[0]
        for i in range(3):
            g.col[i] = float(col[i])
        g.intensity = float(intensity)
        g.use_qmc, g.qmc_dim = int(bool(qmc)), int(qmc_dim)       # Option "use_qmc": quasi-Monte Carlo branches, inray->i / inray->d
        if qmc_instance is not None:
            inst = np.ascontiguousarray(qmc_instance, dtype=np.int32)
            assert len(inst) == len(pts)
            g.qmc_instance = inst.ctypes.data
        out = np.zeros((len(pts), 3), dtype=np.float64)
        nrays = C.c_uint64(0)
        _check(self.lib.ri_b200_gather_points_f64(self._h(), C.byref(g), _ptr(pts), len(pts), _ptr(out), C.byref(nrays)))
        return out, nrays.value

    def mt_prepare(self, max_words: int, seed: int = 4357) -> "Accel":
        """Build the jump-ahead state table of the MT19937 stream ahead of the first call that needs it (ri_b200_mt_prepare)."""
        _check(self.lib.ri_b200_mt_prepare(self._h(), int(seed), int(max_words)))
        return self

    def shade_trace(self, pr6, env=None) -> np.ndarray:
        """The trace() shadeop (shader.c:895-976) for (P, R) pairs up to the call of the hit surface's shader procedure
        (ri_b200_shade_trace_f64): records of TRACE_REC_DTYPE -- the shader's input block on a hit, the environment colour on a miss."""
        pr = np.ascontiguousarray(pr6, dtype=np.float64).reshape(-1, 6)
        out = np.zeros(len(pr), dtype=TRACE_REC_DTYPE)
        if env is not None:
            env = np.ascontiguousarray(env, dtype=np.float32)
        _check(self.lib.ri_b200_shade_trace_f64(self._h(), _ptr(pr), len(pr), None if env is None else _ptr(env),
                                                0 if env is None else env.shape[1], 0 if env is None else env.shape[0], _ptr(out)))
        return out

    def light_samples(self, nsamples: int, angle: float, points6, env=None, seed: int = 4357, stream_offset: int = 0):
        """The light samples of next_lightsource() (shader.c:1116-1310) at shading points (P, N) (ri_b200_light_samples_f64):
        (L [n,m,3], Cl [n,m,3], visible [n,m] u8, shadow rays traced); visible marks the samples the reference's loop returns."""
        pts = np.ascontiguousarray(points6, dtype=np.float64).reshape(-1, 6)
        g = Light()
        g.nsamples, g.seed, g.stream_offset, g.angle = int(nsamples), int(seed), int(stream_offset), float(angle)
        if env is not None:
            env = np.ascontiguousarray(env, dtype=np.float32)
            g.env_rgba, g.env_width, g.env_height = env.ctypes.data, env.shape[1], env.shape[0]
        m = int(self.lib.ri_b200_light_samples_count(int(nsamples)))
        L = np.zeros((len(pts), m, 3)); Cl = np.zeros((len(pts), m, 3)); vis = np.zeros((len(pts), m), dtype=np.uint8)
        nrays = C.c_uint64(0)
        got = self.lib.ri_b200_light_samples_f64(self._h(), C.byref(g), _ptr(pts), len(pts), _ptr(L), _ptr(Cl), _ptr(vis), C.byref(nrays))
        if got < 0:
            _check(got)
        assert got == m
        return L, Cl, vis, nrays.value

    def set_hit_exchange(self, fn):
        """rng_mode 0 on world > 1 (ri_b200_set_hit_exchange): fn(bucket_hits: np.ndarray[u32]) -> (bucket_base: array of u64, same
        length; frame_hits: int).  fn(None) means "this rank failed before it had counts": fn still takes part in the exchange so the
        other ranks are released, and the frame fails everywhere.  None removes the callback."""
        if fn is None:
            self._hit_cb = None
            _check(self.lib.ri_b200_set_hit_exchange(self._h(), None, None))
            return

        def tramp(_user, hits_p, n, base_p, total_p):
            try:
                if not hits_p and not base_p:     # the library's guard: this rank failed before the exchange
                    fn(None)
                    return 1
                hits = np.ctypeslib.as_array(hits_p, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
                base, total = fn(hits)
                base = np.asarray(base, dtype=np.uint64)
                if len(base) != n:
                    return 2
                for i in range(n):
                    base_p[i] = int(base[i])
                total_p[0] = int(total)
                return 0
            except Exception:                     # an exception must not unwind through the C frame
                import traceback
                traceback.print_exc()
                return 1

        self._hit_cb = HIT_EXCHANGE_FN(tramp)     # keep the trampoline alive as long as the accelerator uses it
        _check(self.lib.ri_b200_set_hit_exchange(self._h(), C.cast(self._hit_cb, C.c_void_p), None))

    def render_ao_peer_dev(self, frame: Frame, d_rgb_shared: int, stream: Optional[int] = None, want_stats: bool = True):
        """This rank's buckets stored at their framebuffer positions in a buffer shared by the ranks of the node (peer memory)."""
        stats = FrameStats()
        _check(self.lib.ri_b200_render_ao_peer_dev(self._h(), C.byref(frame), C.c_void_p(d_rgb_shared),
                                                   C.c_void_p(stream) if stream else None, C.byref(stats) if want_stats else None))
        return stats

    def render_ao_dev(self, frame: Frame, d_rgb, stream: Optional[int] = None, want_stats: bool = True):
        stats = FrameStats()
        _check(self.lib.ri_b200_render_ao_dev(self._h(), C.byref(frame), _ptr(d_rgb),
                                              C.c_void_p(stream) if stream else None, C.byref(stats) if want_stats else None))
        return stats


def sunsky_rgb(sky: "Sunsky", dirs: np.ndarray, device: int = 0) -> np.ndarray:
    """``ri_sunsky_get_sky_rgb`` (sunsky.c:387-408) for a batch of directions, computed on the device."""
    dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
    out = np.zeros_like(dirs)
    _check(load_library().ri_b200_sunsky_rgb(C.byref(sky), _ptr(dirs), len(dirs), _ptr(out), device))
    return out


def hdr_encode(rgb, width: int = 0, height: int = 0, device: int = 0) -> bytes:
    """The Radiance .hdr file lucille's display driver writes for a float framebuffer (hdrdrv.c / rgbe.c), encoded on the device.
    ``rgb``: a [h,w,3] float32 numpy array on the host, or a device pointer / CUDA tensor with ``width`` and ``height`` given."""
    lib = load_library()
    if isinstance(rgb, np.ndarray):
        rgb = np.ascontiguousarray(rgb, dtype=np.float32)
        height, width = rgb.shape[:2]
        src, on_dev = _ptr(rgb), 0
    else:
        src, on_dev = _ptr(rgb), 1
    cap = 64 + 4 * height + width * height * 4 + (width // 64 + 8) * 4 * height + 128
    out = np.zeros(cap, dtype=np.uint8)
    n = lib.ri_b200_hdr_encode(src, width, height, _ptr(out), cap, device, on_dev)
    if n < 0 or n > cap:
        raise B200Error(last_error() if n < 0 else "hdr buffer too small")
    return out[:n].tobytes()


GATHER_OCCLUSION, GATHER_IBL, GATHER_DOME = 0, 1, 2


class Gather(C.Structure):
    """ri_b200_gather_t"""
    _fields_ = [("kind", C.c_int32), ("nsamples", C.c_int32), ("seed", C.c_uint32), ("pad_", C.c_uint32), ("stream_offset", C.c_uint64),
                ("env_rgba", C.c_void_p), ("env_width", C.c_int32), ("env_height", C.c_int32), ("col", C.c_double * 3),
                ("intensity", C.c_double), ("use_qmc", C.c_int32), ("qmc_dim", C.c_int32), ("qmc_instance", C.c_void_p)]


class AoPoints(C.Structure):
    """ri_b200_ao_points_t"""
    _fields_ = [("ntheta", C.c_int32), ("nphi", C.c_int32), ("seed", C.c_uint64), ("eps", C.c_double)]


HIT_EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))


def sockdrv_encode(rgb, frame: Frame, device: int = 0) -> bytes:
    """Byte stream of lucille's socket display driver for the finished frame ``rgb`` [h,w,3] (ri_b200_sockdrv_encode)."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    lib = load_library()
    need = lib.ri_b200_sockdrv_encode(_ptr(rgb), C.byref(frame), None, 0, device, 0)
    if need < 0:
        raise B200Error(last_error())
    out = np.zeros(need, dtype=np.uint8)
    n = lib.ri_b200_sockdrv_encode(_ptr(rgb), C.byref(frame), _ptr(out), need, device, 0)
    if n < 0:
        raise B200Error(last_error())
    return out[:n].tobytes()


def peer_alloc(nbytes: int, device: int = 0):
    """(device pointer, 64-byte IPC handle) of a zeroed buffer other processes of the node can map (ri_b200_peer_alloc)."""
    handle = np.zeros(64, dtype=np.uint8)
    p = load_library().ri_b200_peer_alloc(nbytes, device, _ptr(handle))
    if not p:
        raise B200Error(last_error())
    return int(p), handle.tobytes()


def peer_open(handle: bytes, device: int = 0) -> int:
    h = np.frombuffer(handle, dtype=np.uint8).copy()
    p = load_library().ri_b200_peer_open(_ptr(h), device)
    if not p:
        raise B200Error(last_error())
    return int(p)


def peer_close(p: int, device: int = 0):
    _check(load_library().ri_b200_peer_close(C.c_void_p(p), device))


def peer_free(p: int, device: int = 0):
    _check(load_library().ri_b200_peer_free(C.c_void_p(p), device))


def peer_read(p: int, shape, device: int = 0) -> np.ndarray:
    out = np.zeros(shape, dtype=np.float32)
    _check(load_library().ri_b200_peer_read(C.c_void_p(p), _ptr(out), out.nbytes, device))
    return out


def frame_pixels(frame: Frame) -> np.ndarray:
    """Pixels (x | y << 16) rendered by ``frame.rank`` of ``frame.world`` ranks, in visiting order (host logic only)."""
    lib = load_library()
    n = _check(lib.ri_b200_frame_pixels(C.byref(frame), None, 0))
    out = np.zeros(n, dtype=np.uint32)
    if n:
        _check(lib.ri_b200_frame_pixels(C.byref(frame), _ptr(out), n))
    return out


def make_path_frame(c2w, flength: float, is_rh: bool, width: int, height: int, spp: int = 4, max_vertices: int = 10,
                    seed: int = 1, kd: float = 1.0, Le: float = 1.0, rank: int = 0, world: int = 1, bucket_size: int = 32) -> PathFrame:
    f = PathFrame()
    c = np.asarray(c2w, dtype=np.float64).reshape(16)
    for i in range(16):
        f.c2w[i] = float(c[i])
    f.flength, f.is_rh, f.width, f.height = float(flength), int(bool(is_rh)), int(width), int(height)
    f.spp, f.max_vertices, f.seed, f.kd, f.Le = int(spp), int(max_vertices), int(seed), float(kd), float(Le)
    f.rank, f.world, f.bucket_size = int(rank), int(world), int(bucket_size)
    return f


def make_frame(c2w, flength: float, is_rh: bool, width: int, height: int, xsamples: int, ysamples: int,
               gather_nsamples: int = 64, bucket_size: int = 32, rng_mode: int = 0, seed: int = 4357,
               rank: int = 0, world: int = 1, precision: int = PREC_F64) -> Frame:
    """Frame from camera parameters; ``ntheta = nphi = (int)sqrt(gather_nsamples)`` as ambientocclusion.c:378-387."""
    import math
    f = Frame()
    c = np.asarray(c2w, dtype=np.float64).reshape(16)
    for i in range(16):
        f.c2w[i] = float(c[i])
    f.flength = float(flength)
    f.is_rh = int(bool(is_rh))
    f.width, f.height = int(width), int(height)
    f.xsamples, f.ysamples = int(xsamples), int(ysamples)
    n = int(math.sqrt(float(gather_nsamples)))
    f.ntheta = f.nphi = n
    f.bucket_size = int(bucket_size)
    f.rng_mode = int(rng_mode)
    f.seed = int(seed)
    f.rank, f.world = int(rank), int(world)
    f.precision = int(precision)
    return f
