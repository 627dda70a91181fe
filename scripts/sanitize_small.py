"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck): tiny soup, few rays, tiny frames."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from lucille_b200 import accel, scenes
tris = scenes.triangle_soup(3000, 5)
rays8 = scenes.pinhole_rays(48, 48)
rays6 = scenes.rays_f32_to_f64(rays8)
for flag in (accel.BUILD_HOST, accel.BUILD_DEVICE):
    a = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.PREC_F64 | flag)
    h32 = a.intersect(rays8); h64 = a.intersect(rays6)
    o32 = a.occluded(rays8); o64 = a.occluded(rays6)
    st = a.state(rays6, h64)
    print(flag, int((h32["prim"] != 0xFFFFFFFF).sum()), int(h64["hit"].sum()), int(o32.sum()), int(o64.sum()))
g = np.load(os.path.join("tests", "golden", "c1_scene.npz")); cam = g["cam"]
c1 = accel.Accel.bind().build(g["tris"], accel.PREC_F64 | accel.PREC_F32)
for prec in (accel.PREC_F64, accel.PREC_F32):
    for env in ("0", "1"):
        os.environ["B200_FUSED_AO_TEST"] = env
        fr = accel.make_frame(cam[:16], cam[16], bool(cam[17]), 32, 24, 2, 2, gather_nsamples=16, precision=prec)
        rgb, s = c1.render_ao(fr)
        print("frame", prec, env, float(rgb.mean()), s.nrays)
os.environ.pop("B200_FUSED_AO_TEST", None)
# dirt map / whitted / hit mask / sun-sky-free transports, .hdr encoder, point gathers, peer framebuffer, shared MT stream over 2 ranks
import oracle_lib as ol
fr = accel.make_frame(cam[:16], cam[16], bool(cam[17]), 40, 30, 2, 2, gather_nsamples=16, precision=accel.PREC_F64)
rgb, _ = c1.render_dirtmap(fr); print("dirtmap", float(rgb.mean()))
env = ol.test_texture(16, 16, 2)
rgb, _ = c1.render_whitted(fr, env); print("whitted", float(rgb.mean()))
rgb, _ = c1.render_sample(fr); print("sample", float(rgb.mean()))
print("hdr", len(accel.hdr_encode(rgb)))
pts = np.concatenate([np.random.default_rng(1).uniform(0, 1, (200, 3)), np.tile([0.0, 0.0, 1.0], (200, 1))], axis=1)
for kind in (accel.GATHER_OCCLUSION, accel.GATHER_IBL, accel.GATHER_DOME):
    out, n = a.gather_points(kind, 27, pts, env); print("gather", kind, float(out.mean()), n)
for kind in (accel.GATHER_IBL, accel.GATHER_DOME):
    out, n = a.gather_points(kind, 13, pts, env, qmc=True, qmc_instance=np.arange(200, dtype=np.int32), qmc_dim=2); print("gather qmc", kind, float(out.mean()), n)
# the wavefront forms (the scenes above are small enough for the one-lane-per-ray kernels): forced through the test hook
os.environ["B200_FUSED_AO_TEST"] = "0"
import ctypes
gs = np.load(os.path.join("tests", "golden", "sunsky.npz")); blk = ol.sunsky_block(gs["frame_block"], gs)
sky = accel.Sunsky(); ctypes.memmove(ctypes.byref(sky), ctypes.byref(blk), ctypes.sizeof(blk))
for prec in (accel.PREC_F64, accel.PREC_F32):
    f2 = accel.make_frame(cam[:16], cam[16], bool(cam[17]), 40, 30, 2, 2, gather_nsamples=16, precision=prec)
    print("wavefront sunsky", prec, float(c1.render_sunsky(f2, sky)[0].mean()), "dirtmap", float(c1.render_dirtmap(f2)[0].mean()))
print("wavefront whitted", float(c1.render_whitted(fr, env)[0].mean()), "sample", float(c1.render_sample(fr)[0].mean()))
for kind in (accel.GATHER_OCCLUSION, accel.GATHER_IBL, accel.GATHER_DOME):
    out, n = a.gather_points(kind, 27, pts, env); print("wavefront gather", kind, float(out.mean()), n)
os.environ.pop("B200_FUSED_AO_TEST", None)
print("sockdrv", len(accel.sockdrv_encode(rgb, accel.make_frame(np.eye(4).reshape(16), 1.0, False, 40, 30, 1, 1))))
ptr, _ = accel.peer_alloc(40 * 30 * 12, 0)
import copy, threading
for r in (0, 1):
    f = copy.copy(accel.make_frame(cam[:16], cam[16], bool(cam[17]), 40, 30, 2, 2, gather_nsamples=16, rng_mode=1, precision=accel.PREC_F32))
    f.rank, f.world = r, 2
    c1.render_ao_peer_dev(f, ptr)
print("peer", float(accel.peer_read(ptr, (30, 40, 3), 0).mean())); accel.peer_free(ptr, 0)
from lucille_b200 import distributed
bar, posted, res = threading.Barrier(2), {}, {}
def exch(rank):
    def fn(hits):
        posted[rank] = hits.astype(np.int64); bar.wait(timeout=300)
        per = max(len(v) for v in posted.values()); tab = np.zeros((2, per), dtype=np.int64)
        for r, v in posted.items(): tab[r, :len(v)] = v
        b, t = distributed.bucket_bases(tab, 2); bar.wait(timeout=300)
        return b[rank][:len(hits)], t
    return fn
def run(rank):
    acc = accel.Accel.bind().build(g["tris"], accel.PREC_F64); acc.set_hit_exchange(exch(rank))
    res[rank] = acc.render_ao(accel.make_frame(cam[:16], cam[16], bool(cam[17]), 70, 40, 2, 2, gather_nsamples=16, rank=rank, world=2))[0]
th = [threading.Thread(target=run, args=(r,)) for r in (0, 1)]
[t.start() for t in th]; [t.join() for t in th]
one, _ = c1.render_ao(accel.make_frame(cam[:16], cam[16], bool(cam[17]), 70, 40, 2, 2, gather_nsamples=16))
print("shared stream equal:", bool(np.array_equal(res[0] + res[1], one)))
# round 2: the double-exact hybrid occlusion kernel (fp32-representable and not; forced where the dispatcher would decline), the
# double routines it calls on undecided decisions (rays aimed at vertices and along axes), the point entries, the shade callers
os.environ["B200_HYBRID"] = "2"
for name, tt in (("exact", tris), ("inexact", tris * 1.1 + np.array([0.3, -0.2, 0.1])), ("far", tris * 3.7 + np.array([1000.0, -800.0, 400.0]))):
    h = accel.Accel.bind().build(tt, accel.PREC_F32 | accel.PREC_F64)
    lo, hi = tt.reshape(-1, 3).min(axis=0), tt.reshape(-1, 3).max(axis=0)
    rng = np.random.default_rng(2)
    org = lo + (hi - lo) * rng.uniform(-0.3, 1.3, (6000, 3))
    aim = tt.reshape(-1, 3)[rng.integers(0, 9000, 6000)]
    r = np.concatenate([org, aim - org], axis=1)
    r[::5, 3:6] = [0.0, 0.0, 1.0]
    occ = h.occluded(np.ascontiguousarray(r))
    pts6 = np.concatenate([org[:300], np.tile([0.0, 0.0, 1.0], (300, 1))], axis=1)
    hits = h.intersect(np.ascontiguousarray(r))                      # closest_hybrid_kernel (shared / own records)
    print("hybrid", name, int(occ.sum()), int(hits["hit"].sum()), int(h.occlusion_points(pts6, 4, 4, 7, f64=True).sum()), int(h.occlusion_points(pts6, 4, 4, 7).sum()))
    d = accel.Accel.bind().build(tt, accel.PREC_F64)                 # double records only: hybrid through its own records
    print("hybrid, double-only build", name, int(d.occluded(np.ascontiguousarray(r)).sum()), int(d.intersect(np.ascontiguousarray(r))["hit"].sum()))
os.environ.pop("B200_HYBRID", None)
os.environ["B200_K6"] = "1"                                           # reorder.cuh forced: stable octant sort + permuted result writes
print("k6 point entry", int(a.occlusion_points(np.tile(pts, (30, 1)), 4, 4, 3).sum()))
os.environ.pop("B200_K6")
pr = np.concatenate([pts[:, :3], np.random.default_rng(3).normal(size=(200, 3))], axis=1)
rec = a.shade_trace(pr, env); print("shade_trace", int(rec["hit"].sum()), float(rec["Ci"].sum()))
a.mt_prepare(1 << 21)
L, Cl, vis, nr = a.light_samples(48, 1.2, pts, env); print("light_samples", int(vis.sum()), nr)
L, Cl, vis, nr = a.light_samples(48, 1.2, np.tile(pts, (40, 1)), env); print("light_samples (jump table)", int(vis.sum()), nr)
print("done")
