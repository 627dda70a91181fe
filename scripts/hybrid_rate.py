"""Double-exact occlusion on the C3 scene: the hybrid kernel (csrc/hybrid.cuh: fp32 records with certified decisions, double records
where fp32 cannot decide) against the plain double kernel and the fp32 kernel, on the same AO batch.
    python scripts/hybrid_rate.py [npoints] [variant ...]      variants: exact (the soup as it is), inexact (scaled / shifted in double), far
Prints rates and checks that hybrid == plain double for every ray."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucille_b200 import accel, scenes
import bench

npoints = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
variants = sys.argv[2:] or ["exact", "inexact", "far"]
st = torch.cuda.Stream(); torch.cuda.set_stream(st)


def ev(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


base = scenes.triangle_soup(bench.NTRIS, scenes.SEED_C3)
for v in variants:
    if v == "exact":
        tris, scale, shift = base, 1.0, np.zeros(3)
    elif v == "inexact":
        scale, shift = 1.1, np.array([0.3, -0.2, 0.1]); tris = base * scale + shift
    else:
        scale, shift = 3.7, np.array([1000.0, -800.0, 400.0]); tris = base * scale + shift
    hyb = accel.Accel.bind().build(tris, accel.PREC_F32 | accel.PREC_F64)
    plain = accel.Accel.bind().build(tris, accel.PREC_F64)
    bench.NPOINTS = npoints
    # shading points: primary hits of the bench camera on the ORIGINAL soup, moved with the scene; rays in full double
    a0 = hyb if v == "exact" else accel.Accel.bind().build(base, accel.PREC_F32)
    P, n = bench.primary_points(a0.intersect, base[a0.triorder()])
    rays8 = scenes.ao_rays(P[:npoints], n[:npoints], 8, 8, scenes.SEED_C3)
    rays = scenes.rays_f32_to_f64(rays8)
    Pm = P[:npoints].astype(np.float64) * scale + shift
    rays[:, 0:3] = np.repeat(Pm + 1.0e-6 * n[:npoints].astype(np.float64), 64, axis=0)       # origin = P + 1e-6 N in double
    nr = len(rays)
    d64 = torch.from_numpy(rays).cuda()
    o_h = torch.empty(nr, dtype=torch.uint8, device="cuda"); o_p = torch.empty(nr, dtype=torch.uint8, device="cuda")
    ms_h = ev(lambda: hyb.occluded_dev(d64, nr, o_h, st.cuda_stream, f64=True))
    os.environ["B200_HYBRID"] = "0"                                    # the double kernels alone (a double-only accelerator is hybrid too)
    ms_p = ev(lambda: plain.occluded_dev(d64, nr, o_p, st.cuda_stream, f64=True))
    os.environ.pop("B200_HYBRID")
    same = bool(torch.equal(o_h, o_p))
    line = f"{v:8s} {nr} rays: hybrid {ms_h:.3f} ms = {nr / ms_h / 1e3:.1f} Mrays/s | plain f64 {ms_p:.3f} ms = {nr / ms_p / 1e3:.1f} Mrays/s | identical {same} | occluded {float(o_p.float().mean()):.3f}"
    if v == "exact":
        d32 = torch.from_numpy(rays8).cuda()
        o32 = torch.empty(nr, dtype=torch.uint8, device="cuda")
        ms32 = ev(lambda: hyb.occluded_dev(d32, nr, o32, st.cuda_stream))
        line += f" | fp32 records {nr / ms32 / 1e3:.1f} Mrays/s, differ from double on {float((o32 != o_p).float().mean()):.2e} of the rays"
    print(line, flush=True)
    h_h = torch.empty((nr, 4), dtype=torch.float64, device="cuda"); h_p = torch.empty((nr, 4), dtype=torch.float64, device="cuda")
    ms_ch = ev(lambda: hyb.intersect_dev(d64, nr, h_h, st.cuda_stream, f64=True))
    os.environ["B200_HYBRID"] = "0"
    ms_cp = ev(lambda: plain.intersect_dev(d64, nr, h_p, st.cuda_stream, f64=True))
    os.environ.pop("B200_HYBRID")
    same_c = bool(torch.equal(h_h.view(torch.int64), h_p.view(torch.int64)))
    print(f"{v:8s} closest hit: hybrid {ms_ch:.3f} ms = {nr / ms_ch / 1e3:.1f} Mrays/s | plain f64 {ms_cp:.3f} ms = {nr / ms_cp / 1e3:.1f} Mrays/s | identical records {same_c}", flush=True)
    if not same:
        bad = torch.nonzero(o_h != o_p).flatten()[:5].cpu().numpy()
        print("  MISMATCH at rays", bad, "hybrid", o_h[bad].cpu().numpy(), "plain", o_p[bad].cpu().numpy())
    hyb.free(); plain.free()
