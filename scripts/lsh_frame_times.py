"""The unmodified ambient_occlusion.rib through the reference front end: wall time and the renderer's own "Render frame" timer for
  (a) lsh_b200 --accel b200          batched frame hook (integration/ri_b200_frame_hook.c): ONE ri_b200_render_ao call
  (b) oracle_rib --nthreads 1        the compiled reference, CPU BVH, the only deterministic mode
  (c) oracle_rib --nthreads <all>    the compiled reference on every host core
  (d) lsh_b200 --accel b200 with RI_B200_FRAME=0 at 64x48, 2x2, 16 rays: the per-ray vtable slot (one launch per ri_raytrace)
Prints a markdown table (profiles/r02_lsh_frame.md)."""
import os, subprocess, sys, tempfile, time, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol

ref = ol.REF_DIR
rib = os.path.join(ref, "scenes", "ambient_occlusion.rib")
cores = os.cpu_count() or 1
dig = np.load(os.path.join(ROOT, "tests", "golden", "c1_frame_640x480_digest.npz"))
rows = []
def run(label, exe, args, env=None):
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "f.bin")
        e = dict(os.environ); e.update(env or {})
        t0 = time.perf_counter()
        subprocess.run([os.path.join(ref, exe), rib, "--out", out] + args, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=e, timeout=3600)
        wall = time.perf_counter() - t0
        rgb, sec, nrays = ol.read_frame(out)
        sha = hashlib.sha256(np.ascontiguousarray(rgb, dtype=np.float32).tobytes()).hexdigest()
        rows.append((label, rgb.shape[1], rgb.shape[0], nrays, sec, wall, sha == str(dig["sha256"])))
        print(f"{label}: {rgb.shape[1]}x{rgb.shape[0]} rays {nrays} render-frame {sec:.3f}s wall {wall:.2f}s equals-reference-1-thread-frame {sha == str(dig['sha256'])}", flush=True)
run("lsh_b200 --accel b200 (batched hook, 1st run incl. CUDA context)", "lsh_b200", ["--accel", "b200", "--nthreads", "1"])
run("lsh_b200 --accel b200 (batched hook)", "lsh_b200", ["--accel", "b200", "--nthreads", "1"])
run(f"oracle_rib --nthreads {cores} (reference, all host threads)", "oracle_rib", ["--nthreads", str(cores)])
run("oracle_rib --nthreads 1 (reference)", "oracle_rib", ["--nthreads", "1"])
small = ["--nthreads", "1", "--width", "64", "--height", "48", "--pixelsamples", "2", "--gather", "16"]
run("lsh_b200 --accel b200 RI_B200_FRAME=0 (per-ray slot), 64x48 2x2 16 rays", "lsh_b200", ["--accel", "b200"] + small, {"RI_B200_FRAME": "0"})
run("oracle_rib --nthreads 1 (reference), 64x48 2x2 16 rays", "oracle_rib", small)
print("\n| run | frame | rays (render->stat.nrays) | renderer's 'Render frame' timer | process wall | Mrays/s (timer) | == reference 1-thread 640x480 frame |\n|---|---|---:|---:|---:|---:|---|")
for label, w, h, nrays, sec, wall, same in rows:
    print(f"| {label} | {w}x{h} | {nrays} | {sec:.3f} s | {wall:.2f} s | {nrays / max(sec, 1e-9) / 1e6:.2f} | {same if (w, h) == (640, 480) else 'n/a'} |")
print(f"\nhost threads: {cores}")
