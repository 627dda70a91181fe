"""Host vs device BVH build time (info.build_seconds) and tree equality at 100 K / 1 M / 10 M triangles."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucille_b200 import accel, scenes
for n, seed in ((100_000, scenes.SEED_C2), (1_000_000, scenes.SEED_C3), (10_000_000, 0xB2000005)):
    if len(sys.argv) > 1 and n > int(sys.argv[1]): break
    tris = scenes.triangle_soup(n, seed)
    res = {}
    for name, flag in (("host", accel.BUILD_HOST), ("device", accel.BUILD_DEVICE), ("device2", accel.BUILD_DEVICE)):
        t0 = time.perf_counter()
        a = accel.Accel.bind().build(tris, accel.PREC_F32 | flag)
        wall = time.perf_counter() - t0
        i = a.info()
        res[name] = (i.build_seconds, wall, i.ninner, i.max_depth, a.triorder())
        a.free()
    same = np.array_equal(res["host"][4], res["device"][4]) and res["host"][2:4] == res["device"][2:4]
    print(f"N={n}: host build {res['host'][0]*1e3:.1f} ms (call {res['host'][1]*1e3:.0f} ms) | device build {res['device'][0]*1e3:.1f} ms, "
          f"second call {res['device2'][0]*1e3:.1f} ms (call {res['device2'][1]*1e3:.0f} ms) | inner {res['host'][2]} depth {res['host'][3]} same order {same}", flush=True)
