"""torchrun --nproc-per-node N scripts/dist_frame_check.py : render one synthetic AO frame with tiles sharded over N GPUs, gather over
NCCL, and check on rank 0 that it equals the single-GPU frame bit for bit (counter-based RNG).  Prints device time per rank."""
import math, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lucille_b200 import accel, scenes, distributed

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ntris = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
tris = scenes.triangle_soup(ntris, scenes.SEED_C5)
a = accel.Accel.bind().build(tris, accel.PREC_F32, device=local)
c2w = np.eye(4); c2w[3, :3] = (0.5, 0.5, -2.0)
fr = accel.make_frame(c2w.reshape(16), 1.0 / math.tan(math.radians(40.0) / 2), False, res, res, 4, 4, 64, rng_mode=1, seed=5, precision=accel.PREC_F32)
for it in range(2):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rgb, stats = distributed.render_ao_distributed(a, fr, rank, world)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t = torch.tensor([dt, stats.ms_total * 1e-3, float(stats.nrays)], device="cuda", dtype=torch.float64)
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
    else:
        allt = [t]
    if rank == 0:
        rays = sum(float(x[2]) for x in allt); wall = max(float(x[0]) for x in allt)
        print(f"iter {it}: {world} GPU(s) {res}x{res} 16 spp x 64 AO rays on {ntris} tris: {rays/1e6:.1f} Mrays in {wall*1e3:.1f} ms wall "
              f"(device max {max(float(x[1]) for x in allt)*1e3:.1f} ms) -> {rays/wall/1e6:.1f} Mrays/s incl. gather", flush=True)
# the same frame with the gather fused into the resolve kernels (peer memory)
fb = distributed.PeerFramebuffer(res, res, rank, world, local)
for it in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rgb_p, stats_p = distributed.render_ao_distributed_peer(a, fr, fb)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"peer iter {it}: {float(t.item())*1e3:.1f} ms wall, equal to the NCCL-gathered frame: {bool(np.array_equal(rgb_p, rgb))}", flush=True)
# the reference's single MT19937 stream shared by the ranks (hit counts all-gathered between the eye pass and the gather pass)
fr0 = accel.make_frame(c2w.reshape(16), 1.0 / math.tan(math.radians(40.0) / 2), False, res, res, 2, 2, 64, rng_mode=0, seed=4357, precision=accel.PREC_F32)
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
rgb0, _ = distributed.render_ao_distributed_peer(a, fr0, fb)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
if rank == 0:
    one0, _ = a.render_ao(fr0)
    print(f"MT19937 stream over {world} rank(s): {dt*1e3:.1f} ms wall, equal to the single-GPU rng_mode-0 frame: {bool(np.array_equal(rgb0, one0))}", flush=True)
fb.close()
if rank == 0 and world > 1:
    full, _ = a.render_ao(fr)
    print("equal to single-GPU frame:", bool(np.array_equal(full, rgb)), "mean", float(rgb.mean()), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
