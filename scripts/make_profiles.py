"""Turn what `scripts/gpu_r2p.sh` leaves in gpurun_out/ into the committed summaries under profiles/ (round 2 and later):

    gpurun -- 'bash scripts/gpu_r2p.sh'      # on the GPU box: launch list + one `ncu --set full` capture per kernel, summarised there
    python scripts/make_profiles.py r02      # here: copy the summaries, derive the files bench.py reads

  profiles/<tag>_<kernel>.md / .json     metric table, instruction mix, hottest SASS lines (scripts/ncu_md.py, run on the box)
  profiles/<tag>_kernel_metrics.json     the `roofline.limiter` block of the bench line: what physically limits the timed kernel
  profiles/traffic.json                  DRAM bytes of one launch of the timed kernel (`roofline.traffic`)
  profiles/<tag>_launches.md             kernel launch list of the bench command (shares, not absolutes: cold cache, serialised)
(The round-1 form of this script, which read .ncu-rep files directly, is in the history: profiles/r01_* came from it.)"""
import csv
import glob
import json
import os
import shutil
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out_dir, gp = os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out")

for f in sorted(glob.glob(os.path.join(gp, "summaries", f"{tag}_*"))):
    shutil.copy(f, out_dir)
    print("copied", os.path.basename(f))

timed = json.load(open(os.path.join(out_dir, f"{tag}_occluded_f32_c3.json")))
lsu, issue = timed["lsu_wavefront_frac"], timed["issue_frac"]
km = {
    "source": f"profiles/{tag}_occluded_f32_c3.md (ncu --set full, one launch of the timed kernel on the bench batch; scripts/make_profiles.py)",
    "what_limits": ("L1 data-pipe (LSU) wavefronts, with warp-instruction issue next" if lsu >= issue else "warp-instruction issue, with the L1 data pipe next")
                   + "; DRAM is idle (records are L2-resident)",
    "lsu_wavefront_frac": lsu, "issue_frac": issue, "fma_pipe_cycles_frac": timed["fma_pipe_cycles_frac"],
    "alu_pipe_cycles_frac": timed["alu_pipe_cycles_frac"], "inst_per_ray": timed.get("inst_per_ray"), "lanes_per_inst": timed["lanes_per_inst"],
    "dram_gb_s": timed["dram_gb_s"], "dram_bytes_per_launch": timed["dram_bytes"], "l2_hit_rate": timed["l2_hit_rate"],
    "l1_hit_rate": timed["l1_hit_rate"], "registers": timed["registers"], "ms_under_ncu": timed["ms"],
}
json.dump(km, open(os.path.join(out_dir, f"{tag}_kernel_metrics.json"), "w"), indent=1)
json.dump({"occluded_f32_c3_bytes_per_launch": timed["dram_bytes"], "source": f"profiles/{tag}_occluded_f32_c3.md"},
          open(os.path.join(out_dir, "traffic.json"), "w"))
print("wrote", f"{tag}_kernel_metrics.json", "traffic.json")

lc = os.path.join(gp, f"{tag}_launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5 and r[0].isdigit()]
    d = defaultdict(list)
    for r in rows:
        d[(r[4].split("(")[0].replace("void ", ""), r[8], r[7])].append(float(r[-1].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: kernel launch list of `B200_STREAMED=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-frames`\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 900` (cold-cache, serialised: compare shares, not absolutes).\n\n"
                "| kernel | grid | block | launches | avg ms | share of listed GPU time |\n|---|---|---|---:|---:|---:|\n")
        for (k, grid, block), v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {grid} | {block} | {len(v)} | {sum(v) / len(v) / 1e6:.3f} | {sum(v) / tot * 100:.1f} % |\n")
        f.write("\nThe timed region of bench.py is `steps` launches of `occluded_pool32_kernel<20, 0>` on the 16 Mi-ray batch and nothing else (one launch = one\n"
                "step; `gpu_launches` = steps), so the kernel's share of a step is 100 %.  Everything else in the list is set-up (device BVH build `gb_*`,\n"
                "primary rays, the reference-order counters `trace_batch_kernel<..., 1>` of the roofline formula), the side measurements (closest hit\n"
                "`closest_pool32_kernel`, the double-exact leg `occluded_hybrid_kernel` / `closest_hybrid_kernel` against `occluded_pool_kernel<double>` /\n"
                "`closest_pool_kernel<double>`) and the end-to-end calls (`ao_points_gen_kernel` + `occluded_pool32_kernel<20, 1>` = the point entry;\n"
                "the short `occluded_pool32_kernel<20, 0>` launches = the host ray-batch path in its launch-per-piece form, because ncu serialises the\n"
                "copy stream behind the kernel: outside the profiler an e2e call is ONE streamed launch).\n")
    print("wrote", f"{tag}_launches.md")
