"""ncu report -> committed summary:  python scripts/ncu_md.py <report.ncu-rep> <profiles/name> "<title>" [rays_in_launch]
Writes <name>.md (metric table, instruction mix, hottest SASS lines) and <name>.json (the numbers bench.py's roofline.limiter quotes)."""
import csv, json, os, subprocess, sys
from collections import defaultdict

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
nrays = float(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, vals = rr[0], rr[1], rr[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active"]
name = m["Kernel Name"][0]
to_bytes = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
def fl(k):
    try:
        return float(m[k][0].replace(",", ""))
    except Exception:
        return None
rd = fl("dram__bytes_read.sum") * to_bytes[m["dram__bytes_read.sum"][1]]
wr = fl("dram__bytes_write.sum") * to_bytes[m["dram__bytes_write.sum"][1]]
ms = fl("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[m["gpu__time_duration.sum"][1]]
inst = fl("smsp__inst_executed.sum")
js = {"kernel": name.split("(")[0].replace("void ", ""), "source": os.path.basename(out) + ".md", "ms": ms,
      "dram_bytes": rd + wr, "dram_gb_s": (rd + wr) / (ms * 1e-3) / 1e9,
      "issue_frac": fl("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
      "fma_pipe_cycles_frac": fl("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") / 100.0,
      "alu_pipe_cycles_frac": fl("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active") / 100.0,
      "lsu_wavefront_frac": fl("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100.0,
      "l2_hit_rate": fl("lts__t_sector_hit_rate.pct") / 100.0, "l1_hit_rate": fl("l1tex__t_sector_hit_rate.pct") / 100.0,
      "lanes_per_inst": fl("smsp__thread_inst_executed_per_inst_executed.ratio"),
      "warps_active_frac": fl("sm__warps_active.avg.pct_of_peak_sustained_active") / 100.0, "registers": fl("launch__registers_per_thread")}
if nrays:
    js["rays"] = nrays
    js["inst_per_ray"] = inst / nrays
    js["mrays_s_under_ncu"] = nrays / ms / 1e3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sr = list(csv.reader(src.splitlines()))
h2 = sr[1]
data = [r for r in sr[2:] if len(r) == len(h2) and r[0] != "Address"]
ia, isrc, isamp, iavg = h2.index("Instructions Executed"), h2.index("Source"), h2.index("# Samples"), h2.index("Avg. Threads Executed")
total = sum(int(r[ia]) for r in data)
tots = max(1, sum(int(r[isamp]) for r in data))
opc = defaultdict(int)
for r in data:
    t = r[isrc].split()
    op = t[1] if t[0].startswith("@") else t[0]
    opc[op.split(".")[0]] += int(r[ia])
with open(out + ".md", "w") as f:
    f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on`, one launch.  kernel: `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
    for k in keys:
        if k in m:
            f.write(f"| {k} | {m[k][0]} | {m[k][1]} |\n")
    f.write(f"\nDRAM traffic of the launch: {rd / 1e6:.1f} MB read + {wr / 1e6:.1f} MB written = {(rd + wr) / 1e6:.1f} MB in {ms:.3f} ms = "
            f"{(rd + wr) / (ms * 1e-3) / 1e9:.1f} GB/s.\n")
    if nrays:
        f.write(f"Rays in the launch: {int(nrays)} -> {inst / nrays:.1f} warp instructions per ray, {nrays / ms / 1e3:.1f} Mrays/s under the profiler.\n")
    f.write(f"\n## instruction mix ({total} warp instructions, {len(data)} SASS lines)\n\n| opcode | share of executed warp instructions |\n|---|---:|\n")
    for k, v in sorted(opc.items(), key=lambda kv: -kv[1])[:24]:
        f.write(f"| {k} | {v / total * 100:.1f} % |\n")
    f.write("\n## lines with the most stall samples\n\n| samples | executed | avg threads | SASS |\n|---:|---:|---:|---|\n")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:20]:
        f.write(f"| {int(r[isamp]) / tots * 100:.2f} % | {int(r[ia]) / total * 100:.2f} % | {r[iavg]} | `{r[isrc].strip()[:90]}` |\n")
js["opcode_share"] = {k: v / total for k, v in sorted(opc.items(), key=lambda kv: -kv[1])[:12]}
with open(out + ".json", "w") as f:
    json.dump(js, f, indent=1)
print("wrote", out + ".md", out + ".json")
