"""Group the SASS lines of an ncu source page into straight-line segments (equal execution counts):
python scripts/ncu_segments.py report.ncu-rep [min_share_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))[2:]
tot = sum(int(r[5]) for r in rows); tots = sum(int(r[2]) for r in rows)
print("warp instructions", tot, "stall samples", tots)
segs = []
for i, r in enumerate(rows):
    ex, st, th = int(r[5]), int(r[2]), int(r[6])
    if segs and abs(segs[-1]['ex'] - ex) <= max(1, ex * 0.002):
        s = segs[-1]; s['n'] += 1; s['st'] += st; s['tot'] += ex; s['thr'] += th; s['end'] = i
    else:
        segs.append(dict(start=i, end=i, ex=ex, n=1, st=st, tot=ex, thr=th, first=r[1].strip()))
for s in segs:
    if s['tot'] / tot * 100 < min_share: continue
    print(f"{s['start']:4d}-{s['end']:4d} n={s['n']:3d} execs={s['ex']/1e6:8.2f}M share={s['tot']/tot*100:5.2f}% thr={s['thr']/max(1,s['tot']):5.1f} stall={s['st']/tots*100:5.2f}%  {s['first'][:60]}")
