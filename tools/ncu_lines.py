#!/usr/bin/env python
"""Aggregate an .ncu-rep's source page by CUDA source line: stall samples and warp instructions."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; seen = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) >= 8 and r[0] != "":
        try:
            k = (cur.split('/')[-1], int(r[0]))
            if k not in seen: seen[k] = (r[1].strip(), int(r[6]), int(r[7]))
        except ValueError:
            pass
agg = [(k[0], k[1]) + v for k, v in seen.items()]
ts, ti = sum(a[3] for a in agg), sum(a[4] for a in agg)
print(f"total stall samples {ts}, warp instructions {ti}")
print("--- by instructions")
for a in sorted(agg, key=lambda a: -a[4])[:top]:
    print("%5d smp %9d inst %5.1f%%  %s:%d  %s" % (a[3], a[4], 100.0 * a[4] / ti, a[0], a[1], a[2][:95]))
print("--- by samples")
for a in sorted(agg, key=lambda a: -a[3])[:top // 2]:
    print("%5d smp %9d inst  %s:%d  %s" % (a[3], a[4], a[0], a[1], a[2][:95]))
