#!/usr/bin/env python
"""Aggregate an .ncu-rep's source page by CUDA source line: stall samples (with the dominant stall reasons)
and warp instructions.   usage: ncu_lines.py report.ncu-rep [top] [kernel-name-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3: cmd += ["-k", "regex:" + sys.argv[3]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; seen = {}; stall_cols = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No":
        hdr = r
        stall_cols = [(i, c[6:]) for i, c in enumerate(r) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr and len(r) >= 8 and r[0] != "":
        try:
            k = (cur.split('/')[-1], int(r[0]))
            if k not in seen:
                st = {}
                for i, name in stall_cols:
                    try: st[name] = int(r[i])
                    except (ValueError, IndexError): pass
                seen[k] = (r[1].strip(), int(r[6]), int(r[7]), st)
        except ValueError:
            pass
agg = [(k[0], k[1]) + v for k, v in seen.items()]
ts, ti = sum(a[3] for a in agg), sum(a[4] for a in agg)
tot = {}
for a in agg:
    for n, v in a[5].items(): tot[n] = tot.get(n, 0) + v
print(f"total stall samples {ts}, warp instructions {ti}")
print("stall reasons: " + "  ".join(f"{n} {v}" for n, v in sorted(tot.items(), key=lambda x: -x[1]) if v))
def why(st):
    return " ".join(f"{n}:{v}" for n, v in sorted(st.items(), key=lambda x: -x[1])[:3] if v)
print("--- by instructions")
for a in sorted(agg, key=lambda a: -a[4])[:top]:
    print("%5d smp %9d inst %5.1f%%  %s:%d  %s" % (a[3], a[4], 100.0 * a[4] / ti, a[0], a[1], a[2][:95]))
print("--- by samples")
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print("%5d smp %9d inst  %-22s %-40s %s" % (a[3], a[4], "%s:%d" % (a[0], a[1]), why(a[5]), a[2][:80]))
