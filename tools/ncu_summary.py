#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): key metrics per captured launch, plus the
hottest SASS lines by stall samples.  Usage: tools/ncu_summary.py report.ncu-rep [--top 25]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_write_bytes.sum",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("== launch", d.get("ID"), d.get("Kernel Name", "")[:60])
        for k in KEYS:
            if k in d:
                print(f"  {k:70s} {d[k]:>16s} {units[hdr.index(k)]}")
        stalls = [(float(v.replace(",", "")), h) for h, v in d.items()
                  if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.2f}")


def source(rep, top):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    recs = []
    for r in rows:
        if len(r) > 5 and r[0] == "Address":
            if hdr is not None:
                break               # first launch only
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            recs.append(dict(zip(hdr, r)))
    if not recs:
        return
    tot = sum(int(r["# Samples"] or 0) for r in recs)
    tot_inst = sum(int(r["Instructions Executed"] or 0) for r in recs)
    print(f"== source: {len(recs)} SASS lines, {tot} stall samples, {tot_inst} warp instructions")
    for r in sorted(recs, key=lambda r: -int(r["# Samples"] or 0))[:top]:
        print(f"  {int(r['# Samples']):6d} smp {int(r['Instructions Executed']):9d} inst  {r['Source'].strip()[:90]}")


if __name__ == "__main__":
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw(rep)
    source(rep, top)
