#!/usr/bin/env python
"""Print the key numbers of a bench.py JSON line (last line of the given file)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("n_gpus", d.get("n_gpus"), "value %.4g" % d["value"], "ms/step %.5f" % d["ms_per_step"])
if "roofline" in d:
    r = d["roofline"]
    print("roofline", r["kernel"], "frac %.3f" % r["frac"], "achieved %.0f %s" % (r["achieved"], r["unit"]))
print("e2e %.4g" % d["e2e"]["value"], "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
if "cpu_baseline" in d:
    print("cpu_baseline %.4g on %d cores (%s)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"]))
for o in d.get("other_kernels", []):
    print("  %-72s %9.1f us  frac %.3f" % (o["kernel"][:72], o["us"], o["frac"]))
