#!/bin/bash
# usage: tools/ncu_durations.sh <skip> <count> <stepbench spec...> : per-launch gpu__time_duration and instruction count
skip=$1; cnt=$2; shift 2
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --launch-skip $skip -c $cnt --csv python tools/stepbench.py "$@" 2>/dev/null \
  | python -c "
import csv, sys
rows = [r for r in csv.reader(sys.stdin) if len(r) > 10 and r[0].isdigit()]
for r in rows:
    print(r[4][:60].ljust(60), r[-3].ljust(26), r[-1])
"
