#!/bin/bash
# One gpurun call that produces every ncu artefact summarised under profiles/ (run from the repo root on the GPU box):
#   tools/profile_round.sh r01
# Each capture runs only after the same command has exited 0 without ncu.
tag=${1:-r01}
out=gpurun_out
set -x
B="python bench.py --steps 120 --warmup 20 --no-graph --no-cpu-baseline --no-other-kernels --e2e-steps 5"
$B > $out/${tag}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:rt_step_kernel -s 140 -c 1 -f -o $out/${tag}_step $B > $out/${tag}_step.log 2>&1
python tools/stage_clock.py 4096 > $out/${tag}_stage_clock.txt 2>&1
python tools/convbench.py 592 > $out/${tag}_conv_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:rt_conv1_tc -s 3 -c 1 -f -o $out/${tag}_conv1 python tools/convbench.py 592 > $out/${tag}_conv1.log 2>&1
python tools/densebench.py > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:rt_dense_kernel -s 4 -c 1 -f -o $out/${tag}_dense python tools/densebench.py > $out/${tag}_dense.log 2>&1
python tools/volbench.py > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:rt_volumes -s 2 -c 1 -f -o $out/${tag}_volumes python tools/volbench.py > $out/${tag}_volumes.log 2>&1
python tools/gaebench.py > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:rt_gae -s 2 -c 1 -f -o $out/${tag}_gae python tools/gaebench.py > $out/${tag}_gae.log 2>&1
ls -la $out | tail -20
