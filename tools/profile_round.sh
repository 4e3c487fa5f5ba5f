#!/bin/bash
# One gpurun call that produces the ncu artefacts summarised under profiles/ (run from the repo root on the GPU box):
#   tools/profile_round.sh r02
# Each capture runs only after the same command has exited 0 without ncu.
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out
set -x
B="python bench.py --steps 120 --warmup 20 --no-graph --no-cpu-baseline --no-other-kernels --no-ppo --no-dense --e2e-steps 5"
$B > $out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file $out/launches.csv $B > $out/launches.log 2>&1
python tools/stepbench.py 4096 nopdl:4096 eager:4096 nopdl,eager:4096 8192 65536 > $out/stepbench.txt 2>&1
python tools/stepbench.py eager:4096 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:rt_step_kernel -s 300 -c 1 -f -o $out/step4096 python tools/stepbench.py eager:4096 > $out/ncu4096.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rt_step_kernel -s 230 -c 1 -f -o $out/step65536 python tools/stepbench.py eager:65536 > $out/ncu65536.log 2>&1
python tools/stage_clock.py 4096 > $out/stage_clock.txt 2>&1
python tools/volbench.py > $out/vol.txt 2>&1 && ncu --set full --clock-control none -k regex:rt_volumes -s 2 -c 1 -f -o $out/volumes python tools/volbench.py > $out/volumes.log 2>&1
ls -la $out | tail -20
