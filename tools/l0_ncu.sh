#!/bin/bash
# GPU box: ncu --set full capture with source-level stall samples of ONE launch of a kernel of the default bench step.
# usage: tools/l0_ncu.sh <kernel-regex> <tag> [launch-skip]
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --calib-frames 0 --no-variants --no-workloads --no-python-ref"
$B > /dev/null 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$1 --launch-skip ${3:-3} --launch-count 1 -f -o /tmp/$2 $B > gpurun_out/$2_ncu.log 2>&1
ncu -i /tmp/$2.ncu-rep --page raw --csv > gpurun_out/$2.raw.csv 2>/dev/null
ncu -i /tmp/$2.ncu-rep --page source --csv > gpurun_out/$2.src.csv 2>/dev/null
python profiles/ncu_tools.py metrics /tmp/$2.ncu-rep | head -30
python profiles/ncu_tools.py hot /tmp/$2.ncu-rep 0 45
