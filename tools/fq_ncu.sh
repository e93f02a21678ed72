#!/bin/bash
# ncu --set full capture (source-level stall samples) of one exe_mode-0 (EPI_FQ) 3x3 convolution launch of the calibration
# leg; CSV exports only
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-variants --no-workloads --no-python-ref"
$B > /dev/null 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:conv_umma_kernelILi3ELi4E --launch-skip 4 --launch-count 1 -f -o /tmp/fqconv $B > gpurun_out/fq_ncu.log 2>&1
ncu -i /tmp/fqconv.ncu-rep --page raw --csv > gpurun_out/fq_conv.raw.csv 2>/dev/null
ncu -i /tmp/fqconv.ncu-rep --page source --csv > gpurun_out/fq_conv.src.csv 2>/dev/null
python profiles/ncu_tools.py metrics /tmp/fqconv.ncu-rep
python profiles/ncu_tools.py hot /tmp/fqconv.ncu-rep 0 30
