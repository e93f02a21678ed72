# refresh the judged profile artefacts (run on the GPU box; writes gpurun_out/)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --calib-frames 0"
$B > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none -k regex:conv_umma --launch-skip 5 --launch-count 5 -f -o /tmp/full $B > gpurun_out/ncu_full.log 2>&1
python profiles/ncu_tools.py metrics /tmp/full.ncu-rep > gpurun_out/ncu_metrics.txt
