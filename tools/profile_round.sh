# refresh the judged profile artefacts (run on the GPU box only after the plain command exited 0; writes gpurun_out/)
#   launches.csv          ncu launch list (gpu__time_duration) of one default bench command
#   ncu_full_y.raw.csv    ncu --set full of the 4 kernels of one default step (y only)
#   ncu_full_both.raw.csv the same for --outputs both
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --calib-frames 0 --no-python-ref"
$B > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
for o in y both; do
  ncu --set full --clock-control none -k regex:conv_ --launch-skip 4 --launch-count 4 -f -o /tmp/full_$o $B --outputs $o --no-variants --no-workloads > gpurun_out/ncu_full_$o.log 2>&1
  ncu -i /tmp/full_$o.ncu-rep --page raw --csv > gpurun_out/ncu_full_$o.raw.csv 2>/dev/null
  python profiles/ncu_tools.py metrics /tmp/full_$o.ncu-rep > gpurun_out/ncu_metrics_$o.txt
done
