# ncu captures of the final-layer kernel (run only after the plain command exited 0); only the CSV exports are kept
# usage: CAPS="both y q5" bash tools/l4_ncu.sh <tag>
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --calib-frames 0"
N="ncu --set full --import-source on --clock-control none -k regex:conv_umma_kernel --launch-skip 15 --launch-count 1 -f"
$B --outputs both > /dev/null 2>&1 || exit 1
for o in ${CAPS:-both y}; do
  n=l4_$1_$o
  timeout 300 $N -o /tmp/$n $B --outputs $o > gpurun_out/$n.log 2>&1
  ncu -i /tmp/$n.ncu-rep --page raw --csv > gpurun_out/$n.raw.csv 2>/dev/null
  ncu -i /tmp/$n.ncu-rep --page source --csv > gpurun_out/$n.src.csv 2>/dev/null
  rm -f /tmp/$n.ncu-rep
done
