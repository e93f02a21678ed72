#!/bin/bash
# GPU box: kernel durations of the calibration leg (256 frames of 3x80x960): ncu launch list -> per-kernel-name totals
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-variants --no-workloads --no-python-ref"
$B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('calib', d['calibration'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/calib_launches.csv $B > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/calib_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
seq = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > vi]
tot = collections.OrderedDict()
for k, v in seq:
    k = k.split("(")[0][:70]
    a = tot.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, v) in tot.items():
    print("%6d x %-72s %10.1f us total  %8.1f us avg" % (n, k, v / 1e3, v / 1e3 / n))
PY
