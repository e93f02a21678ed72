#!/bin/bash
# GPU box: parity tests, then one default bench line; prints a compact summary.  Usage: tools/gpu_quick.sh <tag> [bench args]
TAG=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -4 gpurun_out/${TAG}_tests.log
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("ms/step", round(d["ms_per_step"], 4), "stack", {k: round(v, 3) for k, v in d["stack"].items()})
    print("layers", [(round(l["avg_ms"], 4), round(l["achieved_gbs"])) for l in d["layers"]])
    print("single", {k: round(v["ms_per_step"], 4) for k, v in (d.get("output_variants") or {}).items()})
    print("e2e", d.get("e2e"))
    print("clocks", d.get("clocks"))
except Exception as e:
    print("bench failed:", e)
    print(open("gpurun_out/${TAG}_bench.err").read()[-2000:])
PY
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    for w in d.get("workloads") or []:
        print("wl", w.get("workload"), w.get("frames_per_step"), w.get("skipped") or "%.4f ms %.1f us/frame frac %.3f" % (w["ms_per_step"], w["us_per_frame"], w["frac_of_hbm_peak"]))
    print("calib", d.get("calibration"))
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no workloads:", e)
PY
