"""Time the REAL reference (the unmodified Python myQL + models, staged in oracle/_ref by oracle/stage_ref.py, or
/root/reference in the build container) on this host's cores: the sim.py path (exe_mode 1) or the test.py path
(exe_mode 0), frame by frame as upstream runs it (batch 1, quan_func.py:349-356).  MEASUREMENT INFRASTRUCTURE ONLY:
called by bench.py's cpu_baseline / --impl reference legs in a subprocess (scratch cwd: the reference keeps its state in
./output_pt); prints ONE JSON line.

    python oracle/time_ref.py --case tests/golden/x2_fixture.npz --h 1080 --w 1920 --frames 1 [--mode 1]
"""
import argparse
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True)
    ap.add_argument("--h", type=int, required=True)
    ap.add_argument("--w", type=int, required=True)
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--mode", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    if "SESRQ_REFERENCE_ROOT" not in os.environ:
        os.environ["SESRQ_REFERENCE_ROOT"] = "/root/reference" if os.path.isdir("/root/reference") else os.path.join(HERE, "_ref")
    nthr = a.threads or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count())
    os.environ["OMP_NUM_THREADS"] = str(nthr)          # a launcher's value (torchrun: 1) must not leak in
    os.environ["MKL_NUM_THREADS"] = str(nthr)
    import numpy as np
    import torch
    torch.set_num_threads(nthr)
    sys.path.insert(0, HERE)
    import ref_harness as rh
    if not os.path.isfile(os.path.join(rh.REF, "myQL", "quan_func.py")):
        print(json.dumps({"unavailable": "reference not staged at %s" % rh.REF}))
        return
    z = np.load(os.path.abspath(a.case))
    mflag = int(z["mflag"])
    cin = z["wf0"].shape[1]
    case = os.path.abspath(a.case)
    os.chdir(tempfile.mkdtemp(prefix="sesrq_ref_"))
    ns = rh.import_reference(mflag, dumps=False)
    model = rh.build_float_model(ns, mflag, sim=(a.mode == 1))
    if a.mode == 1:
        rh.write_thresholds([float(v) for v in z["scales"]], [int(v) for v in z["zeros"]])
    os.makedirs("output_pt/weight", exist_ok=True)
    model = rh.quantize_graph(ns, model, a.mode)
    g = torch.Generator().manual_seed(4321)
    with torch.no_grad():
        model(torch.rand(1, cin, 16, 32, generator=g))                      # warm-up: module import, first-call costs
        times = []
        for _ in range(a.frames):
            x = torch.rand(1, cin, a.h, a.w, generator=g)
            t0 = time.perf_counter()
            model(x)
            times.append(time.perf_counter() - t0)
    print(json.dumps({"s_per_frame_best": min(times), "s_per_frame_mean": sum(times) / len(times), "frames": a.frames,
                      "threads": torch.get_num_threads(), "mode": a.mode, "mflag": mflag, "h": a.h, "w": a.w,
                      "reference_root": os.path.basename(rh.REF), "case": os.path.basename(case)}))


if __name__ == "__main__":
    main()
