#!/usr/bin/env python
"""bench.py -- SESR int8 hardware-sim throughput on B200 (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--batch B]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (5 fused conv layers: quantise -> int8 conv -> PE tree -> requant ->
ReLU/residual -> pixel shuffle) over one batch of synthetic frames per GPU.  Default workload = BASELINE.json's
headline metric: SESR x2 int8 sim on 1080p input (1920x1080 -> 3840x2160).  Frames are sharded data-parallel,
per-GPU batch fixed (weak scaling), no data-path collective.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (golden case holding weights + thresholds, H, W, default per-GPU batch)
    "sesr_x2_int8_sim_1080p": ("x2_fixture", 1080, 1920, 16),
    "sesr_x4_int8_sim_540x960": ("x4_fixture", 540, 960, 64),
    "nrdm3_int8_sim_80x960": ("nrdm_fixture", 80, 960, 512),
    "sesr_x2_int8_sim_80x960": ("x2_fixture", 80, 960, 256),
    "sesr_x4_int8_sim_80x960": ("x4_fixture", 80, 960, 256),
}
METRIC = "output Mpixels/sec (SESR x2 int8 sim)"


def layer_bytes_per_px(cin, cout, want_q5=True, want_y=True):
    """Algorithmic HBM bytes per LR pixel, per layer (SURVEY 8d per-layer-fused model): every activation
    tensor crosses HBM once per direction as int8 NHWC16, fp32 input read once, shortcut re-read once."""
    last = 16 + (cout if want_q5 else 0) + (4 * cout if want_y else 0)
    return [4 * cin + 16, 32, 32, 48, last]


def load_case(name):
    from sesr_pytorch_quantize_b200.fixtures import load_golden
    return load_golden(name)


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def workload_config(workload, case, B, H, W, r, cin, cout, outputs, world, input_mb):
    """The `config` object, shared by the product arm and the reference arm (same keys and values for the same flags)."""
    return {"workload": workload, "frames_per_gpu_per_step": B, "H": H, "W": W, "upscale": r, "cin": cin, "cout": cout,
            "outputs": {"both": "q5 int8 NCHW + fp32 pixel-shuffled y", "y": "fp32 pixel-shuffled y (what model(inps) returns, sim.py:205)",
                        "q5": "q5 int8 NCHW (input.5.pt)"}[outputs],
            "weights": "reference checkpoint collapsed (tests/golden/%s.npz), reference min/max thresholds" % case,
            "input": "torch.rand seed 4321+rank, %.0f MB per step > 126 MB L2 (no L2 flush needed)" % input_mb,
            "parallelism": "dp%d frames sharded, no collective" % world}


def bind_to_gpu_numa(torch, index):
    """Bind this rank to the CPUs local to its GPU (sysfs local_cpulist of the PCI device), so that pinned host buffers
    are first-touched on the GPU's NUMA node.  Returns a short description for the bench line."""
    try:
        p = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        node = open(base + "/numa_node").read().strip()
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"pci": bus, "numa_node": int(node), "cpus": len(cpus)}
    except (OSError, ValueError, AttributeError) as e:
        return {"unavailable": str(e)[:120]}


def copy_ceiling(torch, dev, h2d_bytes, d2h_bytes, reps, barrier):
    """What the PCIe / host-memory path alone allows for one step's transfers: the step's H2D and D2H volumes as two
    concurrent pinned copies on two streams, no kernels (all ranks at once).  Returns ms per step."""
    hi = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    ho = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    di = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    do = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    def once():
        with torch.cuda.stream(s1):
            di.copy_(hi, non_blocking=True)
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)
    once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / reps
    barrier()
    return dt * 1e3


def python_reference_time(case, H, W, frames=1, mode=1, timeout=900):
    """The REAL reference (unmodified Python myQL + models staged in oracle/_ref) timed on this host in a subprocess
    (oracle/time_ref.py).  Returns the parsed JSON or {"unavailable": why}."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "time_ref.py"), "--case", os.path.join(ROOT, "tests", "golden", case + ".npz"),
           "--h", str(H), "--w", str(W), "--frames", str(frames), "--mode", str(mode), "--threads", str(host_cores())]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if p.returncode != 0 or not lines:
            return {"unavailable": "oracle/time_ref.py failed: " + (p.stderr.strip().splitlines() or ["?"])[-1][:200]}
        return json.loads(lines[-1])
    except (OSError, subprocess.TimeoutExpired, ValueError) as e:
        return {"unavailable": str(e)[:200]}


def sample_clocks_start(gpu_index):
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                 "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
    except OSError:
        return None, f.name
    return proc, f.name


def clock_rows(path):
    """Rows nvidia-smi has written so far."""
    try:
        return sum(1 for r in open(path) if r.strip())
    except OSError:
        return 0


def sample_clocks_wait_ready(proc, path, timeout=4.0):
    """Block until the sampler has produced its first row (nvidia-smi takes a few hundred ms to come up); returns the
    number of rows written before the load starts -- those are idle-clock samples and are dropped."""
    if proc is None:
        return 0
    t0 = time.time()
    while clock_rows(path) == 0 and time.time() - t0 < timeout:
        time.sleep(0.02)
    return clock_rows(path)


def sample_clocks_stop(proc, path, skip_rows=0):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    if proc is not None:
        proc.terminate()
        try:
            proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            proc.kill()
    try:
        rows = [r.strip().split(",") for r in open(path) if r.strip()][skip_rows:]
        os.unlink(path)
    except OSError:
        rows = []
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in rows:
        try:
            sm.append(float(r[1]))
            mx.append(float(r[2]))
            for k, nm in enumerate(names):
                if r[5 + k].strip().lower().startswith("active"):
                    reasons.add(nm)
        except (ValueError, IndexError):
            continue
    if sm:
        out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm), reasons=sorted(reasons))
    return out


class NvmlClockSampler:
    """SM clock and clock-event reasons read through NVML from a background thread, fast enough (about one sample per
    millisecond) to land INSIDE a timed region of a few milliseconds, which nvidia-smi's 100 ms period cannot."""

    def __init__(self, torch, index):
        self.ok, self.samples, self._stop, self._thread = False, [], False, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            p = torch.cuda.get_device_properties(index)
            bus = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self._read()
            self.ok = True
        except Exception:                                    # no NVML: the caller falls back to nvidia-smi
            self.ok = False

    def _read(self):
        return (float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)), int(self._reasons(self.h)))

    def start(self):
        if not self.ok:
            return
        import threading
        self.samples, self._stop = [], False

        def loop():
            while not self._stop:
                try:
                    self.samples.append(self._read())
                except Exception:
                    break
                time.sleep(0.0005)
        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.ok or self._thread is None:
            return out
        self._stop = True
        self._thread.join(timeout=2)
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        if self.samples:
            mask = 0
            for _, m in self.samples:
                mask |= m
            out.update(sm_mhz=statistics.median(c for c, _ in self.samples), sm_max_mhz=self.max_mhz, samples=len(self.samples),
                       reasons=sorted(k for k, b in bits.items() if mask & b), source="NVML, ~1 sample/ms inside the timed region")
        return out


def emit(line):
    """Print the ONE JSON line on the real stdout (everything else -- NCCL banners, warnings -- goes to stderr)."""
    sys.stdout.flush()
    os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


# libraries (NCCL prints its version banner on stdout) must not pollute the one-line contract: route fd 1 to stderr
# until the result is printed
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def cpu_reference_time(case, H, W, frames, reps):
    """Times the CPU restatement of the reference path (oracle/sesr_oracle.c, all host threads) on `frames`
    frames of the workload shape.  Returns (best seconds, threads)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import c_oracle
    import sesr_oracle as orc
    g = load_case(case)
    onet = orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r)
    cn = c_oracle.CNet(onet)
    c_oracle.set_threads(host_cores())           # explicit: torchrun exports OMP_NUM_THREADS=1
    x = np.random.default_rng(4321).random((frames, g.wf[0].shape[1], H, W), dtype=np.float32)
    best = None
    for _ in range(reps):
        t = time.perf_counter()
        cn.forward(x)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return best, c_oracle.threads()


def run_reference(args, case, H, W, r, B, cin, cout):
    """`--impl reference`: the reference path on this box's host cores with the product arm's config (same frames per
    step).  Primary number = the C restatement (oracle/sesr_oracle.c, all host threads: it is 17-27x faster than the
    Python reference, so ratios against it are conservative); the real Python reference (oracle/_ref) is timed on one
    frame beside it and reported in `python_reference`.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    out_px = H * W * r * r
    frames = B
    for _ in range(max(args.warmup, 0)):
        cpu_reference_time(case, H, W, min(frames, 2), 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_time(case, H, W, frames, 1)
    dt = (time.perf_counter() - t0) / args.steps
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    val = frames * out_px / dt / 1e6
    pyref = None
    if not args.no_python_ref:
        pr = python_reference_time(case, H, W, frames=1, mode=1)
        pyref = pr if "unavailable" in pr else {
            "value": out_px / pr["s_per_frame_best"] / 1e6, "unit": "Mpix/s", "cores": pr["threads"], "kind": "reference",
            "s_per_frame": pr["s_per_frame_best"],
            "sample": "1 frame of %dx%d through the unmodified myQL + models (sim.py:82-114,205 path, dumps off), staged in oracle/%s" % (H, W, pr["reference_root"])}
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.workload, case, B, H, W, r, cin, cout, args.outputs, max(args.gpus, 1),
                                  B * cin * H * W * 4 / 1e6),
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": c_oracle.threads(), "kind": "port",
                         "sample": "%d frame(s) of %dx%d per step, C restatement of the reference path "
                                   "(oracle/sesr_oracle.c, OpenMP, %d threads set explicitly)" % (frames, H, W, c_oracle.threads())},
        "python_reference": pyref,
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sesr_x2_int8_sim_1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU per step (0 = workload default)")
    ap.add_argument("--kernel-impl", default="auto", choices=["auto", "generic"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--calib-frames", type=int, default=256,
                    help="frames of 80x960 for the calibration timing (BASELINE config 2); 0 = skip")
    ap.add_argument("--no-variants", action="store_true", help="skip the y-only / q5-only timings of the default run")
    ap.add_argument("--outputs", default="y", choices=["both", "y", "q5"],
                    help="what the last layer materialises: fp32 pixel-shuffled y (what model(inps) returns, sim.py:205; default), "
                         "int8 q5 (input.5.pt, written upstream only with INPUT_W_FLG), or both")
    ap.add_argument("--no-python-ref", action="store_true", help="skip timing the real Python reference (oracle/_ref)")
    ap.add_argument("--no-workloads", action="store_true", help="skip the other BASELINE shapes / batch sweep")
    args = ap.parse_args()
    case, H, W, defB = WORKLOADS[args.workload]
    B = args.batch or defB
    g = load_case(case)
    r = g.r
    cin, cout = g.wf[0].shape[1], g.wf[4].shape[0]
    out_px_frame = H * W * r * r

    if args.impl == "reference":
        run_reference(args, case, H, W, r, B, cin, cout)
        return

    import torch
    import torch.distributed as dist
    from sesr_pytorch_quantize_b200 import Int8Net, compile_network, IMPL_AUTO, IMPL_GENERIC, load_library

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sesrq hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa(torch, local)          # before the pinned buffers are allocated (first touch)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load_library()

    cnet = compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], g.scales, g.zeros, r)
    net = Int8Net(cnet, device=dev, impl=IMPL_GENERIC if args.kernel_impl == "generic" else IMPL_AUTO)
    impls = [net.layer_info(i) for i in range(5)]

    gen = torch.Generator().manual_seed(4321 + rank)
    x_host = torch.rand(B, cin, H, W, generator=gen).pin_memory()
    x = x_host.to(dev, non_blocking=True)
    qs, ys = net.out_shapes(B, H, W)
    want_q5, want_y = args.outputs in ("both", "q5"), args.outputs in ("both", "y")
    q5 = torch.empty(qs, dtype=torch.int8, device=dev) if want_q5 else None
    if want_q5 and os.environ.get("SESRQ_BENCH_Q5_OFFSET"):      # address-aliasing experiments: shift the q5 base
        off = int(os.environ["SESRQ_BENCH_Q5_OFFSET"])
        n = 1
        for d in qs:
            n *= d
        _q5_backing = torch.empty(n + off + 4096, dtype=torch.int8, device=dev)
        q5 = _q5_backing[off:off + n].view(qs)
    y = torch.empty(ys, dtype=torch.float32, device=dev) if want_y else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        net(x, want_q5=want_q5, want_y=want_y, q5=q5, y=y)
    barrier()

    # ---- timed region: K steps, inputs resident in HBM, per-kernel events on the launching stream ----
    net.profile(True)
    net.profile_read()
    proc, cpath = sample_clocks_start(local)
    idle_rows = sample_clocks_wait_ready(proc, cpath)
    nvml = NvmlClockSampler(torch, local)
    launches0 = lib.sesrq_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    nvml.start()
    e0.record()
    for _ in range(args.steps):
        net(x, want_q5=want_q5, want_y=want_y, q5=q5, y=y)
    e1.record()
    barrier()
    clocks_timed = nvml.stop()
    ms_total = e0.elapsed_time(e1)
    launches = lib.sesrq_launch_count() - launches0
    layer_ms, layer_launches = net.profile_read()
    net.profile(False)
    # The K timed steps last ~10 ms, less than one nvidia-smi period (100 ms): keep the SAME step running (untimed, after the
    # timed region has been closed) until the sampler has seen the GPU under this load a few times, at most ~1.5 s.
    t_load = time.time()
    sus_steps = 0
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    while clock_rows(cpath) < idle_rows + 1 + 6 and time.time() - t_load < 1.5 and proc is not None:
        for _ in range(args.steps):
            net(x, want_q5=want_q5, want_y=want_y, q5=q5, y=y)
        sus_steps += args.steps
        torch.cuda.synchronize(dev)
    u1.record()
    torch.cuda.synchronize(dev)
    sus_ms = u0.elapsed_time(u1) / sus_steps if sus_steps else None
    smi = sample_clocks_stop(proc, cpath, skip_rows=idle_rows + 1)          # +1: the row that straddles the start of the load
    smi["sampled_over"] = "the same step repeated back to back for up to 1.5 s after the timed region (nvidia-smi -lms 100; idle rows dropped)"
    smi["ms_per_step"] = sus_ms                                             # the step's duration under that sustained load (includes a host sync every K steps)
    smi["steps"] = sus_steps
    if clocks_timed["samples"] > 0:
        clocks = clocks_timed                                               # what the contract asks for: DURING the timed region
        clocks["sustained"] = smi                                           # and what a second of the same load does to the clocks
    else:
        clocks = smi
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B * out_px_frame / (ms_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (largest share of the step), live numbers ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    lbytes = layer_bytes_per_px(cin, cout, want_q5, want_y)
    lr_px_launch = [(B * H * W * args.steps) / max(layer_launches[i], 1) for i in range(5)]
    per_layer = []
    fused12 = bool(net.fused) and layer_launches[2] == 0
    for i in range(5):
        avg_ms = layer_ms[i] / max(layer_launches[i], 1)
        # layers 1 + 2 in one launch (csrc/sesrq_chain.cu): the kernel is credited with both layers' ALGORITHMIC bytes
        # (fusion does not change them), so its figure may exceed the HBM peak; what it really moves
        # (the intermediate codes stay in shared memory) is listed beside it
        alg = lbytes[i] + (lbytes[2] if (fused12 and i == 1) else 0)
        gbs = alg * lr_px_launch[i] / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        rec = {"layer": i, "impl": "umma" if impls[i][0] == 2 else "generic", "pe_exact": impls[i][1],
               "avg_ms": avg_ms, "launches": layer_launches[i], "alg_bytes_per_lr_px": alg,
               "achieved_gbs": gbs, "share": layer_ms[i] / max(sum(layer_ms), 1e-9)}
        if fused12 and i == 1:
            rec.update({"fused_layers": [1, 2], "hbm_bytes_per_lr_px": 32,
                        "hbm_gbs": 32 * lr_px_launch[i] / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0})
        if fused12 and i == 2:
            rec["fused_into"] = 1
        per_layer.append(rec)
    dom = max(range(5), key=lambda i: layer_ms[i])
    traffic = None
    if not args.batch:                                     # the ncu capture was taken on the default batch
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload + ":" + args.outputs, {}).get("layer%d" % dom)
        except (OSError, ValueError):
            pass
    roofline = {"bound": "hbm", "kernel": "layer %d (%s)" % (dom, per_layer[dom]["impl"]),
                "achieved": per_layer[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": per_layer[dom]["achieved_gbs"] / peak, "traffic": traffic, "peak_source": peak_src}
    stack_gbs = sum(lbytes) * B * H * W / (ms_step * 1e-3) / 1e9
    total_macs = sum(L.ksize * L.ksize * L.cin * L.cout for L in cnet.layers)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside) ----
    e2e = None
    if not args.no_e2e:
        q5_h = torch.empty(qs, dtype=torch.int8).pin_memory() if want_q5 else None
        y_h = torch.empty(ys, dtype=torch.float32).pin_memory() if want_y else None
        net.forward_host(x_host, q5_h, y_h)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.e2e_steps):
            net.forward_host(x_host, q5_h, y_h)
        s1.record()
        barrier()
        te = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te.item()) / args.e2e_steps
        h2d_b = x_host.numel() * 4
        d2h_b = (q5_h.numel() if want_q5 else 0) + (y_h.numel() * 4 if want_y else 0)
        tcp = torch.tensor([copy_ceiling(torch, dev, h2d_b, d2h_b, args.e2e_steps, barrier)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tcp, op=dist.ReduceOp.MAX)
        ceil_ms = float(tcp.item())
        e2e = {"value": world * B * out_px_frame / (e_ms * 1e-3) / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
               "ms_per_step": e_ms, "api": "sesrq_net_forward_i8_host (Int8Net.forward_host), pinned host buffers",
               "copy_ceiling": {"ms_per_step": ceil_ms, "value": world * B * out_px_frame / (ceil_ms * 1e-3) / 1e6,
                                "gbs_per_rank": (h2d_b + d2h_b) / (ceil_ms * 1e-3) / 1e9,
                                "what": "the same H2D + D2H volumes as two concurrent pinned copies, no kernels, all %d rank(s) at once" % world},
               "frac_of_copy_ceiling": ceil_ms / e_ms, "affinity": affinity}

    # ---- calibration (BASELINE config 2): fake-quant forward + running min/max over the rank's shard of the frames,
    # one all_reduce(MAX); then the KL calibrator's second pass (2048-bin histograms on device), one all_reduce(SUM)
    # and the on-device threshold search.  Device-timed, max over ranks. ----
    calib = None
    if args.calib_frames > 0:
        from sesr_pytorch_quantize_b200.fq import FakeQuantNet
        from sesr_pytorch_quantize_b200.calib import MinMaxCalibrator, KLCalibrator, shard_range
        ch, cw = 80, 960
        lo, hi = shard_range(args.calib_frames, rank, world)
        # ONE global calibration set (seed 1234, the goldens' recipe), every rank takes its contiguous shard: the merged
        # statistics -- and so the thresholds -- are the same for every world size
        cgen = torch.Generator().manual_seed(1234)
        frames = torch.rand(args.calib_frames, cin, ch, cw, generator=cgen)[lo:hi].to(dev)
        fq = FakeQuantNet([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], r,
                          residual=True, observe_out=(r == 1), device=dev)
        fq(frames)                                                  # warm-up at the timed shape (module load, workspace)
        fq.reset_observers()
        barrier()
        c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        c0.record()
        mmc = MinMaxCalibrator(fq)
        mmc.observe(frames)
        mm_scales, mm_zeros, _, _ = mmc.finish()                    # includes the all_reduce and the D2H of 12 floats
        c1.record()
        fq.reset_observers()
        klc = KLCalibrator(fq)
        klc.observe_pass1(frames)
        klc.begin_pass2()
        klc.observe_pass2(frames)
        kl_scales, kl_zeros, kl_idx = klc.finish()
        c2.record()
        barrier()
        tc = torch.tensor([c0.elapsed_time(c1), c1.elapsed_time(c2)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        import hashlib
        thr_hash = hashlib.sha256(repr(([float(v) for v in mm_scales], [int(v) for v in mm_zeros])).encode()).hexdigest()[:16]
        calib = {"frames": args.calib_frames, "shape": [cin, ch, cw], "minmax_s": float(tc[0]) * 1e-3, "kl_s": float(tc[1]) * 1e-3,
                 "minmax_frames_per_s": args.calib_frames / (float(tc[0]) * 1e-3), "kl_bin_counts": kl_idx,
                 "minmax_threshold_hash": thr_hash, "minmax_zeros": [int(v) for v in mm_zeros],
                 "frame_set": "global: torch.rand seed 1234, %d frames, contiguous shards" % args.calib_frames,
                 "collectives": "all_reduce(MAX) over 12 floats; all_reduce(SUM) over int64 [6,2048]" if world > 1 else "none (1 rank)"}
        # roofline of the min/max pass.  Algorithmic bytes: every activation of the fp32 fake-quant graph crosses HBM once
        # per direction as fp32, 4 * (Cin + 4 * 16 + Cout) B/px (the reference's tensors; 316 B/px for SESR x2).  What the
        # batched pass really moves: each hidden activation is written as fp32 by the conv epilogue (its per-frame range
        # is only known when the whole frame is done), read back and written as operand bytes by the quantise pass, and
        # read by the next conv: 4*Cin + 16 + 16 for layer 0, 64 + 64 + 16 + 16 per hidden layer, 64 + 64 + 4*Cout at the end
        cal_px = (hi - lo) * ch * cw
        alg_b = 4 * (cin + 4 * 16 + cout)
        real_b = (4 * cin + 32) + 4 * 160 + (128 + 4 * cout)
        mm_s = float(tc[0]) * 1e-3
        calib["roofline"] = {"bound": "hbm", "what": "min/max calibration pass (fake-quant forward of all frames)",
                             "alg_bytes_per_lr_px": alg_b, "moved_bytes_per_lr_px": real_b,
                             "achieved": alg_b * cal_px / mm_s / 1e9, "moved_gbs": real_b * cal_px / mm_s / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": alg_b * cal_px / mm_s / 1e9 / peak,
                             "frac_moved": real_b * cal_px / mm_s / 1e9 / peak}
        fq.close()

    # ---- the same step for the other output selections (y = what model(inps) returns; q5 = input.5.pt, SURVEY 8d's
    # base figure of 168 B/px for x2; both = 216 B/px) ----
    def time_net(net_, x_, wq, wy, q5_, y_, steps):
        for _ in range(3):
            net_(x_, want_q5=wq, want_y=wy, q5=q5_, y=y_)
        torch.cuda.synchronize(dev)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(steps):
            net_(x_, want_q5=wq, want_y=wy, q5=q5_, y=y_)
        s1.record()
        torch.cuda.synchronize(dev)
        return s0.elapsed_time(s1) / steps

    single = None
    if world == 1 and not args.no_variants:
        single = {}
        q5v = q5 if q5 is not None else torch.empty(qs, dtype=torch.int8, device=dev)
        yv = y if y is not None else torch.empty(ys, dtype=torch.float32, device=dev)
        for name, wq, wy in (("y_only", False, True), ("q5_only", True, False), ("both", True, True)):
            ms = time_net(net, x, wq, wy, q5v, yv, args.steps)
            nbytes = sum(layer_bytes_per_px(cin, cout, wq, wy))
            single[name] = {"value": B * out_px_frame / (ms * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": ms,
                            "alg_bytes_per_lr_px": nbytes,
                            "frac_of_hbm_peak": nbytes * B * H * W / (ms * 1e-3) / 1e9 / peak}
        del q5v, yv

    # ---- the other BASELINE.json shapes and the batch sweep of config 5, device-timed like `value` (same outputs) ----
    workloads = None
    if world == 1 and not args.no_workloads and not args.batch and args.workload == "sesr_x2_int8_sim_1080p":
        workloads = []
        del x, q5, y
        torch.cuda.empty_cache()
        plan = [(wl, WORKLOADS[wl][3]) for wl in ("sesr_x4_int8_sim_540x960", "nrdm3_int8_sim_80x960", "sesr_x2_int8_sim_80x960",
                                                  "sesr_x4_int8_sim_80x960")]
        plan += [("sesr_x2_int8_sim_1080p", b) for b in (1, 8, 64, 512)] + [("sesr_x2_int8_sim_80x960", 1), ("nrdm3_int8_sim_80x960", 1),
                                                                             ("nrdm3_int8_sim_80x960", 64)]
        for wl, wb in plan:
            wcase, wH, wW, _ = WORKLOADS[wl]
            wg = load_case(wcase)
            wcin, wcout, wr = wg.wf[0].shape[1], wg.wf[4].shape[0], wg.r
            need = wb * wH * wW * (4 * wcin + (wcout if want_q5 else 0) + (4 * wcout if want_y else 0) + 64) * 1.1
            free_b, _tot = torch.cuda.mem_get_info(dev)
            if need > free_b:
                workloads.append({"workload": wl, "frames_per_step": wb, "skipped": "needs %.0f GB" % (need / 1e9)})
                continue
            wnet = Int8Net(compile_network([torch.from_numpy(w) for w in wg.wf], [torch.from_numpy(b) for b in wg.bf],
                                           wg.scales, wg.zeros, wr), device=dev)
            wx = torch.rand(wb, wcin, wH, wW, generator=torch.Generator().manual_seed(4321)).to(dev) if wb * wH * wW * wcin < 5e8 \
                else torch.rand(wb, wcin, wH, wW, device=dev)
            wqs, wys = wnet.out_shapes(wb, wH, wW)
            wq5 = torch.empty(wqs, dtype=torch.int8, device=dev) if want_q5 else None
            wy_ = torch.empty(wys, dtype=torch.float32, device=dev) if want_y else None
            ms = time_net(wnet, wx, want_q5, want_y, wq5, wy_, max(3, min(args.steps, 10)))
            nbytes = sum(layer_bytes_per_px(wcin, wcout, want_q5, want_y))
            workloads.append({"workload": wl, "frames_per_step": wb, "H": wH, "W": wW, "ms_per_step": ms, "us_per_frame": ms * 1e3 / wb,
                              "value": wb * wH * wW * wr * wr / (ms * 1e-3) / 1e6, "unit": "Mpix/s (output)",
                              "alg_bytes_per_lr_px": nbytes, "frac_of_hbm_peak": nbytes * wb * wH * wW / (ms * 1e-3) / 1e9 / peak,
                              "input_mb": wx.numel() * 4 / 1e6})
            wnet.close()
            del wx, wq5, wy_, wnet
            torch.cuda.empty_cache()

    # ---- BASELINE config 1: one 80x960 frame, launch-latency bound -> eager launches vs one CUDA-graph replay ----
    small = None
    if world == 1 and not args.no_workloads and not args.batch:
        sg = load_case("x2_fixture")
        snet = Int8Net(compile_network([torch.from_numpy(w) for w in sg.wf], [torch.from_numpy(b) for b in sg.bf], sg.scales,
                                       sg.zeros, sg.r), device=dev)
        sx = torch.from_numpy(sg.x).to(dev)
        sq, sy = snet.out_shapes(1, 80, 960)
        sy_t = torch.empty(sy, dtype=torch.float32, device=dev)
        eager_ms = time_net(snet, sx, False, True, None, sy_t, 200)
        gnet = snet.capture(sx, want_q5=False, want_y=True)
        for _ in range(5):
            gnet()
        torch.cuda.synchronize(dev)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(200):
            gnet()
        g1.record()
        torch.cuda.synchronize(dev)
        graph_ms = g0.elapsed_time(g1) / 200
        same = bool(torch.equal(gnet.y, sy_t))
        small = {"workload": "sesr_x2_int8_sim_80x960, 1 frame (BASELINE config 1), y only", "eager_us_per_frame": eager_ms * 1e3,
                 "cuda_graph_us_per_frame": graph_ms * 1e3, "graph_equals_eager": same,
                 "note": "back-to-back frames on one stream; launches overlap through programmatic dependent launch; 40 tiles on 148 SMs"}
        snet.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        secs, thr = cpu_reference_time(case, H, W, 1, 2)
        cpu = {"value": out_px_frame / secs / 1e6, "unit": "Mpix/s", "cores": thr, "kind": "port",
               "sample": "1 frame of %dx%d (best of 2), C restatement of the reference path with OpenMP (%d threads set explicitly)" % (H, W, thr),
               "s_per_frame": secs}
        if not args.no_python_ref:
            pr = python_reference_time(case, H, W, frames=1, mode=1)
            cpu["python_reference"] = pr if "unavailable" in pr else {
                "value": out_px_frame / pr["s_per_frame_best"] / 1e6, "unit": "Mpix/s", "cores": pr["threads"], "kind": "reference",
                "s_per_frame": pr["s_per_frame_best"],
                "sample": "1 frame of %dx%d through the unmodified myQL + models (sim.py path, dumps off) from oracle/%s" % (H, W, pr["reference_root"])}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args.workload, case, B, H, W, r, cin, cout, args.outputs, world, B * cin * H * W * 4 / 1e6),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "calibration": calib, "output_variants": single, "workloads": workloads, "small_frame": small,
            "stack": {"alg_bytes_per_lr_px": sum(lbytes), "achieved_gbs": stack_gbs, "frac_of_hbm_peak": stack_gbs / peak,
                      "int8_tops": 2 * total_macs * B * H * W / (ms_step * 1e-3) / 1e12, "us_per_frame": ms_step * 1e3 / B},
            "layers": per_layer,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
