"""Run one of the reference's scripts UNMODIFIED (runpy, run_name "__main__") against this package's import root.

    python tests/run_ref_script.py <script.py> <mflag> <workdir> <result.npz>

The import path is [package root (define, myQL, models), tests/stubs (self_dataset, cv2, skimage), ...]: nothing of the
reference is importable.  Environment accommodations, none of which touches the script:
  * define.MFLAG is the reference's configuration knob (its users edit define.py); it is set here;
  * the scripts pin CUDA_VISIBLE_DEVICES=1 AFTER their imports (sim.py:27-28); CUDA is initialised beforehand so that
    a single-GPU box still sees its device 0;
  * torch >= 2.6 defaults torch.load to weights_only=True, which the reference's full-pickle checkpoint
    (x2sesr.pth.tar) predates: TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1;
  * cwd = <workdir>, which holds model_params/ and the rand_*_Input fixtures (cwd-relative paths in the scripts).
Afterwards the script's result tensors and the constants the run used are written to <result.npz>.
"""
import os
import runpy
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    script, mflag, workdir, result = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"
    sys.path[:0] = [os.path.join(ROOT, "sesr_pytorch_quantize_b200"), os.path.join(HERE, "stubs"), ROOT]
    import torch
    torch.cuda.init()
    torch.cuda.device_count()       # cached: the scripts' later CUDA_VISIBLE_DEVICES=1 must not hide a lone device 0
    import define
    define.MFLAG = mflag
    os.chdir(workdir)
    ns = runpy.run_path(script, run_name="__main__")
    from myQL.state import state
    out = {"mflag": mflag, "script": os.path.basename(script)}
    for k in ("gfake", "inps"):
        v = ns.get(k)
        if isinstance(v, torch.Tensor):
            out[k] = v.detach().float().cpu().numpy()
        elif isinstance(v, np.ndarray):
            out[k] = v
    for k in ("totalpsnr", "totalssim", "totalnum"):
        if k in ns:
            out[k] = float(ns[k])
    model = ns.get("model")
    out["plan"] = getattr(model, "fusion_plan", lambda: "n/a")()
    if hasattr(model, "graph"):
        for node in model.graph.nodes:
            if node.op == "call_function" and "bias" in node.kwargs:
                out["b%d" % node.kwargs["func_id"]] = np.asarray(node.kwargs["bias"], dtype=np.float32)
    for i, c in state.weight_codes.items():
        out["W%d" % i] = c.numpy()
        out["sw%d" % i] = state.weight_scale[i]
    for i in range(6):
        f = "output_pt/input/input.%d.scale.pt" % i
        if os.path.isfile(f):
            out["scale%d" % i] = float(torch.load(f))
            out["zero%d" % i] = int(torch.load("output_pt/input/input.%d.zero.pt" % i))
    np.savez(result, **out)


if __name__ == "__main__":
    main()
