"""Drive the UNMODIFIED reference (gui-yupeng/sesr-pytorch-quantize, mounted read-only at
/root/reference) on CPU to produce golden vectors.  TEST INFRASTRUCTURE ONLY.

This file restates the driver blocks of the reference scripts (sim.py:82-114,
test.py:79-113,186-217) as calls into the reference's own functions, with the
accommodations listed in SURVEY.md section 8c:
  * matplotlib is stubbed (quan_func.py:9 imports it, only used when *_HIST_PNG are set);
  * define.* dump flags are patched BEFORE myQL is imported (imported by value, quan_func.py:11-14);
  * state files are cwd-relative (./output_pt/**), so everything runs inside a scratch cwd;
  * checkpoints / fixtures are loaded with map_location="cpu".
It only works where /root/reference exists (the build container); nothing under tests/,
bench.py or smoke() imports it at run time -- they read the committed fixtures instead.
"""
import os
import sys
import types

REF = os.environ.get("SESRQ_REFERENCE_ROOT", "/root/reference")


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("cla", "hist", "savefig"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def import_reference(mflag, dumps=True):
    """Import reference modules with define patched. Returns a namespace of reference symbols."""
    assert os.path.isdir(REF), "reference tree not present: %s" % REF
    _stub_matplotlib()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import define  # reference define.py:1-36
    define.MFLAG = mflag
    define.WEIGHT_W_FLG = False          # hex txt of weights: slow, not needed
    define.INPUT_W_FLG = bool(dumps)
    define.BIAS_W_FLG = bool(dumps)
    define.BIAS_QUAN_W_FLG = bool(dumps)
    define.OUTPUT_PE_W_FLG = bool(dumps)
    define.OUTPUT_PE_ADD_W_FLG = bool(dumps)
    define.REQUAN_FACTOR_W_FLG = bool(dumps)
    define.WEIGHT_W_HIST_PNG = False
    define.INPUT_W_HIST_PNG = False
    import myQL.quan_func as qf
    import myQL.graph_modify as gm
    import myQL.quan_classes as qc
    qf.MFLAG = mflag
    from models import sesr_sim, sesr, sesr_arch_sim, sesr_arch, nrdm_3_sim, nrdm_3
    ns = types.SimpleNamespace(define=define, qf=qf, gm=gm, qc=qc, sesr_sim=sesr_sim, sesr=sesr,
                               sesr_arch_sim=sesr_arch_sim, sesr_arch=sesr_arch,
                               nrdm_3_sim=nrdm_3_sim, nrdm_3=nrdm_3)
    return ns


def build_float_model(ns, mflag, sim, state_dict=None):
    """sim.py:30-77 / test.py:29-76: construct, load checkpoint, collapse (CPU)."""
    import torch
    if mflag == 3:
        model = (ns.nrdm_3_sim if sim else ns.nrdm_3).nr()
        ck = "model_params/nrdm_3_raw_G.pth"
    elif mflag == 5:
        model = (ns.sesr_sim if sim else ns.sesr).sesr()
        ck = "model_params/x4sesr.pth"
    elif mflag == 6:
        model = (ns.sesr_arch_sim if sim else ns.sesr_arch).sesr()
        ck = "model_params/x2sesr.pth.tar"
    else:
        raise ValueError(mflag)
    model.train()
    if state_dict is None:
        sd = torch.load(os.path.join(REF, ck), map_location="cpu", weights_only=False)
        if mflag == 6:
            sd = sd["state_dict"]
    else:
        sd = state_dict
    model = model.float()
    model.load_state_dict(sd, strict=False)
    model.collapse()
    return model


def quantize_graph(ns, model, qmode):
    """sim.py:82-114 (qmode 1) / test.py:79-106 (qmode 0), verbatim call sequence."""
    import torch
    d, qf, gm, qc = ns.define, ns.qf, ns.gm, ns.qc
    model = qf.quantize_model_weight(model, d.QUAN_BIT, qmode)
    mapping = qc.NodeInsertMapping()
    quan_FP = qc.FunctionPackage(qf.quantize_asymmetrical_by_tensor, {'width': d.QUAN_BIT, 'exe_mode': qmode})
    mapping.add_config(qc.NodeInsertMappingElement(torch.nn.Conv2d, quan_FP))
    if qmode == 0:
        mapping.add_config(qc.NodeInsertMappingElement(torch.nn.PixelShuffle, quan_FP))
    model = gm.insert_before(model_input=model, insert_mapping=mapping, has_func_id=True)
    rm = qc.NodeInsertMapping()
    rm.add_config(qc.NodeInsertMappingElement(torch.nn.Conv2d, qc.FunctionPackage(qf.reshape_input_for_hardware_pe, {'pe_num': d.PE})))
    model = gm.insert_before(model_input=model, insert_mapping=rm)
    if qmode == 1:
        qm = qc.NodeInsertMapping()
        qm.add_config(qc.NodeInsertMappingElement(torch.nn.Conv2d, qc.FunctionPackage(qf.requan_conv2d_output, {'exe_mode': qmode})))
        model = gm.insert_after(model_input=model, insert_mapping=qm)
    bm = qc.NodeInsertMapping()
    bm.add_config(qc.NodeInsertMappingElement(torch.nn.Conv2d, qc.FunctionPackage(
        qf.PEs_and_bias_adder, {'pe_add_width': d.PE_ADD_BIT, 'pe_acc_width': d.PE_ACC_BIT,
                                'bias_width': d.BIAS_BIT, 'pe_num': d.PE, 'exe_mode': qmode})))
    model = gm.insert_bias_bypass(model_input=model, insert_mapping=bm)
    return model


def reset_minmax():
    """test.py:108-113."""
    import torch
    for i in range(6):
        if os.path.isfile("output_pt/input/input.{}.max_val.pt".format(i)):
            torch.save(-1000, "output_pt/input/input.{}.max_val.pt".format(i))
            torch.save(1000, "output_pt/input/input.{}.min_val.pt".format(i))


def calibration_tail(ns):
    """test.py:186-217: running min/max files -> scale/zero files. Returns (scales, zeros, mins, maxs)."""
    import torch
    QB = ns.define.QUAN_BIT
    scales, zeros, mins, maxs = [], [], [], []
    for i in range(6):
        inp_max = torch.load("output_pt/input/input.{}.max_val.pt".format(i))
        inp_min = torch.load("output_pt/input/input.{}.min_val.pt".format(i))
        mins.append(inp_min); maxs.append(inp_max)
        if i == 5:
            inp_min = 0                         # test.py:207
        quan_max = 2 ** (QB - 1) - 1
        quan_min = 0 - 2 ** (QB - 1)
        quan_scale = (inp_max - inp_min) / (quan_max - quan_min)
        quan_zero = quan_min - round(inp_min / quan_scale)
        torch.save(quan_scale, "output_pt/input/input.{}.scale.pt".format(i))
        torch.save(quan_zero, "output_pt/input/input.{}.zero.pt".format(i))
        scales.append(quan_scale); zeros.append(quan_zero)
    return scales, zeros, mins, maxs


def write_thresholds(scales, zeros):
    import torch
    os.makedirs("output_pt/input", exist_ok=True)
    for i, (s, z) in enumerate(zip(scales, zeros)):
        torch.save(float(s), "output_pt/input/input.{}.scale.pt".format(i))
        torch.save(int(z), "output_pt/input/input.{}.zero.pt".format(i))
