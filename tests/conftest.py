import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "sesr_pytorch_quantize_b200")
# PKG is the drop-in import root for the reference's scripts: `define`, `myQL`, `models`
for p in (ROOT, os.path.join(ROOT, "oracle"), PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

ALL_CASES = ["x4_fixture", "nrdm_fixture", "x2_fixture", "x4_small", "nrdm_small", "x2_small",
             "synth_clamp", "synth_zero", "synth_neg", "nr_small", "dm_small"]
SMALL_CASES = [c for c in ALL_CASES if not c.endswith("_fixture")]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


from sesr_pytorch_quantize_b200.fixtures import Golden, load_golden  # noqa: E402,F401


@pytest.fixture(params=ALL_CASES)
def golden(request):
    return load_golden(request.param)


@pytest.fixture(params=SMALL_CASES)
def golden_small(request):
    return load_golden(request.param)


# ---- the reference's driver blocks (sim.py:82-114 / test.py:79-106) on the myQL mirror -----------------
MODEL_OF_MFLAG = {3: ("nrdm_3_sim", "nrdm_3", "nr"), 5: ("sesr_sim", "sesr", "sesr"), 6: ("sesr_arch_sim", "sesr_arch", "sesr")}


# the denoise-only / demosaic-only checkpoints: float variants are models.nr.nr / models.dm.dm (test.py:29-36), the
# integer-sim variant is the NRDM one (same architecture without the AddOp; upstream has no nr_sim / dm_sim)
MODEL_OF_CASE = {"nr_small": (("nrdm_3_sim", "nr"), ("nr", "nr")), "dm_small": (("nrdm_3_sim", "nr"), ("dm", "dm"))}


def collapsed_model(g, sim, device="cpu"):
    """Model of the golden case's family, collapsed, carrying the golden's collapsed float weights."""
    import importlib
    import torch
    if g.name in MODEL_OF_CASE:
        mod_name, cls = MODEL_OF_CASE[g.name][0 if sim else 1]
    else:
        sim_mod, float_mod, cls = MODEL_OF_MFLAG[int(g["mflag"])]
        mod_name = sim_mod if sim else float_mod
    mod = importlib.import_module("models." + mod_name)
    model = getattr(mod, cls)()
    model.collapse()
    convs = [model.conv_first.conv_expand] + [b.conv_expand for b in model.residual_block] + [model.conv_last.conv_expand]
    with torch.no_grad():
        for c, w, b in zip(convs, g.wf, g.bf):
            c.weight.copy_(torch.from_numpy(w))
            c.bias.copy_(torch.from_numpy(b))
    return model.to(device)


def quantize_graph(model, qmode, pe_add_bit=None):
    """The four insert passes exactly as sim.py:82-114 (qmode 1) / test.py:79-106 (qmode 0) call them."""
    import torch
    import define as d
    from myQL.quan_func import (quantize_model_weight, quantize_asymmetrical_by_tensor, reshape_input_for_hardware_pe,
                                PEs_and_bias_adder, requan_conv2d_output)
    from myQL.quan_classes import NodeInsertMapping, FunctionPackage, NodeInsertMappingElement
    from myQL.graph_modify import insert_before, insert_bias_bypass, insert_after
    model = quantize_model_weight(model, d.QUAN_BIT, qmode)
    mapping = NodeInsertMapping()
    quan_fp = FunctionPackage(quantize_asymmetrical_by_tensor, {'width': d.QUAN_BIT, 'exe_mode': qmode})
    mapping.add_config(NodeInsertMappingElement(torch.nn.Conv2d, quan_fp))
    if qmode == 0:
        mapping.add_config(NodeInsertMappingElement(torch.nn.PixelShuffle, quan_fp))
    model = insert_before(model_input=model, insert_mapping=mapping, has_func_id=True)
    rm = NodeInsertMapping()
    rm.add_config(NodeInsertMappingElement(torch.nn.Conv2d, FunctionPackage(reshape_input_for_hardware_pe, {'pe_num': d.PE})))
    model = insert_before(model_input=model, insert_mapping=rm)
    if qmode == 1:
        qm = NodeInsertMapping()
        qm.add_config(NodeInsertMappingElement(torch.nn.Conv2d, FunctionPackage(requan_conv2d_output, {'exe_mode': qmode})))
        model = insert_after(model_input=model, insert_mapping=qm)
    bm = NodeInsertMapping()
    bm.add_config(NodeInsertMappingElement(torch.nn.Conv2d, FunctionPackage(
        PEs_and_bias_adder, {'pe_add_width': pe_add_bit or d.PE_ADD_BIT, 'pe_acc_width': d.PE_ACC_BIT,
                             'bias_width': d.BIAS_BIT, 'pe_num': d.PE, 'exe_mode': qmode})))
    return insert_bias_bypass(model_input=model, insert_mapping=bm)


@pytest.fixture
def fresh_state(tmp_path, monkeypatch):
    """Scratch cwd (the reference's state files are cwd-relative) and an empty myQL.state registry."""
    monkeypatch.chdir(tmp_path)
    from myQL.state import state
    state.reset()
    yield state
    state.reset()
