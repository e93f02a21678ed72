"""Host-side logic of the myQL mirror (no GPU): import surface of the reference, the fx insert
passes produce the reference's node chain, fusion classification, one-time weight / bias
quantisation against the reference's dumps, the file-compat state registry, error behaviour."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, collapsed_model, quantize_graph, SMALL_CASES


def test_import_surface_matches_reference():
    import define
    for name in ("MFLAG", "TEST_RAW_ADD_NOISE", "PE", "QUAN_BIT", "BIAS_BIT", "PE_ACC_BIT", "PE_ADD_BIT", "REQUAN_BIT",
                 "REQUAN_N_MAX", "WEIGHT_W_FLG", "INPUT_W_FLG", "BIAS_W_FLG", "BIAS_QUAN_W_FLG", "OUTPUT_PE_W_FLG",
                 "OUTPUT_PE_ADD_W_FLG", "REQUAN_FACTOR_W_FLG", "WEIGHT_W_HIST_PNG", "INPUT_W_HIST_PNG"):
        assert hasattr(define, name), name
    assert (define.PE, define.QUAN_BIT, define.BIAS_BIT, define.PE_ACC_BIT, define.PE_ADD_BIT, define.REQUAN_BIT,
            define.REQUAN_N_MAX) == (4, 8, 16, 18, 20, 16, 32)
    from myQL import quan_func, quan_classes, graph_modify, quan_modules
    for name in ("quantize_model_weight", "quantize_asymmetrical_by_tensor", "reshape_input_for_hardware_pe",
                 "PEs_and_bias_adder", "requan_conv2d_output", "quantize_symmetrical_by_tensor",
                 "quantize_bias_with_scale", "reshape_ouput_for_hardware_pe", "quan_layer_between_const",
                 "float_to_hex", "remove_suffix"):
        assert callable(getattr(quan_func, name)), name
    for name in ("NodeInsertMapping", "FunctionPackage", "NodeInsertMappingElement", "NodeInsertConfig"):
        assert hasattr(quan_classes, name)
    for name in ("insert_before", "insert_bias_bypass", "insert_after", "get_insert_config"):
        assert callable(getattr(graph_modify, name))
    for name in ("myMinMaxObserver", "conv_forward_naive", "sesr_forward_sim"):
        assert hasattr(quan_modules, name)
    from models import sesr_sim, sesr, sesr_arch_sim, sesr_arch, nrdm_3_sim, nrdm_3, nr, dm, nrdm_6
    assert len(nrdm_6.nr().residual_block) == 6
    for mod, cls, cin, cout in ((sesr_sim, "sesr", 1, 16), (sesr, "sesr", 1, 16), (sesr_arch_sim, "sesr", 3, 12),
                                (sesr_arch, "sesr", 3, 12), (nrdm_3_sim, "nr", 3, 3), (nrdm_3, "nr", 3, 3),
                                (nr, "nr", 3, 3), (dm, "dm", 3, 3), (nrdm_6, "nr", 3, 3)):
        m = getattr(mod, cls)()
        assert m.conv_first.conv_expand.in_channels == cin
        m.before_quantization()
        assert m.conv_last.conv_expand.out_channels == cout and m.conv_last.conv_expand.kernel_size == (5, 5)


def test_collapse_equals_uncollapsed_forward():
    from models import sesr_arch
    torch.manual_seed(0)
    m = sesr_arch.sesr().eval()
    x = torch.rand(1, 3, 12, 17)
    with torch.no_grad():
        y0 = m(x)
        m.collapse()
        y1 = m(x)
    assert y1.shape == (1, 3, 24, 34)
    assert torch.allclose(y0, y1, atol=2e-5)


@pytest.mark.parametrize("qmode", [1, 0])
@pytest.mark.parametrize("case", ["x4_small", "x2_small", "nrdm_small"])
def test_insert_passes_build_reference_chain(case, qmode, fresh_state):
    from myQL import quan_func as qf
    g = load_golden(case)
    gm = quantize_graph(collapsed_model(g, sim=(qmode == 1)), qmode)
    assert isinstance(gm, torch.fx.GraphModule)
    names = []
    for n in gm.graph.nodes:
        if n.op == "call_function":
            names.append(getattr(n.target, "__name__", str(n.target)))
        elif n.op == "call_module":
            names.append(type(dict(gm.named_modules())[n.target]).__name__)
    per_layer = ["quantize_asymmetrical_by_tensor", "reshape_input_for_hardware_pe", "Conv2d", "PEs_and_bias_adder"]
    if qmode == 1:
        per_layer.append("requan_conv2d_output")
    want = []
    for i in range(5):
        want += per_layer + ["ReLU" if i < 4 else "Identity"]
        if qmode == 0 and i == 3:
            want.append("add")                                  # AddOp traced through
    if int(g["mflag"]) != 3:
        want += (["quantize_asymmetrical_by_tensor"] if qmode == 0 else []) + ["PixelShuffle"]
    assert names == want
    # kwargs carry the running func_id and the bypassed bias list; the conv biases are zeroed
    quants = [n for n in gm.graph.nodes if n.op == "call_function" and n.target is qf.quantize_asymmetrical_by_tensor]
    assert [n.kwargs["func_id"] for n in quants] == list(range(len(quants)))
    merges = [n for n in gm.graph.nodes if n.op == "call_function" and n.target is qf.PEs_and_bias_adder]
    for i, n in enumerate(merges):
        assert n.kwargs["func_id"] == i and n.kwargs["exe_mode"] == qmode
        assert np.array_equal(np.asarray(n.kwargs["bias"], dtype=np.float32), g.bf[i])
    for m in gm.modules():
        if type(m) is torch.nn.Conv2d:
            assert float(m.bias.detach().abs().max()) == 0.0
    assert gm.fusion_plan() == ("int8" if qmode == 1 else "fq")


def test_partial_graph_is_not_fused(fresh_state):
    import define as d
    from myQL.quan_func import quantize_model_weight, quantize_asymmetrical_by_tensor
    from myQL.quan_classes import NodeInsertMapping, FunctionPackage, NodeInsertMappingElement
    from myQL.graph_modify import insert_before
    g = load_golden("x4_small")
    model = quantize_model_weight(collapsed_model(g, sim=True), d.QUAN_BIT, 1)
    mp = NodeInsertMapping()
    mp.add_config(NodeInsertMappingElement(torch.nn.Conv2d, FunctionPackage(quantize_asymmetrical_by_tensor,
                                                                         {'width': 8, 'exe_mode': 1})))
    gm = insert_before(model_input=model, insert_mapping=mp, has_func_id=True)
    assert gm.fusion_plan() == "graph"


@pytest.mark.parametrize("qmode", [1, 0])
def test_rewired_five_conv_graph_is_not_fused(qmode, fresh_state):
    """Five convs with all the right node kinds but a different dataflow (layer 2 reads layer 0's output instead of
    layer 1's) must not be executed as the hard-wired SESR chain: the analyser checks the wiring, not the node counts."""
    from myQL import quan_func as qf
    g = load_golden("x4_small")
    gm = quantize_graph(collapsed_model(g, sim=(qmode == 1)), qmode)
    assert gm.fusion_plan() == ("int8" if qmode == 1 else "fq")
    quants = [n for n in gm.graph.nodes if n.op == "call_function" and n.target is qf.quantize_asymmetrical_by_tensor]
    q1, q2 = quants[1], quants[2]
    q2.args = (q1.args[0],) + tuple(q2.args[1:])          # layer 2 now consumes what layer 1 consumes
    gm.graph.lint()
    gm.recompile()
    gm._sesrq_plan = None
    assert gm.fusion_plan() == "graph"


@pytest.mark.parametrize("case", SMALL_CASES)
def test_weight_quantisation_equals_reference_dumps(case, fresh_state):
    import define as d
    from myQL.quan_func import quantize_model_weight
    g = load_golden(case)
    for qmode in (1, 0):
        qm = quantize_model_weight(collapsed_model(g, sim=True), d.QUAN_BIT, qmode)
        convs = [m for m in qm.modules() if type(m) is torch.nn.Conv2d]
        for i, c in enumerate(convs):
            assert np.array_equal(fresh_state.weight_codes[i].numpy(), g["W%d" % i])
            assert fresh_state.weight_scale[i] == float(g["sw%d" % i])
            want = g["W%d" % i].astype(np.float32)
            if qmode == 0:
                want = (torch.from_numpy(want) * float(g["sw%d" % i])).numpy()
            assert np.array_equal(c.weight.detach().numpy(), want)


def test_compile_from_state_equals_reference_constants(fresh_state):
    from sesr_pytorch_quantize_b200.compile import compile_from_codes
    for case in SMALL_CASES:
        g = load_golden(case)
        cn = compile_from_codes([torch.from_numpy(g["W%d" % i]) for i in range(5)], [float(g["sw%d" % i]) for i in range(5)],
                                [torch.from_numpy(b) for b in g.bf], g.scales, g.zeros, g.r)
        assert [L.m16 for L in cn.layers] + [cn.res_m16] == [int(v) for v in g["M"]]
        assert [L.shift_n for L in cn.layers] + [cn.res_shift_n] == [int(v) for v in g["n"]]
        for i, L in enumerate(cn.layers):
            assert L.add_const == [int(v) for v in g["K%d" % i]]


def test_float_to_hex_and_weight_hex_image(fresh_state, monkeypatch):
    from myQL import quan_func as qf
    assert qf.float_to_hex(-1, 8) == "ff" and qf.float_to_hex(127, 8) == "7f" and qf.float_to_hex(-128, 8) == "80"
    assert qf.float_to_hex(-1, 16) == "ffff" and qf.float_to_hex(5, 12) == "005" and qf.float_to_hex(-2, 20) == "ffffe"
    assert qf.remove_suffix("conv_first.conv_expand.weight") == "conv_first.conv_expand"
    monkeypatch.setattr(qf, "WEIGHT_W_FLG", True)
    w = torch.zeros(3, 1, 3, 3)
    w[0, 0, 0, 0], w[1, 0, 0, 0], w[2, 0, 2, 2] = 1.0, -1.0, 0.5
    codes = qf.quantize_symmetrical_by_tensor(w, 8, 1, func_id=0)
    assert codes[0, 0, 0, 0] == 127 and codes[1, 0, 0, 0] == -127 and codes[2, 0, 2, 2] == 64
    lines = open("output_txt/weight/conv.weight.0.txt").read().split("\n")
    # header = number of 16-weight words: oc, ic padded to 4 -> 4*4*3*3/16 = 9 (quan_func.py:82-111)
    assert lines[0] == "09"
    assert lines[1] == "7f810000" + "00000000" * 3                # tap (0,0): ic 0 -> oc0..3, then ic 1..3 (zero pad)
    assert lines[9] == "00004000" + "00000000" * 3                # tap (2,2)
    assert os.path.isfile("output_pt/weight/conv.weight.0.pt") and os.path.isfile("output_pt/weight/conv.weight.0.scale.pt")


def test_quan_layer_between_const_branches():
    from myQL.quan_func import quan_layer_between_const as qbc
    import sesr_oracle as orc
    for v in (0.0081, 0.5, 0.999, 1.0, 1.5, 3.999, 100.25, 2 ** -20, 2 ** -17 * 1.3, 1e-9):
        assert qbc(v, 16, 32) == orc.quan_layer_between_const(v)
        m, n = qbc(v, 16, 32)
        assert 0 <= m < 2 ** 16 and n <= 32
    with pytest.raises(AssertionError):
        qbc(0.5, 32, 32)


def test_state_registry_file_compat(fresh_state):
    st = fresh_state
    with pytest.raises(KeyError):
        st.scale(0)
    # thresholds written by a reference-style calibration tail are picked up from ./output_pt
    os.makedirs("output_pt/input")
    torch.save(0.0039, "output_pt/input/input.0.scale.pt")
    torch.save(-128, "output_pt/input/input.0.zero.pt")
    assert st.scale(0) == 0.0039 and st.zero(0) == -128
    v0 = st.version
    st.set_threshold(1, 0.5, -100)
    assert st.version > v0
    assert torch.load("output_pt/input/input.1.scale.pt") == 0.5 and torch.load("output_pt/input/input.1.zero.pt") == -100
    # running min/max observer: first value is stored, later ones only widen; a reset written into the files
    # (test.py:108-113) is honoured
    st.observe(2, -1.0, 2.0)
    st.observe(2, -0.5, 3.0)
    assert (st.run_min[2], st.run_max[2]) == (-1.0, 3.0)
    assert torch.load("output_pt/input/input.2.max_val.pt") == 3.0
    torch.save(-1000, "output_pt/input/input.2.max_val.pt")
    torch.save(1000, "output_pt/input/input.2.min_val.pt")
    st.observe(2, 0.25, 0.75)
    assert (st.run_min[2], st.run_max[2]) == (0.25, 0.75)


def test_calibration_tail_equals_reference(fresh_state):
    from sesr_pytorch_quantize_b200.compile import thresholds_from_minmax
    for case in ("x4_fixture", "x2_fixture", "nrdm_fixture", "synth_neg"):
        g = load_golden(case)
        scales, zeros = thresholds_from_minmax(g["mins"], g["maxs"])
        assert scales == g.scales and zeros == g.zeros


def test_ops_fail_loudly_without_cuda_tensors(fresh_state):
    from myQL import quan_func as qf
    from sesr_pytorch_quantize_b200 import SesrqError
    x = torch.zeros(1, 16, 4, 4)
    fresh_state.set_threshold(1, 1.0, -128, write_file=False)
    with pytest.raises(SesrqError):
        qf.quantize_asymmetrical_by_tensor(x, 8, 1, func_id=1)
    with pytest.raises(SesrqError):
        qf.reshape_input_for_hardware_pe(x, pe_num=4)
    with pytest.raises(SesrqError):
        qf.requan_conv2d_output(x, 1, 1)
    with pytest.raises(SesrqError):
        qf.quantize_asymmetrical_by_tensor(x.cuda() if torch.cuda.is_available() else x, 8, 0, func_id=1)
    with pytest.raises(AssertionError):
        qf.reshape_input_for_hardware_pe(torch.zeros(4, 4), pe_num=4)
    with pytest.raises(AssertionError):
        qf.quantize_bias_with_scale(torch.ones(4), 0.1, 16, 1, 0)      # "bias elements are all equal"
    with pytest.raises(AssertionError):
        qf.quantize_symmetrical_by_tensor(torch.zeros(2, 2, 3, 3), 8, 1, func_id=0)


def test_observer_module_has_no_cpu_fallback():
    from myQL.quan_modules import myMinMaxObserver
    from sesr_pytorch_quantize_b200 import SesrqError
    ob = myMinMaxObserver()
    assert float(ob.min_val) == float("inf") and float(ob.max_val) == float("-inf")
    assert ob(torch.empty(0)).numel() == 0                         # empty input is returned untouched (reference :27-28)
    with pytest.raises(SesrqError):
        ob(torch.tensor([1.0, -2.0, 3.0]))
    ob.reset_min_max_vals()
    assert float(ob.min_val) == float("inf")
