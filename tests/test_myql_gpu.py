"""The myQL mirror on the GPU, driven exactly as sim.py:82-114,205 / test.py:79-106,146 drive the reference:
  * the fused graph (one libsesrq network call) reproduces the reference's sim output bit for bit;
  * the un-fused per-node path (one CUDA op per reference callable, taken when dump flags are set) reproduces
    every tensor the reference dumps to ./output_pt (input.{i}.pt, pe_output{i}_{p}.pt, pe_add_output{i}.pt,
    requan factors) bit for bit;
  * the fake-quant graph matches the reference's mode-0 output within the stated tolerance and leaves the running
    min/max files the calibration tail (test.py:186-217) reads."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, collapsed_model, quantize_graph, SMALL_CASES

pytestmark = pytest.mark.gpu

DUMP_FLAGS = ("INPUT_W_FLG", "BIAS_W_FLG", "BIAS_QUAN_W_FLG", "OUTPUT_PE_W_FLG", "OUTPUT_PE_ADD_W_FLG", "REQUAN_FACTOR_W_FLG")


def set_thresholds(state, g):
    for i in range(6):
        state.set_threshold(i, g.scales[i], g.zeros[i])


@pytest.mark.parametrize("case", SMALL_CASES)
def test_fused_sim_graph_bit_exact(case, fresh_state):
    g = load_golden(case)
    set_thresholds(fresh_state, g)
    gm = quantize_graph(collapsed_model(g, sim=True, device="cuda"), 1, pe_add_bit=g.pe_add_bit)
    assert gm.fusion_plan() == "int8"
    y = gm(torch.from_numpy(g.x).cuda())
    assert np.array_equal(y.cpu().numpy(), g["yout"])
    # thresholds changed (e.g. a new calibration) -> the fused net is rebuilt, not reused
    fresh_state.set_threshold(5, g.scales[5] * 2, g.zeros[5])
    y2 = gm(torch.from_numpy(g.x).cuda())
    assert not np.array_equal(y2.cpu().numpy(), g["yout"])


@pytest.mark.parametrize("case", SMALL_CASES)
def test_per_node_graph_reproduces_reference_dumps(case, fresh_state, monkeypatch, capsys):
    import define
    from myQL import quan_func as qf
    for f in DUMP_FLAGS:
        monkeypatch.setattr(define, f, True)
        monkeypatch.setattr(qf, f, True)
    g = load_golden(case)
    set_thresholds(fresh_state, g)
    gm = quantize_graph(collapsed_model(g, sim=True, device="cuda"), 1, pe_add_bit=g.pe_add_bit)
    assert gm.fusion_plan() == "graph"
    y = gm(torch.from_numpy(g.x).cuda())
    assert np.array_equal(y.cpu().numpy(), g["yout"])

    def ld(p):
        v = torch.load(p, map_location="cpu")
        return v.numpy() if hasattr(v, "numpy") else v

    for i in range(6):
        assert np.array_equal(ld("output_pt/input/input.%d.pt" % i).astype(np.int8), g["q%d" % i]), "input.%d.pt" % i
    for i in range(5):
        assert np.array_equal(ld("output_pt/pe_add/pe_add_output%d.pt" % i).astype(np.int32), g["pe_add%d" % i])
        for p in range(4):
            assert np.array_equal(ld("output_pt/pe_out/pe_output%d_%d.pt" % (i, p)).astype(np.int32), g["pe%d_%d" % (i, p)])
        assert np.array_equal(ld("output_pt/bias/conv.bias.quan%d.pt" % i).reshape(-1).astype(np.int32), g["K%d" % i])
    tags = ["0_1", "1_2", "2_3", "3_4", "4_5", "res"]
    assert [int(ld("output_pt/requan_factor/requan_%s.pt" % t)) for t in tags] == [int(v) for v in g["M"]]
    assert [int(ld("output_pt/requan_factor/n_%s.pt" % t)) for t in tags] == [int(v) for v in g["n"]]
    assert np.array_equal(fresh_state.shortcut.cpu().numpy(), g["shortcut"])
    out = capsys.readouterr().out
    # the reference prints when an 18-bit PE clamp fires (:358-361); which side fires is in the golden dumps (the
    # synthetic stress case and -- with its real weights -- the denoise-only checkpoint)
    hit_hi = any((g["pe%d_%d" % (i, p)] == 2 ** 17 - 1).any() for i in range(5) for p in range(4))
    hit_lo = any((g["pe%d_%d" % (i, p)] == -2 ** 17).any() for i in range(5) for p in range(4))
    if case == "synth_clamp":
        assert hit_hi or hit_lo
    assert ("max_overflow" in out) == hit_hi and ("min_overflow" in out) == hit_lo


def test_stand_alone_ops_semantics(fresh_state):
    from myQL import quan_func as qf
    # reshape_input_for_hardware_pe: entry p keeps channels c == p (mod 4), zeros elsewhere (:298-318)
    x = torch.arange(2 * 6 * 3 * 5, dtype=torch.float32).reshape(2, 6, 3, 5).cuda() + 1
    s = qf.reshape_input_for_hardware_pe(x, pe_num=4)
    assert s.shape == (8, 6, 3, 5)
    for b in range(2):
        for p in range(4):
            for c in range(6):
                want = x[b, c] if c % 4 == p else torch.zeros_like(x[b, c])
                assert torch.equal(s[4 * b + p, c], want)
    # layers 1..3 quantiser: clamp(round(x + z)) half-to-even, returns q - max(z, -128) (:275-293)
    fresh_state.set_threshold(2, 1.0, -128, write_file=False)
    v = torch.tensor([[[[0.5, 1.5, 2.5, 254.5, 255.5, 300.0, -3.0, 0.49999]]]]).cuda()
    u = qf.quantize_asymmetrical_by_tensor(v, 8, 1, func_id=2)
    assert u.flatten().tolist() == [0.0, 2.0, 2.0, 254.0, 255.0, 255.0, 0.0, 0.0]
    fresh_state.set_threshold(3, 1.0, -131, write_file=False)       # z < -128: operand shifted by -128, not z
    u = qf.quantize_asymmetrical_by_tensor(v, 8, 1, func_id=3)
    assert u.flatten().tolist() == [0.0, 0.0, 0.0, 252.0, 252.0, 255.0, 0.0, 0.0]
    # requant: fp32 RN(acc * M16) * 2^-n (:529) -- products above 2^24 round
    fresh_state.set_threshold(1, 2.0, -128, write_file=False)
    fresh_state.weight_scale[1] = 0.001
    acc = torch.tensor([[[[100001.0, -77777.0, 3.0]]]]).cuda()
    y = qf.requan_conv2d_output(acc, 1, 1)
    m16, n16 = fresh_state.requan["1_2"]
    want = (acc.cpu() * m16) * 2 ** (-n16)
    assert torch.equal(y.cpu(), want)


@pytest.mark.parametrize("case", SMALL_CASES)
def test_fake_quant_graph_against_reference_mode0(case, fresh_state):
    g = load_golden(case)
    gm = quantize_graph(collapsed_model(g, sim=False, device="cuda"), 0, pe_add_bit=g.pe_add_bit)
    assert gm.fusion_plan() == "fq"
    y = gm(torch.from_numpy(g.x).cuda()).cpu().numpy()
    ref = g["fq_out"]
    mse = float(((y.astype(np.float64) - ref) ** 2).mean())
    s5 = (g["fq_maxs"][5] - g["fq_mins"][5]) / 255
    assert mse == 0 or 10 * np.log10(1.0 / mse) >= 60.0                      # PSNR at range 1.0
    assert float(np.abs(y - ref).max()) <= 8 * s5
    # the observer files the calibration tail reads (quan_func.py:184-195)
    mins = [torch.load("output_pt/input/input.%d.min_val.pt" % i) for i in range(6)]
    maxs = [torch.load("output_pt/input/input.%d.max_val.pt" % i) for i in range(6)]
    assert np.allclose(mins, g["fq_mins"], rtol=1e-5, atol=1e-7) and np.allclose(maxs, g["fq_maxs"], rtol=1e-5, atol=1e-7)


def test_observer_module_on_device():
    from myQL.quan_modules import myMinMaxObserver
    ob = myMinMaxObserver().cuda()
    gen = torch.Generator().manual_seed(5)
    a = torch.randn(3, 7, 129, generator=gen)
    b = torch.randn(100001, generator=gen) * 3
    assert ob(a.cuda()) is not None
    ob(b.cuda())
    assert float(ob.min_val) == float(torch.minimum(a.min(), b.min()))
    assert float(ob.max_val) == float(torch.maximum(a.max(), b.max()))


def test_quan_modules_debug_helpers(fresh_state):
    """conv_forward_naive (step-wise 18-bit clamp, reference quan_modules.py:43-99) and sesr_forward_sim (:101-201): the
    reference's hand-unrolled debugging path.  Without clamp events it must agree with the fused network (for the x4 model
    the final ReLU of sesr_forward_sim is a no-op because z5 = -128)."""
    import torch.nn.functional as F
    from myQL.quan_modules import conv_forward_naive, sesr_forward_sim
    from myQL.quan_func import quantize_model_weight
    import define
    gen = torch.Generator().manual_seed(0)
    x = torch.randint(0, 255, (1, 3, 9, 11), generator=gen).float().cuda()
    w = torch.randint(-20, 20, (4, 3, 3, 3), generator=gen).float().cuda()
    assert torch.equal(conv_forward_naive(x, w, 1, 1, 0), F.conv2d(x, w, padding=1))
    big = torch.full((1, 1, 3, 3), 255.0).cuda()
    wbig = torch.full((1, 1, 3, 3), 127.0).cuda()
    assert float(conv_forward_naive(big, wbig, 1, 1, 0).max()) == 2 ** 17 - 1          # 9 * 255 * 127 saturates at 18 bits

    g = load_golden("x4_small")
    set_thresholds(fresh_state, g)
    model = collapsed_model(g, sim=True, device="cuda")
    quantize_model_weight(model, define.QUAN_BIT, 1)                                    # fills state.weight_codes / scales
    for i, b in enumerate(g.bf):
        fresh_state.bias[i] = torch.from_numpy(b)
    y = sesr_forward_sim(torch.from_numpy(g.x).cuda())
    assert np.array_equal(y.cpu().numpy(), g["yout"])


@pytest.mark.parametrize("case", SMALL_CASES)
def test_per_node_fake_quant_graph_equals_fused(case, fresh_state, monkeypatch):
    """exe_mode 0 node by node (the stand-alone CUDA op of every reference callable: quantiser :175-215, PE split,
    Conv2d, reshape_ouput_for_hardware_pe :330-333, PEs_and_bias_adder :431-479) must give the fused fake-quant
    network's output bit for bit, and leave the same observer files."""
    import define
    from myQL import quan_func as qf
    g = load_golden(case)
    x = torch.from_numpy(g.x).cuda()
    gm = quantize_graph(collapsed_model(g, sim=False, device="cuda"), 0, pe_add_bit=g.pe_add_bit)
    assert gm.fusion_plan() == "fq"
    y_fused = gm(x).cpu().numpy()
    mm_fused = [(torch.load("output_pt/input/input.%d.min_val.pt" % i), torch.load("output_pt/input/input.%d.max_val.pt" % i))
                for i in range(6)]
    fresh_state.reset()
    for f in os.listdir("output_pt/input"):
        os.unlink(os.path.join("output_pt/input", f))
    monkeypatch.setattr(define, "INPUT_W_FLG", True)          # any dump flag selects the per-node path
    monkeypatch.setattr(qf, "INPUT_W_FLG", True)
    monkeypatch.setattr(define, "MFLAG", int(g["mflag"]))      # MFLAG 3 observes the last conv's output (:460-479)
    gm2 = quantize_graph(collapsed_model(g, sim=False, device="cuda"), 0, pe_add_bit=g.pe_add_bit)
    assert gm2.fusion_plan() == "graph"
    y_node = gm2(x).cpu().numpy()
    assert np.array_equal(y_node, y_fused)
    mm_node = [(torch.load("output_pt/input/input.%d.min_val.pt" % i), torch.load("output_pt/input/input.%d.max_val.pt" % i))
               for i in range(6)]
    assert mm_node == mm_fused
    # the reference also stores this call's scale / zero and (INPUT_W_FLG) the fake-quantised tensor of every domain
    for i in range(5):
        lo, hi = mm_node[i]
        assert torch.load("output_pt/input/input.%d.scale.pt" % i) == (hi - lo) / 255
        assert os.path.isfile("output_pt/input/input.%d.pt" % i)


def test_mode0_stand_alone_callables(fresh_state):
    from myQL import quan_func as qf
    gen = torch.Generator().manual_seed(9)
    x = (torch.rand(1, 3, 9, 13, generator=gen) * 3 - 1).cuda()
    y = qf.quantize_asymmetrical_by_tensor(x, 8, 0, func_id=2)
    lo, hi = float(x.min()), float(x.max())
    s = (hi - lo) / 255
    z = -128 - round(lo / s)
    want = (torch.clamp(torch.round(x.cpu() / s + z), -128, 127) - z) * s           # reference :207,:215 on the CPU
    assert torch.equal(y.cpu(), want)
    assert fresh_state.scale(2) == s and fresh_state.zero(2) == z
    with pytest.raises(AssertionError, match="Input tensor is all equal,3"):
        qf.quantize_asymmetrical_by_tensor(torch.ones(1, 1, 4, 4).cuda(), 8, 0, func_id=3)
    # reshape_ouput_for_hardware_pe mode 0: float clamp per PE entry, then the four entries are added (:330-333, :386)
    conv = (torch.randn(4, 2, 3, 3, generator=gen) * 50).cuda()
    out = qf.reshape_ouput_for_hardware_pe(conv, None, 4, 0.5, -3, 2.0, 0, 0)      # 4-bit accumulator: [-8, 7]
    fmax, fmin = (7 + 3) * 0.5 * 2.0, (-8 + 3) * 0.5 * 2.0
    c = torch.clamp(conv.cpu(), fmin, fmax)
    assert torch.equal(out.cpu(), ((c[0:1] + c[1:2]) + c[2:3]) + c[3:4])


def test_sim_per_node_writes_shortcut_file(fresh_state, monkeypatch):
    import define
    from myQL import quan_func as qf
    g = load_golden("x4_small")
    set_thresholds(fresh_state, g)
    monkeypatch.setattr(define, "RESIDUAL_W_FLG", True)
    gm = quantize_graph(collapsed_model(g, sim=True, device="cuda"), 1, pe_add_bit=g.pe_add_bit)
    assert gm.fusion_plan() == "graph"
    gm(torch.from_numpy(g.x).cuda())
    sc = torch.load("output_pt/residual/shortcut_tensor.pt", map_location="cpu")      # reference :546-549
    assert np.array_equal(sc.numpy(), g["shortcut"])
