"""GPU box: fused (layers 1+2 in one launch) vs per-layer path, bit-exact on a set of shapes, plus step timing.
usage: python tools/chain_check.py [case ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sesr_pytorch_quantize_b200 import Int8Net, compile_network
from sesr_pytorch_quantize_b200.fixtures import load_golden


def make(case, fuse):
    os.environ["SESRQ_FUSE"] = "1" if fuse else "0"
    g = load_golden(case)
    cnet = compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], g.scales, g.zeros, g.r)
    return Int8Net(cnet, device="cuda:0"), g


def main():
    cases = sys.argv[1:] or ["x2_fixture", "x4_fixture", "nrdm_fixture"]
    shapes = [(2, 64, 200), (1, 5, 7), (3, 37, 131), (4, 80, 960), (2, 270, 500), (1, 1080, 1920), (16, 1080, 1920)]
    ok = True
    for case in cases:
        ref, g = make(case, False)
        fus, _ = make(case, True)
        cin = g.wf[0].shape[1]
        for (B, H, W) in shapes:
            gen = torch.Generator().manual_seed(B * 1000 + H)
            x = torch.rand(B, cin, H, W, generator=gen).cuda()
            q0, y0 = ref(x)
            torch.cuda.synchronize()
            q1, y1 = fus(x)
            torch.cuda.synchronize()
            same = torch.equal(q0, q1) and torch.equal(y0, y1)
            ok &= same
            msg = ""
            if not same:
                d = (q0 != q1)
                idx = d.nonzero()
                msg = " mismatches %d of %d, first %s" % (int(d.sum()), d.numel(), idx[:5].tolist())
            ts = []
            for net in (ref, fus):
                for _ in range(3): net(x, want_q5=False)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10): net(x, want_q5=False)
                e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 10)
            print("%-13s %2dx%4dx%4d  %s  per-layer %.4f ms  fused %.4f ms%s" % (case, B, H, W, "OK  " if same else "DIFF", ts[0], ts[1], msg), flush=True)
    print("ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)


main()
