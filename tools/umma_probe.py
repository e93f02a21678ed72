"""Stand-alone probe of the tcgen05 path (run under `timeout` on the GPU box): one small case, layer by
layer against the oracle, so a descriptor mistake shows up as a located mismatch rather than a hang."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import sesr_oracle as orc
from conftest import load_golden
from sesr_pytorch_quantize_b200 import Int8Net, compile_network, IMPL_AUTO

case = sys.argv[1] if len(sys.argv) > 1 else "x4_fixture"
shape = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (1, 40, 300)
g = load_golden(case)
cnet = compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], g.scales, g.zeros, g.r,
                       pe_add_bits=g.pe_add_bit)
net = Int8Net(cnet, impl=IMPL_AUTO, fuse=False)      # layer by layer: every intermediate domain is read back
print("layer impl/pe_exact:", [net.layer_info(i) for i in range(5)], flush=True)
B, H, W = shape
x = torch.rand(B, g.wf[0].shape[1], H, W, generator=torch.Generator().manual_seed(5))
onet = orc.derive_net(g.wf, g.bf, g.scales, g.zeros, g.r, pe_add_bit=g.pe_add_bit)
tr = {}
q5_ref, y_ref = orc.forward_int(onet, x.numpy(), trace=tr)
q5, y = net(x.cuda())
torch.cuda.synchronize()
print("forward done", flush=True)
ok = True
for d in range(1, 5):
    q = net.read_activation(d).cpu().numpy()
    bad = int((q != tr["q%d" % d]).sum())
    print("q%d mismatches: %d / %d" % (d, bad, q.size))
    if bad:
        ok = False
        idx = np.argwhere(q != tr["q%d" % d])[:8]
        for i in idx:
            print("   at", tuple(i), "got", q[tuple(i)], "want", tr["q%d" % d][tuple(i)])
        break
bad5 = int((q5.cpu().numpy() != q5_ref).sum())
print("q5 mismatches:", bad5, " y equal:", bool(np.array_equal(y.cpu().numpy(), y_ref)))
sys.exit(0 if ok and bad5 == 0 else 1)
