#!/bin/bash
# GPU box: ncu --set full capture (with source-level stall samples) of one fused-pair launch on 16 x 1080p; CSV exports only
mkdir -p gpurun_out
cat > /tmp/chain_one.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
os.environ["SESRQ_FUSE"] = "1"
import torch
from sesr_pytorch_quantize_b200 import Int8Net, compile_network
from sesr_pytorch_quantize_b200.fixtures import load_golden
g = load_golden("x2_fixture")
cnet = compile_network([torch.from_numpy(w) for w in g.wf], [torch.from_numpy(b) for b in g.bf], g.scales, g.zeros, g.r)
net = Int8Net(cnet, device="cuda:0")
x = torch.rand(16, 3, 1080, 1920).cuda()
for _ in range(3): net(x, want_q5=False)
torch.cuda.synchronize()
PY
python /tmp/chain_one.py || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_chain2 --launch-skip 2 --launch-count 1 -f -o /tmp/chain python /tmp/chain_one.py > gpurun_out/chain_ncu.log 2>&1
ncu -i /tmp/chain.ncu-rep --page raw --csv > gpurun_out/chain.raw.csv 2>/dev/null
ncu -i /tmp/chain.ncu-rep --page source --csv > gpurun_out/chain.src.csv 2>/dev/null
python profiles/ncu_tools.py metrics /tmp/chain.ncu-rep
python profiles/ncu_tools.py hot /tmp/chain.ncu-rep 0 40
