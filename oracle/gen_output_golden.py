#!/usr/bin/env python
"""Golden hashes for the hardware test-vector export (build container only; TEST INFRASTRUCTURE).

For each small golden case the reference's ./output_pt dump files are rebuilt from tests/golden/<case>.npz (they ARE the
reference's dumps, see gen_golden.py), then the UNMODIFIED reference scripts output.py and output_end2end.py are run
on them in a scratch cwd, and the sha256 of every ./output_txt file is stored in tests/golden/output_txt_hashes.json.
tests/test_output_export.py runs the mirror (sesr_pytorch_quantize_b200/output.py, output_end2end.py) on the same
dumps and compares the hashes.

    python oracle/gen_output_golden.py
"""
import hashlib
import json
import os
import runpy
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("SESRQ_REFERENCE_ROOT", "/root/reference")
CASES = ["x4_small", "x2_small", "nrdm_small"]


def write_output_pt(g):
    """Rebuild the reference's dump tree from a golden case (same tensors, same dtypes, same file names)."""
    import numpy as np
    import torch
    for d in ("input", "bias", "requan_factor", "pe_out", "pe_add"):
        os.makedirs(os.path.join("output_pt", d), exist_ok=True)
    for i in range(6):
        torch.save(torch.from_numpy(g["q%d" % i].astype(np.float32)), "output_pt/input/input.%d.pt" % i)
    for i in range(5):
        torch.save(torch.from_numpy(g["K%d" % i].astype(np.float32))[None, :, None, None], "output_pt/bias/conv.bias.quan%d.pt" % i)
        torch.save(torch.from_numpy(g["pe_add%d" % i].astype(np.float32)), "output_pt/pe_add/pe_add_output%d.pt" % i)
        for p in range(4):
            torch.save(torch.from_numpy(g["pe%d_%d" % (i, p)].astype(np.float32)), "output_pt/pe_out/pe_output%d_%d.pt" % (i, p))
    tags = ["0_1", "1_2", "2_3", "3_4", "4_5", "res"]
    for t, M, n in zip(tags, g["M"], g["n"]):
        torch.save(int(M), "output_pt/requan_factor/requan_%s.pt" % t)
        torch.save(int(n), "output_pt/requan_factor/n_%s.pt" % t)


def hash_tree(top):
    out = {}
    for root, _, files in os.walk(top):
        for f in sorted(files):
            p = os.path.join(root, f)
            out[os.path.relpath(p, top)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    return out


def worker(case):
    sys.path[:0] = [os.path.join(ROOT, "tests"), ROOT, HERE]
    from conftest import load_golden
    write_output_pt(load_golden(case))
    sys.path.insert(0, REF)
    runpy.run_path(os.path.join(REF, "output.py"), run_name="__main__")
    h = {"output.py": hash_tree("output_txt")}
    import shutil
    shutil.rmtree("output_txt")
    runpy.run_path(os.path.join(REF, "output_end2end.py"), run_name="__main__")
    h["output_end2end.py"] = hash_tree("output_txt")
    print(json.dumps(h))


def main():
    if len(sys.argv) > 1:
        worker(sys.argv[1])
        return
    allh = {}
    for case in CASES:
        with tempfile.TemporaryDirectory(prefix="sesrq_out_") as tmp:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), case], cwd=tmp, check=True, capture_output=True, text=True)
            allh[case] = json.loads(r.stdout.strip().splitlines()[-1])
            print(case, sum(len(v) for v in allh[case].values()), "files")
    with open(os.path.join(ROOT, "tests", "golden", "output_txt_hashes.json"), "w") as f:
        json.dump(allh, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
