"""Stage the files of the reference that its own scripts need into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so
the staged tree travels to the GPU box like a built .so).  Run by __graft_entry__.build() whenever /root/reference is
present (the build container); on the GPU box the pre-staged copy is used as is.

What it is for (test / measurement infrastructure only -- no product file imports or reads oracle/_ref):
  * tests/test_scripts_gpu.py runs the UNMODIFIED sim.py / test.py against this package's import root;
  * bench.py's cpu_baseline / `--impl reference` leg times the real Python reference (myQL + models) on the box's
    host cores (kind "_ref") next to the C restatement (kind "port").
Nothing is copied into tracked paths.
"""
import os
import shutil
import sys

REF = os.environ.get("SESRQ_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

FILES = ["sim.py", "test.py", "define.py", "rand_DM_Input_80x960.pt", "rand_SR_Input_80x960.pt"]
DIRS = ["myQL", "models"]
PARAMS = ["nrdm_3_raw_G.pth", "x4sesr.pth", "x2sesr.pth.tar", "nr_G.pth", "dm_G.pth", "nrdm_6_G.pth", "nrdm_3_qat_G.pth"]


def stage(verbose=True):
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copy2(os.path.join(REF, f), os.path.join(DST, f))
    for d in DIRS:
        dst = os.path.join(DST, d)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, d), dst, ignore=shutil.ignore_patterns("__pycache__"))
    os.makedirs(os.path.join(DST, "model_params"), exist_ok=True)
    for f in PARAMS:
        shutil.copy2(os.path.join(REF, "model_params", f), os.path.join(DST, "model_params", f))
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print("staged %d reference files into %s" % (n, DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
