"""Test stub for the reference's self_dataset.py (needs h5py / scipy.io and a private dataset): the same class name and
constructor, a handful of synthetic frames.  Frames follow the recipe the golden fixtures were calibrated on
(oracle/gen_golden.py: ONE generator seeded 1234, N sequential torch.rand(1, Cin, H, W) draws)."""
import os

import torch
from torch.utils.data import Dataset

from define import TEST_RAW_ADD_NOISE  # noqa: F401  (the reference module imports it too)

_CIN = {1: 3, 2: 3, 3: 3, 4: 3, 5: 1, 6: 3}
_COUT = {1: 3, 2: 3, 3: 3, 4: 3, 5: 1, 6: 3}
_R = {1: 1, 2: 1, 3: 1, 4: 1, 5: 4, 6: 2}


class TestDataset(Dataset):
    __test__ = False            # not a pytest class

    def __init__(self, mflag):
        self.mflag = mflag
        n = int(os.environ.get("SESRQ_STUB_FRAMES", "3"))
        h = int(os.environ.get("SESRQ_STUB_H", "80"))
        w = int(os.environ.get("SESRQ_STUB_W", "960"))
        g = torch.Generator().manual_seed(1234)
        self.frames = [torch.rand(1, _CIN[mflag], h, w, generator=g)[0] for _ in range(n)]
        r = _R[mflag]
        gg = torch.Generator().manual_seed(4321)
        self.gts = [torch.rand(_COUT[mflag], h * r, w * r, generator=gg) for _ in range(n)]

    def __len__(self):
        return len(self.frames)

    def __getitem__(self, i):
        return self.frames[i], self.gts[i]
