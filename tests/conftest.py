import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

ALL_CASES = ["x4_fixture", "nrdm_fixture", "x2_fixture", "x4_small", "nrdm_small", "x2_small",
             "synth_clamp", "synth_zero", "synth_neg"]
SMALL_CASES = [c for c in ALL_CASES if not c.endswith("_fixture")]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


class Golden:
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLD, name + ".npz"))
        self.z = {k: z[k] for k in z.files}
        self.hashes = json.loads(bytes(self.z["hashes_json"]).decode())
        self.scales = [float(v) for v in self.z["scales"]]
        self.zeros = [int(v) for v in self.z["zeros"]]
        self.r = int(self.z["r"])
        self.pe_add_bit = int(self.z["pe_add_bit"])
        self.x = self.z["x"]
        self.wf = [self.z["wf%d" % i] for i in range(5)]
        self.bf = [self.z["bf%d" % i] for i in range(5)]

    def __getitem__(self, k):
        return self.z[k]

    def __contains__(self, k):
        return k in self.z


_cache = {}


def load_golden(name):
    if name not in _cache:
        _cache[name] = Golden(name)
    return _cache[name]


@pytest.fixture(params=ALL_CASES)
def golden(request):
    return load_golden(request.param)


@pytest.fixture(params=SMALL_CASES)
def golden_small(request):
    return load_golden(request.param)
