"""Reader for the committed reference vectors (tests/golden/*.npz, produced by oracle/gen_golden.py from the unmodified
reference): collapsed float weights / biases of the three shipped networks, the reference-calibrated thresholds and the
reference's outputs.  Used by the tests, by bench.py (weights + thresholds of the benchmarked network) and by smoke()."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


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
