"""Every import statement of the reference's sim.py / test.py resolves against this package's import root with the
path order INTEGRATION.md prescribes (package root first) -- names included (`from X import a, b` needs X.a and X.b).
Third parties that are absent from this image and irrelevant to the path (cv2, skimage) and the private dataset module
(self_dataset: h5py + data files) resolve to tests/stubs; everything else must be the mirror or a real installed module.
The scripts are parsed, not executed (tests/test_scripts_gpu.py executes them on the GPU box)."""
import ast
import importlib
import os
import sys

import pytest

from conftest import PKG, ROOT

STUBS = os.path.join(ROOT, "tests", "stubs")
CANDIDATES = [os.path.join(ROOT, "oracle", "_ref"), "/root/reference"]
REF = next((d for d in CANDIDATES if os.path.isfile(os.path.join(d, "sim.py"))), None)
STUBBED = {"cv2", "skimage", "self_dataset"}
MIRRORED = {"define", "myQL", "models"}


def top_level_imports(path):
    tree = ast.parse(open(path, encoding="utf-8").read())
    for node in tree.body:
        if isinstance(node, ast.Import):
            for a in node.names:
                yield a.name, []
        elif isinstance(node, ast.ImportFrom):
            yield node.module, [a.name for a in node.names]


@pytest.mark.skipif(REF is None, reason="reference scripts not available (oracle/_ref not staged)")
@pytest.mark.parametrize("script", ["sim.py", "test.py"])
def test_every_import_of_the_script_resolves_to_the_mirror(script, monkeypatch):
    monkeypatch.syspath_prepend(STUBS)
    monkeypatch.syspath_prepend(PKG)
    seen = 0
    for module, names in top_level_imports(os.path.join(REF, script)):
        mod = importlib.import_module(module)
        top = module.split(".")[0]
        origin = os.path.realpath(getattr(mod, "__file__", "") or "")
        if top in MIRRORED:
            assert origin.startswith(os.path.realpath(PKG)), "%s resolved to %s, not to the mirror" % (module, origin)
        elif top in STUBBED:
            assert origin.startswith(os.path.realpath(STUBS)), "%s resolved to %s" % (module, origin)
        else:
            assert not origin.startswith(os.path.realpath(REF)), "%s came from the reference checkout" % module
        for n in names:
            if n == "*":
                continue
            if not hasattr(mod, n):
                importlib.import_module(module + "." + n)       # `from models import dm` names a submodule
        seen += 1
    assert seen >= 15
    assert not any(os.path.realpath(p).startswith(os.path.realpath(REF)) for p in sys.path if p)
