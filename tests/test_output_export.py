"""Hardware test-vector export: the mirrors of the reference's output.py / output_end2end.py write
byte-identical ./output_txt files.  Goldens: sha256 of what the UNMODIFIED reference scripts wrote for the same dumps
(oracle/gen_output_golden.py -> tests/golden/output_txt_hashes.json).  CPU only."""
import hashlib
import json
import os
import runpy
import shutil

import pytest

from conftest import GOLD, PKG, load_golden

import gen_output_golden as gog

HASHES = json.load(open(os.path.join(GOLD, "output_txt_hashes.json")))


def tree(top):
    return gog.hash_tree(top)


@pytest.mark.parametrize("case", sorted(HASHES))
def test_export_scripts_byte_identical(case, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    gog.write_output_pt(load_golden(case))
    runpy.run_path(os.path.join(PKG, "output.py"), run_name="__main__")          # as the README says: python output.py
    assert tree("output_txt") == HASHES[case]["output.py"]
    shutil.rmtree("output_txt")
    runpy.run_path(os.path.join(PKG, "output_end2end.py"), run_name="__main__")
    assert tree("output_txt") == HASHES[case]["output_end2end.py"]


def test_hex_formatting_matches_scalar_path():
    import numpy as np
    import output as out
    rng = np.random.default_rng(0)
    for bits in (8, 16, 18, 20, 5.0):
        lim = int(2 ** (bits - 1))
        v = rng.integers(-lim, lim, size=(7, 33))
        rows = out.hex_rows(v, bits)
        assert rows == ["".join(out.float_to_hex(x, bits) for x in r) for r in v]
    assert out.float_to_hex(-1, 18) == "3ffff" and out.float_to_hex(-2, 20) == "ffffe" and out.float_to_hex(21, 5.0) == "15"
