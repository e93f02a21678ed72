"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/sesrq.h declares.
No compute entry point is called here (no GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from sesr_pytorch_quantize_b200 import _lib


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sesrq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sesrq_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) >= 20
    assert sorted(_lib.SIGNATURES) == names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), "libsesrq.so does not export %s" % name


def test_load_and_error_strings():
    lib = _lib.load_library()
    assert lib.sesrq_abi_version() == _lib.ABI_VERSION
    assert lib.sesrq_strerror(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert len(lib.sesrq_strerror(code)) > 3
    assert isinstance(lib.sesrq_launch_count(), int)


def test_handle_entry_points_reject_a_null_handle():
    """Error behaviour of the entry points that take a net handle: a negative SESRQ_E* code (never a crash, never a
    positive value that could be mistaken for a result)."""
    lib = _lib.load_library()
    assert lib.sesrq_net_set_fusion(None, 1) == -1            # SESRQ_EINVAL
    assert lib.sesrq_net_layer_layout(None, 0) == -1
    assert lib.sesrq_net_fused_layers(None) == 0
    assert lib.sesrq_net_layer_info(None, 0, None, None) == -1
    assert lib.sesrq_ws_bytes(None, 1, 8, 8) == 0
    assert lib.sesrq_net_profile_enable(None, 1) == -1


def test_struct_layout_matches_header():
    # sizes the C compiler produces for the two descriptor structs (checked with a tiny C probe)
    import subprocess, tempfile, textwrap
    code = textwrap.dedent('''
        #include <stdio.h>
        #include "sesrq.h"
        int main(void){ printf("%zu %zu %zu\\n", sizeof(SesrqLayerDesc), sizeof(SesrqNetDesc), sizeof(SesrqFqDesc)); return 0; }
    ''')
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(code)
        exe = os.path.join(d, "p")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [ctypes.sizeof(_lib.LayerDesc), ctypes.sizeof(_lib.NetDesc), ctypes.sizeof(_lib.FqDesc)]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sesr_pytorch_quantize_b200 import Int8Net, SesrqError
    with pytest.raises(SesrqError):
        Int8Net(None)
