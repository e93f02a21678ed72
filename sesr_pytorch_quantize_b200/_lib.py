"""ctypes binding of libsesrq.so (the C ABI declared in include/sesrq.h).

There is no CPU fallback: if the shared library is missing, importing the binding raises, and any
compute entry point raises SesrqError when no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsesrq.so")

NUM_LAYERS = 5
MAX_CH = 16
NUM_DOMAINS = 6
KL_BINS = 2048
KL_TARGET_BINS = 128
ABI_VERSION = 1

IMPL_AUTO, IMPL_GENERIC, IMPL_UMMA = 0, 1, 2


class SesrqError(RuntimeError):
    pass


class LayerDesc(C.Structure):
    _fields_ = [
        ("ksize", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("relu", C.c_int32),
        ("z_in", C.c_int32), ("m16", C.c_int32), ("shift_n", C.c_int32), ("reserved0", C.c_int32),
        ("h_weight", C.c_void_p),
        ("add_const", C.c_int32 * MAX_CH),
    ]


class NetDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_layers", C.c_int32), ("r_shuffle", C.c_int32),
        ("pe_num", C.c_int32), ("pe_acc_bits", C.c_int32), ("pe_add_bits", C.c_int32),
        ("quan_bits", C.c_int32), ("impl", C.c_int32),
        ("scale_in", C.c_double), ("scale_out", C.c_double),
        ("z_out", C.c_int32), ("res_m16", C.c_int32), ("res_shift_n", C.c_int32),
        ("reserved0", C.c_int32),
        ("layers", LayerDesc * NUM_LAYERS),
    ]


class FqDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_layers", C.c_int32), ("r_shuffle", C.c_int32),
        ("residual", C.c_int32), ("observe_out", C.c_int32),
        ("pe_acc_bits", C.c_int32), ("pe_add_bits", C.c_int32), ("reserved0", C.c_int32),
        ("ksize", C.c_int32 * NUM_LAYERS), ("cin", C.c_int32 * NUM_LAYERS),
        ("cout", C.c_int32 * NUM_LAYERS), ("relu", C.c_int32 * NUM_LAYERS),
        ("h_weight", C.c_void_p * NUM_LAYERS),
        ("w_scale", C.c_double * NUM_LAYERS),
        ("h_bias", C.c_void_p * NUM_LAYERS),
    ]


_vp, _i, _sz, _f, _d = C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_double

# name -> (restype, argtypes); every symbol include/sesrq.h declares
SIGNATURES = {
    "sesrq_abi_version": (_i, []),
    "sesrq_strerror": (C.c_char_p, [_i]),
    "sesrq_last_cuda_error": (C.c_char_p, []),
    "sesrq_launch_count": (C.c_uint64, []),
    "sesrq_net_create": (_i, [C.POINTER(NetDesc), C.POINTER(_vp)]),
    "sesrq_net_destroy": (None, [_vp]),
    "sesrq_net_layer_info": (_i, [_vp, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "sesrq_ws_bytes": (_sz, [_vp, _i, _i, _i]),
    "sesrq_net_forward_i8": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "sesrq_net_forward_i8_host": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "sesrq_net_profile_enable": (_i, [_vp, _i]),
    "sesrq_net_layer_layout": (_i, [_vp, _i]),
    "sesrq_net_set_fusion": (_i, [_vp, _i]),
    "sesrq_net_fused_layers": (_i, [_vp]),
    "sesrq_net_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "sesrq_ws_read_act_i8": (_i, [_vp, _vp, _sz, _i, _i, _i, _i, _vp, _vp]),
    "sesrq_minmax_update": (_i, [_vp, _sz, _vp, _vp]),
    "sesrq_hist_update": (_i, [_vp, _sz, _f, _f, _vp, _vp]),
    "sesrq_kl_threshold": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "sesrq_fq_create": (_i, [C.POINTER(FqDesc), C.POINTER(_vp)]),
    "sesrq_fq_destroy": (None, [_vp]),
    "sesrq_fq_ws_bytes": (_sz, [_vp, _i, _i]),
    "sesrq_fq_ws_bytes_batch": (_sz, [_vp, _i, _i, _i]),
    "sesrq_fq_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sesrq_fq_forward_batch": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sesrq_op_quant_first": (_i, [_vp, _sz, _d, _i, _vp, _vp, _vp]),
    "sesrq_op_quant_next": (_i, [_vp, _sz, _i, _vp, _vp, _vp]),
    "sesrq_op_quant_residual": (_i, [_vp, _vp, _sz, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "sesrq_op_pe_split": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "sesrq_op_pe_merge": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "sesrq_op_requant": (_i, [_vp, _sz, _i, _i, _vp, _vp, _vp]),
    "sesrq_op_requant_final": (_i, [_vp, _sz, _i, _i, _i, _d, _vp, _vp, _vp]),
    "sesrq_op_conv_int": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    "sesrq_op_fakequant": (_i, [_vp, _sz, _d, _i, _vp, _vp]),
    "sesrq_op_conv_fq": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _d, _d, _vp, _vp]),
    "sesrq_op_pe_merge_fq": (_i, [_vp, _i, _i, _i, _f, _f, _f, _f, _vp, _vp, _vp]),
    "sesrq_qat_fakequant_fwd": (_i, [_vp, _sz, _sz, _i, _vp, _vp, _f, _f, _vp, _vp]),
    "sesrq_qat_fakequant_bwd": (_i, [_vp, _vp, _sz, _sz, _i, _vp, _vp, _vp, _vp, _i, _f, _f, _vp, _vp]),
    "sesrq_qat_qparams": (_i, [_vp, _vp, _i, _i, _f, _f, _f, _vp, _vp, _vp]),
    "sesrq_qat_observe": (_i, [_vp, _i, _sz, _i, _f, _vp, _vp, _vp, _vp]),
    "sesrq_kth_abs": (_i, [_vp, _sz, C.c_ulonglong, _vp, _vp, _sz, _vp]),
    "sesrq_minmax_rows": (_i, [_vp, _i, _sz, _vp, _vp, _vp]),
    "sesrq_op_ssim": (_i, [_vp, _vp, _i, _i, _i, _d, _vp, _vp]),
    "sesrq_op_sr_post": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load libsesrq.so and bind every declared symbol.  Raises SesrqError if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SESRQ_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise SesrqError(
            "libsesrq.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the sesrq hot path)" % p)
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.sesrq_abi_version() != ABI_VERSION:
        raise SesrqError("libsesrq ABI %d != binding ABI %d" % (lib.sesrq_abi_version(), ABI_VERSION))
    if path is None:
        _lib = lib
    return lib


def check(code: int, what: str = "sesrq") -> None:
    if code == 0:
        return
    lib = load_library()
    msg = lib.sesrq_strerror(code).decode()
    if code == -3:
        msg += " [" + lib.sesrq_last_cuda_error().decode() + "]"
    raise SesrqError("%s failed: %s (code %d)" % (what, msg, code))
