"""sesr_pytorch_quantize_b200 -- B200-native (sm_100a) SESR / NRDM quantized-inference hot path.

Host side is Python/PyTorch (device memory, streams, torch.distributed plumbing); all arithmetic of
the hot path runs in hand-written CUDA behind the C ABI of libsesrq.so (include/sesrq.h).
The directory of this package is also the drop-in import root for the reference's scripts
(`define`, `myQL`, `models`): see INTEGRATION.md.
"""
from ._lib import SesrqError, load_library, IMPL_AUTO, IMPL_GENERIC, IMPL_UMMA  # noqa: F401
from .compile import (CompiledLayer, CompiledNet, compile_network, collapsed_convs,  # noqa: F401
                      quan_layer_between_const, thresholds_from_minmax)
from .net import GraphedInt8Net, Int8Net  # noqa: F401

__all__ = ["SesrqError", "load_library", "IMPL_AUTO", "IMPL_GENERIC", "IMPL_UMMA", "CompiledLayer",
           "CompiledNet", "compile_network", "collapsed_convs", "quan_layer_between_const",
           "thresholds_from_minmax", "Int8Net", "GraphedInt8Net"]
