"""Device-agnostic spelling of the QAT library (mirror of the reference's models/quantize_utils_pt.py, which differs
from quantize_utils_cuda.py only in not calling `.cuda()` on its buffers, `diff` of the two reference files).  Here the
buffers always follow the module, so both names are the same implementation; the arithmetic still needs CUDA tensors."""
from models.quantize_utils_cuda import *  # noqa: F401,F403
from models.quantize_utils_cuda import (_FakeQuantSTE, _make_quantizer)  # noqa: F401
