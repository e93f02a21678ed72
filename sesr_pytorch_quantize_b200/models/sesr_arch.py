"""sesr_arch.sesr: 3-channel input, x2 PixelShuffle, float / fake-quant variant (AddOp residual).
Mirror of the reference's models/sesr_arch.py (same constructor defaults and attribute names)."""
from models._sesr_family import FiveConvNet


class sesr(FiveConvNet):
    def __init__(self, in_channels=3, out_channels=3, num_channels=16, num_lblocks=3, scaling_factor=2):
        super().__init__(in_channels, out_channels, num_channels, num_lblocks, scaling_factor, sim=False, shuffle=True, spare_add=True)
