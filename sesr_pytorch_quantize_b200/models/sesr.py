"""sesr.sesr: 1-channel input, x4 PixelShuffle, float / fake-quant variant (AddOp residual).
Mirror of the reference's models/sesr.py (same constructor defaults and attribute names)."""
from models._sesr_family import FiveConvNet


class sesr(FiveConvNet):
    def __init__(self, in_channels=1, out_channels=1, num_channels=16, num_lblocks=3, scaling_factor=4):
        super().__init__(in_channels, out_channels, num_channels, num_lblocks, scaling_factor, sim=False, shuffle=True)
