"""nrdm_3.nr: 3-channel input, x1 no upscaling (denoise+demosaic), float / fake-quant variant (AddOp residual).
Mirror of the reference's models/nrdm_3.py (same constructor defaults and attribute names)."""
from models._sesr_family import FiveConvNet


class nr(FiveConvNet):
    def __init__(self, in_channels=3, out_channels=3, num_channels=16, num_lblocks=3, scaling_factor=1):
        super().__init__(in_channels, out_channels, num_channels, num_lblocks, scaling_factor, sim=False, shuffle=False, spare_add=True)
