"""dm.dm: 3 -> 3 channel demosaicker, float / fake-quant variant.  Mirror of the reference's models/dm.py (test.py mflag 2)."""
from models._sesr_family import FiveConvNet


class dm(FiveConvNet):
    def __init__(self, in_channels=3, out_channels=3, num_channels=16, num_lblocks=3, scaling_factor=1):
        super().__init__(in_channels, out_channels, num_channels, num_lblocks, scaling_factor, sim=False, shuffle=False, spare_add=True)
