"""nrdm_3_sim.nr: 3-channel input, x1 no upscaling (denoise+demosaic), integer-sim variant (no AddOp).
Mirror of the reference's models/nrdm_3_sim.py (same constructor defaults and attribute names)."""
from models._sesr_family import FiveConvNet


class nr(FiveConvNet):
    def __init__(self, in_channels=3, out_channels=3, num_channels=16, num_lblocks=3, scaling_factor=1):
        super().__init__(in_channels, out_channels, num_channels, num_lblocks, scaling_factor, sim=True, shuffle=False)
