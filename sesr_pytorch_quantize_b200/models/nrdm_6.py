"""nrdm_6.nr: the 8-conv (6 residual blocks) denoise+demosaic float model.  Mirror of the reference's models/nrdm_6.py
(test.py mflag 4).  Its quantised graph is NOT the 5-conv chain: the reference's own quantisation ops hard-code the
residual hand-off at func_id 3/4 (quan_func.py:244,554,567), so upstream only the float model is usable; here the graph
passes build it and FusedGraphModule reports 'graph' (no fused path)."""
from models._sesr_family import FiveConvNet


class nr(FiveConvNet):
    def __init__(self, in_channels=3, out_channels=3, num_channels=16, num_lblocks=6, scaling_factor=1):
        super().__init__(in_channels, out_channels, num_channels, num_lblocks, scaling_factor, sim=False, shuffle=False, spare_add=True)
