"""Shared body of the five-conv SESR / NRDM networks.  The public modules (sesr, sesr_sim, sesr_arch,
sesr_arch_sim, nrdm_3, nrdm_3_sim) only pick the defaults; `sim=True` drops the AddOp residual because the
integer-sim path re-implements it inside the quantiser of the last conv (reference quan_func.py:244-270)."""
from torch import nn

from models.model_utils_pt import AddOp, CollapsibleLinearBlock, ResidualCollapsibleLinearBlock


class FiveConvNet(nn.Module):
    def __init__(self, in_channels, out_channels, num_channels, num_lblocks, scaling_factor, sim, shuffle, spare_add=False):
        super().__init__()
        self.conv_first = CollapsibleLinearBlock(in_channels=in_channels, out_channels=num_channels,
                                                 tmp_channels=256, kernel_size=5, activation='relu')
        self.residual_block = nn.Sequential(*[
            ResidualCollapsibleLinearBlock(in_channels=num_channels, out_channels=num_channels,
                                           tmp_channels=256, kernel_size=3, activation='relu')
            for _ in range(num_lblocks)])
        if not sim:
            self.add_residual = AddOp()
        self.conv_last = CollapsibleLinearBlock(in_channels=num_channels,
                                                out_channels=out_channels * scaling_factor ** 2,
                                                tmp_channels=256, kernel_size=5, activation='identity')
        if spare_add:
            # never called by forward(), like upstream (models/nrdm_3.py:30,46); kept because prepare() swaps it for a
            # QuantAdd whose buffers are keys of the shipped *_qat_G.pth checkpoints
            self.add_upsampled_input = AddOp()
        if shuffle:
            self.depth_to_space = nn.PixelShuffle(scaling_factor)
        self._sim = sim
        self._shuffle = shuffle

    def collapse(self):
        self.conv_first.collapse()
        for layer in self.residual_block:
            layer.collapse()
        self.conv_last.collapse()

    def before_quantization(self):
        self.collapse()

    def forward(self, input):
        initial_features = self.conv_first(input)
        residual_features = self.residual_block(initial_features)
        if not self._sim:
            residual_features = self.add_residual(residual_features, initial_features)
        output = self.conv_last(residual_features)
        if self._shuffle:
            output = self.depth_to_space(output)
        return output
