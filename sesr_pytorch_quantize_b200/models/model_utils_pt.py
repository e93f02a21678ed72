"""Collapsible linear blocks (reference models/model_utils_pt.py:5-92).

A block is a k x k "expand" conv to `tmp_channels` followed by a 1 x 1 "squeeze" conv; for inference the
pair is folded into a single k x k conv.  collapse() folds by pushing unit impulses through the two convs
-- the same construction as the reference (model_utils_pt.py:39-60), so on the same device the folded
weights are bit-identical to the reference's, which the quantised codes depend on.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

_ACTIVATIONS = {"prelu": nn.PReLU, "relu": nn.ReLU, "identity": nn.Identity}


class CollapsibleLinearBlock(nn.Module):
    def __init__(self, in_channels, out_channels, tmp_channels, kernel_size, activation='prelu'):
        super().__init__()
        if activation not in _ACTIVATIONS:
            raise Exception(f'Activation not supported: {activation}')
        half = int((kernel_size - 1) / 2)
        self.conv_expand = nn.Conv2d(in_channels, tmp_channels, (kernel_size, kernel_size), padding=half, bias=False)
        self.conv_squeeze = nn.Conv2d(tmp_channels, out_channels, (1, 1))
        self.activation = _ACTIVATIONS[activation]()
        self.collapsed = False

    def _linear(self, x):
        x = self.conv_expand(x)
        return x if self.collapsed else self.conv_squeeze(x)

    def forward(self, x):
        return self.activation(self._linear(x))

    def collapse(self):
        if self.collapsed:
            print('Already collapsed')
            return
        k = self.conv_expand.kernel_size[0]
        half = int((k - 1) / 2)
        cin = self.conv_expand.in_channels
        # one unit impulse per input channel, centred in a k x k window
        impulses = torch.eye(cin, device=self.conv_expand.weight.device)[:, :, None, None]
        impulses = F.pad(impulses, (half, half, half, half))
        with torch.no_grad():
            b = self.conv_squeeze.bias
            response = self.conv_squeeze(self.conv_expand(impulses)) - b[None, :, None, None]
        # the response to an impulse is the spatially flipped kernel, indexed [cin, cout]
        folded = torch.flip(response, [2, 3]).permute([1, 0, 2, 3])
        merged = nn.Conv2d(cin, self.conv_squeeze.out_channels, self.conv_expand.kernel_size, padding=half)
        merged.weight = nn.Parameter(folded)
        merged.bias = b
        self.conv_expand = merged
        self.conv_squeeze = nn.Identity()
        self.collapsed = True


class ResidualCollapsibleLinearBlock(CollapsibleLinearBlock):
    """x + squeeze(expand(x)); the identity is folded into the centre tap on collapse (reference :69-88)."""

    def forward(self, x):
        if self.collapsed:
            return self.activation(self.conv_expand(x))
        return self.activation(x + self.conv_squeeze(self.conv_expand(x)))

    def collapse(self):
        if self.collapsed:
            print('Already collapsed')
            return
        super().collapse()
        c = self.conv_expand.kernel_size[0] // 2
        with torch.no_grad():
            for ch in range(self.conv_expand.in_channels):
                self.conv_expand.weight[ch, ch, c, c] += 1.


class AddOp(nn.Module):
    def forward(self, x1, x2):
        return x1 + x2
