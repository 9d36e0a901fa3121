"""ResnetBlockFC (reference: tomosar2height/block/resnet.py:4-54).

out = shortcut(x) + fc_1(relu(fc_0(relu(x)))), with a bias-free linear shortcut iff
size_in != size_out.  Parameters are ordinary fp32 ``nn.Linear`` weights with the reference's
names (fc_0, fc_1, shortcut) so checkpoints load unchanged.
"""
import torch.nn as nn

from ..linear import linear


class ResnetBlockFC(nn.Module):
    def __init__(self, size_in, size_out=None, size_h=None):
        super().__init__()
        size_out = size_in if size_out is None else size_out
        size_h = min(size_in, size_out) if size_h is None else size_h
        self.size_in, self.size_h, self.size_out = size_in, size_h, size_out
        self.fc_0 = nn.Linear(size_in, size_h)
        self.fc_1 = nn.Linear(size_h, size_out)
        self.actvn = nn.ReLU()
        self.shortcut = None if size_in == size_out else nn.Linear(size_in, size_out, bias=False)
        nn.init.zeros_(self.fc_1.weight)  # resnet.py:34 (overwritten by the model-level Xavier init)

    def forward(self, x, x2=None):
        """``x`` (..., size_in), or the two halves ``x | x2`` of a concatenated input (the
        ``torch.cat([net, pooled])`` of pointnet.py:78 is never materialised).  Three tcgen05 GEMMs with
        ReLU-on-load, bias and the residual add fused (t2h_linear_fwd)."""
        hidden = linear(x, self.fc_0.weight, self.fc_0.bias, x2=x2, relu_in=True)
        if self.shortcut is None:
            if x2 is not None:
                raise RuntimeError("ResnetBlockFC: a split input needs a shortcut projection")
            skip = x
        else:
            skip = linear(x, self.shortcut.weight, None, x2=x2)
        return linear(hidden, self.fc_1.weight, self.fc_1.bias, relu_in=True, residual=skip)
