import torch
import torch.nn as nn


def quantize_ste(x):
    """round() with a straight-through gradient (compressai.ops.quantize_ste)."""
    return (torch.round(x) - x).detach() + x


class LowerBound(nn.Module):
    """max(x, bound) (forward semantics only; the custom backward is irrelevant here)."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return torch.max(x, self.bound)


class NonNegativeParametrizer(nn.Module):
    """p_eff = max(p, sqrt(minimum + 2^-36))^2 - 2^-36 (GDN beta/gamma re-parametrisation)."""

    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        bound = (self.minimum + self.reparam_offset ** 2) ** 0.5
        self.lower_bound = LowerBound(bound)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        out = self.lower_bound(x)
        return out ** 2 - self.pedestal
