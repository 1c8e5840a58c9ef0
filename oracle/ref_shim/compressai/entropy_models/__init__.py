import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from compressai.ops import LowerBound


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    offset = property(lambda self: self._offset)
    quantized_cdf = property(lambda self: self._quantized_cdf)
    cdf_length = property(lambda self: self._cdf_length)

    def quantize(self, inputs, mode, means=None):
        if mode not in ("dequantize", "symbols"):
            raise ValueError(f'shim supports eval modes only, got "{mode}"')
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)
        if mode == "dequantize":
            if means is not None:
                outputs += means
            return outputs
        return outputs.int()


class EntropyBottleneck(EntropyModel):
    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        self.matrices = nn.ParameterList()
        self.biases = nn.ParameterList()
        self.factors = nn.ParameterList()
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            self.matrices.append(nn.Parameter(torch.full((channels, filters[i + 1], filters[i]), float(init))))
            self.biases.append(nn.Parameter(torch.empty(channels, filters[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.factors.append(nn.Parameter(torch.zeros(channels, filters[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.Tensor([-self.init_scale, 0, self.init_scale]).repeat(channels, 1, 1))
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2].detach()

    def update(self, force=False, **kw):
        return False  # quantised-CDF tables feed the rANS coder only (out of scope)

    def _logits_cumulative(self, inputs):
        logits = inputs
        for i in range(len(self.filters) + 1):
            logits = torch.matmul(F.softplus(self.matrices[i]), logits) + self.biases[i]
            if i < len(self.filters):
                logits = logits + torch.tanh(self.factors[i]) * torch.tanh(logits)
        return logits

    def forward(self, x, training=None):
        perm = list(range(x.dim()))
        perm[0], perm[1] = 1, 0
        xp = x.permute(*perm).contiguous()
        shape = xp.size()
        values = xp.reshape(xp.size(0), 1, -1)
        outputs = self.quantize(values, "dequantize", self._get_medians())
        lower = self._logits_cumulative(outputs - 0.5)
        upper = self._logits_cumulative(outputs + 0.5)
        likelihood = torch.sigmoid(upper) - torch.sigmoid(lower)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        outputs = outputs.reshape(shape).permute(*perm).contiguous()
        likelihood = likelihood.reshape(shape).permute(*perm).contiguous()
        return outputs, likelihood

    # The rANS coder is unavailable; compress/decompress degrade to quantise/de-quantise
    # so that MLICPlusPlus.compress() can walk the network (mlicpp.py:205-206).
    def compress(self, x, **kw):
        med = self._get_medians().reshape(1, -1, 1, 1)
        return torch.round(x - med).int()

    def decompress(self, strings, size, **kw):
        med = self._get_medians().reshape(1, -1, 1, 1)
        return strings.to(med.dtype) + med


class EntropyBottleneckVbr(EntropyBottleneck):
    """compressai.entropy_models.EntropyBottleneckVbr (1.2.x), restated from its published behaviour (the package is not installable
    here: UNPINNED like the rest of this shim): the factorised prior evaluated on a grid of step `qs` instead of 1 --
    quantise  z_hat = round((z - median) / qs) * qs + median,  likelihood = sigmoid(c(z_hat + qs / 2)) - sigmoid(c(z_hat - qs / 2)).
    Only reachable with vr_entbttlnck=True (mlicpp_vbr.py:104-117,253-259,553-559); qs=None is the plain bottleneck."""

    def update_variable(self, force=False, qs=1.0):
        return False  # quantised-CDF tables feed the rANS coder only

    def forward(self, x, training=None, qs=None, ste=False):
        if qs is None:
            return super().forward(x, training)
        perm = list(range(x.dim()))
        perm[0], perm[1] = 1, 0
        xp = x.permute(*perm).contiguous()
        shape = xp.size()
        values = xp.reshape(xp.size(0), 1, -1)
        med = self._get_medians()
        outputs = torch.round((values - med) / qs) * qs + med
        half = 0.5 * qs
        lower = self._logits_cumulative(outputs - half)
        upper = self._logits_cumulative(outputs + half)
        likelihood = torch.sigmoid(upper) - torch.sigmoid(lower)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        outputs = outputs.reshape(shape).permute(*perm).contiguous()
        likelihood = likelihood.reshape(shape).permute(*perm).contiguous()
        return outputs, likelihood

    def compress(self, x, qs=None, **kw):
        if qs is None:
            return super().compress(x)
        med = self._get_medians().reshape(1, -1, 1, 1)
        return torch.round((x - med) / qs).int(), qs

    def decompress(self, strings, size, qs=None, **kw):
        if isinstance(strings, tuple):          # what compress(qs=...) above returned
            strings = strings[0]
        med = self._get_medians().reshape(1, -1, 1, 1)
        if qs is None:
            return strings.to(med.dtype) + med
        return strings.to(med.dtype) * qs + med


class GaussianConditional(EntropyModel):
    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))
        self.tail_mass = float(tail_mass)
        self.lower_bound_scale = LowerBound(scale_bound)

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        # quantised CDFs (rANS only) are not rebuilt; mark as initialised
        n = len(self.scale_table)
        self._offset = torch.zeros(n, dtype=torch.int32)
        self._cdf_length = torch.zeros(n, dtype=torch.int32)
        self._quantized_cdf = torch.zeros(n, 1, dtype=torch.int32)
        return True

    @staticmethod
    def _standardized_cumulative(inputs):
        return 0.5 * torch.erfc(-(2 ** -0.5) * inputs)

    def _likelihood(self, inputs, scales, means=None):
        values = inputs - means if means is not None else inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((0.5 - values) / scales)
        lower = self._standardized_cumulative((-0.5 - values) / scales)
        return upper - lower

    def forward(self, inputs, scales, means=None, training=None):
        outputs = self.quantize(inputs, "dequantize", means)
        likelihood = self._likelihood(outputs, scales, means)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        return outputs, likelihood

    def build_indexes(self, scales):
        scales = self.lower_bound_scale(scales)
        indexes = scales.new_full(scales.size(), len(self.scale_table) - 1).int()
        for s in self.scale_table[:-1]:
            indexes -= (scales <= s).int()
        return indexes
