"""Entropy-coder hand-off of compress() / decompress(): quantised CDF tables and the range-ANS coder.

The reference delegates all of this to CompressAI 1.2.6 (un-vendored, not installable here; requirements.txt:31):
  * GaussianConditional.update_scale_table / EntropyBottleneck.update build `_quantized_cdf`, `_cdf_length`, `_offset`
    (called through MLIC++/models/mlicpp.py:470-475);
  * BufferedRansEncoder / RansDecoder code the symbol lists (MLIC++/models/mlicpp.py:212-216,279-280,303-304).
Both are restated here from the package's published algorithm; the arithmetic of the coder itself lives in
mlic_b200/csrc/rans.cpp behind the C ABI (include/mlic_b200.h).  The table builders run once per update() on the host in
plain torch (they are not on the forward path).  PARITY with CompressAI's bytes is UNPINNED (nothing to run against);
tests/test_coder.py pins round trips and table invariants."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib

TAIL_MASS = 1e-9


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pmf_to_quantized_cdf(pmf):
    """compressai._CXX.pmf_to_quantized_cdf(pmf, 16): 1-D float pmf (tail mass last) -> int32 cdf of len(pmf) + 1."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(p.size + 1, dtype=np.int32)
    r = _lib.lib().mlic_pmf_to_quantized_cdf(_ptr(p), int(p.size), _ptr(out))
    if r:
        raise ValueError(f"pmf_to_quantized_cdf failed ({r}): pmf must be finite, non-negative and not all zero")
    return out


def _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
    """EntropyModel._pmf_to_cdf: one row per table, padded with zeros to max_length + 2."""
    cdf = np.zeros((len(pmf_length), max_length + 2), dtype=np.int32)
    for i, n in enumerate(pmf_length):
        prob = np.concatenate([pmf[i, :n], tail_mass[i]])
        c = pmf_to_quantized_cdf(prob)
        cdf[i, :c.size] = c
    return cdf


def _std_cum(x):
    return 0.5 * torch.erfc(-(2 ** -0.5) * x)


def _norm_ppf(p):
    """scipy.stats.norm.ppf through erfinv (double precision)."""
    return math.sqrt(2.0) * float(torch.erfinv(torch.tensor(2.0 * p - 1.0, dtype=torch.float64)))


def gaussian_tables(scale_table):
    """GaussianConditional.update(): -> (quantized_cdf int32 [L, max+2], cdf_length int32 [L], offset int32 [L])."""
    st = torch.as_tensor(scale_table, dtype=torch.float32).cpu()
    multiplier = -_norm_ppf(TAIL_MASS / 2)
    pmf_center = torch.ceil(st * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = int(pmf_length.max())
    samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
    scale = st.unsqueeze(1)
    upper = _std_cum((0.5 - samples) / scale)
    lower = _std_cum((-0.5 - samples) / scale)
    pmf = (upper - lower).numpy()
    tail = (2 * lower[:, :1]).numpy()
    cdf = _pmf_to_cdf(pmf, tail, pmf_length.tolist(), max_length)
    return cdf, (pmf_length + 2).numpy().astype(np.int32), (-pmf_center).numpy().astype(np.int32)


def bottleneck_tables(eb, qs=1.0):
    """EntropyBottleneck.update() on the module's parameters (matrices / biases / factors / quantiles); qs != 1: the tables of
    EntropyBottleneckVbr.update_variable(qs) -- the same construction on a grid of step qs around the medians (symbols are
    round((z - median) / qs); restated from CompressAI's published behaviour, unpinned like the rest of the coder)."""
    q = eb.quantiles.detach().float().cpu()
    qs = float(qs)
    medians = q[:, 0, 1]
    minima = torch.clamp(torch.ceil((medians - q[:, 0, 0]) / qs).int(), min=0)
    maxima = torch.clamp(torch.ceil((q[:, 0, 2] - medians) / qs).int(), min=0)
    pmf_start = medians - minima * qs
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max())
    samples = torch.arange(max_length)[None, :] * qs + pmf_start[:, None, None]       # [C,1,L]

    def cum(x):
        logits = x
        for i in range(5):
            m = getattr(eb.matrices, str(i)).detach().float().cpu()
            b = getattr(eb.biases, str(i)).detach().float().cpu()
            logits = torch.matmul(torch.nn.functional.softplus(m), logits) + b
            if i < 4:
                f = getattr(eb.factors, str(i)).detach().float().cpu()
                logits = logits + torch.tanh(f) * torch.tanh(logits)
        return logits

    lower, upper = cum(samples - 0.5 * qs), cum(samples + 0.5 * qs)
    # CompressAI EntropyBottleneck.update(): the difference is taken on the side of the sigmoid where it does not cancel
    # (sign = -sign(lower + upper)); equal to sigmoid(upper) - sigmoid(lower) in exact arithmetic, not in fp32 in the upper tail
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    cdf = _pmf_to_cdf(pmf.numpy(), tail.numpy(), pmf_length.tolist(), max_length)
    return cdf, (pmf_length + 2).numpy().astype(np.int32), (-minima).numpy().astype(np.int32)


def encode_with_indexes(symbols, indexes, cdf, cdf_lengths, offsets):
    """BufferedRansEncoder.encode_with_indexes(...) + flush() -> bytes."""
    sym, idx, cdf, ln, off = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1), _i32(cdf), _i32(cdf_lengths), _i32(offsets)
    if sym.size != idx.size:
        raise ValueError("symbols and indexes differ in length")
    L = _lib.lib()
    cap = int(L.mlic_rans_encode_bound(sym.size))
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    r = L.mlic_rans_encode(_ptr(sym), _ptr(idx), sym.size, _ptr(cdf), cdf.shape[1], _ptr(ln), _ptr(off), cdf.shape[0], _ptr(out), cap,
                           C.byref(n))
    if r:
        raise ValueError(f"range encoder failed ({r})")
    return out[:n.value].tobytes()


class RansDecoder:
    """compressai.ans.RansDecoder: set_stream(bytes) then decode_stream(indexes, cdf, cdf_lengths, offsets) any number of times."""

    def __init__(self):
        self._d = None

    def set_stream(self, data):
        self.close()
        buf = np.frombuffer(bytes(data), dtype=np.uint8)
        self._d = _lib.lib().mlic_rans_decoder_create(_ptr(buf), buf.size)
        if not self._d:
            raise ValueError("not a range-coder stream")

    def decode_stream(self, indexes, cdf, cdf_lengths, offsets):
        idx, cdf, ln, off = _i32(indexes).reshape(-1), _i32(cdf), _i32(cdf_lengths), _i32(offsets)
        out = np.empty(idx.size, dtype=np.int32)
        r = _lib.lib().mlic_rans_decode_stream(self._d, _ptr(idx), idx.size, _ptr(cdf), cdf.shape[1], _ptr(ln), _ptr(off), cdf.shape[0],
                                               _ptr(out))
        if r:
            raise ValueError(f"range decoder failed ({r})")
        return out

    def decode_with_indexes(self, data, indexes, cdf, cdf_lengths, offsets):
        self.set_stream(data)
        return self.decode_stream(indexes, cdf, cdf_lengths, offsets)

    def close(self):
        if self._d:
            _lib.lib().mlic_rans_decoder_destroy(self._d)
            self._d = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
