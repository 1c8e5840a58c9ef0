"""Host range-ANS coder and quantised-CDF tables behind compress() / decompress() (mlic_b200/csrc/rans.cpp, mlic_b200/coder.py).
CompressAI (the reference's coder) is absent here, so these pin the restated algorithm by its own invariants: exact round trips
(escape / bypass paths included), table construction rules, and code lengths against the ideal -log2 p."""
import math

import numpy as np
import pytest
import torch

import mlic_b200
from mlic_b200 import coder


def test_pmf_to_quantized_cdf_rules():
    # plain case: proportional, monotone, ends at 2^16
    c = coder.pmf_to_quantized_cdf([0.5, 0.25, 0.125, 0.125])
    assert c.tolist() == [0, 32768, 49152, 57344, 65536]
    # zero-probability bins steal one count from the smallest bin with frequency > 1 (ryg_rans rule)
    c = coder.pmf_to_quantized_cdf([0.0, 0.7, 0.0, 0.3, 0.0])
    f = np.diff(c)
    assert c[0] == 0 and c[-1] == 65536 and (f >= 1).all() and f[0] == 1 and f[2] == 1 and f[4] == 1
    # unnormalised input is rescaled to 2^16
    assert coder.pmf_to_quantized_cdf([2.0, 2.0])[1] == 32768
    with pytest.raises(ValueError):
        coder.pmf_to_quantized_cdf([0.0, 0.0])
    with pytest.raises(ValueError):
        coder.pmf_to_quantized_cdf([0.5, float("nan")])


def test_gaussian_tables_follow_compressai_update():
    st = mlic_b200.get_scale_table()
    cdf, ln, off = coder.gaussian_tables(st)
    assert cdf.shape[0] == 64 and ln.shape == (64,) and off.shape == (64,)
    mult = -coder._norm_ppf(0.5e-9)                        # ~6.1 sigma each side
    assert 6.0 < mult < 6.2
    centers = np.ceil(st.numpy() * np.float32(mult)).astype(int)
    assert (off == -centers).all() and (ln == 2 * centers + 3).all() and cdf.shape[1] == ln.max()
    for t in (0, 17, 63):
        row = cdf[t, :ln[t]]
        assert row[0] == 0 and row[-1] == 65536 and (np.diff(row) >= 1).all()
        assert (cdf[t, ln[t]:] == 0).all()
        # symmetric pmf around the centre (up to the frequency stealing of empty bins)
        f = np.diff(row)[:-1]
        assert abs(int(f[: centers[t]].sum()) - int(f[centers[t] + 1:].sum())) <= 2 * centers[t] + 2
        # the centre bin holds erf(0.5 / (sigma sqrt 2)) of the mass
        want = math.erf(0.5 / (float(st[t]) * math.sqrt(2.0)))
        assert abs(f[centers[t]] / 65536 - want) < 2e-3 + 0.02 * want


def test_bottleneck_tables_from_parameters():
    net = mlic_b200.get_model("MLICPP_S")
    cdf, ln, off = coder.bottleneck_tables(net.entropy_bottleneck)
    assert cdf.shape == (net.N, 23) and (ln == 23).all() and (off == -10).all()       # quantiles (-10, 0, 10) at init
    assert (cdf[:, 0] == 0).all() and (cdf[:, -1] == 65536).all() and (np.diff(cdf, axis=1) >= 1).all()


@pytest.mark.parametrize("seed", [0, 1])
def test_round_trip_with_escapes_and_split_decoding(seed):
    st = mlic_b200.get_scale_table()
    cdf, ln, off = coder.gaussian_tables(st)
    rng = np.random.default_rng(seed)
    n = 50000
    idx = rng.integers(0, 64, n).astype(np.int32)
    sym = np.round(rng.standard_normal(n) * st.numpy()[idx]).astype(np.int32)
    # out-of-range values on both sides (bypass code), including ones that need 8 nibbles
    sym[::97] = rng.integers(-40000, 40000, sym[::97].size)
    sym[5], sym[6], sym[7] = 2 ** 30, -(2 ** 30), 0
    s = coder.encode_with_indexes(sym, idx, cdf, ln, off)
    assert len(s) % 4 == 0
    d = coder.RansDecoder()
    d.set_stream(s)
    cuts = [0, 1, 130, 20000, n]
    out = np.concatenate([d.decode_stream(idx[a:b], cdf, ln, off) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.array_equal(out, sym)
    assert np.array_equal(coder.RansDecoder().decode_with_indexes(s, idx, cdf, ln, off), sym)


def test_code_length_is_close_to_the_table_entropy():
    st = mlic_b200.get_scale_table()
    cdf, ln, off = coder.gaussian_tables(st)
    rng = np.random.default_rng(3)
    n = 200000
    idx = rng.integers(8, 40, n).astype(np.int32)
    sym = np.round(rng.standard_normal(n) * st.numpy()[idx]).astype(np.int32)
    v = sym - off[idx]
    inside = (v >= 0) & (v < ln[idx] - 2)
    assert inside.all()
    p = (cdf[idx, v + 1] - cdf[idx, v]) / 65536.0
    ideal_bits = float(-np.log2(p).sum())
    got_bits = 8 * len(coder.encode_with_indexes(sym, idx, cdf, ln, off))
    assert ideal_bits <= got_bits <= ideal_bits * 1.0005 + 128


def test_empty_and_errors():
    st = mlic_b200.get_scale_table()
    cdf, ln, off = coder.gaussian_tables(st)
    s = coder.encode_with_indexes([], [], cdf, ln, off)
    assert len(s) == 8                                     # the flushed 64-bit state
    assert coder.RansDecoder().decode_with_indexes(s, [], cdf, ln, off).size == 0
    with pytest.raises(ValueError):
        coder.encode_with_indexes([0, 1], [0], cdf, ln, off)
    with pytest.raises(ValueError):
        coder.encode_with_indexes([0], [64], cdf, ln, off)  # no such table
    with pytest.raises(ValueError):
        coder.RansDecoder().set_stream(b"abc")


def test_truncated_or_mismatched_stream_is_an_error_not_zeros():
    """A stream that ends before its symbols do (truncated file, wrong tables or gain on the decoder side) must fail: the decoder
    used to zero-fill past the end and return plausible symbols with rc 0."""
    st = mlic_b200.get_scale_table()
    cdf, ln, off = coder.gaussian_tables(st)
    rng = np.random.default_rng(3)
    n = 20000
    idx = rng.integers(0, 64, n).astype(np.int32)
    sym = np.round(rng.standard_normal(n) * st.numpy()[idx]).astype(np.int32)
    s = coder.encode_with_indexes(sym, idx, cdf, ln, off)
    assert np.array_equal(coder.RansDecoder().decode_with_indexes(s, idx, cdf, ln, off), sym)      # the whole stream is fine
    with pytest.raises(ValueError):
        coder.RansDecoder().decode_with_indexes(s[:len(s) // 2], idx, cdf, ln, off)                # truncated
    with pytest.raises(ValueError):                                                                # wider tables than it was coded with
        coder.RansDecoder().decode_with_indexes(s, np.full(n, 63, np.int32), cdf, ln, off)


def test_update_fills_the_reference_buffers():
    net = mlic_b200.get_model("MLICPP_S")
    assert net.gaussian_conditional._quantized_cdf.numel() == 0
    assert net.update(force=True) is True and net.update() is False
    gc, eb = net.gaussian_conditional, net.entropy_bottleneck
    assert gc.scale_table.shape == (64,) and gc._quantized_cdf.shape[0] == 64 and gc._cdf_length.shape == (64,) and gc._offset.shape == (64,)
    assert eb._quantized_cdf.shape == (net.N, 23) and eb._cdf_length.dtype == torch.int32
    # a state_dict carrying tables loads into a fresh model (buffers are resized: models/mlicpp.py:461-468)
    other = mlic_b200.get_model("MLICPP_S")
    other.load_state_dict(net.state_dict())
    assert torch.equal(other.gaussian_conditional._quantized_cdf, gc._quantized_cdf)


def _reference_encode(symbols, indexes, cdf, ln, off):
    """The published algorithm of compressai.ans with Python integers (exact division, no reciprocal, no tables): 64-bit
    rANS, 16-bit precision, 32-bit words emitted backwards, 4-bit bypass escapes -- an independent statement of what
    mlic_rans_encode must write, byte for byte."""
    L, P = 1 << 31, 16
    ops = []                                              # (start, range, bits) in coding order
    for sv, t in zip(symbols, indexes):
        mx = int(ln[t]) - 2
        v = int(sv) - int(off[t])
        raw = None
        if v < 0:
            raw, v = -2 * v - 1, mx
        elif v >= mx:
            raw, v = 2 * (v - mx), mx
        ops.append((int(cdf[t][v]), int(cdf[t][v + 1] - cdf[t][v]), P))
        if raw is not None:
            nb = 0
            while nb < 8 and (raw >> (4 * nb)) != 0:
                nb += 1
            val = nb
            while val >= 15:
                ops.append((15, 1, 4))
                val -= 15
            ops.append((val, 1, 4))
            ops.extend(((raw >> (4 * j)) & 15, 1, 4) for j in range(nb))
    x, words = L, []
    for start, rng, bits in reversed(ops):
        if x >= ((L >> bits) << 32) * rng:
            words.append(x & 0xFFFFFFFF)
            x >>= 32
        x = ((x // rng) << bits) + (x % rng) + start
    words += [x >> 32, x & 0xFFFFFFFF]
    return np.array(words[::-1], dtype="<u4").tobytes()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_encoder_bytes_equal_an_exact_integer_reference(seed):
    """The encoder divides by reciprocal multiplication and writes in place; the bytes must equal the plain-integer
    statement of the algorithm, on every frequency class: wide and narrow tables, frequency-1 tails, escapes of 1..8 nibbles."""
    st = mlic_b200.get_scale_table()
    cdf, ln, off = coder.gaussian_tables(st)
    rng = np.random.default_rng(seed)
    n = 6000
    idx = rng.integers(0, 64, n).astype(np.int32)
    sym = np.round(rng.standard_normal(n) * st.numpy()[idx] * 1.5).astype(np.int32)
    sym[::53] = rng.integers(-70000, 70000, sym[::53].size)
    sym[3], sym[4], sym[5], sym[6] = 2 ** 30, -(2 ** 30), 2 ** 31 - 1 - int(ln[idx[5]]), 0
    got = coder.encode_with_indexes(sym, idx, cdf, ln, off)
    assert got == _reference_encode(sym.tolist(), idx.tolist(), np.asarray(cdf), np.asarray(ln), np.asarray(off))
    assert np.array_equal(coder.RansDecoder().decode_with_indexes(got, idx, cdf, ln, off), sym)


def test_decoder_rejects_a_table_that_is_not_a_cdf():
    st = mlic_b200.get_scale_table()
    cdf, ln, off = coder.gaussian_tables(st)
    s = coder.encode_with_indexes([0, 1, -1], [5, 5, 5], cdf, ln, off)
    bad = np.array(cdf, dtype=np.int32, copy=True)
    bad[5, 2] = bad[5, 1] - 1                              # not monotone
    with pytest.raises(ValueError):
        coder.RansDecoder().decode_with_indexes(s, [5, 5, 5], bad, ln, off)
    bad = np.array(cdf, dtype=np.int32, copy=True)
    bad[5, int(ln[5]) - 1] = 65535                         # does not end at 2^16
    with pytest.raises(ValueError):
        coder.RansDecoder().decode_with_indexes(s, [5, 5, 5], bad, ln, off)
    assert coder.RansDecoder().decode_with_indexes(s, [5, 5, 5], cdf, ln, off).tolist() == [0, 1, -1]


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_tables_against_the_integer_reference(seed):
    """Tables of arbitrary shape (2..300 bins, Dirichlet masses incl. near-one and frequency-1 bins) through
    pmf_to_quantized_cdf: encoder bytes == integer reference, decoder inverts them, also in split calls."""
    rng = np.random.default_rng(100 + seed)
    n_tables, stride = 12, 304
    cdf = np.zeros((n_tables, stride), dtype=np.int32)
    ln = np.zeros(n_tables, dtype=np.int32)
    off = rng.integers(-150, 5, n_tables).astype(np.int32)
    for t in range(n_tables):
        bins = int(rng.integers(2, 301))
        pmf = rng.dirichlet(np.full(bins, rng.choice([0.05, 0.5, 5.0]))).astype(np.float32) + 1e-12
        if t == 0:
            pmf = np.array([1e-9, 1.0, 1e-9], dtype=np.float32)      # one bin takes 65534 of the 65536
            bins = 3
        tail = np.float32(1e-9)
        c = coder.pmf_to_quantized_cdf(np.concatenate([pmf, [tail]]))
        cdf[t, :len(c)] = c
        ln[t] = len(c)
        assert c[0] == 0 and c[-1] == 65536 and (np.diff(c) >= 1).all()
    n = 5000
    idx = rng.integers(0, n_tables, n).astype(np.int32)
    sym = (off[idx] + rng.integers(-3, ln[idx] + 2)).astype(np.int32)          # inside and just outside every table
    got = coder.encode_with_indexes(sym, idx, cdf, ln, off)
    assert got == _reference_encode(sym.tolist(), idx.tolist(), cdf, ln, off)
    d = coder.RansDecoder()
    d.set_stream(got)
    out = np.concatenate([d.decode_stream(idx[:777], cdf, ln, off), d.decode_stream(idx[777:], cdf, ln, off)])
    assert np.array_equal(out, sym)
