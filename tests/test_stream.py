"""Bit-stream container (mlic_b200/stream.py) against bytes written by the reference's own functions
(tests/golden/stream_container.*, from oracle/make_stream_golden.py), and the file-level codec calls on the GPU."""
import io
import json
import math
import os

import pytest
import torch

from conftest import GOLDEN
from mlic_b200 import stream


def _golden():
    meta = json.load(open(os.path.join(GOLDEN, "stream_container.json")))
    blob = open(os.path.join(GOLDEN, "stream_container.bin"), "rb").read()
    return meta, blob


@pytest.mark.parametrize("case", ["plain", "vbr"])
def test_container_bytes_equal_the_reference(case):
    meta, blob = _golden()
    m = meta[case]
    want = blob[m["offset"]:m["offset"] + m["length"]]
    strings = [[bytes.fromhex(s[0])] for s in m["strings"]]
    f = io.BytesIO()
    stream.write_uints(f, tuple(m["header"]))
    assert stream.write_body(f, m["shape"], strings) == m["body_bytes"]
    assert f.getvalue() == want
    g = io.BytesIO(want)
    assert list(stream.read_uints(g, len(m["header"]))) == m["header"]
    got, shape = stream.read_body(g)
    assert got == strings and list(shape) == m["shape"] and g.read() == b""


def test_empty_string_and_padding_rules():
    f = io.BytesIO()
    stream.write_body(f, (1, 1), [[b""], [b"ab"]])
    got, shape = stream.read_body(io.BytesIO(f.getvalue()))
    assert got == [[b""], [b"ab"]] and shape == (1, 1)
    x = torch.rand(1, 3, 100, 128)
    p, H, W = stream.pad_to_64(x)
    assert (H, W) == (100, 128) and p.shape == (1, 3, 128, 128) and float(p[:, :, 100:].abs().max()) == 0
    assert stream.pad_to_64(torch.rand(1, 3, 64, 192))[0].shape == (1, 3, 64, 192)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["MLICPP_S", "MLICPP_S_VBR"])
def test_file_round_trip_on_the_engine(name, tmp_path):
    import mlic_b200
    from oracle import weights
    net = mlic_b200.get_model(name)
    net.load_state_dict(weights.seeded_state_dict(net.state_dict(), 1234, y_gain=16.0, sigma_spread=6.0))
    net.update(force=True)
    net = net.to("cuda").set_precision("fp32")
    img = weights.synthetic_image(1, 128, 192, seed=41)[:, :, :100, :150].cuda()      # not a multiple of 64
    x, H, W = stream.pad_to_64(img)
    if name.endswith("VBR"):
        bpp, _ = stream.compress_one_image_vbr(net, x, str(tmp_path), H, W, "a.bin", level=3)
        x_hat, _ = stream.decompress_one_image_vbr(net, str(tmp_path), "a.bin")
        ref = net.compress(x, stage=2, s=3)["x_hat"]
    else:
        bpp, _ = stream.compress_one_image(net, x, str(tmp_path), H, W, "a.bin")
        x_hat, _ = stream.decompress_one_image(net, str(tmp_path), "a.bin")
        ref = net.compress(x)["x_hat"]
    assert x_hat.shape == (1, 3, 100, 150) and torch.equal(x_hat, ref[:, :, :100, :150])
    assert bpp == os.path.getsize(tmp_path / "a.bin") * 8 / (100 * 150) and bpp > 0
    assert stream.psnr(img, img) == float("inf") and 0 < stream.psnr(x_hat, img) < 60


def test_eval_helpers_on_cpu():
    """testing.py:264-295 Gaussian (sigma 0.5, normalised, zero padding) and the 8-bit PSNR of metrics.py:26-33."""
    one = torch.ones(1, 3, 8, 8)
    b = stream.gaussian_blur3(one)
    e, c = math.exp(-2.0), math.exp(-4.0)
    tot = 1 + 4 * e + 4 * c
    assert b.shape == one.shape and float((b[:, :, 1:-1, 1:-1] - 1).abs().max()) < 1e-6
    assert float(b[0, 0, 0, 0]) == pytest.approx((1 + 2 * e + c) / tot, rel=1e-6)       # corner: zero padding
    delta = torch.zeros(1, 3, 5, 5)
    delta[:, :, 2, 2] = 1
    k = stream.gaussian_blur3(delta)[0, 1, 1:4, 1:4]
    assert float(k[1, 1]) == pytest.approx(1 / tot, rel=1e-6) and float(k[0, 1]) == pytest.approx(e / tot, rel=1e-6)
    a = torch.full((1, 3, 4, 4), 0.5)
    assert stream.to_uint8(a).unique().tolist() == [127]                                # truncation, not rounding
    assert stream.psnr_8bit(a, a) == float("inf")
    assert stream.psnr_8bit(a, a + 2.0 / 255) == pytest.approx(20 * math.log10(255) - 10 * math.log10(4.0), abs=1e-9)
    assert stream.psnr_8bit(torch.full((1, 3, 2, 2), 1.7), torch.ones(1, 3, 2, 2)) == float("inf")   # clamp first


@pytest.mark.gpu
def test_eval_loops_on_the_engine(tmp_path):
    """stream.test_model / test_model_vbr (testing.py:338-520 without the perceptual metrics)."""
    import mlic_b200
    from oracle import weights

    def build(name):
        net = mlic_b200.get_model(name)
        net.load_state_dict(weights.seeded_state_dict(net.state_dict(), 1234, y_gain=16.0, sigma_spread=6.0))
        net.update(force=True)
        return net.to("cuda").set_precision("fp32")

    imgs = [weights.synthetic_image(1, 128, 192, seed=51)[:, :, :100, :150], weights.synthetic_image(1, 64, 128, seed=52)]
    net = build("MLICPP_S")
    lines = []
    r = stream.test_model(imgs, net, str(tmp_path / "a"), cons=1e9, log=lines.append)
    assert len(r["images"]) == 2 and len(lines) == 2 and all(rec["blurs"] == 0 for rec in r["images"])
    assert r["images"][0]["bpp"] == os.path.getsize(tmp_path / "a" / "0") * 8 / (100 * 150)
    assert r["avg"]["bpp"] == pytest.approx(sum(rec["bpp"] for rec in r["images"]) / 2)
    assert all(5 < rec["psnr"] < 60 and rec["enc_time"] > 0 and rec["dec_time"] > 0 for rec in r["images"])
    r2 = stream.test_model(imgs[:1], net, str(tmp_path / "b"), cons=1e-6, max_blur=2)            # the blur-until-it-fits loop
    assert r2["images"][0]["blurs"] == 2 and r2["images"][0]["bpp"] > 0
    vbr = build("MLICPP_S_VBR")
    rv = stream.test_model_vbr(imgs[1:], vbr, str(tmp_path / "v"))
    assert sorted(rv) == list(range(6)) and all(len(v["images"]) == 1 and v["avg"]["bpp"] > 0 for v in rv.values())
    assert os.path.exists(tmp_path / "v" / "000_lv05")
    rg = stream.test_model_vbr(imgs[1:], vbr, str(tmp_path / "g"), custom_scales=[0.25])          # forced gain
    assert list(rg) == [0.25] and rg[0.25]["avg"]["psnr"] > 5
    # the forced fractional gain reaches the decoder exactly (float32 bits in the third header word), so the decoded image is the
    # encoder's own reconstruction at that gain, not one dequantised with int(0.25) = 0
    with open(tmp_path / "g" / "000_g0.25", "rb") as f:
        assert stream.read_uints(f, 3)[2] == 0x3E800000
    padded, H, W = stream.pad_to_64(imgs[1])
    enc = vbr.compress(padded.cuda(), stage=2, s=0, inputscale=0.25)
    dec = vbr.decompress(enc["strings"], enc["shape"], stage=2, s=0, inputscale=0.25)
    x_file, _ = stream.decompress_one_image_vbr(vbr, str(tmp_path / "g"), "000_g0.25", force=True)
    assert torch.equal(x_file.cpu(), dec["x_hat"][:, :, :H, :W].cpu())


def test_vbr_header_word_carries_integer_levels_and_fractional_gains():
    import struct
    assert stream._level_word(3, False) == (3, 3) and stream._level_word(2.0, True) == (2, 2)
    word, g = stream._level_word(0.3, True)
    assert word >= stream._GAIN_BITS and struct.unpack("<f", struct.pack("<I", word))[0] == g == torch.tensor(0.3).item()
    with pytest.raises(ValueError):
        stream._level_word(0.3, False)           # not an index into the gain table
    with pytest.raises(ValueError):
        stream._level_word(-1.5, True)
