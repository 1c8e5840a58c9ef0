"""Bit-stream container (mlic_b200/stream.py) against bytes written by the reference's own functions
(tests/golden/stream_container.*, from oracle/make_stream_golden.py), and the file-level codec calls on the GPU."""
import io
import json
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
