"""Bit-stream container and the per-image codec calls of the reference's evaluation path.

Byte-for-byte the reference's container: big-endian uint32 fields and raw strings
  [H, W (, level)] [shape_h, shape_w, n_strings] n_strings x ([len] bytes)
(MLIC++/utils/utils.py:28-77 `write_uints / read_uints / write_bytes / read_bytes / write_body / read_body`;
MLIC++/utils/testing.py:203-262 `compress_one_image[_vbr] / decompress_one_image[_vbr]`; padding to multiples of 64 with
zeros and cropping back: testing.py:130-141).  As in the reference, `write_body` stores the FIRST string of each list:
one image per file.  Pinned against the reference's own functions by tests/golden/stream_container.* (written by
oracle/make_stream_golden.py from the unmodified reference source)."""
import math
import os
import struct
from pathlib import Path

import torch
import torch.nn.functional as F


def write_uints(fd, values, fmt=">{:d}I"):
    fd.write(struct.pack(fmt.format(len(values)), *values))
    return len(values) * 4


def read_uints(fd, n, fmt=">{:d}I"):
    return struct.unpack(fmt.format(n), fd.read(n * struct.calcsize("I")))


def write_bytes(fd, values, fmt=">{:d}s"):
    if len(values) == 0:
        return 0
    fd.write(struct.pack(fmt.format(len(values)), values))
    return len(values)


def read_bytes(fd, n, fmt=">{:d}s"):
    return struct.unpack(fmt.format(n), fd.read(n * struct.calcsize("s")))[0]


def write_body(fd, shape, out_strings):
    cnt = write_uints(fd, (shape[0], shape[1], len(out_strings)))
    for s in out_strings:
        cnt += write_uints(fd, (len(s[0]),))
        cnt += write_bytes(fd, s[0])
    return cnt


def read_body(fd):
    shape = read_uints(fd, 2)
    n_strings = read_uints(fd, 1)[0]
    lstrings = []
    for _ in range(n_strings):
        lstrings.append([read_bytes(fd, read_uints(fd, 1)[0])])
    return lstrings, shape


def pad_to_64(img):
    """testing.py:130-137: zero padding on the right / bottom up to the next multiple of 64 -> (padded, H, W)."""
    H, W = img.shape[-2:]
    pad_h = 0 if H % 64 == 0 else 64 * (H // 64 + 1) - H
    pad_w = 0 if W % 64 == 0 else 64 * (W // 64 + 1) - W
    return F.pad(img, (0, pad_w, 0, pad_h), mode="constant", value=0), H, W


def psnr(a, b):
    mse = float(((a.double().clamp(0, 1) - b.double()) ** 2).mean())
    return 10.0 * math.log10(1.0 / mse) if mse > 0 else float("inf")


def compress_one_image(model, x, stream_path, H, W, img_name):
    """testing.py:203-215 -> (bpp of the file over the unpadded H x W, cost_time)."""
    out = model.compress(x)
    output = os.path.join(stream_path, img_name)
    with Path(output).open("wb") as f:
        write_uints(f, (H, W))
        write_body(f, out["shape"], out["strings"])
    return float(os.path.getsize(output)) * 8 / (H * W), out["cost_time"]


def decompress_one_image(model, stream_path, img_name):
    """testing.py:218-230 -> (x_hat cropped to the original size, cost_time)."""
    with Path(os.path.join(stream_path, img_name)).open("rb") as f:
        original_size = read_uints(f, 2)
        strings, shape = read_body(f)
    out = model.decompress(strings, shape)
    return out["x_hat"][:, :, 0:original_size[0], 0:original_size[1]], out["cost_time"]


def compress_one_image_vbr(model, x, stream_path, H, W, img_name, level=0, force=False):
    """testing.py:232-247"""
    out = model.compress(x, stage=2, s=int(level), inputscale=0 if not force else level)
    output = os.path.join(stream_path, img_name)
    with Path(output).open("wb") as f:
        write_uints(f, (H, W, int(level)))
        write_body(f, out["shape"], out["strings"])
    return float(os.path.getsize(output)) * 8 / (H * W), out["cost_time"]


def decompress_one_image_vbr(model, stream_path, img_name, force=False):
    """testing.py:250-262"""
    with Path(os.path.join(stream_path, img_name)).open("rb") as f:
        H, W, level = read_uints(f, 3)
        strings, shape = read_body(f)
    out = model.decompress(strings, shape, s=int(level), stage=2, inputscale=0 if not force else level)
    return out["x_hat"][:, :, 0:H, 0:W], out["cost_time"]
