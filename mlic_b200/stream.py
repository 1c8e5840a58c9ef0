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


_GAIN_BITS = 0x00800000        # third header word >= this: not a level but the float32 bit pattern of a forced gain


def _level_word(level, force):
    """Third header word of a VBR stream.  The reference writes `level` itself (testing.py:244, struct ">3I"): an integer.
    A fractional forced gain cannot be written by the reference at all (struct.pack raises on a float); here it travels as
    its float32 bit pattern (every positive normal float32 is >= 0x00800000 as an integer, far above any level), so that
    the decoder dequantises with exactly the gain the encoder used.  -> (word, the value handed to compress / decompress)"""
    if float(level) == int(level) and 0 <= int(level) < _GAIN_BITS:
        return int(level), int(level)
    if not force:
        raise ValueError(f"level {level!r} is not an index into the gain table")
    g32 = struct.unpack("<f", struct.pack("<f", float(level)))[0]
    word = struct.unpack("<I", struct.pack("<f", g32))[0]
    if not (g32 > 0 and word >= _GAIN_BITS):
        raise ValueError(f"forced gain {level!r} is not a positive normal float32")
    return word, g32


def compress_one_image_vbr(model, x, stream_path, H, W, img_name, level=0, force=False):
    """testing.py:232-247"""
    word, lv = _level_word(level, force)
    out = model.compress(x, stage=2, s=int(lv), inputscale=0 if not force else lv)
    output = os.path.join(stream_path, img_name)
    with Path(output).open("wb") as f:
        write_uints(f, (H, W, word))
        write_body(f, out["shape"], out["strings"])
    return float(os.path.getsize(output)) * 8 / (H * W), out["cost_time"]


def decompress_one_image_vbr(model, stream_path, img_name, force=False):
    """testing.py:250-262"""
    with Path(os.path.join(stream_path, img_name)).open("rb") as f:
        H, W, word = read_uints(f, 3)
        strings, shape = read_body(f)
    level = struct.unpack("<f", struct.pack("<I", word))[0] if word >= _GAIN_BITS else word
    out = model.decompress(strings, shape, s=int(level), stage=2, inputscale=0 if not force else level)
    return out["x_hat"][:, :, 0:H, 0:W], out["cost_time"]


# ---------------------------------------------------------------------------------------------------------------------
# The evaluation loop of the reference (testing.py:338-424 `test_model`, :427-520 `test_model_vbr`) without the metrics
# whose packages are absent here (MS-SSIM / LPIPS / DISTS, the DeepSpeed FLOP profile): per image pad -> compress -> file
# -> decompress -> crop, bpp from the file size, PSNR on the 8-bit images the reference saves, coder + network latencies.

def gaussian_blur3(img, sigma=0.5):
    """testing.py:264-295: normalised 3x3 Gaussian (sigma 0.5), depthwise, zero padding 1."""
    k = torch.arange(3, dtype=torch.float32) - 1.0
    g = torch.exp(-(k[:, None] ** 2 + k[None, :] ** 2) / (2.0 * sigma ** 2))
    g = (g / g.sum()).to(img.device, img.dtype)
    C = img.shape[1]
    return F.conv2d(img, g.expand(C, 1, 3, 3).contiguous(), padding=1, groups=C)


def to_uint8(x):
    """utils.py:86-87 `torch2img`: clamp to [0,1], then torchvision's float -> PIL rule (x * 255 truncated to a byte)."""
    return x.detach().clamp(0, 1).mul(255).to(torch.uint8)


def psnr_8bit(a, b):
    """metrics.py:26-33 on the 8-bit images: 20 log10(255) - 10 log10(mse)."""
    mse = float(((to_uint8(a).double() - to_uint8(b).double()) ** 2).mean())
    return 20.0 * math.log10(255.0) - 10.0 * math.log10(mse) if mse > 0 else float("inf")


class _Mean:
    def __init__(self):
        self.sum, self.n = 0.0, 0

    def update(self, v):
        self.sum += float(v)
        self.n += 1

    @property
    def avg(self):
        return self.sum / max(self.n, 1)


def _device_of(net):
    return next(net.parameters()).device


def test_model(images, net, save_dir, cons=0.100, max_blur=None, log=None):
    """testing.py:338-424.  `images`: iterable of [1,3,H,W] tensors (or dicts with "image").  As the reference does, an
    image whose stream is above `cons` bpp is blurred (3x3 Gaussian) and coded again until it fits; `max_blur` bounds that
    loop (None = unbounded, the reference's behaviour).  -> {"avg": {...}, "images": [per-image records]}."""
    dev = _device_of(net)
    os.makedirs(save_dir, exist_ok=True)
    keys = ("bpp", "psnr", "enc_time", "dec_time")
    avg = {k: _Mean() for k in keys}
    records = []
    for i, d in enumerate(images):
        ori = (d["image"] if isinstance(d, dict) else d).to(dev)
        img, blurs = ori, 0
        while True:
            padded, H, W = pad_to_64(img)
            if i == 0 and blurs == 0:                                          # the reference's warm-up pass
                compress_one_image(net, padded, save_dir, H, W, str(i))
            net.update_resolutions(16, 16)                                     # "avoid resolution leakage" (:381,385)
            bpp, enc = compress_one_image(net, padded, save_dir, H, W, str(i))
            net.update_resolutions(16, 16)
            x_hat, dec = decompress_one_image(net, save_dir, str(i))
            if bpp <= cons or (max_blur is not None and blurs >= max_blur):
                break
            img, blurs = gaussian_blur3(img), blurs + 1
        rec = {"bpp": bpp, "psnr": psnr_8bit(x_hat, ori), "enc_time": enc, "dec_time": dec, "blurs": blurs}
        for k in keys:
            avg[k].update(rec[k])
        records.append(rec)
        if log is not None:
            log(f"Image[{i}] | Bpp: {bpp:.2f} | PSNR: {rec['psnr']:.4f} | Encoding Latency: {enc:.4f} | Decoding Latency: {dec:.4f}")
    return {"avg": {k: avg[k].avg for k in keys}, "images": records}


def test_model_vbr(images, net, save_dir, custom_scales=None, log=None):
    """testing.py:427-520: every image at every level (or at the given gains, `force`).  The reference reads a
    non-existent `net.Gains` here and its stage-2 decompress() raises; this loop runs the working calls.
    -> {level: {"avg": {...}, "images": [...]}}."""
    dev = _device_of(net)
    os.makedirs(save_dir, exist_ok=True)
    force = custom_scales is not None
    levels = list(custom_scales) if force else list(range(len(net.lmbda)))
    keys = ("bpp", "psnr", "enc_time", "dec_time")
    out = {lv: {"avg": {k: _Mean() for k in keys}, "images": []} for lv in levels}
    for i, d in enumerate(images):
        ori = (d["image"] if isinstance(d, dict) else d).to(dev)
        padded, H, W = pad_to_64(ori)
        for lv in levels:
            name = f"{i:03}_lv{lv:02}" if not force else f"{i:03}_g{lv:g}"
            net.update_resolutions(16, 16)
            bpp, enc = compress_one_image_vbr(net, padded, save_dir, H, W, name, level=lv, force=force)
            net.update_resolutions(16, 16)
            x_hat, dec = decompress_one_image_vbr(net, save_dir, name, force=force)
            rec = {"bpp": bpp, "psnr": psnr_8bit(x_hat, ori), "enc_time": enc, "dec_time": dec}
            for k in keys:
                out[lv]["avg"][k].update(rec[k])
            out[lv]["images"].append(rec)
            if log is not None:
                log(f"Image[{i}] level {lv} | Bpp: {bpp:.2f} | PSNR: {rec['psnr']:.4f}")
    return {lv: {"avg": {k: v["avg"][k].avg for k in keys}, "images": v["images"]} for lv, v in out.items()}
