"""TEST INFRASTRUCTURE ONLY -- deterministic, order-independent weights and inputs.

The reference initialises with PyTorch defaults in construction order; to make
weights reproducible on the GPU box (where /root/reference does not exist) and
identical between the reference model, the oracle and the CUDA engine, every
floating-point *parameter* is overwritten by a value that depends only on
(seed, tensor name, shape).  Magnitudes follow the PyTorch/CompressAI defaults
(U(+-1/sqrt(fan_in)) for conv/linear, GDN beta~1 / gamma~0.1*I, EntropyBottleneck
init) with extra jitter so that no term is trivially zero (GDN off-diagonals,
EB factors/medians, relative-position table).
"""
import json
import math
import os
import zlib

import torch

_SKIP_SUFFIX = ("relative_position_index", "target", "scale_bound", "scale_table", "_offset",
                "_quantized_cdf", "_cdf_length", "pedestal", "bound", "attn_mask")


def _gen(seed, name):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def _u(shape, lo, hi, g):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def fill_value(name, shape, seed, sd_shapes=None):
    """Return the deterministic fp32 tensor for parameter `name`, or None to leave it alone."""
    leaf = name.rsplit(".", 1)[-1]
    if leaf in _SKIP_SUFFIX or name.endswith(_SKIP_SUFFIX):
        return None
    g = _gen(seed, name)
    shape = tuple(shape)
    ped = 2.0 ** -36
    if leaf == "beta":                     # GDN beta (stored re-parametrised)
        return torch.sqrt(1.0 + _u(shape, 0.0, 0.3, g) + ped)
    if leaf == "gamma":                    # GDN gamma [C,C]
        C = shape[0]
        return torch.sqrt(0.1 * torch.eye(C) + _u(shape, 0.0, 0.004, g) + ped)
    if leaf == "relative_position_table":
        return _u(shape, -0.04, 0.04, g)
    if ".matrices." in name:               # EntropyBottleneck
        fo = shape[1]
        scale = 10.0 ** (1.0 / 5.0)
        init = math.log(math.expm1(1.0 / scale / fo))
        return init + _u(shape, -0.2, 0.2, g)
    if ".biases." in name:
        return _u(shape, -0.5, 0.5, g)
    if ".factors." in name:
        return _u(shape, -0.3, 0.3, g)
    if leaf == "quantiles":
        q = torch.empty(shape, dtype=torch.float32)
        q[:, :, 0] = -10.0
        q[:, :, 2] = 10.0
        q[:, :, 1] = _u(shape[:2], -0.4, 0.4, g)
        return q
    if leaf == "Gain":
        return None
    if "norm1." in name or "norm2." in name:  # LayerNorm affine
        return 1.0 + _u(shape, -0.1, 0.1, g) if leaf == "weight" else _u(shape, -0.1, 0.1, g)
    if leaf == "weight" and len(shape) >= 2:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        b = 1.0 / math.sqrt(fan_in)
        return _u(shape, -b, b, g)
    if leaf == "bias" and len(shape) == 1:
        fan_in = None
        if sd_shapes is not None:
            w = sd_shapes.get(name[:-4] + "weight")
            if w is not None and len(w) >= 2:
                fan_in = 1
                for d in w[1:]:
                    fan_in *= d
        b = 1.0 / math.sqrt(fan_in) if fan_in else 0.05
        return _u(shape, -b, b, g)
    return _u(shape, -0.05, 0.05, g)


def seeded_state_dict(state_dict, seed=1234, y_gain=1.0, sigma_spread=0.0):
    """Overwrite every float parameter of `state_dict` (name -> tensor) deterministically.

    y_gain scales g_a's last pointwise conv (weight and bias) so that |y| is large enough for
    non-trivial symbols (random-init latents otherwise round to 0).
    sigma_spread > 0 widens the scale half of every EntropyParameters output layer (rows [:C] of
    fusion.6: weight x20, bias += U(0, sigma_spread)) so CDF indexes cover more than entry 0.
    """
    shapes = {k: tuple(v.shape) for k, v in state_dict.items()}
    out = {}
    for k, v in state_dict.items():
        nv = None
        if torch.is_floating_point(v):
            nv = fill_value(k, v.shape, seed, shapes)
        out[k] = v.clone() if nv is None else nv.to(v.dtype)
    if y_gain != 1.0:
        for k in list(out):
            if k.startswith("g_a.analysis_transform.6."):
                out[k] = out[k] * float(y_gain) if k.endswith(("point_conv.weight", "point_conv.bias")) or \
                    k in ("g_a.analysis_transform.6.weight", "g_a.analysis_transform.6.bias") else out[k]
    if sigma_spread > 0.0:
        for k in list(out):
            if k.startswith("entropy_parameters_") and ".fusion.6." in k:
                C = out[k].shape[0] // 2
                t = out[k].clone()
                if k.endswith("weight"):
                    t[:C] *= 20.0
                else:
                    t[:C] += _u((C,), 0.0, float(sigma_spread), _gen(seed, k + "#spread"))
                out[k] = t
    return out


def template_state_dict(name):
    """state_dict template of reference model `name` (keys, shapes and the small buffers weights.py leaves alone), read from
    oracle/shapes/<name>.npz -- dumped from the unmodified reference by oracle/make_shapes.py.  Float parameters are zeros."""
    import numpy as np
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shapes", f"{name}.npz")
    z = np.load(path)
    meta = json.loads(bytes(z["__meta__"]).decode())
    out = {}
    for k, m in meta.items():
        dt = getattr(torch, m["dtype"])
        out[k] = torch.from_numpy(z[k]).to(dt).reshape(m["shape"]) if m.get("stored") else torch.zeros(m["shape"], dtype=dt)
    return out


def reference_state_dict(name, seed=1234, y_gain=1.0, sigma_spread=0.0):
    """The seeded weights of `name` built from the committed template alone (no model object, no product package)."""
    return seeded_state_dict(template_state_dict(name), seed, y_gain=y_gain, sigma_spread=sigma_spread)


def synthetic_image(B, H, W, seed=2024, kind="smooth"):
    """Deterministic fp32 NCHW image batch in [0,1].

    kind="rand": iid uniform; kind="smooth": bilinear-upsampled 1/16-res noise + 10 % iid noise.
    """
    outs = []
    for i in range(B):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed + i)
        if kind == "rand":
            outs.append(torch.rand(1, 3, H, W, generator=g))
        else:
            lo = torch.rand(1, 3, max(H // 16, 2), max(W // 16, 2), generator=g)
            up = torch.nn.functional.interpolate(lo, size=(H, W), mode="bilinear", align_corners=False)
            outs.append((0.9 * up + 0.1 * torch.rand(1, 3, H, W, generator=g)).clamp(0, 1))
    return torch.cat(outs, 0).contiguous()
