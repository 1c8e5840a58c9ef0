"""Drop-in model API of the reference (MLIC++/models/model_loader.py:4-18): `get_model(name)` returns an
nn.Module with the reference's state_dict (same dotted names and shapes) whose forward / compress /
net_decoder_forward run on the B200 engine (libmlic_b200.so, hand-written sm_100a kernels) through the
C ABI in include/mlic_b200.h.  There is no PyTorch or CPU compute path behind these methods.

Reference methods mirrored (paths relative to /root/reference/MLIC++):
  forward              models/mlicpp.py:79-185   (VBR: models/mlicpp_vbr.py:137-336)
  compress             models/mlicpp.py:199-290  (network walk; the rANS coder stays a host component)
  net_decoder_forward  models/mlicpp.py:380-459
  update / load_state_dict / update_resolutions   models/mlicpp.py:187-197,461-475
"""
import ctypes as C
import math
import time
import types

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .params import KIND_CODE, MODEL_TABLE, SDVBR_GAINS, SDVBR_LAMBDAS, VBR_GAINS, VBR_LAMBDAS, build_entries


def model_config(name):
    """config/config.py:19-62 (context_window / act are carried for API parity; the engine fixes them)."""
    cfg = MODEL_TABLE[name]
    return types.SimpleNamespace(N=cfg["N"], M=cfg["M"], slice_num=cfg["slice_num"], context_window=5, act=nn.GELU)


def get_scale_table(min=0.11, max=256, levels=64):  # noqa: A002  (utils/func.py:16-19)
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def _relpos_index(ws=5):
    """modules/layers/attention.py:28-38"""
    ch, cw = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    co = torch.stack([ch, cw]).flatten(1)
    rel = (co[:, :, None] - co[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def _init_tensor(entry):
    kind, arg = entry.init
    shape = entry.shape
    dt = {"float32": torch.float32, "int32": torch.int32, "int64": torch.int64}[entry.dtype]
    if kind == "uniform_fan":
        b = 1.0 / math.sqrt(arg) if arg else 0.05
        return (torch.rand(shape) * 2 - 1) * b
    if kind == "uniform":
        return (torch.rand(shape) * 2 - 1) * arg
    if kind == "const":
        return torch.full(shape, float(arg), dtype=dt)
    if kind == "gdn_gamma":
        return torch.sqrt(0.1 * torch.eye(shape[0]) + 2.0 ** -36)
    if kind == "eb_quantiles":
        return torch.tensor([-10.0, 0.0, 10.0]).repeat(shape[0], 1, 1)
    if kind == "eb_target":
        t = math.log(2 / 1e-9 - 1)
        return torch.tensor([-t, 0.0, t])
    if kind == "eb_matrix":
        scale = 10.0 ** (1.0 / 5.0)
        return torch.full(shape, math.log(math.expm1(1.0 / scale / arg)))
    if kind == "trunc_normal":
        return nn.init.trunc_normal_(torch.zeros(shape), std=arg)
    if kind == "relpos_index":
        return _relpos_index(5)
    if kind == "vbr_gain":
        return torch.tensor(SDVBR_GAINS if shape[0] == len(SDVBR_GAINS) else VBR_GAINS, dtype=torch.float32)
    if kind == "empty":
        return torch.zeros(shape, dtype=dt)
    raise ValueError(kind)


class _Tree(nn.Module):
    """Generic container: gives the parameter tree the reference's attribute / index structure."""

    def __getitem__(self, i):
        return self._modules[str(i)]

    def __len__(self):
        return len(self._modules)


class MLICPlusPlus(nn.Module):
    """B200 engine behind the reference's MLICPlusPlus interface (models/mlicpp.py:12-76)."""

    KIND = "base"

    def __init__(self, config, name=None, _vr_entbttlnck=False, **kwargs):
        super().__init__()
        self.N, self.M = int(config.N), int(config.M)
        self.slice_num = int(config.slice_num)
        self.context_window = getattr(config, "context_window", 5)
        self.slice_ch = self.M // self.slice_num
        assert self.slice_ch * self.slice_num == self.M          # models/mlicpp.py:21
        self.model_name = name or self._name_for(config)
        self.precision = "bf16"          # "bf16" fast mode | "fp32" validation mode
        self.tensor_cores = True
        self.fuse = True                 # bf16 mode: depthwise 3x3 / x^2 produced inside the GEMM kernel
        for key, ent in build_entries(self.model_name, bool(_vr_entbttlnck)).items():
            self._place(key, _init_tensor(ent), ent.is_param)
        # attributes the reference exposes on sub-modules
        for i in range(self.slice_num):
            lc = self.local_context[i]
            lc.attn_mask, lc.H, lc.W = None, -1, -1
        self._engine = None
        self._engine_sig = None
        self._sig_tensors = None
        self._ws = {}
        self._profile = False
        self._trace = False
        self.last_launch_count = 0

    def _name_for(self, config):
        for nm, cfg in MODEL_TABLE.items():
            if (cfg["N"], cfg["M"], cfg["slice_num"], cfg["kind"]) == (self.N, self.M, self.slice_num, self.KIND):
                return nm
        raise ValueError("no registered MLIC++ configuration matches this config")

    def _place(self, key, tensor, is_param):
        parts = key.split(".")
        mod = self
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Tree())
            mod = mod._modules[p]
        if is_param:
            mod.register_parameter(parts[-1], nn.Parameter(tensor))
        else:
            mod.register_buffer(parts[-1], tensor)

    # ------------------------------------------------------------------ state handling
    def load_state_dict(self, state_dict, strict=True):
        """models/mlicpp.py:461-468: CDF buffers are resized to the checkpoint's before loading."""
        own = dict(self.named_buffers())
        for k in ("_quantized_cdf", "_offset", "_cdf_length", "scale_table"):
            for pre in ("gaussian_conditional.", "entropy_bottleneck."):
                full = pre + k
                if full in state_dict and full in own and own[full].shape != state_dict[full].shape:
                    own[full].resize_(state_dict[full].shape)
        out = super().load_state_dict(state_dict, strict=strict)
        self.invalidate_engine()
        return out

    def update(self, scale_table=None, force=False):
        """models/mlicpp.py:470-475: gaussian_conditional.update_scale_table(scale_table) and CompressionModel.update():
        fills `scale_table` and the quantised CDF tables (`_quantized_cdf`, `_cdf_length`, `_offset`) of both entropy
        models, which compress() / decompress() hand to the range coder (mlic_b200/coder.py)."""
        from . import coder
        if scale_table is None:
            scale_table = get_scale_table()
        gc, eb = self.gaussian_conditional, self.entropy_bottleneck
        if not force and gc.scale_table.numel() == len(scale_table) and gc._offset.numel() > 0 and eb._offset.numel() > 0:
            return False
        gc.scale_table.resize_(len(scale_table))
        gc.scale_table.copy_(torch.as_tensor(scale_table, dtype=torch.float32))
        for mod, tabs in ((gc, coder.gaussian_tables(gc.scale_table)), (eb, coder.bottleneck_tables(eb))):
            for name, arr in zip(("_quantized_cdf", "_cdf_length", "_offset"), tabs):
                buf = getattr(mod, name)
                t = torch.from_numpy(arr)
                buf.resize_(t.shape)
                buf.copy_(t)
        self.invalidate_engine()
        return True

    def update_resolutions(self, H, W, device=None):
        """models/mlicpp.py:187-197.  The engine evaluates the checkerboard window mask analytically, so there is
        no [L,25,25] tensor to rebuild; the cached grid size is recorded for API parity."""
        for i in range(self.slice_num):
            self.local_context[i].H, self.local_context[i].W = H, W

    def aux_loss(self):
        """CompressAI CompressionModel.aux_loss: sum |logits_cumulative(quantiles) - target| (training-side helper,
        plain torch on the parameters; not part of the forward hot path)."""
        eb = self.entropy_bottleneck
        logits = eb.quantiles
        for i in range(5):
            logits = torch.matmul(nn.functional.softplus(getattr(eb.matrices, str(i))), logits) + getattr(eb.biases, str(i))
            if i < 4:
                logits = logits + torch.tanh(getattr(eb.factors, str(i))) * torch.tanh(logits)
        return torch.abs(logits - eb.target).sum()

    def set_profile(self, on):
        """Bracket every tcgen05 GEMM launch with CUDA events (bench.py's live roofline measurement)."""
        self._profile = bool(on)
        return self

    def profile_read(self, reset=True):
        """-> (summed ms, summed algorithmic FLOPs, launches) of the tcgen05 GEMM launches since the last reset."""
        if self._engine is None:
            return 0.0, 0.0, 0
        out = (C.c_double * 3)()
        _lib.check(_lib.lib().mlic_profile_read(self._engine, out, 1 if reset else 0))
        return float(out[0]), float(out[1]), int(out[2])

    def profile_read_top(self, reset=True):
        """-> (summed ms, algorithmic FLOPs per launch, launches) of the heaviest tcgen05 GEMM shape since the last reset."""
        if self._engine is None:
            return 0.0, 0.0, 0
        out = (C.c_double * 3)()
        _lib.check(_lib.lib().mlic_profile_read_top(self._engine, out, 1 if reset else 0))
        return float(out[0]), float(out[1]), int(out[2])

    def trace_dump(self, path):
        """With `_trace` set, writes "label<TAB>microseconds" for every launch since the last dump (development aid)."""
        _lib.check(_lib.lib().mlic_trace_dump(self._engine, str(path).encode()))

    def set_precision(self, precision):
        """"bf16" (fast mode), "fp32" (validation mode), or one precision per stage as a tuple (g_a, entropy model, g_s) --
        "mixed" is ("fp32", "fp32", "bf16"): everything that decides a symbol (g_a -> y, h_a / h_s, the slice loop -> mu, sigma)
        runs in the bit-exact fp32 mode, the synthesis transform (64 % of the FLOPs) on the tensor cores.  The stages are then
        separate engine calls joined by the fp32 `y` / `y_hat` tensors (device inputs only; host tensors are uploaded first)."""
        if precision == "mixed":
            precision = ("fp32", "fp32", "bf16")
        if isinstance(precision, (tuple, list)):
            precision = tuple(precision)
            if len(precision) != 3 or any(q not in ("bf16", "fp32") for q in precision):
                raise ValueError("per-stage precision is a 3-tuple of 'bf16' / 'fp32' (g_a, entropy model, g_s)")
            if len(set(precision)) == 1:
                precision = precision[0]
        elif precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16', 'fp32', 'mixed' or a 3-tuple of 'bf16' / 'fp32'")
        self.precision = precision
        return self

    def _run_staged(self, mode, x, B, H, W, gain, want, zq=1.0):
        """A full call with per-stage precisions: consecutive stages of equal precision share one engine call."""
        if H % 64 or W % 64:
            raise ValueError("H and W must be multiples of 64 (the reference pads, utils/testing.py:130-137)")
        if not torch.cuda.is_available():
            raise _lib.MlicError("mlic_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        pa, pe, ps = self.precision
        dev = x.device if (x is not None and x.is_cuda) else torch.device("cuda", torch.cuda.current_device())
        host = x is not None and not x.is_cuda
        if host:
            x = x.to(dev, non_blocking=True)
        keep = self.precision
        out = {}
        try:
            y = None
            if mode != _lib.MODE_DECODER and pa != pe:
                self.precision = pa
                o = self._run(mode, x, B, H, W, 0.0, ("y",), stages=1)
                y = o["y"]
                if "y" in want:
                    out["y"] = y
            self.precision = pe
            first = 2 if (y is not None or mode == _lib.MODE_DECODER) else 3
            if pe == ps:
                o = self._run(mode, None if y is not None else x, B, H, W, gain, tuple(want), stages=first | 4, y=y, zq=zq)
                out.update(o)
            else:
                w2 = tuple(set(want) | {"y_hat"})
                o = self._run(mode, None if first == 2 else x, B, H, W, gain, w2, stages=first, y=y, zq=zq)
                y_hat = o["y_hat"]
                out.update({k: v for k, v in o.items() if k != "y_hat" or "y_hat" in want})
                self.precision = ps
                out.update(self._run(_lib.MODE_FORWARD, None, B, H, W, 0.0, (), stages=4, y_hat=y_hat))
        finally:
            self.precision = keep
        if host:
            out = {k: v.cpu() for k, v in out.items()}
        return out

    # ------------------------------------------------------------------ engine plumbing
    def _signature(self):
        """(data_ptr, version) of every state tensor: in-place edits, loads and device moves re-pack the engine weights.
        The tensor list itself is cached (walking the module tree costs ~3 ms per call for MLICPP_L); it is dropped by
        load_state_dict / update / .to() -- after replacing a Parameter OBJECT by hand, call invalidate_engine()."""
        if self._sig_tensors is None:
            self._sig_tensors = list(self.state_dict(keep_vars=True).values())
        return tuple((v.data_ptr(), v._version) for v in self._sig_tensors)

    def invalidate_engine(self):
        self._sig_tensors = None
        self._engine_sig = None

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate_engine()
        return out

    def _sync_engine(self, device):
        if not torch.cuda.is_available():
            raise _lib.MlicError("mlic_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        L = _lib.lib()
        sig = (str(device), self._signature())
        if self._engine is not None and sig == self._engine_sig:
            return L
        if self._engine is None:
            h = C.c_void_p()
            _lib.check(L.mlic_engine_create(self.N, self.M, self.slice_num, KIND_CODE[self.KIND], C.byref(h)))
            self._engine = h
        with torch.cuda.device(device):
            for k, v in self.state_dict().items():
                t = v.detach().to("cpu", torch.float32).contiguous()
                shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
                _lib.check(L.mlic_engine_set_param(self._engine, k.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))
            _lib.check(L.mlic_engine_finalize(self._engine))
        self._engine_sig = sig
        return L

    def __del__(self):
        try:
            if getattr(self, "_engine", None) is not None:
                _lib.lib().mlic_engine_destroy(self._engine)
                self._engine = None
        except Exception:
            pass

    def _gain(self, stage, s, inputscale, absolute=False):
        return 0.0

    def _run(self, mode, x, B, H, W, gain=0.0, want=(), stages=7, y=None, y_hat=None, zq=1.0):
        """One engine call.  x: CUDA tensor (device path) or CPU tensor (host path through mlic_run_host).
        stages: bit mask of include/mlic_b200.h option "stages" (1 g_a | 2 entropy model | 4 g_s); `y` / `y_hat` are the
        device INPUTS of the calls that start after g_a / at g_s (row-band sharding, mlic_b200/dist.py)."""
        if isinstance(self.precision, tuple):
            if stages != 7:
                raise _lib.MlicError("stage-subset calls take one precision (set_precision('bf16' | 'fp32'))")
            return self._run_staged(mode, x, B, H, W, gain, want, zq)
        hmul = 64 if stages & 2 else 16
        if H % hmul or W % 64:
            raise ValueError("H and W must be multiples of 64 (the reference pads, utils/testing.py:130-137)")
        if not torch.cuda.is_available():
            raise _lib.MlicError("mlic_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        host = x is not None and not x.is_cuda
        src = x if x is not None else (y if y is not None else y_hat)
        dev = torch.device("cuda", torch.cuda.current_device()) if (host or src is None) else src.device
        if stages != 7 and host:
            raise _lib.MlicError("stage subsets run on device tensors")
        L = self._sync_engine(dev)
        _lib.check(L.mlic_engine_set_option(self._engine, b"tensor_cores", 1 if self.tensor_cores else 0))
        _lib.check(L.mlic_engine_set_option(self._engine, b"profile", 1 if self._profile else 0))
        _lib.check(L.mlic_engine_set_option(self._engine, b"fuse", 1 if self.fuse else 0))
        _lib.check(L.mlic_engine_set_option(self._engine, b"trace", 1 if self._trace else 0))
        _lib.check(L.mlic_engine_set_option(self._engine, b"stages", int(stages)))
        _lib.check(L.mlic_engine_set_option_f(self._engine, b"z_qstep", float(zq)))
        prec = _lib.PREC_BF16 if self.precision == "bf16" else _lib.PREC_FP32
        odev = "cpu" if host else dev
        h, w, hz, wz = H // 16, W // 16, H // 64, W // 64
        out = {}
        buf = _lib.Buffers()
        keep = []
        if x is not None:
            x = x.detach().to(torch.float32).contiguous()
            buf.x = x.data_ptr()

        def new(name, shape, dtype=torch.float32):
            t = torch.empty(shape, dtype=dtype, device=odev, pin_memory=host)
            out[name] = t
            setattr(buf, name, t.data_ptr())

        def given(name, t):
            t = t.detach().to(dev, torch.float32).contiguous()
            if tuple(t.shape) != (B, self.M, h, w):
                raise ValueError(f"{name} must be [{B},{self.M},{h},{w}], got {tuple(t.shape)}")
            keep.append(t)
            setattr(buf, name, t.data_ptr())

        if stages & 4:
            new("x_hat", (B, 3, H, W))
        if mode == _lib.MODE_FORWARD and stages & 2:
            new("y_likelihoods", (B, self.M, h, w))
            new("z_likelihoods", (B, self.N, hz, wz))
            if "rd_sums" in want and stages == 7:
                new("rd_sums", (2,), torch.float64)
        if mode == _lib.MODE_COMPRESS and stages & 2:
            n = 2 * self.slice_num * B * self.slice_ch * h * (w // 2)
            new("symbols", (n,), torch.int32)
            new("indexes", (n,), torch.int32)
            new("z_symbols", (B, self.N, hz, wz), torch.int32)
        if y is not None:
            given("y", y)
        elif "y" in want:
            new("y", (B, self.M, h, w))
        if y_hat is not None:
            given("y_hat", y_hat)
        elif "y_hat" in want:
            new("y_hat", (B, self.M, h, w))
        with torch.cuda.device(dev):
            if host:
                _lib.check(L.mlic_run_host(self._engine, mode, prec, B, H, W, float(gain), C.byref(buf), 1))
            else:
                need = C.c_size_t()
                _lib.check(L.mlic_workspace_bytes(self._engine, mode, prec, B, H, W, C.byref(need)))
                key = str(dev)
                ws = self._ws.get(key)
                if ws is None or ws.numel() < need.value:
                    self._ws[key] = ws = None
                    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
                    self._ws[key] = ws
                stream = torch.cuda.current_stream(dev).cuda_stream
                try:
                    _lib.check(L.mlic_run(self._engine, mode, prec, B, H, W, float(gain), C.byref(buf), C.c_void_p(ws.data_ptr()),
                                          ws.numel(), C.c_void_p(stream)))
                finally:
                    if stages != 7:
                        L.mlic_engine_set_option(self._engine, b"stages", 7)
                for t in keep:                       # the inputs are read by kernels queued on this stream
                    t.record_stream(torch.cuda.current_stream(dev))
        self.last_launch_count = int(L.mlic_last_launch_count(self._engine))
        return out

    # row-band stage calls (mlic_b200/dist.py).  H is the height of the band in image rows.
    @torch.no_grad()
    def analysis_band(self, x):
        """g_a alone on a band of image rows (a multiple of 16) -> y [B,M,rows/16,W/16] fp32."""
        B, _, H, W = x.shape
        return self._run(_lib.MODE_FORWARD, x, B, H, W, 0.0, ("y",), stages=1)["y"]

    @torch.no_grad()
    def entropy_from_y(self, y, gain=0.0):
        """h_a, EntropyBottleneck, h_s and the slice loop on a whole latent -> (likelihoods dict, y_hat)."""
        B, _, h, w = y.shape
        o = self._run(_lib.MODE_FORWARD, None, B, 16 * h, 16 * w, gain, ("y_hat",), stages=2, y=y)
        return {"y_likelihoods": o["y_likelihoods"], "z_likelihoods": o["z_likelihoods"]}, o["y_hat"]

    @torch.no_grad()
    def synthesis_band(self, y_hat):
        """g_s alone on a band of latent rows -> x_hat [B,3,16*rows,16*w]."""
        B, _, h, w = y_hat.shape
        return self._run(_lib.MODE_FORWARD, None, B, 16 * h, 16 * w, 0.0, (), stages=4, y_hat=y_hat)["x_hat"]

    def graphed(self, B, H, W, fn=None, device=None):
        """-> GraphedCall: `fn` (default: forward) captured once for [B,3,H,W] device inputs, replayed per call."""
        return GraphedCall(self, B, H, W, fn, device)

    # ------------------------------------------------------------------ the reference's public methods
    @torch.no_grad()
    def forward(self, x, *, taps=()):
        """models/mlicpp.py:79-185 -> {"x_hat", "likelihoods": {"y_likelihoods", "z_likelihoods"}}"""
        B, _, H, W = x.shape
        self.update_resolutions(H // 16, W // 16)
        o = self._run(_lib.MODE_FORWARD, x, B, H, W, 0.0, taps)
        res = {"x_hat": o["x_hat"], "likelihoods": {"y_likelihoods": o["y_likelihoods"], "z_likelihoods": o["z_likelihoods"]}}
        res.update({k: o[k] for k in taps})
        return res

    def _tables(self, mod):
        if mod._offset.numel() == 0:
            raise RuntimeError("call update(force=True) before compress() / decompress() (models/mlicpp.py:470-475)")
        return (mod._quantized_cdf.detach().cpu().numpy(), mod._cdf_length.detach().cpu().numpy().reshape(-1),
                mod._offset.detach().cpu().numpy().reshape(-1))

    def _to_host(self, *tensors):
        """Device int32 tensors -> numpy views of cached PINNED host buffers (one async copy each, one sync): the 2 x 10 MB
        symbol / index lists of a 1080p image cost ~3.5 ms through pageable memory, under 1 ms this way.  The views are
        valid until the next call."""
        if not hasattr(self, "_pin"):
            self._pin = {}
        out = []
        for i, t in enumerate(tensors):
            if not t.is_cuda:
                out.append(t.numpy())
                continue
            buf = self._pin.get(i)
            if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
                buf = self._pin[i] = torch.empty(max(t.numel(), 1), dtype=t.dtype, pin_memory=True)
            view = buf[:t.numel()].view(t.shape)
            view.copy_(t, non_blocking=True)
            out.append(view.numpy())
        if any(t.is_cuda for t in tensors):
            torch.cuda.current_stream(next(t for t in tensors if t.is_cuda).device).synchronize()
        return out

    def _z_tables(self, zq=1.0):
        """Quantised CDF tables of the hyper prior: those of update() for the plain bottleneck, per-step tables (cached) for the
        variable-rate one (CompressAI EntropyBottleneckVbr.update_variable)."""
        if zq == 1.0:
            return self._tables(self.entropy_bottleneck)
        from . import coder
        if not hasattr(self, "_zq_tables"):
            self._zq_tables = {}
        key = (float(zq), tuple((v.data_ptr(), v._version) for v in self.entropy_bottleneck.parameters()))
        if key not in self._zq_tables:
            self._zq_tables.clear()
            self._zq_tables[key] = coder.bottleneck_tables(self.entropy_bottleneck, float(zq))
        return self._zq_tables[key]

    def _strings(self, o, B, zq=1.0):
        """The coder side of compress() (models/mlicpp.py:205-206,279-280): ONE y string for the whole batch (the symbol
        lists are flattened over [B,C,H,W/2] per half-slice), one z string per image (EntropyBottleneck.compress)."""
        from . import coder
        sym, idx, zs = self._to_host(o["symbols"], o["indexes"], o["z_symbols"])
        y_string = coder.encode_with_indexes(sym, idx, *self._tables(self.gaussian_conditional))
        ztab = self._z_tables(zq)
        zidx = np.broadcast_to(np.arange(self.N, dtype=np.int32)[:, None, None], zs.shape[1:])
        return [[y_string], [coder.encode_with_indexes(zs[b], zidx, *ztab) for b in range(B)]]

    @torch.no_grad()
    def compress(self, x, *, taps=()):
        """models/mlicpp.py:199-290 -> {"strings": [[y_string], z_strings], "shape", "cost_time"} plus the coder's inputs
        (`symbols`, `indexes`: flat int32 in the reference's list order A0,N0,A1,N1,...; `z_symbols`) and `x_hat`."""
        t0 = time.time()
        B, _, H, W = x.shape
        self.update_resolutions(H // 16, W // 16)
        o = self._run(_lib.MODE_COMPRESS, x, B, H, W, 0.0, taps)
        if x.is_cuda:
            torch.cuda.synchronize(x.device)
        strings = self._strings(o, B) if self.gaussian_conditional._offset.numel() else None
        o.update(strings=strings, shape=(H // 64, W // 64), cost_time=time.time() - t0)
        return o

    @torch.no_grad()
    def decompress(self, strings, shape, *, taps=(), _gain=0.0, _zq=1.0):
        """models/mlicpp.py:292-378 -> {"x_hat", "cost_time"}: z strings decoded on the host, then the decoder-side walk with
        the range decoder inside the slice loop (C ABI mlic_decompress)."""
        from . import coder
        t0 = time.time()
        if not torch.cuda.is_available():
            raise _lib.MlicError("mlic_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        y_string, z_strings = strings[0][0], strings[1]
        B, (hz, wz) = len(z_strings), shape
        H, W = 64 * int(hz), 64 * int(wz)
        ztab = self._z_tables(_zq)
        zidx = np.broadcast_to(np.arange(self.N, dtype=np.int32)[:, None, None], (self.N, hz, wz))
        dec = coder.RansDecoder()
        zs = np.stack([dec.decode_with_indexes(s, zidx, *ztab).reshape(self.N, hz, wz) for s in z_strings])
        dec.close()
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.MlicError("move the model to a CUDA device before decompress()")
        self.update_resolutions(H // 16, W // 16)
        L = self._sync_engine(dev)
        ytab = self._tables(self.gaussian_conditional)
        cdf, ln, off = (np.ascontiguousarray(t, dtype=np.int32) for t in ytab)
        _lib.check(L.mlic_engine_set_cdf(self._engine, cdf.ctypes.data_as(C.c_void_p), cdf.shape[1], ln.ctypes.data_as(C.c_void_p),
                                         off.ctypes.data_as(C.c_void_p), cdf.shape[0]))
        for name, val in ((b"tensor_cores", self.tensor_cores), (b"profile", False), (b"fuse", self.fuse), (b"trace", self._trace), (b"stages", 7)):
            _lib.check(L.mlic_engine_set_option(self._engine, name, int(val)))
        _lib.check(L.mlic_engine_set_option_f(self._engine, b"z_qstep", float(_zq)))
        # per-stage precisions: the decoder must reproduce the encoder's mu / sigma, so the whole walk runs in the entropy stage's
        one = self.precision[1] if isinstance(self.precision, tuple) else self.precision
        prec = _lib.PREC_BF16 if one == "bf16" else _lib.PREC_FP32
        out = {"x_hat": torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)}
        if "y_hat" in taps:
            out["y_hat"] = torch.empty((B, self.M, H // 16, W // 16), dtype=torch.float32, device=dev)
        z_dev = torch.from_numpy(zs.astype(np.int32)).to(dev)
        ybuf = np.frombuffer(bytes(y_string), dtype=np.uint8)
        with torch.cuda.device(dev):
            need = C.c_size_t()
            _lib.check(L.mlic_workspace_bytes(self._engine, _lib.MODE_DECOMPRESS, prec, B, H, W, C.byref(need)))
            key = str(dev)
            ws = self._ws.get(key)
            if ws is None or ws.numel() < need.value:
                self._ws[key] = ws = None
                ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
                self._ws[key] = ws
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(L.mlic_decompress(self._engine, prec, B, H, W, float(_gain), ybuf.ctypes.data_as(C.c_void_p), ybuf.size,
                                         C.c_void_p(z_dev.data_ptr()), C.c_void_p(out["x_hat"].data_ptr()),
                                         C.c_void_p(out["y_hat"].data_ptr()) if "y_hat" in out else None,
                                         C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(stream)))
            torch.cuda.synchronize(dev)
        self.last_launch_count = int(L.mlic_last_launch_count(self._engine))
        out["cost_time"] = time.time() - t0
        return out

    @torch.no_grad()
    def net_decoder_forward(self, x):
        """models/mlicpp.py:380-459: decoder-side network walk (z_hat = 0); x is used for its shape only."""
        B, _, H, W = x.shape
        self.update_resolutions(H // 16, W // 16)
        if x.is_cuda:
            with torch.cuda.device(x.device):
                o = self._run(_lib.MODE_DECODER, None, B, H, W, 0.0)
            return o["x_hat"]
        raise _lib.MlicError("net_decoder_forward expects a CUDA tensor (its values are not read)")


class GraphedCall:
    """One network call of fixed shape captured in a CUDA graph (the slice loop is ~500 launches of 10-250 us: replaying
    them removes the host launch gaps, 9.0 -> 7.8 ms for one 1920x1088 forward).  `fn(x)` is any device-path call of the
    model (`net`, `net.net_decoder_forward`, a lambda around a VBR forward with its gain passed as `inputscale` -- a capturing
    stream allows no device -> host read of `Gain[s]` ...).  Replay copies x into the captured input
    and returns the captured output tensors, which the next replay overwrites.  The graph bakes in the engine's packed
    weights and workspace: after load_state_dict / update / .to() build a new GraphedCall (the call checks and raises)."""

    def __init__(self, net, B, H, W, fn=None, device=None):
        if not torch.cuda.is_available():
            raise _lib.MlicError("mlic_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device if device is not None else "cuda")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.net, self.fn = net, (fn if fn is not None else net)
        self.x = torch.zeros(B, 3, H, W, device=dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                      # packs the weights, sizes the workspace, sets kernel attributes: all
                self.fn(self.x)                     # of it outside the capture
            side.synchronize()
            self._sig = net._engine_sig
            self._ws = net._ws.get(str(dev))        # keeps the captured workspace alive if the model later grows its own
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self.out = self.fn(self.x)
            self.launches = net.last_launch_count
        torch.cuda.current_stream(dev).wait_stream(side)

    def __call__(self, x):
        if self.net._engine_sig != self._sig or (str(self.x.device), self.net._signature()) != self._sig:
            raise RuntimeError("the model's parameters changed after the graph was captured; build a new GraphedCall")
        if tuple(x.shape) != tuple(self.x.shape):
            raise ValueError(f"captured for input {tuple(self.x.shape)}, got {tuple(x.shape)}")
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out


class MLICPlusPlusSD(MLICPlusPlus):
    """models/mlicpp_small_decoder.py:16-83 (dense encoder, quarter-width decoder)."""
    KIND = "sd"


class MLICPlusPlusVbr(MLICPlusPlus):
    """models/mlicpp_vbr.py:14-117: 6 gain levels, stage-2 gain-scaled quantisation (no_quantoffset=True)."""
    KIND = "vbr"
    LAMBDAS = VBR_LAMBDAS

    def __init__(self, config, name=None, vr_entbttlnck=None, **kwargs):
        super().__init__(config, name=name, _vr_entbttlnck=bool(vr_entbttlnck), **kwargs)
        self.lmbda = list(self.LAMBDAS)
        self.levels = len(self.lmbda)
        self.no_quantoffset = True
        self.vr_entbttlnck = vr_entbttlnck          # mlicpp_vbr.py:103-117: variable-rate hyper prior (EntropyBottleneckVbr + gayn2zqstep)

    def _scale(self, s, inputscale, absolute):
        if inputscale != 0:
            return float(inputscale)
        if absolute:                                   # mlicpp_vbr.py:540-543
            assert s in range(0, self.levels), f"s should in range(0, {self.levels}), but get s:{s}"
            return abs(float(self.Gain[s].detach()))
        s = max(0, min(int(s), self.Gain.numel() - 1))  # mlicpp_vbr.py:122-135
        return float(self.Gain[s].detach())

    def _zqstep(self, scale):
        """mlicpp_vbr.py:255-256,554-555: z_qstep = LowerBound(0.5)(gayn2zqstep(1 / scale)), a 1-10-10-1 ReLU network with a Softplus
        on top, evaluated on the host in fp32 exactly as the reference's nn.Sequential does (CPU torch); 1.0 without vr_entbttlnck."""
        if not self.vr_entbttlnck:
            return 1.0
        g = self.gayn2zqstep
        t = 1.0 / torch.tensor([float(scale)], dtype=torch.float32)
        for j in (0, 2, 4):
            lin = getattr(g, str(j))
            t = nn.functional.linear(t, lin.weight.detach().float().cpu(), lin.bias.detach().float().cpu())
            t = torch.relu(t) if j < 4 else nn.functional.softplus(t)
        return float(torch.clamp(t, min=float(self.lower_bound_zqstep.bound))[0])

    @torch.no_grad()
    def forward(self, x, stage=2, s=1, inputscale=0, *, taps=()):
        if stage not in (1, 2):
            raise ValueError(f"Invalid stage (stage={stage}) parameter for this model.")     # mlicpp_vbr.py:119-120
        B, _, H, W = x.shape
        gain = self._scale(s, inputscale, False) if stage == 2 else 0.0
        zq = self._zqstep(gain) if stage == 2 else 1.0             # stage 1 runs the plain bottleneck (mlicpp_vbr.py:160)
        o = self._run(_lib.MODE_FORWARD, x, B, H, W, gain, taps, zq=zq)
        res = {"x_hat": o["x_hat"], "likelihoods": {"y_likelihoods": o["y_likelihoods"], "z_likelihoods": o["z_likelihoods"]}}
        res.update({k: o[k] for k in taps})
        return res

    @torch.no_grad()
    def compress(self, x, stage=2, s=1, inputscale=0, *, taps=()):
        t0 = time.time()
        B, _, H, W = x.shape
        scale = self._scale(s, inputscale, True)
        zq = self._zqstep(scale) if stage == 2 else 1.0            # mlicpp_vbr.py:550-559
        o = self._run(_lib.MODE_COMPRESS, x, B, H, W, scale, taps, zq=zq)
        if x.is_cuda:
            torch.cuda.synchronize(x.device)
        strings = self._strings(o, B, zq) if self.gaussian_conditional._offset.numel() else None
        o.update(strings=strings, shape=(H // 64, W // 64), cost_time=time.time() - t0)
        return o

    @torch.no_grad()
    def decompress(self, strings, shape, stage=2, s=1, inputscale=0, *, taps=()):
        """models/mlicpp_vbr.py:889-1040 (stage 2 arithmetic: indexes from sigma * gain, y_hat = symbols / gain + means).  With
        vr_entbttlnck the z strings are decoded on the step compress() used (the reference's decompress() omits `qs` there,
        mlicpp_vbr.py:905, and cannot read its own vr streams)."""
        scale = self._scale(s, inputscale, True)
        return super().decompress(strings, shape, taps=taps, _gain=scale, _zq=self._zqstep(scale) if stage == 2 else 1.0)


class MLICPlusPlusSDVbr(MLICPlusPlusVbr):
    """models/mlicpp_sd_vbr.py:19-127: the small-decoder network (dense encoder, quarter-width h_s / g_s / contexts,
    LRP-Old) under the Vbr methods, 5 gain levels. forward / compress / decompress bodies are those of mlicpp_vbr.py."""
    KIND = "sdvbr"
    LAMBDAS = SDVBR_LAMBDAS


_CLASSES = {"base": MLICPlusPlus, "sd": MLICPlusPlusSD, "vbr": MLICPlusPlusVbr, "sdvbr": MLICPlusPlusSDVbr}


def get_model(name):
    """models/model_loader.py:4-18 (+ MLICPP_L_VBR, which BASELINE.json names)."""
    if name not in MODEL_TABLE:
        raise KeyError(f"unknown model '{name}'; known: {sorted(MODEL_TABLE)}")
    return _CLASSES[MODEL_TABLE[name]["kind"]](model_config(name), name=name)
