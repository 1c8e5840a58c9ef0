"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the MLIC++ network-forward path.

A functional (state_dict-driven) torch/CPU restatement of what the reference's
`MLICPlusPlus`, `MLICPlusPlusSD` and `MLICPlusPlusVbr` compute in `forward`,
`compress` (symbol / CDF-index generation) and `net_decoder_forward`.  It runs
anywhere torch runs (the GPU box has no /root/reference), is the checker for the
CUDA engine's parity tests, and is the `cpu_baseline` / `--impl reference` arm of
bench.py.  It is never imported by the product package.

Parity pin: the reference holds no tests or golden vectors.  This
restatement is pinned against outputs of the reference's own model code executed
in the build container (tests/golden/*.npz, produced by oracle/make_golden.py over
the CompressAI shim in oracle/ref_shim) -- see tests/test_oracle.py.

Every function cites the reference file:line it follows (paths relative to
/root/reference/MLIC++).
"""
import math

import torch
import torch.nn.functional as F

# config/config.py:19-62 (+ SURVEY F4: MLICPP_L_VBR := Vbr class over the L config)
MODEL_TABLE = {
    "MLICPP_L": dict(N=192, M=320, slices=10, kind="base"),
    "MLICPP_M": dict(N=160, M=256, slices=8, kind="base"),
    "MLICPP_S": dict(N=96, M=160, slices=5, kind="base"),
    "MLICPP_S2": dict(N=128, M=128, slices=2, kind="base"),
    "MLICPP_M_SMALL_DEC": dict(N=192, M=320, slices=10, kind="sd"),
    "MLICPP_S_VBR": dict(N=96, M=160, slices=5, kind="vbr"),
    "MLICPP_L_VBR": dict(N=192, M=320, slices=10, kind="vbr"),
    "MLICPP_M_SMALL_DEC_VBR": dict(N=192, M=320, slices=10, kind="sdvbr"),   # mlicpp_sd_vbr.py:19-127
}

SCALE_MIN, SCALE_MAX, SCALE_LEVELS = 0.11, 256.0, 64
LIK_FLOOR = 1e-9


def scale_table():
    """utils/func.py:16-19"""
    return torch.exp(torch.linspace(math.log(SCALE_MIN), math.log(SCALE_MAX), SCALE_LEVELS))


def parity_masks(H, W, dtype=torch.float32):
    """utils/ckbd.py:35-45 -- anchor = (row+col) odd, non-anchor = (row+col) even."""
    r = torch.arange(H).view(H, 1)
    c = torch.arange(W).view(1, W)
    anchor = ((r + c) % 2 == 1).to(dtype).view(1, 1, H, W)
    return anchor, 1.0 - anchor


def squeeze_parity(t, anchor):
    """utils/ckbd.py:47-59 -- keep, per row, the columns of one parity -> [B,C,H,W/2]."""
    B, C, H, W = t.shape
    out = t.new_zeros(B, C, H, W // 2)
    if anchor:
        out[:, :, 0::2] = t[:, :, 0::2, 1::2]
        out[:, :, 1::2] = t[:, :, 1::2, 0::2]
    else:
        out[:, :, 0::2] = t[:, :, 0::2, 0::2]
        out[:, :, 1::2] = t[:, :, 1::2, 1::2]
    return out


def unsqueeze_parity(t, anchor):
    """utils/ckbd.py:61-73"""
    B, C, H, Wh = t.shape
    out = t.new_zeros(B, C, H, Wh * 2)
    if anchor:
        out[:, :, 0::2, 1::2] = t[:, :, 0::2]
        out[:, :, 1::2, 0::2] = t[:, :, 1::2]
    else:
        out[:, :, 0::2, 0::2] = t[:, :, 0::2]
        out[:, :, 1::2, 1::2] = t[:, :, 1::2]
    return out


def gaussian_likelihood(y_hat, sigma, mu):
    """CompressAI GaussianConditional._likelihood + lower bound (call site mlicpp.py:132)."""
    s = torch.clamp_min(sigma, SCALE_MIN)
    v = torch.abs(y_hat - mu)
    c = -(2 ** -0.5)
    upper = 0.5 * torch.erfc(c * ((0.5 - v) / s))
    lower = 0.5 * torch.erfc(c * ((-0.5 - v) / s))
    return torch.clamp_min(upper - lower, LIK_FLOOR)


def cdf_indexes(sigma, table=None):
    """CompressAI GaussianConditional.build_indexes (call site utils/ckbd.py:128)."""
    table = scale_table() if table is None else table
    s = torch.clamp_min(sigma, SCALE_MIN)
    idx = torch.full(s.shape, SCALE_LEVELS - 1, dtype=torch.int32)
    for k in range(SCALE_LEVELS - 1):
        idx -= (s <= table[k]).to(torch.int32)
    return idx


class Oracle:
    """Functional MLIC++ over a reference-shaped state_dict (fp32 or fp64, CPU)."""

    def __init__(self, name, state_dict, dtype=torch.float32):
        cfg = MODEL_TABLE[name]
        self.name, self.kind = name, cfg["kind"]
        self.sd, self.vbr = self.kind in ("sd", "sdvbr"), self.kind in ("vbr", "sdvbr")
        self.N, self.M, self.S = cfg["N"], cfg["M"], cfg["slices"]
        self.C = self.M // self.S
        self.dtype = dtype
        self.w = {k: (v.detach().cpu().to(dtype) if torch.is_floating_point(v) else v.detach().cpu())
                  for k, v in state_dict.items()}
        self.table = scale_table()
        self.trace = None
        self._mask_cache = {}
        # mlicpp_vbr.py:103-117: the variable-rate hyper prior exists iff its gain -> step network is in the state_dict
        self.vr_entbttlnck = self.vbr and "gayn2zqstep.0.weight" in self.w

    # ---------------------------------------------------------------- primitive layers
    def _conv(self, x, p, stride=1, pad=0, groups=1):
        return F.conv2d(x, self.w[p + ".weight"], self.w[p + ".bias"], stride=stride, padding=pad, groups=groups)

    def _ds(self, x, p, stride=1):
        """modules/layers/conv.py:46-63 -- depthwise 3x3 (stride here) then pointwise 1x1, no act between."""
        x = self._conv(x, p + ".depth_conv", stride=stride, pad=1, groups=x.shape[1])
        return self._conv(x, p + ".point_conv")

    def _c3(self, x, p, stride=1, dense=False):
        """modules/layers/conv.py:22-32"""
        return self._conv(x, p, stride=stride, pad=1) if dense else self._ds(x, p, stride)

    def _subpel(self, x, p):
        """CompressAI subpel_conv3x3 = Conv2d(C->4C,3,pad1) + PixelShuffle(2) (res_blk.py:107,111)."""
        return F.pixel_shuffle(self._conv(x, p + ".0", pad=1), 2)

    def _gdn(self, x, p, inverse):
        """CompressAI GDN with NonNegativeParametrizer (res_blk.py:76,110)."""
        ped = 2.0 ** -36
        beta = torch.clamp_min(self.w[p + ".beta"], (1e-6 + ped) ** 0.5) ** 2 - ped
        gamma = torch.clamp_min(self.w[p + ".gamma"], 2.0 ** -18) ** 2 - ped
        C = x.shape[1]
        norm = F.conv2d(x * x, gamma.reshape(C, C, 1, 1), beta)
        return x * (torch.sqrt(norm) if inverse else torch.rsqrt(norm))

    def _rb(self, x, p, dense=False):
        """res_blk.py:142-154 -- GELU(conv2(GELU(conv1 x))) + skip(x)."""
        o = F.gelu(self._c3(x, p + ".conv1", dense=dense))
        o = F.gelu(self._c3(o, p + ".conv2", dense=dense))
        idn = self._conv(x, p + ".skip") if (p + ".skip.weight") in self.w else x
        return o + idn

    def _rbws(self, x, p, dense=False):
        """res_blk.py:82-93 -- GDN(conv2(GELU(conv1_s2 x))) + skip_1x1_s2(x)."""
        o = F.gelu(self._c3(x, p + ".conv1", stride=2, dense=dense))
        o = self._gdn(self._c3(o, p + ".conv2", dense=dense), p + ".gdn", False)
        return o + self._conv(x, p + ".skip", stride=2)

    def _rbu(self, x, p):
        """res_blk.py:113-121 -- IGDN(conv(GELU(subpel x))) + upsample(x)."""
        o = F.gelu(self._subpel(x, p + ".subpel_conv"))
        o = self._gdn(self._c3(o, p + ".conv"), p + ".igdn", True)
        return o + self._subpel(x, p + ".upsample")

    # ---------------------------------------------------------------- transforms
    def g_a(self, x):
        """transform/analysis.py:9-17 (analysis_old.py:10-16 = dense, SD)."""
        d = self.sd
        p = "g_a.analysis_transform."
        for i in (0, 2, 4):
            x = self._rbws(x, p + str(i), d)
            x = self._rb(x, p + str(i + 1), d)
        return self._c3(x, p + "6", stride=2, dense=d)

    def h_a(self, y):
        """transform/analysis.py:33-43"""
        d = self.sd
        p = "h_a.reduction."
        x = y
        for i, s in zip((0, 2, 4, 6, 8), (1, 1, 2, 1, 2)):
            x = self._c3(x, p + str(i), stride=s, dense=d)
            if i != 8:
                x = F.gelu(x)
        return x

    def h_s(self, z_hat):
        """transform/synthesis.py:18-28"""
        p = "h_s.increase."
        x = F.gelu(self._c3(z_hat, p + "0"))
        x = F.gelu(self._subpel(x, p + "2"))
        x = F.gelu(self._c3(x, p + "4"))
        x = F.gelu(self._subpel(x, p + "6"))
        return self._c3(x, p + "8")

    def g_s(self, y_hat):
        """transform/synthesis.py:59-68"""
        p = "g_s.synthesis_transform."
        x = self._rb(y_hat, p + "0")
        for i in (1, 3, 5):
            x = self._rbu(x, p + str(i))
            x = self._rb(x, p + str(i + 1))
        return self._subpel(x, p + "7")

    def z_qstep(self, scale):
        """mlicpp_vbr.py:255-256,554-555: LowerBound(0.5)(Softplus(Linear(ReLU(Linear(ReLU(Linear(1 / scale))))))); None without vr_entbttlnck."""
        if not self.vr_entbttlnck or scale is None:
            return None
        t = 1.0 / scale.reshape(1)
        for j in (0, 2, 4):
            t = F.linear(t, self.w[f"gayn2zqstep.{j}.weight"], self.w[f"gayn2zqstep.{j}.bias"])
            t = torch.relu(t) if j < 4 else F.softplus(t)
        return torch.clamp(t, min=float(self.w["lower_bound_zqstep.bound"]))[0]

    def entropy_bottleneck(self, z, qs=None):
        """CompressAI EntropyBottleneck eval forward (mlicpp.py:96-98): returns (z_hat, z_lik).  qs: quantisation step of
        EntropyBottleneckVbr (restated, unpinned): z_hat = round((z - med) / qs) qs + med, likelihood over [z_hat - qs/2, z_hat + qs/2]."""
        B, Cc, H, W = z.shape
        med = self.w["entropy_bottleneck.quantiles"][:, 0, 1].view(1, Cc, 1, 1)
        if qs is None:
            z_hat = torch.round(z - med) + med
            half = 0.5
        else:
            z_hat = torch.round((z - med) / qs) * qs + med
            half = 0.5 * qs
        v = z_hat.permute(1, 0, 2, 3).reshape(Cc, 1, -1)

        def cum(t):
            for i in range(5):
                t = torch.matmul(F.softplus(self.w[f"entropy_bottleneck.matrices.{i}"]), t)
                t = t + self.w[f"entropy_bottleneck.biases.{i}"]
                if i < 4:
                    t = t + torch.tanh(self.w[f"entropy_bottleneck.factors.{i}"]) * torch.tanh(t)
            return t

        lik = torch.sigmoid(cum(v + half)) - torch.sigmoid(cum(v - half))
        lik = torch.clamp_min(lik, LIK_FLOOR).reshape(Cc, B, H, W).permute(1, 0, 2, 3).contiguous()
        return z_hat, lik

    # ---------------------------------------------------------------- entropy-model pieces
    def _ep(self, x, p):
        """transform/entropy.py:10-29 -- 1x1 chain in->320->256->128->2C, GELU between."""
        for i in (0, 2, 4):
            x = F.gelu(self._conv(x, f"{p}.fusion.{i}"))
        return self._conv(x, f"{p}.fusion.6")

    def _lrp(self, x, p):
        """transform/quantization.py:33-44 (:12-28 for the SD 'Old' pyramid) -- 0.5*tanh(stack)."""
        idxs = (0, 2, 4, 6) if self.sd else (0, 2, 4)
        for n, i in enumerate(idxs):
            x = self._ds(x, f"{p}.lrp_transform.{i}")
            if n != len(idxs) - 1:
                x = F.gelu(x)
        return 0.5 * torch.tanh(x)

    def _channel_ctx(self, x, p):
        """transform/context.py:118-138 (context_old.py:120-126 = dense, SD)."""
        d = self.sd
        x = F.gelu(self._c3(x, p + ".fushion.0", dense=d))
        x = F.gelu(self._c3(x, p + ".fushion.2", dense=d))
        return self._c3(x, p + ".fushion.4", dense=d)

    def local_mask(self, H, W):
        """transform/context.py:43-65 -- [L,25,25]: 0 iff both window taps are in-image anchors, else -100."""
        key = (H, W)
        if key not in self._mask_cache:
            a, _ = parity_masks(H, W)
            win = F.unfold(a, kernel_size=5, padding=2)          # [1,25,L]; zero outside the image
            win = win[0].t()                                      # [L,25]
            both = win.unsqueeze(2) * win.unsqueeze(1)            # [L,25,25]
            self._mask_cache[key] = torch.where(both > 0.5, 0.0, -100.0).to(self.dtype)
        return self._mask_cache[key]

    def _local_ctx(self, x, p):
        """transform/context.py:67-112 (semantics: SURVEY appendix A.4)."""
        B, C, H, W = x.shape
        L, heads, P = H * W, 2, 25
        d = C // heads
        w = self.w
        t = x.permute(0, 2, 3, 1).reshape(B, L, C)
        t = F.layer_norm(t, (C,), w[p + ".norm1.weight"], w[p + ".norm1.bias"], 1e-5)
        f = F.linear(t, w[p + ".qkv_proj.weight"], w[p + ".qkv_proj.bias"])          # [B,L,3C], feature u*C+c
        f = f.reshape(B, H, W, 3 * C).permute(0, 3, 1, 2)                            # [B,3C,H,W]
        win = F.unfold(f, kernel_size=5, padding=2)                                  # [B,3C*25,L], zero-padded taps
        win = win.reshape(B, 3, d, heads, P, L)                                      # channel c = dd*heads + hh
        q, k, v = (win[:, u].permute(0, 4, 2, 3, 1) for u in range(3))               # [B,L,heads,P,d]
        q = q * (d ** -0.5)
        att = q @ k.transpose(-2, -1)                                                # [B,L,heads,P,P]
        rel = w[p + ".relative_position_table"][w[p + ".relative_position_index"].reshape(-1)]
        att = att + rel.reshape(P, P, heads).permute(2, 0, 1)[None, None]
        att = att + self.local_mask(H, W)[None, :, None]
        att = torch.softmax(att, dim=-1)
        o = (att @ v)                                                                # [B,L,heads,P,d]
        o = o.permute(0, 1, 3, 2, 4).reshape(B * L, 5, 5, C).permute(0, 3, 1, 2)      # out channel c = hh*d + dd
        o = F.conv2d(o, w[p + ".fusion.weight"], w[p + ".fusion.bias"]).reshape(B, L, 2 * C)
        o = F.linear(o, w[p + ".proj.weight"], w[p + ".proj.bias"])
        m = F.layer_norm(o, (2 * C,), w[p + ".norm2.weight"], w[p + ".norm2.bias"], 1e-5)
        m = F.linear(F.gelu(F.linear(m, w[p + ".mlp.fc1.weight"], w[p + ".mlp.fc1.bias"])),
                     w[p + ".mlp.fc2.weight"], w[p + ".mlp.fc2.bias"])
        o = o + m
        return o.permute(0, 2, 1).reshape(B, 2 * C, H, W)

    def _qkv(self, x, p):
        """1x1 conv then depthwise 3x3 (context.py:149-160,204-215)."""
        x = self._conv(x, p + ".0")
        return self._conv(x, p + ".1", pad=1, groups=x.shape[1])

    def _mlp_dw(self, x, p):
        """1x1 -> GELU -> dw3x3 -> GELU -> 1x1 (context.py:161-167,217-223)."""
        x = F.gelu(self._conv(x, p + ".0"))
        x = F.gelu(self._conv(x, p + ".2", pad=1, groups=x.shape[1]))
        return self._conv(x, p + ".4")

    def _inter_ctx(self, x, p, heads):
        """transform/context.py:226-245."""
        B, D, H, W = x.shape
        N = H * W
        hd = D // heads
        q = self._qkv(x, p + ".queries").reshape(B, heads, hd, N)
        k = self._qkv(x, p + ".keys").reshape(B, heads, hd, N)
        v = self._qkv(x, p + ".values").reshape(B, heads, hd, N)
        k = torch.softmax(k, dim=3)
        q = torch.softmax(q, dim=2)
        ctx = k @ v.transpose(2, 3)                      # [B,heads,hd,hd]
        o = (ctx.transpose(2, 3) @ q).reshape(B, D, H, W)
        a = self._conv(o, p + ".reprojection", pad=2)
        return self._conv(a, p + ".skip") + self._mlp_dw(a, p + ".mlp")

    def _intra_ctx(self, x1, x2, p):
        """transform/context.py:169-193: compute on squeezed halves, scatter once."""
        B, C, H, W = x1.shape
        am, nm = parity_masks(H, W, x1.dtype)
        heads, hd, Nh = 2, C // 2, H * W // 2
        q = squeeze_parity(self._qkv(x1 * nm, p + ".queries"), False).reshape(B, heads, hd, Nh)
        k = squeeze_parity(self._qkv(x1 * am, p + ".keys"), True).reshape(B, heads, hd, Nh)
        v = squeeze_parity(self._qkv(x2, p + ".values"), True).reshape(B, heads, hd, Nh)
        k = torch.softmax(k, dim=3)
        q = torch.softmax(q, dim=2)
        ctx = k @ v.transpose(2, 3)
        o = (ctx.transpose(2, 3) @ q).reshape(B, C, H, W // 2)
        o = unsqueeze_parity(o, False)
        a = self._conv(o, p + ".reprojection", pad=2)
        return a + self._mlp_dw(a, p + ".mlp")

    # ---------------------------------------------------------------- the slice loop
    def _gain(self, s, inputscale=0):
        """mlicpp_vbr.py:122-135 (forward, eval) / :537-544 (compress: abs)."""
        if not self.vbr:
            return None
        if inputscale != 0:
            return torch.tensor(float(inputscale), dtype=self.dtype)
        s = max(0, min(int(s), self.w["Gain"].numel() - 1))
        return self.w["Gain"][s]

    def _entropy_loop(self, y, hyper, mode, gain=None, rec=None):
        """mlicpp.py:104-176 (forward) / :209-277 (compress) / :396-453 (net_decoder_forward).

        mode: "forward" | "compress" | "decoder".  Returns (y_hat, y_lik|None, symbols, indexes).
        VBR (stage 2, no_quantoffset): mlicpp_vbr.py:260-336 (forward), utils/ckbd.py:76-90,146-158 (compress).
        """
        B, _, H, W = hyper.shape
        C, S = self.C, self.S
        am, nm = parity_masks(H, W, self.dtype)
        hyper_means = hyper[:, hyper.shape[1] // 2:]
        y_hat_slices, liks, syms, idxs = [], [], [], []
        g = gain
        r = None if g is None else (1.0 / g)

        def quant(vals, mu, sigma, anchor):
            mask = am if anchor else nm
            if mode == "decoder":                       # mlicpp.py:405,418: both halves <- means_anchor
                return None
            if mode == "forward":
                if g is None:
                    return torch.round(vals - mu) + mu               # mlicpp.py:117,134
                return torch.round((vals - mu) * g) * r + mu         # mlicpp_vbr.py:277,294
            # compress: squeeze -> symbols/indexes -> de-quantise -> unsqueeze
            v_s, mu_s, sg_s = (squeeze_parity(t, anchor) for t in (vals, mu, sigma))
            if g is None:
                idx = cdf_indexes(sg_s, self.table)                  # ckbd.py:127-128
                sym = torch.round(v_s - mu_s).to(torch.int32)        # ckbd.py:129
                deq = sym.to(self.dtype) + mu_s                      # ckbd.py:132
            else:
                idx = cdf_indexes(sg_s * g, self.table)              # ckbd.py:82 / :151
                if anchor:                                           # ckbd.py:87 (mean subtracted twice, no gain)
                    sym = torch.round((v_s - mu_s) - mu_s).to(torch.int32)
                else:                                                # ckbd.py:155
                    sym = torch.round(v_s - mu_s).to(torch.int32)
                deq = sym.to(self.dtype) * r + mu_s                  # ckbd.py:90 / :158
            syms.append(sym.reshape(-1))
            idxs.append(idx.reshape(-1))
            return unsqueeze_parity(deq, anchor)

        for i in range(S):
            y_i = y[:, i * C:(i + 1) * C] if y is not None else None
            prev = y_hat_slices
            if i == 0:
                ctx_a = hyper
            else:
                cat_prev = torch.cat(prev, 1)
                inter = self._inter_ctx(cat_prev, f"global_inter_context.{i}", heads=(C * i) // 32)
                chan = self._channel_ctx(cat_prev, f"channel_context.{i}")
                ctx_a = torch.cat([inter, chan, hyper], 1)
            pa = self._ep(ctx_a, f"entropy_parameters_anchor.{i}")
            sig_a, mu_a = pa[:, :C] * am, pa[:, C:] * am
            if mode == "decoder":
                a = mu_a.clone()
            else:
                a = quant(y_i * am, mu_a, sig_a, True)   # stays zero off-parity (round(0-0)+0)
            a = a + am * self._lrp(torch.cat([hyper_means] + prev + [a], 1), f"lrp_anchor.{i}")
            local = self._local_ctx(a, f"local_context.{i}")
            if i == 0:
                ctx_n = torch.cat([local, hyper], 1)
            else:
                intra = self._intra_ctx(prev[-1], a, f"global_intra_context.{i}")
                ctx_n = torch.cat([local, intra, inter, chan, hyper], 1)
            pn = self._ep(ctx_n, f"entropy_parameters_nonanchor.{i}")
            sig_n, mu_n = pn[:, :C] * nm, pn[:, C:] * nm
            if mode == "decoder":
                n = mu_a.clone()                          # sic: mlicpp.py:418 uses means_anchor
            else:
                n = quant(y_i * nm, mu_n, sig_n, False)
            if mode == "forward":
                sig, mu = sig_a + sig_n, mu_a + mu_n
                if g is None:
                    y_q = torch.round(y_i - mu) + mu
                    liks.append(gaussian_likelihood(y_q, sig, mu))            # mlicpp.py:132
                else:
                    yg, sg, mg = y_i * g, sig * g, mu * g                      # mlicpp_vbr.py:292
                    liks.append(gaussian_likelihood(torch.round(yg - mg) + mg, sg, mg))
            yh = a + n
            yh = yh + nm * self._lrp(torch.cat([hyper_means] + prev + [yh], 1), f"lrp_nonanchor.{i}")
            y_hat_slices.append(yh)
            if rec is not None:
                rec[f"mu_a{i}"], rec[f"sig_a{i}"], rec[f"mu_n{i}"], rec[f"sig_n{i}"] = mu_a, sig_a, mu_n, sig_n
                rec[f"local{i}"] = local
                rec[f"y_hat{i}"] = yh
                if i > 0:
                    rec[f"inter{i}"], rec[f"chan{i}"], rec[f"intra{i}"] = inter, chan, intra
        y_hat = torch.cat(y_hat_slices, 1)
        y_lik = torch.cat(liks, 1) if liks else None
        sym = torch.cat(syms) if syms else None
        idx = torch.cat(idxs) if idxs else None
        return y_hat, y_lik, sym, idx

    # ---------------------------------------------------------------- public entry points
    @torch.no_grad()
    def forward(self, x, s=1, inputscale=0, trace=False):
        """MLICPlusPlus.forward (mlicpp.py:79-185); VBR: forward(x, stage=2, s, inputscale) (mlicpp_vbr.py:137)."""
        x = x.to(self.dtype)
        rec = {} if trace else None
        y = self.g_a(x)
        z = self.h_a(y)
        z_hat, z_lik = self.entropy_bottleneck(z, self.z_qstep(self._gain(s, inputscale)))
        hyper = self.h_s(z_hat)
        y_hat, y_lik, _, _ = self._entropy_loop(y, hyper, "forward", self._gain(s, inputscale), rec)
        x_hat = self.g_s(y_hat)
        out = {"x_hat": x_hat, "likelihoods": {"y_likelihoods": y_lik, "z_likelihoods": z_lik}}
        if trace:
            rec.update(y=y, z=z, z_hat=z_hat, hyper=hyper, y_hat=y_hat)
            out["trace"] = rec
        return out

    @torch.no_grad()
    def compress_symbols(self, x, s=1, inputscale=0, trace=False):
        """Network walk of MLICPlusPlus.compress (mlicpp.py:199-277) up to, not including, the rANS coder.

        Returns the flat int32 `symbols` / `indexes` in the reference's list order
        (A0,N0,A1,N1,...; each half-slice flattened over [B,C,H,W/2]) plus z symbols, y_hat and x_hat.
        """
        x = x.to(self.dtype)
        rec = {} if trace else None
        y = self.g_a(x)
        z = self.h_a(y)
        med = self.w["entropy_bottleneck.quantiles"][:, 0, 1].view(1, -1, 1, 1)
        g = self._gain(s, inputscale)
        if g is not None:
            g = torch.abs(g)                               # mlicpp_vbr.py:543
        qs = self.z_qstep(g)                               # mlicpp_vbr.py:553-559 (EntropyBottleneckVbr.compress / decompress with qs)
        if qs is None:
            z_sym = torch.round(z - med).to(torch.int32)
            z_hat = z_sym.to(self.dtype) + med
        else:
            z_sym = torch.round((z - med) / qs).to(torch.int32)
            z_hat = z_sym.to(self.dtype) * qs + med
        hyper = self.h_s(z_hat)
        y_hat, _, sym, idx = self._entropy_loop(y, hyper, "compress", g, rec)
        out = {"symbols": sym, "indexes": idx, "z_symbols": z_sym, "y_hat": y_hat, "x_hat": self.g_s(y_hat)}
        if trace:
            rec.update(y=y, z=z, hyper=hyper)
            out["trace"] = rec
        return out

    @torch.no_grad()
    def decoder_forward(self, x):
        """MLICPlusPlus.net_decoder_forward (mlicpp.py:380-459): decoder walk with z_hat = 0."""
        B, _, H, W = x.shape
        z_hat = torch.zeros(B, self.N, H // 64, W // 64, dtype=self.dtype)
        hyper = self.h_s(z_hat)
        y_hat, _, _, _ = self._entropy_loop(None, hyper, "decoder")
        return self.g_s(y_hat)


def rd_stats(out, x):
    """loss/rd_loss.py:37-48 -- (bpp, mse, psnr) with the unpadded-target pixel count."""
    Nb, _, H, W = x.shape
    npx = Nb * H * W
    bpp = sum(float(torch.log(l.double()).sum()) for l in out["likelihoods"].values()) / (-math.log(2) * npx)
    mse = float(((out["x_hat"].double() - x.double()) ** 2).mean())
    psnr = 10.0 * math.log10(1.0 / mse) if mse > 0 else float("inf")
    return bpp, mse, psnr
