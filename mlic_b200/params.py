"""Parameter inventory of the MLIC++ model family.

The drop-in boundary is the reference's `state_dict`: same dotted names, same shapes
(reference: MLIC++/models/mlicpp.py:13-76, mlicpp_small_decoder.py:16-83,
mlicpp_vbr.py:14-117, plus the CompressAI 1.2.6 modules those instantiate).  This module
derives that inventory from a model name; `models.py` materialises it as nn.Parameters and
the CUDA engine consumes the tensors by name.
"""
import math
from collections import OrderedDict

# reference: MLIC++/config/config.py:19-62 ; MLICPP_L_VBR per SURVEY.md F4 (Vbr class over the L config)
MODEL_TABLE = {
    "MLICPP_L": dict(N=192, M=320, slice_num=10, kind="base"),
    "MLICPP_M": dict(N=160, M=256, slice_num=8, kind="base"),
    "MLICPP_S": dict(N=96, M=160, slice_num=5, kind="base"),
    "MLICPP_S2": dict(N=128, M=128, slice_num=2, kind="base"),
    "MLICPP_M_SMALL_DEC": dict(N=192, M=320, slice_num=10, kind="sd"),
    "MLICPP_S_VBR": dict(N=96, M=160, slice_num=5, kind="vbr"),
    "MLICPP_L_VBR": dict(N=192, M=320, slice_num=10, kind="vbr"),
    "MLICPP_M_SMALL_DEC_VBR": dict(N=192, M=320, slice_num=10, kind="sdvbr"),   # mlicpp_sd_vbr.py:19-127
}
KIND_CODE = {"base": 0, "sd": 1, "vbr": 2, "sdvbr": 3}
VBR_GAINS = (0.06556, 0.13944, 0.19293, 0.37268, 0.51801, 1.00000)   # mlicpp_vbr.py:86-91
VBR_LAMBDAS = (0.0005, 0.0035, 0.0067, 0.025, 0.0483, 0.18)          # mlicpp_vbr.py:83
SDVBR_GAINS = (0.002424, 0.06556, 0.13944, 0.51801, 1.00000)          # mlicpp_sd_vbr.py:95-100 (5 levels)
SDVBR_LAMBDAS = (0.0002, 0.0005, 0.0035, 0.0483, 0.18)                # mlicpp_sd_vbr.py:92


class Entry:
    """One state_dict entry: shape, whether it is a Parameter, and how to initialise it."""
    __slots__ = ("shape", "is_param", "init", "dtype")

    def __init__(self, shape, is_param=True, init=("uniform_fan", None), dtype="float32"):
        self.shape, self.is_param, self.init, self.dtype = tuple(shape), is_param, init, dtype


class _Builder:
    def __init__(self):
        self.e = OrderedDict()

    def conv(self, p, cin, cout, k, groups=1):
        fan = (cin // groups) * k * k
        self.e[p + ".weight"] = Entry((cout, cin // groups, k, k), init=("uniform_fan", fan))
        self.e[p + ".bias"] = Entry((cout,), init=("uniform_fan", fan))

    def linear(self, p, cin, cout):
        self.e[p + ".weight"] = Entry((cout, cin), init=("uniform_fan", cin))
        self.e[p + ".bias"] = Entry((cout,), init=("uniform_fan", cin))

    def layernorm(self, p, c):
        self.e[p + ".weight"] = Entry((c,), init=("const", 1.0))
        self.e[p + ".bias"] = Entry((c,), init=("const", 0.0))

    def dsconv(self, p, cin, cout):
        """DepthWiseConv: depthwise 3x3 (+bias) then pointwise 1x1 (+bias)  (modules/layers/conv.py:46-63)."""
        self.conv(p + ".depth_conv", cin, cin, 3, groups=cin)
        self.conv(p + ".point_conv", cin, cout, 1)

    def conv3(self, p, cin, cout, dense=False):
        self.conv(p, cin, cout, 3) if dense else self.dsconv(p, cin, cout)

    def subpel(self, p, cin, cout):
        self.conv(p + ".0", cin, cout * 4, 3)

    def gdn(self, p, c):
        ped = 2.0 ** -36
        self.e[p + ".beta"] = Entry((c,), init=("const", math.sqrt(1.0 + ped)))
        self.e[p + ".gamma"] = Entry((c, c), init=("gdn_gamma", None))
        for r, minimum in (("beta_reparam", 1e-6), ("gamma_reparam", 0.0)):
            self.e[f"{p}.{r}.pedestal"] = Entry((1,), False, ("const", ped))
            self.e[f"{p}.{r}.lower_bound.bound"] = Entry((1,), False, ("const", math.sqrt(minimum + ped)))

    def res_block(self, p, cin, cout, dense=False):
        self.conv3(p + ".conv1", cin, cout, dense)
        self.conv3(p + ".conv2", cout, cout, dense)
        if cin != cout:
            self.conv(p + ".skip", cin, cout, 1)

    def res_block_stride(self, p, cin, cout, dense=False):
        self.conv3(p + ".conv1", cin, cout, dense)
        self.conv3(p + ".conv2", cout, cout, dense)
        self.gdn(p + ".gdn", cout)
        self.conv(p + ".skip", cin, cout, 1)

    def res_block_up(self, p, cin, cout):
        self.subpel(p + ".subpel_conv", cin, cout)
        self.conv3(p + ".conv", cout, cout)
        self.gdn(p + ".igdn", cout)
        self.subpel(p + ".upsample", cin, cout)

    def qkv_branch(self, p, dim):
        self.conv(p + ".0", dim, dim, 1)
        self.conv(p + ".1", dim, dim, 3, groups=dim)

    def dw_mlp(self, p, cin, hidden, cout):
        self.conv(p + ".0", cin, hidden, 1)
        self.conv(p + ".2", hidden, hidden, 3, groups=hidden)
        self.conv(p + ".4", hidden, cout, 1)


def build_entries(name, vr_entbttlnck=False):
    """name -> OrderedDict[str, Entry] covering every state_dict key of the reference model.  vr_entbttlnck (Vbr kinds only): the
    variable-rate hyper-prior branch of mlicpp_vbr.py:104-117 -- its gain -> z quantisation-step network and lower bound."""
    cfg = MODEL_TABLE[name]
    N, M, S, kind = cfg["N"], cfg["M"], cfg["slice_num"], cfg["kind"]
    C = M // S
    assert C * S == M, "M must be divisible by slice_num"      # mlicpp.py:21
    sd = kind in ("sd", "sdvbr")
    b = _Builder()
    e = b.e

    # CompressionModel / EntropyBottleneck(N) (CompressAI): filters (1,3,3,3,3,1)
    e["entropy_bottleneck.quantiles"] = Entry((N, 1, 3), init=("eb_quantiles", None))
    for nm in ("_offset", "_quantized_cdf", "_cdf_length"):
        e["entropy_bottleneck." + nm] = Entry((0,), False, ("empty", None), "int32")
    e["entropy_bottleneck.target"] = Entry((3,), False, ("eb_target", None))
    e["entropy_bottleneck.likelihood_lower_bound.bound"] = Entry((1,), False, ("const", 1e-9))
    filt = (1, 3, 3, 3, 3, 1)
    for i in range(5):
        e[f"entropy_bottleneck.matrices.{i}"] = Entry((N, filt[i + 1], filt[i]), init=("eb_matrix", filt[i + 1]))
    for i in range(5):
        e[f"entropy_bottleneck.biases.{i}"] = Entry((N, filt[i + 1], 1), init=("uniform", 0.5))
    for i in range(4):
        e[f"entropy_bottleneck.factors.{i}"] = Entry((N, filt[i + 1], 1), init=("const", 0.0))

    # g_a  (analysis.py:9-17 ; analysis_old.py:10-16 dense for SD)
    p = "g_a.analysis_transform."
    cin = 3
    for i in (0, 2, 4):
        b.res_block_stride(p + str(i), cin, N, sd)
        b.res_block(p + str(i + 1), N, N, sd)
        cin = N
    b.conv3(p + "6", N, M, sd)
    # h_a  (analysis.py:33-43)
    p = "h_a.reduction."
    for i, ci in zip((0, 2, 4, 6, 8), (M, N, N, N, N)):
        b.conv3(p + str(i), ci, N, sd)
    # g_s  (synthesis.py:59-68); SD: SynthesisTransform(N//4, M)  (mlicpp_small_decoder.py:36)
    Ns = N // 4 if sd else N
    p = "g_s.synthesis_transform."
    b.res_block(p + "0", M, M)
    b.res_block_up(p + "1", M, Ns)
    for i in (2, 4):
        b.res_block(p + str(i), Ns, Ns)
        b.res_block_up(p + str(i + 1), Ns, Ns)
    b.res_block(p + "6", Ns, Ns)
    b.subpel(p + "7", Ns, 3)
    # h_s  (synthesis.py:18-28); SD: HyperSynthesis(M//4, N)  (mlicpp_small_decoder.py:37)
    Mh = M // 4 if sd else M
    p = "h_s.increase."
    b.conv3(p + "0", N, Mh)
    b.subpel(p + "2", Mh, Mh)
    b.conv3(p + "4", Mh, Mh * 3 // 2)
    b.subpel(p + "6", Mh * 3 // 2, Mh * 3 // 2)
    b.conv3(p + "8", Mh * 3 // 2, Mh * 2)

    # GaussianConditional(None)
    for nm in ("_offset", "_quantized_cdf", "_cdf_length"):
        e["gaussian_conditional." + nm] = Entry((0,), False, ("empty", None), "int32")
    e["gaussian_conditional.scale_table"] = Entry((0,), False, ("empty", None))
    e["gaussian_conditional.scale_bound"] = Entry((1,), False, ("const", 0.11))
    e["gaussian_conditional.likelihood_lower_bound.bound"] = Entry((1,), False, ("const", 1e-9))
    e["gaussian_conditional.lower_bound_scale.bound"] = Entry((1,), False, ("const", 0.11))

    # LocalContext x S  (context.py:11-41)
    for i in range(S):
        p = f"local_context.{i}"
        e[p + ".relative_position_table"] = Entry((81, 2), init=("trunc_normal", 0.02))
        e[p + ".relative_position_index"] = Entry((25, 25), False, ("relpos_index", None), "int64")
        b.linear(p + ".qkv_proj", C, 3 * C)
        b.linear(p + ".proj", 2 * C, 2 * C)
        b.linear(p + ".mlp.fc1", 2 * C, 4 * C)
        b.linear(p + ".mlp.fc2", 4 * C, 2 * C)
        b.layernorm(p + ".norm1", C)
        b.layernorm(p + ".norm2", 2 * C)
        b.conv(p + ".fusion", C, 2 * C, 5)
    # ChannelContext  (context.py:115-127); SD: hidden [96,96], dense (context_old.py:120-126)
    hid = (96, 96) if sd else (192, 128)
    for i in range(1, S):
        p = f"channel_context.{i}.fushion."
        b.conv3(p + "0", C * i, hid[0], sd)
        b.conv3(p + "2", hid[0], hid[1], sd)
        b.conv3(p + "4", hid[1], C * 4, sd)
    # LinearGlobalInterContext(dim=C*i, out_dim=2C)  (context.py:195-224)
    for i in range(1, S):
        p = f"global_inter_context.{i}"
        D, O = C * i, 2 * C
        for br in ("keys", "queries", "values"):
            b.qkv_branch(f"{p}.{br}", D)
        b.conv(p + ".reprojection", D, O * 3 // 2, 5)
        b.dw_mlp(p + ".mlp", O * 3 // 2, O * 2, O)
        b.conv(p + ".skip", O * 3 // 2, O, 1)
    # LinearGlobalIntraContext(dim=C)  (context.py:140-167)
    for i in range(1, S):
        p = f"global_intra_context.{i}"
        for br in ("keys", "queries", "values"):
            b.qkv_branch(f"{p}.{br}", C)
        b.conv(p + ".reprojection", C, 2 * C, 5)
        b.dw_mlp(p + ".mlp", 2 * C, 4 * C, 2 * C)
    # EntropyParameters  (mlicpp.py:57-66 ; entropy.py:10-18); SD uses M//4 (mlicpp_small_decoder.py:39)
    Me = M // 4 if sd else M
    for tag, extra0, extra in (("anchor", 0, 6), ("nonanchor", 2, 10)):
        for i in range(S):
            p = f"entropy_parameters_{tag}.{i}.fusion."
            cin = Me * 2 + C * (extra if i else extra0)
            for j, (ci, co) in zip((0, 2, 4, 6), ((cin, 320), (320, 256), (256, 128), (128, 2 * C))):
                b.conv(p + str(j), ci, co, 1)
    # LatentResidualPrediction (quantization.py:30-45); SD: ...Old pyramid (quantization.py:9-23)
    for tag in ("anchor", "nonanchor"):
        for i in range(S):
            p = f"lrp_{tag}.{i}.lrp_transform."
            cin = Me + (i + 1) * C
            if sd:
                d = abs(C - cin)
                widths = (cin, cin - d // 4, cin - d // 2, cin - d * 3 // 4, C)
                idx = (0, 2, 4, 6)
            else:
                widths = (cin, 224, 128, C)
                idx = (0, 2, 4)
            for j, ci, co in zip(idx, widths[:-1], widths[1:]):
                b.dsconv(p + str(j), ci, co)
    # VBR extras (mlicpp_vbr.py:83-101)
    if kind in ("vbr", "sdvbr"):                       # mlicpp_sd_vbr.py:92-110 for the SD variant
        e["Gain"] = Entry((len(SDVBR_GAINS if sd else VBR_GAINS),), init=("vbr_gain", None))
        for j, (ci, co) in zip((0, 2, 4), ((2, 12), (12, 12), (12, 1))):
            b.linear(f"QuantABCD.{j}", ci, co)
        if vr_entbttlnck:                              # mlicpp_vbr.py:105-117 / mlicpp_sd_vbr.py:114-126
            for j, (ci, co) in zip((0, 2, 4), ((1, 10), (10, 10), (10, 1))):
                b.linear(f"gayn2zqstep.{j}", ci, co)
            e["lower_bound_zqstep.bound"] = Entry((1,), False, ("const", 0.5))
    elif vr_entbttlnck:
        raise ValueError("vr_entbttlnck belongs to the Vbr models (mlicpp_vbr.py, mlicpp_sd_vbr.py)")
    return e
