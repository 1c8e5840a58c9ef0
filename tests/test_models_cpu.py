"""Host-side mirror of the reference model API (no GPU needed): registry, state_dict layout, errors."""
import numpy as np
import pytest
import torch

import mlic_b200
from mlic_b200 import _lib, params
from mlic_b200.dist import shard_range


def test_registry_matches_reference_configs():
    # config/config.py:19-62
    assert {k: (v["N"], v["M"], v["slice_num"]) for k, v in mlic_b200.MODEL_TABLE.items()} == {
        "MLICPP_L": (192, 320, 10), "MLICPP_M": (160, 256, 8), "MLICPP_S": (96, 160, 5), "MLICPP_S2": (128, 128, 2),
        "MLICPP_M_SMALL_DEC": (192, 320, 10), "MLICPP_S_VBR": (96, 160, 5), "MLICPP_L_VBR": (192, 320, 10),
        "MLICPP_M_SMALL_DEC_VBR": (192, 320, 10)}
    with pytest.raises(KeyError):
        mlic_b200.get_model("MLICPP_XL")


@pytest.mark.parametrize("name,nparams,nkeys", [("MLICPP_L", 41.72e6, 1261), ("MLICPP_S", 11.79e6, 711),
                                                 ("MLICPP_M_SMALL_DEC", 25.07e6, 1251), ("MLICPP_L_VBR", None, 1268),
                                                 ("MLICPP_M_SMALL_DEC_VBR", 25072266, 1258)])
def test_state_dict_layout(name, nparams, nkeys):
    net = mlic_b200.get_model(name)
    sd = net.state_dict()
    ent = params.build_entries(name)
    assert set(sd) == set(ent) and len(sd) == nkeys
    for k, e in ent.items():
        assert tuple(sd[k].shape) == e.shape, k
    if nparams:                                       # SURVEY.md section 6 [probe] parameter counts of the reference
        assert sum(p.numel() for p in net.parameters()) == pytest.approx(nparams, rel=2e-3)
    # reference attribute surface (mlicpp.py:23-76)
    assert net.slice_ch * net.slice_num == net.M
    assert len(net.local_context) == net.slice_num
    conv1 = net.g_a.analysis_transform[0].conv1
    if name.startswith("MLICPP_M_SMALL_DEC"):         # dense encoder (analysis_old.py:10-16)
        assert conv1.weight.shape == (net.N, 3, 3, 3)
    else:                                             # depthwise-separable (conv.py:46-63)
        assert conv1.depth_conv.weight.shape == (3, 1, 3, 3) and conv1.point_conv.weight.shape == (net.N, 3, 1, 1)


def test_load_state_dict_and_update():
    net = mlic_b200.get_model("MLICPP_S")
    assert net.gaussian_conditional.scale_table.numel() == 0
    assert net.update(force=True) is True
    t = net.gaussian_conditional.scale_table
    assert t.numel() == 64 and float(t[0]) == pytest.approx(0.11, rel=1e-6) and float(t[-1]) == pytest.approx(256, rel=1e-5)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    other = mlic_b200.get_model("MLICPP_S")
    other.load_state_dict(sd)                         # resizes the empty scale_table buffer like mlicpp.py:461-468
    for k, v in other.state_dict().items():
        assert torch.equal(v, sd[k]), k
    bad = dict(sd)
    bad.pop("g_a.analysis_transform.6.point_conv.bias")
    with pytest.raises(RuntimeError):
        other.load_state_dict(bad)


def test_sd_vbr_surface():
    """models/mlicpp_sd_vbr.py:92-110: 5 levels, own gain / lambda tables, the Vbr call signatures."""
    net = mlic_b200.get_model("MLICPP_M_SMALL_DEC_VBR")
    assert isinstance(net, mlic_b200.MLICPlusPlusSDVbr) and isinstance(net, mlic_b200.MLICPlusPlusVbr)
    assert net.levels == 5 and net.lmbda == [0.0002, 0.0005, 0.0035, 0.0483, 0.18]
    np.testing.assert_allclose(net.Gain.detach().numpy(), [0.002424, 0.06556, 0.13944, 0.51801, 1.0], rtol=1e-6)
    assert net.lrp_anchor[0].lrp_transform[6].point_conv.weight.shape[0] == net.slice_ch      # LRP-Old pyramid, 4 convs
    with pytest.raises(AssertionError):
        net._scale(5, 0, True)                                       # mlicpp_sd_vbr.py compress: s in range(levels)
    with pytest.raises(ValueError):
        net.forward(torch.zeros(1, 3, 64, 64), stage=0)


def test_vbr_surface():
    net = mlic_b200.get_model("MLICPP_S_VBR")
    assert net.levels == 6 and net.no_quantoffset is True
    np.testing.assert_allclose(net.Gain.detach().numpy(), [0.06556, 0.13944, 0.19293, 0.37268, 0.51801, 1.0], rtol=1e-6)
    with pytest.raises(ValueError):
        net.forward(torch.zeros(1, 3, 64, 64), stage=3)
    with pytest.raises(AssertionError):
        net._scale(7, 0, True)
    assert net._scale(9, 0, False) == pytest.approx(1.0)          # forward clips the level (mlicpp_vbr.py:123)
    assert net._scale(0, 0.25, True) == 0.25


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    net = mlic_b200.get_model("MLICPP_S")
    with pytest.raises(_lib.MlicError):
        net(torch.zeros(1, 3, 64, 64))
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 60, 64))
    with pytest.raises(_lib.MlicError):                      # decompress runs on the CUDA engine too
        net.decompress([[b""], [b""]], (1, 1))


def test_aux_loss_matches_compressai_definition():
    net = mlic_b200.get_model("MLICPP_S")
    loss = net.aux_loss()
    assert loss.requires_grad and float(loss) > 0
    loss.backward()
    assert net.entropy_bottleneck.quantiles.grad is not None


def test_shard_range():
    for n in (1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_set_precision_accepts_per_stage_tuples():
    """set_precision: "bf16" | "fp32" | "mixed" | (g_a, entropy model, g_s); a uniform tuple collapses to its value; nothing runs on CPU."""
    net = mlic_b200.get_model("MLICPP_S")
    assert net.set_precision("mixed").precision == ("fp32", "fp32", "bf16")
    assert net.set_precision(("bf16", "fp32", "bf16")).precision == ("bf16", "fp32", "bf16")
    assert net.set_precision(["fp32", "fp32", "fp32"]).precision == "fp32"
    for bad in ("fp16", ("fp32", "bf16"), ("fp32", "bf16", "tf32"), 3):
        with pytest.raises((ValueError, TypeError)):
            net.set_precision(bad)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.MlicError):                  # the staged call has no CPU fallback either
            net.set_precision("mixed")(torch.zeros(1, 3, 64, 64))


def test_variable_rate_hyper_prior_model_layout_and_tables():
    """MLICPlusPlusVbr(config, vr_entbttlnck=True): the extra state_dict entries of mlicpp_vbr.py:105-117, the gain -> step network, and
    the per-step z tables (EntropyBottleneckVbr.update_variable): a finer grid has more bins and still sums to 2^16 per channel."""
    import numpy as np
    from conftest import vr_model
    from mlic_b200 import coder
    net = vr_model()
    keys = list(net.state_dict().keys())
    assert keys[-7:] == [f"gayn2zqstep.{j}.{w}" for j in (0, 2, 4) for w in ("weight", "bias")] + ["lower_bound_zqstep.bound"]
    assert tuple(net.gayn2zqstep[0].weight.shape) == (10, 1) and tuple(net.gayn2zqstep[4].weight.shape) == (1, 10)
    assert "gayn2zqstep.0.weight" not in mlic_b200.get_model("MLICPP_S_VBR").state_dict()
    qs = [net._zqstep(net._scale(s, 0, True)) for s in range(6)]
    assert all(q >= 0.5 for q in qs) and mlic_b200.get_model("MLICPP_S_VBR")._zqstep(0.5) == 1.0
    with pytest.raises(ValueError):
        mlic_b200.models.MLICPlusPlus(mlic_b200.models.model_config("MLICPP_S"), name="MLICPP_S", _vr_entbttlnck=True)
    c1, l1, o1 = coder.bottleneck_tables(net.entropy_bottleneck, 1.0)
    c2, l2, o2 = coder.bottleneck_tables(net.entropy_bottleneck, 0.5)
    assert (l2 > l1).all() and (o2 <= o1).all()
    for cdf, ln in ((c1, l1), (c2, l2)):
        for ch in range(0, cdf.shape[0], 17):
            assert cdf[ch, 0] == 0 and cdf[ch, ln[ch] - 1] == 65536 and (np.diff(cdf[ch, :ln[ch]]) > 0).all()
