"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (/root/reference/MLIC++).

Works only in the build container.  Puts `ref_shim/` (CompressAI + timm stand-ins) and
the reference tree on sys.path, patches torch.cuda.synchronize to a no-op on CUDA-less
hosts (mlicpp.py:200,282 call it unconditionally), and registers MLICPP_L_VBR the way
SURVEY.md F4 defines it (MLICPlusPlusVbr over model_config("MLICPP_L")).
"""
import os
import sys

import torch

REF_ROOT = os.environ.get("MLIC_REFERENCE", "/root/reference/MLIC++")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shim")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


def _prepare():
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for p in (REF_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None


def get_reference_model(name, vr_entbttlnck=False):
    _prepare()
    import config.config as cf          # reference: config/config.py:19-62
    import models as ref_models         # reference: models/__init__.py
    if vr_entbttlnck:                   # model_loader.py never passes it: the class is built directly (mlicpp_vbr.py:15, mlicpp_sd_vbr.py:20)
        cls = {"MLICPP_S_VBR": ref_models.MLICPlusPlusVbr, "MLICPP_M_SMALL_DEC_VBR": ref_models.MLICPlusPlusSDVbr}[name]
        net = cls(config=cf.model_config(name), vr_entbttlnck=True)
    elif name == "MLICPP_L_VBR":        # not registered upstream (model_loader.py:8-15)
        net = ref_models.MLICPlusPlusVbr(config=cf.model_config("MLICPP_L"))
    else:
        net = ref_models.get_model(name)
    return net.eval()


def recorded_symbols():
    """(symbols_list, indexes_list) captured by the stub rANS encoder during the last compress()."""
    _prepare()
    import compressai.ans as ans
    return ans.LAST.get("symbols"), ans.LAST.get("indexes")
