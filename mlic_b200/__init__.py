"""mlic_b200 -- B200-native (sm_100a) implementation of the MLIC++ network-forward path behind the
reference's model API (LuZWCHA/MLIC, MLIC++/models/model_loader.py).  See DESIGN.md."""
from .models import (MLICPlusPlus, MLICPlusPlusSD, MLICPlusPlusSDVbr, MLICPlusPlusVbr, get_model, get_scale_table,  # noqa: F401
                     model_config)
from .params import MODEL_TABLE  # noqa: F401

__all__ = ["get_model", "model_config", "get_scale_table", "MLICPlusPlus", "MLICPlusPlusSD", "MLICPlusPlusVbr",
           "MLICPlusPlusSDVbr",
           "MODEL_TABLE"]
