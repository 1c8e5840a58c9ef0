from .ops import LowerBound, quantize_ste, NonNegativeParametrizer  # noqa: F401
