from .ilrma import GaussILRMA, tILRMA  # noqa: F401
from .iva import AuxLaplaceIVA, AuxGaussIVA  # noqa: F401
