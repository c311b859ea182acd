"""bipymc/utils/dblgauss_rv.py mirror."""
from ..targets import BimodeGauss_2D  # noqa: F401
