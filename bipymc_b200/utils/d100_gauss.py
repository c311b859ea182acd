"""bipymc/utils/d100_gauss.py mirror."""
from ..targets import Gauss_100D  # noqa: F401
