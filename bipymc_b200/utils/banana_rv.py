"""bipymc/utils/banana_rv.py mirror."""
from ..targets import Banana_2D  # noqa: F401
