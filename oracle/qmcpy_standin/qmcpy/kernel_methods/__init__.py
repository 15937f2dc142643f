"""Stand-in for qmcpy.kernel_methods (tests only): restated in oracle/primitives.py."""
from oracle.primitives import bernoulli_poly, weighted_walsh_funcs  # noqa: F401
from . import shift_invar_ops  # noqa: F401
from . import util  # noqa: F401
