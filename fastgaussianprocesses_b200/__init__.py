"""fastgaussianprocesses_b200: B200-native structured-covariance hot path of FastGPs (drop-in API)."""
from . import _lib  # noqa: F401
from .sequences import Lattice, DigitalNetB2  # noqa: F401
from .fast_gp import FastGPLattice, FastGPDigitalNetB2  # noqa: F401
from .standard_gp import StandardGP  # noqa: F401

__version__ = "0.1.0"
