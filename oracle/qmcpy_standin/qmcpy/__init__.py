"""TEST INFRASTRUCTURE ONLY -- a stand-in for the third-party `qmcpy` package (absent from this image; no network),
exposing exactly the symbols the reference `fastgps` touches so that the UNMODIFIED reference
under /root/reference can be imported to generate golden fixtures (tests/golden/make_golden.py).

It is NOT qmcpy: generating vectors / matrices and shifts are this repo's own defaults (explicitly overridable),
and the primitives are the restatements in oracle/primitives.py.  Never imported by the product package.
"""
import numpy as np

from oracle import primitives as _p
from . import kernel_methods  # noqa: F401
from . import discrete_distribution  # noqa: F401
from .discrete_distribution import AbstractDiscreteDistribution

DiscreteDistribution = AbstractDiscreteDistribution

fftbr_torch = _p.fftbr_torch
ifftbr_torch = _p.ifftbr_torch
fwht_torch = _p.fwht_torch


def _rng(seed):
    if isinstance(seed, np.random.SeedSequence):
        return np.random.Generator(np.random.PCG64(seed))
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


class Lattice(AbstractDiscreteDistribution):
    def __init__(self, dimension=1, replications=None, seed=None, randomize="SHIFT", generating_vector=None,
                 order="NATURAL", shift=None):
        super().__init__(dimension=dimension, replications=replications, seed=seed, d_limit=np.inf, n_limit=np.inf)
        self.order = str(order).upper()
        self.randomize = str(randomize).upper()
        if self.randomize == "TRUE":
            self.randomize = "SHIFT"
        self.gen_vec = _p.default_lattice_gen_vec(self.d) if generating_vector is None else np.asarray(generating_vector, dtype=np.uint64)
        if shift is not None:
            self.shift = np.asarray(shift, dtype=np.float64)
        elif self.randomize == "SHIFT":
            self.shift = _rng(seed).random(self.d)
        else:
            self.shift = np.zeros(self.d)

    def __call__(self, n=None, n_min=None, n_max=None, **kw):
        if n is not None:
            n_min, n_max = 0, n
        assert self.order == "NATURAL"
        return _p.lattice_points(self.gen_vec, self.shift, int(n_min), int(n_max))


class DigitalNetB2(AbstractDiscreteDistribution):
    def __init__(self, dimension=1, replications=None, seed=None, randomize="DS", generating_matrices=None,
                 order="NATURAL", t=63, dshift=None, m_max=32):
        super().__init__(dimension=dimension, replications=replications, seed=seed, d_limit=np.inf, n_limit=np.inf)
        self.order = str(order).upper()
        self.randomize = str(randomize).upper()
        if self.randomize == "TRUE":
            self.randomize = "DS"
        self.t = int(t)
        self.gen_mats = _p.default_dnb2_gen_mats(self.d, self.t, m_max) if generating_matrices is None else np.asarray(generating_matrices, dtype=np.uint64)
        if dshift is not None:
            self.rshift = np.asarray(dshift, dtype=np.uint64)
        elif self.randomize in ("DS", "LMS_DS"):
            self.rshift = _rng(seed).integers(0, 2 ** self.t, size=self.d, dtype=np.uint64)
        else:
            self.rshift = np.zeros(self.d, dtype=np.uint64)

    def __call__(self, n=None, n_min=None, n_max=None, return_binary=False, **kw):
        if n is not None:
            n_min, n_max = 0, n
        assert self.order == "NATURAL"
        xb, x = _p.dnb2_points(self.gen_mats, self.rshift, self.t, int(n_min), int(n_max))
        return xb if return_binary else x


class IIDStdUniform(AbstractDiscreteDistribution):
    def __init__(self, dimension=1, replications=None, seed=None):
        super().__init__(dimension=dimension, replications=replications, seed=seed, d_limit=np.inf, n_limit=np.inf)
        self._g = _rng(seed)

    def __call__(self, n=None, n_min=None, n_max=None, **kw):
        if n is not None:
            n_min, n_max = 0, n
        return self._g.random((int(n_max) - int(n_min), self.d))
