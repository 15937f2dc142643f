"""Point-sequence generator specs whose samples are produced ON THE GPU by the K1 kernels.

They expose the attributes the reference reads from `qmcpy.Lattice` / `qmcpy.DigitalNetB2`
(fast_gp_lattice.py:219-223, fast_gp_digital_net_b2.py:214-225: `d`, `order`, `replications`, `randomize`, `t`) and the
call signature it uses (abstract_gp.py:308, fast_gp_digital_net_b2.py:267), so they can be passed wherever the reference
takes a qmcpy sequence.  Generating vectors / matrices are explicit, overridable inputs; the defaults are this
package's own (NOT qmcpy's data files) -- see DESIGN.md "parity status".
"""
import numpy as np
import torch

from . import _lib

# Odd integers < 2^20 (first eight: an order-2 base-2 embedded lattice rule commonly used for n <= 2^20)
_DEFAULT_Z = [1, 182667, 469891, 498753, 110745, 446247, 250185, 118627]


def _generator(seed):
    if isinstance(seed, np.random.SeedSequence):
        return np.random.Generator(np.random.PCG64(seed))
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))


def default_generating_vector(d):
    z = list(_DEFAULT_Z[:d])
    if d > len(z):
        rng = np.random.Generator(np.random.PCG64(20201))
        z += [int(v) * 2 + 1 for v in rng.integers(1, 2 ** 19, size=d - len(z))]
    return np.asarray(z, dtype=np.uint64)


def default_generating_matrices(d, t, m_max=32):
    """Sobol' (Joe-Kuo) direction numbers from scipy's table as t-bit, MSB-first column integers, shape (d, m_max)."""
    from scipy.stats import qmc
    assert 1 <= m_max <= 64 and m_max <= t < 64
    sv = qmc.Sobol(d, scramble=False, bits=m_max)._sv.astype(np.uint64)
    return sv << np.uint64(t - m_max)


class _SequenceBase(object):
    order = "NATURAL"
    replications = 1

    def __init__(self, dimension):
        assert isinstance(dimension, (int, np.integer)) and dimension >= 1
        assert dimension <= _lib.MAX_D, "dimension %d exceeds the fused-kernel limit %d" % (dimension, _lib.MAX_D)
        self.d = int(dimension)

    def __call__(self, n=None, n_min=None, n_max=None, return_binary=False, device=None, **kwargs):
        """Host-array interface of the reference's sequences (numpy out).  The GP classes use `.generate` instead."""
        if n is not None:
            n_min, n_max = 0, n
        dev = torch.device("cuda") if device is None else torch.device(device)
        x, xb = self.generate(int(n_min), int(n_max), dev)
        return (xb if return_binary else x).cpu().numpy()


class Lattice(_SequenceBase):
    """Rank-1 lattice in NATURAL (radical-inverse, extensible) order with an optional shift mod 1."""

    def __init__(self, dimension=1, seed=None, randomize="SHIFT", generating_vector=None, shift=None, order="NATURAL",
                 replications=None):
        super().__init__(dimension)
        assert str(order).upper() == "NATURAL", "only the NATURAL (radical inverse) order diagonalises the Gram matrix"
        assert replications in (None, 1)
        r = str(randomize).upper()
        self.randomize = "SHIFT" if r in ("SHIFT", "TRUE") else "FALSE"
        assert r in ("SHIFT", "TRUE", "FALSE")
        self.gen_vec = default_generating_vector(self.d) if generating_vector is None else np.asarray(generating_vector, dtype=np.uint64)
        assert self.gen_vec.shape == (self.d,)
        if shift is not None:
            self.shift = np.asarray(shift, dtype=np.float64)
        elif self.randomize == "SHIFT":
            self.shift = _generator(seed).random(self.d)
        else:
            self.shift = np.zeros(self.d)
        assert self.shift.shape == (self.d,) and ((self.shift >= 0) & (self.shift < 1)).all()

    def generate(self, n_min, n_max, device):
        x = _lib.lattice_points(self.gen_vec, self.shift, n_min, n_max, device)
        return x, x


class DigitalNetB2(_SequenceBase):
    """Base-2 digital net in NATURAL order with an optional digital shift; points are t-bit integers xb and x = xb 2^-t."""

    def __init__(self, dimension=1, seed=None, randomize="DS", generating_matrices=None, t=52, dshift=None, order="NATURAL",
                 replications=None, m_max=32):
        super().__init__(dimension)
        assert str(order).upper() == "NATURAL"
        assert replications in (None, 1)
        r = str(randomize).upper()
        assert r in ("DS", "TRUE", "FALSE"), "randomize must be 'DS' or 'FALSE' (LMS needs explicit generating_matrices)"
        self.randomize = "DS" if r in ("DS", "TRUE") else "FALSE"
        self.t = int(t)
        assert 1 <= self.t < 64
        self.gen_mats = default_generating_matrices(self.d, self.t, m_max) if generating_matrices is None else np.asarray(generating_matrices, dtype=np.uint64)
        assert self.gen_mats.ndim == 2 and self.gen_mats.shape[0] == self.d and self.gen_mats.shape[1] <= 64
        if dshift is not None:
            self.rshift = np.asarray(dshift, dtype=np.uint64)
        elif self.randomize == "DS":
            self.rshift = _generator(seed).integers(0, 2 ** self.t, size=self.d, dtype=np.uint64)
        else:
            self.rshift = np.zeros(self.d, dtype=np.uint64)
        assert self.rshift.shape == (self.d,)
        self._C_dev = {}

    def device_matrices(self, device):
        device = torch.device(device)
        C = self._C_dev.get(device)
        if C is None:
            C = torch.from_numpy(self.gen_mats.astype(np.int64)).to(device).contiguous()
            self._C_dev[device] = C
        return C

    def generate(self, n_min, n_max, device):
        xb, x = _lib.dnb2_points(self.device_matrices(device), self.rshift, self.t, n_min, n_max)
        return x, xb


class HostSequence(object):
    """Adapter for a user-supplied qmcpy-style sequence object (the reference's `seqs` argument, fast_gp_lattice.py:219-223,
    fast_gp_digital_net_b2.py:214-225).  Its points are an INPUT: they come from the user's own host generator exactly as
    in the reference (abstract_gp.py:307-309, fast_gp_digital_net_b2.py:266-269) and are copied to the GPU once."""

    def __init__(self, seq, family):
        self.seq = seq
        self.family = int(family)
        self.d = int(seq.d)
        self.order = str(seq.order).upper()
        rep = getattr(seq, "replications", 1)
        self.replications = 1 if rep is None else int(rep)
        self.randomize = str(seq.randomize).upper()
        if self.family == 1:
            self.t = int(seq.t)

    def generate(self, n_min, n_max, device):
        if self.family == 0:
            x = torch.from_numpy(np.ascontiguousarray(self.seq(n_min=int(n_min), n_max=int(n_max)), dtype=np.float64)).to(device)
            return x, x
        xb = torch.from_numpy(np.ascontiguousarray(self.seq(n_min=int(n_min), n_max=int(n_max), return_binary=True)).astype(np.int64)).to(device)
        return xb * 2 ** (-self.t), xb
