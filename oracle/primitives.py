"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the third-party primitives the
reference's structured-covariance hot path calls (`qmcpy`, absent from this image).

Nothing under ``fastgaussianprocesses_b200/`` may import this module.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs use it,
and only as the checker or the thing timed as the CPU baseline.

PARITY STATUS: "parity partially pinned".
  * `qmcpy` (pyproject.toml:39 ``>=1.6.3b0``; conda_env.yml:149 ``==1.6.2.1``) is not installable here,
    so its primitives are RESTATED from what the reference itself pins:
      - transforms: the doubling recursion of fastgps/util.py:121-126 and :173-178 with
        ``get_omega`` (fast_gp_lattice.py:261-262, fast_gp_digital_net_b2.py:264-265) and the
        length-1 identity (util.py:170) define ft uniquely for every power of two;
      - lattice kernel: fast_gp_lattice.py:267-273 + the standard Bernoulli polynomials;
      - net kernel: fast_gp_digital_net_b2.py:289-301 + the Walsh-series definition of the
        digitally-shift-invariant kernel of smoothness alpha (checked here against a brute-force
        Walsh sum, tests/test_oracle.py);
      - point ORDER: "NATURAL" = radical-inverse order (the only order for which
        util.py:345-352's dense==fast identity holds).
  * Generating vectors / matrices and the RNG stream of the random shift are qmcpy data and are
    NOT pinned: every function here takes them as explicit inputs and is bit-exact *given them*.
  * The reference's own orchestration (fastgps/*.py, unmodified) is run on top of these
    primitives by tests/golden/make_golden.py to produce the committed fixtures.
"""
from fractions import Fraction
from math import comb

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# points
# --------------------------------------------------------------------------------------

def bitrev64(i: np.ndarray) -> np.ndarray:
    """Reverse the 64 bits of each uint64 (radical inverse numerator over 2^64)."""
    v = np.asarray(i, dtype=np.uint64).copy()
    masks = [
        (0x5555555555555555, 1), (0x3333333333333333, 2), (0x0F0F0F0F0F0F0F0F, 4),
        (0x00FF00FF00FF00FF, 8), (0x0000FFFF0000FFFF, 16), (0x00000000FFFFFFFF, 32),
    ]
    for mk, sh in masks:
        mk = np.uint64(mk)
        sh = np.uint64(sh)
        v = ((v >> sh) & mk) | ((v & mk) << sh)
    return v


def lattice_points(z, shift, n_min: int, n_max: int) -> np.ndarray:
    """Rank-1 lattice in NATURAL (radical-inverse) order with a shift mod 1.

    x[i,j] = ( frac(phi_2(i) * z_j) + shift_j ) mod 1, phi_2 the base-2 radical inverse.
    Follows the structure abstract_gp.py:307-309 requires of ``seq(n_min,n_max)`` (float64 (n,d) in [0,1))
    and SURVEY App. B.1.  frac(phi_2(i) z_j) is evaluated exactly:
    phi_2(i) = bitrev64(i)/2^64, so frac(.) = ((z_j * bitrev64(i)) mod 2^64) / 2^64 and the
    wrap-around uint64 product has at most ceil(log2(i+1)) significant bits (exact in float64 for i < 2^53).
    The only rounding is the single IEEE add of the shift; values >= 1 are reduced by an exact subtraction.
    """
    z = np.asarray(z, dtype=np.uint64).reshape(1, -1)
    shift = np.asarray(shift, dtype=np.float64).reshape(1, -1)
    i = np.arange(n_min, n_max, dtype=np.uint64)
    r = bitrev64(i).reshape(-1, 1)
    with np.errstate(over="ignore"):
        prod = r * z  # wraps mod 2^64
    # exact conversion: prod is a multiple of 2^(64-mbits) with <= 53 significant bits
    frac = (prod >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    s = frac + shift
    s = np.where(s >= 1.0, s - 1.0, s)
    return s


def dnb2_points(C, dshift, t: int, n_min: int, n_max: int):
    """Base-2 digital net in NATURAL order with a digital shift.

    xb[i,j] = XOR_{k : bit k of i set} C[j,k]  XOR dshift[j]   (t-bit integers, t < 64)
    x = xb * 2^-t (fast_gp_digital_net_b2.py:266-273, int64 -> float64 conversion rounds to nearest even).
    """
    C = np.asarray(C, dtype=np.uint64)
    d, mmax = C.shape
    dshift = np.asarray(dshift, dtype=np.uint64).reshape(1, d)
    i = np.arange(n_min, n_max, dtype=np.uint64)
    xb = np.zeros((len(i), d), dtype=np.uint64)
    for k in range(mmax):
        bit = ((i >> np.uint64(k)) & np.uint64(1)).astype(bool)
        if not bit.any():
            continue
        xb[bit] ^= C[:, k].reshape(1, d)
    assert n_max <= (1 << mmax)
    xb ^= dshift
    xb = xb.astype(np.int64)
    x = xb.astype(np.float64) * 2.0 ** (-t)
    return xb, x


DEFAULT_LATTICE_Z = [1, 182667, 469891, 498753, 110745, 446247, 250185, 118627]


def default_lattice_gen_vec(d: int) -> np.ndarray:
    """A usable (NOT qmcpy's) default generating vector: odd integers < 2^20."""
    z = list(DEFAULT_LATTICE_Z[:d])
    if d > len(z):
        rng = np.random.Generator(np.random.PCG64(20201))
        extra = rng.integers(1, 2 ** 19, size=d - len(z)) * 2 + 1
        z += [int(v) for v in extra]
    return np.asarray(z, dtype=np.uint64)


def default_dnb2_gen_mats(d: int, t: int, mmax: int = 32) -> np.ndarray:
    """Sobol' (Joe-Kuo) generating matrices as t-bit MSB-first column integers, from scipy's table."""
    from scipy.stats import qmc
    bits = max(mmax, 1)
    sv = qmc.Sobol(d, scramble=False, bits=bits)._sv.astype(np.uint64)  # (d,bits), value bit (bits-1-row)
    assert t >= bits
    return sv << np.uint64(t - bits)


# --------------------------------------------------------------------------------------
# transforms (defined by the reference's own doubling recursion)
# --------------------------------------------------------------------------------------

def _omega(m: int, device=None):
    # fast_gp_lattice.py:261-262
    return torch.exp(-torch.pi * 1j * torch.arange(2 ** m, device=device) / 2 ** m)


def fftbr_torch(x: torch.Tensor) -> torch.Tensor:
    """Orthonormal FFT, bit-reversed-order input -> natural-order output, along the last dim.

    Built stage by stage from util.py:121-126: ft(x[:2h]) = cat(ft(x[:h]) + w*ft(x[h:2h]), ft(x[:h]) - w*ft(x[h:2h]))/sqrt(2),
    w = omega(log2 h), length-1 transform = identity (util.py:170).  Differentiable (torch ops only).
    """
    n = x.size(-1)
    assert n & (n - 1) == 0
    m = n.bit_length() - 1
    y = x.to(torch.complex128)
    batch = y.shape[:-1]
    for s in range(m):
        h = 1 << s
        y = y.reshape(*batch, n // (2 * h), 2, h)
        w = _omega(s, device=y.device)
        a = y[..., 0, :]
        b = w * y[..., 1, :]
        y = torch.cat([a + b, a - b], dim=-1) / np.sqrt(2)
    return y.reshape(*batch, n)


def ifftbr_torch(x: torch.Tensor) -> torch.Tensor:
    """Inverse of `fftbr_torch`: natural-order input -> bit-reversed-order output (util.py:341-343 round trip)."""
    n = x.size(-1)
    assert n & (n - 1) == 0
    m = n.bit_length() - 1
    y = x.to(torch.complex128)
    batch = y.shape[:-1]
    for s in range(m - 1, -1, -1):
        h = 1 << s
        y = y.reshape(*batch, n // (2 * h), 2, h)
        wc = _omega(s, device=y.device).conj()
        u = y[..., 0, :]
        v = y[..., 1, :]
        a = (u + v) / np.sqrt(2)
        b = wc * (u - v) / np.sqrt(2)
        y = torch.cat([a, b], dim=-1)
    return y.reshape(*batch, n)


def fwht_torch(x: torch.Tensor) -> torch.Tensor:
    """Orthonormal Sylvester-ordered Walsh-Hadamard transform (same recursion with omega = 1,
    fast_gp_digital_net_b2.py:264-265); self-inverse."""
    n = x.size(-1)
    assert n & (n - 1) == 0
    m = n.bit_length() - 1
    y = x
    batch = y.shape[:-1]
    for s in range(m):
        h = 1 << s
        y = y.reshape(*batch, n // (2 * h), 2, h)
        a = y[..., 0, :]
        b = y[..., 1, :]
        y = torch.cat([a + b, a - b], dim=-1) / np.sqrt(2)
    return y.reshape(*batch, n)


# --------------------------------------------------------------------------------------
# kernels
# --------------------------------------------------------------------------------------

def bernoulli_numbers(nmax: int):
    B = [Fraction(0)] * (nmax + 1)
    B[0] = Fraction(1)
    for m in range(1, nmax + 1):
        B[m] = -sum(comb(m + 1, k) * B[k] for k in range(m)) / (m + 1)
    return B  # B_1 = -1/2 convention


def bernoulli_poly_coeffs(order: int):
    """Coefficients c[k] of x^k in B_order(x) as exact fractions."""
    B = bernoulli_numbers(order)
    return [comb(order, order - k) * B[order - k] for k in range(order + 1)]


BERNOULLIPOLYSDICT = {n: bernoulli_poly_coeffs(n) for n in range(1, 11)}


def bernoulli_poly(order: int, x):
    """Standard Bernoulli polynomial B_order(x), Horner in float64 (fast_gp_lattice.py:273 call site)."""
    c = [float(v) for v in bernoulli_poly_coeffs(int(order))]
    y = 0 * x + c[-1]
    for k in range(len(c) - 2, -1, -1):
        y = y * x + c[k]
    return y


WALSH_AT_ZERO = {2: 5.0 / 2.0, 3: 43.0 / 18.0, 4: 701.0 / 294.0}


def _digit_sign_sum(xb: torch.Tensor, t: int) -> torch.Tensor:
    """sum_{a>=0} (-1)^{x_{a+1}} 2^{-3a}, x_{a+1} the (a+1)-th binary digit of xb*2^-t (digits past t are 0)."""
    total = torch.full(xb.shape, 8.0 / 7.0, dtype=torch.float64, device=xb.device)
    for a in range(min(t, 22)):  # 2*8^-22 < 2^-64: below float64 resolution of the O(1) sum
        bit = (xb >> (t - 1 - a)) & 1
        total = total - 2.0 * bit.to(torch.float64) * 8.0 ** (-a)
    return total


def weighted_walsh_funcs(alpha: int, xb: torch.Tensor, t: int) -> torch.Tensor:
    """W_alpha(x) = sum_{k>=0} wal_k(x) 2^{-mu_alpha(k)} for x = xb*2^-t, alpha in {2,3,4} (closed forms, SURVEY App. B.2;
    call site fast_gp_digital_net_b2.py:300)."""
    assert alpha in (2, 3, 4)
    assert not torch.is_floating_point(xb)
    xf = xb.to(torch.float64) * 2.0 ** (-t)
    zero = xb == 0
    safe = torch.where(zero, torch.ones_like(xb), xb)
    # beta = -floor(log2 xf) = t - floor(log2 xb), from the integer bit length (exact)
    fl = torch.zeros_like(safe)
    tmp = safe.clone()
    for sh in (32, 16, 8, 4, 2, 1):
        big = tmp >= (1 << sh)
        fl = fl + big.to(fl.dtype) * sh
        tmp = torch.where(big, tmp >> sh, tmp)
    beta = (t - fl).to(torch.float64)
    p1 = 1 - 2.0 ** (-beta)
    if alpha == 2:
        w = -beta * xf + 2.5 * p1
    elif alpha == 3:
        p2 = 1 - 2.0 ** (-2 * beta)
        w = beta * xf ** 2 - 5 * p1 * xf + (43.0 / 18.0) * p2
    else:
        p2 = 1 - 2.0 ** (-2 * beta)
        p3 = 1 - 2.0 ** (-3 * beta)
        s = _digit_sign_sum(xb, t)
        w = (-(2.0 / 3.0) * beta * xf ** 3 + 5 * p1 * xf ** 2 - (43.0 / 9.0) * p2 * xf
             + (701.0 / 294.0) * p3 + beta * (s / 48.0 - 1.0 / 42.0))
    return torch.where(zero, torch.full_like(w, WALSH_AT_ZERO[alpha]), w)


def walsh_series_bruteforce(alpha: int, xb: int, t: int, extra_bits: int = 8) -> float:
    """Direct (slow) evaluation of sum_k wal_k(x) 2^{-mu_alpha(k)}, truncated at k < 2^(t+extra_bits). Tests only."""
    K = 1 << (t + extra_bits)
    k = np.arange(K, dtype=np.int64)
    # wal_k(x) = (-1)^{sum_i k_i x_{i+1}}: bit i of k (LSB=0) pairs with digit i+1 of x = bit (t-1-i) of xb
    sgn = np.zeros(K, dtype=np.int64)
    for i in range(t):
        if (xb >> (t - 1 - i)) & 1:
            sgn ^= (k >> i) & 1
    wal = 1.0 - 2.0 * sgn
    mu = np.zeros(K, dtype=np.int64)
    kk = k.copy()
    for _ in range(alpha):
        nz = kk > 0
        top = np.zeros(K, dtype=np.int64)
        top[nz] = np.floor(np.log2(kk[nz])).astype(np.int64) + 1
        mu += top
        kk = np.where(nz, kk - (np.int64(1) << np.maximum(top - 1, 0)), kk)
        kk = np.where(nz, kk, 0)
    return float(np.sum(wal * 2.0 ** (-mu.astype(np.float64))))
