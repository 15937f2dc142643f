"""CPU tests that PIN the oracle (oracle/) before it is trusted as the checker:
 * against the fixtures produced by the unmodified reference (tests/golden/make_golden.py);
 * against definitions (brute-force Walsh series, torch.fft on bit-reversed input, dense solves)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import primitives as P
from oracle.fgp_oracle import OracleFastGP


def rel(a, b):
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def build_oracle(g):
    fam = str(g["family"])
    if fam == "lattice":
        x = P.lattice_points(g["z"], g["shift"], 0, int(g["n"]))
        o = OracleFastGP("lattice", x, alpha=int(g["alpha"]), scale=float(g["scale0"][0]), lengthscales=g["lengthscales0"], noise=float(g["noise0"][0]))
    else:
        xb, x = P.dnb2_points(g["C"], g["dshift"], int(g["t"]), 0, int(g["n"]))
        o = OracleFastGP("dnb2", x, xb=xb, t=int(g["t"]), alpha=int(g["alpha"]), scale=float(g["scale0"][0]), lengthscales=g["lengthscales0"], noise=float(g["noise0"][0]))
    return o


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_port_matches_reference_fixture(case):
    g = load_golden(case)
    o = build_oracle(g)
    assert np.array_equal(o.x.numpy(), g["x"])  # points: bit-exact
    if "xb" in g:
        assert np.array_equal(o.xb.numpy(), g["xb"])
    o.add_y(torch.from_numpy(g["y"]))
    assert rel(o.k1parts(), g["k1parts"]) < 1e-14
    assert rel(o.lam(), g["lam0"]) < 1e-13
    assert rel(o.ytilde, g["ytilde"]) < 1e-13
    loss, norm, logdet = o.mll_loss()
    assert abs(loss.item() - float(g["loss0"])) <= 1e-12 * abs(float(g["loss0"]))
    loss.backward()
    assert rel(o.raw_scale.grad, g["grad_raw_scale0"]) < 1e-9
    assert rel(o.raw_lengthscales.grad, g["grad_raw_lengthscales0"]) < 1e-9
    o.raw_scale.grad = None
    o.raw_lengthscales.grad = None
    with torch.no_grad():
        assert rel(o.coeffs(), g["coeffs0"]) < 1e-10
    xt = torch.from_numpy(g["xtest"])[:256]
    assert rel(o.post_mean(xt), g["pmean0"][:256]) < 1e-10  # cancellation-limited (coeffs ~ 1/noise)
    assert rel(o.post_var(xt), g["pvar0"][:256]) < 1e-9
    res = o.fit(iterations=int(g["fit_iterations"]))
    assert res["iterations"] == int(g["fit_last_iteration"])
    # reference stores metric_val = -loss (abstract_gp.py:261,285)
    assert np.allclose(-res["loss_hist"], g["loss_hist"], rtol=1e-8, atol=0)
    assert rel(o.scale.detach(), g["scale1"]) < 1e-9
    assert rel(o.lengthscales.detach(), g["lengthscales1"]) < 1e-9
    assert rel(o.post_mean(xt), g["pmean1"][:256]) < 1e-8


def test_transforms_match_definitions():
    g = torch.Generator().manual_seed(1)
    for m in range(0, 11):
        n = 1 << m
        x = torch.randn(3, n, generator=g)
        br = torch.tensor([int(format(i, "0%db" % m)[::-1], 2) if m else 0 for i in range(n)])
        ref = torch.fft.fft(x[..., br].to(torch.complex128), norm="ortho")
        assert (P.fftbr_torch(x) - ref).abs().max() < 1e-13
        z = torch.randn(3, n, generator=g) + 1j * torch.randn(3, n, generator=g)
        ref = torch.fft.ifft(z, norm="ortho")[..., br]
        assert (P.ifftbr_torch(z) - ref).abs().max() < 1e-13
        H = torch.ones(1, 1)
        for _ in range(m):
            H = torch.cat([torch.cat([H, H], 1), torch.cat([H, -H], 1)], 0)
        assert (P.fwht_torch(x) - x @ H.T / np.sqrt(n)).abs().max() < 1e-12


def test_doubling_recursion_util_121_126():
    # the recursion that defines the transform in the reference (util.py:121-126)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(512, generator=g)
    a, b = P.fftbr_torch(x[:256]), P.fftbr_torch(x[256:])
    w = torch.exp(-torch.pi * 1j * torch.arange(256) / 256)
    full = torch.cat([a + w * b, a - w * b]) / np.sqrt(2)
    assert (P.fftbr_torch(x) - full).abs().max() < 1e-13


@pytest.mark.parametrize("alpha", [2, 3, 4])
def test_weighted_walsh_closed_forms_vs_bruteforce(alpha):
    t = 6
    xb = torch.arange(0, 1 << t, dtype=torch.int64)
    w = P.weighted_walsh_funcs(alpha, xb, t).numpy()
    for v in [0, 1, 2, 3, 5, 17, 31, 32, 33, 63]:
        bf = P.walsh_series_bruteforce(alpha, v, t, extra_bits=10)
        # truncation of the series at k < 2^(t+10): tail <= sum_{k>=2^16} 2^{-mu} ~ 2^-16 * O(1)
        assert abs(w[v] - bf) < 2e-4, (alpha, v, w[v], bf)


def test_bernoulli_fourier_identity():
    # c*B_{2a}(x) = 2 sum_{h>=1} cos(2 pi h x) / h^{2a}
    x = torch.linspace(0, 1, 33)
    for a in (1, 2, 3, 4):
        c = (-1) ** (a + 1) * (2 * np.pi) ** (2 * a) / float(np.prod(np.arange(1, 2 * a + 1, dtype=np.float64)))
        h = torch.arange(1, 20001, dtype=torch.float64)[:, None]
        s = 2 * (torch.cos(2 * np.pi * h * x[None, :]) / h ** (2 * a)).sum(0)
        tol = 2e-4 if a == 1 else 1e-9
        assert (c * P.bernoulli_poly(2 * a, x) - s).abs().max() < tol


@pytest.mark.parametrize("family", ["lattice", "dnb2"])
def test_fast_solve_equals_dense(family):
    n, d = 256, 3
    if family == "lattice":
        x = P.lattice_points(P.default_lattice_gen_vec(d), [0.1, 0.7, 0.33], 0, n)
        o = OracleFastGP("lattice", x, alpha=2, scale=1.3, lengthscales=[0.5, 1.0, 0.2], noise=1e-6)
    else:
        t = 40
        xb, x = P.dnb2_points(P.default_dnb2_gen_mats(d, t), [123456789, 987654321, 55555], t, 0, n)
        o = OracleFastGP("dnb2", x, xb=xb, t=t, alpha=2, scale=1.3, lengthscales=[0.5, 1.0, 0.2], noise=1e-6)
    y = torch.sin(6 * o.x.sum(1))
    o.add_y(y)
    with torch.no_grad():
        K = o.kernel(o.xb[:, None, :], o.xb[None, :, :]) + o.noise * torch.eye(n)
        ev = torch.linalg.eigvalsh(K)
        lam = o.full_lam().real.sort().values
        assert ((ev - lam).abs() / ev.abs().max()).max() < 1e-12
        assert rel(o.coeffs(), torch.linalg.solve(K, y)) < 1e-8
        assert abs(o.norm_logdet()[1].item() - torch.logdet(K).item()) < 1e-8 * abs(torch.logdet(K).item())
