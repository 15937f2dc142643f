"""GPU parity tests of the drop-in Python API (FastGPLattice / FastGPDigitalNetB2) against the fixtures written by the
UNMODIFIED reference (tests/golden/make_golden.py).  The test bodies follow the reference's own doctests
(fast_gp_lattice.py:24-121, fast_gp_digital_net_b2.py:24-116): get_x_next -> add_y_next -> posterior -> fit -> doubling.
Tolerances: points bit-exact; 1e-10 relative (norm-wise) unless a looser one is stated with its reason."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden

pytestmark = pytest.mark.gpu
dev = "cuda:0"


def rel(a, b):
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def make_gp(g, **kw):
    import fastgaussianprocesses_b200 as fgp
    fam = str(g["family"])
    d, alpha = int(g["d"]), int(g["alpha"])
    args = dict(alpha=alpha, scale=float(g["scale0"][0]), lengthscales=torch.from_numpy(g["lengthscales0"]).clone(), noise=float(g["noise0"][0]), device=dev)
    args.update(kw)
    if fam == "lattice":
        seq = fgp.Lattice(d, generating_vector=g["z"], shift=g["shift"])
        return fgp.FastGPLattice(seq, **args)
    seq = fgp.DigitalNetB2(d, generating_matrices=g["C"], dshift=g["dshift"], t=int(g["t"]))
    return fgp.FastGPDigitalNetB2(seq, **args)


@pytest.mark.parametrize("path", ["fused", "generic"])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_api_matches_reference_fixture(case, path, monkeypatch):
    """path: "fused" = device-side CUDA-graph fit loop; "generic" = Python loop with torch.optim.Rprop (the route taken for
    user optimisers / transforms).  Both must reproduce the reference's trajectory."""
    if path == "generic":
        monkeypatch.setenv("FGP_B200_GENERIC_FIT", "1")
    g = load_golden(case)
    n, d = int(g["n"]), int(g["d"])
    gp = make_gp(g)
    x = gp.get_x_next(n)
    assert x.is_cuda and np.array_equal(x.cpu().numpy(), g["x"])  # bit-exact points
    if "xb" in g:
        assert np.array_equal(gp.get_xb(0, n).cpu().numpy(), g["xb"])
    gp.add_y_next(torch.from_numpy(g["y"]))  # host buffer in, as a reference user would pass it
    assert rel(gp.get_k1parts(0, 0)[:, 0, 0, :], g["k1parts"]) < 1e-13
    assert rel(gp.get_lam(0, 0), g["lam0"]) < 1e-10
    assert rel(gp.get_ytilde(0), g["ytilde"]) < 1e-12
    # coefficients K^-1 y are conditioned like 1/noise: 1e-9 norm-wise
    assert rel(gp.coeffs, g["coeffs0"]) < 1e-9
    xt = torch.from_numpy(g["xtest"])
    pm = gp.post_mean(xt)
    assert pm.shape == (xt.shape[0],)
    ymax = float(np.abs(g["y"]).max())
    # pmean = sum of n terms of size |coeffs| ~ |y|/noise cancelling down to |y|: compare on the scale of y
    # achieved (profiles/PARITY_r02.json, norm-wise relative): pmean <= 3.3e-11, pvar <= 9.3e-11, both on lattice_d2_n1024_a2 whose
    # spectrum spans 11 decades (reference self-spread there: 8.3e-12 / 1.4e-11, profiles/r2_reference_spread.json)
    assert rel(pm, g["pmean0"]) < 1e-10
    pv = gp.post_var(xt)
    sc = float(g["scale0"][0])
    assert rel(pv, g["pvar0"]) < 2e-10
    mcov = g["pcov0"].shape[0]
    pc = gp.post_cov(xt[:mcov], xt[:mcov // 2])
    assert pc.shape == g["pcov0"].shape
    assert rel(pc, g["pcov0"]) < 2e-10  # achieved <= 7.4e-11
    pcs = gp.post_cov(xt[:mcov], xt[:mcov])
    assert (pcs.diagonal() >= 0).all()
    assert torch.allclose(pcs.diagonal(), pv[:mcov], atol=1e-7 * sc)
    assert abs(float(gp.post_cubature_mean()) - float(g["pcmean0"])) < 1e-8 * ymax
    assert abs(float(gp.post_cubature_var()) - float(g["pcvar0"])) < 1e-8 * sc
    pvf = gp.post_var(xt[:64], n=2 * n)
    assert rel(pvf, g["pvar0_future"]) < 2e-9  # achieved 6.0e-10 on lattice_d2_n1024_a2 (spectrum of the doubled point set), <= 4.2e-11 elsewhere
    # fit: same loss trajectory, same stopping iteration, same hyperparameters
    data = gp.fit(iterations=int(g["fit_iterations"]), verbose=0, store_hists=True)
    assert data["iterations"] == int(g["fit_last_iteration"])
    # achieved: loss trajectory <= 7.7e-11, hyperparameter trajectories <= 2e-16 (Rprop only uses the sign of the gradient),
    # posterior after the fit <= 2.5e-11 / 9.7e-11
    assert rel(data["loss_hist"], g["loss_hist"]) < 1e-10
    assert rel(data["scale_hist"], g["scale_hist"]) < 1e-10
    assert rel(data["lengthscales_hist"], g["lengthscales_hist"]) < 1e-10
    assert rel(gp.scale, g["scale1"]) < 1e-10
    assert rel(gp.lengthscales, g["lengthscales1"]) < 1e-10
    assert rel(gp.post_mean(xt), g["pmean1"]) < 1e-10
    assert rel(gp.post_var(xt), g["pvar1"]) < 2e-10


@pytest.mark.parametrize("case", ["lattice_d2_n1024_a2", "dnb2_d2_n1024_a2"])
def test_api_doubling_matches_reference_fixture(case):
    """Incremental doubling (util.py:113-132,173-183): after fit, add the next n points and re-read the spectrum."""
    g = load_golden(case)
    n = int(g["n"])
    gp = make_gp(g, scale=float(g["scale1"][0]), lengthscales=torch.from_numpy(g["lengthscales1"]).clone())
    gp.add_y_next(_f_ackley(gp.get_x_next(n)))
    assert rel(gp.y, g["y"]) < 1e-13
    x2 = gp.get_x_next(2 * n)
    assert x2.shape == (n, int(g["d"]))
    gp.add_y_next(_f_ackley(x2))
    assert rel(gp.get_lam(0, 0), g["lam_2n"]) < 1e-9
    assert rel(gp.get_ytilde(0), g["ytilde_2n"]) < 1e-10
    xt = torch.from_numpy(g["xtest"])[:64]
    ymax = float(np.abs(g["y"]).max())
    assert float((gp.post_mean(xt).cpu() - torch.from_numpy(g["pmean_2n"])).abs().max()) < 1e-6 * ymax
    sc1 = float(g["scale1"][0])
    assert float((gp.post_var(xt).cpu() - torch.from_numpy(g["pvar_2n"])).abs().max()) < 1e-6 * max(float(np.abs(g["pvar_2n"]).max()), sc1 * 1e-6) + 1e-8 * sc1


def _f_ackley(x, a=20, b=0.2, c=2 * np.pi, scaling=32.768):
    x = 2 * scaling * x - scaling
    t1 = a * torch.exp(-b * torch.sqrt(torch.mean(x ** 2, 1)))
    t2 = torch.exp(torch.mean(torch.cos(c * x), 1))
    return -t1 - t2 + a + np.exp(1)


def test_api_guards_and_errors():
    import fastgaussianprocesses_b200 as fgp
    gp = fgp.FastGPLattice(3, seed_for_seq=7, device=dev)
    with pytest.raises(AssertionError):
        gp.get_x_next(100)  # not a power of two (abstract_fast_gp.py:36)
    x = gp.get_x_next(64)
    with pytest.raises(AssertionError):
        gp.add_y_next(torch.zeros(48))  # total samples must be a power of two (abstract_fast_gp.py:40)
    gp = fgp.FastGPLattice(3, seed_for_seq=7, device=dev)
    with pytest.raises(AssertionError):
        gp.fit()  # cannot fit without data
    gp.add_y_next(torch.sin(x.sum(1)))
    with pytest.raises(AssertionError):
        gp.post_var(x[:4], n=96)
    with pytest.raises(AssertionError):
        gp.post_mean(torch.zeros(4, 2))
    # the lattice kernel lives on the unit cube (fast_gp_lattice.py:264-265): x = 1 is admissible and equals x = 0, outside raises
    edge = torch.tensor([[1.0, 0.25, 0.5], [0.0, 0.25, 0.5]])
    pm = gp.post_mean(edge)
    assert abs(float(pm[0] - pm[1])) <= 1e-12 * max(1.0, float(pm.abs().max()))
    pv = gp.post_var(edge)
    assert abs(float(pv[0] - pv[1])) <= 1e-10
    for bad in (torch.tensor([[1.0 + 1e-9, 0.2, 0.2]]), torch.tensor([[-0.1, 0.2, 0.2]])):
        with pytest.raises(AssertionError, match="\\[0,1\\]"):
            gp.post_mean(bad)
        with pytest.raises(AssertionError, match="\\[0,1\\]"):
            gp.post_var(bad)
        with pytest.raises(AssertionError, match="\\[0,1\\]"):
            gp.post_cov(bad, bad)
    with pytest.raises(RuntimeError):
        fgp.FastGPLattice(3, device="cpu")
    assert fgp.FastGPLattice(3, num_tasks=2, shape_batch=[2], device=dev).shape_batch == (2,)  # several tasks with batched outputs: test_multitask_gpu.py
    assert fgp.FastGPLattice(3, adaptive_nugget=True, device=dev).adaptive_nugget  # the identity for one task (util.py:286-290); several: test_multitask_gpu.py
    assert gp.post_mean(torch.rand(5, 3)).shape == (5,)
    assert gp.post_mean(torch.rand(5, 3), task=[0]).shape == (1, 5)
    assert gp.post_mean(torch.zeros(0, 3)).shape == (0,)


@pytest.mark.parametrize("family", ["lattice", "dnb2"])
def test_api_batched_outputs_and_hyperparameters(family):
    """shape_batch with per-batch hyperparameters: every batch element must equal an independent single GP."""
    import fastgaussianprocesses_b200 as fgp
    d, n, Bt = 3, 256, 4
    mk = (lambda **kw: fgp.FastGPLattice(fgp.Lattice(d, seed=11), device=dev, **kw)) if family == "lattice" else \
         (lambda **kw: fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(d, seed=11), device=dev, **kw))
    scale = torch.tensor([[0.5], [1.0], [2.0], [3.0]])
    ls = torch.rand(Bt, d, generator=torch.Generator().manual_seed(3)) + 0.2
    noise = torch.full((Bt, 1), 1e-6)
    gpb = mk(scale=scale, lengthscales=ls, noise=noise, shape_batch=[Bt])
    x = gpb.get_x_next(n)
    freqs = torch.arange(1, Bt + 1, device=x.device, dtype=torch.float64)
    yb = torch.cos(2 * np.pi * x.sum(1)[None, :] * freqs[:, None])
    gpb.add_y_next(yb)
    xt = torch.rand(33, d, generator=torch.Generator().manual_seed(4))
    pmb, pvb = gpb.post_mean(xt), gpb.post_var(xt)
    assert pmb.shape == (Bt, 33) and pvb.shape == (Bt, 33)
    datab = gpb.fit(iterations=6, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
    losses = []
    for b in range(Bt):
        gp1 = mk(scale=float(scale[b, 0]), lengthscales=ls[b].clone(), noise=1e-6)
        gp1.get_x_next(n)
        gp1.add_y_next(yb[b])
        assert rel(gp1.post_mean(xt), pmb[b]) < 1e-9
        assert float((gp1.post_var(xt) - pvb[b]).abs().max()) < 1e-9 * float(scale[b, 0])
        d1 = gp1.fit(iterations=6, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
        losses.append(d1["loss_hist"])
        assert rel(d1["lengthscales_hist"], datab["lengthscales_hist"][:, b]) < 1e-9
        assert rel(d1["scale_hist"], datab["scale_hist"][:, b]) < 1e-9
    # the batched loss is the sum of the independent losses (abstract_gp.py:253-260)
    assert np.allclose(torch.stack(losses).sum(0).numpy(), datab["loss_hist"].numpy(), rtol=1e-9)


def test_api_accepts_qmcpy_style_sequence_objects():
    """The reference passes qmcpy sequence objects; any object with their interface is accepted (points then come from its
    own host generator, abstract_gp.py:307-309) and gives the same GP as this package's GPU-side spec with equal inputs."""
    import os
    import sys
    import fastgaussianprocesses_b200 as fgp
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "qmcpy_standin"))
    import qmcpy  # tests-only stand-in with the qmcpy call signature
    d, n = 3, 512
    for fam in ("lattice", "dnb2"):
        if fam == "lattice":
            host_seq = qmcpy.Lattice(dimension=d, seed=11)
            gp_h = fgp.FastGPLattice(host_seq, device=dev)
            gp_g = fgp.FastGPLattice(fgp.Lattice(d, generating_vector=host_seq.gen_vec, shift=host_seq.shift), device=dev)
        else:
            host_seq = qmcpy.DigitalNetB2(dimension=d, seed=11, t=52)
            gp_h = fgp.FastGPDigitalNetB2(host_seq, device=dev)
            gp_g = fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(d, generating_matrices=host_seq.gen_mats, dshift=host_seq.rshift, t=52), device=dev)
        xh, xg = gp_h.get_x_next(n), gp_g.get_x_next(n)
        assert torch.equal(xh, xg)
        y = torch.cos(2 * np.pi * xh).sum(1)
        gp_h.add_y_next(y)
        gp_g.add_y_next(y)
        dh = gp_h.fit(iterations=8, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
        dg = gp_g.fit(iterations=8, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
        assert np.allclose(dh["loss_hist"].numpy(), dg["loss_hist"].numpy(), rtol=1e-9)  # stored points vs generator mode
        xt = torch.rand(17, d, generator=torch.Generator().manual_seed(2))
        assert rel(gp_h.post_mean(xt), gp_g.post_mean(xt)) < 1e-7


def test_fit_with_user_optimizer_and_transforms_takes_the_generic_path():
    import fastgaussianprocesses_b200 as fgp
    d, n = 2, 256
    sp = (lambda x: torch.log(torch.expm1(x)), lambda x: torch.nn.functional.softplus(x))  # user transforms (not the defaults)
    gp = fgp.FastGPLattice(fgp.Lattice(d, seed=3), device=dev, tfs_lengthscales=sp, noise=1e-6)
    x = gp.get_x_next(n)
    gp.add_y_next(torch.sin(2 * np.pi * x[:, 0]) + 0.1 * torch.cos(2 * np.pi * x[:, 1]))
    opt = torch.optim.Adam(gp.parameters(), lr=0.05)
    data = gp.fit(optimizer=opt, iterations=25, verbose=0, store_loss_hist=True, stop_crit_wait_iterations=100)
    lh = -data["loss_hist"]
    assert torch.isfinite(lh).all() and lh[-1] < lh[0]
    # autograd through the strategy object's seam (util.py:364-370) agrees with a finite difference
    cache = gp.get_inv_log_det_cache()
    norm, logdet = cache.get_norm_term_logdet_term()
    loss = 0.5 * (norm.sum() + logdet.sum())
    gp.zero_grad()
    loss.backward()
    g = gp.raw_lengthscales.grad.clone()
    eps = 1e-6
    with torch.no_grad():
        gp.raw_lengthscales[0] += eps
    n1, l1 = cache.get_norm_term_logdet_term()
    with torch.no_grad():
        gp.raw_lengthscales[0] -= 2 * eps
    n0, l0 = cache.get_norm_term_logdet_term()
    fd = float(0.5 * ((n1.sum() + l1.sum()) - (n0.sum() + l0.sum())) / (2 * eps))
    assert abs(fd - float(g[0])) <= 1e-5 * max(1.0, abs(fd))


def test_verbose_table_is_the_same_on_both_fit_paths(capsys, monkeypatch):
    import fastgaussianprocesses_b200 as fgp
    outs = []
    for generic in (False, True):
        if generic:
            monkeypatch.setenv("FGP_B200_GENERIC_FIT", "1")
        gp = fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(2, seed=5), device=dev, noise=1e-8)
        x = gp.get_x_next(128)
        gp.add_y_next(torch.cos(2 * np.pi * x).sum(1))
        gp.fit(iterations=7, verbose=2, verbose_indent=2, stop_crit_wait_iterations=100)
        outs.append(capsys.readouterr().out)
    assert outs[0] == outs[1] and "iter of 7.0e+00" in outs[0] and outs[0].count("\n") == 2 + 5


def test_repeated_fit_and_growth_reuse_the_device_loop():
    import fastgaussianprocesses_b200 as fgp
    gp = fgp.FastGPLattice(fgp.Lattice(2, seed=9), device=dev, noise=1e-6)
    f = lambda x: torch.cos(2 * np.pi * x).sum(1)
    gp.add_y_next(f(gp.get_x_next(256)))
    d1 = gp.fit(iterations=5, verbose=0, store_loss_hist=True, stop_crit_wait_iterations=100)
    ctx = gp._fit_route[1]
    d2 = gp.fit(iterations=5, verbose=0, store_loss_hist=True, stop_crit_wait_iterations=100)
    assert gp._fit_route[1] == ctx  # same pooled buffers and CUDA graphs
    assert float(d2["loss_hist"][0]) >= float(d1["loss_hist"].max()) - 1e-9  # restarts from the best iterate of the first fit
    # another GP object of the same shape (another randomisation of the lattice) reuses the context too, and is not disturbed by it
    gp_b = fgp.FastGPLattice(fgp.Lattice(2, seed=10), device=dev, noise=1e-6)
    gp_b.add_y_next(f(gp_b.get_x_next(256)) * 2.0)
    ls_a = gp.lengthscales.detach().clone()
    db = gp_b.fit(iterations=5, verbose=0, store_loss_hist=True, stop_crit_wait_iterations=100)
    assert gp_b._fit_route[1] == ctx and db["iterations"] == 5 and torch.equal(gp.lengthscales.detach(), ls_a)
    gp_c = fgp.FastGPLattice(fgp.Lattice(2, seed=10), device=dev, noise=1e-6)
    gp_c.add_y_next(f(gp_c.get_x_next(256)) * 2.0)
    dc = gp_c.fit(iterations=5, verbose=0, store_loss_hist=True, stop_crit_wait_iterations=100)
    assert torch.equal(dc["loss_hist"], db["loss_hist"]) and torch.equal(gp_c.lengthscales.detach(), gp_b.lengthscales.detach())  # no state leaks between fits
    gp.add_y_next(f(gp.get_x_next(512)))
    d3 = gp.fit(iterations=5, verbose=0, store_loss_hist=True, stop_crit_wait_iterations=100)
    assert gp._fit_route[1] != ctx and d3["iterations"] == 5
    xt = torch.rand(9, 2, generator=torch.Generator().manual_seed(1))
    assert torch.isfinite(gp.post_mean(xt)).all() and (gp.post_var(xt) >= 0).all()


def test_add_y_next_arms_the_fit_context_and_hands_it_back():
    """add_y_next acquires the pooled fit context while y is in flight (fast_gp.py:_prearm_fit); fit() takes it if it still fits, a changed
    parameter layout falls back to a fresh loop, and a GP that goes away returns the context to the pool."""
    import gc
    import fastgaussianprocesses_b200 as fgp
    from fastgaussianprocesses_b200.fast_gp import _FitContext
    f = lambda x: torch.cos(2 * np.pi * x).sum(1)
    mk = lambda seed: fgp.FastGPLattice(fgp.Lattice(2, seed=seed), device=dev, noise=1e-6)
    gp = mk(3)
    gp.add_y_next(f(gp.get_x_next(128)))
    pre = gp._prearmed_loop
    assert pre is not None and pre.ctx is not None
    ctx_id = id(pre.ctx)
    d1 = gp.fit(iterations=4, verbose=0, store_loss_hist=True, stop_crit_wait_iterations=100)
    assert gp._fit_route[1] == ctx_id and gp._prearmed_loop is None and d1["iterations"] == 4
    # the same problem without arming: identical trajectory
    os.environ["FGP_B200_NO_PREARM"] = "1"
    try:
        gq = mk(3)
        gq.add_y_next(f(gq.get_x_next(128)))
        assert gq._prearmed_loop is None
        d1q = gq.fit(iterations=4, verbose=0, store_loss_hist=True, stop_crit_wait_iterations=100)
    finally:
        del os.environ["FGP_B200_NO_PREARM"]
    assert torch.equal(d1["loss_hist"], d1q["loss_hist"]) and torch.equal(gp.lengthscales.detach(), gq.lengthscales.detach())
    # armed, then the parameter layout changes: fit() drops the armed loop and still does its job
    gp2 = mk(4)
    gp2.add_y_next(f(gp2.get_x_next(128)))
    gp2.raw_noise.requires_grad_(False)
    noise0 = gp2.noise.detach().clone()
    d2 = gp2.fit(iterations=3, verbose=0, stop_crit_wait_iterations=100)
    assert d2["iterations"] == 3 and gp2._prearmed_loop is None and torch.equal(gp2.noise.detach(), noise0)
    # an armed GP that is never fitted hands its context back when it goes away
    gp3 = mk(5)
    gp3.add_y_next(f(gp3.get_x_next(128)))
    c3 = gp3._prearmed_loop.ctx
    assert all(c3 is not c for c in _FitContext.POOL.get(c3.key, []))
    del gp3, pre
    gc.collect()
    assert any(c3 is c for c in _FitContext.POOL.get(c3.key, []))


def test_sharded_posterior_single_process_is_identity():
    import fastgaussianprocesses_b200 as fgp
    from fastgaussianprocesses_b200.distributed import post_mean_sharded, post_var_sharded
    gp = fgp.FastGPLattice(fgp.Lattice(3, seed=4), device=dev, noise=1e-6)
    x = gp.get_x_next(128)
    gp.add_y_next(torch.cos(2 * np.pi * x).sum(1))
    xt = torch.rand(21, 3, generator=torch.Generator().manual_seed(8))
    assert torch.equal(post_mean_sharded(gp, xt), gp.post_mean(xt))
    assert torch.equal(post_var_sharded(gp, xt), gp.post_var(xt))


@pytest.mark.parametrize("family", ["lattice", "dnb2"])
def test_shared_hyperparameters_over_a_batch_of_outputs_vs_oracle(family):
    """shape_batch=[3] with ONE hyperparameter set: the loss sums the norm terms and weighs the log-determinant by the
    number of outputs (abstract_gp.py:253-260); trajectory and posterior against the CPU oracle."""
    import fastgaussianprocesses_b200 as fgp
    from oracle import primitives as P
    from oracle.fgp_oracle import OracleFastGP
    d, n = 3, 512
    if family == "lattice":
        seq = fgp.Lattice(d, seed=21)
        gp = fgp.FastGPLattice(seq, device=dev, shape_batch=[3], noise=1e-6)
        o = OracleFastGP("lattice", P.lattice_points(seq.gen_vec, seq.shift, 0, n), alpha=2, noise=1e-6)
    else:
        seq = fgp.DigitalNetB2(d, seed=21)
        gp = fgp.FastGPDigitalNetB2(seq, device=dev, shape_batch=[3], noise=1e-6)
        xb, xh = P.dnb2_points(seq.gen_mats, seq.rshift, seq.t, 0, n)
        o = OracleFastGP("dnb2", xh, xb=xb, t=seq.t, alpha=2, noise=1e-6)
    x = gp.get_x_next(n)
    fr = torch.tensor([1.0, 2.0, 3.0], device=x.device)[:, None]
    y = torch.cos(2 * np.pi * x.sum(1)[None, :] * fr) + 0.2 * fr
    gp.add_y_next(y)
    o.add_y(y.cpu())
    ro = o.fit(iterations=6, stop_crit_wait_iterations=100)
    rg = gp.fit(iterations=6, verbose=0, store_loss_hist=True, stop_crit_wait_iterations=100)
    assert rg["iterations"] == ro["iterations"] == 6
    assert np.allclose(-rg["loss_hist"].numpy(), ro["loss_hist"], rtol=1e-9)
    assert rel(gp.lengthscales, o.lengthscales) < 1e-8
    xt = torch.rand(11, d, generator=torch.Generator().manual_seed(6))
    pm = gp.post_mean(xt)
    assert pm.shape == (3, 11)
    ref = torch.stack([o.post_mean(xt, coeffs=o.solve(y[i].cpu())) for i in range(3)])
    assert float((pm.cpu() - ref).abs().max()) < 1e-7 * float(y.abs().max())


def test_fit_zero_iterations_evaluates_the_loss_once():
    import fastgaussianprocesses_b200 as fgp
    gp = fgp.FastGPLattice(fgp.Lattice(2, seed=1), device=dev, noise=1e-6)
    x = gp.get_x_next(64)
    gp.add_y_next(torch.cos(2 * np.pi * x).sum(1))
    s0 = gp.scale.detach().clone()
    data = gp.fit(iterations=0, verbose=0, store_loss_hist=True)
    assert data["iterations"] == 0 and data["loss_hist"].shape == (1,) and torch.equal(gp.scale.detach(), s0)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_gcv_and_cv_losses_match_reference_fixture(case):
    """GCV (both families) and CV (net) losses through the autograd route: value, gradient w.r.t. the raw parameters and a
    short fit trajectory against the unmodified reference (tests/golden/make_golden.py)."""
    g = load_golden(case)
    n = int(g["n"])
    for metric in (["GCV", "CV"] if "cv_loss0" in g else ["GCV"]):
        key = metric.lower()
        gp = make_gp(g)
        gp.add_y_next(torch.from_numpy(g["y"]))
        gp.get_x_next(n)
        cache = gp.get_inv_log_det_cache()
        if metric == "GCV":
            numer, denom = cache.get_gcv_numer_denom()
            loss = (numer / denom).sum()
        else:
            loss = ((gp.coeffs / cache.get_inv_diag()) ** 2).sum(-1, keepdim=True).sum()
        if metric == "CV":  # coeffs is a no-grad cache; the differentiable value comes from the fit route
            loss, _, _, _ = gp._autograd_loss("CV", None, 1, 1, 0.0)
        assert abs(float(loss) - float(g[key + "_loss0"])) <= 1e-8 * abs(float(g[key + "_loss0"]))
        loss.backward()
        # GCV is invariant to the scale up to the nugget: that derivative is a cancellation residue ~1e-9 of the loss
        gs_ref = float(g[key + "_grad_raw_scale0"][0])
        assert abs(float(gp.raw_scale.grad[0]) - gs_ref) <= 1e-6 * abs(gs_ref) + 1e-10 * abs(float(g[key + "_loss0"]))
        assert rel(gp.raw_lengthscales.grad, g[key + "_grad_raw_lengthscales0"]) < 1e-6
        gp.zero_grad()
        data = gp.fit(loss_metric=metric, iterations=8, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
        assert np.allclose(data["loss_hist"].numpy(), g[key + "_loss_hist"], rtol=1e-6)
        assert rel(data["lengthscales_hist"], g[key + "_lengthscales_hist"]) < 1e-6


def test_masked_mll_fit_equals_fit_on_the_selected_outputs():
    """fit(masks=...) optimises only y[..., *masks] (abstract_gp.py:226-231,257-259): with one shared hyperparameter set it
    must follow the trajectory of a GP that only holds those outputs."""
    import fastgaussianprocesses_b200 as fgp
    d, n = 2, 256
    mk = lambda **kw: fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(d, seed=13), device=dev, noise=1e-6, **kw)
    gpa = mk(shape_batch=[3])
    x = gpa.get_x_next(n)
    fr = torch.tensor([1.0, 2.0, 3.0], device=x.device)[:, None]
    y = torch.cos(2 * np.pi * x.sum(1)[None, :] * fr)
    gpa.add_y_next(y)
    da = gpa.fit(masks=torch.tensor([0, 2]), iterations=6, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
    gpb = mk(shape_batch=[2])
    gpb.get_x_next(n)
    gpb.add_y_next(y[[0, 2]])
    db = gpb.fit(iterations=6, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
    assert np.allclose(da["loss_hist"].numpy(), db["loss_hist"].numpy(), rtol=1e-8)
    assert rel(da["lengthscales_hist"], db["lengthscales_hist"]) < 1e-7


@pytest.mark.parametrize("family,d,m,batch,iters", [("lattice", 8, 20, (), 12), ("lattice", 4, 14, (), 25), ("lattice", 2, 13, (3,), 20), ("lattice", 8, 18, (5,), 6),
                                                    ("dnb2", 4, 16, (), 20), ("dnb2", 3, 14, (2,), 15), ("lattice", 3, 16, (), 40)])
def test_persistent_kernel_fit_matches_the_three_launch_route(monkeypatch, family, d, m, batch, iters):
    """fgp_fit_iterations (ONE cooperative launch per chunk of iterations: grid barriers between the passes, the fit step in the
    tail, early stop on the device) walks the same tiles with the same arithmetic as the per-pass kernels; only the number of
    threads that share a tile's partial sums can differ (the persistent kernel runs every pass with one block size), so the
    whole fit trajectory -- losses, hyperparameter histories, stopping iteration, final parameters -- agrees to round-off."""
    import fastgaussianprocesses_b200 as fgp
    n = 1 << m

    def run(coop):
        monkeypatch.setenv("FGP_COOP", "1" if coop else "0")
        if family == "lattice":
            gp = fgp.FastGPLattice(fgp.Lattice(d, seed=11), device=dev, noise=1e-6, shape_batch=torch.Size(batch),
                                   shape_lengthscales=torch.Size(batch + (d,)), shape_scale=torch.Size(batch + (1,)))
        else:
            gp = fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(d, seed=11), device=dev, noise=1e-6, shape_batch=torch.Size(batch),
                                        shape_lengthscales=torch.Size(batch + (d,)), shape_scale=torch.Size(batch + (1,)))
        x = gp.get_x_next(n)
        y = torch.cos(2 * np.pi * x).sum(1) + torch.sin(2 * np.pi * x[:, 0]) * x[:, -1]
        if batch:
            y = torch.stack([y * (1.0 + 0.3 * k) + 0.1 * k * torch.sin(4 * np.pi * x[:, 0]) for k in range(batch[0])])
        gp.add_y_next(y)
        data = gp.fit(iterations=iters, verbose=0, store_hists=True, stop_crit_wait_iterations=7)
        assert gp._fit_route[0] == ("coop" if coop else "graph")  # FGP_COOP=1 opts into the persistent kernel (default: three launches per iteration)
        return data, gp.raw_scale.detach().clone(), gp.raw_lengthscales.detach().clone()

    d1, s1, l1 = run(True)
    d0, s0, l0 = run(False)
    assert d1["iterations"] == d0["iterations"]
    for key in ("loss_hist", "scale_hist", "lengthscales_hist"):
        assert rel(d1[key], d0[key]) < 1e-12, key
    assert rel(s1, s0) < 1e-12 and rel(l1, l0) < 1e-12


@pytest.mark.parametrize("family,d,m", [("lattice", 5, 17), ("dnb2", 5, 15)])
def test_programmatic_dependent_launch_route_is_bit_identical(monkeypatch, family, d, m):
    """FGP_PDL=2 launches the three per-pass kernels with the programmatic-stream-serialization attribute (the next kernel's CTAs become
    resident as the current ones exit, pass C triggers before its serial fit tail) -- same kernels, same arithmetic: the fit trajectory must
    be bit-identical, launched eagerly and replayed from a CUDA graph captured with the attribute."""
    import fastgaussianprocesses_b200 as fgp
    n = 1 << m

    def run(pdl, graph):
        monkeypatch.setenv("FGP_PDL", str(pdl))
        if graph:
            monkeypatch.delenv("FGP_B200_NO_GRAPH", raising=False)
        else:
            monkeypatch.setenv("FGP_B200_NO_GRAPH", "1")
        gp = (fgp.FastGPLattice(fgp.Lattice(d, seed=23), device=dev, noise=1e-6) if family == "lattice"
              else fgp.FastGPDigitalNetB2(fgp.DigitalNetB2(d, seed=23), device=dev, noise=1e-6))
        x = gp.get_x_next(n)
        gp.add_y_next(torch.cos(2 * np.pi * x).sum(1) + torch.sin(2 * np.pi * x[:, 0]) * x[:, -1])
        data = gp.fit(iterations=18, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
        return data, gp.raw_lengthscales.detach().clone()

    d2g, l2g = run(2, True)   # first use of this shape: the pooled context captures its graphs with the attribute
    d2, l2 = run(2, False)
    d0, l0 = run(0, False)
    for a, la in ((d2g, l2g), (d2, l2)):
        assert a["iterations"] == d0["iterations"]
        for key in ("loss_hist", "scale_hist", "lengthscales_hist"):
            assert torch.equal(torch.as_tensor(a[key]), torch.as_tensor(d0[key])), key
        assert torch.equal(la, l0)
