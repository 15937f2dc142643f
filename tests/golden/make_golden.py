"""Generate the golden fixtures in this directory by running the UNMODIFIED reference (/root/reference/fastgps)
on top of the tests-only qmcpy stand-in (oracle/qmcpy_standin).  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has neither /root/reference nor qmcpy: tests read the committed .npz files, never this script's imports.
Every fixture stores the generator inputs (z / shift, gen matrices / digital shift / t), so the CUDA path and the
oracle port are evaluated on IDENTICAL inputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "qmcpy_standin"))
sys.path.insert(0, "/root/reference")

torch.set_default_dtype(torch.float64)
torch.set_num_threads(4)

import fastgps  # noqa: E402  (the reference, unmodified)
import qmcpy  # noqa: E402  (stand-in)


def f_ackley(x, a=20, b=0.2, c=2 * np.pi, scaling=32.768):
    # the reference doctests' test function (fast_gp_lattice.py:14-22)
    x = 2 * scaling * x - scaling
    t1 = a * torch.exp(-b * torch.sqrt(torch.mean(x ** 2, 1)))
    t2 = torch.exp(torch.mean(torch.cos(c * x), 1))
    return -t1 - t2 + a + np.exp(1)


def f_smooth(x):
    # smooth, periodic-friendly (multitask/fgp_lattice.ipynb cell 4 style)
    j = torch.arange(1, x.shape[1] + 1)
    return torch.cos(2 * np.pi * x).mul(1.0 / j).sum(1) + torch.sin(2 * np.pi * x[:, 0]) * torch.cos(2 * np.pi * x[:, -1])


def run_case(name, family, d, n, m_test, alpha, f, scale=1.0, lengthscales=1.0, noise=None, fit_iterations=30, seed=7):
    if family == "lattice":
        seq = qmcpy.Lattice(dimension=d, seed=seed)
        kw = {} if noise is None else {"noise": noise}
        gp = fastgps.FastGPLattice(seq, alpha=alpha, scale=scale, lengthscales=lengthscales, **kw)
        gen = {"z": seq.gen_vec.astype(np.uint64), "shift": seq.shift}
    else:
        seq = qmcpy.DigitalNetB2(dimension=d, seed=seed)
        kw = {} if noise is None else {"noise": noise}
        gp = fastgps.FastGPDigitalNetB2(seq, alpha=alpha, scale=scale, lengthscales=lengthscales, **kw)
        gen = {"C": seq.gen_mats.astype(np.uint64), "dshift": seq.rshift.astype(np.uint64), "t": np.int64(seq.t)}
    x = gp.get_x_next(n)
    y = f(x)
    gp.add_y_next(y)
    rng = torch.Generator().manual_seed(17)
    xt = torch.rand((m_test, d), generator=rng)
    out = dict(gen)
    out.update(family=family, d=d, n=n, alpha=alpha, x=x.numpy(), y=y.numpy(), xtest=xt.numpy(),
               scale0=gp.scale.detach().numpy(), lengthscales0=gp.lengthscales.detach().numpy(), noise0=gp.noise.detach().numpy())
    if family == "dnb2":
        out["xb"] = gp.get_xb(0).numpy()
    out["k1parts"] = gp.get_k1parts(0, 0).detach().numpy()[:, 0, 0, :]
    # one MLL evaluation with autograd gradients at the initial hyperparameters (abstract_gp.py:239-260,294)
    os.environ["FASTGP_FORCE_RECOMPILE"] = "True"
    cache = gp.get_inv_log_det_cache()
    norm_term, logdet = cache.get_norm_term_logdet_term()
    loss = 0.5 * (norm_term.sum() + logdet.sum() + n * np.log(2 * np.pi))
    loss.backward()
    out.update(norm_term0=norm_term.detach().numpy(), logdet0=logdet.detach().numpy(), loss0=loss.item(),
               grad_raw_scale0=gp.raw_scale.grad.numpy().copy(), grad_raw_lengthscales0=gp.raw_lengthscales.grad.numpy().copy())
    gp.zero_grad()
    del os.environ["FASTGP_FORCE_RECOMPILE"]
    out["lam0"] = gp.get_lam(0, 0).detach().numpy()
    out["ytilde"] = gp.get_ytilde(0).detach().numpy()
    out["coeffs0"] = gp.coeffs.detach().numpy()
    out["pmean0"] = gp.post_mean(xt).numpy()
    out["pvar0"] = gp.post_var(xt).numpy()
    mcov = min(m_test, 32)
    out["pcov0"] = gp.post_cov(xt[:mcov], xt[:mcov // 2]).numpy()
    out["pcmean0"] = gp.post_cubature_mean().numpy()
    out["pcvar0"] = gp.post_cubature_var().numpy()
    out["pvar0_future"] = gp.post_var(xt[:64], n=2 * n).numpy()
    data = gp.fit(iterations=fit_iterations, verbose=0, store_hists=True)
    out.update(fit_iterations=fit_iterations, fit_last_iteration=data["iterations"], loss_hist=data["loss_hist"].numpy(),
               scale_hist=data["scale_hist"].numpy(), lengthscales_hist=data["lengthscales_hist"].numpy(),
               scale1=gp.scale.detach().numpy(), lengthscales1=gp.lengthscales.detach().numpy())
    out["coeffs1"] = gp.coeffs.detach().numpy()
    out["pmean1"] = gp.post_mean(xt).numpy()
    out["pvar1"] = gp.post_var(xt).numpy()
    # incremental doubling (util.py:113-132,173-183): add the next n points and re-read lambda / posterior
    x2 = gp.get_x_next(2 * n)
    gp.add_y_next(f(x2))
    out["lam_2n"] = gp.get_lam(0, 0).detach().numpy()
    out["ytilde_2n"] = gp.get_ytilde(0).detach().numpy()
    out["pmean_2n"] = gp.post_mean(xt[:64]).numpy()
    out["pvar_2n"] = gp.post_var(xt[:64]).numpy()
    # GCV / CV losses (abstract_gp.py:242-251,262-273; util.py:371-394) on fresh objects: value + autograd gradient at the
    # initial hyperparameters and a short fit trajectory.  CV only for the net: the reference's lattice CV loss is complex
    # (get_inv_diag divides by the complex lam) and fails at `loss.item()<...`.
    for metric in (["GCV", "CV"] if family == "dnb2" else ["GCV"]):
        if family == "lattice":
            gp2 = fastgps.FastGPLattice(qmcpy.Lattice(dimension=d, seed=seed), alpha=alpha, scale=scale, lengthscales=lengthscales, **kw)
        else:
            gp2 = fastgps.FastGPDigitalNetB2(qmcpy.DigitalNetB2(dimension=d, seed=seed), alpha=alpha, scale=scale, lengthscales=lengthscales, **kw)
        gp2.add_y_next(f(gp2.get_x_next(n)))
        os.environ["FASTGP_FORCE_RECOMPILE"] = "True"
        cache2 = gp2.get_inv_log_det_cache()
        if metric == "GCV":
            numer, denom = cache2.get_gcv_numer_denom()
            loss2 = (numer / denom).sum()
        else:
            coeffs2 = gp2.coeffs
            del os.environ["FASTGP_FORCE_RECOMPILE"]
            inv_diag = cache2.get_inv_diag()
            os.environ["FASTGP_FORCE_RECOMPILE"] = "True"
            loss2 = ((coeffs2 / inv_diag) ** 2).sum(-1, keepdim=True).sum()
        loss2.backward()
        key = metric.lower()
        out[key + "_loss0"] = loss2.item()
        out[key + "_grad_raw_scale0"] = gp2.raw_scale.grad.numpy().copy()
        out[key + "_grad_raw_lengthscales0"] = gp2.raw_lengthscales.grad.numpy().copy()
        gp2.zero_grad()
        if "FASTGP_FORCE_RECOMPILE" in os.environ:
            del os.environ["FASTGP_FORCE_RECOMPILE"]
        data2 = gp2.fit(loss_metric=metric, iterations=8, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
        out[key + "_loss_hist"] = data2["loss_hist"].numpy()
        out[key + "_lengthscales_hist"] = data2["lengthscales_hist"].numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss0", out["loss0"], "its", data["iterations"], os.path.getsize(path) // 1024, "KiB")


def run_case_multitask(name, family, d, n, T, alpha, m_test=64, noise=1e-6, fit_iterations=8, seed=11, derivatives=None, derivatives_coeffs=None,
                       shared_seq=False, adaptive_nugget=False):
    """num_tasks = T: block eigen-solve util.py:301-323,354-363.  n is one size for every task
    or a list of per-task sizes (ragged arrays are then stored per task: x_0, x_1, ..., y_0, ...)."""
    ns = [int(n)] * T if np.isscalar(n) else [int(v) for v in n]
    ragged = len(set(ns)) > 1
    seeds = np.random.SeedSequence(seed).spawn(T)
    if shared_seq:  # derivative observations at the same points as the function values
        seeds = [seeds[0]] * T
    dkw = {} if derivatives is None else {"derivatives": derivatives, "derivatives_coeffs": derivatives_coeffs}
    if adaptive_nugget:  # util.py:286-290
        dkw["adaptive_nugget"] = True
    if family == "lattice":
        seqs = [qmcpy.Lattice(dimension=d, seed=sd) for sd in seeds]
        gp = fastgps.FastGPLattice(seqs, num_tasks=T, alpha=alpha, noise=noise, **dkw)
        gen = {"z": np.stack([s.gen_vec.astype(np.uint64) for s in seqs]), "shift": np.stack([s.shift for s in seqs])}
    else:
        seqs = [qmcpy.DigitalNetB2(dimension=d, seed=sd) for sd in seeds]
        gp = fastgps.FastGPDigitalNetB2(seqs, num_tasks=T, alpha=alpha, noise=noise, **dkw)
        gen = {"C": np.stack([s.gen_mats.astype(np.uint64) for s in seqs]), "dshift": np.stack([s.rshift.astype(np.uint64) for s in seqs]), "t": np.int64(seqs[0].t)}
    f = lambda x, l: torch.cos(2 * np.pi * x).sum(1) + 0.4 * l * torch.sin(2 * np.pi * x[:, 0]) + 0.1 * l
    if derivatives is not None:  # task l observes sum_p coeff_p D^{beta_p} f0, f0 = prod_j (1 + sin(2 pi x_j) / (j+2))
        def f(x, l):
            out = 0
            for beta, cf in zip(gp.derivatives[l], gp.derivatives_coeffs[l]):
                term = torch.ones(len(x))
                for j in range(d):
                    a = 2 * np.pi * x[:, j]
                    dj = [1 + torch.sin(a) / (j + 2), 2 * np.pi * torch.cos(a) / (j + 2), -(2 * np.pi) ** 2 * torch.sin(a) / (j + 2)][int(beta[j])]
                    term = term * dj
                out = out + cf * term
            return out
    xs = gp.get_x_next(ns)
    ys = [f(xs[l], l) for l in range(T)]
    gp.add_y_next(ys)
    xt = torch.rand((m_test, d), generator=torch.Generator().manual_seed(17))
    out = dict(gen)
    out.update(family=family, d=d, n=max(ns), ns=np.array(ns), T=T, alpha=alpha, noise0=noise, xtest=xt.numpy())
    if adaptive_nugget:
        out["adaptive_nugget"] = np.int64(1)
    if derivatives is not None:
        out.update({"deriv_%d" % l: gp.derivatives[l].numpy() for l in range(T)})
        out.update({"dcoef_%d" % l: gp.derivatives_coeffs[l].numpy() for l in range(T)})
        out["kernel_d01"] = gp.kernel(xt[:8, None, :], xt[None, :5, :], gp.derivatives[0], gp.derivatives[T - 1], gp.derivatives_coeffs[0], gp.derivatives_coeffs[T - 1]).detach().numpy()
        out["kernel_pairs_d10"] = gp.kernel(xt[:8], xt[8:16], gp.derivatives[T - 1], gp.derivatives[0], gp.derivatives_coeffs[T - 1], gp.derivatives_coeffs[0]).detach().numpy()
    if ragged:
        out.update({"x_%d" % l: xs[l].numpy() for l in range(T)})
        out.update({"y_%d" % l: ys[l].numpy() for l in range(T)})
    else:
        out.update(x=np.stack([x.numpy() for x in xs]), y=np.stack([y.numpy() for y in ys]))
    os.environ["FASTGP_FORCE_RECOMPILE"] = "True"
    norm_term, logdet = gp.get_inv_log_det_cache().get_norm_term_logdet_term()
    loss = 0.5 * (norm_term.sum() + logdet.sum() + sum(ns) * np.log(2 * np.pi))
    loss.backward()
    out.update(loss0=loss.item(), norm_term0=norm_term.detach().numpy(), logdet0=logdet.detach().numpy(),
               grad_raw_scale0=gp.raw_scale.grad.numpy().copy(), grad_raw_lengthscales0=gp.raw_lengthscales.grad.numpy().copy(),
               grad_raw_factor0=gp.raw_factor_task_kernel.grad.numpy().copy() if gp.raw_factor_task_kernel.grad is not None else np.zeros(0),
               grad_raw_noise_task0=gp.raw_noise_task_kernel.grad.numpy().copy() if gp.raw_noise_task_kernel.grad is not None else np.zeros(0))
    gp.zero_grad()
    del os.environ["FASTGP_FORCE_RECOMPILE"]
    out["kmat_tasks0"] = gp.gram_matrix_tasks.detach().numpy()
    out["coeffs0"] = gp.coeffs.detach().numpy()
    out["pmean0"] = gp.post_mean(xt).numpy()
    out["pmean0_task1"] = gp.post_mean(xt, task=1).numpy()
    out["pvar0"] = gp.post_var(xt).numpy()
    out["pcov0"] = gp.post_cov(xt[:8], xt[:5]).numpy()
    out["pcmean0"] = gp.post_cubature_mean().numpy()
    out["pcvar0"] = gp.post_cubature_var().numpy()
    out["pccov0"] = gp.post_cubature_cov().numpy()
    if ragged:  # "future" sizes: posterior variance after doubling every task (abstract_gp.py:394,408)
        out["pvar0_n2"] = gp.post_var(xt, n=2 * gp.n).numpy()
        out["pcvar0_n2"] = gp.post_cubature_var(n=2 * gp.n).numpy()
    data = gp.fit(iterations=fit_iterations, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
    out.update(fit_iterations=fit_iterations, fit_last_iteration=data["iterations"], loss_hist=data["loss_hist"].numpy(),
               scale_hist=data["scale_hist"].numpy(), lengthscales_hist=data["lengthscales_hist"].numpy(),
               task_kernel_hist=data["task_kernel_hist"].numpy(), pmean1=gp.post_mean(xt).numpy(), pvar1=gp.post_var(xt).numpy())
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss0", out["loss0"], "its", data["iterations"], os.path.getsize(path) // 1024, "KiB")


def run_deriv_cases():
    """derivative-informed kernels (function values + gradient / mixed derivative sums).
    The noise levels keep the per-frequency blocks well enough conditioned for the reference's own Schur recursion
    (util.py:301-323) to be accurate: at noise 1e-6 the lattice gradient case has blocks of condition number ~6e7 and the
    reference's logdet is 3e-8 (relative) off a dense float64 factorization of its own kernel matrix, its inverse 1e-4."""
    e = torch.eye(2, dtype=int)
    z2 = torch.zeros(2, dtype=int)
    run_case_multitask("dv_lattice_grad_d2_n64_a3", "lattice", 2, 64, 3, 3, derivatives=[z2, e[0], e[1]], shared_seq=True, noise=1e-2)
    run_case_multitask("dv_lattice_mixed_d2_ragged_a3", "lattice", 2, [128, 32], 2, 3, noise=1e-3,
                       derivatives=[z2, torch.tensor([[1, 0], [0, 2], [1, 1]])], derivatives_coeffs=[torch.ones(1), torch.tensor([1., -0.5, 0.25])])
    e3 = torch.eye(3, dtype=int)
    run_case_multitask("dv_dnb2_grad_d3_n64_a3", "dnb2", 3, 64, 4, 3, derivatives=[torch.zeros(3, dtype=int), e3[0], e3[1], e3[2]], shared_seq=True, noise=1e-6)
    run_case_multitask("dv_dnb2_mixed_d2_ragged_a4", "dnb2", 2, [64, 128], 2, 4, noise=1e-6,
                       derivatives=[z2, torch.tensor([[1, 0], [1, 1]])], derivatives_coeffs=[torch.ones(1), torch.tensor([2., -1.])])


def run_case_standard(name, d, ns, kernel_class, T=None, batch=(), noise=1e-4, fit_iterations=8, seed=7, m_test=48):
    """StandardGP (standard_gp.py:11-439, util.py:207-267): the reference's dense GP on explicit points."""
    nt = 1 if T is None else T
    ns = [int(ns)] * nt if np.isscalar(ns) else [int(v) for v in ns]
    seqs = [qmcpy.DigitalNetB2(dimension=d, seed=sd) for sd in np.random.SeedSequence(seed).spawn(nt)]
    kw = dict(shape_batch=torch.Size(batch), shape_lengthscales=torch.Size(batch + (d,))) if batch else {}
    gp = fastgps.StandardGP(seqs if T is not None else seqs[0], num_tasks=T, kernel_class=kernel_class, noise=noise, **kw)
    xs = gp.get_x_next(ns if T is not None else ns[0])
    xs = xs if T is not None else [xs]
    f = lambda x, l: f_ackley(x) * (1 + 0.2 * l) + 0.3 * l * torch.sin(2 * np.pi * x[:, 0])
    ys = [f(xs[l], l) for l in range(nt)]
    if batch:
        ys = [torch.stack([y * (1 + 0.5 * k) + k for k in range(batch[0])]) for y in ys]
    gp.add_y_next(ys if T is not None else ys[0])
    xt = torch.rand((m_test, d), generator=torch.Generator().manual_seed(17))
    out = dict(d=d, ns=np.array(ns), T=nt, solo=T is None, kernel_class=kernel_class, noise0=noise, batch=np.array(batch, dtype=np.int64), xtest=xt.numpy())
    out.update({"x_%d" % l: xs[l].numpy() for l in range(nt)})
    out.update({"y_%d" % l: ys[l].numpy() for l in range(nt)})
    os.environ["FASTGP_FORCE_RECOMPILE"] = "True"
    norm_term, logdet = gp.get_inv_log_det_cache().get_norm_term_logdet_term()
    loss = 0.5 * (norm_term.sum() + logdet.sum() + sum(ns) * np.log(2 * np.pi))
    loss.backward()
    out.update(loss0=loss.item(), norm_term0=norm_term.detach().numpy(), logdet0=logdet.detach().numpy(),
               grad_raw_scale0=gp.raw_scale.grad.numpy().copy(), grad_raw_lengthscales0=gp.raw_lengthscales.grad.numpy().copy())
    gp.zero_grad()
    del os.environ["FASTGP_FORCE_RECOMPILE"]
    out["coeffs0"] = gp.coeffs.detach().numpy()
    out["pmean0"] = gp.post_mean(xt).numpy()
    out["pvar0"] = gp.post_var(xt).numpy()
    out["pcov0"] = gp.post_cov(xt[:8], xt[:5]).numpy()
    if not batch:  # with batch dims the reference's diagonal clamp indexes the batch axis (abstract_gp.py:460-464) and fails
        out["pcov0_eq"] = gp.post_cov(xt[:6], xt[:6]).numpy()
    if kernel_class == "Gaussian":
        out["pcmean0"] = gp.post_cubature_mean().numpy()
        out["pcvar0"] = gp.post_cubature_var().numpy()
        if not batch:  # same indexing failure with batch dims (standard_gp.py:424-428)
            out["pccov0"] = gp.post_cubature_cov().numpy()
    numer, denom = gp.get_inv_log_det_cache().get_gcv_numer_denom()
    out["gcv_loss0"] = (numer / denom).sum().item()
    data = gp.fit(iterations=fit_iterations, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
    out.update(fit_iterations=fit_iterations, fit_last_iteration=data["iterations"], loss_hist=data["loss_hist"].numpy(),
               scale_hist=data["scale_hist"].numpy(), lengthscales_hist=data["lengthscales_hist"].numpy(),
               pmean1=gp.post_mean(xt).numpy(), pvar1=gp.post_var(xt).numpy())
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss0", out["loss0"], "its", data["iterations"], os.path.getsize(path) // 1024, "KiB")


def run_case_multitask_batch(name, family, d, n, T, alpha, batch, hyper_batch=None, m_test=24, noise=1e-4, fit_iterations=6, seed=23, derivatives=None):
    """num_tasks = T with batched outputs y_l (*batch, n_l).  hyper_batch = None: one shared hyperparameter
    set; otherwise the leading shape (a tail of batch) of all five hyperparameters, initialised to different values per set."""
    ns = [int(n)] * T if np.isscalar(n) else [int(v) for v in n]
    seeds = np.random.SeedSequence(seed).spawn(T)
    if derivatives is not None:
        seeds = [seeds[0]] * T
    batch = tuple(batch)
    kw = {"shape_batch": list(batch)}
    if derivatives is not None:
        kw["derivatives"] = derivatives
    if hyper_batch is not None:
        hb = tuple(hyper_batch)
        g = torch.Generator().manual_seed(5)
        kw.update(scale=0.5 + torch.rand(hb + (1,), generator=g), lengthscales=0.3 + torch.rand(hb + (d,), generator=g), noise=noise * (1 + torch.rand(hb + (1,), generator=g)))
        if derivatives is None:
            kw.update(factor_task_kernel=0.5 + torch.rand(hb + (T, 1), generator=g), noise_task_kernel=0.5 + torch.rand(hb + (T,), generator=g))
    else:
        kw["noise"] = noise
    # every object gets its own copies: an identity transform makes the constructor's tensor the parameter itself, which fit() updates in place
    fresh_kw = lambda: {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
    if family == "lattice":
        seqs = [qmcpy.Lattice(dimension=d, seed=sd) for sd in seeds]
        gp = fastgps.FastGPLattice(seqs, num_tasks=T, alpha=alpha, **fresh_kw())
        gen = {"z": np.stack([s.gen_vec.astype(np.uint64) for s in seqs]), "shift": np.stack([s.shift for s in seqs])}
    else:
        seqs = [qmcpy.DigitalNetB2(dimension=d, seed=sd) for sd in seeds]
        gp = fastgps.FastGPDigitalNetB2(seqs, num_tasks=T, alpha=alpha, **fresh_kw())
        gen = {"C": np.stack([s.gen_mats.astype(np.uint64) for s in seqs]), "dshift": np.stack([s.rshift.astype(np.uint64) for s in seqs]), "t": np.int64(seqs[0].t)}
    nb = int(np.prod(batch))

    def f(x, l):  # batch element b: a different amplitude / phase; with derivatives task l observes D^{beta_l} of it
        rows = []
        for b in range(nb):
            if derivatives is None:
                rows.append((1 + 0.5 * b) * torch.cos(2 * np.pi * x).sum(1) + 0.4 * l * torch.sin(2 * np.pi * x[:, 0] + b) + 0.1 * l)
            else:
                term = torch.ones(len(x))
                for j in range(d):
                    a = 2 * np.pi * x[:, j]
                    amp = (1 + 0.25 * b) / (j + 2)
                    term = term * [1 + amp * torch.sin(a), 2 * np.pi * amp * torch.cos(a), -(2 * np.pi) ** 2 * amp * torch.sin(a)][int(gp.derivatives[l][0][j])]
                rows.append(term)
        return torch.stack(rows, 0).reshape(batch + (len(x),))
    xs = gp.get_x_next(ns)
    ys = [f(xs[l], l) for l in range(T)]
    gp.add_y_next(ys)
    xt = torch.rand((m_test, d), generator=torch.Generator().manual_seed(17))
    out = dict(gen)
    out.update(family=family, d=d, ns=np.array(ns), T=T, alpha=alpha, batch=np.array(batch), xtest=xt.numpy(),
               hyper_batch=np.array(hyper_batch if hyper_batch is not None else [], dtype=np.int64))
    for k in ("scale", "lengthscales", "noise", "factor_task_kernel", "noise_task_kernel"):
        out[k + "0"] = getattr(gp, k).detach().numpy().copy()  # copy: an identity transform returns the parameter itself, which fit() updates in place
    if derivatives is not None:
        out.update({"deriv_%d" % l: gp.derivatives[l].numpy() for l in range(T)})
    out.update({"x_%d" % l: xs[l].numpy() for l in range(T)})
    out.update({"y_%d" % l: ys[l].numpy() for l in range(T)})
    os.environ["FASTGP_FORCE_RECOMPILE"] = "True"
    norm_term, logdet = gp.get_inv_log_det_cache().get_norm_term_logdet_term()
    d_out = nb
    loss = 0.5 * (norm_term.sum() + d_out / logdet.numel() * logdet.sum() + d_out * sum(ns) * np.log(2 * np.pi))  # abstract_gp.py:253-256
    loss.backward()
    gr = lambda p: p.grad.numpy().copy() if p.grad is not None else np.zeros(0)
    out.update(loss0=loss.item(), norm_term0=norm_term.detach().numpy(), logdet0=logdet.detach().numpy(),
               grad_raw_scale0=gr(gp.raw_scale), grad_raw_lengthscales0=gr(gp.raw_lengthscales), grad_raw_noise0=gr(gp.raw_noise),
               grad_raw_factor0=gr(gp.raw_factor_task_kernel), grad_raw_noise_task0=gr(gp.raw_noise_task_kernel))
    gp.zero_grad()
    del os.environ["FASTGP_FORCE_RECOMPILE"]
    out["coeffs0"] = gp.coeffs.detach().numpy()
    out["pmean0"] = gp.post_mean(xt).numpy()
    out["pmean0_task1"] = gp.post_mean(xt, task=1).numpy()
    out["pvar0"] = gp.post_var(xt).numpy()
    out["pvar0_task0"] = gp.post_var(xt, task=0).numpy()
    out["pcov0"] = gp.post_cov(xt[:6], xt[:4]).numpy()
    out["pcov0_t10"] = gp.post_cov(xt[:6], xt[:4], task0=1, task1=[0]).numpy()
    out["pcmean0"] = gp.post_cubature_mean().numpy()
    out["pcvar0"] = gp.post_cubature_var().numpy()
    out["pccov0"] = gp.post_cubature_cov().numpy()
    out["pvar0_n2"] = gp.post_var(xt, n=2 * gp.n).numpy()
    data = gp.fit(iterations=fit_iterations, verbose=0, store_hists=True, stop_crit_wait_iterations=100)
    out.update(fit_iterations=fit_iterations, fit_last_iteration=data["iterations"], loss_hist=data["loss_hist"].numpy(),
               scale_hist=data["scale_hist"].numpy(), lengthscales_hist=data["lengthscales_hist"].numpy(), noise_hist=data["noise_hist"].numpy(),
               task_kernel_hist=data["task_kernel_hist"].numpy(), pmean1=gp.post_mean(xt).numpy(), pvar1=gp.post_var(xt).numpy())
    # the other losses of fit (abstract_gp.py:242-273) on fresh objects: GCV, CV, and fits restricted to a subset of the batch by masks
    masks = torch.tensor([[0, nb - 1]]) if len(batch) == 1 else torch.tensor([[0, batch[0] - 1], [batch[1] - 1, 0]])
    out["masks"] = masks.numpy()
    for tag, fkw in (("gcv", {"loss_metric": "GCV"}), ("cv", {"loss_metric": "CV"}), ("mllmask", {"masks": masks}), ("gcvmask", {"loss_metric": "GCV", "masks": masks})):
        cls = fastgps.FastGPLattice if family == "lattice" else fastgps.FastGPDigitalNetB2
        gp2 = cls(seqs, num_tasks=T, alpha=alpha, **fresh_kw())
        gp2.get_x_next(ns)
        gp2.add_y_next(ys)
        try:
            d2 = gp2.fit(iterations=3, verbose=0, store_hists=True, stop_crit_wait_iterations=100, **fkw)
        except Exception as e:
            print("   ", tag, "not available in the reference:", type(e).__name__, str(e)[:80])
            continue
        out.update({"var_%s_loss_hist" % tag: d2["loss_hist"].numpy(), "var_%s_lengthscales_hist" % tag: d2["lengthscales_hist"].numpy(),
                    "var_%s_task_kernel_hist" % tag: d2["task_kernel_hist"].numpy()})
        print("   ", tag, d2["loss_hist"].numpy())
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss0", out["loss0"], "its", data["iterations"], {k: out[k].shape for k in ("pmean0", "pvar0", "pcov0", "pcmean0", "pcvar0", "pccov0", "logdet0")},
          os.path.getsize(path) // 1024, "KiB")


def run_multitask_batch_cases():
    run_case_multitask_batch("mb_lattice_T2_d2_ragged_a2_b3_shared", "lattice", 2, [64, 32], 2, 2, (3,))
    run_case_multitask_batch("mb_dnb2_T2_d3_n64_a2_b2x3_hyper3", "dnb2", 3, 64, 2, 2, (2, 3), hyper_batch=(3,), noise=1e-6)
    run_case_multitask_batch("mb_lattice_T3_d2_ragged_a3_b2_hyper2", "lattice", 2, [32, 128, 64], 3, 3, (2,), hyper_batch=(2,))
    run_case_multitask_batch("mb_lattice_dv_d2_n32_a3_b2_hyper2", "lattice", 2, 32, 2, 3, (2,), hyper_batch=(2,), noise=1e-3,
                             derivatives=[torch.zeros(2, dtype=int), torch.tensor([1, 0])])


def run_nugget_cases():
    run_case_multitask("mt_lattice_T3_d2_ragged_a2_nugget", "lattice", 2, [64, 256, 128], 3, 2, noise=1e-4, adaptive_nugget=True)
    run_case_multitask("mt_dnb2_T2_d3_n128_a2_nugget", "dnb2", 3, 128, 2, 2, noise=1e-5, adaptive_nugget=True)


def run_standard_cases():
    run_case_standard("sg_gaussian_d2_n64", 2, 64, "Gaussian")
    run_case_standard("sg_matern52_d3_T2_n48_20", 3, [48, 20], "Matern52", T=2)
    run_case_standard("sg_gaussian_d2_n32_batch2", 2, 32, "Gaussian", batch=(2,))
    run_case_standard("sg_matern12_d1_n40", 1, 40, "Matern12", noise=1e-3)


if __name__ == "__main__":
    if "--standard-only" in sys.argv:
        run_standard_cases()
        sys.exit(0)
    if "--multitask-batch-only" in sys.argv:
        run_multitask_batch_cases()
        sys.exit(0)
    if "--nugget-only" in sys.argv:
        run_nugget_cases()
        sys.exit(0)
    if "--multitask-only" in sys.argv:
        run_case_multitask("mt_lattice_T2_d2_n256_a2", "lattice", 2, 256, 2, 2)
        run_case_multitask("mt_dnb2_T3_d3_n128_a2", "dnb2", 3, 128, 3, 2)
        run_case_multitask("mt_lattice_T3_d2_n64_a3", "lattice", 2, 64, 3, 3)
    if "--multitask-only" in sys.argv or "--deriv-only" in sys.argv:
        run_deriv_cases()
        if "--deriv-only" in sys.argv:
            sys.exit(0)
    if "--multitask-only" in sys.argv or "--ragged-only" in sys.argv:
        run_case_multitask("mt_lattice_T3_d2_ragged_a2", "lattice", 2, [64, 256, 128], 3, 2)
        run_case_multitask("mt_dnb2_T2_d3_ragged_a2", "dnb2", 3, [256, 32], 2, 2)
        sys.exit(0)
    # C1: FastGPLattice d=2 n=2^10 alpha=2, 2^12 test points (BASELINE.json configs[0])
    run_case("lattice_d2_n1024_a2", "lattice", 2, 2 ** 10, 2 ** 12, 2, f_ackley, fit_iterations=40)
    run_case("lattice_d3_n256_a3", "lattice", 3, 2 ** 8, 2 ** 8, 3, f_smooth, scale=2.5,
             lengthscales=torch.tensor([0.2, 0.7, 1.3]), noise=1e-6, fit_iterations=15)
    run_case("lattice_d8_n4096_a2", "lattice", 8, 2 ** 12, 2 ** 8, 2, f_smooth, scale=1.5, lengthscales=0.4, noise=1e-6, fit_iterations=12)
    run_case("lattice_d2_n64_a1", "lattice", 2, 2 ** 6, 2 ** 6, 1, f_smooth, noise=1e-4, fit_iterations=8)
    run_case("dnb2_d2_n1024_a2", "dnb2", 2, 2 ** 10, 2 ** 12, 2, f_ackley, fit_iterations=40)
    run_case("dnb2_d4_n4096_a2", "dnb2", 4, 2 ** 12, 2 ** 8, 2, f_smooth, scale=2.0, lengthscales=0.6, noise=1e-8, fit_iterations=12)
    run_case("dnb2_d3_n256_a3", "dnb2", 3, 2 ** 8, 2 ** 8, 3, f_smooth, scale=0.7, lengthscales=torch.tensor([0.3, 1.0, 1.9]), noise=1e-8, fit_iterations=10)
    run_case("dnb2_d3_n256_a4", "dnb2", 3, 2 ** 8, 2 ** 8, 4, f_smooth, noise=1e-8, fit_iterations=10)
    run_case("dnb2_d2_n128_a1", "dnb2", 2, 2 ** 7, 2 ** 7, 1, f_smooth, noise=1e-6, fit_iterations=10)
    run_case_multitask("mt_lattice_T2_d2_n256_a2", "lattice", 2, 256, 2, 2)
    run_case_multitask("mt_dnb2_T3_d3_n128_a2", "dnb2", 3, 128, 3, 2)
    run_case_multitask("mt_lattice_T3_d2_n64_a3", "lattice", 2, 64, 3, 3)
    run_case_multitask("mt_lattice_T3_d2_ragged_a2", "lattice", 2, [64, 256, 128], 3, 2)
    run_case_multitask("mt_dnb2_T2_d3_ragged_a2", "dnb2", 3, [256, 32], 2, 2)
    run_deriv_cases()
    run_multitask_batch_cases()
    run_nugget_cases()
    run_standard_cases()
