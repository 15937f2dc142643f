"""StandardGP: dense-Cholesky Gaussian process regression with the reference's API (fastgps/standard_gp.py:11-439,
strategy object util.py:207-267).

north_star keeps `StandardGP` in the drop-in surface but it is NOT on the structured-covariance hot path (SURVEY.md
section 8 lists no row for it): there is no fast transform to exploit, the O(n^3) factorization is a library call in the
reference (torch.linalg.cholesky) and stays one here -- torch on whatever device the user names (CUDA on the GPU box;
the CPU works too, which is what the `-m "not gpu"` parity test uses).  What this file adds is the host-side mirror:
same constructor arguments, kernels (Gaussian, Matern 1/2, 3/2, 5/2; derivative observations through autograd as in
standard_gp.py:293-344), adaptive nugget and jitter doubling (util.py:218-243), posterior mean / variance / covariance,
Gaussian-kernel cubature (standard_gp.py:345-439), losses MLL / GCV / CV and the shared fit() loop.

Shapes follow the reference: hyperparameter batch dims lead, `num_tasks=None` drops the task axis of every output.
"""
import numpy as np
import torch

from . import sequences
from .fast_gp import AbstractFastGP, DEFAULT_TFS_ID, DEFAULT_TFS_LOG_EXP, _prod


class _DataSequence(object):
    """A fixed design handed in through `data={"x": ..., "y": ...}` (role of util.py:6-15 DummyDiscreteDistrib)."""
    order, replications, randomize = "GIVEN", 1, "FALSE"

    def __init__(self, x):
        assert isinstance(x, torch.Tensor) and x.ndim == 2
        self.x = x.detach().clone()
        self.n, self.d = (int(v) for v in x.shape)

    def generate(self, n_min, n_max, device):
        assert n_min == 0 and n_max == self.n, "trying to generate samples other than the one provided is invalid"
        x = self.x.to(device)
        return x, x


class _HostPoints(object):
    """Adapter for a user-supplied qmcpy-style generator: any point set is admissible for a dense GP."""

    def __init__(self, seq):
        self.seq = seq
        self.d = int(seq.d)
        rep = getattr(seq, "replications", 1)
        self.replications = 1 if rep is None else int(rep)
        self.order = str(getattr(seq, "order", "GIVEN")).upper()
        self.randomize = str(getattr(seq, "randomize", "FALSE")).upper()

    def generate(self, n_min, n_max, device):
        x = torch.from_numpy(np.ascontiguousarray(self.seq(n_min=int(n_min), n_max=int(n_max)), dtype=np.float64)).to(device)
        return x, x


class _GpuNetPoints(object):
    """Default design: this package's base-2 digital net (the reference's default is qmcpy.DigitalNetB2 in Gray-code
    order, standard_gp.py:232; the order of a space-filling design is immaterial to a dense GP).  Points come from the K1
    kernel on a CUDA device and from the same integer recurrence in numpy elsewhere."""
    order, replications = "NATURAL", 1

    def __init__(self, d, seed):
        self.spec = sequences.DigitalNetB2(d, seed=seed, randomize="DS")
        self.d, self.randomize = d, "DS"

    def generate(self, n_min, n_max, device):
        if torch.device(device).type == "cuda":
            x, _ = self.spec.generate(n_min, n_max, device)
            return x, x
        s = self.spec
        i = np.arange(int(n_min), int(n_max), dtype=np.uint64)
        xb = np.tile(s.rshift[None, :], (len(i), 1))
        for k in range(s.gen_mats.shape[1]):
            bit = ((i >> np.uint64(k)) & np.uint64(1)).astype(bool)
            xb[bit] ^= s.gen_mats[None, :, k]
        x = torch.from_numpy(xb.astype(np.float64) * 2.0 ** (-s.t)).to(device)
        return x, x


class _DenseInverseLogDetCache(object):
    """Strategy object of `get_inv_log_det_cache` for the dense GP (util.py:207-267): `__call__() -> (K^-1, logdet)`,
    `gram_matrix_solve`, `get_norm_term_logdet_term`, `get_gcv_numer_denom`, `get_inv_diag`."""

    def __init__(self, gp, n):
        self.fgp = gp
        self.n = n
        self.nvec = [int(v) for v in n.tolist()]
        self._key = None

    def _gram(self, jitter):
        gp, T, ns = self.fgp, self.fgp.num_tasks, self.nvec
        kt = gp.gram_matrix_tasks
        noise = gp.noise[..., 0]
        blocks = [[None] * T for _ in range(T)]
        tr0 = None
        for l0 in range(T):
            for l1 in range(l0 + 1):
                k = gp._kernel(gp.get_x(l0, ns[l0])[:, None, :], gp.get_x(l1, ns[l1])[None, :, :], gp.derivatives[l0], gp.derivatives[l1],
                               gp.derivatives_coeffs[l0], gp.derivatives_coeffs[l1])
                if l0 == l1:
                    nl = noise
                    if gp.adaptive_nugget:  # util.py:218-229: the nugget of task l scales with trace(K_ll) / trace(K_00)
                        tr = torch.diagonal(k, dim1=-2, dim2=-1).sum(-1)
                        tr0 = tr if l0 == 0 else tr0
                        nl = noise * tr / tr0
                    k = k + jitter * nl[..., None, None] * torch.eye(ns[l0], device=gp.device)
                blocks[l0][l1] = k
        rows = [torch.cat([kt[..., l0, l1, None, None] * (blocks[l0][l1] if l1 <= l0 else blocks[l1][l0].transpose(-2, -1)) for l1 in range(T)], -1) for l0 in range(T)]
        bshape = torch.broadcast_shapes(*[r.shape[:-2] for r in rows])
        return torch.cat([r.expand(tuple(bshape) + tuple(r.shape[-2:])) for r in rows], -2)

    def __call__(self):
        gp = self.fgp
        key = gp._param_key()
        if self._key != key or torch.is_grad_enabled():
            jitter = 1.0
            while True:  # util.py:230-243: double the nugget until the factorization goes through
                L, info = torch.linalg.cholesky_ex(self._gram(jitter))
                if not bool((info != 0).any()):
                    break
                jitter *= 2.0
            self.logdet = 2 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
            self.thetainv = torch.cholesky_inverse(L)
            self._key = None if torch.is_grad_enabled() and self.thetainv.requires_grad else key
        return self.thetainv, self.logdet

    def gram_matrix_solve(self, y):
        assert y.size(-1) == sum(self.nvec)
        thetainv, _ = self()
        return torch.einsum("...ij,...j->...i", thetainv, y)

    def get_norm_term_logdet_term(self):
        y = torch.cat(self.fgp._y, -1)
        thetainv, logdet = self()
        v = torch.einsum("...ij,...j->...i", thetainv, y)
        return (y * v).sum(-1, keepdim=True), logdet[..., None]

    def get_gcv_numer_denom(self):
        y = torch.cat(self.fgp._y, -1)
        thetainv, _ = self()
        v = torch.einsum("...ij,...j->...i", thetainv, y)
        tr = torch.diagonal(thetainv, dim1=-2, dim2=-1).sum(-1)[..., None]
        return (v ** 2).sum(-1, keepdim=True), (tr / thetainv.size(-1)) ** 2

    def get_inv_diag(self):
        return torch.diagonal(self()[0], dim1=-2, dim2=-1)


class _DenseEngine(object):
    """What AbstractFastGP.fit() asks of a non-fused model: the three losses of abstract_gp.py:242-273 on the dense factor."""

    def __init__(self, gp):
        self.gp = gp

    def loss(self, loss_metric, d_out, mll_const, masks=None, cv_weights=1):
        gp = self.gp
        cache = gp.get_inv_log_det_cache()
        sb = list(gp.shape_batch)
        if loss_metric == "MLL":
            norm, logdet = cache.get_norm_term_logdet_term()
            if masks is None:
                term1, term2 = norm.sum(), d_out / _prod(logdet.shape) * logdet.sum()
            else:
                term1, term2 = norm[..., *masks, 0].sum(), logdet.expand(sb + [1])[..., *masks, 0].sum()
            loss = 1 / 2 * (term1 + term2 + mll_const)
            return loss, term1, term2, -loss
        if loss_metric == "GCV":
            numer, denom = cache.get_gcv_numer_denom()
            term1, term2 = (numer, denom) if masks is None else (numer[..., *masks, :], denom.expand(sb + [1])[..., *masks, :])
            loss = (term1 / term2).sum()
            return loss, term1, term2, loss
        coeffs = cache.gram_matrix_solve(torch.cat(gp._y, -1))
        sq = ((coeffs / cache.get_inv_diag()) ** 2 * cv_weights).sum(-1, keepdim=True)
        loss = sq.sum() if masks is None else sq[..., *masks, 0].sum()
        nan = torch.nan * torch.ones(1)
        return loss, nan, nan, loss


class StandardGP(AbstractFastGP):
    """Dense Gaussian process regression; drop-in for fastgps.StandardGP (standard_gp.py:11-439).  Differences: `seqs` may be
    an int (dimension; the design is then this package's digital net), a qmcpy-style generator object, or a list of them;
    `compile_dist_func*` are accepted and ignored."""
    _FAMILY = 2
    _DENSE = True
    _XBDTYPE = torch.float64
    _FTOUTDTYPE = torch.float64
    _t = 0
    KERNEL_CLASSES = ["gaussian", "matern12", "matern32", "matern52"]

    def __init__(self, seqs, num_tasks=None, seed_for_seq=None, scale=1., lengthscales=1., noise=1e-4, factor_task_kernel=1.,
                 rank_factor_task_kernel=None, noise_task_kernel=1., device="cuda", tfs_scale=DEFAULT_TFS_LOG_EXP,
                 tfs_lengthscales=DEFAULT_TFS_LOG_EXP, tfs_noise=DEFAULT_TFS_LOG_EXP, tfs_factor_task_kernel=DEFAULT_TFS_ID,
                 tfs_noise_task_kernel=DEFAULT_TFS_LOG_EXP, requires_grad_scale=True, requires_grad_lengthscales=True,
                 requires_grad_noise=False, requires_grad_factor_task_kernel=None, requires_grad_noise_task_kernel=None,
                 shape_batch=torch.Size([]), shape_scale=torch.Size([1]), shape_lengthscales=None, shape_noise=torch.Size([1]),
                 shape_factor_task_kernel=None, shape_noise_task_kernel=None, derivatives=None, derivatives_coeffs=None,
                 kernel_class="Gaussian", adaptive_nugget=True, data=None, compile_dist_func=False, compile_dist_func_kwargs={}):
        ntasks = 1 if num_tasks is None else num_tasks
        if data is not None:
            assert isinstance(seqs, int), "passing in data requires seqs (the first argument) is a int specifying the dimension"
            assert isinstance(data, dict) and "x" in data and "y" in data, "data must be a dict with keys 'x' and 'y'"
            xs = [data["x"]] if isinstance(data["x"], torch.Tensor) else data["x"]
            ys = [data["y"]] if isinstance(data["y"], torch.Tensor) else data["y"]
            assert isinstance(xs, list) and len(xs) == ntasks and all(isinstance(x, torch.Tensor) and x.ndim == 2 and x.size(1) == seqs for x in xs), \
                "data['x'] should be a list of 2d tensors of length num_tasks with each number of columns equal to the dimension"
            assert isinstance(ys, list) and len(ys) == ntasks and all(isinstance(y, torch.Tensor) and y.ndim >= 1 for y in ys), \
                "data['y'] should be a list of tensors of length num_tasks"
            seqs = [_DataSequence(x) for x in xs]
        kernel_class = str(kernel_class).lower()
        assert kernel_class in self.KERNEL_CLASSES, "kernel_class must in %s" % str(self.KERNEL_CLASSES)
        self.kernel_class_pending = kernel_class
        assert isinstance(compile_dist_func, bool)
        super().__init__(seqs, num_tasks, seed_for_seq, 2, scale, lengthscales, noise, factor_task_kernel, rank_factor_task_kernel,
                         noise_task_kernel, device, tfs_scale, tfs_lengthscales, tfs_noise, tfs_factor_task_kernel, tfs_noise_task_kernel,
                         requires_grad_scale, requires_grad_lengthscales, requires_grad_noise, requires_grad_factor_task_kernel,
                         requires_grad_noise_task_kernel, shape_batch, shape_scale, shape_lengthscales, shape_noise,
                         shape_factor_task_kernel, shape_noise_task_kernel, derivatives, derivatives_coeffs, False, {}, adaptive_nugget)
        self.kernel_class = kernel_class
        self.available_kernel_classes = list(self.KERNEL_CLASSES)
        self._mt = _DenseEngine(self)
        if data is not None:
            self.add_y_next([y.to(self.device) for y in ys], task=torch.arange(self.num_tasks))

    # ------------------------------------------------------------------------------------------------ hooks of the shared ctor
    @staticmethod
    def _default_sequence(d, seed):
        return _GpuNetPoints(d, seed)

    @staticmethod
    def _adopt_sequence(seq):
        if isinstance(seq, (_DataSequence, _GpuNetPoints, _HostPoints)):
            return seq
        if isinstance(seq, (sequences.Lattice, sequences.DigitalNetB2)):
            class _Spec(_HostPoints):
                def generate(self, n_min, n_max, device):
                    x, _ = self.seq.generate(n_min, n_max, device)
                    return x, x
            return _Spec(seq)
        return _HostPoints(seq)

    def get_x_next(self, n, task=None):
        """abstract_gp.py:310-330 (no power-of-two requirement for a dense GP)."""
        if isinstance(n, (int, np.integer)):
            n = [int(n)]
        n = [int(v) for v in (n.tolist() if isinstance(n, torch.Tensor) else n)]
        inttask, task = self._parse_task(task)
        assert len(n) == len(task)
        cur = self.n.tolist()
        assert all(n[i] >= cur[int(l)] for i, l in enumerate(task)), "maximum sequence index must be greater than the current number of samples"
        out = [self.xxb_seqs[int(l)][cur[int(l)]:n[i]][0] for i, l in enumerate(task)]
        return out[0] if inttask else out

    def add_y_next(self, y_next, task=None):
        if isinstance(y_next, torch.Tensor):
            y_next = [y_next]
        _, task = self._parse_task(task)
        assert isinstance(y_next, list) and len(y_next) == len(task)
        assert all(y.shape[:-1] == self.shape_batch for y in y_next)
        for i, l in enumerate(task):
            self._y[int(l)] = torch.cat([self._y[int(l)], y_next[i].to(self.device)], -1)
        self.n = torch.tensor([y.size(-1) for y in self._y], dtype=int, device=self.device)
        self._nint = int(self.n.max())
        self.m = torch.where(self.n == 0, -1, torch.log2(self.n.clamp(min=1))).to(int)
        cur = self.n.tolist()
        for key in list(self.inv_log_det_cache_dict.keys()):
            if any(k < c for k, c in zip(key, cur)):
                del self.inv_log_det_cache_dict[key]
        self._epoch += 1

    def get_inv_log_det_cache(self, n=None):
        n = self._n_tensor(n)
        assert n.shape == (self.num_tasks,) and (n >= self.n).all()
        key = tuple(n.tolist())
        if key not in self.inv_log_det_cache_dict:
            self.inv_log_det_cache_dict[key] = _DenseInverseLogDetCache(self, n)
        return self.inv_log_det_cache_dict[key]

    def _n_tensor(self, n):
        if n is None:
            return self.n
        if isinstance(n, (int, np.integer)):
            return torch.tensor([int(n)] * self.num_tasks, dtype=int, device=self.device)
        return torch.as_tensor(n, dtype=int, device=self.device).reshape(-1)

    @property
    def coeffs(self):
        return self.get_inv_log_det_cache().gram_matrix_solve(torch.cat(self._y, -1))

    # ------------------------------------------------------------------------------------------------ kernel
    def kernel(self, x, z, beta0=None, beta1=None, c0=None, c1=None):
        zero = torch.zeros((1, self.d), dtype=int, device=self.device)
        beta0 = zero if beta0 is None else torch.atleast_2d(beta0)
        beta1 = zero if beta1 is None else torch.atleast_2d(beta1)
        c0 = torch.ones(len(beta0), device=self.device) if c0 is None else c0
        c1 = torch.ones(len(beta1), device=self.device) if c1 is None else c1
        return self._kernel(x.to(self.device), z.to(self.device), beta0, beta1, c0, c1)

    def _base_kernel(self, x, z):
        """scale * k0(x, z) with the hyperparameter batch dims in front (standard_gp.py:316-331)."""
        extra = max(x.ndim, z.ndim) - 1
        ls = self.lengthscales
        ls = ls.reshape(tuple(ls.shape[:-1]) + (1,) * extra + (ls.size(-1),))
        scale = self.scale
        scale = scale.reshape(tuple(scale.shape[:-1]) + (1,) * extra)
        q = ((x - z) ** 2 / (2 * ls)).sum(-1)
        if self.kernel_class == "gaussian":
            return scale * torch.exp(-q)
        r = torch.sqrt(q)
        if self.kernel_class == "matern12":
            return scale * torch.exp(-r)
        if self.kernel_class == "matern32":
            return scale * (1 + np.sqrt(3) * r) * torch.exp(-np.sqrt(3) * r)
        return scale * (1 + np.sqrt(5) * r + 5 * r ** 2 / 3) * torch.exp(-np.sqrt(5) * r)

    def _kernel(self, x, z, beta0, beta1, c0, c1):
        """sum_{t0,t1} c0[t0] c1[t1] d^beta0[t0]_x d^beta1[t1]_z k(x, z); derivatives by autograd, one coordinate leaf per
        dimension as in standard_gp.py:293-344."""
        assert c0.ndim == 1 and c1.ndim == 1 and beta0.shape == (len(c0), self.d) and beta1.shape == (len(c1), self.d)
        assert x.size(-1) == self.d and z.size(-1) == self.d
        if not ((beta0 > 0).any() or (beta1 > 0).any()):
            return float(c0.sum() * c1.sum()) * self._base_kernel(x, z) if (len(c0) * len(c1) > 1 or c0[0] != 1 or c1[0] != 1) else self._base_kernel(x, z)
        with torch.enable_grad():
            lead = torch.broadcast_shapes(x.shape[:-1], z.shape[:-1])
            xs = [x[..., j].expand(lead).clone().requires_grad_(True) for j in range(self.d)]
            zs = [z[..., j].expand(lead).clone().requires_grad_(True) for j in range(self.d)]
            base = self._base_kernel(torch.stack(xs, -1), torch.stack(zs, -1))
            out = 0
            for i0 in range(len(c0)):
                for i1 in range(len(c1)):
                    part = base
                    for leaves, beta in ((xs, beta0[i0]), (zs, beta1[i1])):
                        for j in range(self.d):
                            for _ in range(int(beta[j])):
                                part = torch.autograd.grad(part, leaves[j], grad_outputs=torch.ones_like(part), create_graph=True)[0]
                    out = out + c0[i0] * c1[i1] * part
        return out

    # ------------------------------------------------------------------------------------------------ posterior
    def _cross(self, x, task, n):
        """(..., T, N, sum_l n_l): K_task[task, l] k(x, X_l) over all training tasks l (abstract_gp.py:375)."""
        kt = self.gram_matrix_tasks
        return torch.stack([torch.cat([kt[..., int(t), l, None, None] * self._kernel(x[:, None, :], self.get_x(l, int(n[l]))[None, :, :], self.derivatives[int(t)],
                                                                                    self.derivatives[l], self.derivatives_coeffs[int(t)], self.derivatives_coeffs[l])
                                       for l in range(self.num_tasks)], -1) for t in task], -3)

    def post_mean(self, x, task=None, eval=True):
        assert x.ndim == 2 and x.size(1) == self.d, "x must a torch.Tensor with shape (-1,d)"
        inttask, task = self._parse_task(task)
        with torch.set_grad_enabled(torch.is_grad_enabled() and not eval):
            coeffs = self.coeffs
            pmean = torch.einsum("...i,...i->...", self._cross(x.to(self.device), task, self.n), coeffs[..., None, None, :])
        return pmean[..., 0, :] if inttask else pmean

    def post_var(self, x, task=None, n=None, eval=True):
        n = self._n_tensor(n)
        assert x.ndim == 2 and x.size(1) == self.d, "x must a torch.Tensor with shape (-1,d)"
        inttask, task = self._parse_task(task)
        x = x.to(self.device)
        with torch.set_grad_enabled(torch.is_grad_enabled() and not eval):
            kt = self.gram_matrix_tasks
            knew = torch.stack([kt[..., int(t), int(t), None] * self._kernel(x, x, self.derivatives[int(t)], self.derivatives[int(t)],
                                                                            self.derivatives_coeffs[int(t)], self.derivatives_coeffs[int(t)]) for t in task], -2)
            km = self._cross(x, task, n)
            thetainv, _ = self.get_inv_log_det_cache(n)()
            sol = torch.einsum("...ij,...tnj->...tni", thetainv, km)
            pvar = (knew - (sol * km).sum(-1)).clamp(min=0)
        return pvar[..., 0, :] if inttask else pvar

    def post_cov(self, x0, x1, task0=None, task1=None, n=None, eval=True):
        n = self._n_tensor(n)
        assert x0.ndim == 2 and x0.size(1) == self.d, "x must a torch.Tensor with shape (-1,d)"
        assert x1.ndim == 2 and x1.size(1) == self.d, "z must a torch.Tensor with shape (-1,d)"
        inttask0, task0 = self._parse_task(task0)
        inttask1, task1 = self._parse_task(task1)
        x0, x1 = x0.to(self.device), x1.to(self.device)
        equal = torch.equal(x0, x1) and torch.equal(task0, task1)
        with torch.set_grad_enabled(torch.is_grad_enabled() and not eval):
            kt = self.gram_matrix_tasks
            knew = torch.stack([torch.stack([kt[..., int(t0), int(t1), None, None] * self._kernel(x0[:, None, :], x1[None, :, :], self.derivatives[int(t0)], self.derivatives[int(t1)],
                                                                                                  self.derivatives_coeffs[int(t0)], self.derivatives_coeffs[int(t1)])
                                             for t1 in task1], -3) for t0 in task0], -4)
            k1 = self._cross(x0, task0, n)
            k2 = k1 if equal else self._cross(x1, task1, n)
            thetainv, _ = self.get_inv_log_det_cache(n)()
            sol = torch.einsum("...ij,...tmj->...tmi", thetainv, k2)
            kmat = knew - torch.einsum("...sni,...tmi->...stnm", k1, sol)
            if equal:
                dg = torch.diagonal(torch.diagonal(kmat, dim1=-4, dim2=-3), dim1=-3, dim2=-2)  # views: (..., T, N) diagonal entries
                dg.clamp_(min=0)
        if inttask0 and inttask1:
            return kmat[..., 0, 0, :, :]
        if inttask0:
            return kmat[..., 0, :, :, :]
        if inttask1:
            return kmat[..., :, 0, :, :]
        return kmat

    # ------------------------------------------------------------------------------------------------ cubature (Gaussian kernel)
    def _kints(self, task, n, unit_cube):
        """(..., T, sum_l n_l): integrals of K_task[task, l] k(., X_l) over the unit cube (or R^d), standard_gp.py:357-362."""
        assert self.kernel_class == "gaussian", "so far, we have only worked out integrals for the Gaussian kernel"
        ls = self.lengthscales[..., None, :]
        lo, hi = (0.0, 1.0) if unit_cube else (-float("inf"), float("inf"))
        kt = self.gram_matrix_tasks
        parts = []
        for l in range(self.num_tasks):
            nrm = torch.distributions.Normal(self.get_x(l, int(n[l])), torch.sqrt(ls))
            mass = nrm.cdf(torch.tensor([hi], device=self.device)) - nrm.cdf(torch.tensor([lo], device=self.device))
            parts.append(self.scale * (torch.sqrt(2 * torch.pi * ls) * mass).prod(-1))
        return torch.cat([kt[..., task, l, None] * parts[l][..., None, :] for l in range(self.num_tasks)], -1)

    def _kint2(self):
        l_d = self.lengthscales + torch.zeros(self.d, device=self.device)
        t = 2 * (-1 + torch.exp(-1 / (2 * l_d))) * l_d + torch.sqrt(2 * np.pi * l_d) * torch.erf(1 / torch.sqrt(2 * l_d))
        return t.prod(-1)

    def post_cubature_mean(self, task=None, eval=True, integrate_unit_cube=True):
        inttask, task = self._parse_task(task)
        with torch.set_grad_enabled(torch.is_grad_enabled() and not eval):
            pcmean = (self._kints(task, self.n, integrate_unit_cube) * self.coeffs[..., None, :]).sum(-1)
        return pcmean[..., 0] if inttask else pcmean

    def post_cubature_var(self, task=None, n=None, eval=True, integrate_unit_cube=True):
        assert integrate_unit_cube, "undefinted posterior variance when integrating first term over all reals"
        n = self._n_tensor(n)
        inttask, task = self._parse_task(task)
        with torch.set_grad_enabled(torch.is_grad_enabled() and not eval):
            thetainv, _ = self.get_inv_log_det_cache(n)()
            kints = self._kints(task, n, True)
            v = torch.einsum("...ij,...tj->...ti", thetainv, kints)
            kt = self.gram_matrix_tasks
            tval = self.scale * kt[..., task, task] * self._kint2()[..., None]
            pcvar = (tval - (kints * v).sum(-1)).clamp(min=0)
        return pcvar[..., 0] if inttask else pcvar

    def post_cubature_cov(self, task0=None, task1=None, n=None, eval=True, integrate_unit_cube=True):
        assert integrate_unit_cube, "undefinted posterior variance when integrating first term over all reals"
        n = self._n_tensor(n)
        inttask0, task0 = self._parse_task(task0)
        inttask1, task1 = self._parse_task(task1)
        with torch.set_grad_enabled(torch.is_grad_enabled() and not eval):
            thetainv, _ = self.get_inv_log_det_cache(n)()
            k0, k1 = self._kints(task0, n, True), self._kints(task1, n, True)
            v = torch.einsum("...ij,...tj->...ti", thetainv, k1)
            kt = self.gram_matrix_tasks
            tval = self.scale[..., None] * kt[..., task0, :][..., :, task1] * self._kint2()[..., None, None]
            pccov = tval - torch.einsum("...si,...ti->...st", k0, v)
            if torch.equal(task0, task1):
                torch.diagonal(pccov, dim1=-2, dim2=-1).clamp_(min=0)
        if inttask0 and inttask1:
            return pccov[..., 0, 0]
        if inttask0:
            return pccov[..., 0, :]
        if inttask1:
            return pccov[..., :, 0]
        return pccov

    # fast-transform seams that do not exist for a dense GP
    def _ft_unstable(self, x):
        raise NotImplementedError("StandardGP has no fast transform")

    _ift_unstable = _ft_unstable
