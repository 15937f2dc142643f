"""TEST INFRASTRUCTURE ONLY -- CPU (torch float64) restatement of the reference's structured-covariance hot path
for the single-task, no-derivative case.

Used as (i) the parity checker for the CUDA path in tests/ and __graft_entry__.smoke(), and (ii) the CPU baseline
that bench.py times on the GPU box's host cores (`cpu_baseline.kind = "port"`, `--impl reference`), because
/root/reference and qmcpy do not exist there.  The product package never imports it.

It deliberately keeps the reference's algorithmic shape: materialised kernel parts, log2(n)-pass torch transforms,
autograd backward through the transform, torch.optim.Rprop, per-iteration `.item()`.

PARITY STATUS: pinned against the UNMODIFIED reference (fastgps/*.py run on oracle/qmcpy_standin) through the
fixtures in tests/golden/ (tests/test_oracle.py); the qmcpy primitives underneath are restated, see primitives.py.
"""
import numpy as np
import torch

from . import primitives as P


class OracleFastGP:
    """family: "lattice" (FastGPLattice) or "dnb2" (FastGPDigitalNetB2)."""

    def __init__(self, family, x, xb=None, t=None, alpha=2, scale=1.0, lengthscales=1.0, noise=None):
        assert family in ("lattice", "dnb2")
        assert torch.get_default_dtype() == torch.float64  # abstract_gp.py:46
        self.family = family
        self.x = torch.as_tensor(x)
        self.n, self.d = self.x.shape
        self.xb = self.x if family == "lattice" else torch.as_tensor(xb)  # _XBDTYPE fast_gp_lattice.py:123 / fast_gp_digital_net_b2.py:118
        self.t = t
        self.alpha = int(alpha)
        if noise is None:
            noise = 1e-8 if family == "lattice" else 1e-16  # ctor defaults fast_gp_lattice.py:132, fast_gp_digital_net_b2.py:127
        ls = torch.as_tensor(lengthscales, dtype=torch.float64) * torch.ones(self.d)
        # raw parameters through the default (log, exp) transforms, abstract_gp.py:88,100,111
        self.raw_scale = torch.log(torch.tensor([float(scale)])).requires_grad_(True)
        self.raw_lengthscales = torch.log(ls).clone().requires_grad_(True)
        self.raw_noise = torch.log(torch.tensor([float(noise)]))  # requires_grad_noise=False by default
        self.y = None
        self._k1parts = None

    # ---- hyperparameters (abstract_gp.py:622-639)
    @property
    def scale(self):
        return torch.exp(self.raw_scale)

    @property
    def lengthscales(self):
        return torch.exp(self.raw_lengthscales)

    @property
    def noise(self):
        return torch.exp(self.raw_noise)

    # ---- kernels
    def _ominus(self, x, z):
        if self.family == "lattice":
            return (x - z) % 1  # fast_gp_lattice.py:263-266
        def to_b(v):  # fast_gp_digital_net_b2.py:270-271
            return torch.floor((v % 1) * 2 ** self.t).to(torch.int64) if torch.is_floating_point(v) else v
        return to_b(x) ^ to_b(z)  # fast_gp_digital_net_b2.py:274-288

    def _parts_from_delta(self, delta):
        if self.family == "lattice":  # fast_gp_lattice.py:267-273 with beta=kappa=0
            order = 2 * self.alpha
            coeff = (-1) ** (self.alpha + 1) * np.exp(2 * self.alpha * np.log(2 * np.pi) - float(torch.lgamma(torch.tensor(order + 1.0))))
            return coeff * P.bernoulli_poly(order, delta)
        # fast_gp_digital_net_b2.py:289-301 with beta=kappa=0
        if self.alpha == 1:
            return 6 * (1 / 6 - 2 ** (torch.log2(delta).floor() - self.t - 1))
        return P.weighted_walsh_funcs(self.alpha, delta, self.t) - 1

    def kernel_parts(self, x, z):
        return self._parts_from_delta(self._ominus(x, z))  # abstract_fast_gp.py:173-180

    def kernel_from_parts(self, parts):
        return self.scale * (1 + self.lengthscales * parts).prod(-1)  # abstract_fast_gp.py:181-191

    def kernel(self, x, z):
        return self.kernel_from_parts(self.kernel_parts(x, z))

    def k1parts(self):
        if self._k1parts is None:  # util.py:50-62
            self._k1parts = self.kernel_parts(self.xb, self.xb[:1]).detach()
        return self._k1parts

    # ---- transforms (abstract_fast_gp.py:197-228)
    def _ft_unstable(self, x):
        return P.fftbr_torch(x) if self.family == "lattice" else P.fwht_torch(x)

    def _ift_unstable(self, x):
        return P.ifftbr_torch(x) if self.family == "lattice" else P.fwht_torch(x)

    def ft(self, x):
        xmean = x.mean(-1)
        y = self._ft_unstable(x - xmean[..., None])
        y[..., 0] += xmean * np.sqrt(x.size(-1))
        return y

    def ift(self, x):
        xmean = x.mean(-1)
        y = self._ift_unstable(x - xmean[..., None])
        y[..., 0] += xmean * np.sqrt(x.size(-1))
        return y

    # ---- data
    def add_y(self, y):
        self.y = torch.as_tensor(y)
        assert self.y.shape[-1] == self.n
        self.ytilde = self.ft(self.y)  # util.py:168-172

    # ---- eigen-solve
    def lam(self):
        k1 = self.kernel_from_parts(self.k1parts())  # util.py:102
        return self.ft(k1)  # util.py:103

    def full_lam(self):
        return np.sqrt(self.n) * self.lam() + self.noise  # util.py:285,293 (K_task = [[1]])

    def norm_logdet(self):
        lam = self.full_lam()
        logdet = torch.log(torch.abs(lam)).sum(-1)  # util.py:299
        ztilde = self.ytilde * (1 / lam)  # util.py:300,357-360
        norm = (self.ytilde.conj() * ztilde).real.sum(-1)  # util.py:369
        return norm, logdet

    def mll_loss(self):
        norm, logdet = self.norm_logdet()
        d_out = int(np.prod(self.y.shape[:-1])) if self.y.ndim > 1 else 1
        norm = norm.sum()
        logdet = d_out * logdet.sum() / max(logdet.numel(), 1) if logdet.ndim > 0 else d_out * logdet
        return 0.5 * (norm + logdet + d_out * self.n * np.log(2 * np.pi)), norm, logdet  # abstract_gp.py:235,255-260

    def solve(self, y):
        yt = self.ft(y)
        return self.ift(yt / self.full_lam()).real  # util.py:338-344

    def coeffs(self):
        return self.solve(self.y)  # util.py:419-425

    # ---- fit (abstract_gp.py:236-298, defaults)
    def fit(self, iterations=5000, lr=1e-1, stop_crit_improvement_threshold=5e-2, stop_crit_wait_iterations=10, store_hist=True):
        params = [p for p in (self.raw_scale, self.raw_lengthscales, self.raw_noise) if p.requires_grad]
        opt = torch.optim.Rprop(params, lr=lr)  # abstract_fast_gp.py:53-57
        logtol = np.log(1 + stop_crit_improvement_threshold)
        best, save, wait = np.inf, np.inf, 0
        hist = []
        for i in range(iterations + 1):
            loss, _, _ = self.mll_loss()
            lv = loss.item()
            if lv < best:
                best = lv
                best_params = [p.data.clone() for p in (self.raw_scale, self.raw_lengthscales, self.raw_noise)]
            if (save - lv) > logtol:
                wait = 0
                save = best
            else:
                wait += 1
            stop = i == iterations or wait == stop_crit_wait_iterations
            if store_hist:
                hist.append(lv)
            if stop:
                break
            loss.backward()
            opt.step()
            opt.zero_grad()
        for p, b in zip((self.raw_scale, self.raw_lengthscales, self.raw_noise), best_params):
            p.data.copy_(b)
        return {"iterations": i, "loss_hist": np.asarray(hist)}

    # ---- posterior (abstract_gp.py:352-416)
    @torch.no_grad()
    def post_mean(self, xs, coeffs=None, chunk=None):
        c = self.coeffs() if coeffs is None else coeffs
        xs = torch.as_tensor(xs)
        if chunk is None:
            chunk = max(1, int(2 ** 27 // (self.n * self.d)))  # keep the materialised (chunk,n,d) parts near 1 GiB
        out = []
        for s in range(0, xs.shape[0], chunk):
            kmat = self.kernel(xs[s:s + chunk, None, :], self.xb[None, :, :])  # abstract_gp.py:375
            out.append(torch.einsum("...i,...i->...", kmat, c))  # abstract_gp.py:377
        return torch.cat(out, -1)

    @torch.no_grad()
    def post_var(self, xs, chunk=None):
        xs = torch.as_tensor(xs)
        if chunk is None:
            chunk = max(1, int(2 ** 27 // (self.n * self.d)))
        out = []
        for s in range(0, xs.shape[0], chunk):
            xc = xs[s:s + chunk]
            knew = self.kernel(xc, xc)  # abstract_gp.py:407
            kmat = self.kernel(xc[:, None, :], self.xb[None, :, :])  # abstract_gp.py:408
            tmat = self.solve(kmat)  # abstract_gp.py:409-411
            diag = knew - (tmat * kmat).sum(-1)  # abstract_gp.py:412
            diag[diag < 0] = 0  # abstract_gp.py:413
            out.append(diag)
        return torch.cat(out, -1)
