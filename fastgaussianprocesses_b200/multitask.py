"""Multi-task block eigen-solve (reference util.py:275-370 multi-task branch,
abstract_gp.py:352-474, abstract_fast_gp.py:65-154) for tasks that all hold the same number of points.

Every block K_task[l0,l1] * K(X_l0, X_l1) of the Gram matrix of T randomisations of one lattice / digital net is
diagonalised by the same fast transform, so the whole matrix reduces to n independent T x T Hermitian systems
    Lam_k[l0,l1] = K_task[l0,l1] (sqrt(n) ft(k1^(l0,l1))_k + noise [l0 == l1]).
The transforms are the CUDA kernels of libfgp_b200 behind torch.autograd (`_FTFunction`); the kernel parts, cross
kernels and posterior-mean products are the K2 / K5 kernels; the n small T x T factorizations are batched torch.linalg
calls (library code, as the reference's own Schur-complement recursion is torch code).  This path is parity-tested
against reference fixtures but not fused or tuned; unequal task sizes raise NotImplementedError.
"""
import numpy as np
import torch

from . import _lib


class MultiTaskEngine(object):
    def __init__(self, gp):
        self.gp = gp
        self._parts = {}
        self._lam_key = None
        self._solve_cache = None

    # ------------------------------------------------------------------------------------------------ sizes
    @property
    def T(self):
        return self.gp.num_tasks

    def n_equal(self, n=None):
        gp = self.gp
        ns = [int(v) for v in (gp.n if n is None else n).tolist()]
        if len(set(ns)) != 1 or ns[0] == 0:
            raise NotImplementedError("multi-task solves need the same (non-zero) number of points in every task (got %s)" % ns)
        return ns[0]

    def xpts(self, l, n):
        x, xb = self.gp.xxb_seqs[l][:int(n)]
        return (xb if self.gp._FAMILY == 1 else x).contiguous()

    # ------------------------------------------------------------------------------------------------ spectrum
    def parts(self, l0, l1, n):
        """(n,d) kernel parts of the points of task l0 against the FIRST point of task l1 (util.py:50-62)."""
        key = (l0, l1, int(n))
        p = self._parts.get(key)
        if p is None:
            gp = self.gp
            x0, xb0 = gp.xxb_seqs[l0][:int(n)]
            x1, xb1 = gp.xxb_seqs[l1][:1]
            with torch.no_grad():
                if gp._FAMILY == 0:
                    p = _lib.lattice_kernel_parts(x0.contiguous(), x1[0].cpu().numpy(), gp._alpha_list)
                else:
                    p = _lib.dnb2_kernel_parts(xb0.contiguous(), xb1[0].cpu().numpy(), gp._alpha_list, gp._t)
            self._parts[key] = p
        return p

    def lam_blocks(self, n):
        """Lam (n, T, T), differentiable w.r.t. every raw parameter (util.py:279-298)."""
        gp, T = self.gp, self.T
        scale, ls, noise, kt = gp.scale, gp.lengthscales, gp.noise, gp.gram_matrix_tasks
        rows = [[None] * T for _ in range(T)]
        for l0 in range(T):
            for l1 in range(l0, T):
                k1 = scale * (1 + ls[..., None, :] * self.parts(l0, l1, n)).prod(-1)
                lam = np.sqrt(n) * gp.ft(k1)
                if l0 == l1:
                    lam = lam + noise
                rows[l0][l1] = lam * kt[..., l0, l1, None]
                if l1 > l0:
                    rows[l1][l0] = rows[l0][l1].conj()
        L = torch.stack([torch.stack(r, -1) for r in rows], -2)  # (n, T, T): [k, l0, l1]
        return L

    def factor(self, n, grad=False):
        """(Lam^-1 (n,T,T), logdet) -- cached on the hyperparameter state when no gradient is needed."""
        if grad:
            L = self.lam_blocks(n)
            return torch.linalg.inv(L), torch.linalg.slogdet(L)[1].sum(-1)
        key = (int(n),) + self.gp._param_key()
        if self._lam_key != key:
            with torch.no_grad():
                L = self.lam_blocks(n)
                self._solve_cache = (torch.linalg.inv(L), torch.linalg.slogdet(L)[1].sum(-1))
            self._lam_key = key
        return self._solve_cache

    def ytilde(self):
        gp = self.gp
        n = self.n_equal()
        key = (n, id(gp._y[0]), tuple(y.data_ptr() for y in gp._y))
        if getattr(self, "_yt_key", None) != key:
            with torch.no_grad():
                self._yt = torch.stack([gp.ft(gp._y[l]) for l in range(self.T)], -2)  # (T, n)
            self._yt_key = key
        return self._yt

    def solve_tilde(self, A, zt):
        """A (n,T,T), zt (..., T, n) -> (..., T, n)."""
        return torch.einsum("kij,...jk->...ik", A, zt.to(A.dtype))

    def gram_matrix_solve(self, y, n=None, A=None):
        """K^-1 y for y (..., T*n) (util.py:338-344 multi-task)."""
        gp = self.gp
        n = self.n_equal() if n is None else int(n)
        if A is None:
            A, _ = self.factor(n)
        y = y.to(gp.device)
        ys = y.reshape(y.shape[:-1] + (self.T, n))
        zt = self.solve_tilde(A, gp.ft(ys))
        return gp.ift(zt).real.reshape(y.shape)

    def norm_logdet(self, grad):
        n = self.n_equal()
        A, logdet = self.factor(n, grad=grad)
        yt = self.ytilde()
        zt = self.solve_tilde(A, yt)
        norm = (yt.conj() * zt).real.sum((-1, -2))[..., None]
        return norm, logdet[..., None], A, zt

    def loss(self, loss_metric, d_out, mll_const):
        """The reference's losses on the block spectrum (abstract_gp.py:242-273); returns (loss, term1, term2, metric_val)."""
        n = self.n_equal()
        norm, logdet, A, zt = self.norm_logdet(grad=True)
        if loss_metric == "MLL":
            term1 = norm.sum()
            term2 = d_out / logdet.numel() * logdet.sum()
            loss = 1 / 2 * (term1 + term2 + mll_const)
            return loss, term1, term2, -loss
        if loss_metric == "GCV":
            numer = (zt.conj() * zt).real.sum((-1, -2))[..., None]
            tr_k_inv = torch.diagonal(A, dim1=-2, dim2=-1).real.sum((-1, -2))[..., None]
            denom = ((tr_k_inv / (self.T * n)) ** 2).real
            loss = (numer / denom).sum()
            return loss, numer, denom, loss
        raise NotImplementedError("loss_metric='CV' needs the O(n^2 log n) inverse diagonal of the reference for several tasks (util.py:386-393); not built")

    # ------------------------------------------------------------------------------------------------ posterior
    def _host(self):
        gp = self.gp
        with torch.no_grad():
            assert gp.scale.numel() == 1 and gp.lengthscales.ndim == 1 and gp.noise.numel() == 1, "multi-task GPs take one hyperparameter set"
            ls = gp.lengthscales.expand(gp.d) if gp.lengthscales.numel() == 1 else gp.lengthscales
            return float(gp.scale.reshape(-1)[0]), ls.cpu().numpy(), gp.gram_matrix_tasks.cpu().numpy()

    def coeffs(self):
        gp = self.gp
        key = (tuple(int(v) for v in gp.n.tolist()),) + gp._param_key()
        if getattr(self, "_coeffs_key", None) != key:
            with torch.no_grad():
                self._coeffs = self.gram_matrix_solve(torch.cat(gp._y, -1))
            self._coeffs_key = key
        return self._coeffs

    def post_mean(self, x, task):
        gp, T = self.gp, self.T
        n = self.n_equal()
        scale, ls, kt = self._host()
        c = self.coeffs().reshape(T, n)
        N = x.shape[0]
        if N == 0:
            return torch.empty((len(task), 0), dtype=torch.float64, device=gp.device)
        # one on-the-fly kernel-vector product per training task, then the T x T task kernel mixes them
        base = torch.stack([_lib.post_mean(gp._FAMILY, x, self.xpts(l, n), gp._alpha_list, gp._t, scale, ls, c[l:l + 1].contiguous())[0] for l in range(T)], 0)
        ktd = torch.from_numpy(kt).to(gp.device)
        return ktd[task.to(gp.device)] @ base  # (len(task), N)

    def _cross_rows(self, x, t, n, scale, ls, kt):
        """K_task[t, l1] k(x, X_l1) for all l1, concatenated: (N, T*n)."""
        gp = self.gp
        return torch.cat([kt[t, l] * _lib.cross_kernel(gp._FAMILY, x, self.xpts(l, n), gp._alpha_list, gp._t, scale, ls) for l in range(self.T)], -1)

    def _kxx(self, scale, ls):
        gp = self.gp
        one = torch.zeros((1, gp.d), dtype=torch.float64, device=gp.device)
        return float(_lib.kernel_pairs(gp._FAMILY, one, one.clone(), gp._alpha_list, gp._t, scale, ls)[0])

    def post_var(self, x, task, n):
        gp = self.gp
        scale, ls, kt = self._host()
        A, _ = self.factor(n)
        kxx = self._kxx(scale, ls)
        out = []
        with torch.no_grad():
            for t in task.tolist():
                km = self._cross_rows(x, t, n, scale, ls, kt)
                sol = self.gram_matrix_solve(km, n=n, A=A)
                out.append((kt[t, t] * kxx - (sol * km).sum(-1)).clamp_(min=0))
        return torch.stack(out, 0)

    def post_cov(self, x0, x1, task0, task1, n, equal):
        gp = self.gp
        scale, ls, kt = self._host()
        A, _ = self.factor(n)
        fam, al, tt = gp._FAMILY, gp._alpha_list, gp._t
        with torch.no_grad():
            knew = _lib.cross_kernel(fam, x0, x1 if fam == 0 else gp._convert_to_b(x1), al, tt, scale, ls)
            k1 = {t: self._cross_rows(x0, t, n, scale, ls, kt) for t in set(task0.tolist())}
            sol2 = {t: self.gram_matrix_solve(k1[t] if (equal and t in k1) else self._cross_rows(x1, t, n, scale, ls, kt), n=n, A=A) for t in set(task1.tolist())}
            out = torch.empty((len(task0), len(task1), x0.shape[0], x1.shape[0]), dtype=torch.float64, device=gp.device)
            for i0, t0 in enumerate(task0.tolist()):
                for i1, t1 in enumerate(task1.tolist()):
                    out[i0, i1] = kt[t0, t1] * knew - k1[t0] @ sol2[t1].T
                    if equal and t0 == t1 and i0 == i1:
                        out[i0, i1].diagonal().clamp_(min=0)
        return out

    def post_cubature_mean(self, task):
        gp = self.gp
        n = self.n_equal()
        scale, _, kt = self._host()
        with torch.no_grad():
            sums = scale * self.coeffs().reshape(self.T, n).sum(-1)  # (T)
            return torch.from_numpy(kt).to(gp.device)[task.to(gp.device)] @ sums

    def post_cubature_cov(self, task0, task1, n):
        """abstract_fast_gp.py:110-154 for equal task sizes: scale K_task - scale^2 K_task (n A_0) K_task."""
        gp = self.gp
        scale, _, kt = self._host()
        with torch.no_grad():
            A, _ = self.factor(n)
            ktd = torch.from_numpy(kt).to(gp.device).to(A.dtype)
            term = (ktd[task0.to(gp.device)] @ (n * A[0]) @ ktd[:, task1.to(gp.device)]).real
            return scale * ktd.real[task0.to(gp.device)][:, task1.to(gp.device)] - scale ** 2 * term


class MultiTaskInverseLogDetCache(object):
    """Strategy object of `get_inv_log_det_cache` for several tasks (util.py:275-394)."""

    def __init__(self, gp, n):
        self.fgp = gp
        self.n = n
        self.nint = gp._mt.n_equal(n)
        self.task_order = torch.arange(gp.num_tasks, device=gp.device)
        self.inv_task_order = torch.arange(gp.num_tasks, device=gp.device)

    def __call__(self):
        A, logdet = self.fgp._mt.factor(self.nint)
        return A.permute(1, 2, 0), logdet  # (T, T, n) as the reference lays it out

    def gram_matrix_solve(self, y):
        return self.fgp._mt.gram_matrix_solve(y, n=self.nint)

    def get_norm_term_logdet_term(self):
        norm, logdet, _, _ = self.fgp._mt.norm_logdet(grad=torch.is_grad_enabled())
        return norm, logdet

    def get_gcv_numer_denom(self):
        _, numer, denom, _ = self.fgp._mt.loss("GCV", 1, 0.0)
        return numer, denom
