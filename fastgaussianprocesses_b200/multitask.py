"""Multi-task block eigen-solve (reference util.py:275-370 multi-task branch,
abstract_gp.py:352-474, abstract_fast_gp.py:65-154) for tasks that all hold the same number of points.

Every block K_task[l0,l1] * K(X_l0, X_l1) of the Gram matrix of T randomisations of one lattice / digital net is
diagonalised by the same fast transform, so the whole matrix reduces to n independent T x T Hermitian systems
    Lam_k[l0,l1] = K_task[l0,l1] (sqrt(n) ft(k1^(l0,l1))_k + noise [l0 == l1]).
The transforms are the CUDA kernels of libfgp_b200 behind torch.autograd (`_FTFunction`); the kernel parts, cross
kernels and posterior-mean products are the K2 / K5 kernels; the n small R x R systems are inverted by the
fgp_block_inv_logdet kernel (one thread per system, Gauss-Jordan with partial pivoting in registers; `_BlockInvLogdet`).  Tasks of different (power-of-two)
sizes fold into n_min independent R x R systems, R = sum_l n_l / n_min, because sub-sampling a lattice aliases the
frequencies kappa and kappa mod n_l.  This path is parity-tested against reference fixtures but not fused or tuned.

Batched outputs (shape_batch): y_l is (*shape_batch, n_l); every hyperparameter may carry leading dimensions that match a
tail of shape_batch.  Bh below is the broadcast of those leading dimensions (empty for one shared hyperparameter set): the
systems are (*Bh, n_min, R, R), data-dependent results (coeffs, posterior mean) are (*shape_batch, ...), data-independent
ones (variances, covariances) are (*Bh, ...), as in the reference (abstract_gp.py:352-474).
"""
import numpy as np
import torch

from . import _lib


class _BlockInvLogdet(torch.autograd.Function):
    """A_k = L_k^-1 and log|det L_k| of the n_min per-frequency R x R systems, computed by the fgp_block_inv_logdet kernel (one thread per
    system, Gauss-Jordan with partial pivoting in registers).  Backward: dL = -A^H gA A^H + g_logdet A^H, two small batched products on the
    saved inverse."""

    @staticmethod
    def forward(ctx, L):
        A, logdet = _lib.block_inv_logdet(L)
        ctx.save_for_backward(A)
        return A, logdet

    @staticmethod
    def backward(ctx, gA, gld):
        (A,) = ctx.saved_tensors
        AH = A.mH
        g = None
        if gA is not None:
            g = -torch.einsum("kij,kjl,klm->kim", AH, gA.to(A.dtype), AH)
        if gld is not None:
            t = gld.to(A.dtype)[:, None, None] * AH
            g = t if g is None else g + t
        return g


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

    def sizes(self, n=None):
        """Per-task sizes and the layout of the folded system (util.py:277-323): tasks in order of decreasing size; a task
        of n_l points occupies r_l = n_l / n_min consecutive rows of the (R, n_min) frequency layout, row a holding the
        natural-order frequencies [a n_min, (a+1) n_min)."""
        gp = self.gp
        if n is None:
            ns = [int(v) for v in gp.n.tolist()]
        elif isinstance(n, (int, np.integer)):
            ns = [int(n)] * self.T
        else:
            ns = [int(v) for v in (n.tolist() if isinstance(n, torch.Tensor) else n)]
        assert all(v == 0 or (v & (v - 1)) == 0 for v in ns), "task sizes must be powers of two"
        active = [l for l in sorted(range(self.T), key=lambda l: -ns[l]) if ns[l] > 0]
        assert active, "no data"
        nmin = min(ns[l] for l in active)
        r = {l: ns[l] // nmin for l in active}
        off, acc = {}, 0
        for l in active:
            off[l] = acc
            acc += r[l]
        return ns, active, nmin, r, off, acc

    def xpts(self, l, n):
        x, xb = self.gp.xxb_seqs[l][:int(n)]
        return (xb if self.gp._FAMILY == 1 else x).contiguous()

    # ------------------------------------------------------------------------------------------------ spectrum
    def parts(self, l0, l1, n):
        """(n,d) kernel parts of the first n points of task l0 against the FIRST point of task l1 (util.py:50-62)."""
        key = (l0, l1, int(n))
        p = self._parts.get(key)
        if p is None:
            gp = self.gp
            x0, xb0 = gp.xxb_seqs[l0][:int(n)]
            x1, xb1 = gp.xxb_seqs[l1][:1]
            with torch.no_grad():
                if gp._has_derivs:  # (n, terms, d): one slice per pair of derivative multi-indices
                    p = _lib.deriv_kernel_parts(gp._FAMILY, (x0 if gp._FAMILY == 0 else xb0).contiguous(),
                                                (x1 if gp._FAMILY == 0 else xb1)[0].cpu().numpy(), self.terms(l0, l1), gp._t)
                elif gp._FAMILY == 0:
                    p = _lib.lattice_kernel_parts(x0.contiguous(), x1[0].cpu().numpy(), gp._alpha_list)
                else:
                    p = _lib.dnb2_kernel_parts(xb0.contiguous(), xb1[0].cpu().numpy(), gp._alpha_list, gp._t)
            self._parts[key] = p
        return p

    def terms(self, l0, l1):
        gp = self.gp
        return gp._deriv_terms(gp.derivatives[l0], gp.derivatives[l1], gp.derivatives_coeffs[l0], gp.derivatives_coeffs[l1])

    def k1(self, l0, l1, n):
        """First kernel column of task l0 against the first point of task l1, differentiable w.r.t. scale and lengthscales
        (abstract_fast_gp.py:181-191: sum over the derivative terms of scale prod_j (ind_j + ls_j part_j))."""
        gp = self.gp
        scale, ls = gp.scale, gp.lengthscales
        p = self.parts(l0, l1, n)
        if not gp._has_derivs:
            return scale * (1 + ls[..., None, :] * p).prod(-1)
        tm = self.terms(l0, l1)
        return scale * ((tm.ind + ls[..., None, None, :] * p).prod(-1) * tm.w).sum(-1)

    def cross(self, x, t, l, n, scale, ls):
        """k(x, X_l[:n]) for test task t (N, n)."""
        gp = self.gp
        if gp._has_derivs:
            return _lib.deriv_cross_kernel(gp._FAMILY, x, self.xpts(l, n), self.terms(t, l), gp._t, scale, ls)
        return _lib.cross_kernel(gp._FAMILY, x, self.xpts(l, n), gp._alpha_list, gp._t, scale, ls)

    def hyper_shape(self):
        """Bh: broadcast of the leading (batch) dimensions of the five hyperparameters."""
        gp = self.gp
        return tuple(torch.broadcast_shapes(gp.raw_scale.shape[:-1], gp.raw_lengthscales.shape[:-1], gp.raw_noise.shape[:-1],
                                            gp.raw_factor_task_kernel.shape[:-2], gp.raw_noise_task_kernel.shape[:-1]))

    def lam_system(self, n=None):
        """Lam (*Bh, n_min, R, R), differentiable w.r.t. every raw parameter.  Block (t0, t1), n_t0 >= n_t1: the length-n_t0 vector
        lam = K_task[t0,t1] sqrt(n_t1) ft(k1^(t0,t1)) (+ noise on the diagonal) couples frequency kappa of task t0 with
        frequency kappa mod n_t1 of task t1 (sub-sampling aliases frequencies), util.py:279-323."""
        gp = self.gp
        ns, active, nmin, r, off, R = self.sizes(n)
        scale, ls, noise, kt = gp.scale, gp.lengthscales, gp.noise, gp.gram_matrix_tasks
        dev = gp.device
        Bh = self.hyper_shape()
        tr00 = None
        if gp.adaptive_nugget:  # util.py:286-290: the noise of task l is scaled by |trace(block l) / trace(block 0)|
            assert ns[0] > 0, "adaptive_nugget needs data for task 0"
            tr00 = (np.sqrt(ns[0]) * gp.ft(self.k1(0, 0, ns[0]))).sum(-1, keepdim=True)
        bidx, rows, cols, vals = [], [], [], []
        ar = torch.arange(nmin, device=dev)
        for i0, t0 in enumerate(active):
            for t1 in active[i0:]:
                nbig = ns[t0]
                if t0 <= t1:
                    lam = gp.ft(self.k1(t0, t1, nbig))
                else:  # the reference keeps one spectrum per unordered pair and conjugates it (util.py:284)
                    lam = gp.ft(self.k1(t1, t0, nbig)).conj()
                lam = np.sqrt(ns[t1]) * lam
                if t0 == t1:
                    lam = lam + (noise if tr00 is None else noise * (lam.sum(-1, keepdim=True) / tr00).abs())
                V = (lam * kt[..., t0, t1, None]).expand(Bh + (nbig,)).reshape(Bh + (r[t0], nmin)).movedim(-1, 0).movedim(-1, 0)  # (r, n_min, *Bh)
                for a0 in range(r[t0]):
                    a1 = a0 if t0 == t1 else a0 % r[t1]
                    bidx.append(ar)
                    rows.append(torch.full((nmin,), off[t0] + a0, device=dev))
                    cols.append(torch.full((nmin,), off[t1] + a1, device=dev))
                    vals.append(V[a0])
                    if t0 != t1:
                        bidx.append(ar)
                        rows.append(torch.full((nmin,), off[t1] + a1, device=dev))
                        cols.append(torch.full((nmin,), off[t0] + a0, device=dev))
                        vals.append(V[a0].conj())
        vals = torch.cat(vals)  # (entries * n_min, *Bh)
        L = torch.zeros((nmin, R, R) + Bh, dtype=vals.dtype, device=dev)
        L = L.index_put((torch.cat(bidx), torch.cat(rows), torch.cat(cols)), vals)
        return L.permute(tuple(range(3, 3 + len(Bh))) + (0, 1, 2)) if len(Bh) else L

    def factor(self, n=None, grad=False):
        """(Lam^-1 (*Bh,n_min,R,R), logdet (*Bh)) -- cached on the hyperparameter state when no gradient is needed."""
        if grad:
            L = self.lam_system(n)
            A, ld = _BlockInvLogdet.apply(L.reshape((-1,) + L.shape[-2:]))
            return A.reshape(L.shape), ld.reshape(L.shape[:-2]).sum(-1)
        key = (tuple(self.sizes(n)[0]),) + self.gp._param_key()
        if self._lam_key != key:
            with torch.no_grad():
                L = self.lam_system(n)
                A, ld = _lib.block_inv_logdet(L.reshape((-1,) + L.shape[-2:]).contiguous())
                self._solve_cache = (A.reshape(L.shape), ld.reshape(L.shape[:-2]).sum(-1))
            self._lam_key = key
        return self._solve_cache

    def fold(self, parts, n=None):
        """per-task spectra [(..., n_l)] in task order -> (..., R, n_min) in the folded layout."""
        ns, active, nmin, r, off, R = self.sizes(n)
        return torch.cat([parts[l].reshape(parts[l].shape[:-1] + (r[l], nmin)) for l in active], -2)

    def unfold(self, z, n=None):
        ns, active, nmin, r, off, R = self.sizes(n)
        out = [None] * self.T
        for l in active:
            out[l] = z[..., off[l]:off[l] + r[l], :].reshape(z.shape[:-2] + (ns[l],))
        for l in range(self.T):
            if out[l] is None:
                out[l] = z.new_zeros(z.shape[:-2] + (0,))
        return out

    def ytilde(self):
        gp = self.gp
        key = (tuple(int(v) for v in gp.n.tolist()), tuple(y.data_ptr() for y in gp._y))
        if getattr(self, "_yt_key", None) != key:
            with torch.no_grad():
                self._yt = self.fold([gp.ft(gp._y[l]) if gp._y[l].size(-1) > 0 else gp._y[l].to(gp._FTOUTDTYPE) for l in range(self.T)])
            self._yt_key = key
        return self._yt

    def solve_tilde(self, A, zt):
        """A (*Bh,n_min,R,R), zt (..., R, n_min) -> (..., R, n_min); the leading dimensions broadcast."""
        return torch.einsum("...kij,...jk->...ik", A, zt.to(A.dtype))

    def gram_matrix_solve(self, y, n=None, A=None):
        """K^-1 y for y (..., sum_l n_l), tasks concatenated in task order (util.py:338-344)."""
        gp = self.gp
        ns = self.sizes(n)[0]
        if A is None:
            A, _ = self.factor(n)
        y = y.to(gp.device)
        ys = y.split(ns, dim=-1)
        zt = self.solve_tilde(A, self.fold([gp.ft(v) if v.size(-1) > 0 else v.to(gp._FTOUTDTYPE) for v in ys], n))
        zs = self.unfold(zt, n)
        return torch.cat([gp.ift(v).real if v.size(-1) > 0 else v.real for v in zs], -1)

    def norm_logdet(self, grad):
        A, logdet = self.factor(None, grad=grad)
        yt = self.ytilde()
        zt = self.solve_tilde(A, yt)
        norm = (yt.conj() * zt).real.sum((-1, -2))[..., None]
        return norm, logdet[..., None], A, zt

    def loss(self, loss_metric, d_out, mll_const, masks=None, cv_weights=1):
        """The reference's losses on the block spectrum (abstract_gp.py:242-273); returns (loss, term1, term2, metric_val).  `masks`
        (one index row per leading batch dimension) restricts the sums to a subset of the batched outputs."""
        gp = self.gp
        sb = list(gp.shape_batch)
        norm, logdet, A, zt = self.norm_logdet(grad=True)
        if loss_metric == "MLL":
            if masks is None:
                term1 = norm.sum()
                term2 = d_out / logdet.numel() * logdet.sum()
            else:
                term1 = norm[..., *masks, 0].sum()
                term2 = logdet.expand(sb + [1])[..., *masks, 0].sum()
            loss = 1 / 2 * (term1 + term2 + mll_const)
            return loss, term1, term2, -loss
        if loss_metric == "GCV":
            numer = (zt.conj() * zt).real.sum((-1, -2))[..., None]
            tr_k_inv = torch.diagonal(A, dim1=-2, dim2=-1).real.sum((-1, -2))[..., None]
            denom = ((tr_k_inv / int(gp.n.sum())) ** 2).real
            if masks is None:
                term1, term2 = numer, denom
            else:
                term1 = numer[..., *masks, :]
                term2 = denom.expand(sb + [1])[..., *masks, :]
            loss = (term1 / term2).sum()
            return loss, term1, term2, loss
        # CV: leave-one-out residuals coeffs_i / (K^-1)_ii; the diagonal of K^-1 from solving against the identity, O(n^2 log n) as in the
        # reference (util.py:386-393)
        ns = self.sizes()[0]
        nsum = sum(ns)
        coeffs = torch.cat([gp.ift(v).real if v.size(-1) > 0 else v.real for v in self.unfold(zt)], -1)
        Bh = self.hyper_shape()
        eye = torch.eye(nsum, device=gp.device).reshape((nsum,) + (1,) * len(Bh) + (nsum,))
        kinv = self.gram_matrix_solve(eye, A=A)  # (nsum, *Bh, nsum)
        inv_diag = kinv.movedim(0, -2).diagonal(dim1=-2, dim2=-1)
        squared_sums = ((coeffs / inv_diag) ** 2 * cv_weights).sum(-1, keepdim=True)
        loss = squared_sums.sum() if masks is None else squared_sums[..., *masks, 0].sum()
        nan = torch.nan * torch.ones(1)
        return loss, nan, nan, loss

    # ------------------------------------------------------------------------------------------------ posterior
    def _host(self):
        """(Bh, [(scale, lengthscales (d), K_task (T,T)) per hyperparameter set, flattened over Bh]) on the host: the kernels of
        libfgp_b200 take one hyperparameter set per call."""
        gp = self.gp
        Bh = self.hyper_shape()
        with torch.no_grad():
            sc = gp.scale.expand(Bh + (1,)).reshape(-1).cpu().numpy()
            ls = gp.lengthscales.expand(Bh + (gp.d,)).reshape(-1, gp.d).cpu().numpy()
            kt = gp.gram_matrix_tasks.expand(Bh + (self.T, self.T)).reshape(-1, self.T, self.T).cpu().numpy()
        return Bh, [(float(sc[h]), np.ascontiguousarray(ls[h]), kt[h]) for h in range(len(sc))]

    def coeffs(self):
        gp = self.gp
        key = (tuple(int(v) for v in gp.n.tolist()),) + gp._param_key()
        if getattr(self, "_coeffs_key", None) != key:
            with torch.no_grad():
                self._coeffs = self.gram_matrix_solve(torch.cat(gp._y, -1))
            self._coeffs_key = key
        return self._coeffs

    def post_mean(self, x, task):
        """(*shape_batch, len(task), N)."""
        gp, T = self.gp, self.T
        ns = self.sizes()[0]
        Bh, sets = self._host()
        H = len(sets)
        sb = tuple(gp.shape_batch)
        N = x.shape[0]
        if N == 0:
            return torch.empty(sb + (len(task), 0), dtype=torch.float64, device=gp.device)
        # coefficient rows grouped by hyperparameter set: Bh is a tail of shape_batch, so (*shape_batch, n_l) = (lead, H, n_l)
        c = [v.reshape(-1, H, v.shape[-1]) for v in self.coeffs().split(ns, dim=-1)]
        lead = c[0].shape[0]
        out = torch.zeros((lead, H, len(task), N), dtype=torch.float64, device=gp.device)
        for h, (scale, ls, kt) in enumerate(sets):
            if gp._has_derivs:  # the kernel depends on the test task: chunked cross tiles times the coefficients
                step = max(1, (1 << 24) // max(1, max(ns)))
                for i, t in enumerate(task.tolist()):
                    for l in range(T):
                        if ns[l] == 0:
                            continue
                        for r0 in range(0, N, step):
                            out[:, h, i, r0:r0 + step] += kt[t, l] * (c[l][:, h, :] @ self.cross(x[r0:r0 + step].contiguous(), t, l, ns[l], scale, ls).T)
                continue
            # one on-the-fly kernel-vector product per training task, then the T x T task kernel mixes them
            base = torch.stack([_lib.post_mean(gp._FAMILY, x, self.xpts(l, ns[l]), gp._alpha_list, gp._t, scale, ls, c[l][:, h, :].contiguous())
                                if ns[l] > 0 else torch.zeros((lead, N), dtype=torch.float64, device=gp.device) for l in range(T)], 0)  # (T, lead, N)
            ktd = torch.from_numpy(kt).to(gp.device)
            out[:, h] = torch.einsum("it,tln->lin", ktd[task.to(gp.device)], base)
        return out.reshape(sb + (len(task), N))

    def _cross_rows(self, x, t, ns, scale, ls, kt):
        """K_task[t, l1] k(x, X_l1) for all l1, concatenated: (N, sum_l n_l)."""
        gp = self.gp
        return torch.cat([kt[t, l] * self.cross(x, t, l, ns[l], scale, ls) for l in range(self.T) if ns[l] > 0], -1)

    def _kxx(self, t, scale, ls):
        """k(x, x) of task t: a constant, the kernels are shift invariant."""
        gp = self.gp
        one = torch.zeros((1, gp.d), dtype=torch.float64, device=gp.device)
        if gp._has_derivs:
            z = one.clone() if gp._FAMILY == 0 else torch.zeros((1, gp.d), dtype=torch.int64, device=gp.device)
            return float(_lib.deriv_cross_kernel(gp._FAMILY, one, z, self.terms(t, t), gp._t, scale, ls)[0, 0])
        return float(_lib.kernel_pairs(gp._FAMILY, one, one.clone(), gp._alpha_list, gp._t, scale, ls)[0])

    def _knew(self, x0, x1, t0, t1, scale, ls):
        gp = self.gp
        z = x1 if gp._FAMILY == 0 else gp._convert_to_b(x1)
        if gp._has_derivs:
            return _lib.deriv_cross_kernel(gp._FAMILY, x0, z.contiguous(), self.terms(t0, t1), gp._t, scale, ls)
        return _lib.cross_kernel(gp._FAMILY, x0, z, gp._alpha_list, gp._t, scale, ls)

    def _inverse_sets(self, n):
        A, _ = self.factor(n)
        return A.reshape((-1,) + A.shape[-3:])

    def post_var(self, x, task, n):
        """(*Bh, len(task), N)."""
        Bh, sets = self._host()
        A = self._inverse_sets(n)
        ns = self.sizes(n)[0]
        out = []
        with torch.no_grad():
            for h, (scale, ls, kt) in enumerate(sets):
                for t in task.tolist():
                    km = self._cross_rows(x, t, ns, scale, ls, kt)
                    sol = self.gram_matrix_solve(km, n=n, A=A[h])
                    out.append((kt[t, t] * self._kxx(t, scale, ls) - (sol * km).sum(-1)).clamp_(min=0))
        return torch.stack(out, 0).reshape(Bh + (len(task), x.shape[0]))

    def post_cov(self, x0, x1, task0, task1, n, equal):
        """(*Bh, len(task0), len(task1), N0, N1)."""
        gp = self.gp
        Bh, sets = self._host()
        A = self._inverse_sets(n)
        ns = self.sizes(n)[0]
        out = torch.empty((len(sets), len(task0), len(task1), x0.shape[0], x1.shape[0]), dtype=torch.float64, device=gp.device)
        with torch.no_grad():
            for h, (scale, ls, kt) in enumerate(sets):
                knew = None if gp._has_derivs else self._knew(x0, x1, 0, 0, scale, ls)
                k1 = {t: self._cross_rows(x0, t, ns, scale, ls, kt) for t in set(task0.tolist())}
                sol2 = {t: self.gram_matrix_solve(k1[t] if (equal and t in k1) else self._cross_rows(x1, t, ns, scale, ls, kt), n=n, A=A[h]) for t in set(task1.tolist())}
                for i0, t0 in enumerate(task0.tolist()):
                    for i1, t1 in enumerate(task1.tolist()):
                        out[h, i0, i1] = kt[t0, t1] * (self._knew(x0, x1, t0, t1, scale, ls) if knew is None else knew) - k1[t0] @ sol2[t1].T
                        if equal and t0 == t1 and i0 == i1:
                            out[h, i0, i1].diagonal().clamp_(min=0)
        return out.reshape(Bh + out.shape[1:])

    def post_cubature_mean(self, task):
        """(*shape_batch, len(task)): K_task[task, :] (scale sum_i coeffs_l,i)_l."""
        gp = self.gp
        ns = self.sizes()[0]
        with torch.no_grad():
            sums = gp.scale * torch.stack([c.sum(-1) for c in self.coeffs().split(ns, dim=-1)], -1)  # (*shape_batch, T)
            return torch.einsum("...it,...t->...i", gp.gram_matrix_tasks[..., task.to(gp.device), :], sums)

    def post_cubature_cov(self, task0, task1, n):
        """abstract_fast_gp.py:110-154: scale K_task - scale^2 K_task (sqrt(n_i n_j) A_0[first rows]) K_task, (*Bh, len(task0), len(task1))."""
        gp = self.gp
        with torch.no_grad():
            A, _ = self.factor(n)
            ns, active, nmin, r, off, R = self.sizes(n)
            Bh = self.hyper_shape()
            ktd = gp.gram_matrix_tasks.expand(Bh + (self.T, self.T)).to(A.dtype)
            scale = gp.scale.expand(Bh + (1,))[..., None]
            idx = torch.tensor([off[l] for l in active], device=gp.device)
            act = torch.tensor(active, device=gp.device)
            nv = torch.tensor([float(ns[l]) for l in active], device=gp.device)
            mid = torch.sqrt(nv[:, None] * nv[None, :]) * A[..., 0, :, :][..., idx, :][..., :, idx]
            t0, t1 = task0.to(gp.device), task1.to(gp.device)
            term = (ktd[..., t0, :][..., :, act] @ mid @ ktd[..., act, :][..., :, t1]).real
            return scale * ktd.real[..., t0, :][..., :, t1] - scale ** 2 * term


class MultiTaskInverseLogDetCache(object):
    """Strategy object of `get_inv_log_det_cache` for several tasks (util.py:275-394)."""

    def __init__(self, gp, n):
        self.fgp = gp
        self.n = n
        self.nvec = [int(v) for v in n.tolist()]
        self.nint = max(self.nvec)
        self.task_order = torch.arange(gp.num_tasks, device=gp.device)
        self.inv_task_order = torch.arange(gp.num_tasks, device=gp.device)

    def __call__(self):
        A, logdet = self.fgp._mt.factor(self.nvec)
        return A.movedim(-3, -1), logdet  # (..., R, R, n_min) as the reference lays it out

    def gram_matrix_solve(self, y):
        return self.fgp._mt.gram_matrix_solve(y, n=self.nvec)

    def get_norm_term_logdet_term(self):
        norm, logdet, _, _ = self.fgp._mt.norm_logdet(grad=torch.is_grad_enabled())
        return norm, logdet

    def get_gcv_numer_denom(self):
        _, numer, denom, _ = self.fgp._mt.loss("GCV", 1, 0.0)
        return numer, denom
