"""Host-side mirror of the reference's fast GP classes for the structured-covariance hot path.

Reference surface being mirrored (same names, argument meaning, shapes and assertion behaviour):
    fastgps/abstract_gp.py      AbstractGP      (ctor :13-150, fit :152-306, get_x_next/add_y_next :310-351,
                                                  post_mean :352-380, post_var :381-416, post_cov :417-474,
                                                  post_error/post_ci :475-526, properties :610-706)
    fastgps/abstract_fast_gp.py AbstractFastGP  (ctor :12-31, power-of-two guards :32-52, default optimizer :53-57,
                                                  get_inv_log_det_cache :58-64, cubature :65-154, get_* :155-172,
                                                  _kernel* :173-196, ft/ift :197-228)
    fastgps/util.py             _XXbSeq :16-38, _K1PartsSeq :40-62, _LamCaches :64-141, _YtildeCache :164-183,
                                _FastInverseLogDetCache :275-394 (single-task branch), _CoeffsCache :396-425

Everything numerical is done by libfgp_b200.so (include/fgp_b200.h) on CUDA tensors; torch is used for memory,
streams, the tiny hyperparameter transforms and (outside the fast path) the optimizer.  There is no CPU fallback.

Scope: the fused device-side path covers one task without derivative information (rows a1-a15); several
tasks (equal or different power-of-two sizes), derivative observations, GCV / CV losses and masked fits run on the same CUDA
transforms through torch.autograd (multitask.py, _FTFunction) -- rows (f)2-(f)4, with batched outputs (shape_batch), one shared or
one hyperparameter set per batch element, and the adaptive nugget (util.py:286-290; the identity for a single task).
"""
import functools
import math
import os
import weakref
from fractions import Fraction
from typing import List, Tuple, Union

import numpy as np
import scipy.stats
import torch

from . import _lib
from . import multitask
from . import sequences


def _tf_log(x):
    return torch.log(x)


def _tf_exp(x):
    return torch.exp(x)


def _tf_id(x):
    return x


DEFAULT_TFS_LOG_EXP = (_tf_log, _tf_exp)
DEFAULT_TFS_ID = (_tf_id, _tf_id)


def _prod(shape):
    out = 1
    for s in shape:
        out *= int(s)
    return out


def _bshape(*shapes):
    """torch.broadcast_shapes for plain shape tuples (that function goes through several Python layers: ~20 us per call in fit()'s set-up)."""
    nd = max(len(sh) for sh in shapes)
    out = [1] * nd
    for sh in shapes:
        for k in range(1, len(sh) + 1):
            v = int(sh[-k])
            if v != 1:
                if out[-k] != 1 and out[-k] != v:
                    raise RuntimeError("shapes %s do not broadcast" % (shapes,))
                out[-k] = v
    return torch.Size(out)


class _XXbSeq(object):
    """Growing cache of the points of one sequence, generated on the GPU (util.py:16-38)."""

    def __init__(self, fgp, seq):
        # a weak back-reference: the GP keeps these objects in a numpy object array (as the reference does), which the cycle collector
        # does not traverse -- a strong one would make every GP object, its device data and its armed fit context immortal
        self.fgp = weakref.proxy(fgp)
        self.seq = seq
        self.n = 0
        self.x = torch.empty((0, seq.d), device=fgp.device)
        self.xb = torch.empty((0, seq.d), dtype=fgp._XBDTYPE, device=fgp.device)

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            i = slice(None, int(i), None)
        if isinstance(i, torch.Tensor):
            assert i.numel() == 1
            i = slice(None, int(i.item()), None)
        assert isinstance(i, slice)
        stop = int(i.stop)
        if stop > self.n:
            x_next, xb_next = self.fgp._sample(self.seq, self.n, stop)
            if x_next.data_ptr() == xb_next.data_ptr():
                self.x = self.xb = torch.vstack([self.x, x_next])
            else:
                self.x = torch.vstack([self.x, x_next])
                self.xb = torch.vstack([self.xb, xb_next])
            self.n = stop
        i = slice(None if i.start is None else int(i.start), stop, None)
        return self.x[i], self.xb[i]


class _FTFunction(torch.autograd.Function):
    """The unitary transforms of libfgp_b200 with autograd (kind 0: FFT-BRO, 1: inverse FFT-BRO, 2: FWHT): the backward of
    y = T x is the adjoint T^H g, i.e. the inverse transform (real part for a real input).  This is the differentiable
    seam the reference injects at fast_gp_lattice.py:224-225 / fast_gp_digital_net_b2.py:226; it carries the GCV / CV
    losses and masked fits, whose gradients are not in the fused MLL kernel."""

    @staticmethod
    def forward(ctx, x, kind):
        ctx.kind = kind
        ctx.real_in = not x.is_complex()
        if kind == 0:
            return _lib.fftbr(x)
        if kind == 1:
            return _lib.ifftbr(x)
        return _lib.fwht(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        if ctx.kind == 2:
            return _lib.fwht(g), None
        gi = _lib.ifftbr(g) if ctx.kind == 0 else _lib.fftbr(g)
        return (gi.real if ctx.real_in else gi), None


class _MLLFunction(torch.autograd.Function):
    """loss_b = wn_b*norm_b + wl_b*logdet_b of B hyperparameter sets with the analytic gradient of the fused kernel
    (replaces autograd through util.py:285-300,354-370 and the transform)."""

    @staticmethod
    def forward(ctx, scale_B, ls_B, noise_B, fgp, n, ysq, weights, want_grad):
        xpts = fgp._xpts(n)
        out, _ = _lib.mll_grad(fgp._FAMILY, xpts, fgp._alpha_list, fgp._t, ysq, scale_B.detach().contiguous(),
                               ls_B.detach().contiguous(), noise_B.detach().contiguous(), want_grad=want_grad,
                               weights=weights, z=fgp._zgen, C=fgp._Cgen(n))
        ctx.save_for_backward(out)
        ctx.d = fgp.d
        ctx.have_grad = want_grad
        lossb = weights[:, 0] * out[:, 0] + weights[:, 1] * out[:, 1]
        norm, logdet = out[:, 0].clone(), out[:, 1].clone()
        ctx.mark_non_differentiable(norm, logdet)
        return lossb, norm, logdet

    @staticmethod
    def backward(ctx, g, _gn, _gl):
        (out,) = ctx.saved_tensors
        assert ctx.have_grad, "gradients were not requested in the forward pass"
        d = ctx.d
        return g * out[:, 3], g[:, None] * out[:, 4:4 + d], g * out[:, 2], None, None, None, None, None


# Pinned host buffers for polling the device-side fit state.  cudaHostAlloc costs ~1 ms, more than 10 fit iterations at
# n = 2^20, so the buffers are pooled per process instead of being allocated per fit loop.
_PIN_POOL = []


def _pin_acquire():
    return _PIN_POOL.pop() if _PIN_POOL else torch.zeros((3, 32), dtype=torch.float64).pin_memory()


def _pin_release(buf):
    if buf is not None and len(_PIN_POOL) < 64:
        _PIN_POOL.append(buf)


class _FitContext(object):
    """Everything a device-side fit loop touches on the GPU, owned by a process-wide pool and keyed on the problem's shape:
    raw-parameter staging, effective hyperparameters, |ytilde|^2, workspace, state block, history rows -- and the CUDA graphs
    captured over exactly these buffers.  A fit() copies its parameters and |ytilde|^2 in (a few KB + 8n bytes, device to device),
    runs, copies the fitted parameters out and hands the context back, so the second and every later fit of a given shape in a
    process pays neither allocations nor graph capture (measured: ~1.2 ms of a 2.7 ms fit(iterations=20) at n = 2^20)."""
    POOL = {}
    ST_HEADER = 32

    def __init__(self, key, fgp, shapes, req, B, n, tau, d_out, hist_flags, cap):
        dev = fgp.device
        self.key = key
        self.B, self.d, self.n, self.tau, self.d_out = B, fgp.d, n, tau, d_out
        self.family = fgp._FAMILY
        f64 = dict(dtype=torch.float64, device=dev)
        self.raw = tuple(torch.zeros(tuple(sh), **f64) for sh in shapes)
        self.scale_B = torch.zeros(B, **f64)
        self.ls_B = torch.zeros((B, fgp.d), **f64)
        self.noise_B = torch.zeros(B, **f64)
        self.ysq = torch.zeros((B, n), **f64)
        self.out = torch.zeros((B, fgp.d + 4), **f64)
        self.ws = _lib.mll_workspace(fgp._FAMILY, n, fgp.d, B, dev)
        # filled on the device: torch.tensor(list, device=cuda) is a pageable H2D copy that makes the host wait for the stream
        self.weights = torch.empty((B, 2), **f64)
        self.weights[:, 0] = 0.5
        self.weights[:, 1] = 0.5 * d_out / B
        self.P = sum(r.numel() for r in self.raw)
        self.state = torch.zeros(_lib.fit_state_doubles(self.P, B), **f64)
        self.pin = _pin_acquire()
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.hist_flags, self.hist_capacity = hist_flags, cap
        self.loss_hist = torch.zeros((cap, 3), **f64)
        mk = lambda flag, r: torch.zeros((cap, r.numel()), **f64) if flag else None
        self.scale_hist, self.ls_hist, self.noise_hist = (mk(f, r) for f, r in zip(hist_flags, self.raw))
        L = _lib.FitLayout()
        L.B, L.d = B, fgp.d
        L.n_scale = _prod(shapes[0][:-1])
        L.n_ls_b, L.n_ls_d = _prod(shapes[1][:-1]), int(shapes[1][-1])
        L.n_noise = _prod(shapes[2][:-1])
        L.req_scale, L.req_ls, L.req_noise = (int(r) for r in req)
        L.tau = tau
        L.raw_scale, L.raw_ls, L.raw_noise = (r.data_ptr() for r in self.raw)
        L.scale_B, L.ls_B, L.noise_B = self.scale_B.data_ptr(), self.ls_B.data_ptr(), self.noise_B.data_ptr()
        L.state = self.state.data_ptr()
        L.loss_hist = self.loss_hist.data_ptr()
        L.scale_hist = None if self.scale_hist is None else self.scale_hist.data_ptr()
        L.ls_hist = None if self.ls_hist is None else self.ls_hist.data_ptr()
        L.noise_hist = None if self.noise_hist is None else self.noise_hist.data_ptr()
        self.layout = L
        # the whole iteration as one C call (fgp_fit_iteration); ctypes arrays and device tensors are kept alive on self
        self._alpha_arr = (_lib._i32 * fgp.d)(*fgp._alpha_list)
        self._z_arr = (_lib._u64 * fgp.d)(*fgp._zgen) if fgp._zgen is not None else None
        self.xpts = fgp._xpts(n).contiguous()
        self._C = fgp._Cgen(n)
        Pb = _lib.FitProblem()
        Pb.family = fgp._FAMILY
        Pb.x_dev = self.xpts.data_ptr()
        Pb.z_host = _lib._c.cast(self._z_arr, _lib._vp) if self._z_arr is not None else None
        Pb.C_dev = self._C.data_ptr() if self._C is not None else None
        Pb.mmax = int(self._C.shape[1]) if self._C is not None else 0
        Pb.n, Pb.d, Pb.t = n, fgp.d, int(fgp._t)
        Pb.alpha_host = _lib._c.cast(self._alpha_arr, _lib._vp)
        Pb.ysq_dev = self.ysq.data_ptr()
        Pb.weights_dev = self.weights.data_ptr()
        Pb.table_dev = _lib.fft_table(n, dev).data_ptr() if fgp._FAMILY == 0 else None
        Pb.workspace_dev = self.ws.data_ptr()
        Pb.out_dev = self.out.data_ptr()
        self.problem = Pb
        self.generator = self._z_arr is not None or self._C is not None
        self.graphs = {}
        self.streams = {}
        self.warmed = False
        self.kernels_per_iteration = None

    @classmethod
    def acquire(cls, fgp, pshape, hist_flags, hist_capacity):
        with torch.no_grad():
            tau = fgp._tau_host()
        raws = (fgp.raw_scale, fgp.raw_lengthscales, fgp.raw_noise)
        shapes = tuple(tuple(r.shape) for r in raws)
        req = tuple(bool(r.requires_grad) for r in raws)
        B, n = _prod(pshape), fgp._nint
        d_out = _prod(fgp.shape_batch)
        cap = 64
        while cap < max(int(hist_capacity), 1):
            cap *= 2
        hist_flags = tuple(bool(f) for f in hist_flags)
        # what the captured kernels read besides the pooled buffers: the generator inputs, or -- without them -- the stored points
        if fgp._zgen is not None:
            gen = ("z",) + tuple(fgp._zgen)
        elif fgp._Cgen(n) is not None:
            gen = ("C", fgp._Cgen(n).data_ptr())
        else:
            gen = ("x", fgp._xpts(n).data_ptr())
        key = (str(fgp.device), fgp._FAMILY, n, fgp.d, B, tuple(fgp._alpha_list), int(fgp._t), gen, shapes, req, tau, d_out, hist_flags, cap)
        free = cls.POOL.setdefault(key, [])
        return free.pop() if free else cls(key, fgp, shapes, req, B, n, tau, d_out, hist_flags, cap)

    def release(self):
        free = self.POOL.setdefault(self.key, [])
        if len(free) < 4:
            free.append(self)
        else:
            _pin_release(self.pin)
            self.pin = None


class _FusedFitLoop(object):
    """Device-side fit() loop (the product's fast path): every iteration is ONE C call (fgp_fit_iteration: fused eigen-solve whose
    last CTA reduces the partial sums and runs Rprop and the early-stop state machine), replayed from CUDA graphs held by a pooled
    `_FitContext`; the host only polls the `stopped` flag between chunks (include/fgp_b200.h, K4/K4b).  FGP_COOP=1 opts into the
    persistent cooperative kernel instead (a chunk of iterations per launch).
    Used when the loss is MLL, the optimiser is the default Rprop and the transforms are the default (log, exp)."""
    GRAPH_ITERS = 16
    ST_STOPPED, ST_LAST_ITER, ST_HEADER = 4, 5, 32

    @staticmethod
    def eligible(fgp):
        if not fgp._default_tfs:
            return False
        if fgp.raw_factor_task_kernel.requires_grad or fgp.raw_noise_task_kernel.requires_grad:
            return False
        if fgp.raw_factor_task_kernel.numel() != 0 or fgp.raw_noise_task_kernel.numel() != 1:
            return False
        B = _prod(_bshape(fgp.raw_scale.shape[:-1], fgp.raw_lengthscales.shape[:-1], fgp.raw_noise.shape[:-1]))
        ok = lambda p: _prod(p.shape[:-1]) in (1, B) and p.is_contiguous() and p.dtype == torch.float64
        return ok(fgp.raw_scale) and ok(fgp.raw_lengthscales) and ok(fgp.raw_noise) and B <= 65535

    def __init__(self, fgp, hist_flags=(False, False, False), hist_capacity=0):
        self.fgp = fgp
        pshape = fgp._pshape()
        self.pshape = pshape
        ysq = fgp._get_ysq(pshape)
        c = self.ctx = _FitContext.acquire(fgp, pshape, hist_flags, hist_capacity)
        self.B, self.d, self.n, self.d_out = c.B, c.d, c.n, c.d_out
        # |ytilde|^2 in: a device-to-device copy into the buffer the captured kernels read (the parameters follow in begin())
        with torch.no_grad():
            c.ysq.copy_(ysq)
        self.state, self.layout, self.problem = c.state, c.layout, c.problem
        self.loss_hist, self.scale_hist, self.ls_hist, self.noise_hist = c.loss_hist, c.scale_hist, c.ls_hist, c.noise_hist
        self.hist_capacity = c.hist_capacity
        self.state_host, self.state_host2, self.events = c.pin[2], c.pin[:2], c.events
        self.launches = 0
        self.replayed = 0
        # the Stream object snapshot events are recorded on and close() waits for: kept by the pooled context per raw stream handle
        # (torch.cuda.current_stream() alone is ~20 us of host time)
        with _lib.on_device(fgp.device):
            handle = _lib._stream()
        self.stream = c.streams.get(handle)
        if self.stream is None:
            self.stream = c.streams[handle] = torch.cuda.current_stream(fgp.device)
        # persistent cooperative kernel (opt-in): k iterations per launch, no graph
        self.multi = _lib.fit_iterations_per_launch(fgp._FAMILY, self.n) > 1 and os.environ.get("FGP_B200_NO_MULTI") != "1"
        self.use_graph = not self.multi and os.environ.get("FGP_B200_NO_GRAPH") != "1"
        if self.multi:
            c.kernels_per_iteration = 1
        elif not c.warmed:
            # eager warm-up of every kernel before the first capture (a capture must not meet lazy module loading); `stopped`
            # is raised so that the fit step changes nothing (the state block is zero: tickets start at 0 as fit_init leaves them)
            with torch.cuda.device(fgp.device):
                c.state[self.ST_STOPPED] = 1.0
                c0 = _lib.launch_count()
                self._iteration()
                c.kernels_per_iteration = _lib.launch_count() - c0
                torch.cuda.synchronize(fgp.device)
            c.warmed = True
        self.kernels_per_iteration = c.kernels_per_iteration

    # algorithmic bytes of one iteration: points read by the first and last pass, workspace written and
    # read twice, |ytilde|^2 read once
    @property
    def algorithmic_bytes(self):
        e = 16 if self.fgp._FAMILY == 0 else 8
        pts = 0 if self.ctx.generator else 2 * 8 * self.n * self.d
        return self.B * (pts + 4 * e * self.n + 8 * self.n)

    def kernel_algorithmic_bytes(self, name):
        e = 16 if self.fgp._FAMILY == 0 else 8
        n, d, B = self.n, self.d, self.B
        if self.ctx.generator:
            d = 0
        if name == "mll_coop":
            return self.algorithmic_bytes
        return {"mll_passA": B * (8 * n * d + e * n), "mll_passB": B * (2 * e * n + 8 * n), "mll_passC": B * (8 * n * d + e * n),
                "mll_single": B * (16 * n * d + 8 * n)}.get(name, 0)

    def _iteration(self):
        _lib.fit_iteration(self.problem, self.layout)

    def begin(self, iterations, stop_wait, logtol, lr):
        o = _lib.FitOptions()
        o.iterations, o.stop_wait, o.hist_capacity = int(iterations), int(stop_wait), int(self.hist_capacity)
        o.logtol = float(logtol)
        o.half_const = 0.5 * self.d_out * self.n * float(np.log(2 * np.pi))
        o.wn, o.wl = 0.5, 0.5 * self.d_out / self.B
        o.lr, o.etaminus, o.etaplus, o.step_min, o.step_max = float(lr), 0.5, 1.2, 1e-6, 50.0
        fgp = self.fgp
        with _lib.on_device(fgp.device):  # the GP's parameters -> the pooled staging buffers, inside the init launch
            _lib.fit_init_from(self.layout, o, fgp.raw_scale.data, fgp.raw_lengthscales.data, fgp.raw_noise.data)

    def _graph(self, k):
        graphs = self.ctx.graphs
        g = graphs.get(k)
        if g is None:
            # raw capture on a side stream: the torch.cuda.graph() context manager also runs gc.collect() and
            # empty_cache(), tens of milliseconds -- more than a whole short fit
            g = torch.cuda.CUDAGraph()
            with torch.cuda.device(self.fgp.device):
                cur = torch.cuda.current_stream()
                side = torch.cuda.Stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    g.capture_begin()
                    try:
                        for _ in range(k):
                            self._iteration()
                    finally:
                        g.capture_end()
                cur.wait_stream(side)
            graphs[k] = g
        return g

    def replay(self, k):
        """Enqueue k <= GRAPH_ITERS iterations."""
        if self.multi:
            with _lib.on_device(self.fgp.device):
                _lib.fit_iterations(self.problem, self.layout, k)
            self.replayed += k
            self.launches += 1
            return
        if not self.use_graph:
            with _lib.on_device(self.fgp.device):
                for _ in range(k):
                    self._iteration()
        else:
            # one graph per chunk length, captured on first use and kept by the pooled context (a fit of K iterations replays two graphs:
            # chunks of GRAPH_ITERS and the remainder)
            self._graph(k).replay()
        self.replayed += k
        self.launches += k * self.kernels_per_iteration

    def snapshot(self, slot):
        """Stream-ordered copy of the state header into pinned slot `slot` plus an event: lets the host look at chunk j
        while chunk j+1 is already enqueued (kernels of iterations after the stop decision exit immediately)."""
        self.state_host2[slot].copy_(self.state[:self.ST_HEADER], non_blocking=True)
        self.events[slot].record(self.stream)

    def wait_snapshot(self, slot):
        self.events[slot].synchronize()
        return self.state_host2[slot]

    def step(self):
        self.replay(1)

    def read_state(self):
        self.state_host.copy_(self.state[:self.ST_HEADER], non_blocking=True)
        torch.cuda.current_stream(self.fgp.device).synchronize()
        return self.state_host

    def check(self, st):
        if float(st[self.ST_STOPPED]) == 2.0:  # raised by a grid barrier of the persistent kernel that gave up waiting
            raise _lib.FgpError("libfgp_b200: the device-side fit loop reported a failed grid barrier")

    def finish(self):
        """Best iterate -> staged parameters (abstract_gp.py:297-298) -> the GP's own parameter storages."""
        fgp, c = self.fgp, self.ctx
        with _lib.on_device(fgp.device):
            _lib.fit_finish_to(self.layout, fgp.raw_scale.data, fgp.raw_lengthscales.data, fgp.raw_noise.data)
        fgp._epoch += 1

    def kernel_times(self, reps=10, flush=None):
        """Per-kernel device time of one iteration (eager launches, CUDA events between kernels; L2 flushed first)."""
        acc = {}
        with torch.cuda.device(self.fgp.device):
            for _ in range(reps):
                if flush is not None:
                    flush.zero_()
                _lib.profile_begin()
                self._iteration()
                for name, ms in _lib.profile_end():
                    acc.setdefault(name, []).append(ms)
        return [{"name": k, "ms": float(np.mean(v)), "alg_bytes": self.kernel_algorithmic_bytes(k)} for k, v in acc.items()]

    def close(self, finish=True):
        """Hand the context back to the pool.  finish: copy the best iterate into the GP first (an open-ended `fit_stepper` run ends
        here; fit() has already done it)."""
        if self.ctx is None:
            return
        if finish:
            self.finish()
        self.stream.synchronize()  # nothing may still be reading or writing the pooled buffers
        self.ctx.release()
        self.ctx = None

    def __del__(self):
        try:
            if getattr(self, "ctx", None) is not None:
                self.close(finish=False)
        except Exception:
            pass


class _FastInverseLogDetCache(object):
    """Strategy object handed out by `get_inv_log_det_cache` (util.py:275-394, single-task branch)."""

    def __init__(self, fgp, n):
        self.fgp = fgp
        self.n = n
        self.nint = int(n[0])
        self.task_order = torch.zeros(1, dtype=int, device=fgp.device)
        self.inv_task_order = torch.zeros(1, dtype=int, device=fgp.device)
        self._key = None

    def _lam_full(self):
        """(B,n) eigenvalues sqrt(n)*ft(k1)+noise (times the task kernel), cached on the hyperparameter state."""
        key = self.fgp._param_key()
        if self._key != key or os.environ.get("FASTGP_FORCE_RECOMPILE") == "True":
            with torch.no_grad():
                scale_B, ls_B, noise_B, pshape = self.fgp._hyper()
                B = scale_B.numel()
                ysq = torch.zeros((B, self.nint), dtype=torch.float64, device=self.fgp.device)
                _, lam = _lib.mll_grad(self.fgp._FAMILY, self.fgp._xpts(self.nint), self.fgp._alpha_list, self.fgp._t, ysq,
                                       scale_B.contiguous(), ls_B.contiguous(), noise_B.contiguous(), want_grad=False,
                                       want_lam=True, z=self.fgp._zgen, C=self.fgp._Cgen(self.nint))
            self.lam = lam
            self.pshape = pshape
            self._key = key
        return self.lam

    def __call__(self):
        lam = self._lam_full()
        lamp = lam.reshape(tuple(self.pshape) + (self.nint,))
        self.logdet = torch.log(torch.abs(lamp)).sum(-1)
        self.inv = (1 / lamp)[..., None, None, :]
        return self.inv, self.logdet

    def gram_matrix_solve(self, y):
        """K^-1 y along the last dim; batch dims of y broadcast against the hyperparameter batch shape."""
        assert y.size(-1) == self.nint
        lam = self._lam_full()
        B = lam.shape[0]
        y = y.to(self.fgp.device)
        if B == 1:
            return _lib.gram_solve(self.fgp._FAMILY, y.contiguous(), lam[0])
        k = len(self.pshape)
        assert tuple(y.shape[-1 - k:-1]) == tuple(self.pshape), "y batch dims must end with the hyperparameter batch shape %s" % (tuple(self.pshape),)
        lead = y.shape[:-1 - k]
        yb = y.reshape((_prod(lead), B, self.nint))
        out = torch.empty_like(yb)
        for b in range(B):
            out[:, b, :] = _lib.gram_solve(self.fgp._FAMILY, yb[:, b, :].contiguous(), lam[b])
        return out.reshape(y.shape)

    def get_norm_term_logdet_term(self):
        """(norm_term[...,1], logdet[...,1]); differentiable w.r.t. the raw hyperparameters when grad mode is on."""
        fgp = self.fgp
        assert self.nint == fgp._nint, "norm term needs the current data size"
        ytilde = fgp.get_ytilde(0)
        scale_B, ls_B, noise_B, pshape = fgp._hyper()
        B = scale_B.numel()
        k = len(pshape)
        sb = tuple(fgp.shape_batch)
        lead = _prod(sb[:len(sb) - k])
        want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in (scale_B, ls_B, noise_B))
        # one "hyperparameter set" per y column so the norm term keeps the reference's per-column shape
        ysq_cols = (torch.view_as_real(ytilde).pow(2).sum(-1) if ytilde.is_complex() else ytilde ** 2).reshape(lead * B, self.nint).contiguous()
        w_norm = torch.tensor([1.0, 0.0], device=fgp.device).expand(lead * B, 2).contiguous()
        rep = lambda v: v.expand((lead,) + tuple(v.shape)).reshape((lead * B,) + tuple(v.shape[1:]))
        norm_cols, _, _ = _MLLFunction.apply(rep(scale_B), rep(ls_B), rep(noise_B), fgp, self.nint, ysq_cols, w_norm, want_grad)
        w_ld = torch.tensor([0.0, 1.0], device=fgp.device).expand(B, 2).contiguous()
        ld, _, _ = _MLLFunction.apply(scale_B, ls_B, noise_B, fgp, self.nint, torch.zeros((B, self.nint), device=fgp.device), w_ld, want_grad)
        return norm_cols.reshape(sb + (1,)), ld.reshape(tuple(pshape) + (1,))

    def get_gcv_numer_denom(self):
        """util.py:371-380, differentiable through the autograd transform."""
        fgp = self.fgp
        assert self.nint == fgp._nint, "GCV needs the current data size"
        ytilde = fgp.get_ytilde(0)
        inv = 1 / fgp._lam_autograd(self.nint)
        ztilde = ytilde * inv
        numer = (ztilde.conj() * ztilde).real.sum(-1, keepdim=True)
        tr_k_inv = inv.real.sum(-1, keepdim=True)
        denom = ((tr_k_inv / self.nint) ** 2).real
        return numer, denom

    def get_inv_diag(self):
        """util.py:381-385 (single task): mean of 1 / (lam~ sqrt(n)), lam~ = ft(k1) without the noise, as in the reference."""
        _, lam_tilde = self.fgp._lam_autograd(self.nint, with_tilde=True)
        return (1 / (lam_tilde * np.sqrt(self.nint))).mean(-1, keepdim=True)


class AbstractFastGP(torch.nn.Module):
    _FAMILY = None  # 0 lattice, 1 digital net
    _DENSE = False  # StandardGP (standard_gp.py): same parameter handling and fit loop, dense torch algebra instead of the CUDA path
    _XBDTYPE = None
    _FTOUTDTYPE = None
    _DEFAULT_NOISE = None

    def __init__(self, seqs, num_tasks, seed_for_seq, alpha, scale, lengthscales, noise, factor_task_kernel,
                 rank_factor_task_kernel, noise_task_kernel, device, tfs_scale, tfs_lengthscales, tfs_noise,
                 tfs_factor_task_kernel, tfs_noise_task_kernel, requires_grad_scale, requires_grad_lengthscales,
                 requires_grad_noise, requires_grad_factor_task_kernel, requires_grad_noise_task_kernel, shape_batch,
                 shape_scale, shape_lengthscales, shape_noise, shape_factor_task_kernel, shape_noise_task_kernel,
                 derivatives, derivatives_coeffs, compile_fts, compile_fts_kwargs, adaptive_nugget):
        super().__init__()
        assert torch.get_default_dtype() == torch.float64, "fast transforms do not work without torch.float64 precision"
        self.device = torch.device(device)
        if not self._DENSE:
            _lib.load()  # fail loudly when the CUDA library is missing: there is no CPU path
            if self.device.type != "cuda":
                raise RuntimeError("fastgaussianprocesses_b200 computes on a CUDA device only (got device=%r); there is no CPU fallback" % (device,))
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if num_tasks is None:
            solo_task, default_task, num_tasks = True, 0, 1
        else:
            assert isinstance(num_tasks, int) and num_tasks > 0
            solo_task, default_task = False, torch.arange(num_tasks)
        if derivatives is not None or derivatives_coeffs is not None:  # abstract_gp.py:59-62
            rank_factor_task_kernel = 1
            tfs_noise_task_kernel = DEFAULT_TFS_ID
            noise_task_kernel = 0.
        self.num_tasks = num_tasks
        self.default_task = default_task
        self.solo_task = solo_task
        # sequences
        if isinstance(seqs, (int, np.integer)):
            seqs = np.array([self._default_sequence(int(seqs), seed) for seed in np.random.SeedSequence(seed_for_seq).spawn(num_tasks)], dtype=object)
        elif isinstance(seqs, (list, tuple)):
            seqs = np.array(list(seqs), dtype=object)
        elif not isinstance(seqs, np.ndarray):
            seqs = np.array([seqs], dtype=object)
        assert seqs.shape == (num_tasks,), "seqs should be a length num_tasks=%d list" % num_tasks
        seqs = np.array([self._adopt_sequence(s) for s in seqs], dtype=object)
        assert self._DENSE or all(seqs[i].order == "NATURAL" for i in range(num_tasks)), "each seq should be in 'NATURAL' order "
        assert all(seqs[i].replications == 1 for i in range(num_tasks)), "each seq should have only 1 replication"
        self.d = seqs[0].d
        assert self._DENSE or self.d <= _lib.MAX_D, "dimension %d exceeds the fused-kernel limit %d" % (self.d, _lib.MAX_D)
        self.seqs = seqs
        self.n = torch.zeros(self.num_tasks, dtype=int, device=self.device)
        self.m = -1 * torch.ones(self.num_tasks, dtype=int, device=self.device)
        self._nint = 0
        # derivatives (abstract_gp.py:63-72): per task a (p,d) table of derivative multi-indices and p coefficients
        if derivatives is None:
            derivatives = [torch.zeros((1, self.d), dtype=torch.int64, device=self.device) for _ in range(self.num_tasks)]
        if isinstance(derivatives, torch.Tensor):
            derivatives = [derivatives]
        assert isinstance(derivatives, list) and len(derivatives) == self.num_tasks
        derivatives = [(deriv[None, :] if deriv.ndim == 1 else deriv).to(self.device) for deriv in derivatives]
        assert all((derivatives[i].ndim == 2 and derivatives[i].size(1) == self.d) for i in range(self.num_tasks))
        self.derivatives = derivatives
        if derivatives_coeffs is None:
            derivatives_coeffs = [torch.ones(len(self.derivatives[i]), device=self.device) for i in range(self.num_tasks)]
        assert isinstance(derivatives_coeffs, list) and len(derivatives_coeffs) == self.num_tasks
        derivatives_coeffs = [c.to(self.device) for c in derivatives_coeffs]
        assert all((derivatives_coeffs[i].ndim == 1 and len(derivatives_coeffs[i])) == len(self.derivatives[i]) for i in range(self.num_tasks))
        self.derivatives_coeffs = derivatives_coeffs
        self._has_derivs = any((self.derivatives[i] > 0).any() or (self.derivatives_coeffs[i] != 1).any() or len(self.derivatives[i]) != 1 for i in range(self.num_tasks))
        self._deriv_cache = {}
        # alpha
        assert (np.isscalar(alpha) and alpha % 1 == 0) or (isinstance(alpha, torch.Tensor) and alpha.shape == (self.d,)), "alpha should be an int or a torch.Tensor of length d"
        if np.isscalar(alpha):
            alpha = int(alpha) * torch.ones(self.d, dtype=int, device=self.device)
        self.alpha = alpha.to(self.device)
        self._alpha_list = [int(a) for a in self.alpha.tolist()]
        # shape_batch
        if isinstance(shape_batch, (list, tuple)):
            shape_batch = torch.Size(shape_batch)
        assert isinstance(shape_batch, torch.Size)
        self.shape_batch = shape_batch
        self.ndim_batch = len(self.shape_batch)
        # hyperparameters (abstract_gp.py:77-139): one raw (transformed) torch Parameter each, the trailing dimensions fixed by the
        # parameter, any leading ones matching the tail of shape_batch. Task kernel: F F^T + diag(v), F of rank 0 for a single task
        if shape_factor_task_kernel is None and not isinstance(factor_task_kernel, torch.Tensor):
            if rank_factor_task_kernel is None:
                rank_factor_task_kernel = 0 if self.num_tasks == 1 else 1
            assert isinstance(rank_factor_task_kernel, int) and 0 <= rank_factor_task_kernel <= self.num_tasks
            shape_factor_task_kernel = [self.num_tasks, rank_factor_task_kernel]
        T = self.num_tasks
        tfs_hint = " should be a tuple of two callables, the transform and inverse transform"
        for name, value, shape, tfs, rgrad, trailing, lower in (
                ("scale", scale, shape_scale, tfs_scale, requires_grad_scale, lambda s: s[-1] == 1, ">"),
                ("lengthscales", lengthscales, shape_lengthscales or [self.d], tfs_lengthscales, requires_grad_lengthscales, lambda s: s[-1] in (self.d, 1), ">"),
                ("noise", noise, shape_noise, tfs_noise, requires_grad_noise, lambda s: s[-1] == 1, ">"),
                ("factor_task_kernel", factor_task_kernel, shape_factor_task_kernel, tfs_factor_task_kernel, requires_grad_factor_task_kernel,
                 lambda s: 0 <= s[-1] <= T and s[-2] == T, None),
                ("noise_task_kernel", noise_task_kernel, shape_noise_task_kernel or [T], tfs_noise_task_kernel, requires_grad_noise_task_kernel,
                 lambda s: s[-1] in (T, 1), ">=")):
            self._add_hyperparameter(name, value, shape, tfs, (T > 1) if rgrad is None else rgrad, trailing, 2 if name == "factor_task_kernel" else 1,
                                     lower, tfs_hint)
        self._default_tfs = (tuple(tfs_scale) == DEFAULT_TFS_LOG_EXP and tuple(tfs_lengthscales) == DEFAULT_TFS_LOG_EXP and
                             tuple(tfs_noise) == DEFAULT_TFS_LOG_EXP)
        # storage and caches
        self._y = [torch.empty(0, device=self.device) for _ in range(self.num_tasks)]
        self.xxb_seqs = np.array([_XXbSeq(self, self.seqs[i]) for i in range(self.num_tasks)], dtype=object)
        self.inv_log_det_cache_dict = {}
        self.adaptive_nugget = bool(adaptive_nugget)
        # generator mode of the fused eigen-solve: a lattice spec of this package carries its generating vector, so the
        # first kernel column is regenerated from the point index and the points are never read (include/fgp_b200.h)
        s0 = self.seqs[0]
        self._zgen = [int(v) for v in s0.gen_vec] if (self._FAMILY == 0 and isinstance(s0, sequences.Lattice) and os.environ.get("FGP_B200_NO_GEN") != "1") else None
        self._netgen = self._FAMILY == 1 and isinstance(s0, sequences.DigitalNetB2) and os.environ.get("FGP_B200_NO_GEN") != "1"
        # several tasks, or derivative observations (their kernels are sums over derivative terms): block eigen-solve route
        self._mt = multitask.MultiTaskEngine(self) if ((self.num_tasks > 1 or self._has_derivs) and not self._DENSE) else None
        if self._has_derivs:  # abstract_gp.py:147-150
            self.raw_noise_task_kernel.requires_grad_(False)
            self.raw_factor_task_kernel.requires_grad_(False)
            assert (self.gram_matrix_tasks == 1).all()
        self._epoch = 0
        self._coeffs = None
        self._coeffs_key = None
        self._ytilde = None
        self._ytilde_n = -1
        self._ysq = None
        self._data_epoch = 0        # bumped by add_y_next
        self._fused_loop = None     # the device-side fit loop while a fit / fit_stepper is open
        self._prearmed_loop = None  # the loop add_y_next armed for the next fit (fast_gp.py:_prearm_fit)
        if self.num_tasks == 1:
            self._tau_host()
        # the injected transforms of the reference (abstract_fast_gp.py:26-27)
        self.ft_unstable = self._ft_unstable
        self.ift_unstable = self._ift_unstable

    def __setattr__(self, name, value):
        # bookkeeping attributes (underscore names, sizes) skip nn.Module's parameter / buffer / submodule checks: ~5 us each, a dozen per
        # add_y_next + fit; parameters and anything else still go through nn.Module
        if name[0] == "_" or name in ("n", "m"):
            object.__setattr__(self, name, value)
        else:
            super().__setattr__(name, value)

    # ------------------------------------------------------------------------------------------------ state keys
    def _add_hyperparameter(self, name, value, shape, tfs, requires_grad, trailing_ok, ntrail, sign, tfs_hint):
        """Validate one hyperparameter of abstract_gp.py:77-139 and register tfs[0](value) as self.raw_<name> (self.tf_<name> = tfs[1]).
        A tensor value fixes the shape; a scalar is broadcast to `shape`, whose last `ntrail` dimensions must pass `trailing_ok` and whose
        leading dimensions must equal the tail of shape_batch; `sign` is ">" / ">=" / None for the positivity assertion."""
        assert np.isscalar(value) or isinstance(value, torch.Tensor), name + " must be a scalar or torch.Tensor"
        if isinstance(value, torch.Tensor):
            shape = value.shape
        if isinstance(shape, (list, tuple)):
            shape = torch.Size(shape)
        assert isinstance(shape, torch.Size) and len(shape) >= ntrail and trailing_ok(shape)
        lead = len(shape) - ntrail
        assert lead == 0 or shape[:lead] == self.shape_batch[-lead:]
        value = (value * torch.ones(shape, device=self.device) if np.isscalar(value) else value).to(self.device)
        if sign is not None:
            assert ((value > 0) if sign == ">" else (value >= 0)).all(), name + " must be positive"
        assert len(tfs) == 2 and callable(tfs[0]) and callable(tfs[1]), "tfs_" + name + tfs_hint
        setattr(self, "tf_" + name, tfs[1])
        setattr(self, "raw_" + name, torch.nn.Parameter(tfs[0](value), requires_grad=requires_grad))

    def _param_key(self):
        ps = (self.raw_scale, self.raw_lengthscales, self.raw_noise, self.raw_factor_task_kernel, self.raw_noise_task_kernel)
        return (self._epoch,) + tuple((id(p), p._version, p.data_ptr()) for p in ps)

    def _hyper(self):
        """Effective per-set hyperparameters (scale_B (B,), ls_B (B,d), noise_B (B,), batch shape), autograd-connected.
        The 1x1 task kernel multiplies lam after the noise is added (util.py:293,298), so it folds into scale and noise."""
        scale, ls, noise = self.scale, self.lengthscales, self.noise
        tau = self.gram_matrix_tasks[..., 0, 0]
        pshape = torch.broadcast_shapes(scale.shape[:-1], ls.shape[:-1], noise.shape[:-1], tau.shape)
        B = _prod(pshape)
        scale_B = (scale[..., 0] * tau).expand(pshape).reshape(B)
        noise_B = (noise[..., 0] * tau).expand(pshape).reshape(B)
        ls_B = ls.expand(tuple(pshape) + (self.d,)).reshape(B, self.d)
        return scale_B, ls_B, noise_B, pshape

    def _pshape(self):
        """Batch shape of the hyperparameter sets (what `_hyper` returns last), from the parameter shapes alone: no device work."""
        return _bshape(self.raw_scale.shape[:-1], self.raw_lengthscales.shape[:-1], self.raw_noise.shape[:-1],
                       self.raw_factor_task_kernel.shape[:-2], self.raw_noise_task_kernel.shape[:-1])

    def _tau_host(self):
        """K_task[0,0] of a single task as a host float, cached on the task-kernel parameters' identity and version: reading it
        from the device is a host-device synchronisation in the middle of fit()'s set-up."""
        ps = (self.raw_factor_task_kernel, self.raw_noise_task_kernel)
        key = tuple((p.data_ptr(), p._version, p.numel()) for p in ps)
        c = getattr(self, "_tau_cache", None)
        if c is None or c[0] != key:
            with torch.no_grad():
                c = (key, float(self.gram_matrix_tasks.reshape(-1)[0]))
            self._tau_cache = c
        return c[1]

    def _hyper_host(self):
        """Host copies of the effective hyperparameters, cached on the parameter state (one device-to-host synchronisation per
        change of the parameters instead of one per posterior / kernel call)."""
        key = self._param_key()
        c = getattr(self, "_hyper_host_cache", None)
        if c is None or c[0] != key:
            with torch.no_grad():
                scale_B, ls_B, noise_B, pshape = self._hyper()
                c = (key, (scale_B.cpu().numpy(), ls_B.cpu().numpy(), noise_B.cpu().numpy(), pshape))
            self._hyper_host_cache = c
        return c[1]

    def _check_unit_cube(self, x):
        """fast_gp_lattice.py:264-265: the lattice kernel takes points of the unit cube (its |x - z| form is only valid there)."""
        if self._FAMILY == 0 and x.numel():
            assert bool(((0 <= x) & (x <= 1)).all()), "x should have all elements in [0,1]"

    def _Cgen(self, n):
        """Device generating matrices for the net generator mode (None: read the stored points)."""
        if not self._netgen:
            return None
        s0 = self.seqs[0]
        if int(n) > (1 << int(s0.gen_mats.shape[1])):
            return None
        return s0.device_matrices(self.device)

    def _xpts(self, n):
        x, xb = self.xxb_seqs[0][:int(n)]
        return xb if self._FAMILY == 1 else x

    # ------------------------------------------------------------------------------------------------ data
    def _sample(self, seq, n_min, n_max):
        return seq.generate(int(n_min), int(n_max), self.device)

    def get_x_next(self, n: Union[int, torch.Tensor], task: Union[int, torch.Tensor] = None):
        n_og = n
        if isinstance(n, (int, np.integer)):
            n = torch.tensor([n], dtype=int, device=self.device)
        if isinstance(n, list):
            n = torch.tensor(n, dtype=int, device=self.device)
        assert isinstance(n, torch.Tensor) and torch.logical_or(n == 0, n & (n - 1) == 0).all(), "maximum sequence index must be a power of 2"
        if task is None:
            task = self.default_task
        inttask = isinstance(task, int)
        if inttask:
            task = torch.tensor([task], dtype=int)
        if isinstance(task, list):
            task = torch.tensor(task, dtype=int)
        assert isinstance(n, torch.Tensor) and isinstance(task, torch.Tensor) and n.ndim == task.ndim == 1 and len(n) == len(task)
        assert (n >= self.n[task]).all(), "maximum sequence index must be greater than the current number of samples"
        x_next = [self.xxb_seqs[int(l)][int(self.n[l]):int(n[i])][0] for i, l in enumerate(task)]
        return x_next[0] if inttask else x_next

    def add_y_next(self, y_next: Union[torch.Tensor, List], task: Union[int, torch.Tensor] = None):
        if isinstance(y_next, torch.Tensor):
            y_next = [y_next]
        if task is None:
            task = self.default_task
        if isinstance(task, int):
            tasks = [task]
        else:
            if isinstance(task, list):
                task = torch.tensor(task, dtype=int)
            assert isinstance(task, torch.Tensor) and task.ndim == 1
            tasks = [int(l) for l in task.tolist()]
        assert isinstance(y_next, list) and len(y_next) == len(tasks)
        assert all(y_next[i].shape[:-1] == self.shape_batch for i in range(len(y_next)))
        # sizes from the shapes alone, then the copies FIRST: everything after them on the host overlaps the transfer
        ncur = [int(self._y[i].size(-1)) for i in range(self.num_tasks)]
        for i, l in enumerate(tasks):
            ncur[l] += int(y_next[i].size(-1))
        assert all(nl == 0 or (nl & (nl - 1)) == 0 for nl in ncur), "total samples must be power of 2"
        for i, l in enumerate(tasks):
            yi = y_next[i]
            if yi.device.type == "cpu" and yi.is_pinned():
                # a PINNED host buffer is copied asynchronously (stream-ordered): the host goes on to enqueue the transforms below while the
                # copy is in flight.  The source is kept alive until then; as with any non_blocking copy it must not be rewritten before
                # the next synchronising call.
                self._pinned_src = yi
                yi = yi.to(self.device, non_blocking=True)
            fresh = y_next[i].device != self.device  # our own device copy: no need to copy it again when it is the first block
            yi = yi.to(self.device)
            self._y[l] = yi if (fresh and self._y[l].numel() == 0 and yi.dtype == self._y[l].dtype) else torch.cat([self._y[l], yi], -1)
        self._nint = max(ncur)
        mcur = [-1 if nl == 0 else nl.bit_length() - 1 for nl in ncur]
        if self.num_tasks == 1:  # device fills: torch.tensor(list, device=cuda) is a pageable copy the host waits for
            self.n = torch.full((1,), ncur[0], dtype=int, device=self.device)
            self.m = torch.full((1,), mcur[0], dtype=int, device=self.device)
        else:
            self.n = torch.tensor(ncur, dtype=int, device=self.device)
            self.m = torch.tensor(mcur, dtype=int, device=self.device)
        for key in list(self.inv_log_det_cache_dict.keys()):
            if any(k < c for k, c in zip(key, ncur)):
                del self.inv_log_det_cache_dict[key]
        self._epoch += 1
        # single-task fast path: enqueue ytilde = ft(y) and |ytilde|^2 right away.  They are what fit() and coeffs need first, and here
        # their host-side launch work overlaps the (asynchronous) host-to-device copy of y instead of delaying the first fit iteration.
        self._data_epoch = getattr(self, "_data_epoch", 0) + 1
        if self._mt is None and not self._DENSE and self._nint > 0 and os.environ.get("FGP_B200_NO_PREFETCH") != "1":
            with torch.no_grad():
                self._get_ysq(self._pshape())
                if self._nint > 1 and os.environ.get("FGP_B200_NO_PREARM") != "1":
                    self._prearm_fit()

    # ------------------------------------------------------------------------------------------------ properties
    @property
    def total_parameters(self):
        return sum(p.numel() for p in self.parameters())

    @property
    def total_tuneable_parameters(self):
        return sum((p.numel() if p.requires_grad else 0) for p in self.parameters())

    @property
    def scale(self):
        return self.tf_scale(self.raw_scale)

    @property
    def lengthscales(self):
        return self.tf_lengthscales(self.raw_lengthscales)

    @property
    def noise(self):
        return self.tf_noise(self.raw_noise)

    @property
    def factor_task_kernel(self):
        return self.tf_factor_task_kernel(self.raw_factor_task_kernel)

    @property
    def noise_task_kernel(self):
        return self.tf_noise_task_kernel(self.raw_noise_task_kernel)

    @property
    def gram_matrix_tasks(self):
        f = self.factor_task_kernel
        kmat = torch.einsum("...il,...kl->...ik", f, f)
        return kmat + self.noise_task_kernel[..., None] * torch.eye(self.num_tasks, device=self.device)

    @property
    def coeffs(self):
        r"""Coefficients $\mathsf{K}^{-1} \boldsymbol{y}$ (util.py:419-425)."""
        if self._mt is not None:
            return self._mt.coeffs()
        key = (self._nint,) + self._param_key()
        if self._coeffs is None or self._coeffs_key != key or os.environ.get("FASTGP_FORCE_RECOMPILE") == "True":
            with torch.no_grad():
                self._coeffs = self.get_inv_log_det_cache().gram_matrix_solve(self._y[0])
            self._coeffs_key = key
        return self._coeffs

    @property
    def x(self):
        xs = [self.get_x(l) for l in range(self.num_tasks)]
        return xs[0] if self.solo_task else xs

    @property
    def y(self):
        return self._y[0] if self.solo_task else self._y

    def get_x(self, task, n=None):
        assert 0 <= task < self.num_tasks
        if n is None:
            n = self._nint
        assert n >= 0
        return self.xxb_seqs[task][:int(n)][0]

    def get_xb(self, task, n=None):
        assert 0 <= task < self.num_tasks
        if n is None:
            n = self._nint
        assert n >= 0
        return self.xxb_seqs[task][:int(n)][1]

    # ------------------------------------------------------------------------------------------------ transforms
    def ft(self, x):
        """Mean-stabilised orthonormal fast transform along the last dim (abstract_fast_gp.py:197-212)."""
        x = x.to(self.device)
        xmean = x.mean(-1)
        y = self.ft_unstable(x - xmean[..., None])
        if y.requires_grad:
            y = y.clone()  # autograd forbids in-place edits of a view of a custom Function's output
        y[..., 0] += xmean * np.sqrt(x.size(-1))
        return y

    def ift(self, x):
        """Mean-stabilised orthonormal inverse fast transform along the last dim (abstract_fast_gp.py:213-228)."""
        x = x.to(self.device)
        xmean = x.mean(-1)
        y = self.ift_unstable(x - xmean[..., None])
        if y.requires_grad:
            y = y.clone()
        y[..., 0] += xmean * np.sqrt(x.size(-1))
        return y

    def get_ytilde(self, task):
        assert 0 <= task < self.num_tasks
        if self._ytilde is None or self._ytilde_n != self._nint:
            y = self._y[task]
            self._ytilde = self.ft(y) if self._nint > 1 else y.clone().to(self._FTOUTDTYPE)
            self._ytilde_n = self._nint
            self._ysq = None
        return self._ytilde

    def _get_ysq(self, pshape):
        """(B,n) sums of |ytilde|^2 over the leading batch dims that share one hyperparameter set."""
        B = _prod(pshape)
        n = self._nint
        if ((self._ytilde is None or self._ytilde_n != n) and n > 1 and self._mt is None and self._y[0].dtype == torch.float64
                and B <= 65535 and _prod(self.shape_batch) <= 65535 and os.environ.get("FGP_B200_NO_SPECTRUM_CALL") != "1"):
            # nothing cached yet: ytilde and |ytilde|^2 from ONE C call (fgp_data_spectrum) instead of nine torch launches
            yt, ysq = _lib.data_spectrum(self._FAMILY, self._y[0].reshape(-1, n).contiguous(), B)
            self._ytilde, self._ytilde_n, self._ysq = yt.reshape(tuple(self.shape_batch) + (n,)), n, ysq
            return ysq
        ytilde = self.get_ytilde(0)
        if self._ysq is None or self._ysq.shape[0] != B:
            sb = tuple(self.shape_batch)
            lead = _prod(sb[:len(sb) - len(pshape)])
            # |z|^2 from the (re, im) view: torch's complex abs() is a jiterator kernel (NVRTC compile on first use)
            a = torch.view_as_real(ytilde).pow(2).sum(-1) if ytilde.is_complex() else ytilde ** 2
            self._ysq = a.reshape(lead, B, self._nint).sum(0).contiguous()
        return self._ysq

    def get_k1parts(self, task0, task1, n=None):
        assert 0 <= task0 < self.num_tasks and 0 <= task1 < self.num_tasks
        if n is None:
            n = self._nint
        assert n >= 0
        x, xb = self.xxb_seqs[0][:int(n)]
        if self._FAMILY == 0:
            parts = _lib.lattice_kernel_parts(x.contiguous(), x[0].cpu().numpy(), self._alpha_list)
        else:
            parts = _lib.dnb2_kernel_parts(xb.contiguous(), xb[0].cpu().numpy(), self._alpha_list, self._t)
        return parts[:, None, None, :]

    def get_lam(self, task0, task1, n=None):
        """lam~ = ft(k1) in the reference's normalisation (util.py:95-141)."""
        assert 0 <= task0 < self.num_tasks and 0 <= task1 < self.num_tasks
        if n is None:
            n = self._nint
        n = int(n)
        with torch.no_grad():
            parts = self.get_k1parts(task0, task1, n)[:, 0, 0, :].contiguous()
            scale, ls = self.scale, self.lengthscales
            pshape = torch.broadcast_shapes(scale.shape[:-1], ls.shape[:-1])
            B = _prod(pshape)
            k1 = _lib.kernel_from_parts(parts, scale[..., 0].expand(pshape).reshape(B).contiguous(),
                                        ls.expand(tuple(pshape) + (self.d,)).reshape(B, self.d).contiguous())
            lam = self.ft(k1)
        return lam.reshape(tuple(pshape) + (n,))

    def get_inv_log_det_cache(self, n=None):
        if n is None:
            n = self.n
        if isinstance(n, (int, np.integer)):
            n = torch.tensor([int(n)] * self.num_tasks, dtype=int, device=self.device)
        assert isinstance(n, torch.Tensor) and n.shape == (self.num_tasks,) and (n >= self.n).all()
        ntup = tuple(n.tolist())
        if ntup not in self.inv_log_det_cache_dict.keys():
            self.inv_log_det_cache_dict[ntup] = _FastInverseLogDetCache(self, n) if self._mt is None else multitask.MultiTaskInverseLogDetCache(self, n)
        return self.inv_log_det_cache_dict[ntup]

    def get_inv_log_det(self, n=None):
        return self.get_inv_log_det_cache(n)()

    def get_default_optimizer(self, lr):
        if lr is None:
            lr = 1e-1
        return torch.optim.Rprop(self.parameters(), lr=lr)

    # ------------------------------------------------------------------------------------------------ kernel
    def kernel(self, x: torch.Tensor, z: torch.Tensor, beta0: torch.Tensor = None, beta1: torch.Tensor = None,
               c0: torch.Tensor = None, c1: torch.Tensor = None):
        assert isinstance(x, torch.Tensor) and x.size(-1) == self.d
        assert isinstance(z, torch.Tensor) and z.size(-1) == self.d
        if beta0 is None:
            beta0 = torch.zeros((1, self.d), dtype=int, device=self.device)
        if beta0.shape == (len(beta0),):
            beta0 = beta0[None, :]
        assert isinstance(beta0, torch.Tensor) and beta0.ndim == 2 and beta0.size(1) == self.d
        if beta1 is None:
            beta1 = torch.zeros((1, self.d), dtype=int, device=self.device)
        if beta1.shape == (len(beta1),):
            beta1 = beta1[None, :]
        assert isinstance(beta1, torch.Tensor) and beta1.ndim == 2 and beta1.size(1) == self.d
        assert c0 is None or (isinstance(c0, torch.Tensor) and c0.shape == (beta0.size(0),))
        assert c1 is None or (isinstance(c1, torch.Tensor) and c1.shape == (beta1.size(0),))
        if (beta0 != 0).any() or (beta1 != 0).any() or len(beta0) > 1 or len(beta1) > 1:
            c0 = torch.ones(len(beta0), device=self.device) if c0 is None else c0
            c1 = torch.ones(len(beta1), device=self.device) if c1 is None else c1
            return self._kernel_deriv(x, z, beta0, beta1, c0, c1)
        k = self._kernel(x, z)
        for c in (c0, c1):
            if c is not None:
                k = k * c.to(self.device).reshape(())
        return k

    def _kernel(self, x, z, *unused):
        """k(x,z) with numpy-style broadcasting of the leading dims; hyperparameter batch dims come first
        (abstract_fast_gp.py:181-196)."""
        x = x.to(self.device)
        z = z.to(self.device)
        scale_B, ls_B, _, pshape = self._hyper_host()
        B = len(scale_B)
        lead = torch.broadcast_shapes(x.shape[:-1], z.shape[:-1])
        outs = []
        cross = x.ndim == 3 and z.ndim == 3 and x.shape[1] == 1 and z.shape[0] == 1 and torch.is_floating_point(x)
        for b in range(B):
            if cross:
                k = _lib.cross_kernel(self._FAMILY, x[:, 0, :].contiguous(), z[0].contiguous() if self._FAMILY == 0 or not torch.is_floating_point(z) else self._convert_to_b(z[0]).contiguous(),
                                      self._alpha_list, self._t, scale_B[b], ls_B[b])
            else:
                xe = x.expand(tuple(lead) + (self.d,)).reshape(-1, self.d)
                ze = z.expand(tuple(lead) + (self.d,)).reshape(-1, self.d)
                if not torch.is_floating_point(xe):
                    xe = self._convert_from_b(xe)
                k = _lib.kernel_pairs(self._FAMILY, xe.contiguous(), ze.contiguous(), self._alpha_list, self._t, scale_B[b], ls_B[b]).reshape(lead)
            outs.append(k)
        if len(pshape) == 0:
            return outs[0]
        return torch.stack(outs, 0).reshape(tuple(pshape) + tuple(outs[0].shape))

    # ------------------------------------------------------------------------------------------------ derivative kernels
    def _deriv_terms(self, beta0, beta1, c0, c1):
        """Device term tables (`_lib.DerivTerms`) of the derivative kernel sum_{t0,t1} c0[t0] c1[t1] scale prod_j (ind_j + ls_j part_j)
        (abstract_fast_gp.py:173-191 with fast_gp_lattice.py:267-273 / fast_gp_digital_net_b2.py:289-301)."""
        b0 = np.asarray(beta0.cpu(), dtype=np.int64).reshape(-1, self.d)
        b1 = np.asarray(beta1.cpu(), dtype=np.int64).reshape(-1, self.d)
        w0 = np.asarray(c0.detach().cpu(), dtype=np.float64).reshape(-1)
        w1 = np.asarray(c1.detach().cpu(), dtype=np.float64).reshape(-1)
        key = (b0.tobytes(), b1.tobytes(), w0.tobytes(), w1.tobytes())
        tm = self._deriv_cache.get(key)
        if tm is None:
            nt = len(b0) * len(b1)
            ord_ = np.zeros((nt, self.d), dtype=np.int32)
            par = np.zeros((nt, self.d, _lib.DERIV_STRIDE))
            ind = np.zeros((nt, self.d))
            w = np.zeros(nt)
            for t0 in range(len(b0)):
                for t1 in range(len(b1)):
                    t = t0 * len(b1) + t1
                    w[t] = w0[t0] * w1[t1]
                    for j in range(self.d):
                        ind[t, j] = float(b0[t0, j] + b1[t1, j] == 0)
                        ord_[t, j], par[t, j] = self._deriv_part_spec(self._alpha_list[j], int(b0[t0, j]), int(b1[t1, j]))
            tm = self._deriv_cache[key] = _lib.DerivTerms(ord_, par, ind, w, self.device)
        return tm

    def _kernel_deriv(self, x, z, beta0, beta1, c0, c1):
        """Derivative kernel with numpy-style broadcasting of the leading dims (one hyperparameter set)."""
        x = x.to(self.device)
        z = z.to(self.device)
        scale_B, ls_B, _, pshape = self._hyper_host()
        assert len(scale_B) == 1, "derivative kernels take one hyperparameter set"
        tm = self._deriv_terms(beta0, beta1, c0, c1)
        if x.ndim == 3 and z.ndim == 3 and x.shape[1] == 1 and z.shape[0] == 1 and torch.is_floating_point(x):
            zz = z[0] if self._FAMILY == 0 or not torch.is_floating_point(z) else self._convert_to_b(z[0])
            return _lib.deriv_cross_kernel(self._FAMILY, x[:, 0, :].contiguous(), zz.contiguous(), tm, self._t, scale_B[0], ls_B[0])
        # row pairs: the kernels are shift invariant, so evaluate the parts of delta = x (-) z against the origin
        lead = torch.broadcast_shapes(x.shape[:-1], z.shape[:-1])
        delta = self._ominus(x, z).expand(tuple(lead) + (self.d,)).reshape(-1, self.d).contiguous()
        parts = _lib.deriv_kernel_parts(self._FAMILY, delta, [0] * self.d, tm, self._t)
        ls = torch.as_tensor(ls_B[0], device=self.device)
        return (float(scale_B[0]) * ((tm.ind + ls * parts).prod(-1) * tm.w).sum(-1)).reshape(lead)

    # ------------------------------------------------------------------------------------------------ fit
    def _lam_autograd(self, n=None, with_tilde=False):
        """Full eigenvalues (sqrt(n) ft(k1) + noise) K_task with autograd through the differentiable transform, from the cached
        hyperparameter-independent kernel parts (util.py:95-141, :285-298).  Shape (*param batch, n).  with_tilde: also
        return lam~ = ft(k1), the quantity `get_lam` hands out."""
        n = self._nint if n is None else int(n)
        cached = getattr(self, "_k1parts_cache", None)
        if cached is None or cached[0] != n:
            with torch.no_grad():
                cached = (n, self.get_k1parts(0, 0, n)[:, 0, 0, :].contiguous())
            self._k1parts_cache = cached
        parts = cached[1]
        scale, ls, noise = self.scale, self.lengthscales, self.noise
        k1 = scale * (1 + ls[..., None, :] * parts).prod(-1)
        lam_tilde = self.ft(k1)
        lam = (np.sqrt(n) * lam_tilde + noise) * self.gram_matrix_tasks[..., 0, 0][..., None]
        return (lam, lam_tilde) if with_tilde else lam

    def _autograd_loss(self, loss_metric, masks, cv_weights, d_out, mll_const):
        """The reference's three losses written on the differentiable spectrum (abstract_gp.py:242-273, util.py:354-394);
        used for GCV, CV and masked fits.  Returns (loss, term1, term2, metric_val)."""
        n = self._nint
        lam, lam_tilde = self._lam_autograd(n, with_tilde=True)
        inv = 1 / lam
        ytilde = self.get_ytilde(0)
        ztilde = ytilde * inv
        sb = list(self.shape_batch)
        if loss_metric == "MLL":
            norm_term = (ytilde.conj() * ztilde).real.sum(-1, keepdim=True)
            logdet = torch.log(torch.abs(lam)).sum(-1)[..., None]
            if masks is None:
                term1 = norm_term.sum()
                term2 = d_out / _prod(logdet.shape) * logdet.sum()
            else:
                term1 = norm_term[..., *masks, 0].sum()
                term2 = logdet.expand(sb + [1])[..., *masks, 0].sum()
            loss = 1 / 2 * (term1 + term2 + mll_const)
            return loss, term1, term2, -loss
        if loss_metric == "GCV":
            numer = (ztilde.conj() * ztilde).real.sum(-1, keepdim=True)
            tr_k_inv = inv.real.sum(-1, keepdim=True)
            denom = ((tr_k_inv / n) ** 2).real
            if masks is None:
                term1, term2 = numer, denom
            else:
                term1 = numer[..., *masks, :]
                term2 = denom.expand(sb + [1])[..., *masks, :]
            loss = (term1 / term2).sum()
            return loss, term1, term2, loss
        # CV (abstract_gp.py:262-273 with util.py:381-385): the reference divides by 1 / (lam~ sqrt(n)) WITHOUT the noise
        coeffs = self.ift(ztilde).real
        inv_diag = (1 / (lam_tilde * np.sqrt(n))).mean(-1, keepdim=True)
        if inv_diag.is_complex():
            raise TypeError("loss_metric='CV' is complex-valued for FastGPLattice in the reference (util.py:381-385 divides by the complex lam); it is only defined for FastGPDigitalNetB2")
        squared_sums = ((coeffs / inv_diag) ** 2 * cv_weights).sum(-1, keepdim=True)
        loss = squared_sums.sum() if masks is None else squared_sums[..., *masks, 0].sum()
        nan = torch.nan * torch.ones(1)
        return loss, nan, nan, loss

    def _mll_terms(self, want_grad):
        """Weighted MLL per hyperparameter set through the fused kernel; returns (loss, term1, term2) tensors."""
        scale_B, ls_B, noise_B, pshape = self._hyper()
        B = scale_B.numel()
        d_out = _prod(self.shape_batch)
        ysq = self._get_ysq(pshape)
        if getattr(self, "_mllw", None) is None or self._mllw.shape[0] != B or self._mllw_dout != d_out:
            self._mllw = torch.tensor([0.5, 0.5 * d_out / B], device=self.device).expand(B, 2).contiguous()
            self._mllw_dout = d_out
        lossb, norm, logdet = _MLLFunction.apply(scale_B, ls_B, noise_B, self, self._nint, ysq, self._mllw, want_grad)
        return lossb.sum(), norm.sum(), d_out / B * logdet.sum()

    def fit(self,
            loss_metric: str = "MLL",
            iterations: int = 5000,
            lr: float = None,
            optimizer: torch.optim.Optimizer = None,
            stop_crit_improvement_threshold: float = 5e-2,
            stop_crit_wait_iterations: int = 10,
            store_hists: bool = False,
            store_loss_hist: bool = False,
            store_scale_hist: bool = False,
            store_lengthscales_hist: bool = False,
            store_noise_hist: bool = False,
            store_task_kernel_hist: bool = False,
            verbose: int = 5,
            verbose_indent: int = 4,
            masks: torch.Tensor = None,
            cv_weights: torch.Tensor = 1,
            ):
        """Hyperparameter optimisation; arguments and return value as the reference (abstract_gp.py:152-306).
        The MLL value and its gradient come from the fused CUDA eigen-solve (one call per iteration) instead of an
        autograd tape through log2(n) transform passes."""
        assert isinstance(loss_metric, str) and loss_metric.upper() in ["MLL", "GCV", "CV"]
        assert self._nint > 0, "cannot fit without data"
        assert isinstance(iterations, int) and iterations >= 0
        assert isinstance(store_hists, bool), "require bool store_mll_hist"
        assert isinstance(store_loss_hist, bool), "require bool store_loss_hist"
        assert isinstance(store_scale_hist, bool), "require bool store_scale_hist"
        assert isinstance(store_lengthscales_hist, bool), "require bool store_lengthscales_hist"
        assert isinstance(store_noise_hist, bool), "require bool store_noise_hist"
        assert isinstance(store_task_kernel_hist, bool), "require bool store_task_kernel_hist"
        assert (isinstance(verbose, int) or isinstance(verbose, bool)) and verbose >= 0, "require verbose is a non-negative int"
        assert isinstance(verbose_indent, int) and verbose_indent >= 0, "require verbose_indent is a non-negative int"
        assert np.isscalar(stop_crit_improvement_threshold) and 0 < stop_crit_improvement_threshold, "require stop_crit_improvement_threshold is a positive float"
        assert isinstance(stop_crit_wait_iterations, int) and stop_crit_wait_iterations > 0
        assert masks is None or (isinstance(masks, torch.Tensor))
        loss_metric = loss_metric.upper()
        # MLL without masks: fused CUDA eigen-solve with the analytic gradient.  GCV, CV and masked fits: the same transforms
        # behind torch.autograd (_FTFunction), formulas as in the reference.
        autograd_route = loss_metric != "MLL" or masks is not None or self._mt is not None
        if isinstance(cv_weights, torch.Tensor):
            cv_weights = cv_weights.to(self.device)
        fused = (not autograd_route) and optimizer is None and _FusedFitLoop.eligible(self) and os.environ.get("FGP_B200_GENERIC_FIT") != "1"
        logtol = np.log(1 + stop_crit_improvement_threshold)
        store_loss_hist = store_hists or store_loss_hist
        store_scale_hist = store_hists or (store_scale_hist and self.raw_scale.requires_grad)
        store_lengthscales_hist = store_hists or (store_lengthscales_hist and self.raw_lengthscales.requires_grad)
        store_noise_hist = store_hists or (store_noise_hist and self.raw_noise.requires_grad)
        store_task_kernel_hist = store_hists or (store_task_kernel_hist and (self.raw_factor_task_kernel.requires_grad or self.raw_noise_task_kernel.requires_grad))
        if fused:
            return self._fit_fused(iterations, 1e-1 if lr is None else lr, logtol, stop_crit_wait_iterations, store_loss_hist, store_scale_hist,
                                   store_lengthscales_hist, store_noise_hist, store_task_kernel_hist, verbose, verbose_indent)
        if optimizer is None:
            optimizer = self.get_default_optimizer(lr)
        assert isinstance(optimizer, torch.optim.Optimizer)
        if store_loss_hist:
            loss_hist = torch.empty(iterations + 1)
        if store_scale_hist:
            scale_hist = torch.empty(torch.Size([iterations + 1]) + self.raw_scale.shape)
        if store_lengthscales_hist:
            lengthscales_hist = torch.empty(torch.Size([iterations + 1]) + self.raw_lengthscales.shape)
        if store_noise_hist:
            noise_hist = torch.empty(torch.Size([iterations + 1]) + self.raw_noise.shape)
        if store_task_kernel_hist:
            task_kernel_hist = torch.empty(torch.Size([iterations + 1]) + self.gram_matrix_tasks.shape)
        if masks is not None:
            masks = torch.atleast_2d(masks)
            assert masks.ndim == 2
            assert len(masks) <= len(self.shape_batch)
            d_out = torch.empty(self.shape_batch)[..., *masks].numel()
        else:
            d_out = _prod(self.shape_batch)
        if verbose:
            _s = "%16s | %-10s | %-10s | %-10s" % ("iter of %.1e" % iterations, "loss", "term1", "term2")
            print(" " * verbose_indent + _s)
            print(" " * verbose_indent + "~" * len(_s))
        mll_const = d_out * int(self.n.sum()) * np.log(2 * np.pi)
        stop_crit_best_loss = torch.inf
        stop_crit_save_loss = torch.inf
        stop_crit_iterations_without_improvement_loss = 0
        want_grad = any(p.requires_grad for p in self.parameters())
        if self._mt is None:
            self.get_ytilde(0)
        for i in range(iterations + 1):
            if self._DENSE:
                loss, term1, term2, metric_val = self._mt.loss(loss_metric, d_out, mll_const, masks, cv_weights)
            elif self._mt is not None:
                loss, term1, term2, metric_val = self._mt.loss(loss_metric, d_out, mll_const, masks, cv_weights)
            elif autograd_route:
                loss, term1, term2, metric_val = self._autograd_loss(loss_metric, masks, cv_weights, d_out, mll_const)
            else:
                wsum, term1, term2 = self._mll_terms(want_grad)
                loss = wsum + 0.5 * mll_const
                metric_val = -loss
            lossv = loss.item()
            if lossv < stop_crit_best_loss:
                stop_crit_best_loss = lossv
                best_params = {param[0]: param[1].data.clone() for param in self.named_parameters()}
            if (stop_crit_save_loss - lossv) > logtol:
                stop_crit_iterations_without_improvement_loss = 0
                stop_crit_save_loss = stop_crit_best_loss
            else:
                stop_crit_iterations_without_improvement_loss += 1
            break_condition = i == iterations or stop_crit_iterations_without_improvement_loss == stop_crit_wait_iterations
            if store_loss_hist:
                loss_hist[i] = metric_val.item()
            if store_scale_hist:
                scale_hist[i] = self.scale.detach().to(scale_hist.device)
            if store_lengthscales_hist:
                lengthscales_hist[i] = self.lengthscales.detach().to(lengthscales_hist.device)
            if store_noise_hist:
                noise_hist[i] = self.noise.detach().to(noise_hist.device)
            if store_task_kernel_hist:
                task_kernel_hist[i] = self.gram_matrix_tasks.detach().to(task_kernel_hist.device)
            if verbose and (i % verbose == 0 or break_condition):
                _s = "%16.2e | %-10.2e | %-10.2e | %-10.2e" % (i, lossv, term1.item() if term1.numel() == 1 else torch.nan, term2.item() if term2.numel() == 1 else torch.nan)
                print(" " * verbose_indent + _s)
            if break_condition:
                break
            if want_grad:
                loss.backward()
            optimizer.step()
            optimizer.zero_grad()
        for pname, pdata in best_params.items():
            setattr(self, pname, torch.nn.Parameter(pdata, requires_grad=getattr(self, pname).requires_grad))
        self._epoch += 1
        data = {"iterations": i}
        if store_loss_hist:
            data["loss_hist"] = loss_hist[:(i + 1)]
        if store_scale_hist:
            data["scale_hist"] = scale_hist[:(i + 1)]
        if store_lengthscales_hist:
            data["lengthscales_hist"] = lengthscales_hist[:(i + 1)]
        if store_noise_hist:
            data["noise_hist"] = noise_hist[:(i + 1)]
        if store_task_kernel_hist:
            data["task_kernel_hist"] = task_kernel_hist[:(i + 1)]
        return data

    def _new_fused_loop(self, hist_flags=(False, False, False), hist_capacity=0):
        """Bind this GP to a pooled fit context (buffers + CUDA graphs of its shape): parameters and |ytilde|^2 are copied in."""
        old = getattr(self, "_fused_loop", None)
        if old is not None:
            old.close(finish=False)
        hist_flags = tuple(bool(f) for f in hist_flags)
        pre = getattr(self, "_prearmed_loop", None)
        self._prearmed_loop = None
        if pre is not None:
            # the loop add_y_next armed while the data was in flight: taken if it still fits (same data, same parameter layout, same history
            # rows, enough history capacity); begin() copies the parameter values in either way
            c = pre.ctx
            if (c is not None and pre.data_epoch == self._data_epoch and c.hist_flags == hist_flags and c.hist_capacity >= max(hist_capacity, 1)
                    and pre.param_shapes == tuple(tuple(p.shape) for p in (self.raw_scale, self.raw_lengthscales, self.raw_noise))
                    and pre.param_req == tuple(bool(p.requires_grad) for p in (self.raw_scale, self.raw_lengthscales, self.raw_noise))
                    and pre.tau == self._tau_host()):
                pre.fgp = self  # a strong reference for the duration of the fit (the armed loop only held a weak one)
                self._fused_loop = pre
                return pre
            pre.close(finish=False)
        self.get_ytilde(0)
        loop = _FusedFitLoop(self, hist_flags, max(hist_capacity, 1))
        self._fused_loop = loop
        return loop

    def _prearm_fit(self):
        """Called at the end of add_y_next (one task, default transforms): acquire the pooled fit context and copy |ytilde|^2 into it NOW,
        while the host-to-device copy of y and the transform are still running -- ~75 us of host work that fit() would otherwise do with
        the GPU idle.  Costs nothing on the device but the 8n-byte copy fit() needs anyway; undone by the next add_y_next."""
        pre = getattr(self, "_prearmed_loop", None)
        self._prearmed_loop = None
        if pre is not None:
            pre.close(finish=False)
        if getattr(self, "_fused_loop", None) is not None or not _FusedFitLoop.eligible(self) or os.environ.get("FGP_B200_GENERIC_FIT") == "1":
            return
        loop = _FusedFitLoop(self, (False, False, False), 1)
        loop.data_epoch = self._data_epoch
        loop.param_shapes = tuple(tuple(p.shape) for p in (self.raw_scale, self.raw_lengthscales, self.raw_noise))
        loop.param_req = tuple(bool(p.requires_grad) for p in (self.raw_scale, self.raw_lengthscales, self.raw_noise))
        loop.tau = self._tau_host()
        loop.fgp = weakref.proxy(self)  # no reference cycle: when the GP goes away the loop goes with it and hands its context back
        self._prearmed_loop = loop

    def fit_stepper(self):
        """The fused device-side fit loop armed for an open-ended run: `.step()` = one MLL+gradient+Rprop iteration on the device
        (what fit() replays in chunks); `.close()` copies the best iterate into the parameters and returns the buffers to the pool.
        While the stepper is open the GP's own parameters are NOT updated."""
        assert self._nint > 0, "cannot fit without data"
        assert _FusedFitLoop.eligible(self), "fit_stepper needs the default transforms and parameter layouts"
        loop = self._new_fused_loop()
        loop.begin(2 ** 30, 2 ** 30, np.log(1.05), 1e-1)
        return loop

    def _fit_fused(self, iterations, lr, logtol, stop_wait, store_loss_hist, store_scale_hist, store_lengthscales_hist, store_noise_hist,
                   store_task_kernel_hist, verbose, verbose_indent):
        """fit() on the device: same state machine and outputs as the generic loop below (abstract_gp.py:236-306)."""
        flags = (store_scale_hist, store_lengthscales_hist, store_noise_hist)
        loop = self._new_fused_loop(flags, iterations + 1)
        if verbose:
            _s = "%16s | %-10s | %-10s | %-10s" % ("iter of %.1e" % iterations, "loss", "term1", "term2")
            print(" " * verbose_indent + _s)
            print(" " * verbose_indent + "~" * len(_s))
        loop.begin(iterations, stop_wait, logtol, lr)
        printed = 0
        chunk = min(loop.GRAPH_ITERS, iterations + 1)
        budget = [iterations + 1]  # never enqueue past the iteration budget: launching no-op iterations costs host time

        def enqueue():
            k = min(chunk, budget[0])
            if k > 0:
                loop.replay(k)
                budget[0] -= k
            return k

        if verbose:
            while True:
                enqueue()
                st = loop.read_state()
                last, stopped = int(st[loop.ST_LAST_ITER]), bool(st[loop.ST_STOPPED] != 0)
                loop.check(st)
                rows = loop.loss_hist[printed:last + 1].cpu().numpy()
                for r, row in enumerate(rows):
                    it = printed + r
                    if it % verbose == 0 or (stopped and it == last):
                        print(" " * verbose_indent + "%16.2e | %-10.2e | %-10.2e | %-10.2e" % (it, row[0], row[1], row[2]))
                printed = last + 1
                if stopped:
                    break
        else:
            # one chunk of look-ahead: chunk j+1 is enqueued before the host waits for the state after chunk j
            enqueue()
            loop.snapshot(0)
            j = 0
            while True:
                enqueue()
                loop.snapshot((j + 1) & 1)
                st = loop.wait_snapshot(j & 1)
                last, stopped = int(st[loop.ST_LAST_ITER]), bool(st[loop.ST_STOPPED] != 0)
                loop.check(st)
                if stopped:
                    break
                j += 1
        loop.finish()  # best iterate -> this GP's parameter storages (bumps the cache epoch)
        i = last
        for pname in ("raw_scale", "raw_lengthscales", "raw_noise", "raw_factor_task_kernel", "raw_noise_task_kernel"):
            p = self._parameters[pname]  # fresh Parameter objects as after the reference's fit (abstract_gp.py:297-298), straight into the registry
            self._parameters[pname] = torch.nn.Parameter(p.data, requires_grad=p.requires_grad)
        data = {"iterations": i}
        if store_loss_hist:
            data["loss_hist"] = -(loop.loss_hist[:(i + 1)].cpu()[:, 0])  # one copy of the contiguous rows; the slice and the sign on the host
        if store_scale_hist:
            data["scale_hist"] = loop.scale_hist[:(i + 1)].reshape((i + 1,) + tuple(self.raw_scale.shape)).cpu()
        if store_lengthscales_hist:
            data["lengthscales_hist"] = loop.ls_hist[:(i + 1)].reshape((i + 1,) + tuple(self.raw_lengthscales.shape)).cpu()
        if store_noise_hist:
            data["noise_hist"] = loop.noise_hist[:(i + 1)].reshape((i + 1,) + tuple(self.raw_noise.shape)).cpu()
        if store_task_kernel_hist:
            data["task_kernel_hist"] = self.gram_matrix_tasks.detach().cpu()[None].expand((i + 1,) + tuple(self.gram_matrix_tasks.shape)).clone()
        self._fit_route = ("coop" if loop.multi else "graph", id(loop.ctx))  # diagnostics: which device loop ran, over which pooled context
        loop.close(finish=False)  # buffers and graphs back to the pool: the next fit of this shape (any GP object) reuses them
        self._fused_loop = None
        return data

    # ------------------------------------------------------------------------------------------------ posterior
    def _parse_task(self, task):
        if task is None:
            task = self.default_task
        inttask = isinstance(task, int)
        if inttask:
            task = torch.tensor([task], dtype=int)
        if isinstance(task, list):
            task = torch.tensor(task, dtype=int)
        assert task.ndim == 1 and (task >= 0).all() and (task < self.num_tasks).all()
        return inttask, task

    def _parse_n(self, n):
        if n is None:
            return self._nint if self._mt is None else [int(v) for v in self.n.tolist()]
        if isinstance(n, (list, tuple)):
            n = torch.tensor(n, dtype=int)
        if self._mt is not None:  # per-task sizes (abstract_gp.py:394: an int applies to every task)
            ns = [int(n)] * self.num_tasks if isinstance(n, (int, np.integer)) else [int(v) for v in n.reshape(-1).tolist()]
            cur = [int(v) for v in self.n.tolist()]
            assert len(ns) == self.num_tasks and all((v & (v - 1)) == 0 and v >= c for v, c in zip(ns, cur)), "require n are all power of two greater than or equal to self.n"
            return ns
        if isinstance(n, torch.Tensor):
            vals = set(int(v) for v in n.reshape(-1).tolist())
            assert len(vals) == 1, "one task takes one size (got n=%s)" % sorted(vals)
            n = vals.pop()
        n = int(n)
        assert (n & (n - 1)) == 0 and n >= self._nint, "require n are all power of two greater than or equal to self.n"
        return n

    def post_mean(self, x: torch.Tensor, task: Union[int, torch.Tensor] = None, eval: bool = True):
        """Posterior mean (abstract_gp.py:352-380) as an on-the-fly kernel-vector product: the (N,n) cross-covariance
        is never materialised."""
        assert x.ndim == 2 and x.size(1) == self.d, "x must a torch.Tensor with shape (-1,d)"
        inttask, task = self._parse_task(task)
        if self._mt is not None:
            pmean = self._mt.post_mean(x.to(self.device).contiguous(), task)
            return pmean[..., 0, :] if inttask else pmean
        coeffs = self.coeffs
        x = x.to(self.device).contiguous()
        self._check_unit_cube(x)
        scale_B, ls_B, _, pshape = self._hyper_host()
        B = len(scale_B)
        sb = tuple(self.shape_batch)
        N = x.shape[0]
        xpts = self._xpts(self._nint)
        c = coeffs.reshape(-1, B, self._nint)
        if N == 0:
            pm = torch.empty((c.shape[0], B, 0), dtype=torch.float64, device=self.device)
        elif B == 1:
            pm = _lib.post_mean(self._FAMILY, x, xpts, self._alpha_list, self._t, scale_B[0], ls_B[0], c[:, 0, :].contiguous())
        else:
            pm = torch.empty((c.shape[0], B, N), dtype=torch.float64, device=self.device)
            for b in range(B):
                pm[:, b, :] = _lib.post_mean(self._FAMILY, x, xpts, self._alpha_list, self._t, scale_B[b], ls_B[b], c[:, b, :].contiguous())
        pmean = pm.reshape(sb + (1, N))
        return pmean[..., 0, :] if inttask else pmean

    def post_var(self, x: torch.Tensor, task: Union[int, torch.Tensor] = None, n: Union[int, torch.Tensor] = None, eval: bool = True):
        """Posterior variance (abstract_gp.py:381-416 with the guard of abstract_fast_gp.py:41-46)."""
        n = self._parse_n(n)
        assert x.ndim == 2 and x.size(1) == self.d, "x must a torch.Tensor with shape (-1,d)"
        inttask, task = self._parse_task(task)
        x = x.to(self.device).contiguous()
        self._check_unit_cube(x)
        if self._mt is not None:
            pvar = self._mt.post_var(x, task, n)
            return pvar[..., 0, :] if inttask else pvar
        scale_B, ls_B, _, pshape = self._hyper_host()
        B = len(scale_B)
        lam = self.get_inv_log_det_cache(n)._lam_full()
        xpts = self._xpts(n)
        if x.shape[0] == 0:
            outs = [torch.empty((0,), dtype=torch.float64, device=self.device) for b in range(B)]
        else:
            # lattice with a known generating vector: fused generator form (the points are regenerated inside the first transform pass)
            zgen = self._zgen is not None and all(int(v) < (1 << 32) for v in self._zgen) and os.environ.get("FGP_B200_NO_PVZ") != "1" and _lib.post_var_z_supported(n)
            Cgen = self._Cgen(n) if (self._FAMILY == 1 and os.environ.get("FGP_B200_NO_PVZ") != "1" and _lib.post_var_C_supported(n)) else None
            if zgen:
                shift = self.seqs[0].shift
                outs = [_lib.post_var_z(x, self._zgen, shift, n, self._alpha_list, scale_B[b], ls_B[b], lam[b]) for b in range(B)]
            elif Cgen is not None:  # digital net with known generating matrices: the same fusion with the Walsh-Hadamard transform
                outs = [_lib.post_var_C(x, Cgen, self.seqs[0].rshift, self._t, n, self._alpha_list, scale_B[b], ls_B[b], lam[b]) for b in range(B)]
            else:
                outs = [_lib.post_var(self._FAMILY, x, xpts, self._alpha_list, self._t, scale_B[b], ls_B[b], lam[b]) for b in range(B)]
        pvar = torch.stack(outs, 0).reshape(tuple(pshape) + (1, x.shape[0]))
        return pvar[..., 0, :] if inttask else pvar

    def post_cov(self, x0: torch.Tensor, x1: torch.Tensor, task0: Union[int, torch.Tensor] = None, task1: Union[int, torch.Tensor] = None,
                 n: Union[int, torch.Tensor] = None, eval: bool = True):
        """Posterior covariance matrix (abstract_gp.py:417-474)."""
        n = self._parse_n(n)
        assert x0.ndim == 2 and x0.size(1) == self.d, "x must a torch.Tensor with shape (-1,d)"
        assert x1.ndim == 2 and x1.size(1) == self.d, "z must a torch.Tensor with shape (-1,d)"
        inttask0, task0 = self._parse_task(task0)
        inttask1, task1 = self._parse_task(task1)
        x0 = x0.to(self.device).contiguous()
        x1 = x1.to(self.device).contiguous()
        self._check_unit_cube(x0)
        self._check_unit_cube(x1)
        equal = torch.equal(x0, x1) and torch.equal(task0, task1)
        if self._mt is not None:
            kmat = self._mt.post_cov(x0, x1, task0, task1, n, equal)
        else:
            kmat = self._post_cov_single(x0, x1, n, equal)
        if inttask0 and inttask1:
            return kmat[..., 0, 0, :, :]
        elif inttask0 and not inttask1:
            return kmat[..., 0, :, :, :]
        elif not inttask0 and inttask1:
            return kmat[..., :, 0, :, :]
        return kmat

    def _post_cov_single(self, x0, x1, n, equal):
        """One task: k(x0,x1) - k(x0,X) K^-1 k(X,x1) per hyperparameter set, (*batch, 1, 1, N0, N1)."""
        scale_B, ls_B, _, pshape = self._hyper_host()
        B = len(scale_B)
        cache = self.get_inv_log_det_cache(n)
        lam = cache._lam_full()
        xpts = self._xpts(n)
        fam, al, t = self._FAMILY, self._alpha_list, self._t
        outs = []
        with torch.no_grad():
            for b in range(B):
                knew = _lib.cross_kernel(fam, x0, x1 if fam == 0 else self._convert_to_b(x1), al, t, scale_B[b], ls_B[b])
                k1 = _lib.cross_kernel(fam, x0, xpts, al, t, scale_B[b], ls_B[b])
                k2 = k1 if equal else _lib.cross_kernel(fam, x1, xpts, al, t, scale_B[b], ls_B[b])
                tm = _lib.gram_solve(fam, k2, lam[b])
                kmat = knew - k1 @ tm.T
                if equal:
                    dg = kmat.diagonal()
                    dg.clamp_(min=0)
                outs.append(kmat)
        return torch.stack(outs, 0).reshape(tuple(pshape) + (1, 1) + tuple(outs[0].shape))

    def post_error(self, x: torch.Tensor, task: Union[int, torch.Tensor] = None, n: Union[int, torch.Tensor] = None, confidence: float = 0.99, eval: bool = True):
        assert np.isscalar(confidence) and 0 < confidence < 1, "confidence must be between 0 and 1"
        q = scipy.stats.norm.ppf(1 - (1 - confidence) / 2)
        pvar = self.post_var(x, task=task, n=n, eval=eval)
        pstd = torch.sqrt(pvar)
        perror = q * pstd
        return pvar, q, perror

    def post_ci(self, x: torch.Tensor, task: Union[int, torch.Tensor] = None, confidence: float = 0.99, eval: bool = True):
        # the reference multiplies by the quantile twice (abstract_gp.py:498,523-525); kept for drop-in parity
        assert np.isscalar(confidence) and 0 < confidence < 1, "confidence must be between 0 and 1"
        q = scipy.stats.norm.ppf(1 - (1 - confidence) / 2)
        pmean = self.post_mean(x, task=task, eval=eval)
        pvar, q, perror = self.post_error(x, task=task, confidence=confidence)
        pci_low = pmean - q * perror
        pci_high = pmean + q * perror
        return pmean, pvar, q, pci_low, pci_high

    def post_cubature_mean(self, task: Union[int, torch.Tensor] = None, eval: bool = True):
        """abstract_fast_gp.py:65-81 for one task: scale * sum(coeffs) * K_task."""
        inttask, task = self._parse_task(task)
        if self._mt is not None:
            pcmean = self._mt.post_cubature_mean(task)
            return pcmean[..., 0] if inttask else pcmean
        with torch.no_grad():
            scale_B, _, _, pshape = self._hyper()
            coeffs = self.coeffs
            pcmean = (scale_B.reshape(tuple(pshape) + (1,)) * coeffs).sum(-1)[..., None]
        return pcmean[..., 0] if inttask else pcmean

    def post_cubature_var(self, task: Union[int, torch.Tensor] = None, n: Union[int, torch.Tensor] = None, eval: bool = True):
        """abstract_fast_gp.py:82-109 for one task: s - s^2 n / lam_0 (clamped at 0), s = scale*K_task."""
        n = self._parse_n(n)
        inttask, task = self._parse_task(task)
        if self._mt is not None:
            pcvar = self._mt.post_cubature_cov(task, task, n).diagonal(dim1=-2, dim2=-1).clamp(min=0)
            return pcvar[..., 0] if inttask else pcvar
        with torch.no_grad():
            scale_B, _, _, pshape = self._hyper()
            lam = self.get_inv_log_det_cache(n)._lam_full()
            term = (n / lam[:, 0]).real
            pcvar = (scale_B - scale_B ** 2 * term).clamp_(min=0).reshape(tuple(pshape) + (1,))
        return pcvar[..., 0] if inttask else pcvar

    def post_cubature_cov(self, task0: Union[int, torch.Tensor] = None, task1: Union[int, torch.Tensor] = None, n: Union[int, torch.Tensor] = None, eval: bool = True):
        n = self._parse_n(n)
        inttask0, task0 = self._parse_task(task0)
        inttask1, task1 = self._parse_task(task1)
        if self._mt is not None:
            pccov = self._mt.post_cubature_cov(task0, task1, n)
            if torch.equal(task0, task1):
                pccov.diagonal(dim1=-2, dim2=-1).clamp_(min=0)
        else:
            pccov = self.post_cubature_var(task=[0], n=n)[..., None]
        if inttask0 and inttask1:
            return pccov[..., 0, 0]
        elif inttask0 and not inttask1:
            return pccov[..., 0, :]
        elif not inttask0 and inttask1:
            return pccov[..., :, 0]
        return pccov

    def post_cubature_error(self, task: Union[int, torch.Tensor] = None, n: Union[int, torch.Tensor] = None, confidence: float = 0.99, eval: bool = True):
        assert np.isscalar(confidence) and 0 < confidence < 1, "confidence must be between 0 and 1"
        q = scipy.stats.norm.ppf(1 - (1 - confidence) / 2)
        pcvar = self.post_cubature_var(task=task, n=n, eval=eval)
        pcstd = torch.sqrt(pcvar)
        pcerror = q * pcstd
        return pcvar, q, pcerror

    def post_cubature_ci(self, task: Union[int, torch.Tensor] = None, confidence: float = 0.99, eval: bool = True):
        assert np.isscalar(confidence) and 0 < confidence < 1, "confidence must be between 0 and 1"
        q = scipy.stats.norm.ppf(1 - (1 - confidence) / 2)
        pcmean = self.post_cubature_mean(task=task, eval=eval)
        pcvar, q, pcerror = self.post_cubature_error(task=task, confidence=confidence, eval=eval)
        pcci_low = pcmean - pcerror
        pcci_high = pcmean + pcerror
        return pcmean, pcvar, q, pcci_low, pcci_high


_CTOR_DOC = """
    Args mirror the reference constructor (fast_gp_lattice.py:125-158 / fast_gp_digital_net_b2.py:120-153).  Differences:
    `device` defaults to "cuda" and must be a CUDA device; `seqs` may be an int (dimension), one of this package's
    GPU-side sequence specs (`sequences.Lattice` / `sequences.DigitalNetB2`) or any qmcpy-style sequence object, whose
    points are then taken from its own host generator; `compile_fts*` are accepted and ignored (the transforms are
    hand-written CUDA kernels).
"""


@functools.lru_cache(maxsize=None)
def _bernoulli_poly_coeffs(order):
    """Coefficients (low degree first) of the Bernoulli polynomial B_order(x) = sum_k C(order,k) B_{order-k} x^k, B_1 = -1/2."""
    B = [Fraction(1)]
    for m in range(1, order + 1):
        B.append(-sum(Fraction(math.comb(m + 1, k)) * B[k] for k in range(m)) / (m + 1))
    return tuple(Fraction(math.comb(order, k)) * B[order - k] for k in range(order + 1))


class FastGPLattice(AbstractFastGP):
    """Fast GP regression on rank-1 lattice points with shift-invariant (Bernoulli-polynomial) kernels.
    Drop-in for fastgps.FastGPLattice (fast_gp_lattice.py:7-273) on the single-task path.""" + _CTOR_DOC
    _FAMILY = 0
    _XBDTYPE = torch.float64
    _FTOUTDTYPE = torch.complex128
    _t = 0

    def __init__(self,
                 seqs,
                 num_tasks: int = None,
                 seed_for_seq: int = None,
                 alpha: int = 2,
                 scale: float = 1.,
                 lengthscales: Union[torch.Tensor, float] = 1.,
                 noise: float = 1e-8,
                 factor_task_kernel: Union[torch.Tensor, int] = 1.,
                 rank_factor_task_kernel: int = None,
                 noise_task_kernel: Union[torch.Tensor, float] = 1.,
                 device: torch.device = "cuda",
                 tfs_scale: Tuple[callable, callable] = DEFAULT_TFS_LOG_EXP,
                 tfs_lengthscales: Tuple[callable, callable] = DEFAULT_TFS_LOG_EXP,
                 tfs_noise: Tuple[callable, callable] = DEFAULT_TFS_LOG_EXP,
                 tfs_factor_task_kernel: Tuple[callable, callable] = DEFAULT_TFS_ID,
                 tfs_noise_task_kernel: Tuple[callable, callable] = DEFAULT_TFS_LOG_EXP,
                 requires_grad_scale: bool = True,
                 requires_grad_lengthscales: bool = True,
                 requires_grad_noise: bool = False,
                 requires_grad_factor_task_kernel: bool = None,
                 requires_grad_noise_task_kernel: bool = None,
                 shape_batch: torch.Size = torch.Size([]),
                 shape_scale: torch.Size = torch.Size([1]),
                 shape_lengthscales: torch.Size = None,
                 shape_noise: torch.Size = torch.Size([1]),
                 shape_factor_task_kernel: torch.Size = None,
                 shape_noise_task_kernel: torch.Size = None,
                 derivatives: list = None,
                 derivatives_coeffs: list = None,
                 compile_fts: bool = False,
                 compile_fts_kwargs: dict = {},
                 adaptive_nugget: bool = False,
                 ):
        assert isinstance(alpha, int) and 1 <= alpha <= _lib.MAX_ALPHA, "alpha must be in %s" % list(range(1, _lib.MAX_ALPHA + 1))
        super().__init__(seqs, num_tasks, seed_for_seq, alpha, scale, lengthscales, noise, factor_task_kernel,
                         rank_factor_task_kernel, noise_task_kernel, device, tfs_scale, tfs_lengthscales, tfs_noise,
                         tfs_factor_task_kernel, tfs_noise_task_kernel, requires_grad_scale, requires_grad_lengthscales,
                         requires_grad_noise, requires_grad_factor_task_kernel, requires_grad_noise_task_kernel, shape_batch,
                         shape_scale, shape_lengthscales, shape_noise, shape_factor_task_kernel, shape_noise_task_kernel,
                         derivatives, derivatives_coeffs, compile_fts, compile_fts_kwargs, adaptive_nugget)
        assert all(self.seqs[i].randomize in ['FALSE', 'SHIFT'] for i in range(self.num_tasks)), "each seq should have randomize in ['FALSE','SHIFT']"

    @staticmethod
    def _deriv_part_spec(alpha, beta, kappa):
        """fast_gp_lattice.py:267-273: (-1)^(alpha+kappa+1) (2 pi)^(2 alpha) / order! * B_order(a), order = 2 alpha - beta - kappa,
        as the coefficients (low degree first) of a polynomial in a = (x - z) mod 1; exact rational Bernoulli coefficients."""
        order = 2 * alpha - beta - kappa
        assert 2 <= order, "order must all be at least 2, but got order = %s" % str(order)
        assert order < _lib.DERIV_STRIDE
        coeff = (-1) ** (alpha + kappa + 1) * math.exp(2 * alpha * math.log(2 * math.pi) - math.lgamma(order + 1))
        par = np.zeros(_lib.DERIV_STRIDE)
        for k, c in enumerate(_bernoulli_poly_coeffs(order)):
            par[k] = coeff * float(c)
        return order, par

    @staticmethod
    def _default_sequence(d, seed):
        return sequences.Lattice(d, seed=seed, randomize="SHIFT")

    @staticmethod
    def _adopt_sequence(seq):
        if isinstance(seq, sequences.Lattice):
            return seq
        assert not isinstance(seq, sequences.DigitalNetB2), "each seq should be a lattice sequence"
        return sequences.HostSequence(seq, family=0)

    def get_omega(self, m):
        return torch.exp(-torch.pi * 1j * torch.arange(2 ** m, device=self.device) / 2 ** m)

    def _ft_unstable(self, x):
        # the autograd wrapper costs ~0.1 ms of host time per call: only when a gradient can flow
        return _FTFunction.apply(x, 0) if (torch.is_grad_enabled() and x.requires_grad) else _lib.fftbr(x)

    def _ift_unstable(self, x):
        return _FTFunction.apply(x, 1) if (torch.is_grad_enabled() and x.requires_grad) else _lib.ifftbr(x)

    def _ominus(self, x, z):
        return (x - z) % 1


class FastGPDigitalNetB2(AbstractFastGP):
    """Fast GP regression on base-2 digital nets with digitally-shift-invariant (Walsh) kernels.
    Drop-in for fastgps.FastGPDigitalNetB2 (fast_gp_digital_net_b2.py:7-301) on the single-task path.""" + _CTOR_DOC
    _FAMILY = 1
    _XBDTYPE = torch.int64
    _FTOUTDTYPE = torch.float64

    def __init__(self,
                 seqs,
                 num_tasks: int = None,
                 seed_for_seq: int = None,
                 alpha: int = 2,
                 scale: float = 1.,
                 lengthscales: Union[torch.Tensor, float] = 1.,
                 noise: float = 1e-16,
                 factor_task_kernel: Union[torch.Tensor, int] = 1.,
                 rank_factor_task_kernel: int = None,
                 noise_task_kernel: Union[torch.Tensor, float] = 1.,
                 device: torch.device = "cuda",
                 tfs_scale: Tuple[callable, callable] = DEFAULT_TFS_LOG_EXP,
                 tfs_lengthscales: Tuple[callable, callable] = DEFAULT_TFS_LOG_EXP,
                 tfs_noise: Tuple[callable, callable] = DEFAULT_TFS_LOG_EXP,
                 tfs_factor_task_kernel: Tuple[callable, callable] = DEFAULT_TFS_ID,
                 tfs_noise_task_kernel: Tuple[callable, callable] = DEFAULT_TFS_LOG_EXP,
                 requires_grad_scale: bool = True,
                 requires_grad_lengthscales: bool = True,
                 requires_grad_noise: bool = False,
                 requires_grad_factor_task_kernel: bool = None,
                 requires_grad_noise_task_kernel: bool = None,
                 shape_batch: torch.Size = torch.Size([]),
                 shape_scale: torch.Size = torch.Size([1]),
                 shape_lengthscales: torch.Size = None,
                 shape_noise: torch.Size = torch.Size([1]),
                 shape_factor_task_kernel: torch.Size = None,
                 shape_noise_task_kernel: torch.Size = None,
                 derivatives: list = None,
                 derivatives_coeffs: list = None,
                 compile_fts: bool = False,
                 compile_fts_kwargs: dict = {},
                 adaptive_nugget: bool = False,
                 ):
        super().__init__(seqs, num_tasks, seed_for_seq, alpha, scale, lengthscales, noise, factor_task_kernel,
                         rank_factor_task_kernel, noise_task_kernel, device, tfs_scale, tfs_lengthscales, tfs_noise,
                         tfs_factor_task_kernel, tfs_noise_task_kernel, requires_grad_scale, requires_grad_lengthscales,
                         requires_grad_noise, requires_grad_factor_task_kernel, requires_grad_noise_task_kernel, shape_batch,
                         shape_scale, shape_lengthscales, shape_noise, shape_factor_task_kernel, shape_noise_task_kernel,
                         derivatives, derivatives_coeffs, compile_fts, compile_fts_kwargs, adaptive_nugget)
        assert self.seqs[0].randomize in ['FALSE', 'DS', 'LMS', 'LMS_DS'], "seq should have randomize in ['FALSE','DS','LMS','LMS_DS']"
        ts = [int(self.seqs[i].t) for i in range(self.num_tasks)]
        assert all(t < 64 for t in ts), "each seq must have t<64"
        assert all(t == ts[0] for t in ts), "all seqs should have the same t"
        self.t = self._t = ts[0]
        assert (1 <= self.alpha).all() and (self.alpha <= 4).all()

    @staticmethod
    def _deriv_part_spec(alpha, beta, kappa):
        """fast_gp_digital_net_b2.py:289-301: (-2)^(beta+kappa) ([beta+kappa > 0] + W_order - 1), order = alpha - beta - kappa."""
        order = alpha - beta - kappa
        assert 1 <= order <= 4, "order must all be between 2 and 4, but got order = %s. Try increasing alpha" % str(order)
        par = np.zeros(_lib.DERIV_STRIDE)
        par[0] = float((-2) ** (beta + kappa))
        par[1] = float(beta + kappa > 0)
        return order, par

    @staticmethod
    def _default_sequence(d, seed):
        return sequences.DigitalNetB2(d, seed=seed, randomize="DS")

    @staticmethod
    def _adopt_sequence(seq):
        if isinstance(seq, sequences.DigitalNetB2):
            return seq
        assert not isinstance(seq, sequences.Lattice), "each seq should be a digital net sequence"
        return sequences.HostSequence(seq, family=1)

    def get_omega(self, m):
        return 1

    def _ft_unstable(self, x):
        return _FTFunction.apply(x, 2) if (torch.is_grad_enabled() and x.requires_grad) else _lib.fwht(x)

    _ift_unstable = _ft_unstable

    def _convert_to_b(self, x):
        return torch.floor((x % 1) * 2 ** (self.t)).to(self._XBDTYPE)

    def _convert_from_b(self, xb):
        return xb * 2 ** (-self.t)

    def _ominus(self, x_or_xb, z_or_zb):
        fp_x = torch.is_floating_point(x_or_xb)
        fp_z = torch.is_floating_point(z_or_zb)
        xb = self._convert_to_b(x_or_xb) if fp_x else x_or_xb
        zb = self._convert_to_b(z_or_zb) if fp_z else z_or_zb
        return xb ^ zb
