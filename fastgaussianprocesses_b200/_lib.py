"""ctypes binding of libfgp_b200.so (include/fgp_b200.h).  No CPU fallback: every entry point needs the CUDA library
and CUDA tensors, and fails loudly otherwise."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FGP_B200_LIB") or os.path.join(_HERE, "lib", "libfgp_b200.so")

MAX_D = 32
MAX_ALPHA = 10
DERIV_STRIDE = 24

_c = ctypes
_vp, _i32, _i64, _u64, _f64, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_uint64, _c.c_double, _c.c_size_t

_dp = _c.POINTER(_c.c_double)


class FitLayout(_c.Structure):
    """fgp_fit_layout (include/fgp_b200.h)."""
    _fields_ = [("B", _i32), ("d", _i32), ("n_scale", _i32), ("n_ls_b", _i32), ("n_ls_d", _i32), ("n_noise", _i32),
                ("req_scale", _i32), ("req_ls", _i32), ("req_noise", _i32), ("tau", _f64),
                ("raw_scale", _vp), ("raw_ls", _vp), ("raw_noise", _vp), ("scale_B", _vp), ("ls_B", _vp), ("noise_B", _vp),
                ("state", _vp), ("loss_hist", _vp), ("scale_hist", _vp), ("ls_hist", _vp), ("noise_hist", _vp)]


class FitOptions(_c.Structure):
    """fgp_fit_options (include/fgp_b200.h)."""
    _fields_ = [("iterations", _i32), ("stop_wait", _i32), ("hist_capacity", _i32), ("logtol", _f64), ("half_const", _f64),
                ("wn", _f64), ("wl", _f64), ("lr", _f64), ("etaminus", _f64), ("etaplus", _f64), ("step_min", _f64), ("step_max", _f64)]


class FitProblem(_c.Structure):
    """fgp_fit_problem (include/fgp_b200.h)."""
    _fields_ = [("family", _i32), ("x_dev", _vp), ("z_host", _vp), ("C_dev", _vp), ("mmax", _i32), ("n", _i64), ("d", _i32), ("alpha_host", _vp), ("t", _i32),
                ("ysq_dev", _vp), ("weights_dev", _vp), ("table_dev", _vp), ("workspace_dev", _vp), ("out_dev", _vp)]


# name -> (restype, argtypes); every symbol include/fgp_b200.h declares
SIGNATURES = {
    "fgp_version": (_i32, []),
    "fgp_last_error": (_c.c_char_p, []),
    "fgp_launch_count": (_u64, []),
    "fgp_device_info": (_i32, [_vp, _vp, _vp, _vp]),
    "fgp_lattice_points": (_i32, [_vp, _vp, _i32, _u64, _u64, _vp, _vp]),
    "fgp_dnb2_points": (_i32, [_vp, _i32, _vp, _i32, _i32, _u64, _u64, _vp, _vp, _vp]),
    "fgp_lattice_kernel_parts": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "fgp_dnb2_kernel_parts": (_i32, [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _vp]),
    "fgp_kernel_from_parts": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "fgp_fft_table_bytes": (_sz, [_i64]),
    "fgp_fft_table_init": (_i32, [_i64, _vp, _vp]),
    "fgp_fftbr_r2c": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "fgp_fftbr_c2c": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "fgp_ifftbr_c2c": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "fgp_fwht": (_i32, [_vp, _vp, _i64, _i64, _vp]),
    "fgp_fwht_fused": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "fgp_mll_workspace_bytes": (_sz, [_i32, _i64, _i32, _i32]),
    "fgp_lattice_mll_grad": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "fgp_lattice_mll_grad_z": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "fgp_dnb2_mll_grad_C": (_i32, [_vp, _i32, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "fgp_dnb2_mll_grad": (_i32, [_vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "fgp_fit_state_doubles": (_sz, [_i32, _i32]),
    "fgp_fit_iteration": (_i32, [_c.POINTER(FitProblem), _c.POINTER(FitLayout), _vp]),
    "fgp_fit_iterations": (_i32, [_c.POINTER(FitProblem), _c.POINTER(FitLayout), _i32, _vp]),
    "fgp_fit_iterations_per_launch": (_i32, [_i32, _i64]),
    "fgp_fit_init": (_i32, [_c.POINTER(FitLayout), _c.POINTER(FitOptions), _vp]),
    "fgp_fit_step": (_i32, [_c.POINTER(FitLayout), _vp, _vp]),
    "fgp_fit_finish": (_i32, [_c.POINTER(FitLayout), _vp]),
    "fgp_fit_init_from": (_i32, [_c.POINTER(FitLayout), _c.POINTER(FitOptions), _vp, _vp, _vp, _vp]),
    "fgp_fit_finish_to": (_i32, [_c.POINTER(FitLayout), _vp, _vp, _vp, _vp]),
    "fgp_data_spectrum_workspace_bytes": (_sz, [_i64, _i64]),
    "fgp_data_spectrum": (_i32, [_i32, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "fgp_profile_begin": (_i32, [_vp]),
    "fgp_profile_end": (_i32, [_vp, _i32, _c.POINTER(_c.c_char_p), _c.POINTER(_c.c_float)]),
    "fgp_gram_solve": (_i32, [_i32, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "fgp_post_mean_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "fgp_lattice_post_mean": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _f64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "fgp_dnb2_post_mean": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i32, _f64, _vp, _vp, _i32, _vp, _vp, _vp]),
    "fgp_post_var_workspace_bytes": (_sz, [_i32, _i64, _i64]),
    "fgp_lattice_post_var": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _f64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fgp_dnb2_post_var": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i32, _f64, _vp, _vp, _vp, _vp, _vp]),
    "fgp_lattice_post_var_z_workspace_bytes": (_sz, [_i64, _i64]),
    "fgp_lattice_post_var_z": (_i32, [_vp, _i64, _vp, _vp, _i64, _i32, _vp, _f64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fgp_dnb2_post_var_C_workspace_bytes": (_sz, [_i64, _i64]),
    "fgp_dnb2_post_var_C": (_i32, [_vp, _i64, _vp, _i32, _vp, _i32, _i64, _i32, _vp, _f64, _vp, _vp, _vp, _vp, _vp]),
    "fgp_lattice_cross_kernel": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _f64, _vp, _vp, _vp]),
    "fgp_dnb2_cross_kernel": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i32, _f64, _vp, _vp, _vp]),
    "fgp_kernel_pairs": (_i32, [_i32, _vp, _vp, _i32, _i64, _i32, _vp, _i32, _f64, _vp, _vp, _vp]),
    "fgp_deriv_kernel_parts": (_i32, [_i32, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _vp]),
    "fgp_deriv_cross_kernel": (_i32, [_i32, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _f64, _vp, _vp, _vp]),
    "fgp_block_inv_logdet": (_i32, [_i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "fgp_fp64_peak_probe": (_i32, [_i32, _vp, _vp, _vp]),
}

_lib = None


class FgpError(AssertionError):
    """Raised when the native library reports an error (the reference raises AssertionError at the same places)."""


def load():
    """Load the C-ABI library (no GPU needed to load it; calls need one)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "fastgaussianprocesses_b200: %s is missing -- build it with `python -m fastgaussianprocesses_b200.build` "
                "(there is deliberately no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        raise FgpError("libfgp_b200: %s (code %d)" % (load().fgp_last_error().decode(), rc))


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """The current CUDA stream of the current device as a raw handle.  torch.cuda.current_stream() builds a Stream object through several
    Python layers (~20 us, measured with cProfile inside fit()); the raw getter is what torch's own compiled code paths use."""
    if _raw_stream is not None:
        return _raw_stream(torch._C._cuda_getDevice())
    return torch.cuda.current_stream().cuda_stream


class _NoSwitch(object):
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def on_device(device):
    """`with on_device(dev):` -- torch.cuda.device(dev) only when dev is not already the current device (the context manager costs
    ~15 us of host time per use, and every C call of the fit loop sits inside one)."""
    idx = device.index
    if idx is None or idx == torch._C._cuda_getDevice():
        return _NO_SWITCH
    return torch.cuda.device(device)


def _dev(t, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("fastgaussianprocesses_b200 runs on CUDA tensors only (no CPU fallback); got %r" % (getattr(t, "device", type(t)),))
    if dtype is not None and t.dtype != dtype:
        raise TypeError("expected %s, got %s" % (dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    return t.data_ptr()


def _harr(ctype, vals):
    vals = list(vals)
    return (ctype * max(1, len(vals)))(*vals)


def launch_count():
    return int(load().fgp_launch_count())


def device_info():
    sm, ma, mi, sh = _c.c_int(), _c.c_int(), _c.c_int(), _c.c_size_t()
    _check(load().fgp_device_info(_c.byref(sm), _c.byref(ma), _c.byref(mi), _c.byref(sh)))
    return {"sm_count": sm.value, "cc": (ma.value, mi.value), "smem_optin": sh.value}


# --------------------------------------------------------------------------------------------- K1 points
def lattice_points(z, shift, i0, i1, device):
    d = len(z)
    x = torch.empty((int(i1) - int(i0), d), dtype=torch.float64, device=device)
    if x.numel() == 0:
        return x
    with on_device(x.device):
        _check(load().fgp_lattice_points(_harr(_u64, [int(v) for v in z]), _harr(_f64, [float(v) for v in shift]), d,
                                         int(i0), int(i1), x.data_ptr(), _stream()))
    return x


def dnb2_points(C_dev, dshift, t, i0, i1, want_x=True):
    d, mmax = C_dev.shape
    _dev(C_dev, torch.int64)
    xb = torch.empty((int(i1) - int(i0), d), dtype=torch.int64, device=C_dev.device)
    x = torch.empty((int(i1) - int(i0), d), dtype=torch.float64, device=C_dev.device) if want_x else None
    if xb.numel() == 0:
        return xb, x
    with on_device(C_dev.device):
        _check(load().fgp_dnb2_points(C_dev.data_ptr(), mmax, _harr(_u64, [int(v) for v in dshift]), d, int(t), int(i0), int(i1),
                                      xb.data_ptr(), x.data_ptr() if want_x else None, _stream()))
    return xb, x


# --------------------------------------------------------------------------------------------- K2 kernels
def lattice_kernel_parts(x, z, alpha):
    n, d = x.shape
    out = torch.empty_like(x)
    with on_device(x.device):
        _check(load().fgp_lattice_kernel_parts(_dev(x, torch.float64), n, d, _harr(_f64, z), _harr(_i32, alpha), out.data_ptr(), _stream()))
    return out


def dnb2_kernel_parts(xb, zb, alpha, t):
    n, d = xb.shape
    out = torch.empty((n, d), dtype=torch.float64, device=xb.device)
    with on_device(xb.device):
        _check(load().fgp_dnb2_kernel_parts(_dev(xb, torch.int64), n, d, _harr(_i64, zb), _harr(_i32, alpha), int(t), out.data_ptr(), _stream()))
    return out


def kernel_from_parts(parts, scale, ls):
    """parts (n,d); scale (B,), ls (B,d) device tensors -> (B,n)."""
    n, d = parts.shape
    B = scale.numel()
    out = torch.empty((B, n), dtype=torch.float64, device=parts.device)
    with on_device(parts.device):
        _check(load().fgp_kernel_from_parts(_dev(parts, torch.float64), n, d, B, _dev(scale, torch.float64), _dev(ls, torch.float64),
                                            out.data_ptr(), _stream()))
    return out


def cross_kernel(family, xs, xtrain, alpha, t, scale, ls):
    m, d = xs.shape
    n = xtrain.shape[0]
    out = torch.empty((m, n), dtype=torch.float64, device=xs.device)
    with on_device(xs.device):
        if family == 0:
            _check(load().fgp_lattice_cross_kernel(_dev(xs, torch.float64), m, _dev(xtrain, torch.float64), n, d, _harr(_i32, alpha),
                                                   float(scale), _harr(_f64, ls), out.data_ptr(), _stream()))
        else:
            _check(load().fgp_dnb2_cross_kernel(_dev(xs, torch.float64), m, _dev(xtrain, torch.int64), n, d, _harr(_i32, alpha), int(t),
                                                float(scale), _harr(_f64, ls), out.data_ptr(), _stream()))
    return out


def kernel_pairs(family, x, z, alpha, t, scale, ls):
    """k(x_i, z_i) for row pairs: x (N,d) float64, z (N,d) float64 (or int64 net integers) -> (N,)."""
    N, d = x.shape
    out = torch.empty((N,), dtype=torch.float64, device=x.device)
    z_is_int = z.dtype == torch.int64
    with on_device(x.device):
        _check(load().fgp_kernel_pairs(int(family), _dev(x, torch.float64), _dev(z, torch.int64 if z_is_int else torch.float64), int(z_is_int),
                                       N, d, _harr(_i32, alpha), int(t), float(scale), _harr(_f64, ls), out.data_ptr(), _stream()))
    return out


class DerivTerms(object):
    """Device tables of the derivative terms of one task pair (include/fgp_b200.h, fgp_deriv_*): ord (nt*d) int32,
    par (nt*d*DERIV_STRIDE), ind (nt*d), w (nt) float64."""

    def __init__(self, ord_, par, ind, w, device):
        self.nt, self.d = ord_.shape
        assert par.shape == (self.nt, self.d, DERIV_STRIDE) and ind.shape == (self.nt, self.d) and w.shape == (self.nt,)
        self.ord = torch.as_tensor(ord_, dtype=torch.int32).contiguous().to(device)
        self.par = torch.as_tensor(par, dtype=torch.float64).contiguous().to(device)
        self.ind = torch.as_tensor(ind, dtype=torch.float64).contiguous().to(device)
        self.w = torch.as_tensor(w, dtype=torch.float64).contiguous().to(device)


def deriv_kernel_parts(family, x, z, terms, t):
    """parts (n, nt, d) of the points x (float64 lattice / int64 net) against the single point z (length-d host list)."""
    n, d = x.shape
    out = torch.empty((n, terms.nt, d), dtype=torch.float64, device=x.device)
    with on_device(x.device):
        zh = _harr(_f64, [float(v) for v in z]) if family == 0 else _harr(_i64, [int(v) for v in z])
        _check(load().fgp_deriv_kernel_parts(int(family), _dev(x, torch.float64 if family == 0 else torch.int64), n, d, zh, terms.nt,
                                             _dev(terms.ord, torch.int32), _dev(terms.par, torch.float64), int(t), out.data_ptr(), _stream()))
    return out


def deriv_cross_kernel(family, xs, xtrain, terms, t, scale, ls):
    """K (m,n) = derivative kernel of float test points xs against xtrain (float64 lattice / int64 net)."""
    m, d = xs.shape
    n = xtrain.shape[0]
    out = torch.empty((m, n), dtype=torch.float64, device=xs.device)
    with on_device(xs.device):
        _check(load().fgp_deriv_cross_kernel(int(family), _dev(xs, torch.float64), m, _dev(xtrain, torch.float64 if family == 0 else torch.int64), n, d,
                                             terms.nt, _dev(terms.ord, torch.int32), _dev(terms.par, torch.float64), _dev(terms.ind, torch.float64),
                                             _dev(terms.w, torch.float64), int(t), float(scale), _harr(_f64, ls), out.data_ptr(), _stream()))
    return out


# --------------------------------------------------------------------------------------------- K3 transforms
_tables = {}


def fft_table(n, device):
    """Twiddle tables for length-n FFT-BRO on `device` (built once per (n, device) by a tiny kernel)."""
    device = torch.device(device)
    key = (int(n), device.index if device.index is not None else torch.cuda.current_device())
    tab = _tables.get(key)
    if tab is None:
        nbytes = load().fgp_fft_table_bytes(int(n))
        tab = torch.empty(nbytes // 8, dtype=torch.float64, device=device)
        with on_device(device):
            _check(load().fgp_fft_table_init(int(n), tab.data_ptr(), _stream()))
        _tables[key] = tab
    return tab


def _as2d(x):
    _dev(x)
    n = x.shape[-1]
    return x.reshape(-1, n) if x.numel() else x.reshape(0, n), n


def fftbr(x):
    """Orthonormal FFT, bit-reversed-order input -> natural-order output, last dim.  Real or complex input."""
    x2, n = _as2d(x.contiguous())
    out = torch.empty(x2.shape, dtype=torch.complex128, device=x.device)
    tab = fft_table(n, x.device)
    with on_device(x.device):
        if x2.dtype == torch.float64:
            _check(load().fgp_fftbr_r2c(_dev(x2), out.data_ptr(), x2.shape[0], n, tab.data_ptr(), _stream()))
        elif x2.dtype == torch.complex128:
            _check(load().fgp_fftbr_c2c(_dev(x2), out.data_ptr(), x2.shape[0], n, tab.data_ptr(), _stream()))
        else:
            raise TypeError("fftbr needs float64 or complex128")
    return out.reshape(x.shape)


def ifftbr(x):
    x = x.contiguous()
    if x.dtype == torch.float64:
        x = x.to(torch.complex128)
    x2, n = _as2d(x)
    out = torch.empty_like(x2)
    tab = fft_table(n, x.device)
    with on_device(x.device):
        _check(load().fgp_ifftbr_c2c(_dev(x2, torch.complex128), out.data_ptr(), x2.shape[0], n, tab.data_ptr(), _stream()))
    return out.reshape(x.shape)


_fused_ctl = {}


def fwht(x, fused=None):
    """Orthonormal FWHT along the last dim.  fused=True (or FGP_B200_FUSED_FWHT=1) runs two-pass sizes (n > 2^12) as ONE
    persistent kernel whose intermediate stays in the L2 (fgp_fwht_fused): the same results (to FMA contraction), but measured slower than the
    two launches on B200 (64 x 2^20: 0.45 ms against 0.35 ms, profiles/README.md snapshot j), so it is off by default."""
    x2, n = _as2d(x.contiguous())
    out = torch.empty_like(x2)
    if fused is None:
        fused = os.environ.get("FGP_B200_FUSED_FWHT") == "1"
    with on_device(x.device):
        if fused and n > 4096 and x2.shape[0] < (1 << 24):
            # control block per (device, stream): the kernel leaves it zeroed, calls on one stream are serialised
            key = (x.device.index, _stream())
            ctl = _fused_ctl.get(key)
            if ctl is None or ctl.numel() < x2.shape[0] + 2:
                ctl = _fused_ctl[key] = torch.zeros(max(1026, x2.shape[0] + 2), dtype=torch.int32, device=x.device)
            _check(load().fgp_fwht_fused(_dev(x2, torch.float64), out.data_ptr(), x2.shape[0], n, ctl.data_ptr(), _stream()))
        else:
            _check(load().fgp_fwht(_dev(x2, torch.float64), out.data_ptr(), x2.shape[0], n, _stream()))
    return out.reshape(x.shape)


# --------------------------------------------------------------------------------------------- K4 MLL
_workspaces = {}


def release_workspaces():
    """Drop every cached scratch buffer (they are kept per (kind, device, stream) for the life of the process otherwise)."""
    _workspaces.clear()


def _workspace(kind, nbytes, device):
    """Scratch buffer cached per (kind, device, stream): two streams never share one (calls on one stream are serialised)."""
    device = torch.device(device)
    key = (kind, device.index if device.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() * 8 < nbytes:
        ws = torch.empty((max(nbytes, 256) + 7) // 8, dtype=torch.float64, device=device)
        _workspaces[key] = ws
    return ws


def mll_grad(family, xpts, alpha, t, ysq, scale, ls, noise, want_grad=True, want_lam=False, weights=None, z=None, C=None):
    """Fused MLL terms + gradients.  xpts: (n,d) float64 (lattice) / int64 (net); ysq (B,n); scale (B,), ls (B,d), noise (B,).
    z: lattice generating vector / C: net generating matrices (d, mmax) int64 device tensor -> generator mode (the points are
    regenerated from the index, xpts only gives n, d, device).
    Returns out (B, d+4) = [norm, logdet, dL/dnoise, dL/dscale, dL/dls...] and lam (B,n) or None."""
    n, d = xpts.shape
    B = scale.numel()
    dev = xpts.device
    out = torch.zeros((B, d + 4), dtype=torch.float64, device=dev)
    lam = torch.empty((B, n), dtype=torch.complex128 if family == 0 else torch.float64, device=dev) if want_lam else None
    ws = _workspace("mll", load().fgp_mll_workspace_bytes(family, n, d, B), dev)
    wptr = None if weights is None else _dev(weights, torch.float64)
    with on_device(dev):
        if family == 0 and z is not None:
            tab = fft_table(n, dev)
            _check(load().fgp_lattice_mll_grad_z(_harr(_u64, [int(v) for v in z]), n, d, _harr(_i32, alpha), B, _dev(ysq, torch.float64),
                                                 _dev(scale, torch.float64), _dev(ls, torch.float64), _dev(noise, torch.float64), wptr,
                                                 tab.data_ptr(), ws.data_ptr(), lam.data_ptr() if want_lam else None, out.data_ptr(),
                                                 1 if want_grad else 0, _stream()))
        elif family == 1 and C is not None:
            _check(load().fgp_dnb2_mll_grad_C(_dev(C, torch.int64), int(C.shape[1]), n, d, _harr(_i32, alpha), int(t), B, _dev(ysq, torch.float64),
                                              _dev(scale, torch.float64), _dev(ls, torch.float64), _dev(noise, torch.float64), wptr,
                                              ws.data_ptr(), lam.data_ptr() if want_lam else None, out.data_ptr(), 1 if want_grad else 0, _stream()))
        elif family == 0:
            tab = fft_table(n, dev)
            _check(load().fgp_lattice_mll_grad(_dev(xpts, torch.float64), n, d, _harr(_i32, alpha), B, _dev(ysq, torch.float64),
                                               _dev(scale, torch.float64), _dev(ls, torch.float64), _dev(noise, torch.float64), wptr,
                                               tab.data_ptr(), ws.data_ptr(), lam.data_ptr() if want_lam else None, out.data_ptr(),
                                               1 if want_grad else 0, _stream()))
        else:
            _check(load().fgp_dnb2_mll_grad(_dev(xpts, torch.int64), n, d, _harr(_i32, alpha), int(t), B, _dev(ysq, torch.float64),
                                            _dev(scale, torch.float64), _dev(ls, torch.float64), _dev(noise, torch.float64), wptr,
                                            ws.data_ptr(), lam.data_ptr() if want_lam else None, out.data_ptr(),
                                            1 if want_grad else 0, _stream()))
    return out, lam


def mll_grad_into(family, xpts, alpha, t, ysq, scale, ls, noise, weights, ws, lam, out, want_grad=True, z=None):
    """Allocation-free form of `mll_grad` (all buffers caller-provided) -- safe inside CUDA graph capture."""
    n, d = xpts.shape
    B = scale.numel()
    wptr = None if weights is None else weights.data_ptr()
    lptr = None if lam is None else lam.data_ptr()
    if family == 0 and z is not None:
        tab = fft_table(n, xpts.device)
        _check(load().fgp_lattice_mll_grad_z(_harr(_u64, [int(v) for v in z]), n, d, _harr(_i32, alpha), B, ysq.data_ptr(), scale.data_ptr(), ls.data_ptr(),
                                             noise.data_ptr(), wptr, tab.data_ptr(), ws.data_ptr(), lptr, out.data_ptr(), 1 if want_grad else 0, _stream()))
    elif family == 0:
        tab = fft_table(n, xpts.device)
        _check(load().fgp_lattice_mll_grad(xpts.data_ptr(), n, d, _harr(_i32, alpha), B, ysq.data_ptr(), scale.data_ptr(), ls.data_ptr(),
                                           noise.data_ptr(), wptr, tab.data_ptr(), ws.data_ptr(), lptr, out.data_ptr(), 1 if want_grad else 0, _stream()))
    else:
        _check(load().fgp_dnb2_mll_grad(xpts.data_ptr(), n, d, _harr(_i32, alpha), int(t), B, ysq.data_ptr(), scale.data_ptr(), ls.data_ptr(),
                                        noise.data_ptr(), wptr, ws.data_ptr(), lptr, out.data_ptr(), 1 if want_grad else 0, _stream()))


def mll_workspace(family, n, d, B, device):
    ws = torch.empty((max(load().fgp_mll_workspace_bytes(family, n, d, B), 256) + 7) // 8, dtype=torch.float64, device=device)
    ws[-32:].zero_()  # control words of the persistent kernel (include/fgp_b200.h): zero once, every launch leaves them zeroed
    return ws


def fit_state_doubles(P, B):
    return int(load().fgp_fit_state_doubles(int(P), int(B)))


def fit_iteration(problem, layout):
    _check(load().fgp_fit_iteration(_c.byref(problem), _c.byref(layout), _stream()))


def fit_iterations(problem, layout, k):
    _check(load().fgp_fit_iterations(_c.byref(problem), _c.byref(layout), int(k), _stream()))


def fit_iterations_per_launch(family, n):
    return int(load().fgp_fit_iterations_per_launch(int(family), int(n)))


def fit_init(layout, options):
    _check(load().fgp_fit_init(_c.byref(layout), _c.byref(options), _stream()))


def fit_step(layout, out):
    _check(load().fgp_fit_step(_c.byref(layout), out.data_ptr(), _stream()))


def fit_finish(layout):
    _check(load().fgp_fit_finish(_c.byref(layout), _stream()))


def fit_init_from(layout, options, raw_scale, raw_ls, raw_noise):
    """fgp_fit_init_from: the three parameter tensors (float64, contiguous, shapes of the layout) are copied into the layout's staging first."""
    _check(load().fgp_fit_init_from(_c.byref(layout), _c.byref(options), raw_scale.data_ptr(), raw_ls.data_ptr(), raw_noise.data_ptr(), _stream()))


def fit_finish_to(layout, raw_scale, raw_ls, raw_noise):
    """fgp_fit_finish_to: best iterate -> the layout's staging AND the three given parameter tensors."""
    _check(load().fgp_fit_finish_to(_c.byref(layout), raw_scale.data_ptr(), raw_ls.data_ptr(), raw_noise.data_ptr(), _stream()))


def data_spectrum(family, y, B):
    """y (rows, n) float64 on the GPU, rows = lead * B -> (ytilde (rows, n) complex128 / float64, ysq (B, n)); fgp_data_spectrum."""
    rows, n = y.shape
    dev = y.device
    yt = torch.empty((rows, n), dtype=torch.complex128 if family == 0 else torch.float64, device=dev)
    ysq = torch.empty((B, n), dtype=torch.float64, device=dev)
    ws = _workspace("spectrum", load().fgp_data_spectrum_workspace_bytes(rows, n), dev)
    table = fft_table(n, dev).data_ptr() if family == 0 else None
    with on_device(dev):
        _check(load().fgp_data_spectrum(family, _dev(y, torch.float64), rows, B, n, table, yt.data_ptr(), ysq.data_ptr(), ws.data_ptr(), _stream()))
    return yt, ysq


def profile_begin():
    _check(load().fgp_profile_begin(_stream()))


def profile_end(max_entries=256):
    names = (_c.c_char_p * max_entries)()
    ms = (_c.c_float * max_entries)()
    cnt = load().fgp_profile_end(_stream(), max_entries, names, ms)
    if cnt < 0:
        _check(cnt)
    return [(names[i].decode(), float(ms[i])) for i in range(cnt)]


def gram_solve(family, y, lam):
    """K^-1 y for y (..., n) with full eigenvalues lam (n,)."""
    y2, n = _as2d(y.contiguous())
    R = y2.shape[0]
    out = torch.empty_like(y2)
    with on_device(y.device):
        if family == 0:
            tab = fft_table(n, y.device)
            work = _workspace("solve", R * n * 16, y.device)
            _check(load().fgp_gram_solve(0, _dev(y2, torch.float64), out.data_ptr(), R, n, _dev(lam, torch.complex128), tab.data_ptr(),
                                         work.data_ptr(), _stream()))
        else:
            _check(load().fgp_gram_solve(1, _dev(y2, torch.float64), out.data_ptr(), R, n, _dev(lam, torch.float64), None, None, _stream()))
    return out.reshape(y.shape)


# --------------------------------------------------------------------------------------------- K5 posterior
def post_mean(family, xs, xtrain, alpha, t, scale, ls, coeffs):
    """xs (m,d); coeffs (B,n) -> (B,m)."""
    m, d = xs.shape
    n = xtrain.shape[0]
    B = coeffs.shape[0]
    out = torch.empty((B, m), dtype=torch.float64, device=xs.device)
    ws = _workspace("pmean", load().fgp_post_mean_workspace_bytes(m, n, d, B), xs.device)
    with on_device(xs.device):
        if family == 0:
            _check(load().fgp_lattice_post_mean(_dev(xs, torch.float64), m, _dev(xtrain, torch.float64), n, d, _harr(_i32, alpha),
                                                float(scale), _harr(_f64, ls), _dev(coeffs, torch.float64), B, ws.data_ptr(),
                                                out.data_ptr(), _stream()))
        else:
            _check(load().fgp_dnb2_post_mean(_dev(xs, torch.float64), m, _dev(xtrain, torch.int64), n, d, _harr(_i32, alpha), int(t),
                                             float(scale), _harr(_f64, ls), _dev(coeffs, torch.float64), B, ws.data_ptr(),
                                             out.data_ptr(), _stream()))
    return out


def post_var(family, xs, xtrain, alpha, t, scale, ls, lam):
    m, d = xs.shape
    n = xtrain.shape[0]
    out = torch.empty((m,), dtype=torch.float64, device=xs.device)
    ws = _workspace("pvar", load().fgp_post_var_workspace_bytes(family, m, n), xs.device)
    with on_device(xs.device):
        if family == 0:
            tab = fft_table(n, xs.device)
            _check(load().fgp_lattice_post_var(_dev(xs, torch.float64), m, _dev(xtrain, torch.float64), n, d, _harr(_i32, alpha),
                                               float(scale), _harr(_f64, ls), _dev(lam, torch.complex128), tab.data_ptr(), ws.data_ptr(),
                                               out.data_ptr(), _stream()))
        else:
            _check(load().fgp_dnb2_post_var(_dev(xs, torch.float64), m, _dev(xtrain, torch.int64), n, d, _harr(_i32, alpha), int(t),
                                            float(scale), _harr(_f64, ls), _dev(lam, torch.float64), ws.data_ptr(), out.data_ptr(),
                                            _stream()))
    return out


def post_var_z(xs, z, shift, n, alpha, scale, ls, lam):
    """Fused lattice posterior variance in generator form (fgp_lattice_post_var_z); falls back is the caller's business:
    `post_var_z_supported(n)` says whether the size is a two-pass size."""
    m, d = xs.shape
    out = torch.empty((m,), dtype=torch.float64, device=xs.device)
    ws = _workspace("pvarz", load().fgp_lattice_post_var_z_workspace_bytes(m, n), xs.device)
    with on_device(xs.device):
        tab = fft_table(n, xs.device)
        _check(load().fgp_lattice_post_var_z(_dev(xs, torch.float64), m, _harr(_u64, [int(v) for v in z]), _harr(_f64, [float(v) for v in shift]), n, d,
                                             _harr(_i32, alpha), float(scale), _harr(_f64, ls), _dev(lam, torch.complex128), tab.data_ptr(),
                                             ws.data_ptr(), out.data_ptr(), _stream()))
    return out


def post_var_C(xs, C, dshift, t, n, alpha, scale, ls, lam):
    """Fused digital-net posterior variance in generator form (fgp_dnb2_post_var_C); C: (d, mmax) int64 device tensor."""
    m, d = xs.shape
    out = torch.empty((m,), dtype=torch.float64, device=xs.device)
    ws = _workspace("pvarC", load().fgp_dnb2_post_var_C_workspace_bytes(m, n), xs.device)
    with on_device(xs.device):
        _check(load().fgp_dnb2_post_var_C(_dev(xs, torch.float64), m, _dev(C, torch.int64), int(C.shape[1]), _harr(_u64, [int(v) for v in dshift]), int(t), n, d,
                                          _harr(_i32, alpha), float(scale), _harr(_f64, ls), _dev(lam, torch.float64), ws.data_ptr(), out.data_ptr(), _stream()))
    return out


def post_var_C_supported(n):
    return load().fgp_dnb2_post_var_C_workspace_bytes(2, int(n)) > 0


def post_var_z_supported(n):
    return load().fgp_lattice_post_var_z_workspace_bytes(2, int(n)) > 0


def block_inv_logdet(L):
    """(nm, R, R) complex128 / float64 -> (inverse (nm, R, R), log|det| (nm)); fgp_block_inv_logdet."""
    assert L.ndim == 3 and L.shape[1] == L.shape[2]
    L = L.contiguous()
    A = torch.empty_like(L)
    logdet = torch.empty(L.shape[0], dtype=torch.float64, device=L.device)
    with on_device(L.device):
        _check(load().fgp_block_inv_logdet(1 if L.is_complex() else 0, _dev(L, L.dtype), L.shape[0], L.shape[1], A.data_ptr(), logdet.data_ptr(), _stream()))
    return A, logdet


def fp64_peak_probe(iters, device):
    sink = torch.zeros(8, dtype=torch.float64, device=device)
    flops = _c.c_double()
    with torch.cuda.device(device):
        _check(load().fgp_fp64_peak_probe(int(iters), sink.data_ptr(), _c.byref(flops), _stream()))
    return flops.value
