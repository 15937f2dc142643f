// K4b device side: fit() bookkeeping shared by the stand-alone fit_step kernel and the tail of the fused MLL kernels.
#pragma once
#include "fgp_common.cuh"

namespace fgp {

// slots of the device state block (doubles)
enum {
  ST_BEST = 0, ST_SAVE, ST_WAIT, ST_ITER, ST_STOPPED, ST_LAST_ITER, ST_LAST_LOSS, ST_TERM1, ST_TERM2,
  ST_ITERATIONS, ST_STOP_WAIT, ST_LOGTOL, ST_HALF_CONST, ST_WN, ST_WL, ST_LR, ST_ETAM, ST_ETAP, ST_SMIN, ST_SMAX,
  ST_HIST_CAP, ST_HEADER = 32
};

struct FitLayout {
  int B, d;
  int n_scale, n_ls_b, n_ls_d, n_noise;
  int req_scale, req_ls, req_noise;
  int P;
  double tau;
  double* raw_scale;
  double* raw_ls;
  double* raw_noise;
  double* scale_B;
  double* ls_B;
  double* noise_B;
  double* state;
  double* loss_hist;
  double* scale_hist;
  double* ls_hist;
  double* noise_hist;
  unsigned int* tickets;  // B + 1 counters behind the state block (zeroed by fit_init, self-resetting)
};

struct FitParamPtrs {  // optional caller-side storages of the three raw parameter groups (all null: unused)
  double *scale, *ls, *noise;
};

__device__ __forceinline__ void write_effective(const FitLayout& c) {
  for (int b = threadIdx.x; b < c.B; b += blockDim.x) {
    c.scale_B[b] = c.tau * exp(c.raw_scale[c.n_scale == 1 ? 0 : b]);
    c.noise_B[b] = c.tau * exp(c.raw_noise[c.n_noise == 1 ? 0 : b]);
  }
  for (int e = threadIdx.x; e < c.B * c.d; e += blockDim.x) {
    const int b = e / c.d, j = e - b * c.d;
    c.ls_B[e] = exp(c.raw_ls[(c.n_ls_b == 1 ? 0 : b) * c.n_ls_d + (c.n_ls_d == 1 ? 0 : j)]);
  }
}

// gradient w.r.t. raw element e of parameter group g (0 scale, 1 lengthscales, 2 noise); theta = exp(raw) => dtheta/draw = theta
// s_out: the (d+4) reduced terms of the only set in shared memory (B == 1), else NULL and `out` is read from global memory
__device__ __forceinline__ double raw_grad(const FitLayout& c, const double* out, const double* s_out, int g, int e) {
  const int stride = c.d + 4;
  if (s_out) {
    if (g == 0) return s_out[3] * c.scale_B[0];
    if (g == 2) return s_out[2] * c.noise_B[0];
    const int j0 = c.n_ls_d == 1 ? 0 : e, j1 = c.n_ls_d == 1 ? c.d : e + 1;
    double s = 0.0;
    for (int j = j0; j < j1; ++j) s += s_out[4 + j] * c.ls_B[j];
    return s;
  }
  double s = 0.0;
  if (g == 0) {
    if (c.n_scale == 1) {
      for (int b = 0; b < c.B; ++b) s += __ldcg(out + b * stride + 3) * c.scale_B[b];
    } else {
      s = __ldcg(out + e * stride + 3) * c.scale_B[e];
    }
  } else if (g == 2) {
    if (c.n_noise == 1) {
      for (int b = 0; b < c.B; ++b) s += __ldcg(out + b * stride + 2) * c.noise_B[b];
    } else {
      s = __ldcg(out + e * stride + 2) * c.noise_B[e];
    }
  } else {
    const int eb = e / c.n_ls_d, ej = e - eb * c.n_ls_d;
    const int b0 = c.n_ls_b == 1 ? 0 : eb, b1 = c.n_ls_b == 1 ? c.B : eb + 1;
    const int j0 = c.n_ls_d == 1 ? 0 : ej, j1 = c.n_ls_d == 1 ? c.d : ej + 1;
    for (int b = b0; b < b1; ++b)
      for (int j = j0; j < j1; ++j) s += __ldcg(out + b * stride + 4 + j) * c.ls_B[b * c.d + j];
  }
  return s;
}

// Values of raw element e = threadIdx.x that the fused kernels request BEFORE the completion ticket (one hyperparameter set, P <= blockDim.x):
// the fit step then runs without a single load on its serial path (it was three dependent L2 round trips: Rprop state and raw value, the
// effective values for the chain rule, the raw values again for the new effective ones -- 3.8 us per iteration in the -DFGP_TIMING stamps).
struct FitPrefetch {
  double raw, prev, step;
};
__device__ __forceinline__ double* fit_raw_ptr(const FitLayout& c, int e) {
  const int n_ls = c.n_ls_b * c.n_ls_d;
  return e < c.n_scale ? c.raw_scale + e : (e < c.n_scale + n_ls ? c.raw_ls + (e - c.n_scale) : c.raw_noise + (e - c.n_scale - n_ls));
}
__device__ __forceinline__ FitPrefetch fit_prefetch(const FitLayout& c) {
  FitPrefetch f{0.0, 0.0, 0.0};
  const int e = threadIdx.x;
  if (e < c.P) {
    f.raw = __ldcg(fit_raw_ptr(c, e));
    f.prev = __ldcg(c.state + ST_HEADER + e);
    f.step = __ldcg(c.state + ST_HEADER + c.P + e);
  }
  return f;
}

// One fit() iteration's bookkeeping, executed by ONE CTA (any size >= 32 threads).  `out` is read with cache-global
// loads so the function can run in the tail of the kernel that produced it.  hdr: ST_HEADER doubles of shared memory.
// hdr_loaded: the caller already copied the state header into hdr (and synchronised); s_out: see raw_grad.
__device__ __forceinline__ void fit_step_device(const FitLayout& c, const double* out, double* red, double* hdr, int* flags, bool hdr_loaded = false,
                                                const double* s_out = nullptr, const FitPrefetch* pf = nullptr) {
  double* st = hdr;
  if (!hdr_loaded) {
    for (int e = threadIdx.x; e < ST_HEADER; e += blockDim.x) hdr[e] = __ldcg(c.state + e);
    __syncthreads();
  }
  if (st[ST_STOPPED] != 0.0) return;
  int& s_break = flags[0];
  int& s_newbest = flags[1];
  const int stride = c.d + 4;
  double v[2] = {0.0, 0.0};
  if (s_out) {
    v[0] = s_out[0];
    v[1] = s_out[1];
  } else {
    for (int b = threadIdx.x; b < c.B; b += blockDim.x) {
      v[0] += __ldcg(out + b * stride + 0);
      v[1] += __ldcg(out + b * stride + 1);
    }
    block_sum<2>(v, red);
  }
  if (threadIdx.x == 0) {
    const double wn = st[ST_WN], wl = st[ST_WL];
    const double loss = wn * v[0] + wl * v[1] + st[ST_HALF_CONST];
    const int i = (int)st[ST_ITER];
    int newbest = 0;
    if (loss < st[ST_BEST]) {
      st[ST_BEST] = loss;
      newbest = 1;
    }
    if (st[ST_SAVE] - loss > st[ST_LOGTOL]) {
      st[ST_WAIT] = 0.0;
      st[ST_SAVE] = st[ST_BEST];
    } else {
      st[ST_WAIT] += 1.0;
    }
    const int brk = (i == (int)st[ST_ITERATIONS]) || ((int)st[ST_WAIT] == (int)st[ST_STOP_WAIT]);
    st[ST_LAST_LOSS] = loss;
    st[ST_TERM1] = v[0];
    st[ST_TERM2] = 2.0 * wl * v[1];
    st[ST_LAST_ITER] = (double)i;
    if (c.loss_hist && i < (int)st[ST_HIST_CAP]) {
      c.loss_hist[3 * i + 0] = loss;
      c.loss_hist[3 * i + 1] = v[0];
      c.loss_hist[3 * i + 2] = 2.0 * wl * v[1];
    }
    if (brk) st[ST_STOPPED] = 1.0;
    st[ST_ITER] = (double)(i + 1);
    s_break = brk;
    s_newbest = newbest;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ST_HEADER; e += blockDim.x) c.state[e] = hdr[e];  // header back to global memory
  const int i = (int)st[ST_ITER] - 1;
  const int P = c.P;
  double* prev = c.state + ST_HEADER;
  double* step = prev + P;
  double* best = step + P;
  const int n_ls = c.n_ls_b * c.n_ls_d;
  if (pf) {
    // one set (B == 1, so n_scale == n_noise == n_ls_b == 1), s_out in shared memory, element e = threadIdx.x of the raw parameters in
    // registers: same arithmetic as the general path below, no loads
    const int e = threadIdx.x;
    if (e >= P) return;
    const int g = e < c.n_scale ? 0 : (e < c.n_scale + n_ls ? 1 : 2);
    const int le = g == 0 ? e : (g == 1 ? e - c.n_scale : e - c.n_scale - n_ls);
    const double ex = exp(pf->raw);  // exp(raw): the history row, and with tau the effective value of this iterate
    if (i < (int)st[ST_HIST_CAP]) {
      double* h = g == 0 ? c.scale_hist : (g == 1 ? c.ls_hist : c.noise_hist);
      const int width = g == 0 ? c.n_scale : (g == 1 ? n_ls : c.n_noise);
      if (h) h[(int64_t)i * width + le] = ex;
    }
    if (s_newbest) best[e] = pf->raw;
    if (s_break) return;
    if (!(g == 0 ? c.req_scale : (g == 1 ? c.req_ls : c.req_noise))) return;
    double grad;
    if (g == 0) {
      grad = s_out[3] * (c.tau * ex);
    } else if (g == 2) {
      grad = s_out[2] * (c.tau * ex);
    } else if (c.n_ls_d == 1) {
      grad = 0.0;
      for (int j = 0; j < c.d; ++j) grad += s_out[4 + j] * ex;
    } else {
      grad = 0.0;
      grad += s_out[4 + le] * ex;
    }
    const double etam = st[ST_ETAM], etap = st[ST_ETAP], smin = st[ST_SMIN], smax = st[ST_SMAX];
    const double sp = grad * pf->prev;
    double factor = 1.0;
    if (sp > 0.0) factor = etap;
    if (sp < 0.0) factor = etam;
    double ss = pf->step * factor;
    ss = ss < smin ? smin : (ss > smax ? smax : ss);
    step[e] = ss;
    if (sp < 0.0) grad = 0.0;
    const double sg = grad > 0.0 ? 1.0 : (grad < 0.0 ? -1.0 : 0.0);
    const double nr = pf->raw - sg * ss;
    *fit_raw_ptr(c, e) = nr;
    prev[e] = grad;
    const double en = exp(nr);  // the effective values the next iteration reads (write_effective)
    if (g == 0) {
      c.scale_B[0] = c.tau * en;
    } else if (g == 2) {
      c.noise_B[0] = c.tau * en;
    } else if (c.n_ls_d == 1) {
      for (int j = 0; j < c.d; ++j) c.ls_B[j] = en;
    } else {
      c.ls_B[le] = en;
    }
    return;
  }
  // history rows of the effective hyperparameters at this iterate (abstract_gp.py:285-288)
  if (i < (int)st[ST_HIST_CAP]) {
    if (c.scale_hist)
      for (int e = threadIdx.x; e < c.n_scale; e += blockDim.x) c.scale_hist[(int64_t)i * c.n_scale + e] = exp(c.raw_scale[e]);
    if (c.ls_hist)
      for (int e = threadIdx.x; e < n_ls; e += blockDim.x) c.ls_hist[(int64_t)i * n_ls + e] = exp(c.raw_ls[e]);
    if (c.noise_hist)
      for (int e = threadIdx.x; e < c.n_noise; e += blockDim.x) c.noise_hist[(int64_t)i * c.n_noise + e] = exp(c.raw_noise[e]);
  }
  if (s_newbest) {
    for (int e = threadIdx.x; e < P; e += blockDim.x) {
      const double* src = e < c.n_scale ? c.raw_scale + e : (e < c.n_scale + n_ls ? c.raw_ls + (e - c.n_scale) : c.raw_noise + (e - c.n_scale - n_ls));
      best[e] = *src;
    }
  }
  if (s_break) return;
  // Rprop
  const double etam = st[ST_ETAM], etap = st[ST_ETAP], smin = st[ST_SMIN], smax = st[ST_SMAX];
  for (int e = threadIdx.x; e < P; e += blockDim.x) {
    int g, le;
    double* raw;
    if (e < c.n_scale) {
      g = 0, le = e, raw = c.raw_scale + le;
      if (!c.req_scale) continue;
    } else if (e < c.n_scale + n_ls) {
      g = 1, le = e - c.n_scale, raw = c.raw_ls + le;
      if (!c.req_ls) continue;
    } else {
      g = 2, le = e - c.n_scale - n_ls, raw = c.raw_noise + le;
      if (!c.req_noise) continue;
    }
    double grad = raw_grad(c, out, s_out, g, le);
    const double sp = grad * prev[e];
    double factor = 1.0;
    if (sp > 0.0) factor = etap;
    if (sp < 0.0) factor = etam;
    double ss = step[e] * factor;
    ss = ss < smin ? smin : (ss > smax ? smax : ss);
    step[e] = ss;
    if (sp < 0.0) grad = 0.0;
    const double sg = grad > 0.0 ? 1.0 : (grad < 0.0 ? -1.0 : 0.0);
    *raw = *raw - sg * ss;
    prev[e] = grad;
  }
  __syncthreads();
  write_effective(c);
}


int make_layout(const fgp_fit_layout* in, FitLayout* c);  // fgp_fit.cu

}  // namespace fgp
