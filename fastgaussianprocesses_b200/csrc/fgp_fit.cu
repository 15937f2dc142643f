// K4b: device-side fit() bookkeeping -- loss assembly, early-stop state machine, best-iterate snapshot, Rprop update and
// the raw -> effective hyperparameter transform, in ONE single-CTA kernel per iteration, so that a whole fit() loop
// (abstract_gp.py:236-298 with the default optimiser abstract_fast_gp.py:53-57 and the default (log, exp) transforms
// fast_gp_lattice.py:137-139) can be captured in a CUDA graph with no host round trip per iteration.
//
// Reference semantics reproduced exactly:
//   loss = 1/2 (sum_b norm_b + d_out/B sum_b logdet_b + d_out n log 2 pi)            abstract_gp.py:253-260
//   best / save / wait state machine and break condition                             abstract_gp.py:276-283
//   history rows written BEFORE the update (iteration i sees the parameters it was evaluated at)   :284-289
//   torch.optim.Rprop (etas, step sizes, sign rule, zeroed gradient on a sign flip)  torch/optim/rprop.py
//   parameters restored to the best iterate at the end                               abstract_gp.py:297-298
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

__global__ void __launch_bounds__(256) fit_init_kernel(FitLayout c) {
  double* st = c.state;
  double* prev = st + ST_HEADER;
  double* step = prev + c.P;
  for (int e = threadIdx.x; e < c.P; e += blockDim.x) {
    prev[e] = 0.0;
    step[e] = st[ST_LR];
  }
  if (threadIdx.x == 0) {
    st[ST_BEST] = INFINITY;
    st[ST_SAVE] = INFINITY;
    st[ST_WAIT] = 0.0;
    st[ST_ITER] = 0.0;
    st[ST_STOPPED] = 0.0;
    st[ST_LAST_ITER] = -1.0;
  }
  write_effective(c);
}

// gradient w.r.t. raw element e of parameter group g (0 scale, 1 lengthscales, 2 noise); theta = exp(raw) => dtheta/draw = theta
__device__ __forceinline__ double raw_grad(const FitLayout& c, const double* __restrict__ out, int g, int e) {
  const int stride = c.d + 4;
  double s = 0.0;
  if (g == 0) {
    if (c.n_scale == 1) {
      for (int b = 0; b < c.B; ++b) s += out[b * stride + 3] * c.scale_B[b];
    } else {
      s = out[e * stride + 3] * c.scale_B[e];
    }
  } else if (g == 2) {
    if (c.n_noise == 1) {
      for (int b = 0; b < c.B; ++b) s += out[b * stride + 2] * c.noise_B[b];
    } else {
      s = out[e * stride + 2] * c.noise_B[e];
    }
  } else {
    const int eb = e / c.n_ls_d, ej = e - eb * c.n_ls_d;
    const int b0 = c.n_ls_b == 1 ? 0 : eb, b1 = c.n_ls_b == 1 ? c.B : eb + 1;
    const int j0 = c.n_ls_d == 1 ? 0 : ej, j1 = c.n_ls_d == 1 ? c.d : ej + 1;
    for (int b = b0; b < b1; ++b)
      for (int j = j0; j < j1; ++j) s += out[b * stride + 4 + j] * c.ls_B[b * c.d + j];
  }
  return s;
}

__global__ void __launch_bounds__(256) fit_step_kernel(FitLayout c, const double* __restrict__ out) {
  __shared__ double red[32 * 4];
  __shared__ int s_break, s_newbest;
  double* st = c.state;
  if (st[ST_STOPPED] != 0.0) return;
  const int stride = c.d + 4;
  double v[2] = {0.0, 0.0};
  for (int b = threadIdx.x; b < c.B; b += blockDim.x) {
    v[0] += out[b * stride + 0];
    v[1] += out[b * stride + 1];
  }
  block_sum<2>(v, red);
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
  const int i = (int)st[ST_ITER] - 1;
  const int P = c.P;
  double* prev = st + ST_HEADER;
  double* step = prev + P;
  double* best = step + P;
  const int n_ls = c.n_ls_b * c.n_ls_d;
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
    double grad = raw_grad(c, out, g, le);
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

// copy the best iterate back into the parameters (abstract_gp.py:297-298) and refresh the effective values
__global__ void __launch_bounds__(256) fit_finish_kernel(FitLayout c) {
  double* st = c.state;
  const int P = c.P;
  const double* best = st + ST_HEADER + 2 * P;
  const int n_ls = c.n_ls_b * c.n_ls_d;
  if (st[ST_BEST] < INFINITY) {
    for (int e = threadIdx.x; e < P; e += blockDim.x) {
      double* dst = e < c.n_scale ? c.raw_scale + e : (e < c.n_scale + n_ls ? c.raw_ls + (e - c.n_scale) : c.raw_noise + (e - c.n_scale - n_ls));
      *dst = best[e];
    }
  }
  __syncthreads();
  write_effective(c);
}

static int make_layout(const fgp_fit_layout* in, FitLayout* c) {
  FGP_REQUIRE(in, "fit: null layout");
  FGP_REQUIRE(in->B >= 1 && in->B <= 65535 && in->d >= 1 && in->d <= FGP_MAX_D, "fit: bad B/d");
  FGP_REQUIRE((in->n_scale == 1 || in->n_scale == in->B) && (in->n_noise == 1 || in->n_noise == in->B) &&
                  (in->n_ls_b == 1 || in->n_ls_b == in->B) && (in->n_ls_d == 1 || in->n_ls_d == in->d),
              "fit: raw parameter batch sizes must be 1 or B (lengthscale width 1 or d)");
  FGP_REQUIRE(in->raw_scale && in->raw_ls && in->raw_noise && in->scale_B && in->ls_B && in->noise_B && in->state, "fit: null pointer");
  c->B = in->B;
  c->d = in->d;
  c->n_scale = in->n_scale;
  c->n_ls_b = in->n_ls_b;
  c->n_ls_d = in->n_ls_d;
  c->n_noise = in->n_noise;
  c->req_scale = in->req_scale;
  c->req_ls = in->req_ls;
  c->req_noise = in->req_noise;
  c->P = in->n_scale + in->n_ls_b * in->n_ls_d + in->n_noise;
  c->tau = in->tau;
  c->raw_scale = in->raw_scale;
  c->raw_ls = in->raw_ls;
  c->raw_noise = in->raw_noise;
  c->scale_B = in->scale_B;
  c->ls_B = in->ls_B;
  c->noise_B = in->noise_B;
  c->state = in->state;
  c->loss_hist = in->loss_hist;
  c->scale_hist = in->scale_hist;
  c->ls_hist = in->ls_hist;
  c->noise_hist = in->noise_hist;
  return FGP_OK;
}

}  // namespace fgp

extern "C" {

size_t fgp_fit_state_doubles(int n_raw_params) { return (size_t)fgp::ST_HEADER + 3 * (size_t)(n_raw_params > 0 ? n_raw_params : 0); }

int fgp_fit_init(const fgp_fit_layout* layout, const fgp_fit_options* opt, fgp_stream_t stream) {
  fgp::FitLayout c;
  int rc = fgp::make_layout(layout, &c);
  if (rc) return rc;
  FGP_REQUIRE(opt, "fit_init: null options");
  FGP_REQUIRE(opt->iterations >= 0 && opt->stop_wait > 0 && opt->lr > 0.0, "fit_init: bad options");
  double h[fgp::ST_HEADER];
  memset(h, 0, sizeof(h));
  h[fgp::ST_ITERATIONS] = (double)opt->iterations;
  h[fgp::ST_STOP_WAIT] = (double)opt->stop_wait;
  h[fgp::ST_LOGTOL] = opt->logtol;
  h[fgp::ST_HALF_CONST] = opt->half_const;
  h[fgp::ST_WN] = opt->wn;
  h[fgp::ST_WL] = opt->wl;
  h[fgp::ST_LR] = opt->lr;
  h[fgp::ST_ETAM] = opt->etaminus;
  h[fgp::ST_ETAP] = opt->etaplus;
  h[fgp::ST_SMIN] = opt->step_min;
  h[fgp::ST_SMAX] = opt->step_max;
  h[fgp::ST_HIST_CAP] = (double)opt->hist_capacity;
  // the header is tiny: a by-value kernel parameter would also do, but a stream-ordered copy keeps the graph generic
  FGP_CUDA(cudaMemcpyAsync(c.state, h, sizeof(h), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  FGP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // h lives on this stack frame
  fgp::fit_init_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(c);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

int fgp_fit_step(const fgp_fit_layout* layout, const double* out_dev, fgp_stream_t stream) {
  fgp::FitLayout c;
  int rc = fgp::make_layout(layout, &c);
  if (rc) return rc;
  FGP_REQUIRE(out_dev, "fit_step: null mll output");
  fgp::fit_step_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(c, out_dev);
  FGP_LAUNCH_NAMED("fit_step", (cudaStream_t)stream);
  return FGP_OK;
}

int fgp_fit_finish(const fgp_fit_layout* layout, fgp_stream_t stream) {
  fgp::FitLayout c;
  int rc = fgp::make_layout(layout, &c);
  if (rc) return rc;
  fgp::fit_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(c);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

}  // extern "C"
