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
#include "fgp_fit.cuh"

namespace fgp {

struct FitHeader {
  double v[ST_HEADER];
};

// the three raw parameter groups of a layout, in state order (scale, lengthscales, noise)
__device__ __forceinline__ double* raw_slot(const FitLayout& c, double* s0, double* s1, double* s2, int e) {
  const int n_ls = c.n_ls_b * c.n_ls_d;
  return e < c.n_scale ? s0 + e : (e < c.n_scale + n_ls ? s1 + (e - c.n_scale) : s2 + (e - c.n_scale - n_ls));
}

__global__ void __launch_bounds__(256) fit_init_kernel(FitLayout c, FitHeader h, FitParamPtrs src) {
  double* st = c.state;
  if (threadIdx.x < ST_HEADER) st[threadIdx.x] = h.v[threadIdx.x];  // options by value: no host copy, no synchronisation
  if (src.scale) {  // the caller's parameter storages -> the layout's (one launch instead of three device-to-device copies)
    for (int e = threadIdx.x; e < c.P; e += blockDim.x) *raw_slot(c, c.raw_scale, c.raw_ls, c.raw_noise, e) = *raw_slot(c, src.scale, src.ls, src.noise, e);
  }
  __syncthreads();
  double* prev = st + ST_HEADER;
  double* step = prev + c.P;
  for (int e = threadIdx.x; e < c.P; e += blockDim.x) {
    prev[e] = 0.0;
    step[e] = st[ST_LR];
  }
  for (int e = threadIdx.x; e <= c.B; e += blockDim.x) c.tickets[e] = 0u;
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

__global__ void __launch_bounds__(256) fit_step_kernel(FitLayout c, const double* __restrict__ out) {
  __shared__ double red[32 * 4];
  __shared__ double hdr[ST_HEADER];
  __shared__ int flags[2];
  fit_step_device(c, out, red, hdr, flags);
}

// copy the best iterate back into the parameters (abstract_gp.py:297-298) and refresh the effective values
__global__ void __launch_bounds__(256) fit_finish_kernel(FitLayout c, FitParamPtrs dst) {
  double* st = c.state;
  const int P = c.P;
  const double* best = st + ST_HEADER + 2 * P;
  const bool have = st[ST_BEST] < INFINITY;
  for (int e = threadIdx.x; e < P; e += blockDim.x) {
    double* own = raw_slot(c, c.raw_scale, c.raw_ls, c.raw_noise, e);
    if (have) *own = best[e];
    if (dst.scale) *raw_slot(c, dst.scale, dst.ls, dst.noise, e) = *own;  // and out to the caller's parameter storages
  }
  __syncthreads();
  write_effective(c);
}

int make_layout(const fgp_fit_layout* in, FitLayout* c) {
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
  c->tickets = (unsigned int*)(in->state + ST_HEADER + 3 * (size_t)c->P);
  return FGP_OK;
}

}  // namespace fgp

extern "C" {

size_t fgp_fit_state_doubles(int n_raw_params, int B) {
  return (size_t)fgp::ST_HEADER + 3 * (size_t)(n_raw_params > 0 ? n_raw_params : 0) + ((size_t)(B > 0 ? B : 0) + 2) / 2 + 1;
}

int fgp_fit_init(const fgp_fit_layout* layout, const fgp_fit_options* opt, fgp_stream_t stream) {
  return fgp_fit_init_from(layout, opt, nullptr, nullptr, nullptr, stream);
}

int fgp_fit_init_from(const fgp_fit_layout* layout, const fgp_fit_options* opt, const double* raw_scale_src, const double* raw_ls_src,
                      const double* raw_noise_src, fgp_stream_t stream) {
  fgp::FitLayout c;
  int rc = fgp::make_layout(layout, &c);
  if (rc) return rc;
  FGP_REQUIRE(opt, "fit_init: null options");
  const int nsrc = (raw_scale_src != nullptr) + (raw_ls_src != nullptr) + (raw_noise_src != nullptr);
  FGP_REQUIRE(nsrc == 0 || nsrc == 3, "fit_init_from: give all three parameter sources or none");
  const fgp::FitParamPtrs src{const_cast<double*>(raw_scale_src), const_cast<double*>(raw_ls_src), const_cast<double*>(raw_noise_src)};
  FGP_REQUIRE(opt->iterations >= 0 && opt->stop_wait > 0 && opt->lr > 0.0, "fit_init: bad options");
  fgp::FitHeader hdr;
  double* h = hdr.v;
  memset(h, 0, sizeof(hdr));
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
  fgp::fit_init_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(c, hdr, src);
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
  return fgp_fit_finish_to(layout, nullptr, nullptr, nullptr, stream);
}

int fgp_fit_finish_to(const fgp_fit_layout* layout, double* raw_scale_dst, double* raw_ls_dst, double* raw_noise_dst, fgp_stream_t stream) {
  fgp::FitLayout c;
  int rc = fgp::make_layout(layout, &c);
  if (rc) return rc;
  const int ndst = (raw_scale_dst != nullptr) + (raw_ls_dst != nullptr) + (raw_noise_dst != nullptr);
  FGP_REQUIRE(ndst == 0 || ndst == 3, "fit_finish_to: give all three parameter destinations or none");
  const fgp::FitParamPtrs dst{raw_scale_dst, raw_ls_dst, raw_noise_dst};
  fgp::fit_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(c, dst);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

}  // extern "C"
