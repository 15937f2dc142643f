// K4 pass B (column transform + spectral epilogue + start of the backward transform) and the deterministic finalize.
#include "fgp_mll.cuh"

namespace fgp {

template <bool NET>
__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) mll_passB_kernel(const __grid_constant__ MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red[kRed];
  if (fit_stopped(a)) return;
  const int b = blockIdx.y;
  const double noise = a.noise[b];
  const double dc = a.scale[b] * (double)a.n;  // the DC guess removed in pass A comes back in bin 0
  const int l1 = a.l1, l2 = a.l2, lntr = a.lntrB, LP = a.LPB;
  const int q0 = blockIdx.x << lntr;
  const int64_t boff = (int64_t)b * a.n;
  const double* ysq = a.ysq + boff + q0;
  const double wn = a.weights ? a.weights[2 * b] : 0.5, wl = a.weights ? a.weights[2 * b + 1] : 0.5;
  const int want_grad = a.want_grad;
  double s[3] = {0.0, 0.0, 0.0};
  if (NET) {
    double* sm = (double*)smraw;
    double* W = (double*)a.W + boff + q0;
    double* lamo = a.lam ? a.lam + boff + q0 : nullptr;
    block_wht_io<true>(sm, l2, lntr, LP, wht_sched_up(l2), [&](int tr, int r) -> double { return W[((int64_t)r << l1) + tr]; }, SmemTag{});
    __syncthreads();
    tile_map_r<true>(SmemR{sm, LP}, l2, lntr, [&](int tr, int r, double v) -> double {
      const int64_t k = ((int64_t)r << l1) + tr;
      double lam = v + noise;
      if (k + q0 == 0) lam += dc;
      if (lamo) lamo[k] = lam;
      return spectral_r(lam, ysq[k], wn, wl, s);
    });
    if (want_grad) {
      __syncthreads();
      block_wht_io<true>(sm, l2, lntr, LP, wht_sched_up(l2), SmemTag{}, [&](int tr, int r, double v) { W[((int64_t)r << l1) + tr] = v; });
    }
  } else {
    double2* sm = (double2*)smraw;
    double2* W = (double2*)a.W + boff + q0;
    double2* lamo = a.lam ? (double2*)a.lam + boff + q0 : nullptr;
    const FftTables T = a.T;
    block_fft_fwd_io<true>(sm, l2, lntr, LP, T.stage, [&](int tr, int r) -> double2 { return W[((int64_t)r << l1) + tr]; }, SmemTag{});
    __syncthreads();
    tile_map_c<true>(SmemC{sm, LP}, l2, lntr, [&](int tr, int r, double2 lam) -> double2 {
      const int64_t k = ((int64_t)r << l1) + tr;
      lam.x += noise;
      if (k + q0 == 0) lam.x += dc;
      if (lamo) lamo[k] = lam;
      return spectral_c(lam, ysq[k], wn, wl, s);
    });
    if (want_grad) {
      __syncthreads();
      block_fft_inv_io<true>(sm, l2, lntr, LP, T.stage, SmemTag{}, [&](int tr, int r, double2 v) {
        const double2 w = twiddle_n(T, brev_bits((uint32_t)r, l2) * (uint32_t)(q0 + tr));
        W[((int64_t)r << l1) + tr] = cmulc(w, v);
      });
    }
  }
  __syncthreads();
  reduce_store<3>(s, 3, red, a.partB + ((int64_t)b * a.ctasB + blockIdx.x) * 3);
  if (a.has_fit && !want_grad) mll_fit_tail(a, b, a.ctasB, gridDim.y, true, red);
}

// stand-alone finalize (plain fgp_*_mll_grad calls)
__global__ void __launch_bounds__(256) mll_finalize_kernel(const __grid_constant__ MllArgs a) {
  __shared__ double red[kRed];
  finalize_set(a, blockIdx.x, red);
}

int launch_mll_passB(const MllArgs& a, const PassGeom& g, int B, bool net, cudaStream_t st) {
  int rc;
  if (net) {
    if ((rc = set_smem_attr(mll_passB_kernel<true>, g.smemB))) return rc;
    mll_passB_kernel<true><<<dim3(a.ctasB, B), g.threadsB, g.smemB, st>>>(a);
  } else {
    if ((rc = set_smem_attr(mll_passB_kernel<false>, g.smemB))) return rc;
    mll_passB_kernel<false><<<dim3(a.ctasB, B), g.threadsB, g.smemB, st>>>(a);
  }
  FGP_LAUNCH_NAMED("mll_passB", st);
  return FGP_OK;
}

int launch_mll_finalize(const MllArgs& a, int B, cudaStream_t st) {
  mll_finalize_kernel<<<B, 256, 0, st>>>(a);
  FGP_LAUNCH_NAMED("mll_finalize", st);
  return FGP_OK;
}

}  // namespace fgp
