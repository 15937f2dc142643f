// K4 pass B (column transform + spectral epilogue + start of the backward transform) and the deterministic finalize.
#include "fgp_mll.cuh"

namespace fgp {

template <bool NET>
__global__ void __launch_bounds__(512, 1) mll_passB_kernel(const __grid_constant__ MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red[kRed];
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
    double* W = (double*)a.W + boff + q0;
    double* lamo = a.lam ? a.lam + boff + q0 : nullptr;
    auto gld = [&](int tr, int r) -> double { return W[((int64_t)r << l1) + tr]; };
    auto mid = [&](int tr, int r, double v) -> double {
      const int64_t k = ((int64_t)r << l1) + tr;
      double lam = v + noise;
      if (k + q0 == 0) lam += dc;
      if (lamo) lamo[k] = lam;
      return spectral_r(lam, ysq[k], wn, wl, s);
    };
    auto gst = [&](int tr, int r, double v) {
      if (want_grad) W[((int64_t)r << l1) + tr] = v;
    };
    block_wht_fwd_mid_inv_io<true>((double*)smraw, l2, lntr, LP, gld, mid, gst);
  } else {
    double2* W = (double2*)a.W + boff + q0;
    double2* lamo = a.lam ? (double2*)a.lam + boff + q0 : nullptr;
    const FftTables T = a.T;
    auto gld = [&](int tr, int r) -> double2 { return W[((int64_t)r << l1) + tr]; };
    auto mid = [&](int tr, int r, double2 lam) -> double2 {
      const int64_t k = ((int64_t)r << l1) + tr;
      lam.x += noise;
      if (k + q0 == 0) lam.x += dc;
      if (lamo) lamo[k] = lam;
      return spectral_c(lam, ysq[k], wn, wl, s);
    };
    auto gst = [&](int tr, int r, double2 v) {
      if (!want_grad) return;
      const double2 w = twiddle_n(T, brev_bits((uint32_t)r, l2) * (uint32_t)(q0 + tr));
      W[((int64_t)r << l1) + tr] = cmulc(w, v);
    };
    block_fft_fwd_mid_inv_io<true>((double2*)smraw, l2, lntr, LP, T.stage, gld, mid, gst);
  }
  __syncthreads();
  reduce_store<3>(s, 3, red, a.partB + ((int64_t)b * a.ctasB + blockIdx.x) * 3);
}

// finalize: deterministic reduction of the per-CTA partial sums; thread (j, lane-group) layout keeps it one pass
__global__ void __launch_bounds__(256) mll_finalize_kernel(const __grid_constant__ MllArgs a) {
  __shared__ double red[kRed];
  const int b = blockIdx.x;
  const int d = a.d;
  double* out = a.out + (int64_t)b * (d + 4);
  {
    double s[3] = {0.0, 0.0, 0.0};
    const double* p = a.partB + (int64_t)b * a.ctasB * 3;
    for (int c = threadIdx.x; c < a.ctasB; c += blockDim.x) {
      s[0] += p[c * 3 + 0];
      s[1] += p[c * 3 + 1];
      s[2] += p[c * 3 + 2];
    }
    reduce_store<3>(s, 3, red, out);
  }
  if (!a.want_grad) return;
  // warp w reduces components w, w+8, ... over all pass-C CTAs (fixed order => deterministic)
  const double* p = a.partC + (int64_t)b * a.ctasA * (d + 1);
  const double inv_scale = 1.0 / a.scale[b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int j = warp; j <= d; j += nwarp) {
    double v = 0.0;
    for (int c = lane; c < a.ctasA; c += 32) v += p[(int64_t)c * (d + 1) + j];
    v = warp_sum(v);
    if (lane == 0) out[3 + j] = j == 0 ? v * inv_scale : v;
  }
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
