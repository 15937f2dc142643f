// K4 pass B (column transform + spectral epilogue + start of the backward transform) and the deterministic finalize.
#include "fgp_mll.cuh"

namespace fgp {

template <bool NET>
__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) mll_passB_kernel(const __grid_constant__ MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red[kRed];
  pdl_prologue();
  const double stop_flag = fit_stop_flag(a);
  const int b = blockIdx.y;
  const double noise = a.noise[b];
  const double dc = a.scale[b] * (double)a.n;  // the DC guess removed in pass A comes back in bin 0
  if (stop_flag != 0.0) return;
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
    if (a.hs) {
      // half-spectrum mode (fgp_mll.cuh): block rows of residue class r > L2/2 were not written by pass A, they are the
      // conjugates of class L2 - r; columns q and L1 - q carry the same eigenvalues, so column q counts twice
      const uint32_t L2 = 1u << l2, half2 = L2 >> 1;
      const int half1 = 1 << (l1 - 1);
      block_fft_fwd_io<true>(sm, l2, lntr, LP, T.stage, [&](int tr, int r) -> double2 {
        const uint32_t res = brev_bits((uint32_t)r, l2);
        const bool mir = res > half2;
        const uint32_t row = mir ? brev_bits(L2 - res, l2) : (uint32_t)r;
        const double2 v = W[((int64_t)row << l1) + tr];
        return make_double2(v.x, mir ? -v.y : v.y);
      }, SmemTag{});
      __syncthreads();
      tile_map_c<true>(SmemC{sm, LP}, l2, lntr, [&](int tr, int r, double2 lam) -> double2 {
        const int64_t k = ((int64_t)r << l1) + tr;
        const int q = q0 + tr;
        lam.x += noise;
        if (k + q0 == 0) lam.x += dc;
        double t3[3] = {0.0, 0.0, 0.0};
        const double2 G = spectral_c(lam, ysq[k], wn, wl, t3);
        const double cw = (q == 0 || q == half1) ? 1.0 : (q < half1 ? 2.0 : 0.0);
        s[0] = fma(cw, t3[0], s[0]);
        s[1] = fma(cw, t3[1], s[1]);
        s[2] = fma(cw, t3[2], s[2]);
        // lam is real in exact arithmetic; its computed imaginary part is round-off, but dL/dIm(lam) ~ Im(lam) |y~|^2 / lam^3 is
        // not small where lam is.  A Hermitian (instead of real) dL/dlam back-transforms to a real but not EVEN sequence,
        // and the odd part only cancels in a sum over all points -- pass C sums half of them twice.  Keep the real part.
        return make_double2(G.x, 0.0);
      });
      if (want_grad && (q0 == 0 || q0 == half1)) {
        // the two self-mirrored columns hold both members of every pair (k, n-k): make them exactly equal as well
        __syncthreads();
        double2* col = sm;  // tr == 0
        const int L2i = 1 << l2;
        for (int sidx = threadIdx.x; sidx < (L2i >> 1); sidx += blockDim.x) {
          const int s1 = q0 == 0 ? sidx : sidx;
          const int s2 = q0 == 0 ? (L2i - sidx) & (L2i - 1) : L2i - 1 - sidx;
          if (s1 != s2) {
            const int i1 = padidx<kPSC>(s1), i2 = padidx<kPSC>(s2);
            const double av = 0.5 * (col[i1].x + col[i2].x);
            col[i1].x = av;
            col[i2].x = av;
          }
        }
      }
      if (want_grad) {
        __syncthreads();
        block_fft_inv_io<true>(sm, l2, lntr, LP, T.stage, SmemTag{}, [&](int tr, int r, double2 v) {
          const uint32_t res = brev_bits((uint32_t)r, l2);
          if (res > half2) return;  // pass C never reads the mirrored block rows
          const double2 w = twiddle_n(T, res * (uint32_t)(q0 + tr));
          W[((int64_t)r << l1) + tr] = cmulc(w, v);
        });
      }
    } else {
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

bool pdl_enabled() {
  // measured on B200 (profiles/README.md, snapshot j): 57.1 us per lattice fit iteration with the attribute, 51.4 without -- the
  // early-resident CTAs of the next kernel cost more than the hidden launch latency gains.  Off unless FGP_PDL=1.
  static const bool on = env_int("FGP_PDL", 0) != 0;
  return on;
}

int launch_mll_passB(const MllArgs& a, const PassGeom& g, int B, bool net, cudaStream_t st) {
  int rc;
  if (net) {
    if ((rc = set_smem_attr(mll_passB_kernel<true>, g.smemB))) return rc;
    launch_chain(mll_passB_kernel<true>, dim3(a.ctasB, B), dim3(g.threadsB), g.smemB, st, a);
  } else {
    if ((rc = set_smem_attr(mll_passB_kernel<false>, g.smemB))) return rc;
    launch_chain(mll_passB_kernel<false>, dim3(a.ctasB, B), dim3(g.threadsB), g.smemB, st, a);
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
