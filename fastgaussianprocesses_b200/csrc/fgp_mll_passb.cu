// K4 pass B (column transform + spectral epilogue + start of the backward transform) and the deterministic finalize.
#include "fgp_mll.cuh"

namespace fgp {

template <bool NET>
__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) mll_passB_kernel(const __grid_constant__ MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red[kRed];
  pdl_prologue(a);
  FGP_PSTAMP(21);
  if (fit_stop_flag(a) != 0.0) return;  // uniform over the CTA, before any barrier
  const int tile = blockIdx.x, b = blockIdx.y;
  passB_tile<NET>(a, smraw, red, tile, b);
  FGP_PSTAMP(24);
  if (a.has_fit && !a.want_grad) mll_fit_tail(a, b, a.ctasB, gridDim.y, true, red);
}

// stand-alone finalize (plain fgp_*_mll_grad calls)
__global__ void __launch_bounds__(256) mll_finalize_kernel(const __grid_constant__ MllArgs a) {
  __shared__ double red[kRed];
  finalize_set(a, blockIdx.x, red);
}

int pdl_mode() {
  // measured on B200 (profiles/README.md, snapshot j): 57.1 us per lattice fit iteration with every CTA triggering at entry (mode 1), 51.4
  // without the attribute -- the early-resident CTAs of the next kernel cost more than the hidden launch latency gains.  Mode 2 leaves the
  // trigger to the CTA's exit: the next kernel's CTAs only move into slots that are free for good.  Off unless FGP_PDL is set; read per call.
  return env_int("FGP_PDL", 0);
}

// The persistent cooperative kernel is OPT-IN (FGP_COOP=1).  Measured on B200 (profiles/README.md, round 2): its phases walk the same
// tiles as the three per-pass kernels, but the merged kernel needs more than the 128 registers that two 256-thread CTAs per SM allow
// (1.4 KB of spills; 0.2 KB with factorised twiddles), and the globaltimer stamps show that a pass-A tile alone takes 9.6-12 us
// even without spills -- launch gaps were never the bottleneck, the per-tile dependent chain is.  Best variant 51.1 us per
// iteration against 53.2 us for the three launches from a CUDA graph.  Read on every call (a getenv): tests switch routes.
bool coop_enabled() { return env_int("FGP_COOP", 0) != 0; }
int coop_max_ctas() { return env_int("FGP_COOP_CTAS", 0); }

#ifdef FGP_TIMING
long long* debug_stamp_buffer() {
  static long long* buf = nullptr;
  if (!buf && cudaMalloc(&buf, sizeof(long long) * kStampCtas * kStampSlots) == cudaSuccess) cudaMemset(buf, 0, sizeof(long long) * kStampCtas * kStampSlots);
  return buf;
}
#endif

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

#ifdef FGP_TIMING
// tools-only (-DFGP_TIMING builds): globaltimer stamps of the last mll_coop_kernel launch, (ctas, 16) int64 nanoseconds
extern "C" int fgp_debug_stamps(long long* out_host, int max_ctas) {
  const int n = max_ctas < fgp::kStampCtas ? max_ctas : fgp::kStampCtas;
  return cudaMemcpy(out_host, fgp::debug_stamp_buffer(), sizeof(long long) * n * fgp::kStampSlots, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
#endif
