// K4 (digital net): fused FWHT eigen-solve + MLL + gradients.  Templates live in fgp_mll.cuh.
#include "fgp_mll.cuh"

extern "C" {

int fgp_dnb2_mll_grad(const int64_t* xb_dev, int64_t n, int d, const int* alpha_host, int t, int B, const double* ysq_dev,
                      const double* scale_dev, const double* ls_dev, const double* noise_dev, const double* weights_dev, void* workspace_dev,
                      double* lam_dev, double* out_dev, int want_grad, fgp_stream_t stream) {
  return fgp::mll_common<true>(nullptr, xb_dev, n, d, alpha_host, t, B, ysq_dev, scale_dev, ls_dev, noise_dev, weights_dev, nullptr,
                               workspace_dev, lam_dev, out_dev, want_grad, stream);
}

}  // extern "C"
