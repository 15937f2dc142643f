// K4 (lattice): fused FFT-BRO eigen-solve + MLL + gradients.  Templates live in fgp_mll.cuh.
#include "fgp_mll.cuh"

extern "C" {

size_t fgp_mll_workspace_bytes(int family, int64_t n, int d, int B) {
  using namespace fgp;
  if (!is_pow2(n) || B < 1 || d < 1) return 0;
  const bool net = family != 0;
  const PassGeom g = make_geom(n, !net);
  if (g.l2 == 0) return 256;
  return align256((size_t)B * n * (net ? sizeof(double) : sizeof(double2))) +
         align256((size_t)B * g.ctasB * 3 * sizeof(double)) + align256((size_t)B * g.ctasA * (d + 1) * sizeof(double));
}

int fgp_lattice_mll_grad(const double* x_dev, int64_t n, int d, const int* alpha_host, int B, const double* ysq_dev,
                         const double* scale_dev, const double* ls_dev, const double* noise_dev, const double* weights_dev, const void* table_dev,
                         void* workspace_dev, double* lam_dev, double* out_dev, int want_grad, fgp_stream_t stream) {
  return fgp::mll_common<false>(nullptr, x_dev, n, d, alpha_host, 0, B, ysq_dev, scale_dev, ls_dev, noise_dev, weights_dev, table_dev,
                                workspace_dev, lam_dev, out_dev, want_grad, stream);
}

int fgp_lattice_mll_grad_z(const uint64_t* z_host, int64_t n, int d, const int* alpha_host, int B, const double* ysq_dev,
                           const double* scale_dev, const double* ls_dev, const double* noise_dev, const double* weights_dev, const void* table_dev,
                           void* workspace_dev, double* lam_dev, double* out_dev, int want_grad, fgp_stream_t stream) {
  FGP_REQUIRE(z_host, "mll_grad_z: null generating vector");
  return fgp::mll_common<false>(z_host, nullptr, n, d, alpha_host, 0, B, ysq_dev, scale_dev, ls_dev, noise_dev, weights_dev, table_dev,
                                workspace_dev, lam_dev, out_dev, want_grad, stream);
}

}  // extern "C"
