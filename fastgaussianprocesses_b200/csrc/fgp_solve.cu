// K^-1 y = T^-1( T(y) / lam ) for R right-hand sides sharing one spectrum (util.py:338-344, single task).
#include "fgp_transform.cuh"

namespace fgp {

__global__ void __launch_bounds__(256) divide_c_kernel(double2* __restrict__ v, const double2* __restrict__ lam,
                                                       int64_t R, int64_t n) {
  const int64_t total = R * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const double2 l = lam[e % n];
    const double inv = 1.0 / fma(l.x, l.x, l.y * l.y);
    const double2 x = v[e];
    v[e] = make_double2(fma(x.x, l.x, x.y * l.y) * inv, fma(x.y, l.x, -x.x * l.y) * inv);  // x * conj(l) / |l|^2
  }
}
__global__ void __launch_bounds__(256) divide_r_kernel(double* __restrict__ v, const double* __restrict__ lam, int64_t R,
                                                       int64_t n) {
  const int64_t total = R * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    v[e] = v[e] / lam[e % n];
}
__global__ void __launch_bounds__(256) real_part_kernel(const double2* __restrict__ v, double* __restrict__ out,
                                                        int64_t total) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = v[e].x;
}

}  // namespace fgp

extern "C" {

int fgp_gram_solve(int family, const double* y_dev, double* out_dev, int64_t R, int64_t n, const double* lam_dev,
                   const void* table_dev, void* work_dev, fgp_stream_t stream) {
  using namespace fgp;
  FGP_REQUIRE(y_dev && out_dev && lam_dev, "gram_solve: null pointer");
  FGP_REQUIRE(R >= 0 && is_pow2(n), "gram_solve: bad R/n");
  if (R == 0) return FGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = R * n;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  int rc;
  if (family == 0) {
    FGP_REQUIRE(table_dev && work_dev, "gram_solve: lattice needs a twiddle table and a complex workspace");
    if ((rc = fgp_fftbr_r2c(y_dev, (double*)work_dev, R, n, table_dev, stream))) return rc;
    divide_c_kernel<<<(unsigned)blocks, 256, 0, st>>>((double2*)work_dev, (const double2*)lam_dev, R, n);
    FGP_LAUNCH_CHECK();
    if ((rc = fgp_ifftbr_c2c((const double*)work_dev, (double*)work_dev, R, n, table_dev, stream))) return rc;
    real_part_kernel<<<(unsigned)blocks, 256, 0, st>>>((const double2*)work_dev, out_dev, total);
    FGP_LAUNCH_CHECK();
  } else {
    if ((rc = fgp_fwht(y_dev, out_dev, R, n, stream))) return rc;
    divide_r_kernel<<<(unsigned)blocks, 256, 0, st>>>(out_dev, lam_dev, R, n);
    FGP_LAUNCH_CHECK();
    if ((rc = fgp_fwht(out_dev, out_dev, R, n, stream))) return rc;
  }
  return FGP_OK;
}

}  // extern "C"
