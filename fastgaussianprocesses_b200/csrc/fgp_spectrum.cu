// K3b: the spectrum of the data in ONE call -- ytilde = ft(y), mean-stabilised as the reference does it (abstract_fast_gp.py:197-212:
// transform y - mean(y), then add mean(y) sqrt(n) back to the zero frequency), and |ytilde|^2 summed over the batch rows that share a
// hyperparameter set (what every MLL iteration reads, util.py:364-370).  The host path through torch took nine small launches (mean,
// subtract, two transform passes, add-back, square, two sums, copy), ~0.25 ms of host time in front of the first fit iteration; here it
// is five launches behind one C call.  Row sums are two-stage with a fixed order: the result does not depend on scheduling.
#include "fgp_common.cuh"

namespace fgp {

constexpr int kMeanChunks = 64;

// stage 1: partial[r][c] = sum of chunk c of row r
__global__ void __launch_bounds__(256) row_partial_sums_kernel(const double* __restrict__ y, int64_t n, double* __restrict__ partial) {
  __shared__ double red[32];
  const int64_t r = blockIdx.y;
  const int64_t len = (n + kMeanChunks - 1) / kMeanChunks;
  const int64_t i0 = (int64_t)blockIdx.x * len, i1 = min(n, i0 + len);
  double s[1] = {0.0};
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) s[0] += y[r * n + i];
  block_sum<1>(s, red);
  if (threadIdx.x == 0) partial[r * kMeanChunks + blockIdx.x] = s[0];
}

__device__ __forceinline__ double row_mean(const double* __restrict__ partial, int64_t r, int64_t n) {
  double s = 0.0;
  for (int c = 0; c < kMeanChunks; ++c) s += partial[r * kMeanChunks + c];
  return s / (double)n;
}

// stage 2: centred copy of the rows
__global__ void __launch_bounds__(256) center_rows_kernel(const double* __restrict__ y, int64_t n, const double* __restrict__ partial,
                                                           double* __restrict__ yc) {
  __shared__ double mean;
  const int64_t r = blockIdx.y;
  if (threadIdx.x == 0) mean = row_mean(partial, r, n);
  __syncthreads();
  const double m = mean;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) yc[r * n + i] = y[r * n + i] - m;
}

// after the transform: the zero frequency gets mean sqrt(n) back (written to ytilde), then ysq[b][k] = sum_l |ytilde[l B + b][k]|^2
template <bool CPLX>
__global__ void __launch_bounds__(256) spectrum_finish_kernel(double* __restrict__ yt, int64_t lead, int64_t B, int64_t n,
                                                               const double* __restrict__ partial, double* __restrict__ ysq) {
  const int64_t b = blockIdx.y;
  const double rootn = sqrt((double)n);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t l = 0; l < lead; ++l) {
      const int64_t r = l * B + b;
      if (CPLX) {
        double2 v = ((double2*)yt)[r * n + k];
        if (k == 0) {
          v.x += row_mean(partial, r, n) * rootn;
          ((double2*)yt)[r * n] = v;
        }
        s += fma(v.x, v.x, v.y * v.y);
      } else {
        double v = yt[r * n + k];
        if (k == 0) {
          v += row_mean(partial, r, n) * rootn;
          yt[r * n] = v;
        }
        s = fma(v, v, s);
      }
    }
    ysq[b * n + k] = s;
  }
}

}  // namespace fgp

extern "C" size_t fgp_data_spectrum_workspace_bytes(int64_t rows, int64_t n) {
  if (rows < 0 || n < 0) return 0;
  return (size_t)rows * (size_t)n * sizeof(double) + (size_t)rows * fgp::kMeanChunks * sizeof(double);
}

extern "C" int fgp_data_spectrum(int family, const double* y_dev, int64_t rows, int64_t B, int64_t n, const void* table_dev, double* ytilde_dev,
                                 double* ysq_dev, double* work_dev, fgp_stream_t stream) {
  using namespace fgp;
  FGP_REQUIRE(family == 0 || family == 1, "data_spectrum: family must be 0 (lattice) or 1 (digital net), got %d", family);
  FGP_REQUIRE(y_dev && ytilde_dev && ysq_dev && work_dev, "data_spectrum: null pointer");
  FGP_REQUIRE(rows >= 1 && B >= 1 && rows % B == 0 && B <= 65535 && rows <= 65535,
              "data_spectrum: need 1 <= B <= rows <= 65535 with B dividing rows (got rows=%lld B=%lld)", (long long)rows, (long long)B);
  FGP_REQUIRE(is_pow2(n) && n >= 2, "data_spectrum: n must be a power of two >= 2 (got %lld)", (long long)n);
  FGP_REQUIRE(family == 1 || table_dev, "data_spectrum: the lattice transform needs its twiddle table");
  cudaStream_t st = (cudaStream_t)stream;
  double* yc = work_dev;
  double* partial = work_dev + rows * n;
  row_partial_sums_kernel<<<dim3(kMeanChunks, (unsigned)rows), 256, 0, st>>>(y_dev, n, partial);
  FGP_LAUNCH_NAMED("row_partial_sums", st);
  const unsigned gx = (unsigned)std::min<int64_t>((n + 255) / 256, 1184);
  center_rows_kernel<<<dim3(gx, (unsigned)rows), 256, 0, st>>>(y_dev, n, partial, yc);
  FGP_LAUNCH_NAMED("center_rows", st);
  int rc = family == 0 ? fgp_fftbr_r2c(yc, ytilde_dev, rows, n, table_dev, stream) : fgp_fwht(yc, ytilde_dev, rows, n, stream);
  if (rc) return rc;
  if (family == 0)
    spectrum_finish_kernel<true><<<dim3(gx, (unsigned)B), 256, 0, st>>>(ytilde_dev, rows / B, B, n, partial, ysq_dev);
  else
    spectrum_finish_kernel<false><<<dim3(gx, (unsigned)B), 256, 0, st>>>(ytilde_dev, rows / B, B, n, partial, ysq_dev);
  FGP_LAUNCH_NAMED("spectrum_finish", st);
  return FGP_OK;
}
