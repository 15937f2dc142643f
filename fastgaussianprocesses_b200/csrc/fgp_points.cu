// K1: lattice and base-2 digital-net point generation (write-bandwidth bound, 8*n*d bytes; 16*n*d for the net).
// Replaces the host-side qmcpy generator + H2D copy at abstract_gp.py:307-309 / fast_gp_digital_net_b2.py:266-269.
#include "fgp_common.cuh"

namespace fgp {

// one thread per (point, dim) element, row-major output so a warp writes 256 contiguous bytes
__global__ void __launch_bounds__(256) lattice_points_kernel(UVec z, DVec shift, int d, uint64_t i0, uint64_t count,
                                                             double* __restrict__ x) {
  const uint64_t total = count * (uint64_t)d;
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = e / (uint64_t)d;
    const int j = (int)(e - r * (uint64_t)d);
    const uint64_t i = i0 + r;
    // phi_2(i) = brev64(i) / 2^64 ; frac(phi_2(i) z_j) = (brev64(i) * z_j mod 2^64) / 2^64  (wrap-around product)
    const uint64_t prod = __brevll(i) * z.v[j];
    // prod has at most ceil(log2(i+1)) significant bits below the top: exact in float64 for i < 2^53
    const double frac = (double)(prod >> 11) * 0x1.0p-53;
    double s = frac + shift.v[j];
    if (s >= 1.0) s -= 1.0;
    x[e] = s;
  }
}

// one thread per (point, dim); the generating-matrix columns of all dims are staged in shared memory
__global__ void __launch_bounds__(256) dnb2_points_kernel(const uint64_t* __restrict__ C, int mmax, UVec dshift, int d,
                                                          int t, uint64_t i0, uint64_t count, int64_t* __restrict__ xb,
                                                          double* __restrict__ x) {
  extern __shared__ uint64_t sC[];
  for (int k = threadIdx.x; k < d * mmax; k += blockDim.x) sC[k] = C[k];
  __syncthreads();
  const uint64_t total = count * (uint64_t)d;
  const double sc = exp2((double)-t);
  for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = e / (uint64_t)d;
    const int j = (int)(e - r * (uint64_t)d);
    uint64_t i = i0 + r;
    uint64_t v = dshift.v[j];
    const uint64_t* cj = sC + j * mmax;
    while (i) {
      const int k = __ffsll((long long)i) - 1;
      v ^= cj[k];
      i &= i - 1;
    }
    xb[e] = (int64_t)v;
    if (x) x[e] = __ll2double_rn((int64_t)v) * sc;  // int64 -> float64 (RN) then exact power-of-two scale
  }
}

}  // namespace fgp

extern "C" {

int fgp_lattice_points(const uint64_t* z_host, const double* shift_host, int d, uint64_t i0, uint64_t i1, double* x_dev,
                       fgp_stream_t stream) {
  FGP_REQUIRE(z_host && shift_host && x_dev, "lattice_points: null pointer");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D, "lattice_points: d=%d outside 1..%d", d, FGP_MAX_D);
  FGP_REQUIRE(i1 >= i0 && i1 <= (1ull << 53), "lattice_points: bad index range");
  if (i1 == i0) return FGP_OK;
  fgp::UVec z;
  fgp::DVec s;
  for (int j = 0; j < d; ++j) {
    z.v[j] = z_host[j];
    s.v[j] = shift_host[j];
    FGP_REQUIRE(s.v[j] >= 0.0 && s.v[j] < 1.0, "lattice_points: shift[%d] outside [0,1)", j);
  }
  const uint64_t total = (i1 - i0) * (uint64_t)d;
  uint64_t blocks = (total + 255) / 256;
  const uint64_t cap = (uint64_t)fgp::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  fgp::lattice_points_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(z, s, d, i0, i1 - i0, x_dev);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

int fgp_dnb2_points(const uint64_t* C_dev, int mmax, const uint64_t* dshift_host, int d, int t, uint64_t i0, uint64_t i1,
                    int64_t* xb_dev, double* x_dev, fgp_stream_t stream) {
  FGP_REQUIRE(C_dev && dshift_host && xb_dev, "dnb2_points: null pointer");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D, "dnb2_points: d=%d outside 1..%d", d, FGP_MAX_D);
  FGP_REQUIRE(t >= 1 && t < 64, "dnb2_points: t=%d outside 1..63", t);
  FGP_REQUIRE(mmax >= 1 && mmax <= 64, "dnb2_points: mmax=%d outside 1..64", mmax);
  FGP_REQUIRE(i1 >= i0 && (mmax == 64 || i1 <= (1ull << mmax)), "dnb2_points: index range exceeds 2^mmax");
  if (i1 == i0) return FGP_OK;
  fgp::UVec s;
  for (int j = 0; j < d; ++j) s.v[j] = dshift_host[j];
  const uint64_t total = (i1 - i0) * (uint64_t)d;
  uint64_t blocks = (total + 255) / 256;
  const uint64_t cap = (uint64_t)fgp::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  fgp::dnb2_points_kernel<<<(unsigned)blocks, 256, (size_t)d * mmax * 8, (cudaStream_t)stream>>>(
      C_dev, mmax, s, d, t, i0, i1 - i0, xb_dev, x_dev);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

}  // extern "C"
