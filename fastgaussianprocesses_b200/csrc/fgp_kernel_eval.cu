// K2: kernel parts, product kernel from parts, dense cross-kernel tiles.
// fast_gp_lattice.py:263-273, fast_gp_digital_net_b2.py:270-301, abstract_fast_gp.py:173-196.
#include "fgp_common.cuh"

namespace fgp {

__global__ void __launch_bounds__(256) lattice_parts_kernel(const double* __restrict__ x, int64_t n, int d, DVec z,
                                                            LatPoly P, double* __restrict__ parts) {
  const int64_t total = n * d;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % d);
    parts[e] = lat_part(x[e] - z.v[j], P.q[j], P.alpha[j]);
  }
}

__global__ void __launch_bounds__(256) dnb2_parts_kernel(const int64_t* __restrict__ xb, int64_t n, int d, UVec zb,
                                                         IVec alpha, int t, double* __restrict__ parts) {
  const int64_t total = n * d;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % d);
    parts[e] = dnb2_part((uint64_t)xb[e] ^ zb.v[j], alpha.v[j], t);
  }
}

__global__ void __launch_bounds__(256) kernel_from_parts_kernel(const double* __restrict__ parts, int64_t n, int d, int B,
                                                                const double* __restrict__ scale,
                                                                const double* __restrict__ ls, double* __restrict__ k) {
  const int b = blockIdx.y;
  const double s = scale[b];
  const double* l = ls + (int64_t)b * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double prod = s;
    for (int j = 0; j < d; ++j) prod *= fma(l[j], parts[i * d + j], 1.0);
    k[(int64_t)b * n + i] = prod;
  }
}

// dense K[i,a] = scale prod_j (1 + ls_j part(xs_ij, X_aj)); one thread per output, X tile in registers via L1
template <bool NET>
__global__ void __launch_bounds__(256) cross_kernel_kernel(const double* __restrict__ xs, int64_t m,
                                                           const void* __restrict__ xtrain, int64_t n, int d, LatPoly P,
                                                           IVec alpha, int t, double scale, DVec ls,
                                                           double* __restrict__ K) {
  const int64_t total = m * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / n, a = e - i * n;
    double prod = scale;
    for (int j = 0; j < d; ++j) {
      double part;
      if (NET) {
        const uint64_t xb = dnb2_to_b(xs[i * d + j], t);
        part = dnb2_part(xb ^ (uint64_t)((const int64_t*)xtrain)[a * d + j], alpha.v[j], t);
      } else {
        part = lat_part(xs[i * d + j] - ((const double*)xtrain)[a * d + j], P.q[j], P.alpha[j]);
      }
      prod *= fma(ls.v[j], part, 1.0);
    }
    K[e] = prod;
  }
}


// k[i] = scale prod_j (1 + ls_j part(x_ij, z_ij)) for N row pairs (diagonal terms k(x,x) and elementwise kernel() calls)
template <bool NET>
__global__ void __launch_bounds__(256) pair_kernel_kernel(const double* __restrict__ x, const void* __restrict__ z,
                                                          int z_is_int, int64_t N, int d, LatPoly P, IVec alpha, int t,
                                                          double scale, DVec ls, double* __restrict__ k) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    double prod = scale;
    for (int j = 0; j < d; ++j) {
      double part;
      if (NET) {
        const uint64_t xb = dnb2_to_b(x[i * d + j], t);
        const uint64_t zb = z_is_int ? (uint64_t)((const int64_t*)z)[i * d + j] : dnb2_to_b(((const double*)z)[i * d + j], t);
        part = dnb2_part(xb ^ zb, alpha.v[j], t);
      } else {
        part = lat_part(x[i * d + j] - ((const double*)z)[i * d + j], P.q[j], P.alpha[j]);
      }
      prod *= fma(ls.v[j], part, 1.0);
    }
    k[i] = prod;
  }
}

static inline unsigned grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace fgp

extern "C" {

int fgp_lattice_kernel_parts(const double* x_dev, int64_t n, int d, const double* z_host, const int* alpha_host,
                             double* parts_dev, fgp_stream_t stream) {
  FGP_REQUIRE(x_dev && z_host && alpha_host && parts_dev, "lattice_kernel_parts: null pointer");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D && n >= 0, "lattice_kernel_parts: bad n/d");
  if (n == 0) return FGP_OK;
  fgp::LatPoly P;
  int rc = fgp::fill_lat_poly(alpha_host, d, &P);
  if (rc) return rc;
  fgp::DVec z;
  for (int j = 0; j < d; ++j) z.v[j] = z_host[j];
  fgp::lattice_parts_kernel<<<fgp::grid_for(n * d), 256, 0, (cudaStream_t)stream>>>(x_dev, n, d, z, P, parts_dev);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

int fgp_dnb2_kernel_parts(const int64_t* xb_dev, int64_t n, int d, const int64_t* zb_host, const int* alpha_host, int t,
                          double* parts_dev, fgp_stream_t stream) {
  FGP_REQUIRE(xb_dev && zb_host && alpha_host && parts_dev, "dnb2_kernel_parts: null pointer");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D && n >= 0, "dnb2_kernel_parts: bad n/d");
  FGP_REQUIRE(t >= 1 && t < 64, "dnb2_kernel_parts: t outside 1..63");
  if (n == 0) return FGP_OK;
  fgp::UVec zb;
  fgp::IVec al;
  for (int j = 0; j < d; ++j) {
    zb.v[j] = (uint64_t)zb_host[j];
    al.v[j] = alpha_host[j];
    FGP_REQUIRE(al.v[j] >= 1 && al.v[j] <= 4, "dnb2_kernel_parts: alpha[%d]=%d outside 1..4", j, al.v[j]);
  }
  fgp::dnb2_parts_kernel<<<fgp::grid_for(n * d), 256, 0, (cudaStream_t)stream>>>(xb_dev, n, d, zb, al, t, parts_dev);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

int fgp_kernel_from_parts(const double* parts_dev, int64_t n, int d, int B, const double* scale_dev, const double* ls_dev,
                          double* k_dev, fgp_stream_t stream) {
  FGP_REQUIRE(parts_dev && scale_dev && ls_dev && k_dev, "kernel_from_parts: null pointer");
  FGP_REQUIRE(d >= 1 && n >= 0 && B >= 1 && B <= 65535, "kernel_from_parts: bad n/d/B");
  if (n == 0) return FGP_OK;
  dim3 grid(fgp::grid_for(n), B);
  fgp::kernel_from_parts_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(parts_dev, n, d, B, scale_dev, ls_dev, k_dev);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

static int cross_common(bool net, const double* xs, int64_t m, const void* xt, int64_t n, int d, const int* alpha_host,
                        int t, double scale, const double* ls_host, double* K, fgp_stream_t stream) {
  FGP_REQUIRE(xs && xt && alpha_host && ls_host && K, "cross_kernel: null pointer");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D && m >= 0 && n >= 0, "cross_kernel: bad m/n/d");
  if (m == 0 || n == 0) return FGP_OK;
  fgp::LatPoly P;
  memset(&P, 0, sizeof(P));
  fgp::IVec al;
  fgp::DVec ls;
  for (int j = 0; j < d; ++j) {
    al.v[j] = alpha_host[j];
    ls.v[j] = ls_host[j];
  }
  if (!net) {
    int rc = fgp::fill_lat_poly(alpha_host, d, &P);
    if (rc) return rc;
    fgp::cross_kernel_kernel<false><<<fgp::grid_for(m * n), 256, 0, (cudaStream_t)stream>>>(xs, m, xt, n, d, P, al, t,
                                                                                         scale, ls, K);
  } else {
    FGP_REQUIRE(t >= 1 && t < 64, "cross_kernel: t outside 1..63");
    for (int j = 0; j < d; ++j) FGP_REQUIRE(al.v[j] >= 1 && al.v[j] <= 4, "cross_kernel: net alpha outside 1..4");
    fgp::cross_kernel_kernel<true><<<fgp::grid_for(m * n), 256, 0, (cudaStream_t)stream>>>(xs, m, xt, n, d, P, al, t,
                                                                                        scale, ls, K);
  }
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

int fgp_lattice_cross_kernel(const double* xs_dev, int64_t m, const double* x_dev, int64_t n, int d, const int* alpha_host,
                             double scale, const double* ls_host, double* k_dev, fgp_stream_t stream) {
  return cross_common(false, xs_dev, m, x_dev, n, d, alpha_host, 0, scale, ls_host, k_dev, stream);
}
int fgp_dnb2_cross_kernel(const double* xs_dev, int64_t m, const int64_t* xb_dev, int64_t n, int d, const int* alpha_host,
                          int t, double scale, const double* ls_host, double* k_dev, fgp_stream_t stream) {
  return cross_common(true, xs_dev, m, xb_dev, n, d, alpha_host, t, scale, ls_host, k_dev, stream);
}

int fgp_kernel_pairs(int family, const double* x_dev, const void* z_dev, int z_is_int, int64_t N, int d,
                     const int* alpha_host, int t, double scale, const double* ls_host, double* k_dev, fgp_stream_t stream) {
  FGP_REQUIRE(x_dev && z_dev && alpha_host && ls_host && k_dev, "kernel_pairs: null pointer");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D && N >= 0, "kernel_pairs: bad N/d");
  if (N == 0) return FGP_OK;
  fgp::LatPoly P;
  memset(&P, 0, sizeof(P));
  fgp::IVec al;
  fgp::DVec ls;
  for (int j = 0; j < d; ++j) {
    al.v[j] = alpha_host[j];
    ls.v[j] = ls_host[j];
  }
  if (family == 0) {
    FGP_REQUIRE(!z_is_int, "kernel_pairs: lattice points are float64");
    int rc = fgp::fill_lat_poly(alpha_host, d, &P);
    if (rc) return rc;
    fgp::pair_kernel_kernel<false><<<fgp::grid_for(N), 256, 0, (cudaStream_t)stream>>>(x_dev, z_dev, 0, N, d, P, al, t, scale, ls, k_dev);
  } else {
    FGP_REQUIRE(t >= 1 && t < 64, "kernel_pairs: t outside 1..63");
    for (int j = 0; j < d; ++j) FGP_REQUIRE(al.v[j] >= 1 && al.v[j] <= 4, "kernel_pairs: net alpha outside 1..4");
    fgp::pair_kernel_kernel<true><<<fgp::grid_for(N), 256, 0, (cudaStream_t)stream>>>(x_dev, z_dev, z_is_int, N, d, P, al, t, scale, ls, k_dev);
  }
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

}  // extern "C"
