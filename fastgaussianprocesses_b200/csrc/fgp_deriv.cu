// K2d: derivative-informed kernel parts and dense cross-kernel tiles.
// Reference: fast_gp_lattice.py:267-273 (Bernoulli polynomials of order 2 alpha - beta - kappa, odd orders included),
// fast_gp_digital_net_b2.py:289-301 (Walsh kernels of order alpha - beta - kappa with the (-2)^(beta+kappa) factor),
// abstract_fast_gp.py:181-191 (sum over the derivative terms of prod_j (ind_j + ls_j part_j)).
//
// A "term" is one pair (t0, t1) of derivative multi-indices of the two tasks.  Per (term, dimension) the caller passes
//   lattice: ord = polynomial degree, par[0..ord] = coefficients of coef * B_order(a) in a = frac(delta) (Horner, low first)
//   net:     ord = Walsh order 1..4,  par[0] = (-2)^(beta+kappa), par[1] = [beta+kappa > 0]
//   ind      = [beta0_j + beta1_j == 0]
// so that part = poly(a) (lattice) or par[0] * (par[1] + W_ord(delta) - 1) (net), factor = ind + ls_j * part.
// These kernels exist for coverage of the derivative path (autograd route, multitask.py); they are not tuned.
#include "fgp_common.cuh"

namespace fgp {

constexpr int kDS = FGP_DERIV_STRIDE;

template <bool NET>
__device__ __forceinline__ double deriv_part(double dx, uint64_t db, int ord, const double* __restrict__ par, int t) {
  if (NET) return par[0] * (par[1] + dnb2_part(db, ord, t));
  const double a = dx - floor(dx);
  double r = par[ord];
  for (int p = ord - 1; p >= 0; --p) r = fma(r, a, par[p]);
  return r;
}

// parts[i, term, j] of n points against ONE point z
template <bool NET>
__global__ void __launch_bounds__(256) deriv_parts_kernel(const void* __restrict__ x, int64_t n, int d, int nt, DVec z, UVec zb,
                                                          const int* __restrict__ ord, const double* __restrict__ par, int t,
                                                          double* __restrict__ parts) {
  const int64_t total = n * nt * d;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % d);
    const int64_t r = e / d;
    const int tm = (int)(r % nt);
    const int64_t i = r / nt;
    const int q = tm * d + j;
    double dx = 0.0;
    uint64_t db = 0;
    if (NET)
      db = (uint64_t)((const int64_t*)x)[i * d + j] ^ zb.v[j];
    else
      dx = ((const double*)x)[i * d + j] - z.v[j];
    parts[e] = deriv_part<NET>(dx, db, ord[q], par + (int64_t)q * kDS, t);
  }
}

// K[i,a] = scale sum_term w_term prod_j (ind + ls_j part); xs float test points, xtrain float (lattice) / int64 (net)
template <bool NET>
__global__ void __launch_bounds__(256) deriv_cross_kernel(const double* __restrict__ xs, int64_t m, const void* __restrict__ xtrain,
                                                          int64_t n, int d, int nt, const int* __restrict__ ord,
                                                          const double* __restrict__ par, const double* __restrict__ ind,
                                                          const double* __restrict__ w, int t, double scale, DVec ls,
                                                          double* __restrict__ K) {
  const int64_t total = m * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / n, a = e - i * n;
    double acc = 0.0;
    for (int tm = 0; tm < nt; ++tm) {
      double prod = scale;
      for (int j = 0; j < d; ++j) {
        const int q = tm * d + j;
        double dx = 0.0;
        uint64_t db = 0;
        if (NET)
          db = dnb2_to_b(xs[i * d + j], t) ^ (uint64_t)((const int64_t*)xtrain)[a * d + j];
        else
          dx = xs[i * d + j] - ((const double*)xtrain)[a * d + j];
        prod *= fma(ls.v[j], deriv_part<NET>(dx, db, ord[q], par + (int64_t)q * kDS, t), ind[q]);
      }
      acc = fma(w[tm], prod, acc);
    }
    K[e] = acc;
  }
}

static inline unsigned deriv_grid(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace fgp

extern "C" {

int fgp_deriv_kernel_parts(int family, const void* x_dev, int64_t n, int d, const void* z_host, int nterms, const int* ord_dev,
                           const double* par_dev, int t, double* parts_dev, fgp_stream_t stream) {
  FGP_REQUIRE(x_dev && z_host && ord_dev && par_dev && parts_dev, "deriv_kernel_parts: null pointer");
  FGP_REQUIRE(family == 0 || family == 1, "deriv_kernel_parts: family must be 0 (lattice) or 1 (net)");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D && n >= 0 && nterms >= 1, "deriv_kernel_parts: bad n/d/nterms");
  if (n == 0) return FGP_OK;
  fgp::DVec z;
  fgp::UVec zb;
  memset(&z, 0, sizeof(z));
  memset(&zb, 0, sizeof(zb));
  const unsigned grid = fgp::deriv_grid(n * nterms * d);
  if (family == 0) {
    for (int j = 0; j < d; ++j) z.v[j] = ((const double*)z_host)[j];
    fgp::deriv_parts_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, n, d, nterms, z, zb, ord_dev, par_dev, t, parts_dev);
  } else {
    FGP_REQUIRE(t >= 1 && t < 64, "deriv_kernel_parts: t outside 1..63");
    for (int j = 0; j < d; ++j) zb.v[j] = (uint64_t)((const int64_t*)z_host)[j];
    fgp::deriv_parts_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, n, d, nterms, z, zb, ord_dev, par_dev, t, parts_dev);
  }
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

int fgp_deriv_cross_kernel(int family, const double* xs_dev, int64_t m, const void* x_dev, int64_t n, int d, int nterms,
                           const int* ord_dev, const double* par_dev, const double* ind_dev, const double* w_dev, int t,
                           double scale, const double* ls_host, double* k_dev, fgp_stream_t stream) {
  FGP_REQUIRE(xs_dev && x_dev && ord_dev && par_dev && ind_dev && w_dev && ls_host && k_dev, "deriv_cross_kernel: null pointer");
  FGP_REQUIRE(family == 0 || family == 1, "deriv_cross_kernel: family must be 0 (lattice) or 1 (net)");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D && m >= 0 && n >= 0 && nterms >= 1, "deriv_cross_kernel: bad m/n/d/nterms");
  if (m == 0 || n == 0) return FGP_OK;
  fgp::DVec ls;
  memset(&ls, 0, sizeof(ls));
  for (int j = 0; j < d; ++j) ls.v[j] = ls_host[j];
  const unsigned grid = fgp::deriv_grid(m * n);
  if (family == 0) {
    fgp::deriv_cross_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(xs_dev, m, x_dev, n, d, nterms, ord_dev, par_dev, ind_dev,
                                                                         w_dev, t, scale, ls, k_dev);
  } else {
    FGP_REQUIRE(t >= 1 && t < 64, "deriv_cross_kernel: t outside 1..63");
    fgp::deriv_cross_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(xs_dev, m, x_dev, n, d, nterms, ord_dev, par_dev, ind_dev,
                                                                        w_dev, t, scale, ls, k_dev);
  }
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

}  // extern "C"
