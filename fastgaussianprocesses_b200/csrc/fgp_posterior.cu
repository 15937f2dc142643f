// K5: posterior mean / variance as on-the-fly kernel-vector products (abstract_gp.py:352-380, :381-416).
// The reference materialises parts of shape (m, n, 1, 1, d) (abstract_fast_gp.py:173-180) and then contracts with
// coeffs; here the (m x n) cross-covariance only ever exists as one FP64 register per (test point, thread).
//
// post_mean is FP64-pipe bound.  Lattice alpha = 2 inner loop, per (test point, train point, dim):
//     delta = xs' - X'          DADD      (coordinates pre-scaled by sigma_j = kappa_j^(1/2))
//     v     = sigma_j - |delta| DADD
//     h     = |delta| * v       DMUL      (= kappa_j a (1 - a),  a = |x - X| ; B_4(frac t) = B_4(|t|))
//     f     = 1 - h*h           DFMA      (= (1 + ls_j c B_4(a)) / A_j)
//     prod *= f                 DMUL
// i.e. 5 FP64 issue slots per pair-dimension + 1 DFMA per pair for the coefficient: (5d + 1) slots / pair.
#include "fgp_transform.cuh"

namespace fgp {

constexpr int kPT = 256;   // threads per CTA
constexpr int kTN = 128;   // train points per shared-memory tile

struct PostArgs {
  const double* xs;   // (m,d) test points
  int64_t m;
  const void* x;      // (n,d) train points: double (lattice) / int64 (net)
  int64_t n;
  int d;
  int t;
  double tscale;  // net: 2^-t
  const double* coeffs;  // (B,n)
  int B;
  double* partial;    // (splits, B, m) or the output itself when splits == 1
  int splits;
  int64_t n_per_split;
  // per-dimension polynomial in u = a(1-a):  f_j(u) = sum_p c[j][p] u^p  (ls and the "+1" folded in)
  double c[FGP_MAX_D][FGP_MAX_ALPHA + 1];
  int alpha[FGP_MAX_D];
  double sig[FGP_MAX_D];  // alpha = 2 fast path: sigma_j
  double ls[FGP_MAX_D];   // net: lengthscales
  double pref;            // scale (times prod_j A_j on the fast path)
};

// MODE 0: lattice alpha=2 fast path; 1: lattice generic alpha; 2: net generic alpha; 3: net alpha=2 (branch-free part);
// 4: net alpha=2 with t <= 52 (exact conversion through the 2^52 magic number, fgp_common.cuh)
template <int DT, int R, int MODE>
__global__ void __launch_bounds__(kPT) post_mean_kernel(const __grid_constant__ PostArgs a) {
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  const int d = DT > 0 ? DT : a.d;
  extern __shared__ __align__(16) unsigned char smraw[];
  double* sX = (double*)smraw;          // kTN * d   (lattice: scaled coords; net: int64 bit patterns)
  double* sC = sX + (size_t)kTN * d;    // kTN coefficients
  const int b = blockIdx.z;
  const int64_t i0 = ((int64_t)blockIdx.x * kPT + threadIdx.x) * R;
  double xr[R][DM];
  uint64_t xbr[R][MODE >= 2 ? DM : 1];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t i = i0 + r < a.m ? i0 + r : a.m - 1;
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      if (j >= d) break;
      const double v = a.xs[i * d + j];
      if (MODE == 0) xr[r][j] = v * a.sig[j];
      if (MODE == 1) xr[r][j] = v;
      if (MODE >= 2) xbr[r][j] = dnb2_to_b(v, a.t);
    }
  }
  double acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.0;
  const int64_t a0 = (int64_t)blockIdx.y * a.n_per_split;
  const int64_t a1 = min(a.n, a0 + a.n_per_split);
  const double* coef = a.coeffs + (int64_t)b * a.n;
  for (int64_t base = a0; base < a1; base += kTN) {
    const int cnt = (int)min((int64_t)kTN, a1 - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * d; e += kPT) {
      if (MODE >= 2) {
        ((int64_t*)sX)[e] = ((const int64_t*)a.x)[base * d + e];
      } else {
        const double v = ((const double*)a.x)[base * d + e];
        sX[e] = MODE == 0 ? v * a.sig[e % d] : v;
      }
    }
    for (int e = threadIdx.x; e < cnt; e += kPT) sC[e] = coef[base + e];
    __syncthreads();
#pragma unroll 2
    for (int k = 0; k < cnt; ++k) {
      const double ck = sC[k];
      double prod[R];
#pragma unroll
      for (int r = 0; r < R; ++r) prod[r] = 1.0;
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        if (j >= d) break;
        if (MODE >= 2) {
          const uint64_t Xb = ((const uint64_t*)sX)[k * d + j];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const double part = MODE == 4 ? dnb2_part_a2_t52(xbr[r][j] ^ Xb, a.t, a.tscale)
                                          : (MODE == 3 ? dnb2_part_a2(xbr[r][j] ^ Xb, a.t, a.tscale) : dnb2_part(xbr[r][j] ^ Xb, a.alpha[j], a.t));
            const double f = fma(a.ls[j], part, 1.0);
            prod[r] = j == 0 ? f : prod[r] * f;
          }
        } else {
          const double X = sX[k * d + j];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const double ad = fabs(xr[r][j] - X);
            if (MODE == 0) {
              const double h = ad * (a.sig[j] - ad);
              const double f = fma(-h, h, 1.0);
              prod[r] = j == 0 ? f : prod[r] * f;
            } else {
              const double u = ad * (1.0 - ad);
              double f = a.c[j][a.alpha[j]];
              for (int p = a.alpha[j] - 1; p >= 0; --p) f = fma(f, u, a.c[j][p]);
              prod[r] *= f;
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = fma(prod[r], ck, acc[r]);
    }
  }
  double* out = a.partial + ((int64_t)blockIdx.y * a.B + b) * a.m;
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (i0 + r < a.m) out[i0 + r] = acc[r] * a.pref;
}

__global__ void __launch_bounds__(256) post_mean_reduce_kernel(const double* __restrict__ partial, int splits,
                                                               int64_t total, double* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += partial[(int64_t)k * total + e];
    out[e] = s;
  }
}

// pvar_i = max(0, kxx - sum_k |kt_ik|^2 Re(1/lam_k))
template <bool CPLX>
__global__ void __launch_bounds__(256) post_var_reduce_kernel(const double* __restrict__ kt, const double* __restrict__ lam,
                                                              int64_t n, double kxx, double* __restrict__ out) {
  __shared__ double red[32 * 4];
  const int64_t i = blockIdx.x;
  double s[1] = {0.0};
  if (CPLX) {
    const double2* row = (const double2*)kt + i * n;
    const double2* l = (const double2*)lam;
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
      const double2 v = row[k], lk = l[k];
      s[0] = fma(fma(v.x, v.x, v.y * v.y), lk.x / fma(lk.x, lk.x, lk.y * lk.y), s[0]);
    }
  } else {
    const double* row = kt + i * n;
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
      const double v = row[k];
      s[0] = fma(v * v, 1.0 / lam[k], s[0]);
    }
  }
  block_sum<1>(s, red);
  if (threadIdx.x == 0) {
    const double v = kxx - s[0];
    out[i] = v < 0.0 ? 0.0 : v;
  }
}

// Lattice post_var, two test points per complex transform: row p holds z_a = k(x*_{2p}, X_a) + i k(x*_{2p+1}, X_a).  With
// Z = ft(z), the spectra of the two real sequences are A_k = (Z_k + conj Z_{n-k}) / 2 and B_k = (Z_k - conj Z_{n-k}) / (2i), so
// one length-n complex transform serves two test points: half the transform work and no separate real input array.
// A2: every alpha_j == 2 (branch-free part, as in the post_mean inner loop).  A thread evaluates kPG pairs (2 kPG test points)
// against ONE training point, so the training points are read from the L2 once per 2 kPG test points.
constexpr int kPG = 4;
template <bool A2>
__global__ void __launch_bounds__(256) lattice_cross_pair_kernel(const double* __restrict__ xs, int64_t m, const double* __restrict__ xtrain,
                                                                 int64_t n, int d, LatPoly P, double scale, DVec ls,
                                                                 double2* __restrict__ K) {
  const int64_t pairs = (m + 1) >> 1, groups = (pairs + kPG - 1) / kPG, total = groups * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = e / n, a = e - g * n;
    const double* __restrict__ xa = xtrain + a * d;
    const double* xr[2 * kPG];
    double k[2 * kPG];
#pragma unroll
    for (int q = 0; q < 2 * kPG; ++q) {
      int64_t i = 2 * kPG * g + q;
      if (i >= m) i = m - 1;
      xr[q] = xs + i * d;
      k[q] = scale;
    }
    for (int j = 0; j < d; ++j) {
      const double v = __ldg(xa + j);
      const double l = ls.v[j];
#pragma unroll
      for (int q = 0; q < 2 * kPG; ++q) {
        const double part = A2 ? lat_part_a2(__ldg(xr[q] + j) - v, P.q[j][0], P.q[j][2]) : lat_part(__ldg(xr[q] + j) - v, P.q[j], P.alpha[j]);
        k[q] *= fma(l, part, 1.0);
      }
    }
#pragma unroll
    for (int q = 0; q < kPG; ++q) {
      const int64_t p = kPG * g + q;
      if (p < pairs) K[p * n + a] = make_double2(k[2 * q], 2 * p + 1 < m ? k[2 * q + 1] : 0.0);
    }
  }
}

// partial[p][seg] = sum over the seg-th part of the spectrum of (|A_k|^2, |B_k|^2) Re(1/lam_k); grid (segs, pairs).  One CTA per
// pair left 128 CTAs reading 16 MiB each (2.1 ms for 256 points at n = 2^20, more than the transform); the partial sums are
// added in a fixed order by post_var_pair_final_kernel, so the result does not depend on scheduling.
__global__ void __launch_bounds__(256) post_var_pair_reduce_kernel(const double2* __restrict__ kt, const double2* __restrict__ lam, int64_t n,
                                                                   double* __restrict__ partial) {
  __shared__ double red[32 * 4];
  const int64_t p = blockIdx.y;
  const int segs = gridDim.x, seg = blockIdx.x;
  // |A_{n-k}| = |A_k| and |B_{n-k}| = |B_k|: visit k = 0..n/2 and give the pair (k, n-k) both weights, so every element of the
  // row is read once (the ncu capture of the one-index-at-a-time version showed it HBM-bound at 5.9 TB/s reading each row twice)
  const int64_t half = n >> 1, tot = half + 1;
  const int64_t k0 = tot / segs * seg, k1 = seg == segs - 1 ? tot : tot / segs * (seg + 1);
  const double2* row = kt + p * n;
  double s[2] = {0.0, 0.0};
  for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
    const bool self = k == 0 || k == half;
    const double2 z = row[k], lk = lam[k];
    const double2 zm = self ? z : row[n - k];
    double w = 0.25 * lk.x / fma(lk.x, lk.x, lk.y * lk.y);
    if (!self) {
      const double2 lm = lam[n - k];
      w += 0.25 * lm.x / fma(lm.x, lm.x, lm.y * lm.y);
    }
    const double ar = z.x + zm.x, ai = z.y - zm.y;  // Z_k + conj Z_{n-k}
    const double br = z.x - zm.x, bi = z.y + zm.y;  // Z_k - conj Z_{n-k}
    s[0] = fma(fma(ar, ar, ai * ai), w, s[0]);
    s[1] = fma(fma(br, br, bi * bi), w, s[1]);
  }
  block_sum<2>(s, red);
  if (threadIdx.x == 0) {
    partial[(p * segs + seg) * 2 + 0] = s[0];
    partial[(p * segs + seg) * 2 + 1] = s[1];
  }
}
// pvar_{2p} = max(0, kxx - sum_k |A_k|^2 Re(1/lam_k)), pvar_{2p+1} likewise with B_k
__global__ void __launch_bounds__(256) post_var_pair_final_kernel(const double* __restrict__ partial, int segs, int64_t pairs, int64_t m, double kxx,
                                                                  double* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pairs) return;
  double s0 = 0.0, s1 = 0.0;
  for (int g = 0; g < segs; ++g) {
    s0 += partial[(p * segs + g) * 2 + 0];
    s1 += partial[(p * segs + g) * 2 + 1];
  }
  const double v0 = kxx - s0, v1 = kxx - s1;
  out[2 * p] = v0 < 0.0 ? 0.0 : v0;
  if (2 * p + 1 < m) out[2 * p + 1] = v1 < 0.0 ? 0.0 : v1;
}

// net: partial[i][seg] = sum over the seg-th part of kt_ik^2 / lam_k (grid (segs, points)), finished by post_var_final_kernel
__global__ void __launch_bounds__(256) post_var_reduce_seg_kernel(const double* __restrict__ kt, const double* __restrict__ lam, int64_t n,
                                                                  double* __restrict__ partial) {
  __shared__ double red[32 * 4];
  const int64_t i = blockIdx.y;
  const int segs = gridDim.x, seg = blockIdx.x;
  const int64_t k0 = n / segs * seg, k1 = seg == segs - 1 ? n : n / segs * (seg + 1);
  const double* row = kt + i * n;
  double s[1] = {0.0};
  for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
    const double v = row[k];
    s[0] = fma(v * v, 1.0 / lam[k], s[0]);
  }
  block_sum<1>(s, red);
  if (threadIdx.x == 0) partial[i * segs + seg] = s[0];
}
__global__ void __launch_bounds__(256) post_var_final_kernel(const double* __restrict__ partial, int segs, int64_t cnt, double kxx, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  double s = 0.0;
  for (int g = 0; g < segs; ++g) s += partial[i * segs + g];
  const double v = kxx - s;
  out[i] = v < 0.0 ? 0.0 : v;
}

template <int DT, int R, int MODE>
static int launch_post_mean(const PostArgs& a, dim3 grid, size_t smem, cudaStream_t st) {
  post_mean_kernel<DT, R, MODE><<<grid, kPT, smem, st>>>(a);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

template <int MODE>
static int dispatch_post_mean(const PostArgs& a, cudaStream_t st) {
  const int d = a.d;
  const int R = d <= 8 ? 2 : 1;
  const int64_t per_cta = (int64_t)kPT * R;
  dim3 grid((unsigned)((a.m + per_cta - 1) / per_cta), (unsigned)a.splits, (unsigned)a.B);
  const size_t smem = (size_t)kTN * (d + 1) * sizeof(double);
  switch (d) {
    case 2: return launch_post_mean<2, 2, MODE>(a, grid, smem, st);
    case 3: return launch_post_mean<3, 2, MODE>(a, grid, smem, st);
    case 4: return launch_post_mean<4, 2, MODE>(a, grid, smem, st);
    case 8: return launch_post_mean<8, 2, MODE>(a, grid, smem, st);
    case 16: return launch_post_mean<16, 1, MODE>(a, grid, smem, st);
    default:
      if (R == 2) return launch_post_mean<0, 2, MODE>(a, grid, smem, st);
      return launch_post_mean<0, 1, MODE>(a, grid, smem, st);
  }
}

static int choose_splits(int64_t m, int64_t n, int d, int B) {
  const int R = d <= 8 ? 2 : 1;
  const int64_t ctas = ((m + (int64_t)kPT * R - 1) / ((int64_t)kPT * R)) * B;
  const int64_t want = (int64_t)sm_count() * 4;
  int64_t s = 1;
  if (ctas < want) s = (want + ctas - 1) / ctas;
  const int64_t maxs = (n + kTN - 1) / kTN;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int)s;
}

static int post_mean_common(int family, const double* xs, int64_t m, const void* x, int64_t n, int d,
                            const int* alpha_host, int t, double scale, const double* ls_host, const double* coeffs, int B,
                            void* partial, double* pmean, fgp_stream_t stream) {
  FGP_REQUIRE(xs && x && alpha_host && ls_host && coeffs && pmean, "post_mean: null pointer");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D && m >= 0 && n >= 1 && B >= 1 && B <= 65535, "post_mean: bad m/n/d/B");
  if (m == 0) return FGP_OK;
  PostArgs a;
  memset(&a, 0, sizeof(a));
  a.xs = xs;
  a.m = m;
  a.x = x;
  a.n = n;
  a.d = d;
  a.t = t;
  a.tscale = ldexp(1.0, -t);
  a.coeffs = coeffs;
  a.B = B;
  a.splits = choose_splits(m, n, d, B);
  a.n_per_split = ((n + a.splits - 1) / a.splits + kTN - 1) / kTN * kTN;
  a.splits = (int)((n + a.n_per_split - 1) / a.n_per_split);
  FGP_REQUIRE(a.splits == 1 || partial, "post_mean: null workspace");
  a.partial = a.splits == 1 ? pmean : (double*)partial;
  a.pref = scale;
  int mode;
  if (family == 0) {
    LatPoly P;
    int rc = fill_lat_poly(alpha_host, d, &P);
    if (rc) return rc;
    bool all2 = true;
    for (int j = 0; j < d; ++j) {
      a.alpha[j] = alpha_host[j];
      all2 = all2 && alpha_host[j] == 2;
      for (int p = 0; p <= alpha_host[j]; ++p) a.c[j][p] = ls_host[j] * P.q[j][p];
      a.c[j][0] += 1.0;
    }
    mode = all2 ? 0 : 1;
    if (all2) {
      for (int j = 0; j < d; ++j) {
        const double A = a.c[j][0];  // 1 + ls q0 > 0
        const double Bq = -a.c[j][2];  // ls |q2| > 0
        FGP_REQUIRE(A > 0.0 && Bq >= 0.0, "post_mean: unexpected Bernoulli coefficients");
        a.sig[j] = sqrt(sqrt(Bq / A));
        a.pref *= A;
      }
    }
  } else {
    FGP_REQUIRE(t >= 1 && t < 64, "post_mean: t outside 1..63");
    for (int j = 0; j < d; ++j) {
      a.alpha[j] = alpha_host[j];
      a.ls[j] = ls_host[j];
      FGP_REQUIRE(alpha_host[j] >= 1 && alpha_host[j] <= 4, "post_mean: net alpha outside 1..4");
    }
    bool all2 = true;
    for (int j = 0; j < d; ++j) all2 = all2 && alpha_host[j] == 2;
    mode = all2 ? (t <= 52 ? 4 : 3) : 2;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = mode == 0 ? dispatch_post_mean<0>(a, st)
                     : (mode == 1 ? dispatch_post_mean<1>(a, st)
                                  : (mode == 2 ? dispatch_post_mean<2>(a, st) : (mode == 3 ? dispatch_post_mean<3>(a, st) : dispatch_post_mean<4>(a, st))));
  if (rc) return rc;
  if (a.splits > 1) {
    const int64_t total = (int64_t)B * m;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    post_mean_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>((const double*)partial, a.splits, total, pmean);
    FGP_LAUNCH_CHECK();
  }
  return FGP_OK;
}

static int64_t post_var_chunk(int64_t m, int64_t n) {
  int64_t mc = (int64_t(1) << 27) / n;
  if (mc < 1) mc = 1;
  if (mc > m) mc = m;
  return mc;
}

}  // namespace fgp

extern "C" {

size_t fgp_post_mean_workspace_bytes(int64_t m, int64_t n, int d, int B) {
  if (m <= 0 || n <= 0 || d < 1 || B < 1) return 0;
  const int s = fgp::choose_splits(m, n, d, B);
  return s <= 1 ? 256 : (size_t)s * B * m * sizeof(double);
}

int fgp_lattice_post_mean(const double* xs_dev, int64_t m, const double* x_dev, int64_t n, int d, const int* alpha_host,
                          double scale, const double* ls_host, const double* coeffs_dev, int B, void* partial_dev,
                          double* pmean_dev, fgp_stream_t stream) {
  return fgp::post_mean_common(0, xs_dev, m, x_dev, n, d, alpha_host, 0, scale, ls_host, coeffs_dev, B, partial_dev,
                               pmean_dev, stream);
}

int fgp_dnb2_post_mean(const double* xs_dev, int64_t m, const int64_t* xb_dev, int64_t n, int d, const int* alpha_host,
                       int t, double scale, const double* ls_host, const double* coeffs_dev, int B, void* partial_dev,
                       double* pmean_dev, fgp_stream_t stream) {
  return fgp::post_mean_common(1, xs_dev, m, xb_dev, n, d, alpha_host, t, scale, ls_host, coeffs_dev, B, partial_dev,
                               pmean_dev, stream);
}

size_t fgp_post_var_workspace_bytes(int family, int64_t m, int64_t n) {
  if (m <= 0 || n <= 0) return 0;
  const int64_t mc = fgp::post_var_chunk(m, n);
  // lattice: mc PAIRS of test points per chunk = mc*n complex values + 16 x 2 partial sums per pair (n = 1: a real and a complex row
  // per point); net: mc rows + 16 partial sums per row
  if (family == 0) return (n > 1 ? (size_t)mc * n * 2 + 32 * (size_t)mc + 2 : (size_t)mc * 3 + 2) * sizeof(double);
  return ((size_t)mc * n + 2 + 16 * (size_t)mc) * sizeof(double);
}

static int post_var_common(int family, const double* xs, int64_t m, const void* x, int64_t n, int d, const int* alpha_host,
                           int t, double scale, const double* ls_host, const double* lam, const void* table, void* work,
                           double* pvar, fgp_stream_t stream) {
  using namespace fgp;
  FGP_REQUIRE(xs && x && alpha_host && ls_host && lam && work && pvar, "post_var: null pointer");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D && m >= 0 && is_pow2(n), "post_var: bad m/n/d (n must be a power of two)");
  if (m == 0) return FGP_OK;
  // k(x,x): delta = 0 in every dimension
  double kxx = scale;
  if (family == 0) {
    FGP_REQUIRE(table, "post_var: null twiddle table");
    LatPoly P;
    int rc = fill_lat_poly(alpha_host, d, &P);
    if (rc) return rc;
    for (int j = 0; j < d; ++j) kxx *= 1.0 + ls_host[j] * P.q[j][0];
  } else {
    static const double w0[5] = {0.0, 1.0, 1.5, 43.0 / 18.0 - 1.0, 701.0 / 294.0 - 1.0};
    for (int j = 0; j < d; ++j) {
      FGP_REQUIRE(alpha_host[j] >= 1 && alpha_host[j] <= 4, "post_var: net alpha outside 1..4");
      kxx *= 1.0 + ls_host[j] * w0[alpha_host[j]];
    }
  }
  const int64_t mc = post_var_chunk(m, n);
  cudaStream_t st = (cudaStream_t)stream;
  if (family == 0 && n > 1) {
    // two test points per complex transform, in place in the workspace
    LatPoly P;
    int rc = fill_lat_poly(alpha_host, d, &P);
    if (rc) return rc;
    DVec ls;
    memset(&ls, 0, sizeof(ls));
    for (int j = 0; j < d; ++j) ls.v[j] = ls_host[j];
    bool all2 = true;
    for (int j = 0; j < d; ++j) all2 = all2 && alpha_host[j] == 2;
    const int64_t mc2 = 2 * mc;  // points per chunk: mc pairs = 2*mc*n doubles of the 3*mc*n-double workspace
    double2* kc = (double2*)work;
    for (int64_t i0 = 0; i0 < m; i0 += mc2) {
      const int64_t cnt = m - i0 < mc2 ? m - i0 : mc2;
      const int64_t pairs = (cnt + 1) >> 1;
      int64_t blocks = (((pairs + kPG - 1) / kPG) * n + 255) / 256;
      const int64_t cap = (int64_t)sm_count() * 16;
      if (blocks > cap) blocks = cap;
      if (all2)
        lattice_cross_pair_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(xs + i0 * d, cnt, (const double*)x, n, d, P, scale, ls, kc);
      else
        lattice_cross_pair_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(xs + i0 * d, cnt, (const double*)x, n, d, P, scale, ls, kc);
      FGP_LAUNCH_CHECK();
      if ((rc = fgp_fftbr_c2c((const double*)kc, (double*)kc, pairs, n, table, stream))) return rc;
      // partial sums behind the mc pairs of the chunk (the third n*mc doubles of the workspace are free in this path)
      const int segs = n >= 8192 ? 16 : 1;
      double* partial = (double*)work + 2 * mc * n;
      post_var_pair_reduce_kernel<<<dim3(segs, (unsigned)pairs), 256, 0, st>>>(kc, (const double2*)lam, n, partial);
      FGP_LAUNCH_CHECK();
      post_var_pair_final_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(partial, segs, pairs, cnt, kxx, pvar + i0);
      FGP_LAUNCH_CHECK();
    }
    return FGP_OK;
  }
  double* kreal = (double*)work;
  double* kcplx = kreal + ((mc * n + 1) & ~(int64_t)1);  // complex rows need 16-byte alignment (mc*n may be odd when n = 1)
  for (int64_t i0 = 0; i0 < m; i0 += mc) {
    const int64_t cnt = m - i0 < mc ? m - i0 : mc;
    int rc;
    if (family == 0) {
      if ((rc = fgp_lattice_cross_kernel(xs + i0 * d, cnt, (const double*)x, n, d, alpha_host, scale, ls_host, kreal, stream))) return rc;
      if ((rc = fgp_fftbr_r2c(kreal, kcplx, cnt, n, table, stream))) return rc;
      post_var_reduce_kernel<true><<<(unsigned)cnt, 256, 0, st>>>(kcplx, lam, n, kxx, pvar + i0);
    } else {
      if ((rc = fgp_dnb2_cross_kernel(xs + i0 * d, cnt, (const int64_t*)x, n, d, alpha_host, t, scale, ls_host, kreal, stream))) return rc;
      if ((rc = fgp_fwht(kreal, kreal, cnt, n, stream))) return rc;
      if (n >= 8192 && cnt < 2048) {  // few long rows: several CTAs per row, partial sums behind the chunk's rows
        double* partial = kreal + mc * n;
        post_var_reduce_seg_kernel<<<dim3(16, (unsigned)cnt), 256, 0, st>>>(kreal, lam, n, partial);
        FGP_LAUNCH_CHECK();
        post_var_final_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(partial, 16, cnt, kxx, pvar + i0);
      } else {
        post_var_reduce_kernel<false><<<(unsigned)cnt, 256, 0, st>>>(kreal, lam, n, kxx, pvar + i0);
      }
    }
    FGP_LAUNCH_CHECK();
  }
  return FGP_OK;
}

int fgp_lattice_post_var(const double* xs_dev, int64_t m, const double* x_dev, int64_t n, int d, const int* alpha_host,
                         double scale, const double* ls_host, const double* lam_dev, const void* table_dev, void* work_dev,
                         double* pvar_dev, fgp_stream_t stream) {
  return post_var_common(0, xs_dev, m, x_dev, n, d, alpha_host, 0, scale, ls_host, lam_dev, table_dev, work_dev, pvar_dev,
                         stream);
}

int fgp_dnb2_post_var(const double* xs_dev, int64_t m, const int64_t* xb_dev, int64_t n, int d, const int* alpha_host, int t,
                      double scale, const double* ls_host, const double* lam_dev, void* work_dev, double* pvar_dev,
                      fgp_stream_t stream) {
  return post_var_common(1, xs_dev, m, xb_dev, n, d, alpha_host, t, scale, ls_host, lam_dev, nullptr, work_dev, pvar_dev,
                         stream);
}

}  // extern "C"
