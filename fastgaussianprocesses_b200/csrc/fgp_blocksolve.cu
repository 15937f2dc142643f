// K6: per-frequency block systems of the multi-task / derivative-informed eigen-solve (util.py:301-323, :354-363).
// T randomisations of one lattice / net share the transform, so the (sum_l n_l)^2 Gram matrix reduces to n_min independent R x R systems
// Lam_k (Hermitian for the lattice, symmetric for the net; R = sum_l n_l / n_min <= 16).  The reference eliminates them with a Schur
// recursion written in torch; round 1 of this package used batched torch.linalg.inv / slogdet.  Here one thread owns one system:
// in-place Gauss-Jordan inversion with partial pivoting in registers / local memory, log|det| from the pivots, no library call.
#include "fgp_common.cuh"

namespace fgp {

constexpr int kMaxR = 16;

template <bool CPLX>
struct Num;
template <>
struct Num<false> {
  typedef double T;
  static __device__ __forceinline__ double abs2(double a) { return a * a; }
  static __device__ __forceinline__ double inv(double a) { return 1.0 / a; }
  static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
  static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
  static __device__ __forceinline__ double zero() { return 0.0; }
  static __device__ __forceinline__ double one() { return 1.0; }
};
template <>
struct Num<true> {
  typedef double2 T;
  static __device__ __forceinline__ double abs2(double2 a) { return fma(a.x, a.x, a.y * a.y); }
  static __device__ __forceinline__ double2 inv(double2 a) {
    const double s = 1.0 / fma(a.x, a.x, a.y * a.y);
    return make_double2(a.x * s, -a.y * s);
  }
  static __device__ __forceinline__ double2 mul(double2 a, double2 b) { return cmul(a, b); }
  static __device__ __forceinline__ double2 sub(double2 a, double2 b) { return csub(a, b); }
  static __device__ __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ double2 one() { return make_double2(1.0, 0.0); }
};

// L, A: (nm, R, R) row-major; logdet: (nm).  Compile-time R = RT: the matrix lives in registers.
template <bool CPLX, int RT>
__global__ void __launch_bounds__(128) block_inv_logdet_kernel(const typename Num<CPLX>::T* __restrict__ L, int64_t nm, int Rrt,
                                                               typename Num<CPLX>::T* __restrict__ A, double* __restrict__ logdet) {
  typedef Num<CPLX> N;
  typedef typename N::T T;
  constexpr int RM = RT > 0 ? RT : kMaxR;
  const int R = RT > 0 ? RT : Rrt;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nm) return;
  T a[RM][RM];
  int piv[RM];
  const T* src = L + k * R * R;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    if (i >= R) break;
#pragma unroll
    for (int j = 0; j < RM; ++j) {
      if (j >= R) break;
      a[i][j] = src[i * R + j];
    }
  }
  double ld = 0.0;
#pragma unroll
  for (int c = 0; c < RM; ++c) {
    if (c >= R) break;
    int p = c;
    double best = N::abs2(a[c][c]);
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      if (i <= c || i >= R) continue;
      const double v = N::abs2(a[i][c]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    piv[c] = p;
    if (p != c) {
#pragma unroll
      for (int j = 0; j < RM; ++j) {
        if (j >= R) break;
        // dynamic row index p: select by comparison so that the matrix can stay in registers for compile-time R
        T tp = a[c][j];
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          if (i >= R) break;
          if (i == p) {
            const T t = a[i][j];
            a[i][j] = tp;
            tp = t;
          }
        }
        a[c][j] = tp;
      }
    }
    ld += 0.5 * log(best);
    const T pinv = N::inv(a[c][c]);
    a[c][c] = N::one();
#pragma unroll
    for (int j = 0; j < RM; ++j) {
      if (j >= R) break;
      a[c][j] = N::mul(a[c][j], pinv);
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      if (i >= R || i == c) continue;
      const T f = a[i][c];
      a[i][c] = N::zero();
#pragma unroll
      for (int j = 0; j < RM; ++j) {
        if (j >= R) break;
        a[i][j] = N::sub(a[i][j], N::mul(f, a[c][j]));
      }
    }
  }
  // undo the row exchanges as column exchanges, last first
#pragma unroll
  for (int cc = RM - 1; cc >= 0; --cc) {
    if (cc >= R) continue;
    const int p = piv[cc];
    if (p != cc) {
#pragma unroll
      for (int i = 0; i < RM; ++i) {
        if (i >= R) break;
        T tp = a[i][cc];
#pragma unroll
        for (int j = 0; j < RM; ++j) {
          if (j >= R) break;
          if (j == p) {
            const T t = a[i][j];
            a[i][j] = tp;
            tp = t;
          }
        }
        a[i][cc] = tp;
      }
    }
  }
  T* dst = A + k * R * R;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    if (i >= R) break;
#pragma unroll
    for (int j = 0; j < RM; ++j) {
      if (j >= R) break;
      dst[i * R + j] = a[i][j];
    }
  }
  logdet[k] = ld;
}

// any R <= kMaxR: the same elimination with rolled loops on a local-memory copy (coverage of large folded systems, not speed)
template <bool CPLX>
__global__ void __launch_bounds__(128) block_inv_logdet_generic_kernel(const typename Num<CPLX>::T* __restrict__ L, int64_t nm, int R,
                                                                       typename Num<CPLX>::T* __restrict__ A, double* __restrict__ logdet) {
  typedef Num<CPLX> N;
  typedef typename N::T T;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nm) return;
  T a[kMaxR * kMaxR];
  int piv[kMaxR];
  const T* src = L + k * R * R;
#pragma unroll 1
  for (int e = 0; e < R * R; ++e) a[e] = src[e];
  double ld = 0.0;
#pragma unroll 1
  for (int c = 0; c < R; ++c) {
    int p = c;
    double best = N::abs2(a[c * R + c]);
#pragma unroll 1
    for (int i = c + 1; i < R; ++i) {
      const double v = N::abs2(a[i * R + c]);
      if (v > best) best = v, p = i;
    }
    piv[c] = p;
    if (p != c) {
#pragma unroll 1
      for (int j = 0; j < R; ++j) {
        const T t = a[c * R + j];
        a[c * R + j] = a[p * R + j];
        a[p * R + j] = t;
      }
    }
    ld += 0.5 * log(best);
    const T pinv = N::inv(a[c * R + c]);
    a[c * R + c] = N::one();
#pragma unroll 1
    for (int j = 0; j < R; ++j) a[c * R + j] = N::mul(a[c * R + j], pinv);
#pragma unroll 1
    for (int i = 0; i < R; ++i) {
      if (i == c) continue;
      const T f = a[i * R + c];
      a[i * R + c] = N::zero();
#pragma unroll 1
      for (int j = 0; j < R; ++j) a[i * R + j] = N::sub(a[i * R + j], N::mul(f, a[c * R + j]));
    }
  }
#pragma unroll 1
  for (int c = R - 1; c >= 0; --c) {
    const int p = piv[c];
    if (p == c) continue;
#pragma unroll 1
    for (int i = 0; i < R; ++i) {
      const T t = a[i * R + c];
      a[i * R + c] = a[i * R + p];
      a[i * R + p] = t;
    }
  }
  T* dst = A + k * R * R;
#pragma unroll 1
  for (int e = 0; e < R * R; ++e) dst[e] = a[e];
  logdet[k] = ld;
}

template <bool CPLX>
static void launch_block(const void* L, int64_t nm, int R, void* A, double* logdet, cudaStream_t st) {
  typedef typename Num<CPLX>::T T;
  const unsigned blocks = (unsigned)((nm + 127) / 128);
  switch (R) {
    case 1: block_inv_logdet_kernel<CPLX, 1><<<blocks, 128, 0, st>>>((const T*)L, nm, R, (T*)A, logdet); break;
    case 2: block_inv_logdet_kernel<CPLX, 2><<<blocks, 128, 0, st>>>((const T*)L, nm, R, (T*)A, logdet); break;
    case 3: block_inv_logdet_kernel<CPLX, 3><<<blocks, 128, 0, st>>>((const T*)L, nm, R, (T*)A, logdet); break;
    case 4: block_inv_logdet_kernel<CPLX, 4><<<blocks, 128, 0, st>>>((const T*)L, nm, R, (T*)A, logdet); break;
    default: block_inv_logdet_generic_kernel<CPLX><<<blocks, 128, 0, st>>>((const T*)L, nm, R, (T*)A, logdet); break;
  }
}

}  // namespace fgp

extern "C" int fgp_block_inv_logdet(int cplx, const double* L_dev, int64_t nm, int R, double* A_dev, double* logdet_dev, fgp_stream_t stream) {
  using namespace fgp;
  FGP_REQUIRE(L_dev && A_dev && logdet_dev, "block_inv_logdet: null pointer");
  FGP_REQUIRE(nm >= 0 && R >= 1 && R <= kMaxR, "block_inv_logdet: need nm >= 0 and 1 <= R <= %d (got nm=%lld R=%d)", kMaxR, (long long)nm, R);
  if (nm == 0) return FGP_OK;
  if (cplx)
    launch_block<true>(L_dev, nm, R, A_dev, logdet_dev, (cudaStream_t)stream);
  else
    launch_block<false>(L_dev, nm, R, A_dev, logdet_dev, (cudaStream_t)stream);
  FGP_LAUNCH_NAMED("block_inv_logdet", (cudaStream_t)stream);
  return FGP_OK;
}
