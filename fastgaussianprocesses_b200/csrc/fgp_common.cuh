// Shared helpers for libfgp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "../../include/fgp_b200.h"

namespace fgp {

void set_error(const char* fmt, ...);
void count_launch();
void prof_mark(const char* name, cudaStream_t st);  // no-op unless fgp_profile_begin() armed it
int sm_count();

#define FGP_REQUIRE(cond, ...)                    \
  do {                                            \
    if (!(cond)) {                                \
      fgp::set_error(__VA_ARGS__);                \
      return FGP_EINVAL;                          \
    }                                             \
  } while (0)

#define FGP_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      fgp::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return FGP_ECUDA;                                                                  \
    }                                                                                    \
  } while (0)

#define FGP_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    fgp::count_launch();                                                                 \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      fgp::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return FGP_ECUDA;                                                                  \
    }                                                                                    \
  } while (0)

// launch check that also drops a timing mark (per-kernel durations for bench.py's roofline block)
#define FGP_LAUNCH_NAMED(name, st)       \
  do {                                   \
    FGP_LAUNCH_CHECK();                  \
    fgp::prof_mark(name, st);            \
  } while (0)

static inline bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }
static inline int ilog2(int64_t n) {
  int m = 0;
  while ((int64_t(1) << m) < n) ++m;
  return m;
}

// ---- small by-value parameter blocks (copied into the kernel parameter space; no H2D copies)
struct DVec {
  double v[FGP_MAX_D];
};
struct IVec {
  int v[FGP_MAX_D];
};
struct UVec {
  uint64_t v[FGP_MAX_D];
};

// Bernoulli polynomial B_{2a}(x), written as a polynomial in u = x(1-x) (B_{2a}(1-x) = B_{2a}(x)):
//   c_a * B_{2a}(x) = sum_{p=0..a} q[a][p] u^p   with c_a = (-1)^(a+1) (2 pi)^(2a) / (2a)!  folded in.
// Filled on the host once (exact rational arithmetic), see fgp_kernel_eval.cu.
struct LatPoly {
  // q[j][p], p = 0..alpha_j, for each dimension j
  double q[FGP_MAX_D][FGP_MAX_ALPHA + 1];
  int alpha[FGP_MAX_D];
};
int fill_lat_poly(const int* alpha_host, int d, LatPoly* out);  // returns FGP_OK / FGP_EINVAL

// ---- device helpers
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // conj(a) * b
  return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.x, b.y, -a.y * b.x));
}

// L2 prefetches (no register destination, nothing to wait for).  A fit iteration is a chain of dependent phases, each one memory
// round trip long: data that a LATER phase reads from HBM (|y~|^2 in the spectral epilogue of pass B, the twiddle tables) is requested
// at kernel entry so that the phase finds it in L2.  l2_prefetch_bulk is the TMA form (cp.async.bulk.prefetch.L2: one instruction for a
// contiguous range, 16-byte aligned, a multiple of 16 bytes); l2_prefetch_line takes one 128-byte line.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void l2_prefetch_line(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(__cvta_generic_to_global(p)));
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of NV values per thread; result valid in thread 0.  red: shared scratch of >= NV*32 doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) red[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < nwarp ? red[k * 32 + lane] : 0.0;
      v[k] = warp_sum(t);
    }
  }
}

// ---- kernel-part evaluation (device), shared by K2 / K4 / K5
// lattice: c*B_{2a}(frac(delta)) from delta in (-1,1]: a = |delta| works because B_{2a}(1-x) = B_{2a}(x)
__device__ __forceinline__ double lat_part(double delta, const double* __restrict__ q, int alpha) {
  double a = fabs(delta);
  a = a - floor(a);            // delta = +-1 -> 0
  double u = a * (1.0 - a);
  double r = q[alpha];
  for (int p = alpha - 1; p >= 0; --p) r = fma(r, u, q[p]);
  return r;
}
__device__ __forceinline__ double lat_part_a2(double delta, double q0, double q2) {
  double a = fabs(delta);
  a = a - floor(a);
  double u = a * (1.0 - a);
  return fma(q2 * u, u, q0);
}

// net: W_alpha(delta) - 1 for a t-bit integer delta (SURVEY App. B.2 closed forms)
__device__ __forceinline__ double dnb2_part(uint64_t delta, int alpha, int t) {
  if (alpha == 1) {
    if (delta == 0) return 1.0;
    int fl = 63 - __clzll((long long)delta);
    return 1.0 - 3.0 * exp2((double)(fl - t));  // 6*(1/6 - 2^(fl-t-1))
  }
  if (delta == 0) return alpha == 2 ? 1.5 : (alpha == 3 ? 43.0 / 18.0 - 1.0 : 701.0 / 294.0 - 1.0);
  const int fl = 63 - __clzll((long long)delta);
  const int ibeta = t - fl;
  const double beta = (double)ibeta;
  const double xf = __ull2double_rn(delta) * exp2((double)-t);
  const double p1 = 1.0 - exp2(-beta);
  if (alpha == 2) return fma(-beta, xf, 2.5 * p1) - 1.0;
  const double p2 = 1.0 - exp2(-2.0 * beta);
  if (alpha == 3) return (beta * xf * xf - 5.0 * p1 * xf + (43.0 / 18.0) * p2) - 1.0;
  const double p3 = 1.0 - exp2(-3.0 * beta);
  // s = sum_{a>=0} (-1)^{x_{a+1}} 8^-a = 8/7 - 2 sum_a x_{a+1} 8^-a
  double s = 8.0 / 7.0;
  double w = 2.0;
  const int na = t < 22 ? t : 22;
  for (int a = 0; a < na; ++a) {
    if ((delta >> (t - 1 - a)) & 1ull) s -= w;
    w *= 0.125;
  }
  const double x2 = xf * xf;
  return (-(2.0 / 3.0) * beta * x2 * xf + 5.0 * p1 * x2 - (43.0 / 9.0) * p2 * xf + (701.0 / 294.0) * p3 +
          beta * (s / 48.0 - 1.0 / 42.0)) - 1.0;
}

// net alpha = 2 without branches on alpha: W_2(delta) - 1 = 3/2 - (5/2) 2^-beta - beta x_f, beta = t - floor(log2 delta).
// t <= 52 (delta < 2^52, this package's default nets): the integer converts EXACTLY through the 2^52 magic number (one OR on the high
// word, one DADD) and floor(log2 delta) is the exponent field of that double -- no 64-bit count-leading-zeros and no 64-bit int-to-double
// conversion, both of which are multi-instruction sequences (the inner loop of the net post_mean kernel was 55 instructions per
// (test, train, dimension), issue-bound; about 18 this way).
__device__ __forceinline__ double dnb2_part_a2_t52(uint64_t delta, int t, double tscale) {  // requires t <= 52
  const double dd = __longlong_as_double((long long)(delta | 0x4330000000000000ull)) - 4503599627370496.0;
  const int fl = (int)((unsigned long long)__double_as_longlong(dd) >> 52) - 1023;  // floor(log2 delta), delta > 0
  const int beta = t - fl;
  const double xf = dd * tscale;
  const double pw = __longlong_as_double((long long)(1023 - beta) << 52);  // 2^-beta
  const double r = fma(-(double)beta, xf, fma(-2.5, pw, 1.5));
  return delta ? r : 1.5;
}
__device__ __forceinline__ double dnb2_part_a2(uint64_t delta, int t, double tscale) {
  const int fl = 63 - __clzll((long long)delta);  // -1 when delta == 0
  const int beta = t - fl;
  const double xf = __ull2double_rn(delta) * tscale;
  const double pw = __longlong_as_double((long long)(1023 - beta) << 52);  // 2^-beta
  const double r = fma(-(double)beta, xf, fma(-2.5, pw, 1.5));
  return delta ? r : 1.5;
}

// float test point -> t-bit integer, fast_gp_digital_net_b2.py:270-271: floor((x % 1) * 2^t)
__device__ __forceinline__ uint64_t dnb2_to_b(double x, int t) {
  double f = x - floor(x);
  return (uint64_t)__double2ull_rd(f * exp2((double)t));
}

}  // namespace fgp
