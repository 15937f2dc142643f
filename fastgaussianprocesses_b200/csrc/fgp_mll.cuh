// K4: fused eigen-solve + marginal log-likelihood + analytic hyperparameter gradients (and K^-1 y solves).
//
// Reference path being replaced, per fit() iteration (abstract_gp.py:241-296):
//   _kernel_from_parts(k1parts)            abstract_fast_gp.py:181-191  (reads the cached (n,d) parts)
//   ft(k1)                                  abstract_fast_gp.py:197-212  (log2 n torch passes, autograd tape)
//   lam = sqrt(n) lam~ + noise ; log|lam| ; 1/lam ; sum conj(y~) y~ / lam      util.py:285-300,354-370
//   loss.backward()                         a second transform under autograd
// Here: the first column k1 is evaluated on the fly from the points inside the first transform pass (never stored),
// the spectral epilogue (log-det, quadratic form, dL/dlam) runs in the registers/shared memory of the second pass,
// which immediately starts the backward transform on the same tile, and the last pass contracts the back-transformed
// dL/dk1 with the leave-one-out kernel products.  Global traffic per iteration: 16n B written + 16n B read between
// the passes twice (complex workspace, L2-resident up to n = 2^21), 8n B of |y~|^2, and the points twice.
#pragma once
#include "fgp_transform.cuh"

namespace fgp {

constexpr int kRed = 32 * (FGP_MAX_D + 4);  // doubles of reduction scratch

struct MllArgs {
  const void* x;  // lattice: double (n,d); net: int64 (n,d); NULL in generator mode
  UVec z;         // generator mode (lattice): generating vector, delta_ij = frac(phi2(i) z_j) regenerated from the index
  int64_t n;
  int d;
  int t;          // net only
  double tscale;  // net only: 2^-t
  LatPoly P;      // lattice only
  IVec alpha;     // net only
  const double* ysq;    // (B,n)
  const double* scale;  // (B)
  const double* ls;     // (B,d)
  const double* noise;  // (B)
  const double* weights;  // (B,2) (wn, wl) or NULL for (1/2, 1/2): gradients are those of wn*norm + wl*logdet
  void* W;              // workspace: (B,n) complex (lattice) / real (net)
  double* lam;          // optional (B,n) complex / real
  double* partB;        // (B, ctasB, 3)
  double* partC;        // (B, ctasA, d+1)
  double* out;          // (B, d+4)
  int want_grad;
  int l1, l2, ntrA, lntrB, LPA, LPB;
  int ctasA, ctasB;
  FftTables T;
};

struct Hyp {  // per-CTA hyperparameters and first point, staged in shared memory
  double scale, noise;
  double ls[FGP_MAX_D];
  double x0[FGP_MAX_D];     // lattice
  uint64_t xb0[FGP_MAX_D];  // net
};

template <bool NET>
__device__ __forceinline__ void load_hyp(Hyp& H, const MllArgs& a, int b) {
  if (threadIdx.x == 0) {
    H.scale = a.scale[b];
    H.noise = a.noise[b];
  }
  for (int j = threadIdx.x; j < a.d; j += blockDim.x) {
    H.ls[j] = a.ls[(int64_t)b * a.d + j];
    if (a.x) {
      if (NET)
        H.xb0[j] = (uint64_t)((const int64_t*)a.x)[j];
      else
        H.x0[j] = ((const double*)a.x)[j];
    }
  }
}

// generator mode, lattice: x_i - x_0 = frac(phi2(i) z_j) exactly (the shift cancels), phi2(i) = brev64(i) / 2^64.
// The wrap-around product keeps its ceil(log2 n) <= 32 significant bits at the top of the word.
__device__ __forceinline__ double lat_delta_gen(uint64_t rev, uint64_t zj) {
  return (double)(uint32_t)((rev * zj) >> 32) * 0x1.0p-32;
}

// net alpha = 2 without branches on alpha: W_2(delta) - 1 = 3/2 - (5/2) 2^-beta - beta x_f, beta = t - floor(log2 delta)
__device__ __forceinline__ double dnb2_part_a2(uint64_t delta, int t, double tscale) {
  const int fl = 63 - __clzll((long long)delta);  // -1 when delta == 0
  const int beta = t - fl;
  const double xf = __ull2double_rn(delta) * tscale;
  const double pw = __longlong_as_double((long long)(1023 - beta) << 52);  // 2^-beta
  const double r = fma(-(double)beta, xf, fma(-2.5, pw, 1.5));
  return delta ? r : 1.5;
}

// parts of point i against the first point.  A2: every alpha_j == 2 (straight-line code, no per-dimension loop on alpha)
template <int DT, bool NET, bool A2, bool GEN>
__device__ __forceinline__ void point_parts(const MllArgs& a, const Hyp& H, int64_t i, double* p) {
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  const int d = DT > 0 ? DT : a.d;
  if (GEN && !NET) {
    const uint64_t rev = __brevll((unsigned long long)i);
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      if (j >= d) break;
      const double delta = lat_delta_gen(rev, a.z.v[j]);
      if (A2) {
        const double u = delta * (1.0 - delta);
        p[j] = fma(a.P.q[j][2] * u, u, a.P.q[j][0]);
      } else {
        p[j] = lat_part(delta, a.P.q[j], a.P.alpha[j]);
      }
    }
    return;
  }
  if (NET) {
    const int64_t* row = (const int64_t*)a.x + i * d;
    uint64_t xr[DM];
    if (DT > 0 && DT % 2 == 0) {
#pragma unroll
      for (int j = 0; j < DM; j += 2) {
        const longlong2 v = __ldg((const longlong2*)(row + j));
        xr[j] = (uint64_t)v.x;
        xr[j + 1] = (uint64_t)v.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        if (j >= d) break;
        xr[j] = (uint64_t)__ldg(row + j);
      }
    }
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      if (j >= d) break;
      p[j] = A2 ? dnb2_part_a2(xr[j] ^ H.xb0[j], a.t, a.tscale) : dnb2_part(xr[j] ^ H.xb0[j], a.alpha.v[j], a.t);
    }
  } else {
    const double* row = (const double*)a.x + i * d;
    double xr[DM];
    if (DT > 0 && DT % 2 == 0) {
#pragma unroll
      for (int j = 0; j < DM; j += 2) {
        const double2 v = __ldg((const double2*)(row + j));
        xr[j] = v.x;
        xr[j + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        if (j >= d) break;
        xr[j] = __ldg(row + j);
      }
    }
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      if (j >= d) break;
      p[j] = A2 ? lat_part_a2(xr[j] - H.x0[j], a.P.q[j][0], a.P.q[j][2]) : lat_part(xr[j] - H.x0[j], a.P.q[j], a.P.alpha[j]);
    }
  }
}

template <int DT, bool NET, bool A2, bool GEN>
__device__ __forceinline__ double point_k1(const MllArgs& a, const Hyp& H, int64_t i) {
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  const int d = DT > 0 ? DT : a.d;
  double p[DM];
  point_parts<DT, NET, A2, GEN>(a, H, i, p);
  double k = H.scale;
#pragma unroll
  for (int j = 0; j < DM; ++j) {
    if (j >= d) break;
    k *= fma(H.ls[j], p[j], 1.0);
  }
  return k;
}

// acc[0] += w*k1 ; acc[1+j] += w * dk1/dls_j   (leave-one-out products: factors may cross zero)
template <int DT, bool NET, bool A2, bool GEN>
__device__ __forceinline__ void point_grad(const MllArgs& a, const Hyp& H, int64_t i, double w, double* acc) {
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  const int d = DT > 0 ? DT : a.d;
  double p[DM], left[DM];
  point_parts<DT, NET, A2, GEN>(a, H, i, p);
  double pre = H.scale;
#pragma unroll
  for (int j = 0; j < DM; ++j) {
    if (j >= d) break;
    left[j] = pre;
    pre *= fma(H.ls[j], p[j], 1.0);
  }
  acc[0] = fma(w, pre, acc[0]);
  double suf = w;
#pragma unroll
  for (int j = DM - 1; j >= 0; --j) {
    if (j >= d) continue;
    acc[1 + j] = fma(left[j] * suf, p[j], acc[1 + j]);
    suf *= fma(H.ls[j], p[j], 1.0);
  }
}

// spectral epilogue: lam -> (norm, logdet, dnoise) partial sums and G = dL/dlam (stored in place of lam)
__device__ __forceinline__ double2 spectral_c(double2 lam, double ysq, double wn, double wl, double* s) {
  const double a = lam.x, b = lam.y;
  const double m2 = fma(a, a, b * b);
  const double inv = 1.0 / m2;
  s[0] = fma(ysq, a * inv, s[0]);
  s[1] += 0.5 * log(m2);
  const double yi2 = wn * ysq * inv * inv;
  const double li = wl * inv;
  const double ga = fma(yi2, fma(b, b, -a * a), a * li);
  const double gb = fma(-2.0 * a * b, yi2, b * li);
  s[2] += ga;
  return make_double2(ga, gb);
}
__device__ __forceinline__ double spectral_r(double lam, double ysq, double wn, double wl, double* s) {
  const double inv = 1.0 / lam;
  s[0] = fma(ysq, inv, s[0]);
  s[1] += log(fabs(lam));
  const double g = inv * (wl - wn * ysq * inv);
  s[2] += g;
  return g;
}

template <int NV>
__device__ __forceinline__ void reduce_store(double* v, int nv, double* red, double* dst) {
  // runtime nv <= NV
  for (int k0 = 0; k0 < nv; k0 += 4) {
    double t[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = (k0 + k < nv) ? v[k0 + k] : 0.0;
    block_sum<4>(t, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k0 + k < nv) dst[k0 + k] = t[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// single-pass kernel: one CTA per hyperparameter set, n <= block capacity
// ------------------------------------------------------------------------------------------------------------
template <int DT, bool NET, bool A2, bool GEN>
__global__ void __launch_bounds__(256, 2) mll_single_kernel(MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ Hyp H;
  __shared__ double red[kRed];
  const int b = blockIdx.x;
  const int n = (int)a.n;
  const int l = a.l1;
  const int LP = a.LPA;
  const int d = DT > 0 ? DT : a.d;
  load_hyp<NET>(H, a, b);
  __syncthreads();
  double2* smc = (double2*)smraw;
  double* smr = (double*)smraw;
  const double c = H.scale;  // DC guess removed before the transform (role of abstract_fast_gp.py:209-211)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double k1 = point_k1<DT, NET, A2, GEN>(a, H, i) - c;
    if (NET)
      smr[padidx(i)] = k1;
    else
      smc[padidx(i)] = make_double2(k1, 0.0);
  }
  __syncthreads();
  if (NET)
    block_wht(smr, l, 1, LP);
  else
    block_fft_fwd(smc, l, 1, LP, a.T.stage);
  double s[3] = {0.0, 0.0, 0.0};
  const double* ysq = a.ysq + (int64_t)b * n;
  const double wn = a.weights ? a.weights[2 * b] : 0.5, wl = a.weights ? a.weights[2 * b + 1] : 0.5;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    if (NET) {
      double lam = smr[padidx(k)] + H.noise;
      if (k == 0) lam += c * (double)n;
      if (a.lam) a.lam[(int64_t)b * n + k] = lam;
      smr[padidx(k)] = spectral_r(lam, ysq[k], wn, wl, s);
    } else {
      double2 lam = smc[padidx(k)];
      lam.x += H.noise;
      if (k == 0) lam.x += c * (double)n;
      if (a.lam) ((double2*)a.lam)[(int64_t)b * n + k] = lam;
      smc[padidx(k)] = spectral_c(lam, ysq[k], wn, wl, s);
    }
  }
  double* out = a.out + (int64_t)b * (d + 4);
  reduce_store<3>(s, 3, red, out);
  if (!a.want_grad) return;
  __syncthreads();
  if (NET)
    block_wht(smr, l, 1, LP);
  else
    block_fft_inv(smc, l, 1, LP, a.T.stage);
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  double acc[DM + 1];
#pragma unroll
  for (int j = 0; j <= DM; ++j) acc[j] = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double w = NET ? smr[padidx(i)] : smc[padidx(i)].x;
    point_grad<DT, NET, A2, GEN>(a, H, i, w, acc);
  }
  acc[0] /= H.scale;
  reduce_store<DM + 1>(acc, d + 1, red, out + 3);
}

// ------------------------------------------------------------------------------------------------------------
// two-pass kernels
// ------------------------------------------------------------------------------------------------------------
// pass A: k1 on the fly -> contiguous block transform -> inter-pass twiddle -> workspace
template <int DT, bool NET, bool A2, bool GEN>
__global__ void __launch_bounds__(256, 2) mll_passA_kernel(MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ Hyp H;
  const int b = blockIdx.y;
  load_hyp<NET>(H, a, b);
  __syncthreads();
  double2* smc = (double2*)smraw;
  double* smr = (double*)smraw;
  const int l1 = a.l1, l2 = a.l2, ntr = a.ntrA, LP = a.LPA;
  const int64_t blk0 = (int64_t)blockIdx.x * ntr;
  const int cnt = ntr << l1;
  const int64_t g0 = blk0 << l1;
  const int qmask = (1 << l1) - 1;
  const double c = H.scale;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    const double k1 = point_k1<DT, NET, A2, GEN>(a, H, g0 + e) - c;
    const int si = (e >> l1) * LP + padidx(e & qmask);
    if (NET)
      smr[si] = k1;
    else
      smc[si] = make_double2(k1, 0.0);
  }
  __syncthreads();
  if (NET) {
    block_wht(smr, l1, ntr, LP);
    double* W = (double*)a.W + (int64_t)b * a.n;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) W[g0 + e] = smr[(e >> l1) * LP + padidx(e & qmask)];
  } else {
    block_fft_fwd(smc, l1, ntr, LP, a.T.stage);
    double2* W = (double2*)a.W + (int64_t)b * a.n;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
      const int tr = e >> l1, q = e & qmask;
      const uint32_t bb = (uint32_t)((blk0 + tr) & ((1 << l2) - 1));
      W[g0 + e] = cmul(smc[tr * LP + padidx(q)], twiddle_n(a.T, brev_bits(bb, l2) * (uint32_t)q));
    }
  }
}

// pass B: strided columns -> forward transform -> spectral epilogue -> backward transform of dL/dlam -> workspace
template <bool NET>
__global__ void __launch_bounds__(256, 2) mll_passB_kernel(MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red[kRed];
  __shared__ double s_noise, s_scale;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    s_noise = a.noise[b];
    s_scale = a.scale[b];
  }
  double2* smc = (double2*)smraw;
  double* smr = (double*)smraw;
  const int l1 = a.l1, l2 = a.l2, lntr = a.lntrB, LP = a.LPB;
  const int ntr = 1 << lntr;
  const int q0 = blockIdx.x << lntr;
  const int cnt = ntr << l2;
  const int64_t boff = (int64_t)b * a.n;
  if (NET) {
    const double* W = (const double*)a.W + boff + q0;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
      const int cc = e & (ntr - 1), r = e >> lntr;
      smr[cc * LP + padidx(r)] = W[((int64_t)r << l1) + cc];
    }
  } else {
    const double2* W = (const double2*)a.W + boff + q0;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
      const int cc = e & (ntr - 1), r = e >> lntr;
      smc[cc * LP + padidx(r)] = W[((int64_t)r << l1) + cc];
    }
  }
  __syncthreads();
  if (NET)
    block_wht(smr, l2, ntr, LP);
  else
    block_fft_fwd(smc, l2, ntr, LP, a.T.stage);
  double s[3] = {0.0, 0.0, 0.0};
  const double* ysq = a.ysq + boff + q0;
  const double noise = s_noise;
  const double wn = a.weights ? a.weights[2 * b] : 0.5, wl = a.weights ? a.weights[2 * b + 1] : 0.5;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int cc = e & (ntr - 1), r = e >> lntr;
    const int64_t k = ((int64_t)r << l1) + cc;  // + q0
    const int si = cc * LP + padidx(r);
    if (NET) {
      double lam = smr[si] + noise;
      if (k + q0 == 0) lam += s_scale * (double)a.n;
      if (a.lam) a.lam[boff + q0 + k] = lam;
      smr[si] = spectral_r(lam, ysq[k], wn, wl, s);
    } else {
      double2 lam = smc[si];
      lam.x += noise;
      if (k + q0 == 0) lam.x += s_scale * (double)a.n;
      if (a.lam) ((double2*)a.lam)[boff + q0 + k] = lam;
      smc[si] = spectral_c(lam, ysq[k], wn, wl, s);
    }
  }
  reduce_store<3>(s, 3, red, a.partB + ((int64_t)b * a.ctasB + blockIdx.x) * 3);
  if (!a.want_grad) return;
  __syncthreads();
  if (NET) {
    block_wht(smr, l2, ntr, LP);
    double* W = (double*)a.W + boff + q0;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
      const int cc = e & (ntr - 1), r = e >> lntr;
      W[((int64_t)r << l1) + cc] = smr[cc * LP + padidx(r)];
    }
  } else {
    block_fft_inv(smc, l2, ntr, LP, a.T.stage);
    double2* W = (double2*)a.W + boff + q0;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
      const int cc = e & (ntr - 1), r = e >> lntr;
      const double2 w = twiddle_n(a.T, brev_bits((uint32_t)r, l2) * (uint32_t)(q0 + cc));
      W[((int64_t)r << l1) + cc] = cmulc(w, smc[cc * LP + padidx(r)]);
    }
  }
}

// pass C: contiguous blocks of the back-transformed dL/dlam -> inverse block transform -> contraction with dk1/dtheta
template <int DT, bool NET, bool A2, bool GEN>
__global__ void __launch_bounds__(256, 2) mll_passC_kernel(MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ Hyp H;
  __shared__ double red[kRed];
  const int b = blockIdx.y;
  const int d = DT > 0 ? DT : a.d;
  load_hyp<NET>(H, a, b);
  double2* smc = (double2*)smraw;
  double* smr = (double*)smraw;
  const int l1 = a.l1, ntr = a.ntrA, LP = a.LPA;
  const int64_t blk0 = (int64_t)blockIdx.x * ntr;
  const int cnt = ntr << l1;
  const int64_t g0 = blk0 << l1;
  const int qmask = (1 << l1) - 1;
  if (NET) {
    const double* W = (const double*)a.W + (int64_t)b * a.n;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) smr[(e >> l1) * LP + padidx(e & qmask)] = W[g0 + e];
  } else {
    const double2* W = (const double2*)a.W + (int64_t)b * a.n;
    for (int e = threadIdx.x; e < cnt; e += blockDim.x) smc[(e >> l1) * LP + padidx(e & qmask)] = W[g0 + e];
  }
  __syncthreads();
  if (NET)
    block_wht(smr, l1, ntr, LP);
  else
    block_fft_inv(smc, l1, ntr, LP, a.T.stage);
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  double acc[DM + 1];
#pragma unroll
  for (int j = 0; j <= DM; ++j) acc[j] = 0.0;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int si = (e >> l1) * LP + padidx(e & qmask);
    const double w = NET ? smr[si] : smc[si].x;
    point_grad<DT, NET, A2, GEN>(a, H, g0 + e, w, acc);
  }
  reduce_store<DM + 1>(acc, d + 1, red, a.partC + ((int64_t)b * a.ctasA + blockIdx.x) * (d + 1));
}

// finalize: deterministic reduction of the per-CTA partial sums
static __global__ void __launch_bounds__(256, 2) mll_finalize_kernel(MllArgs a) {
  __shared__ double red[kRed];
  const int b = blockIdx.x;
  const int d = a.d;
  double* out = a.out + (int64_t)b * (d + 4);
  {
    double s[3] = {0.0, 0.0, 0.0};
    const double* p = a.partB + (int64_t)b * a.ctasB * 3;
    for (int c = threadIdx.x; c < a.ctasB; c += blockDim.x) {
      s[0] += p[c * 3 + 0];
      s[1] += p[c * 3 + 1];
      s[2] += p[c * 3 + 2];
    }
    reduce_store<3>(s, 3, red, out);
  }
  if (!a.want_grad) return;
  const double* p = a.partC + (int64_t)b * a.ctasA * (d + 1);
  const double inv_scale = 1.0 / a.scale[b];
  for (int j = 0; j <= d; ++j) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = threadIdx.x; c < a.ctasA; c += blockDim.x) v[0] += p[(int64_t)c * (d + 1) + j];
    block_sum<4>(v, red);
    if (threadIdx.x == 0) out[3 + j] = j == 0 ? v[0] * inv_scale : v[0];
    __syncthreads();
  }
}

template <typename K>
static int set_smem_attr(K kernel, size_t bytes) {
  if (bytes > 24 * 1024) {  // static shared memory counts towards the 48 KiB default limit
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(%zu bytes): %s", bytes, cudaGetErrorString(e));
      return FGP_ECUDA;
    }
  }
  return FGP_OK;
}

template <int DT, bool NET, bool A2, bool GEN>
static int launch_mll(const MllArgs& a, const PassGeom& g, int B, cudaStream_t st) {
  int rc;
  if (g.l2 == 0) {
    if ((rc = set_smem_attr(mll_single_kernel<DT, NET, A2, GEN>, g.smemA))) return rc;
    mll_single_kernel<DT, NET, A2, GEN><<<B, g.threads, g.smemA, st>>>(a);
    FGP_LAUNCH_NAMED("mll_single", st);
    return FGP_OK;
  }
  if ((rc = set_smem_attr(mll_passA_kernel<DT, NET, A2, GEN>, g.smemA))) return rc;
  mll_passA_kernel<DT, NET, A2, GEN><<<dim3(a.ctasA, B), g.threads, g.smemA, st>>>(a);
  FGP_LAUNCH_NAMED("mll_passA", st);
  if ((rc = set_smem_attr(mll_passB_kernel<NET>, g.smemB))) return rc;
  mll_passB_kernel<NET><<<dim3(a.ctasB, B), g.threads, g.smemB, st>>>(a);
  FGP_LAUNCH_NAMED("mll_passB", st);
  if (a.want_grad) {
    if ((rc = set_smem_attr(mll_passC_kernel<DT, NET, A2, GEN>, g.smemA))) return rc;
    mll_passC_kernel<DT, NET, A2, GEN><<<dim3(a.ctasA, B), g.threads, g.smemA, st>>>(a);
    FGP_LAUNCH_NAMED("mll_passC", st);
  }
  mll_finalize_kernel<<<B, 256, 0, st>>>(a);
  FGP_LAUNCH_NAMED("mll_finalize", st);
  return FGP_OK;
}

template <bool NET, bool GEN>
static int dispatch_mll(const MllArgs& a, const PassGeom& g, int B, bool all2, cudaStream_t st) {
  if (all2) {
    switch (a.d) {
      case 2: return launch_mll<2, NET, true, GEN>(a, g, B, st);
      case 4: return launch_mll<4, NET, true, GEN>(a, g, B, st);
      case 8: return launch_mll<8, NET, true, GEN>(a, g, B, st);
      case 16: return launch_mll<16, NET, true, GEN>(a, g, B, st);
      default: break;
    }
  }
  if (GEN) return launch_mll<0, NET, false, GEN>(a, g, B, st);
  switch (a.d) {
    case 2: return launch_mll<2, NET, false, GEN>(a, g, B, st);
    case 4: return launch_mll<4, NET, false, GEN>(a, g, B, st);
    case 8: return launch_mll<8, NET, false, GEN>(a, g, B, st);
    default: return launch_mll<0, NET, false, GEN>(a, g, B, st);
  }
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

template <bool NET>
static int mll_common(const uint64_t* z_host, const void* x, int64_t n, int d, const int* alpha_host, int t, int B, const double* ysq,
                      const double* scale, const double* ls, const double* noise, const double* weights, const void* table, void* workspace,
                      double* lam, double* out, int want_grad, fgp_stream_t stream) {
  const bool net = NET;
  FGP_REQUIRE((x || z_host) && alpha_host && ysq && scale && ls && noise && out, "mll_grad: null pointer");
  FGP_REQUIRE(!(NET && z_host), "mll_grad: generator mode is lattice-only");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D, "mll_grad: d=%d outside 1..%d", d, FGP_MAX_D);
  FGP_REQUIRE(B >= 1 && B <= 65535, "mll_grad: B=%d outside 1..65535", B);
  FGP_REQUIRE(is_pow2(n) && ilog2(n) <= (net ? FGP_MAX_LOG2N_WHT : FGP_MAX_LOG2N_FFT),
              "mll_grad: n=%lld must be a power of two <= 2^%d", (long long)n, net ? FGP_MAX_LOG2N_WHT : FGP_MAX_LOG2N_FFT);
  MllArgs a;
  memset(&a, 0, sizeof(a));
  a.x = z_host ? nullptr : x;
  if (z_host) {
    FGP_REQUIRE(ilog2(n) <= 32, "mll_grad: generator mode needs n <= 2^32");
    for (int j = 0; j < d; ++j) a.z.v[j] = z_host[j];
  }
  a.n = n;
  a.d = d;
  a.t = t;
  a.tscale = ldexp(1.0, -t);
  if (net) {
    FGP_REQUIRE(t >= 1 && t < 64, "mll_grad: t outside 1..63");
    for (int j = 0; j < d; ++j) {
      a.alpha.v[j] = alpha_host[j];
      FGP_REQUIRE(alpha_host[j] >= 1 && alpha_host[j] <= 4, "mll_grad: net alpha[%d]=%d outside 1..4", j, alpha_host[j]);
    }
  } else {
    FGP_REQUIRE(table, "mll_grad: null twiddle table");
    int rc = fill_lat_poly(alpha_host, d, &a.P);
    if (rc) return rc;
    a.T = make_tables(table);
  }
  const PassGeom g = make_geom(n, !net);
  a.ysq = ysq;
  a.scale = scale;
  a.ls = ls;
  a.noise = noise;
  a.weights = weights;
  a.lam = lam;
  a.out = out;
  a.want_grad = want_grad;
  a.l1 = g.l1;
  a.l2 = g.l2;
  a.ntrA = g.ntrA;
  a.lntrB = ilog2(g.ntrB);
  a.LPA = g.LPA;
  a.LPB = g.LPB;
  a.ctasA = (int)g.ctasA;
  a.ctasB = (int)g.ctasB;
  if (g.l2) {
    FGP_REQUIRE(workspace, "mll_grad: null workspace");
    const size_t wbytes = align256((size_t)B * n * (net ? sizeof(double) : sizeof(double2)));
    const size_t pb = align256((size_t)B * a.ctasB * 3 * sizeof(double));
    a.W = workspace;
    a.partB = (double*)((char*)workspace + wbytes);
    a.partC = (double*)((char*)workspace + wbytes + pb);
  }
  bool all2 = true;
  for (int j = 0; j < d; ++j) all2 = all2 && alpha_host[j] == 2;
  if (!NET && z_host) return dispatch_mll<false, true>(a, g, B, all2, (cudaStream_t)stream);
  return dispatch_mll<NET, false>(a, g, B, all2, (cudaStream_t)stream);
}

}  // namespace fgp
