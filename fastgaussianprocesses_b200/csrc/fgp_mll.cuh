// K4: fused eigen-solve + marginal log-likelihood + analytic hyperparameter gradients.
//
// Reference path being replaced, per fit() iteration (abstract_gp.py:241-296):
//   _kernel_from_parts(k1parts)            abstract_fast_gp.py:181-191  (reads the cached (n,d) parts)
//   ft(k1)                                  abstract_fast_gp.py:197-212  (log2 n torch passes, autograd tape)
//   lam = sqrt(n) lam~ + noise ; log|lam| ; 1/lam ; sum conj(y~) y~ / lam      util.py:285-300,354-370
//   loss.backward()                         a second transform under autograd
// Here, three kernels (one when n fits a CTA):
//   pass A  first column k1 evaluated on the fly INTO the registers of the first transform round (from the points, or
//           in generator mode from the point index alone), contiguous block transform, inter-pass twiddle -> W
//   pass B  column transform; its top round leaves lam in registers, where the spectral epilogue (log-det, quadratic
//           form, dL/dlam) runs and the backward transform starts on the same registers -> W
//   pass C  inverse block transform whose last round hands dL/dk1 of 16 consecutive points to the contraction with
//           dk1/dtheta (leave-one-out products), nothing is stored.
// Global traffic per iteration: W written and read twice (16n B complex / 8n B real each way, L2-resident up to
// n = 2^21), 8n B of |y~|^2, and the points twice -- or not at all in generator mode.
#pragma once
#include "fgp_transform.cuh"
#include "fgp_fit.cuh"

#ifndef FGP_K1_SIGMA
#define FGP_K1_SIGMA 1  // six-slot first-column evaluation for lattice / alpha = 2 / generator mode (Hyp::sig); 0: the nine-slot form
#endif

namespace fgp {

constexpr int kRed = 32 * 4;  // doubles of reduction scratch (block_sum<4>)

struct MllArgs {
  const void* x;  // lattice: double (n,d); net: int64 (n,d); NULL in generator mode
  UVec z;         // generator mode (lattice): generating vector, delta_ij = frac(phi2(i) z_j) regenerated from the index
  const uint64_t* C;  // generator mode (net): device (d, mmax) generating-matrix columns, xb_i ^ xb_0 = XOR_{k in bits(i)} C[j][k]
  int mmax;
  int64_t n;
  int d;
  int t;          // net only
  double tscale;  // net only: 2^-t
  LatPoly P;      // lattice only
  IVec alpha;     // net only
  const double* ysq;      // (B,n)
  const double* scale;    // (B)
  const double* ls;       // (B,d)
  const double* noise;    // (B)
  const double* weights;  // (B,2) (wn, wl) or NULL for (1/2, 1/2): gradients are those of wn*norm + wl*logdet
  void* W;                // workspace: (B,n) complex (lattice) / real (net)
  double* lam;            // optional (B,n) complex / real
  double* partB;          // (B, ctasB, 3)
  double* partC;          // (B, ctasA, d+1)
  double* out;            // (B, d+4)
  int want_grad;
  int l1, l2, lntrA, lntrB, LPA, LPB;
  int ctasA, ctasB;
  FftTables T;
  int tab_off;    // net generator mode: byte offset of the XOR-fold tables behind the tile in dynamic shared memory
  int has_fit;    // fused fit iteration: the last CTA to finish reduces the partial sums and runs the fit step
  // Lattice two-pass kernels, half-spectrum mode.  k1 is real and even on Z_n (B_2a(1-t) = B_2a(t) in every dimension), and so
  // are lam, dL/dlam (|y~|^2 is even for real y) and the back-transformed dL/dk1.  In the four-step layout W[b][q] (block row
  // b of residue class r = rev(b), column q) that reads  W[rev(L2-r)][q] = conj(W[b][q])  after pass A,  lam_{n-k} = lam_k,
  // and  W[b][L1-q] = conj(W[b][q])  after pass B's backward half.  So pass A and pass C run only the L2/2+1 blocks with
  // r <= L2/2, pass B only the columns q <= L1/2; mirrored loads are conjugated reads and mirrored contributions weights 2.
  int hs;
  // persistent cooperative kernel (mll_coop_kernel): control words {grid-barrier counter, exit counter, error flag} in the
  // workspace (zero between launches: the last CTA to leave resets them) and the number of fit iterations of one launch
  unsigned int* bar;
  int iters;
  int pdl;  // programmatic dependent launch mode of the per-pass kernels (pdl_mode())
#ifdef FGP_TIMING
  long long* stamps;  // tools-only build: phase stamps of the persistent kernel
#endif
  FitLayout fit;
};

// Programmatic dependent launch: the three kernels of an iteration (and the iterations of a fused fit loop) form one chain
// in the stream, each about one wave long, so the launch latency and CTA ramp-up of a kernel are a visible share of it.
// With the attribute the next kernel's CTAs become resident as soon as every CTA of the current one has started
// (pdl_prologue triggers at once); they wait in cudaGridDependencySynchronize() for its memory before touching anything.
int pdl_mode();  // fgp_mll_passb.cu: 0 = off (default), FGP_PDL=1: every CTA triggers at entry, FGP_PDL=2: the trigger is the CTA's exit
template <typename K>
static cudaError_t launch_chain(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const MllArgs& a) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, a);
}
__device__ __forceinline__ void pdl_prologue(const MllArgs& a) {
  if (a.pdl == 1) cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();  // returns at once in a launch without the attribute
}

struct Hyp {  // per-CTA hyperparameters and first point, staged in shared memory
  double scale, noise;
  double ls[FGP_MAX_D];
  double x0[FGP_MAX_D];     // lattice
  // lattice, alpha = 2, generator mode: 1 + ls_j c B_4(t) = A_j (1 - h^2), h = (sig_j t)(sig_j - sig_j t), A_j = 1 + ls_j q0_j,
  // sig_j = (ls_j |q2_j| / A_j)^(1/4); sig32 = sig 2^-32 (the integer-to-[0,1) scaling folded in), pref = scale prod_j A_j.
  // Six FP64 issue slots per (point, dimension) in the first-column evaluation instead of nine (post_mean's form of the kernel).
  double sig[FGP_MAX_D], sig32[FGP_MAX_D], pref;
  uint64_t xb0[FGP_MAX_D];  // net
  // net generator mode: xb_i ^ xb_0 of tile element e = TA[j][e & 63] ^ TB[j][e >> 6] (shared-memory tables of XOR folds,
  // dimension-major so that consecutive threads read consecutive words)
  const uint64_t* TA;
  const uint64_t* TB;
  int64_t tile_base;
  int nbB;  // entries per dimension of TB
};

// XOR fold of generating-matrix columns over the set bits of v
__device__ __forceinline__ uint64_t dnb2_fold(const uint64_t* __restrict__ Cj, uint64_t v) {
  uint64_t r = 0;
  while (v) {
    const int k = __ffsll((long long)v) - 1;
    r ^= __ldg(Cj + k);
    v &= v - 1;
  }
  return r;
}
// build the generator tables of a tile of 2^tile_log points starting at tile_base (a multiple of the tile size).
// The generating-matrix columns (d * mmax words) are staged in shared memory by ONE parallel load; the folds then run on shared memory with
// predicated XORs (no data-dependent loop of dependent global loads: that was 4 us of prologue per CTA at n = 2^20, up to 20 serial L2
// round trips for the high bits of the tile base -- -DFGP_TIMING stamps, profiles/README.md).
__device__ __forceinline__ void dnb2_build_tables(Hyp& H, const MllArgs& a, uint64_t* tab, int64_t tile_base, int tile_log) {
  const int d = a.d, mmax = a.mmax;
  const int nb = tile_log > 6 ? 1 << (tile_log - 6) : 1;
  uint64_t* TA = tab;
  uint64_t* TB = tab + 64 * d;
  uint64_t* sC = TB + nb * d;  // (d, mmax) copy of the generating matrices
  for (int e = threadIdx.x; e < d * mmax; e += blockDim.x) sC[e] = __ldg(a.C + e);
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * d; e += blockDim.x) {
    const int j = e >> 6, v = e & 63;
    const uint64_t* Cj = sC + j * mmax;
    uint64_t r = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k < mmax) r ^= ((v >> k) & 1) ? Cj[k] : 0ull;
    TA[e] = r;
  }
  for (int e = threadIdx.x; e < nb * d; e += blockDim.x) {
    const int j = e / nb, h = e - j * nb;
    const uint64_t* Cj = sC + j * mmax;
    const uint64_t v = ((uint64_t)tile_base | ((uint64_t)h << 6)) >> 6;  // bits 6 and up of the point index
    uint64_t r = 0;
#pragma unroll 8
    for (int k = 6; k < mmax; ++k) r ^= ((v >> (k - 6)) & 1ull) ? Cj[k] : 0ull;
    TB[e] = r;
  }
  if (threadIdx.x == 0) {
    H.TA = TA;
    H.TB = TB;
    H.tile_base = tile_base;
    H.nbB = nb;
  }
}
static inline size_t dnb2_table_bytes(int d, int tile_log, int mmax) {
  return (size_t)(64 + (tile_log > 6 ? 1 << (tile_log - 6) : 1) + mmax) * d * sizeof(uint64_t);
}

struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};
// hook(): independent loads of the caller (pass C's tile), issued by warp 0 right AFTER its hyperparameter loads are in flight and before it
// waits for them, so that the hyperparameters -- the critical path of the prologue -- are not queued behind them
template <bool NET, typename Hook = NoHook>
__device__ __forceinline__ void load_hyp(Hyp& H, const MllArgs& a, int b, Hook hook = Hook()) {
  // Warp 0 alone, ONE L2 round trip (this runs on the critical path of every CTA of passes A and C: the stamps of a -DFGP_TIMING build
  // showed 2.0 us for the prologue when one thread walked the lengthscales in a loop of dependent loads): lane j takes dimension j
  // (d <= FGP_MAX_D = 32), lane 0 also scale and noise; the prefactor scale * prod_j A_j is a butterfly product over the lanes.
  if (threadIdx.x >= 32) return;
  const int j = threadIdx.x;
  const bool on = j < a.d;
  const double l = on ? __ldcg(a.ls + (int64_t)b * a.d + j) : 0.0;
  double sc = 0.0, nz = 0.0;
  if (j == 0) {
    sc = __ldcg(a.scale + b);
    nz = __ldcg(a.noise + b);
  }
  hook();
  if (on) {
    H.ls[j] = l;
    if (a.x) {
      if (NET)
        H.xb0[j] = (uint64_t)((const int64_t*)a.x)[j];
      else
        H.x0[j] = ((const double*)a.x)[j];
    }
  }
  if (!NET && !a.x) {  // generator mode: constants of the six-slot alpha = 2 form (unused by the other variants)
    const double A = on ? fma(l, a.P.q[j][0], 1.0) : 1.0;
    if (on) {
      const double sg = sqrt(sqrt(-l * a.P.q[j][2] / A));
      H.sig[j] = sg;
      H.sig32[j] = sg * 0x1.0p-32;
    }
    double p = A;
#pragma unroll
    for (int o = 16; o; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
    if (j == 0) H.pref = sc * p;
  }
  if (j == 0) {
    H.scale = sc;
    H.noise = nz;
  }
}

// generator mode, lattice: x_i - x_0 = frac(phi2(i) z_j) exactly (the shift cancels), phi2(i) = brev64(i) / 2^64.
// For i < 2^32 brev64(i) = brev32(i) << 32, so the top word of the wrap-around 64-bit product is the wrap-around 32-bit
// product brev32(i) * (z_j mod 2^32): one IMAD, and its ceil(log2 n) significant bits sit at the top of that word.
__device__ __forceinline__ double lat_delta_gen(uint32_t rev32, uint64_t zj) {
  return (double)(rev32 * (uint32_t)zj) * 0x1.0p-32;
}

// parts of point i against the first point.  A2: every alpha_j == 2 (straight-line code, no per-dimension loop on alpha)
template <int DT, bool NET, bool A2, bool GEN>
__device__ __forceinline__ void point_parts(const MllArgs& a, const Hyp& H, int64_t i, double* p) {
  static_assert(DT > 0, "generic d goes through point_*_generic");
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  const int d = DT > 0 ? DT : a.d;
  if (GEN && !NET) {
    const uint32_t rev = __brev((uint32_t)i);
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
  if (GEN && NET) {
    const int e = (int)(i - H.tile_base);
    const uint64_t* ta = H.TA + (e & 63);
    const uint64_t* tb = H.TB + (e >> 6);
    const int nb = H.nbB;
    const bool t52 = a.t <= 52;
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      if (j >= d) break;
      const uint64_t delta = ta[j * 64] ^ tb[j * nb];
      // t <= 52 (this package's nets): the integer converts exactly through the 2^52 magic number and the exponent field gives floor(log2):
      // 18 instead of 55 instructions per (point, dimension), bit-identical to the general form (fgp_common.cuh)
      p[j] = A2 ? (t52 ? dnb2_part_a2_t52(delta, a.t, a.tscale) : dnb2_part_a2(delta, a.t, a.tscale)) : dnb2_part(delta, a.alpha.v[j], a.t);
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

// generic-d variants (DT == 0): rolled loops, out of line -- they exist for coverage of every (d, alpha), not for speed
template <bool NET, bool GEN>
__device__ __noinline__ double point_part_generic(const MllArgs& a, const Hyp& H, int64_t i, int j) {
  if (GEN && !NET) return lat_part(lat_delta_gen(__brev((uint32_t)i), a.z.v[j]), a.P.q[j], a.P.alpha[j]);
  if (GEN && NET) {
    const int e = (int)(i - H.tile_base);
    return dnb2_part(H.TA[j * 64 + (e & 63)] ^ H.TB[j * H.nbB + (e >> 6)], a.alpha.v[j], a.t);
  }
  if (NET) return dnb2_part((uint64_t)__ldg((const int64_t*)a.x + i * a.d + j) ^ H.xb0[j], a.alpha.v[j], a.t);
  return lat_part(__ldg((const double*)a.x + i * a.d + j) - H.x0[j], a.P.q[j], a.P.alpha[j]);
}
template <bool NET, bool GEN>
__device__ __noinline__ double point_k1_generic(const MllArgs& a, const Hyp& H, int64_t i) {
  double k = H.scale;
#pragma unroll 1
  for (int j = 0; j < a.d; ++j) k *= fma(H.ls[j], point_part_generic<NET, GEN>(a, H, i, j), 1.0);
  return k;
}
template <bool NET, bool GEN>
__device__ __noinline__ void point_grad_generic(const MllArgs& a, const Hyp& H, int64_t i, double w, double* acc) {
  double p[FGP_MAX_D], left[FGP_MAX_D];
  double pre = H.scale;
#pragma unroll 1
  for (int j = 0; j < a.d; ++j) {
    p[j] = point_part_generic<NET, GEN>(a, H, i, j);
    left[j] = pre;
    pre *= fma(H.ls[j], p[j], 1.0);
  }
  acc[0] = fma(w, pre, acc[0]);
  double suf = w;
#pragma unroll 1
  for (int j = a.d - 1; j >= 0; --j) {
    acc[1 + j] = fma(left[j] * suf, p[j], acc[1 + j]);
    suf *= fma(H.ls[j], p[j], 1.0);
  }
}

template <int DT, bool NET, bool A2, bool GEN>
__device__ __forceinline__ double point_k1(const MllArgs& a, const Hyp& H, int64_t i) {
  if constexpr (DT == 0) {
    return point_k1_generic<NET, GEN>(a, H, i);
  } else if constexpr (FGP_K1_SIGMA && GEN && !NET && A2) {
    const uint32_t rev = __brev((uint32_t)i);
    double k = H.pref;
#pragma unroll
    for (int j = 0; j < DT; ++j) {
      const double ts = (double)(rev * (uint32_t)a.z.v[j]) * H.sig32[j];
      const double h = ts * (H.sig[j] - ts);
      k *= fma(-h, h, 1.0);
    }
    return k;
  } else {
    double p[DT];
    point_parts<DT, NET, A2, GEN>(a, H, i, p);
    double k = H.scale;
#pragma unroll
    for (int j = 0; j < DT; ++j) k *= fma(H.ls[j], p[j], 1.0);
    return k;
  }
}

// acc[0] += w*k1 ; acc[1+j] += w * dk1/dls_j   (leave-one-out products: factors may cross zero)
template <int DT, bool NET, bool A2, bool GEN>
__device__ __forceinline__ void point_grad(const MllArgs& a, const Hyp& H, int64_t i, double w, double* acc) {
  if constexpr (DT == 0) {
    point_grad_generic<NET, GEN>(a, H, i, w, acc);
  } else {
    double p[DT], left[DT];
    point_parts<DT, NET, A2, GEN>(a, H, i, p);
    double pre = H.scale;
#pragma unroll
    for (int j = 0; j < DT; ++j) {
      left[j] = pre;
      pre *= fma(H.ls[j], p[j], 1.0);
    }
    acc[0] = fma(w, pre, acc[0]);
    double suf = w;
#pragma unroll
    for (int j = DT - 1; j >= 0; --j) {
      acc[1 + j] = fma(left[j] * suf, p[j], acc[1 + j]);
      suf *= fma(H.ls[j], p[j], 1.0);
    }
  }
}

// spectral epilogue: lam -> (norm, logdet, dnoise) partial sums and G = dL/dlam
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
// The same without the logarithm: |lam|^2 goes into a running product kept as (mantissa product, exponent sum), so that a thread
// takes ONE log for all its eigenvalues (a double-precision log is ~40 FP64 instructions, a third of the epilogue's work); exact
// range handling: m2 = f 2^e with f in [0.5, 1), log prod m2 = log prod f + ln2 sum e.
__device__ __forceinline__ double2 spectral_c_prod(double2 lam, double ysq, double wn, double wl, double& s0, double& s2, double& mant, int& ex) {
  const double a = lam.x, b = lam.y;
  const double m2 = fma(a, a, b * b);
  const double inv = 1.0 / m2;
  s0 = fma(ysq, a * inv, s0);
  const long long bits = __double_as_longlong(m2);
  ex += (int)((bits >> 52) & 0x7ff) - 1022;
  mant *= __longlong_as_double((bits & 0x800fffffffffffffLL) | 0x3fe0000000000000LL);
  const double yi2 = wn * ysq * inv * inv;
  const double li = wl * inv;
  const double ga = fma(yi2, fma(b, b, -a * a), a * li);
  const double gb = fma(-2.0 * a * b, yi2, b * li);
  s2 += ga;
  return make_double2(ga, gb);
}
__device__ __forceinline__ double logprod_flush(double& mant, int& ex) {
  const double v = fma((double)ex, 0.69314718055994530942, log(mant));
  mant = 1.0;
  ex = 0;
  return v;
}
__device__ __forceinline__ double spectral_r(double lam, double ysq, double wn, double wl, double* s) {
  const double inv = 1.0 / lam;
  s[0] = fma(ysq, inv, s[0]);
  s[1] += log(fabs(lam));
  const double g = inv * (wl - wn * ysq * inv);
  s[2] += g;
  return g;
}

// The same for a pass whose shared-memory tile is dead by now: all nv values at once through `scratch` (>= nv * warps doubles of the tile) -- two
// barriers instead of two per group of four (pass C reduces d + 1 = 9 values per CTA: 1.6 us of its critical path in the phase stamps).
// Fixed order: lanes by the shuffle tree, warps ascending.
template <int NV>
__device__ __forceinline__ void reduce_store_all(double* v, int nv, double* scratch, double* dst) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  __syncthreads();  // every reader of the tile is done
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if (k >= nv) break;
    const double r = warp_sum(v[k]);
    if (lane == 0) scratch[k * nwarp + warp] = r;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nv; k += blockDim.x) {
    double r = 0.0;
    for (int w = 0; w < nwarp; ++w) r += scratch[k * nwarp + w];
    dst[k] = r;
    __threadfence();  // the fit tail's completion ticket is taken by thread 0 after a barrier: each writer releases its own store
  }
}

// the real epilogue with the running-product logarithm of spectral_c_prod (zero / denormal eigenvalues take the plain log)
__device__ __forceinline__ double spectral_r_prod(double lam, double ysq, double wn, double wl, double* s, double& mant, int& ex) {
  const double inv = 1.0 / lam;
  s[0] = fma(ysq, inv, s[0]);
  const long long bits = __double_as_longlong(fabs(lam));
  const int e = (int)((bits >> 52) & 0x7ff);
  if (e == 0 || e == 0x7ff) {
    s[1] += log(fabs(lam));
  } else {
    ex += e - 1022;
    mant *= __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3fe0000000000000LL);
  }
  const double g = inv * (wl - wn * ysq * inv);
  s[2] += g;
  return g;
}

// block-reduce nv <= NV per-thread values and let thread 0 store them
template <int NV>
__device__ __forceinline__ void reduce_store(double* v, int nv, double* red, double* dst) {
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

// inside a fused fit loop, iterations enqueued after the stop decision cost only their launches
__device__ __forceinline__ bool fit_stopped(const MllArgs& a) { return a.has_fit && __ldcg(a.fit.state + ST_STOPPED) != 0.0; }
#ifdef FGP_TIMING
#define FGP_TSTAMP(slot)                                                                          \
  do {                                                                                            \
    if (threadIdx.x == 0 && blockIdx.x < 1024 && a.stamps) {                                      \
      long long _t;                                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                      \
      a.stamps[blockIdx.x * 48 + (slot)] = _t;                                                    \
    }                                                                                             \
  } while (0)
#else
#define FGP_TSTAMP(slot) do { } while (0)
#endif
// the same flag as a value: issue the load at kernel entry, test it after the hyperparameter loads are in flight (one L2 round
// trip per kernel instead of two on the critical path of a ~17 us kernel)
__device__ __forceinline__ double fit_stop_flag(const MllArgs& a) { return a.has_fit ? __ldcg(a.fit.state + ST_STOPPED) : 0.0; }

// deterministic reduction of the per-CTA partial sums of set b into out[b] (fixed order); any CTA size that is a
// multiple of 32.  Partials come from other CTAs: cache-global loads, all of them issued before the first reduction so that
// the call costs ONE L2 round trip.  s_out (optional, d+4 shared doubles): the same values for a fit step run by this CTA.
__device__ __forceinline__ void finalize_set(const MllArgs& a, int b, double* red, double* s_out = nullptr) {
  const int d = a.d;
  double* out = a.out + (int64_t)b * (d + 4);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  double s[3] = {0.0, 0.0, 0.0};
  // Loads first, sums after: in a rolled `s += load` loop every load waits for the previous one's use (in-order issue) -- nine dependent
  // L2 round trips per lane in the gradient loop below were 2.7 us of the 8 us serial tail of an iteration (-DFGP_TIMING stamps).
  // The first two pass-B partials of this thread are only REQUESTED here; they are added after the gradient partials below have been requested
  // as well, so that both groups share one round trip
  const double* pB = a.partB + (int64_t)b * a.ctasB * 3;
  double tB[2][3];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
#pragma unroll
    for (int k = 0; k < 3; ++k) tB[i][k] = c < a.ctasB ? __ldcg(pB + c * 3 + k) : 0.0;
  }
  auto sumB = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int k = 0; k < 3; ++k) s[k] += tB[i][k];
    }
    for (int c0 = threadIdx.x + 2 * blockDim.x; c0 < a.ctasB; c0 += 2 * blockDim.x) {
      double t[2][3];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = c0 + i * blockDim.x;
#pragma unroll
        for (int k = 0; k < 3; ++k) t[i][k] = c < a.ctasB ? __ldcg(pB + c * 3 + k) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int k = 0; k < 3; ++k) s[k] += t[i][k];
      }
    }
  };
  // warp w reduces gradient components w, w + nwarp, ... over all pass-C CTAs, four components' loads in flight at a time; the
  // first four are requested before the block reduction above is waited for
  const double* p = a.partC + (int64_t)b * a.ctasA * (d + 1);
  auto load4 = [&](int j0, double* v) {
    // components j0 + k nwarp, k = 0..3, two at a time: the loads of BOTH are requested before either sum (warp 0 owns components 0 and 8 at
    // d = 8: one L2 round trip instead of two); each sum runs over the CTAs in ascending order, so the result does not depend on the split
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
      const int ja = j0 + k * nwarp, jb = ja + nwarp;
      v[k] = 0.0;
      v[k + 1] = 0.0;
      if (ja > d) continue;  // uniform over the warp
      for (int c0 = lane; c0 < a.ctasA; c0 += 32 * 10) {  // 10 (+10) loads in flight per lane
        double ta[10], tb[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          const int c = c0 + 32 * i;
          ta[i] = c < a.ctasA ? __ldcg(p + (int64_t)c * (d + 1) + ja) : 0.0;
          tb[i] = (c < a.ctasA && jb <= d) ? __ldcg(p + (int64_t)c * (d + 1) + jb) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < 10; ++i) {
          v[k] += ta[i];
          v[k + 1] += tb[i];
        }
      }
    }
  };
  double inv_scale = 1.0;
  auto store4 = [&](int j0, const double* v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 + k * nwarp;
      if (j > d) continue;  // uniform over the warp
      double r = warp_sum(v[k]);
      if (lane == 0) {
        r = j == 0 ? r * inv_scale : r;
        out[3 + j] = r;
        if (s_out) s_out[3 + j] = r;
      }
    }
  };
  double v0[4];
  if (a.want_grad) {
    inv_scale = 1.0 / __ldcg(a.scale + b);
    load4(warp, v0);
  }
  sumB();
  block_sum<3>(s, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      out[k] = s[k];
      if (s_out) s_out[k] = s[k];
    }
  }
  if (!a.want_grad) return;
  store4(warp, v0);
  for (int j0 = warp + 4 * nwarp; j0 <= d; j0 += 4 * nwarp) {
    load4(j0, v0);
    store4(j0, v0);
  }
}

// Tail of a fused fit iteration, called by every CTA of the iteration's LAST kernel after its partial sums are stored:
// the last CTA of set b finalizes b; the last finalizer runs the fit step (loss, early stop, Rprop, new hyperparameters).
// ctas_b: CTAs per set in this kernel; B: sets; two_pass: partial sums need reducing.
// Critical path (it is serial: ~8 us of a 51 us iteration in round 1): one atomic ticket, one L2 round trip for the partial
// sums, the fit step on values that are already in shared memory.  The state header is requested BEFORE the ticket (it was
// written by the previous iteration's fit step, long ago), and with one set the second ticket is skipped.
__device__ __forceinline__ void mll_fit_tail(const MllArgs& a, int b, int ctas_b, int B, bool two_pass, double* red) {
  __shared__ int s_last;
  __shared__ double s_hdr[ST_HEADER];
  __shared__ int s_flags[2];
  __shared__ double s_out[FGP_MAX_D + 4];
  const double pre_hdr = threadIdx.x < ST_HEADER ? __ldcg(a.fit.state + threadIdx.x) : 0.0;
  // one set: the fit step's own inputs too (nobody writes them during this kernel) -- see FitPrefetch
  const bool use_pf = two_pass && B == 1 && a.fit.P <= (int)blockDim.x;
  FitPrefetch pf{0.0, 0.0, 0.0};
  if (use_pf) pf = fit_prefetch(a.fit);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // release: this CTA's partial sums before its ticket
    const bool last = atomicAdd(&a.fit.tickets[1 + b], 1u) == (unsigned)(ctas_b - 1);
    if (last) __threadfence();  // acquire, by the observing thread alone: the barrier below orders the other threads' (cache-global) loads after it
    s_last = last;
  }
  __syncthreads();
  if (!s_last) return;
  FGP_TSTAMP(32);
  const bool local = two_pass && B == 1;  // this CTA finalizes the only set: the fit step reads the sums from shared memory
  if (two_pass) finalize_set(a, b, red, local ? s_out : nullptr);
  FGP_TSTAMP(33);
  if (threadIdx.x < ST_HEADER) s_hdr[threadIdx.x] = pre_hdr;
  if (B == 1) {
    if (threadIdx.x == 0) a.fit.tickets[1 + b] = 0u;
    __syncthreads();
    fit_step_device(a.fit, a.out, red, s_hdr, s_flags, true, local ? s_out : nullptr, use_pf ? &pf : nullptr);
    FGP_TSTAMP(34);
    return;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    a.fit.tickets[1 + b] = 0u;
    __threadfence();
    const bool last = atomicAdd(&a.fit.tickets[0], 1u) == (unsigned)(B - 1);
    if (last) a.fit.tickets[0] = 0u;
    s_last = last;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  fit_step_device(a.fit, a.out, red, s_hdr, s_flags, true, nullptr);
}

// ------------------------------------------------------------------------------------------------------------
// single-pass kernel: one CTA per hyperparameter set, n <= block capacity
// Heavy per-element work (kernel evaluation, log/divide epilogue, gradient contraction) runs in rolled element loops
// over shared memory; only the transform rounds are unrolled (instruction-cache footprint, profiles/README.md).
// ------------------------------------------------------------------------------------------------------------
template <int DT, bool NET, bool A2, bool GEN>
__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) mll_single_kernel(const __grid_constant__ MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ Hyp H;
  __shared__ double red[kRed];
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  pdl_prologue(a);
  const double stop_flag = fit_stop_flag(a);
  const int b = blockIdx.x;
  const int n = (int)a.n;
  const int l = a.l1;
  const int LP = a.LPA;
  const int d = DT > 0 ? DT : a.d;
  load_hyp<NET>(H, a, b);
  if (stop_flag != 0.0) return;  // uniform over the CTA, before any barrier
  if (GEN && NET) dnb2_build_tables(H, a, (uint64_t*)(smraw + a.tab_off), 0, l);
  __syncthreads();
  const double c = H.scale;  // DC guess removed before the transform (role of abstract_fast_gp.py:209-211)
  const double noise = H.noise;
  const double* ysq = a.ysq + (int64_t)b * n;
  const double wn = a.weights ? a.weights[2 * b] : 0.5, wl = a.weights ? a.weights[2 * b + 1] : 0.5;
  double s[3] = {0.0, 0.0, 0.0};
  double acc[DM + 1];
#pragma unroll
  for (int j = 0; j <= DM; ++j) acc[j] = 0.0;
  const int want_grad = a.want_grad;
  if (NET) {
    double* sm = (double*)smraw;
    const SmemR S{sm, LP};
    double* lamo = a.lam ? a.lam + (int64_t)b * n : nullptr;
    tile_fill_r<false>(S, l, 0, [&](int, int idx) -> double { return point_k1<DT, NET, A2, GEN>(a, H, idx) - c; });
    __syncthreads();
    block_wht_io<false>(sm, l, 0, LP, wht_sched_up(l), SmemTag{}, SmemTag{});
    __syncthreads();
    tile_map_r<false>(S, l, 0, [&](int, int k, double v) -> double {
      double lam = v + noise;
      if (k == 0) lam += c * (double)n;
      if (lamo) lamo[k] = lam;
      return spectral_r(lam, ysq[k], wn, wl, s);
    });
    if (want_grad) {
      __syncthreads();
      block_wht_io<false>(sm, l, 0, LP, wht_sched_up(l), SmemTag{}, SmemTag{});
      __syncthreads();
      tile_drain_r<false>(S, l, 0, [&](int, int idx, double w) { point_grad<DT, NET, A2, GEN>(a, H, idx, w, acc); });
    }
  } else {
    double2* sm = (double2*)smraw;
    const SmemC S{sm, LP};
    double2* lamo = a.lam ? (double2*)a.lam + (int64_t)b * n : nullptr;
    tile_fill_c<false>(S, l, 0, [&](int, int idx) -> double2 { return make_double2(point_k1<DT, NET, A2, GEN>(a, H, idx) - c, 0.0); });
    __syncthreads();
    block_fft_fwd_io<false>(sm, l, 0, LP, a.T.stage, SmemTag{}, SmemTag{});
    __syncthreads();
    tile_map_c<false>(S, l, 0, [&](int, int k, double2 lam) -> double2 {
      lam.x += noise;
      if (k == 0) lam.x += c * (double)n;
      if (lamo) lamo[k] = lam;
      return spectral_c(lam, ysq[k], wn, wl, s);
    });
    if (want_grad) {
      __syncthreads();
      block_fft_inv_io<false>(sm, l, 0, LP, a.T.stage, SmemTag{}, SmemTag{});
      __syncthreads();
      tile_drain_c<false>(S, l, 0, [&](int, int idx, double2 w) { point_grad<DT, NET, A2, GEN>(a, H, idx, w.x, acc); });
    }
  }
  double* out = a.out + (int64_t)b * (d + 4);
  reduce_store<3>(s, 3, red, out);
  if (want_grad) {
    acc[0] /= H.scale;
    reduce_store<DM + 1>(acc, d + 1, red, out + 3);
  }
  if (a.has_fit) mll_fit_tail(a, b, 1, gridDim.x, false, red);
}

// ------------------------------------------------------------------------------------------------------------
// two-pass tile bodies.  One "tile" is what one CTA of the per-pass kernels does; the persistent cooperative kernel
// below walks the same tiles in the same order of operations, so both routes give bit-identical results.
// Workspace reads are cache-global (__ldcg): in the persistent kernel another CTA rewrote W since this SM last read it.
// ------------------------------------------------------------------------------------------------------------
#ifdef FGP_TIMING
// tools-only build (-DFGP_TIMING, FGP_LIB_DIR): clock stamps of the persistent kernel, read back by fgp_debug_stamps()
constexpr int kStampSlots = 48, kStampCtas = 1024;  // 0-7 globaltimer, 8-15 clock64 (persistent kernel); 16-31 globaltimer of the per-pass kernels
long long* debug_stamp_buffer();  // fgp_mll_passb.cu: device buffer, allocated on first use
#define FGP_STAMP(slot)                                                                           \
  do {                                                                                            \
    if (threadIdx.x == 0 && blockIdx.x < kStampCtas && a.stamps) {                                \
      long long _t;                                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                      \
      a.stamps[blockIdx.x * kStampSlots + (slot)] = _t;                                           \
      a.stamps[blockIdx.x * kStampSlots + 8 + (slot)] = clock64();                                \
    }                                                                                             \
  } while (0)
#define FGP_PSTAMP(slot)                                                                          \
  do {                                                                                            \
    if (threadIdx.x == 0 && blockIdx.x < kStampCtas && a.stamps) {                                \
      long long _t;                                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                      \
      a.stamps[blockIdx.x * kStampSlots + (slot)] = _t;                                           \
    }                                                                                             \
  } while (0)
#else
#define FGP_STAMP(slot) do { } while (0)
#define FGP_PSTAMP(slot) do { } while (0)
#endif

// hyperparameters (and the net generator tables) of set b for a tile starting at point g0
template <bool NET, bool GEN, typename Hook = NoHook>
__device__ __forceinline__ void tile_prologue(const MllArgs& a, Hyp& H, unsigned char* smraw, int b, int64_t g0, int tile_log, Hook hook = Hook()) {
  if (threadIdx.x >= 32) hook();
  __syncthreads();  // the previous tile's readers of H and of the shared-memory tile are done
  load_hyp<NET>(H, a, b, hook);
  if (GEN && NET) dnb2_build_tables(H, a, (uint64_t*)(smraw + a.tab_off), g0, tile_log);
  __syncthreads();
}
template <bool NET>
__device__ __forceinline__ int64_t tileA_block0(const MllArgs& a, int tile) {
  // half-spectrum mode (lntr == 0): tile c takes residue class r = c, i.e. block row rev(c)
  return (!NET && a.hs) ? (int64_t)brev_bits((uint32_t)tile, a.l2) : ((int64_t)tile << a.lntrA);
}

// pass A: k1 on the fly -> shared memory -> contiguous block transform -> inter-pass twiddle -> workspace
template <int DT, bool NET, bool A2, bool GEN>
__device__ __forceinline__ void passA_tile(const MllArgs& a, const Hyp& H, unsigned char* smraw, int tile, int b) {
  const int l1 = a.l1, l2 = a.l2, lntr = a.lntrA, LP = a.LPA;
  const int64_t blk0 = tileA_block0<NET>(a, tile);
  const int64_t g0 = blk0 << l1;
  const double c = H.scale;  // DC guess removed before the transform (role of abstract_fast_gp.py:209-211)
  if (NET) {
    double* sm = (double*)smraw;
    double* W = (double*)a.W + (int64_t)b * a.n + g0;
    tile_fill_r<false>(SmemR{sm, LP}, l1, lntr,
                       [&](int tr, int idx) -> double { return point_k1<DT, NET, A2, GEN>(a, H, g0 + ((int64_t)tr << l1) + idx) - c; });
    __syncthreads();
    FGP_PSTAMP(18);
    FGP_PSTAMP(19);
    // any schedule whose last round is strided (coalesced stores)
    block_wht_io<false>(sm, l1, lntr, LP, wht_sched_up(l1), SmemTag{}, [&](int tr, int idx, double v) { W[((int64_t)tr << l1) + idx] = v; });
  } else {
    double2* sm = (double2*)smraw;
    double2* W = (double2*)a.W + (int64_t)b * a.n + g0;
    const FftTables T = a.T;
    // Half-spectrum mode: the block is REAL (one block per tile): transform it as ONE complex transform of HALF its length.  In the
    // bit-reversed input order the even natural samples are the first half of the block and the odd ones the second half, so
    // z[p] = x[p] + i x[p + L1/2]; with Z = ft(z):  E_k = (Z_k + conj Z_{h-k}) / 2,  O_k = (Z_k - conj Z_{h-k}) / (2i)  are the transforms of
    // the two halves and  X_k = E_k + w^k O_k  (k = 0 .. h, w = exp(-2 pi i / L1)) is the last radix-2 stage of the full transform.  Pass B
    // reads the columns q <= L1/2 only.  Half the butterflies and half the shared-memory traffic of the complex transform of (x, 0).
    const bool rf = a.hs != 0;
    const int lh = l1 - 1, h = 1 << lh;
    {
      // ONE rolled loop (one inlined copy of the kernel evaluation) fills either layout: (x, 0) per element, or component e >> lh of slot e & (h-1)
      const int total = 1 << (lntr + l1);
#pragma unroll kFillUnroll
      for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const double v = point_k1<DT, NET, A2, GEN>(a, H, g0 + e) - c;
        if (rf)
          ((double*)&sm[padidx<kPSC>(e & (h - 1))])[e >> lh] = v;
        else
          sm[(e >> l1) * LP + padidx<kPSC>(e & ((1 << l1) - 1))] = make_double2(v, 0.0);
      }
    }
    __syncthreads();
    FGP_PSTAMP(18);
    if (rf) {
      const SmemC S{sm, LP};
      block_fft_fwd_io<false>(sm, lh, 0, LP, T.stage, SmemTag{}, SmemTag{});
      __syncthreads();
      FGP_PSTAMP(19);
      const uint32_t res = brev_bits((uint32_t)(blk0 & ((1 << l2) - 1)), l2);  // residue class of this block row
      for (int k = threadIdx.x; k <= h; k += blockDim.x) {
        const double2 zk = S(0, k & (h - 1)), zm = S(0, (h - k) & (h - 1));
        const double2 E = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
        const double2 O = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
        const double2 wk = k < h ? __ldg(T.stage + h + k) : make_double2(-1.0, 0.0);  // exp(-2 pi i k / L1): one table entry
        const double2 X = cadd(E, cmul(wk, O));
        W[k] = cmul(X, twiddle_n(T, res * (uint32_t)k));
      }
      // the last pass-B tile also loads the (weight-0) columns just above L1/2: keep them finite
      for (int k = h + 1 + threadIdx.x; k < h + (1 << a.lntrB) && k < (1 << l1); k += blockDim.x) W[k] = make_double2(0.0, 0.0);
      return;
    }
    block_fft_fwd_io<false>(sm, l1, lntr, LP, T.stage, SmemTag{}, [&](int tr, int idx, double2 v) {
      const uint32_t bb = (uint32_t)((blk0 + tr) & ((1 << l2) - 1));
      W[((int64_t)tr << l1) + idx] = cmul(v, twiddle_n(T, brev_bits(bb, l2) * (uint32_t)idx));
    });
  }
}

// pass B: column transform -> lam -> spectral epilogue -> inverse column transform; per-tile partial sums -> partB
template <bool NET>
__device__ __forceinline__ void passB_tile(const MllArgs& a, unsigned char* smraw, double* red, int tile, int b) {
  const double noise = __ldcg(a.noise + b);
  const double dc = __ldcg(a.scale + b) * (double)a.n;  // the DC guess removed in pass A comes back in bin 0
  const int l1 = a.l1, l2 = a.l2, lntr = a.lntrB, LP = a.LPB;
  const int q0 = tile << lntr;
  const int64_t boff = (int64_t)b * a.n;
  const double* ysq = a.ysq + boff + q0;
  const double wn = a.weights ? a.weights[2 * b] : 0.5, wl = a.weights ? a.weights[2 * b + 1] : 0.5;
  const int want_grad = a.want_grad;
  double s[3] = {0.0, 0.0, 0.0};
  // |y~|^2 of this tile (one 64- or 128-byte segment per row) is only read by the spectral epilogue, after the column transform: when it is
  // not in L2 that phase waits for HBM (2.5 -> 5.5 us in the stamps of an L2-flushed iteration).  Request the lines now.
  for (int r = threadIdx.x; r < (1 << l2); r += blockDim.x) l2_prefetch_line(ysq + ((int64_t)r << l1));
  if (NET) {
    double* sm = (double*)smraw;
    double* W = (double*)a.W + boff + q0;
    double* lamo = a.lam ? a.lam + boff + q0 : nullptr;
    block_wht_io<true>(sm, l2, lntr, LP, wht_sched_up(l2), [&](int tr, int r) -> double { return __ldcg(W + ((int64_t)r << l1) + tr); }, SmemTag{});
    __syncthreads();
    FGP_PSTAMP(22);
    double mant = 1.0;  // one logarithm per 64 eigenvalues of a thread instead of one each (a third of this epilogue's FP64 work)
    int ex = 0, cnt = 0;
    tile_map_r<true>(SmemR{sm, LP}, l2, lntr, [&](int tr, int r, double v) -> double {
      const int64_t k = ((int64_t)r << l1) + tr;
      double lam = v + noise;
      if (k + q0 == 0) lam += dc;
      if (lamo) lamo[k] = lam;
      const double g = spectral_r_prod(lam, ysq[k], wn, wl, s, mant, ex);
      if (++cnt == 64) {  // 0.5^64 is far from underflow
        s[1] += logprod_flush(mant, ex);
        cnt = 0;
      }
      return g;
    });
    s[1] += logprod_flush(mant, ex);
    if (want_grad) {
      __syncthreads();
      FGP_PSTAMP(23);
      block_wht_io<true>(sm, l2, lntr, LP, wht_sched_up(l2), SmemTag{}, [&](int tr, int r, double v) { W[((int64_t)r << l1) + tr] = v; });
    }
  } else {
    double2* sm = (double2*)smraw;
    double2* W = (double2*)a.W + boff + q0;
    double2* lamo = a.lam ? (double2*)a.lam + boff + q0 : nullptr;
    const FftTables T = a.T;
    if (a.hs) {
      // half-spectrum mode (MllArgs::hs): block rows of residue class r > L2/2 were not written by pass A, they are the
      // conjugates of class L2 - r; columns q and L1 - q carry the same eigenvalues, so column q counts twice
      const uint32_t L2 = 1u << l2, half2 = L2 >> 1;
      const int half1 = 1 << (l1 - 1);
      block_fft_fwd_io<true>(sm, l2, lntr, LP, T.stage, [&](int tr, int r) -> double2 {
        const uint32_t res = brev_bits((uint32_t)r, l2);
        const bool mir = res > half2;
        const uint32_t row = mir ? brev_bits(L2 - res, l2) : (uint32_t)r;
        const double2 v = __ldcg(W + ((int64_t)row << l1) + tr);
        return make_double2(v.x, mir ? -v.y : v.y);
      }, SmemTag{});
      __syncthreads();
      FGP_PSTAMP(22);
      // a thread stays in one column (the block size is a multiple of the columns per tile), so its weight cw is a constant and the
      // log-determinant terms of its eigenvalues can be taken as one log of their product
      double mant = 1.0, t0 = 0.0, t2 = 0.0;
      int ex = 0, cnt = 0;
      const int qt = q0 + (threadIdx.x & ((1 << lntr) - 1));
      const double cwt = (qt == 0 || qt == half1) ? 1.0 : (qt < half1 ? 2.0 : 0.0);
      tile_map_c<true>(SmemC{sm, LP}, l2, lntr, [&](int tr, int r, double2 lam) -> double2 {
        const int64_t k = ((int64_t)r << l1) + tr;
        lam.x += noise;
        if (k + q0 == 0) lam.x += dc;
        const double2 G = spectral_c_prod(lam, ysq[k], wn, wl, t0, t2, mant, ex);
        if (++cnt == 64) {  // 0.5^64 is far from underflow
          s[1] = fma(0.5 * cwt, logprod_flush(mant, ex), s[1]);
          cnt = 0;
        }
        // lam is real in exact arithmetic; its computed imaginary part is round-off, but dL/dIm(lam) ~ Im(lam) |y~|^2 / lam^3 is
        // not small where lam is.  A Hermitian (instead of real) dL/dlam back-transforms to a real but not EVEN sequence,
        // and the odd part only cancels in a sum over all points -- pass C sums half of them twice.  Keep the real part.
        return make_double2(G.x, 0.0);
      });
      s[0] = cwt * t0;
      s[1] = fma(0.5 * cwt, logprod_flush(mant, ex), s[1]);
      s[2] = cwt * t2;
      if (want_grad && (q0 == 0 || q0 == half1)) {
        // the two self-mirrored columns hold both members of every pair (k, n-k): make them exactly equal as well
        __syncthreads();
        double2* col = sm;  // tr == 0
        const int L2i = 1 << l2;
        for (int sidx = threadIdx.x; sidx < (L2i >> 1); sidx += blockDim.x) {
          const int s1 = sidx;
          const int s2 = q0 == 0 ? (L2i - sidx) & (L2i - 1) : L2i - 1 - sidx;
          if (s1 != s2) {
            const int i1 = padidx<kPSC>(s1), i2 = padidx<kPSC>(s2);
            const double av = 0.5 * (col[i1].x + col[i2].x);
            col[i1].x = av;
            col[i2].x = av;
          }
        }
      }
      if (want_grad) {
        __syncthreads();
        FGP_PSTAMP(23);
        block_fft_inv_io<true>(sm, l2, lntr, LP, T.stage, SmemTag{}, [&](int tr, int r, double2 v) {
          const uint32_t res = brev_bits((uint32_t)r, l2);
          if (res > half2) return;  // pass C never reads the mirrored block rows
          const double2 w = twiddle_n(T, res * (uint32_t)(q0 + tr));
          W[((int64_t)r << l1) + tr] = cmulc(w, v);
        });
      }
    } else {
      block_fft_fwd_io<true>(sm, l2, lntr, LP, T.stage, [&](int tr, int r) -> double2 { return __ldcg(W + ((int64_t)r << l1) + tr); }, SmemTag{});
      __syncthreads();
      tile_map_c<true>(SmemC{sm, LP}, l2, lntr, [&](int tr, int r, double2 lam) -> double2 {
        const int64_t k = ((int64_t)r << l1) + tr;
        lam.x += noise;
        if (k + q0 == 0) lam.x += dc;
        if (lamo) lamo[k] = lam;
        return spectral_c(lam, ysq[k], wn, wl, s);
      });
      if (want_grad) {
        __syncthreads();
        block_fft_inv_io<true>(sm, l2, lntr, LP, T.stage, SmemTag{}, [&](int tr, int r, double2 v) {
          const double2 w = twiddle_n(T, brev_bits((uint32_t)r, l2) * (uint32_t)(q0 + tr));
          W[((int64_t)r << l1) + tr] = cmulc(w, v);
        });
      }
    }
  }
  __syncthreads();
  reduce_store<3>(s, 3, red, a.partB + ((int64_t)b * a.ctasB + tile) * 3);
}

// pass C: contiguous blocks of the back-transformed dL/dlam -> inverse block transform -> contraction with dk1/dtheta
// Half-spectrum pass C, 2^11-point tiles, 256 threads: the tile's loads do not depend on the hyperparameters, so the per-pass kernel issues
// them BEFORE the tile prologue (one L2 round trip in flight behind the other instead of two in a row: the stamps showed 1.6 + 1.4 us).
constexpr int kPreC = 4;
struct PreC {
  double2 xk[kPreC], xm[kPreC];
  bool on = false;
};
template <bool NET>
__device__ __forceinline__ void passC_preload(const MllArgs& a, int tile, int b, PreC& P) {
  if (NET || !a.hs || (1 << (a.l1 - 1)) != kPreC * (int)blockDim.x) return;
  const int h = 1 << (a.l1 - 1);
  const double2* W = (const double2*)a.W + (int64_t)b * a.n + (tileA_block0<NET>(a, tile) << a.l1);
#pragma unroll
  for (int i = 0; i < kPreC; ++i) {
    const int k = threadIdx.x + i * (int)blockDim.x;
    P.xk[i] = __ldcg(W + k);
    P.xm[i] = __ldcg(W + (h - k));
  }
  P.on = true;
}
template <int DT, bool NET, bool A2, bool GEN>
__device__ __forceinline__ void passC_tile(const MllArgs& a, const Hyp& H, unsigned char* smraw, double* red, int tile, int b, const PreC& P = PreC()) {
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  const int d = DT > 0 ? DT : a.d;
  const int l1 = a.l1, lntr = a.lntrA, LP = a.LPA;
  const bool hs = !NET && a.hs;
  const int64_t g0 = tileA_block0<NET>(a, tile) << l1;
  double acc[DM + 1];
#pragma unroll
  for (int j = 0; j <= DM; ++j) acc[j] = 0.0;
  if (NET) {
    double* sm = (double*)smraw;
    const double* W = (const double*)a.W + (int64_t)b * a.n + g0;
    // top stages first: a thread's 16 loads are 2^(l1-4) apart, consecutive threads read consecutive addresses
    block_wht_io<false>(sm, l1, lntr, LP, wht_sched_coalesced(l1), [&](int tr, int idx) -> double { return __ldcg(W + ((int64_t)tr << l1) + idx); }, SmemTag{});
    __syncthreads();
    FGP_PSTAMP(27);
    FGP_PSTAMP(28);
    tile_drain_r<false>(SmemR{sm, LP}, l1, lntr,
                        [&](int tr, int idx, double w) { point_grad<DT, NET, A2, GEN>(a, H, g0 + ((int64_t)tr << l1) + idx, w, acc); });
  } else {
    double2* sm = (double2*)smraw;
    const double2* W = (const double2*)a.W + (int64_t)b * a.n + g0;
    // Half-spectrum mode: the back-transformed row is Hermitian in q (columns above L1/2 were never computed), its inverse transform REAL:
    // the inverse of pass A's split.  E'_k = X_k + conj X_{h-k},  O'_k = conj(w^k) (X_k - conj X_{h-k}),  Z'_k = E'_k + i O'_k  (k < h = L1/2),
    // ONE inverse complex transform of length h, and x[p] = Re z[p], x[p + h] = Im z[p].
    const int lh = l1 - 1, h = 1 << lh;
    if (hs) {
      const SmemC S{sm, LP};
      const FftTables T = a.T;
      const int l2 = a.l2;
      if (P.on) {
#pragma unroll
        for (int i = 0; i < kPreC; ++i) {
          const int k = threadIdx.x + i * (int)blockDim.x;
          const double2 xk = P.xk[i], xm = P.xm[i];
          const double2 o = cmulc(__ldg(T.stage + h + k), make_double2(xk.x - xm.x, xk.y + xm.y));
          S(0, k, make_double2((xk.x + xm.x) - o.y, (xk.y - xm.y) + o.x));
        }
      } else {
        for (int k = threadIdx.x; k < h; k += blockDim.x) {
          const double2 xk = __ldcg(W + k), xm = __ldcg(W + (h - k));
          const double2 o = cmulc(__ldg(T.stage + h + k), make_double2(xk.x - xm.x, xk.y + xm.y));
          S(0, k, make_double2((xk.x + xm.x) - o.y, (xk.y - xm.y) + o.x));
        }
      }
      __syncthreads();
      FGP_PSTAMP(27);
      block_fft_inv_io<false>(sm, lh, 0, LP, T.stage, SmemTag{}, SmemTag{});
    } else {
      block_fft_inv_io<false>(sm, l1, lntr, LP, a.T.stage, [&](int tr, int idx) -> double2 { return __ldcg(W + ((int64_t)tr << l1) + idx); }, SmemTag{});
    }
    __syncthreads();
    FGP_PSTAMP(28);
    {
      // ONE rolled loop (one inlined copy of the gradient contraction) drains either layout
      const int total = 1 << (lntr + l1);
#pragma unroll kFillUnroll
      for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const double* slot = hs ? (const double*)&sm[padidx<kPSC>(e & (h - 1))] + (e >> lh) : (const double*)&sm[(e >> l1) * LP + padidx<kPSC>(e & ((1 << l1) - 1))];
        point_grad<DT, NET, A2, GEN>(a, H, g0 + e, *slot, acc);
      }
    }
    if (hs && tile != 0 && tile != (1 << (a.l2 - 1))) {  // class r stands for r and L2 - r
#pragma unroll
      for (int j = 0; j <= DM; ++j) acc[j] *= 2.0;
    }
  }
  FGP_PSTAMP(29);
  reduce_store_all<DM + 1>(acc, d + 1, (double*)smraw, a.partC + ((int64_t)b * a.ctasA + tile) * (d + 1));
}

// ------------------------------------------------------------------------------------------------------------
// per-pass kernels: one tile per CTA, three launches per evaluation (plain fgp_*_mll_grad calls, and the fit iteration when
// the cooperative route is switched off)
// ------------------------------------------------------------------------------------------------------------
template <int DT, bool NET, bool A2, bool GEN>
__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) mll_passA_kernel(const __grid_constant__ MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ Hyp H;
  pdl_prologue(a);
  FGP_PSTAMP(16);
  const double stop_flag = fit_stop_flag(a);  // tested after the prologue: its load shares the prologue's L2 round trip
  const int tile = blockIdx.x, b = blockIdx.y;
  if (!NET && threadIdx.x == 0) {  // twiddle tables (192 KiB, first used 4 us from now): one bulk L2 prefetch of 1 KiB per CTA
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x, total = 3u * kTabLen * (unsigned)sizeof(double2);
    if (cta * 1024u < total) l2_prefetch_bulk((const char*)a.T.stage + cta * 1024u, 1024u);
  }
  tile_prologue<NET, GEN>(a, H, smraw, b, tileA_block0<NET>(a, tile) << a.l1, a.l1 + a.lntrA);
  if (stop_flag != 0.0) return;  // uniform over the CTA
  FGP_PSTAMP(17);
  passA_tile<DT, NET, A2, GEN>(a, H, smraw, tile, b);
  FGP_PSTAMP(20);
}

template <int DT, bool NET, bool A2, bool GEN>
__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) mll_passC_kernel(const __grid_constant__ MllArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ Hyp H;
  __shared__ double red[kRed];
  pdl_prologue(a);
  FGP_PSTAMP(25);
  const double stop_flag = fit_stop_flag(a);  // tested after the prologue: its load shares the prologue's L2 round trip
  const int tile = blockIdx.x, b = blockIdx.y;
  PreC P;
  // W is valid memory whatever the stop flag says
  tile_prologue<NET, GEN>(a, H, smraw, b, tileA_block0<NET>(a, tile) << a.l1, a.l1 + a.lntrA, [&]() { passC_preload<NET>(a, tile, b, P); });
  if (stop_flag != 0.0) return;  // uniform over the CTA
  FGP_PSTAMP(26);
  passC_tile<DT, NET, A2, GEN>(a, H, smraw, red, tile, b, P);
  FGP_PSTAMP(30);
  // PDL mode 2: this CTA's tile is done; what is left is the serial tail of ONE CTA (ticket, partial sums, fit step: ~4.7 us).  Trigger now, so
  // that the next iteration's pass-A CTAs are launched and resident (waiting in cudaGridDependencySynchronize) while the tail runs.
  if (a.pdl == 2) cudaTriggerProgrammaticLaunchCompletion();
  if (a.has_fit) mll_fit_tail(a, b, a.ctasA, gridDim.y, true, red);
  FGP_PSTAMP(31);
}

// ------------------------------------------------------------------------------------------------------------
// persistent cooperative kernel: a whole fit() iteration -- or a.iters of them -- in ONE launch.  Every CTA is resident
// (cooperative launch, grid <= occupancy x SMs) and walks pass-A tiles, a grid barrier, pass-B tiles, a grid barrier,
// pass-C tiles; the CTA that finishes the last tile reduces the partial sums and runs the fit step (mll_fit_tail); before
// the next iteration a third barrier makes the new hyperparameters visible.  Replaces three ~17 us launches, a third of
// which was launch latency, ramp-up and drain (profiles/README.md, round 1 snapshot j).
// ------------------------------------------------------------------------------------------------------------
// Arrive-and-wait on a monotonically increasing counter.  The wait gives up after ~1 s of device clock and raises the
// error word instead of hanging the GPU if the launch contract (all CTAs resident) were ever broken.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target, double* fit_state) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    for (;;) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if ((int)(v - target) >= 0) break;
      if (clock64() - t0 > (1ll << 31)) {
        atomicExch(bar + 2, 1u);
        if (fit_state) fit_state[ST_STOPPED] = 2.0;  // the host reads 2 as "the device-side loop failed"
        break;
      }
    }
    __threadfence();
  }
  __syncthreads();
}

#ifndef FGP_COOP_PHASE_ATTR
#define FGP_COOP_PHASE_ATTR __forceinline__
#endif
// phase bodies of the persistent kernel (-DFGP_COOP_PHASE_ATTR=__noinline__ gives each its own register allocation; measured
// worse: the argument block is then read through a generic pointer instead of constant-bank operands)
template <int DT, bool NET, bool A2, bool GEN>
__device__ FGP_COOP_PHASE_ATTR void coop_phaseA(const MllArgs& a, Hyp& H, unsigned char* smraw, int tile, int b) {
  tile_prologue<NET, GEN>(a, H, smraw, b, tileA_block0<NET>(a, tile) << a.l1, a.l1 + a.lntrA);
  passA_tile<DT, NET, A2, GEN>(a, H, smraw, tile, b);
}
template <bool NET>
__device__ FGP_COOP_PHASE_ATTR void coop_phaseB(const MllArgs& a, unsigned char* smraw, double* red, int tile, int b, int B) {
  __syncthreads();  // the previous tile's shared-memory readers are done
  passB_tile<NET>(a, smraw, red, tile, b);
  if (a.has_fit && !a.want_grad) mll_fit_tail(a, b, a.ctasB, B, true, red);
}
template <int DT, bool NET, bool A2, bool GEN>
__device__ FGP_COOP_PHASE_ATTR void coop_phaseC(const MllArgs& a, Hyp& H, unsigned char* smraw, double* red, int tile, int b, int B) {
  tile_prologue<NET, GEN>(a, H, smraw, b, tileA_block0<NET>(a, tile) << a.l1, a.l1 + a.lntrA);
  passC_tile<DT, NET, A2, GEN>(a, H, smraw, red, tile, b);
  if (a.has_fit) mll_fit_tail(a, b, a.ctasA, B, true, red);
}

#ifndef FGP_COOP_MINB
#define FGP_COOP_MINB FGP_LB_BLOCKS
#endif
template <int DT, bool NET, bool A2, bool GEN>
__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_COOP_MINB) mll_coop_kernel(const __grid_constant__ MllArgs a, const int B) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ Hyp H;
  __shared__ double red[kRed];
  const int tilesA = a.ctasA * B, tilesB = a.ctasB * B;
  unsigned int nbar = 0;
  FGP_STAMP(0);
  for (int it = 0; it < a.iters; ++it) {
    // written by the fit step before the barrier that ended the previous iteration (or by an earlier launch): every CTA
    // reads the same value, so the whole grid leaves the loop together
    if (fit_stop_flag(a) != 0.0) break;
    for (int w = blockIdx.x; w < tilesA; w += gridDim.x) {
      const int b = w / a.ctasA, tile = w - b * a.ctasA;
      coop_phaseA<DT, NET, A2, GEN>(a, H, smraw, tile, b);
    }
    FGP_STAMP(1);
    grid_barrier(a.bar, ++nbar * gridDim.x, a.has_fit ? a.fit.state : nullptr);
    FGP_STAMP(2);
    for (int w = blockIdx.x; w < tilesB; w += gridDim.x) {
      const int b = w / a.ctasB, tile = w - b * a.ctasB;
      coop_phaseB<NET>(a, smraw, red, tile, b, B);
    }
    FGP_STAMP(3);
    if (a.want_grad) {
      grid_barrier(a.bar, ++nbar * gridDim.x, a.has_fit ? a.fit.state : nullptr);
      FGP_STAMP(4);
      for (int w = blockIdx.x; w < tilesA; w += gridDim.x) {
        const int b = w / a.ctasA, tile = w - b * a.ctasA;
        coop_phaseC<DT, NET, A2, GEN>(a, H, smraw, red, tile, b, B);
      }
      FGP_STAMP(5);
    }
    if (it + 1 < a.iters) grid_barrier(a.bar, ++nbar * gridDim.x, a.has_fit ? a.fit.state : nullptr);
  }
  FGP_STAMP(6);
  // every CTA has passed its last barrier when it gets here; the last one to leave re-arms the control words
  if (threadIdx.x == 0 && atomicAdd(a.bar + 1, 1u) == gridDim.x - 1) {
    a.bar[0] = 0u;
    a.bar[1] = 0u;
    __threadfence();
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

// defined once in fgp_mll_passb.cu
int launch_mll_passB(const MllArgs& a, const PassGeom& g, int B, bool net, cudaStream_t st);
int launch_mll_finalize(const MllArgs& a, int B, cudaStream_t st);
bool coop_enabled();  // FGP_COOP=0 switches the persistent kernel off (three launches per iteration, as in round 1)
int coop_max_ctas();  // FGP_COOP_CTAS caps the grid of the persistent kernel (tuning)

template <int DT, bool NET, bool A2, bool GEN>
static int launch_mll_coop(const MllArgs& a0, const PassGeom& g, int B, size_t smemAC, cudaStream_t st) {
  int rc;
  MllArgs a = a0;
#ifdef FGP_TIMING
  a.stamps = debug_stamp_buffer();
#endif
  int Bk = B;
  const size_t smem = smemAC > g.smemB ? smemAC : g.smemB;
  const int threads = g.threadsA > g.threadsB ? g.threadsA : g.threadsB;
  auto kernel = mll_coop_kernel<DT, NET, A2, GEN>;
  if ((rc = set_smem_attr(kernel, smem))) return rc;
  // resident CTAs of this (kernel, block, shared memory) triple; queried once per variant and shape
  static int cached_threads = 0, cached_per_sm = 0;
  static size_t cached_smem = 0;
  if (cached_threads != threads || cached_smem != smem) {
    int per_sm = 0;
    FGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    cached_threads = threads;
    cached_smem = smem;
    cached_per_sm = per_sm;
  }
  FGP_REQUIRE(cached_per_sm >= 1, "mll_coop: the persistent kernel does not fit an SM (%zu bytes of shared memory)", smem);
  int64_t grid = (int64_t)cached_per_sm * sm_count();
  const int64_t tiles = (int64_t)B * (a.ctasA > a.ctasB ? a.ctasA : a.ctasB);
  if (grid > tiles) grid = tiles;
  if (coop_max_ctas() > 0 && grid > coop_max_ctas()) grid = coop_max_ctas();
  void* params[2] = {(void*)&a, (void*)&Bk};
  FGP_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3((unsigned)grid), dim3(threads), params, smem, st));
  FGP_LAUNCH_NAMED("mll_coop", st);
  return FGP_OK;
}

template <int DT, bool NET, bool A2, bool GEN>
static int launch_mll(const MllArgs& a0, const PassGeom& g, int B, cudaStream_t st) {
  int rc;
#ifdef FGP_TIMING
  MllArgs a = a0;
  a.stamps = debug_stamp_buffer();
#else
  const MllArgs& a = a0;
#endif
  const size_t smemAC = (size_t)a.tab_off + ((GEN && NET) ? dnb2_table_bytes(a.d, g.l2 ? g.l1 + g.lntrA : g.l1, a.mmax) : 0);
  if (g.l2 == 0) {
    if ((rc = set_smem_attr(mll_single_kernel<DT, NET, A2, GEN>, smemAC))) return rc;
    launch_chain(mll_single_kernel<DT, NET, A2, GEN>, dim3(B), dim3(g.threadsA), smemAC, st, a);
    FGP_LAUNCH_NAMED("mll_single", st);
    return FGP_OK;
  }
  if (a.has_fit && a.bar && coop_enabled()) return launch_mll_coop<DT, NET, A2, GEN>(a, g, B, smemAC, st);
  if ((rc = set_smem_attr(mll_passA_kernel<DT, NET, A2, GEN>, smemAC))) return rc;
  launch_chain(mll_passA_kernel<DT, NET, A2, GEN>, dim3(a.ctasA, B), dim3(g.threadsA), smemAC, st, a);
  FGP_LAUNCH_NAMED("mll_passA", st);
  if ((rc = launch_mll_passB(a, g, B, NET, st))) return rc;
  if (a.want_grad) {
    if ((rc = set_smem_attr(mll_passC_kernel<DT, NET, A2, GEN>, smemAC))) return rc;
    launch_chain(mll_passC_kernel<DT, NET, A2, GEN>, dim3(a.ctasA, B), dim3(g.threadsA), smemAC, st, a);
    FGP_LAUNCH_NAMED("mll_passC", st);
  }
  if (a.has_fit) return FGP_OK;  // reduced (and stepped) by the last CTA of the last kernel
  return launch_mll_finalize(a, B, st);
}

// one translation unit per instantiation (fgp_mll_inst_*.cu) so that they compile in parallel
typedef int (*mll_launch_fn)(const MllArgs&, const PassGeom&, int, cudaStream_t);
#define FGP_MLL_DECLARE(name) int name(const MllArgs& a, const PassGeom& g, int B, cudaStream_t st)
FGP_MLL_DECLARE(mll_lat_x_a2_d2);
FGP_MLL_DECLARE(mll_lat_x_a2_d4);
FGP_MLL_DECLARE(mll_lat_x_a2_d8);
FGP_MLL_DECLARE(mll_lat_x_a2_d16);
FGP_MLL_DECLARE(mll_lat_x_gen_alpha);
FGP_MLL_DECLARE(mll_lat_z_a2_d2);
FGP_MLL_DECLARE(mll_lat_z_a2_d4);
FGP_MLL_DECLARE(mll_lat_z_a2_d8);
FGP_MLL_DECLARE(mll_lat_z_a2_d16);
FGP_MLL_DECLARE(mll_lat_z_gen_alpha);
FGP_MLL_DECLARE(mll_net_x_a2_d2);
FGP_MLL_DECLARE(mll_net_x_a2_d4);
FGP_MLL_DECLARE(mll_net_x_a2_d8);
FGP_MLL_DECLARE(mll_net_x_a2_d16);
FGP_MLL_DECLARE(mll_net_x_gen_alpha);
FGP_MLL_DECLARE(mll_net_z_a2_d2);
FGP_MLL_DECLARE(mll_net_z_a2_d4);
FGP_MLL_DECLARE(mll_net_z_a2_d8);
FGP_MLL_DECLARE(mll_net_z_a2_d16);
FGP_MLL_DECLARE(mll_net_z_gen_alpha);

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace fgp
