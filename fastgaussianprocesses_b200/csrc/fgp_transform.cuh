// K3 building blocks: block FFT-BRO (radix-16 register rounds) and block FWHT with register-fused I/O.
//
// A "block transform" is a length L = 2^l transform that lives in one CTA.  It runs as ceil(l/4) ROUNDS; a round loads
// the 2^R (R <= 4) elements of one butterfly group into registers, runs R radix-2 stages on them and hands them on.
// The first round takes its elements from a caller functor (global memory, or values computed on the fly) and the
// last round gives its results to a caller functor (global memory, or a fused epilogue), so a block transform makes
// (rounds - 1) shared-memory round trips and as many __syncthreads -- one for L = 256, two for L = 4096.
//   forward  = decimation in time:  bit-reversed-order input -> natural-order output, no permutation pass
//   inverse  = decimation in frequency with conjugate twiddles: natural-order input -> bit-reversed-order output,
//              rounds of the forward schedule in reverse order
// (exactly the conventions fixed by the reference's doubling recursion, fastgps/util.py:121-126.)
//
// Twiddles: stage (s+u) of a round pairs (c, c|2^u) with w = exp(-i pi (low + cm 2^s) / 2^(s+u)), cm = c mod 2^u.
// That factors as B_u(low) * exp(-i pi cm / 2^u): R table values per thread per round (none in the first round, where
// low = 0) times compile-time 16th roots of unity.
//
// Shared-memory layout: element e of transform tr sits at tr*LP + e + (e>>4) (one-in-sixteen padding, LP odd when
// several transforms share a CTA) -- conflict-free for every round stride and for both thread->group mappings.
//
// Larger n = L1*L2 are done in two passes over global/L2-resident memory (four-step):
//   pass A: contiguous length-L1 blocks b,  then multiply element q1 by w_n^{rev(b) q1}
//   pass B: stride-L1 columns q1, length-L2 transforms, in place.
#pragma once
#include <stdlib.h>
#include "fgp_common.cuh"

namespace fgp {

// launch bounds of the transform / MLL kernels: 256 threads x 2 CTAs per SM = up to 128 registers per thread (no spills);
// a 4096-point complex tile then gives every thread two radix-8 groups per round.  Measured against 512 x 2 (64 registers,
// spills) and 512 x 1 on B200: profiles/README.md.
#ifndef FGP_LB_THREADS
#define FGP_LB_THREADS 256
#endif
#ifndef FGP_LB_BLOCKS
#define FGP_LB_BLOCKS 2
#endif
// the real (FWHT) stand-alone kernels hold 16 doubles per thread: 64 registers, 4 CTAs per SM
#ifndef FGP_LB_BLOCKS_R
#define FGP_LB_BLOCKS_R 4
#endif

#ifndef FGP_FILL_UNROLL
#define FGP_FILL_UNROLL 2
#endif

constexpr int kFillUnroll = FGP_FILL_UNROLL;  // unroll of the rolled element loops (independent chains per thread)
constexpr int kTabLen = 4096;

// padding: one element in 2^PS (PS = 3 for the radix-8 complex rounds, 4 for the radix-16 real rounds)
constexpr int kPSC = 3, kPSR = 4;
constexpr int kRC = 3;  // log2 of the main complex radix: an 8-point round body (~4 KB of SASS) stays resident in the
                        // instruction caches; radix-16 bodies (x first/middle/last variants) did not -- see profiles/README.md
template <int PS>
__host__ __device__ __forceinline__ int padidx(int e) { return e + (e >> PS); }
__host__ __device__ __forceinline__ int padlen(int L, int ntr, int ps) { return L + (L >> ps) + (ntr > 1 ? 1 : 0); }

// twiddle tables, one caller-owned buffer of 3*kTabLen complex values
struct FftTables {
  const double2* stage;  // stage[h+p] = exp(-i pi p / h), h = 2^q, p < h (q <= 11)
  const double2* lo;     // lo[e] = w_n^e,        e < 4096
  const double2* hi;     // hi[e] = w_n^{4096 e}, e < max(1, n/4096)
};
__host__ __device__ __forceinline__ FftTables make_tables(const void* buf) {
  FftTables t;
  t.stage = (const double2*)buf;
  t.lo = t.stage + kTabLen;
  t.hi = t.lo + kTabLen;
  return t;
}
__device__ __forceinline__ double2 twiddle_n(const FftTables& t, uint32_t e) {  // w_n^e, e < n
  const double2 a = __ldg(t.lo + (e & (kTabLen - 1)));
  const double2 b = __ldg(t.hi + (e >> 12));
  return cmul(a, b);
}

// t * exp(-i pi k / 8), k = 0..7 a compile-time constant after unrolling
__device__ __forceinline__ double2 mul_root16(double2 t, int k) {
  constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173, H = 0.70710678118654752440;
  switch (k) {
    case 0: return t;
    case 1: return make_double2(fma(t.y, S1, t.x * C1), fma(-t.x, S1, t.y * C1));
    case 2: return make_double2((t.x + t.y) * H, (t.y - t.x) * H);
    case 3: return make_double2(fma(t.y, C1, t.x * S1), fma(-t.x, C1, t.y * S1));
    case 4: return make_double2(t.y, -t.x);
    case 5: return make_double2(fma(t.y, C1, -t.x * S1), fma(-t.x, C1, -t.y * S1));
    case 6: return make_double2((t.y - t.x) * H, -(t.x + t.y) * H);
    default: return make_double2(fma(t.y, S1, -t.x * C1), fma(-t.x, S1, -t.y * C1));
  }
}
// t * conj(exp(-i pi k / 8)) = t * exp(+i pi k / 8)
__device__ __forceinline__ double2 mul_root16c(double2 t, int k) {
  constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173, H = 0.70710678118654752440;
  switch (k) {
    case 0: return t;
    case 1: return make_double2(fma(-t.y, S1, t.x * C1), fma(t.x, S1, t.y * C1));
    case 2: return make_double2((t.x - t.y) * H, (t.y + t.x) * H);
    case 3: return make_double2(fma(-t.y, C1, t.x * S1), fma(t.x, C1, t.y * S1));
    case 4: return make_double2(-t.y, t.x);
    case 5: return make_double2(fma(-t.y, C1, -t.x * S1), fma(t.x, C1, -t.y * S1));
    case 6: return make_double2(-(t.x + t.y) * H, (t.x - t.y) * H);
    default: return make_double2(fma(-t.y, S1, -t.x * C1), fma(t.x, S1, -t.y * C1));
  }
}

// R radix-2 stages on the 2^R registers of one group.
// TW: twiddles of stage u come from the per-stage table, Wt[(1<<u) - 1 + cm] = exp(-i pi (low + cm 2^s) / 2^(s+u)) -- the
// very values (and the single complex multiply per butterfly) of the reference's radix-2 recursion, so the round-off
// matches it; !TW (first round, low = 0): compile-time 8th/16th roots of unity.
#ifndef FGP_TW_FACTORIZED
#define FGP_TW_FACTORIZED 0
#endif
// twiddle registers per group: 2^R - 1 table values, or (FGP_TW_FACTORIZED) R base values B_u = exp(-i pi low / 2^(s+u)) that
// are combined with compile-time roots exp(-i pi cm / 2^u) (3 loads instead of 7 per radix-8 group; one extra rounding on
// half of the butterflies)
template <int R>
struct TwCount {
#if FGP_TW_FACTORIZED
  static constexpr int value = R > 0 ? R : 1;
#else
  static constexpr int value = (1 << R) - 1 > 0 ? (1 << R) - 1 : 1;
#endif
};

template <int R, bool INV, bool TW>
__device__ __forceinline__ void butterflies(double2 (&v)[1 << R], const double2 (&Wt)[TwCount<R>::value]) {
  constexpr int RAD = 1 << R;
  if (!INV) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
#pragma unroll
      for (int c = 0; c < RAD; ++c) {
        if (c & (1 << u)) continue;
        const int cm = c & ((1 << u) - 1);
        double2 t = v[c | (1 << u)];
#if FGP_TW_FACTORIZED
        if (TW) t = cmul(Wt[u], t);
        t = mul_root16(t, cm << (3 - u));
#else
        if (TW)
          t = cmul(Wt[(1 << u) - 1 + cm], t);
        else
          t = mul_root16(t, cm << (3 - u));
#endif
        v[c | (1 << u)] = csub(v[c], t);
        v[c] = cadd(v[c], t);
      }
    }
  } else {
#pragma unroll
    for (int u = R - 1; u >= 0; --u) {
#pragma unroll
      for (int c = 0; c < RAD; ++c) {
        if (c & (1 << u)) continue;
        const int cm = c & ((1 << u) - 1);
        const double2 a = v[c], b = v[c | (1 << u)];
        v[c] = cadd(a, b);
        double2 t = csub(a, b);
#if FGP_TW_FACTORIZED
        t = mul_root16c(t, cm << (3 - u));
        if (TW) t = cmulc(Wt[u], t);
#else
        if (TW)
          t = cmulc(Wt[(1 << u) - 1 + cm], t);
        else
          t = mul_root16c(t, cm << (3 - u));
#endif
        v[c | (1 << u)] = t;
      }
    }
  }
}

// group decomposition of a round at stage base s: group j of a length-2^l transform owns elements base + (c << s)
struct GroupIdx {
  int tr, low, base;
};
template <int R, bool TRFAST>
__device__ __forceinline__ GroupIdx group_of(int g, int s, int l, int lntr) {
  GroupIdx o;
  int j;
  if (TRFAST) {
    o.tr = g & ((1 << lntr) - 1);
    j = g >> lntr;
  } else {
    o.tr = g >> (l - R);
    j = g & ((1 << (l - R)) - 1);
  }
  o.low = j & ((1 << s) - 1);
  o.base = o.low + ((j >> s) << (s + R));
  return o;
}

template <int R>
__device__ __forceinline__ void load_twiddles(double2 (&Wt)[TwCount<R>::value], const double2* __restrict__ tw, int s, int low) {
#if FGP_TW_FACTORIZED
#pragma unroll
  for (int u = 0; u < R; ++u) Wt[u] = __ldg(tw + (1 << (s + u)) + low);
#else
#pragma unroll
  for (int u = 0; u < R; ++u)
#pragma unroll
    for (int cm = 0; cm < (1 << u); ++cm) Wt[(1 << u) - 1 + cm] = __ldg(tw + (1 << (s + u)) + low + (cm << s));
#endif
}

struct SmemC {
  double2* sm;
  int LP;
  __device__ __forceinline__ double2 operator()(int tr, int idx) const { return sm[tr * LP + padidx<kPSC>(idx)]; }
  __device__ __forceinline__ void operator()(int tr, int idx, double2 v) const { sm[tr * LP + padidx<kPSC>(idx)] = v; }
};
template <class T>
struct is_smemc {
  static constexpr bool value = false;
};
template <>
struct is_smemc<SmemC> {
  static constexpr bool value = true;
};

// One round over all groups of the CTA's 2^lntr transforms.  ld(tr, idx) -> double2 ; st(tr, idx, value).
// TRFAST: consecutive threads take consecutive transforms (column access); otherwise consecutive groups.
// Shared-memory ends use linear addressing: every schedule has s == 0 (the group sits inside one padding block) or
// s >= kPSC (the padded offset of element c is c * (2^s + 2^(s-kPSC))), so one pointer and one stride replace the
// per-element index arithmetic.
template <int R, bool INV, bool TRFAST, class Ld, class St>
__device__ __forceinline__ void fft_round_io(int s, int l, int lntr, const double2* __restrict__ tw, Ld ld, St st) {
  constexpr int RAD = 1 << R;
  const int total = 1 << (lntr + l - R);
  const int stp = s == 0 ? 1 : (1 << s) + (1 << (s - kPSC));
  for (int g = threadIdx.x; g < total; g += blockDim.x) {
    const GroupIdx G = group_of<R, TRFAST>(g, s, l, lntr);
    double2 B[TwCount<R>::value];
    if (s > 0) load_twiddles<R>(B, tw, s, G.low);
    double2 v[RAD];
    if constexpr (is_smemc<Ld>::value) {
      const double2* p = ld.sm + G.tr * ld.LP + padidx<kPSC>(G.base);
#pragma unroll
      for (int c = 0; c < RAD; ++c) v[c] = p[c * stp];
    } else {
#pragma unroll
      for (int c = 0; c < RAD; ++c) v[c] = ld(G.tr, G.base + (c << s));
    }
    if (s > 0)
      butterflies<R, INV, true>(v, B);
    else
      butterflies<R, INV, false>(v, B);
    if constexpr (is_smemc<St>::value) {
      double2* p = st.sm + G.tr * st.LP + padidx<kPSC>(G.base);
#pragma unroll
      for (int c = 0; c < RAD; ++c) p[c * stp] = v[c];
    } else {
#pragma unroll
      for (int c = 0; c < RAD; ++c) st(G.tr, G.base + (c << s), v[c]);
    }
  }
}

#define FGP_R_DISPATCH(r, CALL4, CALL3, CALL2, CALL1) \
  switch (r) {                                        \
    case 4: CALL4; break;                             \
    case 3: CALL3; break;                             \
    case 2: CALL2; break;                             \
    default: CALL1; break;                            \
  }

struct SmemTag {};  // "the data is already in / should stay in shared memory" (the caller synchronises)
template <class T>
struct is_smem_tag {
  static constexpr bool value = false;
};
template <>
struct is_smem_tag<SmemTag> {
  static constexpr bool value = true;
};

// rolled element loops between a caller functor and shared memory: compact code for heavy functors (kernel evaluation,
// log/divide epilogues, gradient contraction) -- unrolling those 8 or 16 times per round overflowed the instruction caches
template <bool TRFAST, class F>
__device__ __forceinline__ void tile_fill_c(const SmemC& S, int l, int lntr, F f) {  // S(tr,idx) = f(tr,idx)
  const int total = 1 << (lntr + l);
#pragma unroll kFillUnroll
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int tr = TRFAST ? (e & ((1 << lntr) - 1)) : (e >> l);
    const int idx = TRFAST ? (e >> lntr) : (e & ((1 << l) - 1));
    S(tr, idx, f(tr, idx));
  }
}
template <bool TRFAST, class F>
__device__ __forceinline__ void tile_map_c(const SmemC& S, int l, int lntr, F f) {  // S(tr,idx) = f(tr,idx,S(tr,idx))
  const int total = 1 << (lntr + l);
#pragma unroll 2
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int tr = TRFAST ? (e & ((1 << lntr) - 1)) : (e >> l);
    const int idx = TRFAST ? (e >> lntr) : (e & ((1 << l) - 1));
    S(tr, idx, f(tr, idx, S(tr, idx)));
  }
}
template <bool TRFAST, class F>
__device__ __forceinline__ void tile_drain_c(const SmemC& S, int l, int lntr, F f) {  // f(tr,idx,S(tr,idx))
  const int total = 1 << (lntr + l);
#pragma unroll kFillUnroll
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int tr = TRFAST ? (e & ((1 << lntr) - 1)) : (e >> l);
    const int idx = TRFAST ? (e >> lntr) : (e & ((1 << l) - 1));
    f(tr, idx, S(tr, idx));
  }
}

// shared-memory to shared-memory round of r <= kRC stages: ONE out-of-line copy per direction, shared by every call site
template <bool INV>
__device__ __noinline__ void fft_round_smem(int r, int s, int l, int lntr, const double2* tw, double2* sm, int LP) {
  const SmemC S{sm, LP};
  switch (r) {
    case 3: fft_round_io<3, INV, false>(s, l, lntr, tw, S, S); break;
    case 2: fft_round_io<2, INV, false>(s, l, lntr, tw, S, S); break;
    default: fft_round_io<1, INV, false>(s, l, lntr, tw, S, S); break;
  }
}

// Round schedule of a length-2^l transform, l >= 2 kRC: stages [0,3) first, then the remainder (l mod 3 stages), then
// threes, ending with [l-3, l); the inverse runs it backwards.  First and last rounds are always full radix-8 rounds, so
// only they exist in (global <-> registers) variants; pass SmemTag for an end that is already / should stay in shared
// memory (the caller synchronises that end).  No trailing sync after a functor end.
template <bool TRFAST, class GLd, class GSt>
__device__ __forceinline__ void block_fft_fwd_io(double2* sm, int l, int lntr, int LP, const double2* tw, GLd gld, GSt gst) {
  constexpr bool IN_S = is_smem_tag<GLd>::value, OUT_S = is_smem_tag<GSt>::value;
  const SmemC S{sm, LP};
  if (l == 0) {  // length-1 transforms are the identity (util.py:170)
    if constexpr (!IN_S) {
      tile_fill_c<TRFAST>(S, 0, lntr, gld);
      __syncthreads();
    }
    if constexpr (!OUT_S) tile_drain_c<TRFAST>(S, 0, lntr, gst);
    return;
  }
  if (l < 2 * kRC) {  // small transforms: everything through shared memory
    if constexpr (!IN_S) {
      tile_fill_c<TRFAST>(S, l, lntr, gld);
      __syncthreads();
    }
    for (int s = 0; s < l; s += kRC) {
      fft_round_smem<false>(l - s >= kRC ? kRC : l - s, s, l, lntr, tw, sm, LP);
      if (s + kRC < l || !OUT_S) __syncthreads();
    }
    if constexpr (!OUT_S) tile_drain_c<TRFAST>(S, l, lntr, gst);
    return;
  }
  if constexpr (IN_S)
    fft_round_smem<false>(kRC, 0, l, lntr, tw, sm, LP);
  else
    fft_round_io<kRC, false, TRFAST>(0, l, lntr, tw, gld, S);
  __syncthreads();
  int s = kRC;
  const int rem = (l - 2 * kRC) % kRC;
  if (rem) {
    fft_round_smem<false>(rem, s, l, lntr, tw, sm, LP);
    s += rem;
    __syncthreads();
  }
  while (s < l - kRC) {
    fft_round_smem<false>(kRC, s, l, lntr, tw, sm, LP);
    s += kRC;
    __syncthreads();
  }
  if constexpr (OUT_S)
    fft_round_smem<false>(kRC, s, l, lntr, tw, sm, LP);
  else
    fft_round_io<kRC, false, TRFAST>(s, l, lntr, tw, S, gst);
}

// Inverse block transform (mirror schedule).
template <bool TRFAST, class GLd, class GSt>
__device__ __forceinline__ void block_fft_inv_io(double2* sm, int l, int lntr, int LP, const double2* tw, GLd gld, GSt gst) {
  constexpr bool IN_S = is_smem_tag<GLd>::value, OUT_S = is_smem_tag<GSt>::value;
  const SmemC S{sm, LP};
  if (l == 0) {
    if constexpr (!IN_S) {
      tile_fill_c<TRFAST>(S, 0, lntr, gld);
      __syncthreads();
    }
    if constexpr (!OUT_S) tile_drain_c<TRFAST>(S, 0, lntr, gst);
    return;
  }
  if (l < 2 * kRC) {
    if constexpr (!IN_S) {
      tile_fill_c<TRFAST>(S, l, lntr, gld);
      __syncthreads();
    }
    for (int s = ((l - 1) / kRC) * kRC; s >= 0; s -= kRC) {
      fft_round_smem<true>(l - s >= kRC ? kRC : l - s, s, l, lntr, tw, sm, LP);
      if (s > 0 || !OUT_S) __syncthreads();
    }
    if constexpr (!OUT_S) tile_drain_c<TRFAST>(S, l, lntr, gst);
    return;
  }
  int s = l - kRC;
  if constexpr (IN_S)
    fft_round_smem<true>(kRC, s, l, lntr, tw, sm, LP);
  else
    fft_round_io<kRC, true, TRFAST>(s, l, lntr, tw, gld, S);
  __syncthreads();
  const int rem = (l - 2 * kRC) % kRC;
  while (s > kRC + rem) {
    s -= kRC;
    fft_round_smem<true>(kRC, s, l, lntr, tw, sm, LP);
    __syncthreads();
  }
  if (rem) {
    fft_round_smem<true>(rem, kRC, l, lntr, tw, sm, LP);
    __syncthreads();
  }
  if constexpr (OUT_S)
    fft_round_smem<true>(kRC, 0, l, lntr, tw, sm, LP);
  else
    fft_round_io<kRC, true, TRFAST>(0, l, lntr, tw, S, gst);
}

// ---------------------------------------------------------------------------------------------------------------
// real Walsh-Hadamard rounds (radix 16; no twiddles; self-inverse; stage order is free)
// ---------------------------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void wht_butterflies(double (&v)[1 << R]) {
  constexpr int RAD = 1 << R;
#pragma unroll
  for (int u = 0; u < R; ++u) {
#pragma unroll
    for (int c = 0; c < RAD; ++c) {
      if (c & (1 << u)) continue;
      const double a = v[c], b = v[c | (1 << u)];
      v[c] = a + b;
      v[c | (1 << u)] = a - b;
    }
  }
}

struct SmemR {
  double* sm;
  int LP;
  __device__ __forceinline__ double operator()(int tr, int idx) const { return sm[tr * LP + padidx<kPSR>(idx)]; }
  __device__ __forceinline__ void operator()(int tr, int idx, double v) const { sm[tr * LP + padidx<kPSR>(idx)] = v; }
};
template <class T>
struct is_smemr {
  static constexpr bool value = false;
};
template <>
struct is_smemr<SmemR> {
  static constexpr bool value = true;
};

// One radix-2^R FWHT round.  Shared-memory ends use linear addressing (every schedule below has s == 0 or s >= kPSR).
template <int R, bool TRFAST, class Ld, class St>
__device__ __forceinline__ void wht_round_io(int s, int l, int lntr, Ld ld, St st) {
  constexpr int RAD = 1 << R;
  const int total = 1 << (lntr + l - R);
  const int stp = s == 0 ? 1 : (1 << s) + (1 << (s - kPSR));
  for (int g = threadIdx.x; g < total; g += blockDim.x) {
    const GroupIdx G = group_of<R, TRFAST>(g, s, l, lntr);
    double v[RAD];
    if constexpr (is_smemr<Ld>::value) {
      const double* p = ld.sm + G.tr * ld.LP + padidx<kPSR>(G.base);
#pragma unroll
      for (int c = 0; c < RAD; ++c) v[c] = p[c * stp];
    } else {
#pragma unroll
      for (int c = 0; c < RAD; ++c) v[c] = ld(G.tr, G.base + (c << s));
    }
    wht_butterflies<R>(v);
    if constexpr (is_smemr<St>::value) {
      double* p = st.sm + G.tr * st.LP + padidx<kPSR>(G.base);
#pragma unroll
      for (int c = 0; c < RAD; ++c) p[c * stp] = v[c];
    } else {
#pragma unroll
      for (int c = 0; c < RAD; ++c) st(G.tr, G.base + (c << s), v[c]);
    }
  }
}

template <bool TRFAST, class F>
__device__ __forceinline__ void tile_fill_r(const SmemR& S, int l, int lntr, F f) {
  const int total = 1 << (lntr + l);
#pragma unroll 1
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int tr = TRFAST ? (e & ((1 << lntr) - 1)) : (e >> l);
    const int idx = TRFAST ? (e >> lntr) : (e & ((1 << l) - 1));
    S(tr, idx, f(tr, idx));
  }
}
template <bool TRFAST, class F>
__device__ __forceinline__ void tile_map_r(const SmemR& S, int l, int lntr, F f) {
  const int total = 1 << (lntr + l);
#pragma unroll 1
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int tr = TRFAST ? (e & ((1 << lntr) - 1)) : (e >> l);
    const int idx = TRFAST ? (e >> lntr) : (e & ((1 << l) - 1));
    S(tr, idx, f(tr, idx, S(tr, idx)));
  }
}
template <bool TRFAST, class F>
__device__ __forceinline__ void tile_drain_r(const SmemR& S, int l, int lntr, F f) {
  const int total = 1 << (lntr + l);
#pragma unroll 1
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int tr = TRFAST ? (e & ((1 << lntr) - 1)) : (e >> l);
    const int idx = TRFAST ? (e >> lntr) : (e & ((1 << l) - 1));
    f(tr, idx, S(tr, idx));
  }
}

static __device__ __noinline__ void wht_round_smem(int r, int s, int l, int lntr, double* sm, int LP) {
  const SmemR S{sm, LP};
  FGP_R_DISPATCH(r, (wht_round_io<4, false>(s, l, lntr, S, S)), (wht_round_io<3, false>(s, l, lntr, S, S)),
                 (wht_round_io<2, false>(s, l, lntr, S, S)), (wht_round_io<1, false>(s, l, lntr, S, S)))
}

// Round schedules for the FWHT (stage order is free): s[k], r[k] = first stage and number of stages of round k.
// When l >= 8 both schedules start and end with full radix-16 rounds (only those exist in global<->register variants).
struct WhtSched {
  int n;
  int s[5], r[5];
};
// bottom-up: [0,4), remainder, fours
__device__ __forceinline__ WhtSched wht_sched_up(int l) {
  WhtSched q;
  q.n = 0;
  if (l < 8) {
    for (int s = 0; s < l; s += 4) q.s[q.n] = s, q.r[q.n] = l - s >= 4 ? 4 : l - s, ++q.n;
    return q;
  }
  const int rem = l & 3;
  q.s[q.n] = 0, q.r[q.n] = 4, ++q.n;
  if (rem) q.s[q.n] = 4, q.r[q.n] = rem, ++q.n;
  for (int s = 4 + rem; s < l; s += 4) q.s[q.n] = s, q.r[q.n] = 4, ++q.n;
  return q;
}
// coalesced at both ends for contiguous tiles (l >= 12): the top four stages first (a thread's elements are 2^(l-4)
// apart, so consecutive threads touch consecutive addresses), then [0,4), the remainder, and upwards, ending with the
// four stages just below the top.  Every round keeps s == 0 or s >= 4 (linear shared-memory addressing).
__device__ __forceinline__ WhtSched wht_sched_coalesced(int l) {
  if (l < 12) return wht_sched_up(l);
  WhtSched q;
  const int rem = l & 3;
  q.n = 0;
  q.s[q.n] = l - 4, q.r[q.n] = 4, ++q.n;
  q.s[q.n] = 0, q.r[q.n] = 4, ++q.n;
  if (rem) q.s[q.n] = 4, q.r[q.n] = rem, ++q.n;
  for (int s = 4 + rem; s < l - 4; s += 4) q.s[q.n] = s, q.r[q.n] = 4, ++q.n;
  return q;
}

// Block FWHT following schedule q.  gld / gst as for the FFT drivers (SmemTag = stays in shared memory).
template <bool TRFAST, class GLd, class GSt>
__device__ __forceinline__ void block_wht_io(double* sm, int l, int lntr, int LP, const WhtSched& q, GLd gld, GSt gst) {
  constexpr bool IN_S = is_smem_tag<GLd>::value, OUT_S = is_smem_tag<GSt>::value;
  const SmemR S{sm, LP};
  if (l < 8) {  // small transforms (and l = 0: identity): everything through shared memory
    if constexpr (!IN_S) {
      tile_fill_r<TRFAST>(S, l, lntr, gld);
      __syncthreads();
    }
    for (int k = 0; k < q.n; ++k) {
      wht_round_smem(q.r[k], q.s[k], l, lntr, sm, LP);
      if (k + 1 < q.n || !OUT_S) __syncthreads();
    }
    if constexpr (!OUT_S) tile_drain_r<TRFAST>(S, l, lntr, gst);
    return;
  }
  if constexpr (IN_S)
    wht_round_smem(4, q.s[0], l, lntr, sm, LP);
  else
    wht_round_io<4, TRFAST>(q.s[0], l, lntr, gld, S);
  __syncthreads();
  for (int k = 1; k < q.n - 1; ++k) {
    wht_round_smem(q.r[k], q.s[k], l, lntr, sm, LP);
    __syncthreads();
  }
  if constexpr (OUT_S)
    wht_round_smem(4, q.s[q.n - 1], l, lntr, sm, LP);
  else
    wht_round_io<4, TRFAST>(q.s[q.n - 1], l, lntr, S, gst);
}

// ---------------------------------------------------------------------------------------------------------------
// two-pass geometry
// ---------------------------------------------------------------------------------------------------------------
struct PassGeom {
  int m;     // log2 n
  int l1;    // log2 of the contiguous block length (pass A)
  int l2;    // log2 of the strided transform length (pass B); 0 => single pass
  int lntrA; // log2 transforms per CTA in pass A
  int lntrB; // log2 columns per CTA in pass B
  int ntrA, ntrB;
  int LPA, LPB;
  int threadsA, threadsB;  // one radix-16 group per thread per round
  int64_t ctasA, ctasB;    // per batch item
  size_t smemA, smemB;
};

// Tile capacities (log2 elements per CTA).  Pass A keeps the contiguous block as long as shared memory allows (64 KiB)
// so that pass B's strided transforms are short; pass B takes 8 adjacent columns (128-byte global segments for complex
// data) whenever that fits.  FGP_CAP_C / FGP_CAP_R / FGP_COLS_LOG2 override the defaults for tuning runs.
static inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
// standalone: the stand-alone real transforms (fgp_fwht) take 128 KiB tiles for n >= 2^24 (pass-B strips of 8 columns instead
// of 4: 2^24 159 -> 146 us, 2^26 1149 -> 753 us on B200); the fused MLL kernels keep 64 KiB tiles (they carry generator tables).
// mll: the fused eigen-solve kernels (fgp_mll.cuh): 2^11-point tiles and 4-column strips up to n = 2^20 -- two tiles per SM in
// flight in every pass (measured at n = 2^20, d = 8 on B200: 47.4 us per iteration against 49.5 us with 2^12 / 8, 58.4 against 62.4 cold)
// batch (mll only): hyperparameter sets evaluated by one call.  From about two single-GP waves of work on (batch * n >= 2^21, n >= 2^18) the
// passes are throughput-bound and the 2^12-point tiles / 8-column strips win again (measured, tools/tune_batched.py: 64 x 2^18 447 -> 392 us per
// batched iteration, 16 x 2^20 462 -> 410, 8 x 2^18 72 -> 65; 4 x 2^18 45 -> 47, 64 x 2^16 138 -> 152: not below)
static inline PassGeom make_geom(int64_t n, bool cplx, bool standalone = false, bool mll = false, int64_t batch = 1) {
  PassGeom g;
  const size_t elem = cplx ? sizeof(double2) : sizeof(double);
  g.m = ilog2(n);
  // defaults from the B200 sweeps in profiles/README.md: 32 KiB real tiles and 16-column pass-B strips while the data is
  // L2-sized (more CTAs in different phases per SM), 64 KiB tiles for the HBM-sized transforms
  static const int capC = env_int("FGP_CAP_C", 0), capR = env_int("FGP_CAP_R", 0), colsEnv = env_int("FGP_COLS_LOG2", -1);
  static const int hardC = env_int("FGP_HARD_C", 0), hardR = env_int("FGP_HARD_R", 13);
  // complex: 2^11-point tiles and 4-column strips while one iteration's CTAs fit one wave (n <= 2^18), 2^12 / 8 above
  const bool fat = mll && ilog2(n) >= 18 && batch * n >= (int64_t(1) << 21);
  const int small_c = mll ? (fat ? 0 : 20) : 18;  // complex tiles: 2^11 points / 4 columns up to here
  int cap = cplx ? (capC ? capC : (g.m <= small_c ? 11 : 12)) : (capR ? capR : (g.m <= 22 ? 12 : ((standalone && g.m >= 25) ? 14 : 13)));
  // largest tile: 64 KiB of elements; the 2^24-point FFT takes 128 KiB pass-B tiles (2 columns instead of 1)
  static const int hardREnv = env_int("FGP_HARD_R", 0);
  const int hard = cplx ? (hardC ? hardC : (g.m >= 24 ? 13 : 12)) : (hardREnv ? hardREnv : ((standalone && g.m >= 24) ? 14 : hardR));
  if (cap > hard) cap = hard;
  const int colsLog = colsEnv >= 0 ? colsEnv : (cplx ? (g.m <= small_c ? 2 : 3) : (g.m <= 22 ? 4 : 3));
  int tileA, tileB = 0;
  if (g.m <= cap) {
    g.l1 = g.m;
    g.l2 = 0;
    tileA = g.m > cap - 1 ? g.m : cap - 1;
  } else {
    g.l1 = cap;
    g.l2 = g.m - g.l1;
    if (g.l2 > hard) {  // keep the strided transform inside one CTA
      g.l2 = hard;
      g.l1 = g.m - g.l2;
    }
    tileA = g.l1;
    tileB = g.l2 + colsLog;
    if (tileB > hard) tileB = hard;
    if (tileB > g.l2 + g.l1) tileB = g.l2 + g.l1;
    if (tileB < g.l2) tileB = g.l2;
  }
  g.lntrA = tileA - g.l1;
  g.ntrA = 1 << g.lntrA;
  g.lntrB = g.l2 ? tileB - g.l2 : 0;
  g.ntrB = 1 << g.lntrB;
  g.LPA = padlen(1 << g.l1, g.ntrA, cplx ? kPSC : kPSR);
  g.LPB = padlen(1 << g.l2, g.ntrB, cplx ? kPSC : kPSR);
  static const int tdiv = env_int("FGP_THREAD_DIV", 1);  // tuning: fewer threads, more groups per thread per round
  // more threads, fewer groups per thread per round: complex tiles run one radix-8 group per thread per round (measured
  // 55.3 -> 51.5 us per lattice fit iteration at n = 2^20, profiles/README.md)
  static const int tmulEnv = env_int("FGP_THREAD_MUL", 0);
  const int tmul = tmulEnv ? tmulEnv : (cplx ? 2 : 1);
  auto thr = [tmul](int tile) {
    int t = (1 << tile) / 16 * tmul / tdiv;
    if (t < 32) t = 32;
    if (t > FGP_LB_THREADS) t = FGP_LB_THREADS;
    return t;
  };
  g.threadsA = thr(tileA);
  g.threadsB = thr(tileB);
  g.ctasA = (n >> g.l1) >> g.lntrA;
  if (g.ctasA < 1) g.ctasA = 1;
  g.ctasB = g.l2 ? ((int64_t(1) << g.l1) >> g.lntrB) : 0;
  g.smemA = (size_t)g.ntrA * g.LPA * elem;
  g.smemB = (size_t)g.ntrB * g.LPB * elem;
  return g;
}

__device__ __forceinline__ uint32_t brev_bits(uint32_t v, int bits) { return bits ? (__brev(v) >> (32 - bits)) : 0u; }

}  // namespace fgp
