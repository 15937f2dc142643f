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
#include "fgp_common.cuh"

namespace fgp {

constexpr int kTabLen = 4096;

__host__ __device__ __forceinline__ int padidx(int e) { return e + (e >> 4); }
__host__ __device__ __forceinline__ int padlen(int L, int ntr) { return L + (L >> 4) + (ntr > 1 ? 1 : 0); }

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

// R radix-2 stages on the 2^R registers of one group.  B[u] = exp(-i pi low / 2^(s+u)) (ignored when !TW: low = 0).
template <int R, bool INV, bool TW>
__device__ __forceinline__ void butterflies(double2 (&v)[1 << R], const double2 (&B)[R > 0 ? R : 1]) {
  constexpr int RAD = 1 << R;
  if (!INV) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
#pragma unroll
      for (int c = 0; c < RAD; ++c) {
        if (c & (1 << u)) continue;
        const int cm = c & ((1 << u) - 1);
        double2 t = v[c | (1 << u)];
        if (TW) t = cmul(B[u], t);
        t = mul_root16(t, cm << (3 - u));
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
        double2 t = mul_root16c(csub(a, b), cm << (3 - u));
        if (TW) t = cmulc(B[u], t);
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
__device__ __forceinline__ void load_bases(double2 (&B)[R > 0 ? R : 1], const double2* __restrict__ tw, int s, int low) {
#pragma unroll
  for (int u = 0; u < R; ++u) B[u] = __ldg(tw + (1 << (s + u)) + low);
}

// One round over all groups of the CTA's 2^lntr transforms.  ld(tr, idx) -> double2 ; st(tr, idx, value).
// TRFAST: consecutive threads take consecutive transforms (column access); otherwise consecutive groups.
template <int R, bool INV, bool TRFAST, class Ld, class St>
__device__ __forceinline__ void fft_round_io(int s, int l, int lntr, const double2* __restrict__ tw, Ld ld, St st) {
  constexpr int RAD = 1 << R;
  const int total = 1 << (lntr + l - R);
  for (int g = threadIdx.x; g < total; g += blockDim.x) {
    const GroupIdx G = group_of<R, TRFAST>(g, s, l, lntr);
    double2 B[R > 0 ? R : 1];
    if (s > 0) load_bases<R>(B, tw, s, G.low);
    double2 v[RAD];
#pragma unroll
    for (int c = 0; c < RAD; ++c) v[c] = ld(G.tr, G.base + (c << s));
    if (s > 0)
      butterflies<R, INV, true>(v, B);
    else
      butterflies<R, INV, false>(v, B);
#pragma unroll
    for (int c = 0; c < RAD; ++c) st(G.tr, G.base + (c << s), v[c]);
  }
}

// forward round + elementwise map + inverse round on the same registers (spectral epilogue between the transforms)
template <int R, bool TRFAST, class Ld, class Mid, class St>
__device__ __forceinline__ void fft_round_fwd_mid_inv(int s, int l, int lntr, const double2* __restrict__ tw, Ld ld, Mid mid, St st) {
  constexpr int RAD = 1 << R;
  const int total = 1 << (lntr + l - R);
  for (int g = threadIdx.x; g < total; g += blockDim.x) {
    const GroupIdx G = group_of<R, TRFAST>(g, s, l, lntr);
    double2 B[R > 0 ? R : 1];
    if (s > 0) load_bases<R>(B, tw, s, G.low);
    double2 v[RAD];
#pragma unroll
    for (int c = 0; c < RAD; ++c) v[c] = ld(G.tr, G.base + (c << s));
    if (s > 0)
      butterflies<R, false, true>(v, B);
    else
      butterflies<R, false, false>(v, B);
#pragma unroll
    for (int c = 0; c < RAD; ++c) v[c] = mid(G.tr, G.base + (c << s), v[c]);
    if (s > 0)
      butterflies<R, true, true>(v, B);
    else
      butterflies<R, true, false>(v, B);
#pragma unroll
    for (int c = 0; c < RAD; ++c) st(G.tr, G.base + (c << s), v[c]);
  }
}

struct SmemC {
  double2* sm;
  int LP;
  __device__ __forceinline__ double2 operator()(int tr, int idx) const { return sm[tr * LP + padidx(idx)]; }
  __device__ __forceinline__ void operator()(int tr, int idx, double2 v) const { sm[tr * LP + padidx(idx)] = v; }
};

#define FGP_R_DISPATCH(r, CALL4, CALL3, CALL2, CALL1) \
  switch (r) {                                        \
    case 4: CALL4; break;                             \
    case 3: CALL3; break;                             \
    case 2: CALL2; break;                             \
    default: CALL1; break;                            \
  }

// Forward block transform.  gld feeds the first round, gst consumes the last; returns WITHOUT a trailing sync
// (the last round does not touch shared memory unless the functor does).
template <bool TRFAST, class GLd, class GSt>
__device__ __forceinline__ void block_fft_fwd_io(double2* sm, int l, int lntr, int LP, const double2* tw, GLd gld, GSt gst) {
  const SmemC S{sm, LP};
  if (l <= 4) {
    FGP_R_DISPATCH(l, (fft_round_io<4, false, TRFAST>(0, l, lntr, tw, gld, gst)), (fft_round_io<3, false, TRFAST>(0, l, lntr, tw, gld, gst)),
                   (fft_round_io<2, false, TRFAST>(0, l, lntr, tw, gld, gst)), (fft_round_io<1, false, TRFAST>(0, l, lntr, tw, gld, gst)))
    return;
  }
  fft_round_io<4, false, TRFAST>(0, l, lntr, tw, gld, S);
  __syncthreads();
  int s = 4;
  while (l - s > 4) {
    fft_round_io<4, false, false>(s, l, lntr, tw, S, S);
    s += 4;
    __syncthreads();
  }
  const int r = l - s;
  FGP_R_DISPATCH(r, (fft_round_io<4, false, TRFAST>(s, l, lntr, tw, S, gst)), (fft_round_io<3, false, TRFAST>(s, l, lntr, tw, S, gst)),
                 (fft_round_io<2, false, TRFAST>(s, l, lntr, tw, S, gst)), (fft_round_io<1, false, TRFAST>(s, l, lntr, tw, S, gst)))
}

// Inverse block transform (mirror schedule).
template <bool TRFAST, class GLd, class GSt>
__device__ __forceinline__ void block_fft_inv_io(double2* sm, int l, int lntr, int LP, const double2* tw, GLd gld, GSt gst) {
  const SmemC S{sm, LP};
  if (l <= 4) {
    FGP_R_DISPATCH(l, (fft_round_io<4, true, TRFAST>(0, l, lntr, tw, gld, gst)), (fft_round_io<3, true, TRFAST>(0, l, lntr, tw, gld, gst)),
                   (fft_round_io<2, true, TRFAST>(0, l, lntr, tw, gld, gst)), (fft_round_io<1, true, TRFAST>(0, l, lntr, tw, gld, gst)))
    return;
  }
  int s = ((l - 1) >> 2) << 2;
  const int r = l - s;
  FGP_R_DISPATCH(r, (fft_round_io<4, true, TRFAST>(s, l, lntr, tw, gld, S)), (fft_round_io<3, true, TRFAST>(s, l, lntr, tw, gld, S)),
                 (fft_round_io<2, true, TRFAST>(s, l, lntr, tw, gld, S)), (fft_round_io<1, true, TRFAST>(s, l, lntr, tw, gld, S)))
  __syncthreads();
  s -= 4;
  while (s > 0) {
    fft_round_io<4, true, false>(s, l, lntr, tw, S, S);
    s -= 4;
    __syncthreads();
  }
  fft_round_io<4, true, TRFAST>(0, l, lntr, tw, S, gst);
}

// Forward transform, spectral map, inverse transform of the same 2^lntr x 2^l tile: the top round of the forward
// transform, the map and the first round of the inverse share registers.
template <bool TRFAST, class GLd, class Mid, class GSt>
__device__ __forceinline__ void block_fft_fwd_mid_inv_io(double2* sm, int l, int lntr, int LP, const double2* tw, GLd gld, Mid mid, GSt gst) {
  const SmemC S{sm, LP};
  if (l <= 4) {
    FGP_R_DISPATCH(l, (fft_round_fwd_mid_inv<4, TRFAST>(0, l, lntr, tw, gld, mid, gst)), (fft_round_fwd_mid_inv<3, TRFAST>(0, l, lntr, tw, gld, mid, gst)),
                   (fft_round_fwd_mid_inv<2, TRFAST>(0, l, lntr, tw, gld, mid, gst)), (fft_round_fwd_mid_inv<1, TRFAST>(0, l, lntr, tw, gld, mid, gst)))
    return;
  }
  fft_round_io<4, false, TRFAST>(0, l, lntr, tw, gld, S);
  __syncthreads();
  int s = 4;
  while (l - s > 4) {
    fft_round_io<4, false, false>(s, l, lntr, tw, S, S);
    s += 4;
    __syncthreads();
  }
  const int r = l - s;
  // each thread reads and rewrites only its own group: no hazard inside the round
  FGP_R_DISPATCH(r, (fft_round_fwd_mid_inv<4, TRFAST>(s, l, lntr, tw, S, mid, S)), (fft_round_fwd_mid_inv<3, TRFAST>(s, l, lntr, tw, S, mid, S)),
                 (fft_round_fwd_mid_inv<2, TRFAST>(s, l, lntr, tw, S, mid, S)), (fft_round_fwd_mid_inv<1, TRFAST>(s, l, lntr, tw, S, mid, S)))
  __syncthreads();
  s -= 4;
  while (s > 0) {
    fft_round_io<4, true, false>(s, l, lntr, tw, S, S);
    s -= 4;
    __syncthreads();
  }
  fft_round_io<4, true, TRFAST>(0, l, lntr, tw, S, gst);
}

// ---------------------------------------------------------------------------------------------------------------
// real Walsh-Hadamard rounds (no twiddles; self-inverse; stage order is irrelevant)
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

template <int R, bool TRFAST, class Ld, class St>
__device__ __forceinline__ void wht_round_io(int s, int l, int lntr, Ld ld, St st) {
  constexpr int RAD = 1 << R;
  const int total = 1 << (lntr + l - R);
  for (int g = threadIdx.x; g < total; g += blockDim.x) {
    const GroupIdx G = group_of<R, TRFAST>(g, s, l, lntr);
    double v[RAD];
#pragma unroll
    for (int c = 0; c < RAD; ++c) v[c] = ld(G.tr, G.base + (c << s));
    wht_butterflies<R>(v);
#pragma unroll
    for (int c = 0; c < RAD; ++c) st(G.tr, G.base + (c << s), v[c]);
  }
}

template <int R, bool TRFAST, class Ld, class Mid, class St>
__device__ __forceinline__ void wht_round_fwd_mid_inv(int s, int l, int lntr, Ld ld, Mid mid, St st) {
  constexpr int RAD = 1 << R;
  const int total = 1 << (lntr + l - R);
  for (int g = threadIdx.x; g < total; g += blockDim.x) {
    const GroupIdx G = group_of<R, TRFAST>(g, s, l, lntr);
    double v[RAD];
#pragma unroll
    for (int c = 0; c < RAD; ++c) v[c] = ld(G.tr, G.base + (c << s));
    wht_butterflies<R>(v);
#pragma unroll
    for (int c = 0; c < RAD; ++c) v[c] = mid(G.tr, G.base + (c << s), v[c]);
    wht_butterflies<R>(v);
#pragma unroll
    for (int c = 0; c < RAD; ++c) st(G.tr, G.base + (c << s), v[c]);
  }
}

struct SmemR {
  double* sm;
  int LP;
  __device__ __forceinline__ double operator()(int tr, int idx) const { return sm[tr * LP + padidx(idx)]; }
  __device__ __forceinline__ void operator()(int tr, int idx, double v) const { sm[tr * LP + padidx(idx)] = v; }
};

// Round schedules for the FWHT (stage order is free): s[k], r[k] = first stage and number of stages of round k.
struct WhtSched {
  int n;
  int s[4], r[4];
};
// bottom-up: [0,4),[4,8),...,remainder on top
__device__ __forceinline__ WhtSched wht_sched_up(int l) {
  WhtSched q;
  q.n = 0;
  for (int s = 0; s < l; s += 4) {
    q.s[q.n] = s;
    q.r[q.n] = l - s >= 4 ? 4 : l - s;
    ++q.n;
  }
  return q;
}
// coalesced at both ends for contiguous tiles: the top four stages first (a thread's elements are 2^(l-4) apart, so
// consecutive threads read consecutive addresses), then the bottom remainder, then upwards, ending just below the top.
__device__ __forceinline__ WhtSched wht_sched_coalesced(int l) {
  WhtSched q;
  if (l <= 8) {
    q = wht_sched_up(l);
    if (q.n == 2) {  // top first, bottom last
      const int s1 = q.s[1], r1 = q.r[1];
      q.s[1] = q.s[0];
      q.r[1] = q.r[0];
      q.s[0] = s1;
      q.r[0] = r1;
    }
    return q;
  }
  const int rem = l & 3;
  q.n = 0;
  q.s[q.n] = l - 4, q.r[q.n] = 4, ++q.n;
  if (rem) q.s[q.n] = 0, q.r[q.n] = rem, ++q.n;
  for (int s = rem; s < l - 4; s += 4) q.s[q.n] = s, q.r[q.n] = 4, ++q.n;
  return q;
}

template <bool TRFAST, class Ld, class St>
__device__ __forceinline__ void wht_round_dispatch(int r, int s, int l, int lntr, Ld ld, St st) {
  FGP_R_DISPATCH(r, (wht_round_io<4, TRFAST>(s, l, lntr, ld, st)), (wht_round_io<3, TRFAST>(s, l, lntr, ld, st)),
                 (wht_round_io<2, TRFAST>(s, l, lntr, ld, st)), (wht_round_io<1, TRFAST>(s, l, lntr, ld, st)))
}

// Block FWHT following schedule q; gld feeds the first round, gst consumes the last.  No trailing sync.
template <bool TRFAST, class GLd, class GSt>
__device__ __forceinline__ void block_wht_io(double* sm, int l, int lntr, int LP, const WhtSched& q, GLd gld, GSt gst) {
  const SmemR S{sm, LP};
  if (q.n <= 1) {
    wht_round_dispatch<TRFAST>(l, 0, l, lntr, gld, gst);
    return;
  }
  wht_round_dispatch<TRFAST>(q.r[0], q.s[0], l, lntr, gld, S);
  __syncthreads();
  for (int k = 1; k < q.n - 1; ++k) {
    wht_round_dispatch<false>(q.r[k], q.s[k], l, lntr, S, S);
    __syncthreads();
  }
  wht_round_dispatch<TRFAST>(q.r[q.n - 1], q.s[q.n - 1], l, lntr, S, gst);
}

// forward FWHT, elementwise map, FWHT again (its own inverse) of the same tile; the last forward round, the map and the
// first backward round share registers.  Bottom-up then top-down.
template <bool TRFAST, class GLd, class Mid, class GSt>
__device__ __forceinline__ void block_wht_fwd_mid_inv_io(double* sm, int l, int lntr, int LP, GLd gld, Mid mid, GSt gst) {
  const SmemR S{sm, LP};
  if (l <= 4) {
    FGP_R_DISPATCH(l, (wht_round_fwd_mid_inv<4, TRFAST>(0, l, lntr, gld, mid, gst)), (wht_round_fwd_mid_inv<3, TRFAST>(0, l, lntr, gld, mid, gst)),
                   (wht_round_fwd_mid_inv<2, TRFAST>(0, l, lntr, gld, mid, gst)), (wht_round_fwd_mid_inv<1, TRFAST>(0, l, lntr, gld, mid, gst)))
    return;
  }
  wht_round_io<4, TRFAST>(0, l, lntr, gld, S);
  __syncthreads();
  int s = 4;
  while (l - s > 4) {
    wht_round_io<4, false>(s, l, lntr, S, S);
    s += 4;
    __syncthreads();
  }
  const int r = l - s;
  FGP_R_DISPATCH(r, (wht_round_fwd_mid_inv<4, TRFAST>(s, l, lntr, S, mid, S)), (wht_round_fwd_mid_inv<3, TRFAST>(s, l, lntr, S, mid, S)),
                 (wht_round_fwd_mid_inv<2, TRFAST>(s, l, lntr, S, mid, S)), (wht_round_fwd_mid_inv<1, TRFAST>(s, l, lntr, S, mid, S)))
  __syncthreads();
  s -= 4;
  while (s > 0) {
    wht_round_io<4, false>(s, l, lntr, S, S);
    s -= 4;
    __syncthreads();
  }
  wht_round_io<4, TRFAST>(0, l, lntr, S, gst);
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

// Tile capacities (log2 elements per CTA): 64 KiB of shared memory at most, 32 KiB preferred so that several CTAs in
// different phases share an SM.  Pass A keeps the contiguous block as long as possible so that pass B's strided
// transforms are short and its column tiles wide (128-byte global segments at n = 2^20).
static inline PassGeom make_geom(int64_t n, bool cplx) {
  PassGeom g;
  const size_t elem = cplx ? sizeof(double2) : sizeof(double);
  const int cap = cplx ? 12 : 13;
  g.m = ilog2(n);
  int tileA, tileB = 0;
  if (g.m <= cap) {
    g.l1 = g.m;
    g.l2 = 0;
    tileA = g.m > cap - 1 ? g.m : cap - 1;
  } else {
    g.l1 = g.m - cap > cap ? g.m - cap : cap;  // l2 <= cap always (n <= 2^(2 cap))
    if (g.l1 > cap) g.l1 = cap;
    g.l2 = g.m - g.l1;
    tileA = g.l1;
    tileB = g.l2 + g.l1 < cap - 1 ? g.l2 + g.l1 : cap - 1;
    if (tileB < g.l2) tileB = g.l2;
  }
  g.lntrA = tileA - g.l1;
  g.ntrA = 1 << g.lntrA;
  g.lntrB = g.l2 ? tileB - g.l2 : 0;
  g.ntrB = 1 << g.lntrB;
  g.LPA = padlen(1 << g.l1, g.ntrA);
  g.LPB = padlen(1 << g.l2, g.ntrB);
  auto thr = [](int tile) {
    int t = (1 << tile) / 16;
    if (t < 32) t = 32;
    if (t > 512) t = 512;
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
