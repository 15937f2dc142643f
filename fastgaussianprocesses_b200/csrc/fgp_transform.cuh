// K3 building blocks: shared-memory block FFT-BRO (radix-16 register rounds) and block FWHT.
//
// A "block transform" is a length L = 2^l (l <= 12 complex / 13 real) transform that lives entirely in one CTA's
// shared memory.  Each round loads 2^R elements (R <= 4) of one butterfly group into registers, runs R radix-2
// stages on them and writes them back in place, so a 4096-point block needs 3 shared-memory round trips.
//   forward  = decimation in time:  bit-reversed-order input -> natural-order output, no permutation pass
//   inverse  = decimation in frequency with conjugate twiddles: natural-order input -> bit-reversed-order output
// (these are exactly the conventions fixed by the reference's doubling recursion, fastgps/util.py:121-126.)
//
// Shared-memory layout: element e of transform tr sits at tr*LP + e + (e>>4).  The one-in-sixteen padding makes
// every radix-16 round (stride 1, 16, 256) conflict-free for 16-byte and 8-byte elements.
//
// Larger n = L1*L2 are done in two passes over global/L2-resident memory (four-step):
//   pass A: contiguous length-L1 blocks b,  then multiply element q1 by w_n^{rev(b) q1}
//   pass B: stride-L1 columns q1, length-L2 transforms, in place.
#pragma once
#include "fgp_common.cuh"

namespace fgp {

constexpr int kBlkLogC = 12;  // complex block transform: up to 4096 points (64 KiB + padding)
constexpr int kBlkLogR = 13;  // real block transform:    up to 8192 points
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

// ---------------------------------------------------------------------------------------------------------------
// complex rounds
// ---------------------------------------------------------------------------------------------------------------
template <int R, bool INV>
__device__ __forceinline__ void fft_round(double2* __restrict__ sm, int s, int l, int ntr, int LP,
                                          const double2* __restrict__ tw) {
  constexpr int RAD = 1 << R;
  const int total = ntr << (l - R);
  const int gmask = (1 << (l - R)) - 1;
  const int lmask = (1 << s) - 1;
  for (int g = threadIdx.x; g < total; g += blockDim.x) {
    const int tr = g >> (l - R);
    const int j = g & gmask;
    const int low = j & lmask;
    const int base = low + ((j >> s) << (s + R));
    double2* p = sm + tr * LP;
    double2 v[RAD];
#pragma unroll
    for (int c = 0; c < RAD; ++c) v[c] = p[padidx(base + (c << s))];
    if (!INV) {
#pragma unroll
      for (int u = 0; u < R; ++u) {
#pragma unroll
        for (int c = 0; c < RAD; ++c) {
          if (c & (1 << u)) continue;
          const int cm = c & ((1 << u) - 1);
          const double2 w = __ldg(tw + (1 << (s + u)) + low + (cm << s));
          const double2 b = cmul(w, v[c | (1 << u)]);
          v[c | (1 << u)] = csub(v[c], b);
          v[c] = cadd(v[c], b);
        }
      }
    } else {
#pragma unroll
      for (int u = R - 1; u >= 0; --u) {
#pragma unroll
        for (int c = 0; c < RAD; ++c) {
          if (c & (1 << u)) continue;
          const int cm = c & ((1 << u) - 1);
          const double2 w = __ldg(tw + (1 << (s + u)) + low + (cm << s));
          const double2 a = v[c], b = v[c | (1 << u)];
          v[c] = cadd(a, b);
          v[c | (1 << u)] = cmulc(w, csub(a, b));
        }
      }
    }
#pragma unroll
    for (int c = 0; c < RAD; ++c) p[padidx(base + (c << s))] = v[c];
  }
}

template <bool INV>
__device__ __forceinline__ void fft_round_dispatch(int r, double2* sm, int s, int l, int ntr, int LP,
                                                   const double2* tw) {
  switch (r) {
    case 4: fft_round<4, INV>(sm, s, l, ntr, LP, tw); break;
    case 3: fft_round<3, INV>(sm, s, l, ntr, LP, tw); break;
    case 2: fft_round<2, INV>(sm, s, l, ntr, LP, tw); break;
    default: fft_round<1, INV>(sm, s, l, ntr, LP, tw); break;
  }
}

// ntr transforms of length 2^l in shared memory; caller has synchronised after filling sm; returns synchronised.
__device__ __forceinline__ void block_fft_fwd(double2* sm, int l, int ntr, int LP, const double2* tw) {
  int s = 0;
  while (s < l) {
    const int r = (l - s) >= 4 ? 4 : (l - s);
    fft_round_dispatch<false>(r, sm, s, l, ntr, LP, tw);
    s += r;
    __syncthreads();
  }
}
__device__ __forceinline__ void block_fft_inv(double2* sm, int l, int ntr, int LP, const double2* tw) {
  // mirror image of the forward schedule so that both use the same (conflict-free) round boundaries
  int top = l;
  const int rem = l & 3;
  // forward rounds are [0,4),[4,8),...,[l-rem,l); run them last-to-first
  if (rem) {
    fft_round_dispatch<true>(rem, sm, l - rem, l, ntr, LP, tw);
    top = l - rem;
    __syncthreads();
  }
  while (top > 0) {
    fft_round_dispatch<true>(4, sm, top - 4, l, ntr, LP, tw);
    top -= 4;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// real Walsh-Hadamard rounds (no twiddles; self-inverse; stage order is irrelevant)
// ---------------------------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void wht_round(double* __restrict__ sm, int s, int l, int ntr, int LP) {
  constexpr int RAD = 1 << R;
  const int total = ntr << (l - R);
  const int gmask = (1 << (l - R)) - 1;
  const int lmask = (1 << s) - 1;
  for (int g = threadIdx.x; g < total; g += blockDim.x) {
    const int tr = g >> (l - R);
    const int j = g & gmask;
    const int base = (j & lmask) + ((j >> s) << (s + R));
    double* p = sm + tr * LP;
    double v[RAD];
#pragma unroll
    for (int c = 0; c < RAD; ++c) v[c] = p[padidx(base + (c << s))];
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
#pragma unroll
    for (int c = 0; c < RAD; ++c) p[padidx(base + (c << s))] = v[c];
  }
}

__device__ __forceinline__ void block_wht(double* sm, int l, int ntr, int LP) {
  int s = 0;
  while (s < l) {
    const int r = (l - s) >= 4 ? 4 : (l - s);
    switch (r) {
      case 4: wht_round<4>(sm, s, l, ntr, LP); break;
      case 3: wht_round<3>(sm, s, l, ntr, LP); break;
      case 2: wht_round<2>(sm, s, l, ntr, LP); break;
      default: wht_round<1>(sm, s, l, ntr, LP); break;
    }
    s += r;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// two-pass geometry
// ---------------------------------------------------------------------------------------------------------------
struct PassGeom {
  int m;    // log2 n
  int l1;   // log2 of the contiguous block length (pass A)
  int l2;   // log2 of the strided transform length (pass B); 0 => single pass
  int ntrA; // transforms per CTA in pass A
  int ntrB; // columns per CTA in pass B
  int LPA, LPB;
  int threads;  // CTA size: one radix-16 butterfly group per thread per round
  int64_t ctasA, ctasB;  // per batch item
  size_t smemA, smemB;
};
// Tile capacity (elements per CTA) is chosen by problem size: small tiles (128 threads, 4 CTAs/SM) keep several CTAs in
// different phases (global load / butterflies / store) resident per SM; the largest sizes need the full-size tile so
// that two passes suffice.  l1 is taken as large as the tile allows so that pass B's column tiles are as wide as possible.
static inline int tile_log(int m, bool cplx) {
  if (cplx) return m <= 22 ? 11 : 12;
  return m <= 22 ? 12 : (m == 23 ? 13 : 14);  // >= 64-byte column segments once the data no longer fits L2
}
static inline PassGeom make_geom(int64_t n, bool cplx, int max_threads = 256) {
  PassGeom g;
  const size_t elem = cplx ? sizeof(double2) : sizeof(double);
  g.m = ilog2(n);
  const int blklog = tile_log(g.m, cplx);
  g.l1 = g.m <= blklog ? g.m : blklog;
  g.l2 = g.m - g.l1;
  const int cap = 1 << blklog;
  g.threads = cap / 16 < max_threads ? cap / 16 : max_threads;
  if (g.threads < 32) g.threads = 32;
  g.ntrA = cap >> g.l1;
  if (g.ntrA < 1) g.ntrA = 1;
  g.ntrB = g.l2 ? (cap >> g.l2) : 1;
  g.LPA = padlen(1 << g.l1, g.ntrA);
  g.LPB = padlen(1 << g.l2, g.ntrB);
  g.ctasA = (n >> g.l1) / g.ntrA;
  if (g.ctasA < 1) g.ctasA = 1;
  g.ctasB = g.l2 ? ((int64_t(1) << g.l1) / g.ntrB) : 0;
  g.smemA = (size_t)g.ntrA * g.LPA * elem;
  g.smemB = (size_t)g.ntrB * g.LPB * elem;
  return g;
}

__device__ __forceinline__ uint32_t brev_bits(uint32_t v, int bits) { return bits ? (__brev(v) >> (32 - bits)) : 0u; }

}  // namespace fgp
