// K5b: fused posterior variance for a rank-1 lattice in generator form (abstract_gp.py:381-416, guard abstract_fast_gp.py:41-46).
//
//   pvar(x*) = k(x*,x*) - sum_k |ft(k(x*, X))_k|^2 / lam_k
//
// The unfused route (fgp_posterior.cu) runs four kernels per chunk of test points: cross-pair kernel -> c2c pass A -> pass B ->
// pair reduction, ~96 n bytes moved per pair of test points and the (n,d) training points read for every group of pairs
// (profiles/README.md, round 1: 43 us per pair at n = 2^20, d = 8).  Here a pair of test points costs TWO kernels and 32 n bytes:
//   pv_passA  evaluates z_i = k(x*_a, X_i) + i k(x*_b, X_i) for the 2^l1 points of one contiguous block straight into the
//             shared-memory tile of the block transform -- the training points are REGENERATED from the point index,
//             x_i - shift = frac(phi2(i) z) (one 32-bit IMAD per coordinate, as in the fit kernels), so nothing but the two test
//             points is read -- then runs the block transform and the inter-pass twiddle and writes the workspace once;
//   pv_passB  runs the column transforms of 2c columns chosen as c MIRROR PAIRS (q, L1 - q), so that both members of every
//             spectral pair (k, n - k) sit in the same shared-memory tile, and reduces
//                 |A_k|^2 = |Z_k + conj Z_{n-k}|^2 / 4,  |B_k|^2 = |Z_k - conj Z_{n-k}|^2 / 4   (the spectra of the two real sequences)
//             against Re(1/lam_k) right there: the spectrum is never written.
// Per-tile partial sums are added in a fixed order by a tiny final kernel: results do not depend on scheduling.
#include "fgp_transform.cuh"
#include "fgp_tma.cuh"

namespace fgp {

struct PvArgs {
  const double* xs;  // (m,d) test points of this chunk
  int64_t m;
  int64_t n;
  int d;
  UVec z;       // generating vector
  DVec shift;   // random shift of the lattice
  int a2;       // every alpha_j == 2
  // a2:  f_j / A_j = 1 - h^2,  h = |x' - t'| (sig_j - |x' - t'|) with coordinates pre-scaled by sig_j (post_mean's 5-slot form)
  DVec sig, sig32;  // sig_j and sig_j 2^-32
  double pref;  // scale * prod_j A_j  (a2) or scale
  // general alpha: f_j(u) = sum_p c[j][p] u^p, u = a(1-a)
  double c[FGP_MAX_D][FGP_MAX_ALPHA + 1];
  int alpha[FGP_MAX_D];
  double2* W;          // (pairs, n) workspace
  const double* winv;  // (n) Re(1/lam_k) / n   (the transforms below are unnormalised)
  double* partial;     // (pairs, tilesB, 2)
  int l1, l2, LPA, LPB, lcB, tilesB;
  FftTables T;
};

__global__ void __launch_bounds__(256) pv_winv_kernel(const double2* __restrict__ lam, int64_t n, double* __restrict__ winv) {
  const double inv_n = 1.0 / (double)n;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const double2 l = lam[k];
    winv[k] = l.x / fma(l.x, l.x, l.y * l.y) * inv_n;
  }
}

// Workspace layout of one pair: pass-B tile t holds the c columns q0 .. q0+c-1 (q0 = t c) in slots 0 .. c-1 and their mirrors L1 - q in slots
// c .. 2c-1 (the self-mirrored column L1/2 takes the slot of q = 0's mirror), all L2 rows: element (row, slot) of tile t at ((t L2 + row) 2c + slot).
__device__ __forceinline__ int64_t pv_woff(int q, int row, int lcB, int l1, int l2) {
  const int c = 1 << lcB, L1 = 1 << l1, half1 = L1 >> 1;
  int tile, slot;
  if (q < half1) {
    tile = q >> lcB;
    slot = q & (c - 1);
  } else if (q == half1) {
    tile = 0;
    slot = c;
  } else {
    const int qm = L1 - q;
    tile = qm >> lcB;
    slot = c + (qm & (c - 1));
  }
  return ((int64_t)tile << (l2 + lcB + 1)) + ((int64_t)row << (lcB + 1)) + slot;
}

template <int DT, bool A2>
__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) pv_passA_kernel(const __grid_constant__ PvArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double xq[2][FGP_MAX_D];  // frac(x* - shift) of the two test points (times sig_j on the a2 path)
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  const int d = DT > 0 ? DT : a.d;
  double2* sm = (double2*)smraw;
  const int64_t p = blockIdx.y;
  const int blk = blockIdx.x;
  const int l1 = a.l1, LP = a.LPA;
  if (threadIdx.x < 2 * d) {
    const int q = threadIdx.x / d, j = threadIdx.x - q * d;
    int64_t i = 2 * p + q;
    if (i >= a.m) i = a.m - 1;  // odd m: the last pair repeats the last point (its second variance is dropped)
    double v = a.xs[i * d + j] - a.shift.v[j];
    v -= floor(v);
    xq[q][j] = A2 ? v * a.sig.v[j] : v;
  }
  __syncthreads();
  const int64_t g0 = (int64_t)blk << l1;
  tile_fill_c<false>(SmemC{sm, LP}, l1, 0, [&](int, int idx) -> double2 {
    // x_i - shift = frac(phi2(i) z_j) exactly: the top word of the wrap-around product brev32(i) * z_j (fgp_mll.cuh, generator mode)
    const uint32_t rev = __brev((uint32_t)(g0 + idx));
    double ka = 1.0, kb = 1.0;
#pragma unroll
    for (int j = 0; j < DM; ++j) {
      if (j >= d) break;
      const double ti = (double)(rev * (uint32_t)a.z.v[j]);  // 2^32 frac(phi2(i) z_j), exact
      const double t = ti * 0x1.0p-32;
      if (A2) {
        const double ts = ti * a.sig32.v[j];  // sig_j 2^-32 folded into one constant
        const double da = fabs(xq[0][j] - ts), db = fabs(xq[1][j] - ts);
        const double ha = da * (a.sig.v[j] - da), hb = db * (a.sig.v[j] - db);
        const double fa = fma(-ha, ha, 1.0), fb = fma(-hb, hb, 1.0);
        ka = j == 0 ? fa : ka * fa;
        kb = j == 0 ? fb : kb * fb;
      } else {
        const double da = fabs(xq[0][j] - t), db = fabs(xq[1][j] - t);  // B_2a(frac delta) = B_2a(|delta|), |delta| < 1
        const double ua = da * (1.0 - da), ub = db * (1.0 - db);
        const int al = a.alpha[j];
        double fa = a.c[j][al], fb = fa;
        for (int q = al - 1; q >= 0; --q) {
          fa = fma(fa, ua, a.c[j][q]);
          fb = fma(fb, ub, a.c[j][q]);
        }
        ka *= fa;
        kb *= fb;
      }
    }
    return make_double2(ka * a.pref, kb * a.pref);
  });
  __syncthreads();
  // scatter on store, never on load: the workspace of a pair is laid out by PASS-B TILE, [tile][row][2c slots] (pv_woff), so that pass B
  // reads each of its tiles as one contiguous block; pass A pays with 16c-byte store runs, which nobody waits for
  double2* W = a.W + p * a.n;
  const FftTables T = a.T;
  const uint32_t rb = brev_bits((uint32_t)blk, a.l2);
  const int lcB = a.lcB, l2 = a.l2;
  block_fft_fwd_io<false>(sm, l1, 0, LP, T.stage, SmemTag{},
                          [&](int, int idx, double2 v) { W[pv_woff(idx, blk, lcB, l1, l2)] = cmul(v, twiddle_n(T, rb * (uint32_t)idx)); });
}

// reduction of one transformed pass-B tile (in shared memory) against Re(1/lam): both members of every pair (k, n-k) are in the tile
__device__ __forceinline__ void pv_tile_reduce(const PvArgs& a, double2* sm, int64_t p, int tile, double* red) {
  const int l1 = a.l1, l2 = a.l2, LP = a.LPB, lcB = a.lcB;
  const int c = 1 << lcB, L1 = 1 << l1, L2 = 1 << l2, half1 = L1 >> 1;
  const int q0 = tile << lcB;
  const SmemC S{sm, LP};
  const double* winv = a.winv;
  const int64_t n = a.n;
  double s[2] = {0.0, 0.0};
  auto acc = [&](double2 zk, double2 zm, bool self, int64_t k) {
    const double w = 0.25 * (self ? winv[k] : winv[k] + winv[n - k]);
    const double ar = zk.x + zm.x, ai = zk.y - zm.y;  // Z_k + conj Z_{n-k}
    const double br = zk.x - zm.x, bi = zk.y + zm.y;  // Z_k - conj Z_{n-k}
    s[0] = fma(fma(ar, ar, ai * ai), w, s[0]);
    s[1] = fma(fma(br, br, bi * bi), w, s[1]);
  };
  for (int e = threadIdx.x; e < (c << l2); e += blockDim.x) {
    const int t = e & (c - 1), r = e >> lcB;
    const int q = q0 + t;
    if (q == 0) {
      const int rp = (L2 - r) & (L2 - 1);  // n - k = (L2 - r) L1
      if (r <= rp) acc(S(0, r), S(0, rp), r == rp, (int64_t)r << l1);
      const int rm = L2 - 1 - r;           // column L1/2: n - k = (L2 - 1 - r) L1 + L1/2
      if (r < rm) acc(S(c, r), S(c, rm), false, ((int64_t)r << l1) + half1);
    } else {
      acc(S(t, r), S(c + t, L2 - 1 - r), false, ((int64_t)r << l1) + q);  // n - k = (L2 - 1 - r) L1 + (L1 - q)
    }
  }
  block_sum<2>(s, red);
  if (threadIdx.x == 0) {
    double* dst = a.partial + (p * a.tilesB + tile) * 2;
    dst[0] = s[0];
    dst[1] = s[1];
  }
}

__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) pv_passB_kernel(const __grid_constant__ PvArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red[32 * 4];
  double2* sm = (double2*)smraw;
  const int64_t p = blockIdx.y;
  const int tile = blockIdx.x;
  const int lc2 = a.lcB + 1;
  const double2* Wt = a.W + p * a.n + ((int64_t)tile << (a.l2 + lc2));  // the tile is one contiguous block (pv_woff)
  block_fft_fwd_io<true>(sm, a.l2, lc2, a.LPB, a.T.stage, [&](int tr, int r) -> double2 { return Wt[(r << lc2) + tr]; }, SmemTag{});
  __syncthreads();
  pv_tile_reduce(a, sm, p, tile, red);
}

// The same pass as a PERSISTENT kernel with TMA-staged, double-buffered tiles: a CTA walks tiles w, w + grid, ...; while it transforms
// and reduces tile w out of its padded working tile, the TMA engine (bulk-asynchronous copies of the contiguous tile, SASS UBLKCP, completion
// on an mbarrier) already fills the other dense staging tile with tile w + grid: no registers and no warp slots are held by loads in flight.
__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) pv_passB_tma_kernel(const __grid_constant__ PvArgs a, int npairs, int stage_elems) {
  extern __shared__ __align__(128) unsigned char smraw_pv[];
  __shared__ double red[32 * 4];
  __shared__ __align__(8) uint64_t bars[2];
  double2* stg0 = (double2*)smraw_pv;
  double2* stg1 = stg0 + stage_elems;
  double2* sm = stg1 + stage_elems;
  const int l2 = a.l2, lc2 = a.lcB + 1;
  const int total = npairs * a.tilesB;
  const unsigned tile_bytes = (unsigned)(sizeof(double2) << (l2 + lc2));
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int w, int buf) {  // warp 0: the tile as 32 bulk copies
    const int p = w / a.tilesB, tile = w - p * a.tilesB;
    const char* src = (const char*)(a.W + (int64_t)p * a.n + ((int64_t)tile << (l2 + lc2)));
    char* dst = (char*)(buf ? stg1 : stg0);
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");
      mbar_arrive_expect_tx(&bars[buf], tile_bytes);
    }
    __syncwarp();
    const unsigned chunk = tile_bytes / 32;
    tma_load_1d(dst + (size_t)lane * chunk, src + (size_t)lane * chunk, chunk, &bars[buf]);
  };
  unsigned parity[2] = {0u, 0u};
  int buf = 0;
  int w = blockIdx.x;
  if (w < total && threadIdx.x < 32) issue(w, 0);
  for (; w < total; w += gridDim.x) {
    if (w + (int)gridDim.x < total && threadIdx.x < 32) issue(w + gridDim.x, buf ^ 1);
    const int p = w / a.tilesB, tile = w - p * a.tilesB;
    mbar_wait(&bars[buf], parity[buf]);
    parity[buf] ^= 1u;
    const double2* stg = buf ? stg1 : stg0;
    block_fft_fwd_io<true>(sm, l2, lc2, a.LPB, a.T.stage, [&](int tr, int r) -> double2 { return stg[(r << lc2) + tr]; }, SmemTag{});
    __syncthreads();
    pv_tile_reduce(a, sm, p, tile, red);
    __syncthreads();  // the working tile and this staging tile are free again
    buf ^= 1;
  }
}

// one WARP per pair: lane-strided partial sums, then a shuffle tree -- a fixed order, so the result does not depend on scheduling
// (one thread per pair made this kernel 30 us of serial loads per chunk, 7 % of the whole post_var)
__global__ void __launch_bounds__(256) pv_final_kernel(const double* __restrict__ partial, int tiles, int64_t pairs, int64_t m, double kxx, double* __restrict__ out) {
  const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= pairs) return;
  double s0 = 0.0, s1 = 0.0;
  for (int g = lane; g < tiles; g += 32) {
    s0 += partial[(p * tiles + g) * 2 + 0];
    s1 += partial[(p * tiles + g) * 2 + 1];
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (lane == 0) {
    const double v0 = kxx - s0, v1 = kxx - s1;
    out[2 * p] = v0 < 0.0 ? 0.0 : v0;
    if (2 * p + 1 < m) out[2 * p + 1] = v1 < 0.0 ? 0.0 : v1;
  }
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Geometry of the two passes: the complex two-pass split of the transforms (make_geom), with pass-B tiles as wide as 64 KiB of shared
// memory allow -- 2^(12 - l2) columns, half of them mirrors.  Measured at d = 8 on B200 (profiles/README.md, round 2): 16 instead of 8
// columns per tile at n = 2^20: 68.0 k -> 77.8 k points/s; n = 2^16: 690 k -> 947 k.
static PassGeom pv_geom(int64_t n) {
  PassGeom g = make_geom(n, true);
  if (!g.l2) return g;
  static const int cols_env = env_int("FGP_PV_COLS_LOG2", 0);
  int lb = cols_env > 0 ? cols_env : 12 - g.l2;
  if (env_int("FGP_PV_TMA", 0) != 0 && lb > 11 - g.l2) lb = 11 - g.l2;  // the staged variant needs room for two staging tiles: 32 KiB tiles
  if (lb > g.l1 - 1) lb = g.l1 - 1;
  if (lb < 1) lb = 1;
  g.lntrB = lb;
  g.ntrB = 1 << lb;
  g.LPB = padlen(1 << g.l2, g.ntrB, kPSC);
  int t = ((1 << (g.l2 + lb)) / 16) * 2;
  if (t < 32) t = 32;
  if (t > FGP_LB_THREADS) t = FGP_LB_THREADS;
  g.threadsB = t;
  g.smemB = (size_t)g.ntrB * g.LPB * sizeof(double2);
  return g;
}

static int64_t pv_chunk_pairs(int64_t pairs, int64_t n) {
  // pairs per chunk: enough tiles per launch for many waves (measured at n = 2^20, d = 8: 1 / 4 / 16 pairs per chunk -> 28.6 k / 52.8 k /
  // 65.6 k points per second, although 16 pairs = 256 MiB no longer fit the L2); FGP_PV_CHUNK overrides
  static const int env = env_int("FGP_PV_CHUNK", 0);
  int64_t c = env > 0 ? env : (int64_t(1) << 24) / n;
  if (c < 16) c = 16;
  if (c * n > (int64_t(1) << 26)) c = (int64_t(1) << 26) / n;  // at most 1 GiB of workspace
  if (c > 4096) c = 4096;
  if (c > pairs) c = pairs;
  return c;
}

template <int DT, bool A2>
static void launch_pvA(const PvArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  pv_passA_kernel<DT, A2><<<grid, threads, smem, st>>>(a);
}


// ------------------------------------------------------------------------------------------------------------------------------
// digital nets: the same two-kernel structure with the real Walsh-Hadamard transform (no pairing: one real sequence per test point)
//   pvn_passA  k(x*, X_i) for the 2^l1 points of a block, the points REGENERATED as xb_i = XOR_{k in bits(i)} C[j][k] ^ dshift_j from two
//              shared-memory XOR-fold tables per tile (as the fit kernels do) -> block FWHT -> workspace in pass-B tile order
//   pvn_passB  column FWHT of a contiguous tile -> sum_k v_k^2 / (n lam_k) in the epilogue
// ------------------------------------------------------------------------------------------------------------------------------
struct PvnArgs {
  const double* xs;
  int64_t m, n;
  int d, t, mmax;
  double tscale;
  const uint64_t* C;  // device (d, mmax) generating-matrix columns
  UVec dshift;
  DVec ls;
  IVec alpha;
  double scale;
  double* W;           // (points, n) in pass-B tile order
  const double* winv;  // (n) 1 / (n lam_k)
  double* partial;     // (points, tilesB)
  int l1, l2, LPA, LPB, lcB, tilesB, tab_off;
};

__global__ void __launch_bounds__(256) pvn_winv_kernel(const double* __restrict__ lam, int64_t n, double* __restrict__ winv) {
  const double inv_n = 1.0 / (double)n;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) winv[k] = inv_n / lam[k];
}

__device__ __forceinline__ uint64_t pvn_fold(const uint64_t* __restrict__ Cj, uint64_t v) {
  uint64_t r = 0;
  while (v) {
    const int k = __ffsll((long long)v) - 1;
    r ^= __ldg(Cj + k);
    v &= v - 1;
  }
  return r;
}

constexpr int kPvnPoints = 4;
// MODE 0: any alpha; 1: every alpha == 2; 2: alpha == 2 and t <= 52 (exact conversion through the 2^52 magic number)
template <int DT, int MODE>
__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) pvn_passA_kernel(const __grid_constant__ PvnArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ uint64_t cst[FGP_MAX_D];  // to_b(x*_j) ^ dshift_j
  constexpr int DM = DT > 0 ? DT : FGP_MAX_D;
  const int d = DT > 0 ? DT : a.d;
  double* sm = (double*)smraw;
  uint64_t* TA = (uint64_t*)(smraw + a.tab_off);  // TA[j*64 + v]: fold over the low 6 index bits; TB[j*nb + h]: over the rest of the tile index
  const int l1 = a.l1, LP = a.LPA;
  const int nb = l1 > 6 ? 1 << (l1 - 6) : 1;
  uint64_t* TB = TA + 64 * d;
  const int blk = blockIdx.x;
  const int64_t g0 = (int64_t)blk << l1;
  for (int e = threadIdx.x; e < 64 * d; e += blockDim.x) TA[e] = pvn_fold(a.C + (int64_t)(e >> 6) * a.mmax, (uint64_t)(e & 63));
  for (int e = threadIdx.x; e < nb * d; e += blockDim.x) {
    const int j = e / nb, h = e - j * nb;
    TB[e] = pvn_fold(a.C + (int64_t)j * a.mmax, (uint64_t)g0 | ((uint64_t)h << 6));
  }
  const int lcB = a.lcB, l2 = a.l2;
  // kPvnPoints test points per CTA, one after the other: the fold tables of the block are built once for all of them
  for (int q = 0; q < kPvnPoints; ++q) {
    const int64_t p = (int64_t)blockIdx.y * kPvnPoints + q;
    if (p >= a.m) break;  // uniform over the CTA
    __syncthreads();      // the previous point's tile and cst are no longer read
    if (threadIdx.x < d) cst[threadIdx.x] = dnb2_to_b(a.xs[p * d + threadIdx.x], a.t) ^ a.dshift.v[threadIdx.x];
    __syncthreads();
    tile_fill_r<false>(SmemR{sm, LP}, l1, 0, [&](int, int idx) -> double {
      const uint64_t* ta = TA + (idx & 63);
      const uint64_t* tb = TB + (idx >> 6);
      double k = a.scale;
#pragma unroll
      for (int j = 0; j < DM; ++j) {
        if (j >= d) break;
        const uint64_t delta = ta[j * 64] ^ tb[j * nb] ^ cst[j];
        const double part = MODE == 2 ? dnb2_part_a2_t52(delta, a.t, a.tscale) : (MODE == 1 ? dnb2_part_a2(delta, a.t, a.tscale) : dnb2_part(delta, a.alpha.v[j], a.t));
        k *= fma(a.ls.v[j], part, 1.0);
      }
      return k;
    });
    __syncthreads();
    // workspace in pass-B tile order: tile = q >> lcB holds 2^lcB adjacent columns of all L2 rows, element (row, slot) at ((tile L2 + row) << lcB) + slot
    double* W = a.W + p * a.n;
    block_wht_io<false>(sm, l1, 0, LP, wht_sched_up(l1), SmemTag{}, [&](int, int idx, double v) {
      W[((int64_t)(idx >> lcB) << (l2 + lcB)) + ((int64_t)blk << lcB) + (idx & ((1 << lcB) - 1))] = v;
    });
  }
}

__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS_R) pvn_passB_kernel(const __grid_constant__ PvnArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red[32 * 4];
  double* sm = (double*)smraw;
  const int64_t p = blockIdx.y;
  const int tile = blockIdx.x;
  const int l1 = a.l1, l2 = a.l2, lcB = a.lcB;
  const double* Wt = a.W + p * a.n + ((int64_t)tile << (l2 + lcB));
  block_wht_io<true>(sm, l2, lcB, a.LPB, wht_sched_up(l2), [&](int tr, int r) -> double { return Wt[(r << lcB) + tr]; }, SmemTag{});
  __syncthreads();
  const double* winv = a.winv + ((int64_t)tile << lcB);
  double s[1] = {0.0};
  tile_drain_r<true>(SmemR{sm, a.LPB}, l2, lcB, [&](int tr, int r, double v) { s[0] = fma(v * v, winv[((int64_t)r << l1) + tr], s[0]); });
  block_sum<1>(s, red);
  if (threadIdx.x == 0) a.partial[p * a.tilesB + tile] = s[0];
}

__global__ void __launch_bounds__(256) pvn_final_kernel(const double* __restrict__ partial, int tiles, int64_t cnt, double kxx, double* __restrict__ out) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per test point, fixed summation order
  const int lane = threadIdx.x & 31;
  if (i >= cnt) return;
  double s = 0.0;
  for (int g = lane; g < tiles; g += 32) s += partial[i * tiles + g];
  s = warp_sum(s);
  if (lane == 0) {
    const double v = kxx - s;
    out[i] = v < 0.0 ? 0.0 : v;
  }
}

static PassGeom pvn_geom(int64_t n) {
  PassGeom g = make_geom(n, false);
  if (!g.l2) return g;
  int lb = 12 - g.l2;  // 32 KiB pass-B tiles
  if (lb > g.l1) lb = g.l1;
  if (lb < 0) lb = 0;
  g.lntrB = lb;
  g.ntrB = 1 << lb;
  g.LPB = padlen(1 << g.l2, g.ntrB, kPSR);
  int t = (1 << (g.l2 + lb)) / 16;
  if (t < 32) t = 32;
  if (t > FGP_LB_THREADS) t = FGP_LB_THREADS;
  g.threadsB = t;
  g.smemB = (size_t)g.ntrB * g.LPB * sizeof(double);
  return g;
}

static int64_t pvn_chunk_points(int64_t m, int64_t n) {
  static const int env = env_int("FGP_PV_CHUNK", 0);
  int64_t c = env > 0 ? env : (int64_t(1) << 25) / n;
  if (c < 32) c = 32;
  if (c * n > (int64_t(1) << 27)) c = (int64_t(1) << 27) / n;  // at most 1 GiB of workspace
  if (c < 1) c = 1;
  if (c > m) c = m;
  return c;
}

template <int DT, int MODE>
static void launch_pvnA(const PvnArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pvn_passA_kernel<DT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaGetLastError();
    attr = true;
  }
  pvn_passA_kernel<DT, MODE><<<grid, threads, smem, st>>>(a);
}
template <int MODE>
static void dispatch_pvnA(const PvnArgs& a, dim3 grid, int threads, size_t smem, cudaStream_t st) {
  switch (a.d) {
    case 2: launch_pvnA<2, MODE>(a, grid, threads, smem, st); break;
    case 4: launch_pvnA<4, MODE>(a, grid, threads, smem, st); break;
    case 8: launch_pvnA<8, MODE>(a, grid, threads, smem, st); break;
    case 16: launch_pvnA<16, MODE>(a, grid, threads, smem, st); break;
    default: launch_pvnA<0, MODE>(a, grid, threads, smem, st); break;
  }
}

}  // namespace fgp

extern "C" {

size_t fgp_lattice_post_var_z_workspace_bytes(int64_t m, int64_t n) {
  using namespace fgp;
  if (m <= 0 || !is_pow2(n)) return 0;
  const PassGeom g = pv_geom(n);
  if (!g.l2) return 0;
  const int64_t pairs = (m + 1) >> 1, cp = pv_chunk_pairs(pairs, n);
  const int64_t tilesB = (int64_t(1) << g.l1) >> g.lntrB;
  return align256((size_t)cp * n * sizeof(double2)) + align256((size_t)n * sizeof(double)) + align256((size_t)cp * tilesB * 2 * sizeof(double));
}

int fgp_lattice_post_var_z(const double* xs_dev, int64_t m, const uint64_t* z_host, const double* shift_host, int64_t n, int d, const int* alpha_host,
                           double scale, const double* ls_host, const double* lam_dev, const void* table_dev, void* work_dev, double* pvar_dev,
                           fgp_stream_t stream) {
  using namespace fgp;
  FGP_REQUIRE(xs_dev && z_host && shift_host && alpha_host && ls_host && lam_dev && table_dev && work_dev && pvar_dev, "post_var_z: null pointer");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D && m >= 0 && is_pow2(n) && ilog2(n) <= FGP_MAX_LOG2N_FFT, "post_var_z: bad m/n/d (n must be a power of two <= 2^%d)", FGP_MAX_LOG2N_FFT);
  if (m == 0) return FGP_OK;
  const PassGeom g = pv_geom(n);
  FGP_REQUIRE(g.l2 >= 1 && g.lntrA == 0 && g.lntrB >= 1, "post_var_z: n=%lld is a single-tile size, use fgp_lattice_post_var", (long long)n);
  PvArgs a;
  memset(&a, 0, sizeof(a));
  LatPoly P;
  int rc = fill_lat_poly(alpha_host, d, &P);
  if (rc) return rc;
  double kxx = scale;
  bool all2 = true;
  a.pref = scale;
  for (int j = 0; j < d; ++j) {
    FGP_REQUIRE(z_host[j] < (uint64_t(1) << 32), "post_var_z: generating vector entries must be below 2^32");
    FGP_REQUIRE(shift_host[j] >= 0.0 && shift_host[j] < 1.0, "post_var_z: shift outside [0,1)");
    a.z.v[j] = z_host[j];
    a.shift.v[j] = shift_host[j];
    a.alpha[j] = alpha_host[j];
    all2 = all2 && alpha_host[j] == 2;
    for (int q = 0; q <= alpha_host[j]; ++q) a.c[j][q] = ls_host[j] * P.q[j][q];
    a.c[j][0] += 1.0;
    kxx *= 1.0 + ls_host[j] * P.q[j][0];  // delta = 0 in every dimension
  }
  if (all2) {
    for (int j = 0; j < d; ++j) {
      const double A = a.c[j][0], Bq = -a.c[j][2];
      if (!(A > 0.0 && Bq >= 0.0)) {
        all2 = false;  // cannot happen for alpha = 2 and positive lengthscales; the general path is always valid
        break;
      }
    }
  }
  if (all2) {
    for (int j = 0; j < d; ++j) {
      const double A = a.c[j][0], Bq = -a.c[j][2];
      a.sig.v[j] = sqrt(sqrt(Bq / A));
      a.sig32.v[j] = a.sig.v[j] * 0x1.0p-32;
      a.pref *= A;
    }
  }
  a.a2 = all2;
  a.n = n;
  a.d = d;
  a.l1 = g.l1;
  a.l2 = g.l2;
  a.LPA = g.LPA;
  a.LPB = g.LPB;
  a.lcB = g.lntrB - 1;
  a.tilesB = (int)((int64_t(1) << g.l1) >> g.lntrB);
  a.T = make_tables(table_dev);
  const int64_t pairs = (m + 1) >> 1, cp = pv_chunk_pairs(pairs, n);
  char* w = (char*)work_dev;
  a.W = (double2*)w;
  w += align256((size_t)cp * n * sizeof(double2));
  double* winv = (double*)w;
  a.winv = winv;
  w += align256((size_t)n * sizeof(double));
  a.partial = (double*)w;
  cudaStream_t st = (cudaStream_t)stream;
  {
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    pv_winv_kernel<<<(unsigned)blocks, 256, 0, st>>>((const double2*)lam_dev, n, winv);
    FGP_LAUNCH_NAMED("pv_winv", st);
  }
  static bool attr_done = false;
  if (!attr_done) {
    const int big = 200 * 1024;
    cudaFuncSetAttribute(pv_passA_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(pv_passA_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(pv_passA_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(pv_passA_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(pv_passA_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(pv_passA_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(pv_passB_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaGetLastError();
    attr_done = true;
  }
  // TMA-staged persistent pass B (two dense staging tiles + the padded working tile); FGP_PV_TMA=1 selects it
  // (read per call so that tests compare both routes in one process)
  const int stage_elems = (2 << a.lcB) << g.l2;
  const size_t smem_tma = 2 * (size_t)stage_elems * sizeof(double2) + g.smemB;
  // opt-in: measured SLOWER than the one-tile-per-CTA kernel on B200 (n = 2^20: 62.6 k against 68.0 k points/s; n = 2^16: 242 k against 690 k) --
  // with ~2 us of work per tile, many small CTAs that the hardware schedules beat two resident CTAs per SM that prefetch
  const bool use_tma = env_int("FGP_PV_TMA", 0) != 0 && smem_tma <= 110 * 1024 && (stage_elems * sizeof(double2)) % 512 == 0;  // 32 copies of >= 16 bytes
  static bool attr_tma = false;
  if (use_tma && !attr_tma) {
    cudaFuncSetAttribute(pv_passB_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaGetLastError();
    attr_tma = true;
  }
  for (int64_t p0 = 0; p0 < pairs; p0 += cp) {
    const int64_t np = pairs - p0 < cp ? pairs - p0 : cp;
    a.xs = xs_dev + 2 * p0 * d;
    a.m = m - 2 * p0 < 2 * np ? m - 2 * p0 : 2 * np;
    const dim3 gridA((unsigned)(1 << g.l2), (unsigned)np);
    if (all2) {
      switch (d) {
        case 2: launch_pvA<2, true>(a, gridA, g.threadsA, g.smemA, st); break;
        case 4: launch_pvA<4, true>(a, gridA, g.threadsA, g.smemA, st); break;
        case 8: launch_pvA<8, true>(a, gridA, g.threadsA, g.smemA, st); break;
        case 16: launch_pvA<16, true>(a, gridA, g.threadsA, g.smemA, st); break;
        default: launch_pvA<0, true>(a, gridA, g.threadsA, g.smemA, st); break;
      }
    } else {
      launch_pvA<0, false>(a, gridA, g.threadsA, g.smemA, st);
    }
    FGP_LAUNCH_NAMED("pv_passA", st);
    if (use_tma) {
      int64_t grid = (int64_t)np * a.tilesB;
      if (grid > 2 * (int64_t)sm_count()) grid = 2 * (int64_t)sm_count();
      pv_passB_tma_kernel<<<(unsigned)grid, g.threadsB, smem_tma, st>>>(a, (int)np, stage_elems);
      FGP_LAUNCH_NAMED("pv_passB_tma", st);
    } else {
      pv_passB_kernel<<<dim3((unsigned)a.tilesB, (unsigned)np), g.threadsB, g.smemB, st>>>(a);
      FGP_LAUNCH_NAMED("pv_passB", st);
    }
    pv_final_kernel<<<(unsigned)((np * 32 + 255) / 256), 256, 0, st>>>(a.partial, a.tilesB, np, a.m, kxx, pvar_dev + 2 * p0);
    FGP_LAUNCH_NAMED("pv_final", st);
  }
  return FGP_OK;
}

size_t fgp_dnb2_post_var_C_workspace_bytes(int64_t m, int64_t n) {
  using namespace fgp;
  if (m <= 0 || !is_pow2(n)) return 0;
  const PassGeom g = pvn_geom(n);
  if (!g.l2) return 0;
  const int64_t cp = pvn_chunk_points(m, n);
  const int64_t tilesB = (int64_t(1) << g.l1) >> g.lntrB;
  return align256((size_t)cp * n * sizeof(double)) + align256((size_t)n * sizeof(double)) + align256((size_t)cp * tilesB * sizeof(double));
}

int fgp_dnb2_post_var_C(const double* xs_dev, int64_t m, const uint64_t* C_dev, int mmax, const uint64_t* dshift_host, int t, int64_t n, int d,
                        const int* alpha_host, double scale, const double* ls_host, const double* lam_dev, void* work_dev, double* pvar_dev,
                        fgp_stream_t stream) {
  using namespace fgp;
  FGP_REQUIRE(xs_dev && C_dev && dshift_host && alpha_host && ls_host && lam_dev && work_dev && pvar_dev, "post_var_C: null pointer");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D && m >= 0 && is_pow2(n) && ilog2(n) <= FGP_MAX_LOG2N_WHT && t >= 1 && t < 64, "post_var_C: bad m/n/d/t");
  FGP_REQUIRE(mmax >= 1 && mmax <= 64 && (mmax == 64 || n <= (int64_t(1) << mmax)), "post_var_C: n exceeds 2^mmax generating-matrix columns");
  if (m == 0) return FGP_OK;
  const PassGeom g = pvn_geom(n);
  FGP_REQUIRE(g.l2 >= 1 && g.lntrA == 0, "post_var_C: n=%lld is a single-tile size, use fgp_dnb2_post_var", (long long)n);
  PvnArgs a;
  memset(&a, 0, sizeof(a));
  static const double w0[5] = {0.0, 1.0, 1.5, 43.0 / 18.0 - 1.0, 701.0 / 294.0 - 1.0};
  double kxx = scale;
  bool all2 = true;
  for (int j = 0; j < d; ++j) {
    FGP_REQUIRE(alpha_host[j] >= 1 && alpha_host[j] <= 4, "post_var_C: net alpha outside 1..4");
    a.alpha.v[j] = alpha_host[j];
    a.ls.v[j] = ls_host[j];
    a.dshift.v[j] = dshift_host[j];
    all2 = all2 && alpha_host[j] == 2;
    kxx *= 1.0 + ls_host[j] * w0[alpha_host[j]];
  }
  a.n = n;
  a.d = d;
  a.t = t;
  a.mmax = mmax;
  a.tscale = ldexp(1.0, -t);
  a.C = C_dev;
  a.scale = scale;
  a.l1 = g.l1;
  a.l2 = g.l2;
  a.LPA = g.LPA;
  a.LPB = g.LPB;
  a.lcB = g.lntrB;
  a.tilesB = (int)((int64_t(1) << g.l1) >> g.lntrB);
  a.tab_off = (int)((g.smemA + 15) & ~(size_t)15);
  const size_t smemA = (size_t)a.tab_off + (size_t)(64 + (g.l1 > 6 ? 1 << (g.l1 - 6) : 1)) * d * sizeof(uint64_t);
  const int64_t cp = pvn_chunk_points(m, n);
  char* w = (char*)work_dev;
  a.W = (double*)w;
  w += align256((size_t)cp * n * sizeof(double));
  double* winv = (double*)w;
  a.winv = winv;
  w += align256((size_t)n * sizeof(double));
  a.partial = (double*)w;
  cudaStream_t st = (cudaStream_t)stream;
  {
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    pvn_winv_kernel<<<(unsigned)blocks, 256, 0, st>>>(lam_dev, n, winv);
    FGP_LAUNCH_NAMED("pvn_winv", st);
  }
  static bool attrB = false;
  if (!attrB) {
    cudaFuncSetAttribute(pvn_passB_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaGetLastError();
    attrB = true;
  }
  const int mode = all2 ? (t <= 52 ? 2 : 1) : 0;
  for (int64_t p0 = 0; p0 < m; p0 += cp) {
    const int64_t np = m - p0 < cp ? m - p0 : cp;
    a.xs = xs_dev + p0 * d;
    a.m = np;
    const dim3 gridA((unsigned)(1 << g.l2), (unsigned)((np + kPvnPoints - 1) / kPvnPoints));
    if (mode == 2)
      dispatch_pvnA<2>(a, gridA, g.threadsA, smemA, st);
    else if (mode == 1)
      dispatch_pvnA<1>(a, gridA, g.threadsA, smemA, st);
    else
      launch_pvnA<0, 0>(a, gridA, g.threadsA, smemA, st);
    FGP_LAUNCH_NAMED("pvn_passA", st);
    pvn_passB_kernel<<<dim3((unsigned)a.tilesB, (unsigned)np), g.threadsB, g.smemB, st>>>(a);
    FGP_LAUNCH_NAMED("pvn_passB", st);
    pvn_final_kernel<<<(unsigned)((np * 32 + 255) / 256), 256, 0, st>>>(a.partial, a.tilesB, np, kxx, pvar_dev + p0);
    FGP_LAUNCH_NAMED("pvn_final", st);
  }
  return FGP_OK;
}

}  // extern "C"
