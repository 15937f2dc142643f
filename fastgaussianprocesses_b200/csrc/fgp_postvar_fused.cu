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
  double2* W = a.W + p * a.n + g0;
  const FftTables T = a.T;
  const uint32_t rb = brev_bits((uint32_t)blk, a.l2);
  block_fft_fwd_io<false>(sm, l1, 0, LP, T.stage, SmemTag{}, [&](int, int idx, double2 v) { W[idx] = cmul(v, twiddle_n(T, rb * (uint32_t)idx)); });
}

__global__ void __launch_bounds__(FGP_LB_THREADS, FGP_LB_BLOCKS) pv_passB_kernel(const __grid_constant__ PvArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ double red[32 * 4];
  double2* sm = (double2*)smraw;
  const int64_t p = blockIdx.y;
  const int tile = blockIdx.x;
  const int l1 = a.l1, l2 = a.l2, LP = a.LPB, lcB = a.lcB;
  const int c = 1 << lcB, L1 = 1 << l1, L2 = 1 << l2, half1 = L1 >> 1;
  const int q0 = tile << lcB;
  // local column tr < c: q = q0 + tr; local column c + t: its mirror L1 - q (the self-mirrored column L1/2 takes the slot of q = 0)
  auto colq = [&](int tr) -> int {
    if (tr < c) return q0 + tr;
    const int q = q0 + tr - c;
    return q == 0 ? half1 : L1 - q;
  };
  const double2* W = a.W + p * a.n;
  block_fft_fwd_io<true>(sm, l2, lcB + 1, LP, a.T.stage, [&](int tr, int r) -> double2 { return W[((int64_t)r << l1) + colq(tr)]; }, SmemTag{});
  __syncthreads();
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

__global__ void __launch_bounds__(256) pv_final_kernel(const double* __restrict__ partial, int tiles, int64_t pairs, int64_t m, double kxx, double* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pairs) return;
  double s0 = 0.0, s1 = 0.0;
  for (int g = 0; g < tiles; ++g) {
    s0 += partial[(p * tiles + g) * 2 + 0];
    s1 += partial[(p * tiles + g) * 2 + 1];
  }
  const double v0 = kxx - s0, v1 = kxx - s1;
  out[2 * p] = v0 < 0.0 ? 0.0 : v0;
  if (2 * p + 1 < m) out[2 * p + 1] = v1 < 0.0 ? 0.0 : v1;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

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

}  // namespace fgp

extern "C" {

size_t fgp_lattice_post_var_z_workspace_bytes(int64_t m, int64_t n) {
  using namespace fgp;
  if (m <= 0 || !is_pow2(n)) return 0;
  const PassGeom g = make_geom(n, true);
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
  const PassGeom g = make_geom(n, true);
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
    pv_passB_kernel<<<dim3((unsigned)a.tilesB, (unsigned)np), g.threadsB, g.smemB, st>>>(a);
    FGP_LAUNCH_NAMED("pv_passB", st);
    pv_final_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(a.partial, a.tilesB, np, a.m, kxx, pvar_dev + 2 * p0);
    FGP_LAUNCH_NAMED("pv_final", st);
  }
  return FGP_OK;
}

}  // extern "C"
