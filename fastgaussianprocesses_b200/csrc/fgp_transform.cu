// K3: stand-alone FFT-BRO (forward / inverse) and FWHT over (batch, n) arrays.
// Replaces qmcpy.fftbr_torch / ifftbr_torch / fwht_torch (fast_gp_lattice.py:224-225, fast_gp_digital_net_b2.py:226).
#include "fgp_transform.cuh"
#include "fgp_tma.cuh"

namespace fgp {


__global__ void fft_table_kernel(double2* stage, double2* lo, double2* hi, double n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kTabLen) return;
  double s, c;
  // stage-major: index e = h + p
  if (e == 0) {
    stage[0] = make_double2(1.0, 0.0);
  } else {
    const int q = 31 - __clz(e);
    const int h = 1 << q, p = e - h;
    sincospi(-(double)p / (double)h, &s, &c);
    stage[e] = make_double2(c, s);
  }
  sincospi(-2.0 * (double)e / n, &s, &c);
  lo[e] = make_double2(c, s);
  const double nh = n / 4096.0;  // hi[e] = w_n^{4096 e} = exp(-2 pi i e / (n/4096))
  if (nh >= 1.0 && (double)e < nh) {
    sincospi(-2.0 * (double)e / nh, &s, &c);
    hi[e] = make_double2(c, s);
  } else {
    hi[e] = make_double2(1.0, 0.0);
  }
}

// `in` and `out` of the transform kernels may be the SAME buffer (in-place calls are part of the C-ABI contract): every CTA reads its
// whole tile before it writes it and tiles are disjoint, but the pointers are deliberately not __restrict__.
// ---- forward -------------------------------------------------------------------------------------------------
// pass A: 2^lntr contiguous length-2^l1 blocks per CTA; the first round reads global memory straight into registers,
// the last round applies the inter-pass twiddle and writes global memory (coalesced: its elements are 2^(l1-4) apart).
template <bool REAL_IN, int MINB>
__global__ void __launch_bounds__(FGP_LB_THREADS, MINB) fft_passA_fwd(const double* in, double2* out, int64_t total_blocks,
                                                         int l1, int l2, int lntr, int LP, double scale, FftTables T) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = (double2*)smraw;
  const int64_t blk0 = (int64_t)blockIdx.x << lntr;
  const int nb = (int)min((int64_t)1 << lntr, total_blocks - blk0);
  const int64_t g0 = blk0 << l1;
  auto gld = [&](int tr, int idx) -> double2 {
    if (tr >= nb) return make_double2(0.0, 0.0);
    const int64_t e = g0 + ((int64_t)tr << l1) + idx;
    if (REAL_IN) return make_double2(in[e] * scale, 0.0);
    const double2 v = ((const double2*)in)[e];
    return make_double2(v.x * scale, v.y * scale);
  };
  auto gst = [&](int tr, int idx, double2 v) {
    if (tr >= nb) return;
    if (l2) {
      const uint32_t b = (uint32_t)((blk0 + tr) & ((1 << l2) - 1));
      v = cmul(v, twiddle_n(T, brev_bits(b, l2) * (uint32_t)idx));
    }
    out[g0 + ((int64_t)tr << l1) + idx] = v;
  };
  block_fft_fwd_io<false>(sm, l1, lntr, LP, T.stage, gld, gst);
}

// pass B: 2^lntr adjacent stride-2^l1 columns per CTA, in place; consecutive threads take consecutive columns.
template <bool INV, int MINB>
__global__ void __launch_bounds__(FGP_LB_THREADS, MINB) fft_passB(const double2* in, double2* out, int l1, int l2, int lntr,
                                                     int LP, FftTables T) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = (double2*)smraw;
  const int64_t col0 = (int64_t)blockIdx.x << lntr;
  const int64_t item = col0 >> l1;
  const int q0 = (int)(col0 & ((1 << l1) - 1));
  const int64_t base = (item << (l1 + l2)) + q0;
  auto gld = [&](int tr, int idx) -> double2 { return in[base + ((int64_t)idx << l1) + tr]; };
  if (!INV) {
    auto gst = [&](int tr, int idx, double2 v) { out[base + ((int64_t)idx << l1) + tr] = v; };
    block_fft_fwd_io<true>(sm, l2, lntr, LP, T.stage, gld, gst);
  } else {
    auto gst = [&](int tr, int idx, double2 v) {
      const double2 w = twiddle_n(T, brev_bits((uint32_t)idx, l2) * (uint32_t)(q0 + tr));
      out[base + ((int64_t)idx << l1) + tr] = cmulc(w, v);
    };
    block_fft_inv_io<true>(sm, l2, lntr, LP, T.stage, gld, gst);
  }
}

template <int MINB>
__global__ void __launch_bounds__(FGP_LB_THREADS, MINB) fft_passA_inv(const double2* in, double2* out, int64_t total_blocks,
                                                         int l1, int lntr, int LP, double scale, FftTables T) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = (double2*)smraw;
  const int64_t blk0 = (int64_t)blockIdx.x << lntr;
  const int nb = (int)min((int64_t)1 << lntr, total_blocks - blk0);
  const int64_t g0 = blk0 << l1;
  auto gld = [&](int tr, int idx) -> double2 {
    if (tr >= nb) return make_double2(0.0, 0.0);
    return in[g0 + ((int64_t)tr << l1) + idx];
  };
  auto gst = [&](int tr, int idx, double2 v) {
    if (tr >= nb) return;
    out[g0 + ((int64_t)tr << l1) + idx] = make_double2(v.x * scale, v.y * scale);
  };
  block_fft_inv_io<false>(sm, l1, lntr, LP, T.stage, gld, gst);
}

// ---- FWHT ------------------------------------------------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(FGP_LB_THREADS, MINB) wht_passA(const double* in, double* out, int64_t total_blocks, int l1,
                                                     int lntr, int LP, double scale) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* sm = (double*)smraw;
  const int64_t blk0 = (int64_t)blockIdx.x << lntr;
  const int nb = (int)min((int64_t)1 << lntr, total_blocks - blk0);
  const int64_t g0 = blk0 << l1;
  auto gld = [&](int tr, int idx) -> double { return tr < nb ? in[g0 + ((int64_t)tr << l1) + idx] * scale : 0.0; };
  auto gst = [&](int tr, int idx, double v) {
    if (tr < nb) out[g0 + ((int64_t)tr << l1) + idx] = v;
  };
  block_wht_io<false>(sm, l1, lntr, LP, wht_sched_coalesced(l1), gld, gst);
}

template <int MINB>
__global__ void __launch_bounds__(FGP_LB_THREADS, MINB) wht_passB(double* __restrict__ data, int l1, int l2, int lntr, int LP) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* sm = (double*)smraw;
  const int64_t col0 = (int64_t)blockIdx.x << lntr;
  const int64_t item = col0 >> l1;
  const int q0 = (int)(col0 & ((1 << l1) - 1));
  double* base = data + (item << (l1 + l2)) + q0;
  auto gld = [&](int tr, int idx) -> double { return base[((int64_t)idx << l1) + tr]; };
  auto gst = [&](int tr, int idx, double v) { base[((int64_t)idx << l1) + tr] = v; };
  block_wht_io<true>(sm, l2, lntr, LP, wht_sched_up(l2), gld, gst);
}

// ---- fused two-pass FWHT of a batch: ONE persistent kernel, the intermediate stays in the L2 -----------------------
// Pass A of all items followed by pass B of all items moves 32n bytes per item through HBM once the batch exceeds the L2
// (8n in, 8n intermediate out, 8n intermediate in, 8n out).  Here the CTAs of one persistent grid draw tiles from a ticket
// counter in the order A(0) .. A(lag-1), then A(s) B(s-lag) alternating, then the remaining B tiles: a B tile of item b is
// drawn after every A tile of items <= b + lag - 1, so it rarely waits, and it reads what was written a few tens of
// microseconds earlier -- an L2 hit.  HBM traffic drops to the algorithmic 16n.  A B tile waits (acquire-spin on done[b])
// only for A tiles with smaller tickets, which are already running and never wait themselves: no deadlock for any grid size.
struct FusedGeom {
  int l1, l2, lntrA, lntrB, LPA, LPB;
  unsigned nA, nB, B, lag;
};
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <int MINB>
__global__ void __launch_bounds__(FGP_LB_THREADS, MINB) wht_fused(const double* in, double* out, FusedGeom g, double scale,
                                                     unsigned* __restrict__ ctl) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* sm = (double*)smraw;
  __shared__ unsigned s_t;
  const unsigned lag = g.lag < g.B ? g.lag : g.B;
  const unsigned head = lag * g.nA, mid = (g.B - lag) * (g.nA + g.nB), total = g.B * (g.nA + g.nB);
  unsigned* done = ctl + 2;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_t = atomicAdd(&ctl[0], 1u);
    __syncthreads();
    unsigned t = s_t;
    if (t >= total) break;
    unsigned b, idx;
    bool phaseB;
    if (t < head) {
      b = t / g.nA, idx = t - b * g.nA, phaseB = false;
    } else if (t < head + mid) {
      t -= head;
      const unsigned slot = t / (g.nA + g.nB), w = t - slot * (g.nA + g.nB);
      if (w < g.nA)
        b = slot + lag, idx = w, phaseB = false;
      else
        b = slot, idx = w - g.nA, phaseB = true;
    } else {
      t -= head + mid;
      const unsigned slot = t / g.nB;
      b = g.B - lag + slot, idx = t - slot * g.nB, phaseB = true;
    }
    const int64_t item0 = (int64_t)b << (g.l1 + g.l2);
    if (!phaseB) {
      const int64_t g0 = item0 + ((int64_t)idx << (g.l1 + g.lntrA));
      const int l1 = g.l1;
      auto gld = [&](int tr, int e) -> double { return in[g0 + ((int64_t)tr << l1) + e] * scale; };
      auto gst = [&](int tr, int e, double v) { out[g0 + ((int64_t)tr << l1) + e] = v; };
      block_wht_io<false>(sm, g.l1, g.lntrA, g.LPA, wht_sched_coalesced(g.l1), gld, gst);
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&done[b], 1u);
      }
    } else {
      if (threadIdx.x == 0) {
        while (ld_acquire_u32(&done[b]) < g.nA) __nanosleep(40);
      }
      __syncthreads();
      double* base = out + item0 + ((int64_t)idx << g.lntrB);
      const int l1 = g.l1;
      auto gld = [&](int tr, int e) -> double { return __ldcg(base + ((int64_t)e << l1) + tr); };  // written by other SMs: not through L1
      auto gst = [&](int tr, int e, double v) { base[((int64_t)e << l1) + tr] = v; };
      block_wht_io<true>(sm, g.l2, g.lntrB, g.LPB, wht_sched_up(g.l2), gld, gst);
    }
  }
  // the last CTA to leave resets the control block, so one zeroed buffer serves every later call on the same stream
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl[1], 1u) == gridDim.x - 1) {
      for (unsigned i = 0; i < g.B; ++i) done[i] = 0u;
      ctl[0] = 0u;
      __threadfence();
      ctl[1] = 0u;
    }
  }
}

// ---- the same fused transform with TMA-staged, double-buffered tiles ---------------------------------------------------
// wht_fused above is latency-bound: a CTA loads its tile with ordinary loads, waits, computes, stores, and only then draws the
// next ticket (8.1 us per 32 KiB tile, profiles/README.md round 1).  Here the loads are bulk-asynchronous copies
// (cp.async.bulk, the TMA engine; SASS UBLKCP) into a DENSE staging tile, completion is tracked by an mbarrier
// (SYNCS.ARRIVE.TRANS64 / SYNCS.PHASECHK), and there are two staging tiles: while the CTA runs the rounds of tile k out of
// its padded working tile, the TMA engine already fills the other staging tile with tile k+1 -- no registers and no warp
// slots are held by loads in flight.  The first round of a tile reads the dense staging tile (conflict-free: in both
// schedules consecutive threads read consecutive words) and writes the padded working tile; the last round stores to global
// memory from registers.  Same round schedules as wht_passA / wht_passB, hence the same bits.
struct FusedTile {
  unsigned b, idx;
  bool phaseB, valid;
};
__device__ __forceinline__ FusedTile fused_decode(const FusedGeom& g, unsigned t) {
  const unsigned lag = g.lag < g.B ? g.lag : g.B;
  const unsigned head = lag * g.nA, mid = (g.B - lag) * (g.nA + g.nB), total = g.B * (g.nA + g.nB);
  FusedTile o;
  o.valid = t < total;
  o.b = o.idx = 0;
  o.phaseB = false;
  if (!o.valid) return o;
  if (t < head) {
    o.b = t / g.nA, o.idx = t - o.b * g.nA;
  } else if (t < head + mid) {
    t -= head;
    const unsigned slot = t / (g.nA + g.nB), w = t - slot * (g.nA + g.nB);
    if (w < g.nA)
      o.b = slot + lag, o.idx = w;
    else
      o.b = slot, o.idx = w - g.nA, o.phaseB = true;
  } else {
    t -= head + mid;
    const unsigned slot = t / g.nB;
    o.b = g.B - lag + slot, o.idx = t - slot * g.nB, o.phaseB = true;
  }
  return o;
}

template <int MINB>
__global__ void __launch_bounds__(FGP_LB_THREADS, MINB) wht_fused_tma(const double* in, double* out, FusedGeom g, double scale,
                                                         unsigned* __restrict__ ctl, int stage_doubles, const __grid_constant__ CUtensorMap tmapB, int box_rows) {
  extern __shared__ __align__(128) unsigned char smraw_tma[];
  double* stg0 = (double*)smraw_tma;               // dense staging tiles (TMA destinations)
  double* stg1 = stg0 + stage_doubles;
  double* sm = stg1 + stage_doubles;               // padded working tile
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ unsigned s_t;
  unsigned* done = ctl + 2;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ncolsB = 1 << g.lntrB;
  // warp 0 issues the copies of a tile into staging buffer `buf`; returns false (nothing issued) when a pass-B tile's item is
  // not transformed yet and `block` is false
  auto issue = [&](const FusedTile& t, int buf, bool block) -> bool {
    bool ready = true;
    if (t.phaseB) {
      if (threadIdx.x == 0) {
        if (block) {
          while (ld_acquire_u32(&done[t.b]) < g.nA) __nanosleep(40);
        } else {
          ready = ld_acquire_u32(&done[t.b]) >= g.nA;
        }
      }
      ready = __shfl_sync(0xffffffffu, (int)ready, 0) != 0;
      if (!ready) return false;
    }
    double* dst = buf ? stg1 : stg0;
    uint64_t* bar = &bars[buf];
    const int64_t item0 = (int64_t)t.b << (g.l1 + g.l2);
    const int lane = threadIdx.x & 31;
    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes (other CTAs' stores, our own reads of dst) before the async proxy
    if (!t.phaseB) {
      const unsigned tile_bytes = 8u << (g.l1 + g.lntrA);
      if (lane == 0) mbar_arrive_expect_tx(bar, tile_bytes);
      __syncwarp();
      const double* src = in + item0 + ((int64_t)t.idx << (g.l1 + g.lntrA));
      const unsigned chunk = tile_bytes / 32;  // one 1/32 of the tile per lane (>= 16 bytes for every geometry used)
      tma_load_1d((char*)dst + (size_t)lane * chunk, (const char*)src + (size_t)lane * chunk, chunk, bar);
    } else {
      // the tile is 2^l2 rows of ncolsB adjacent columns of item b seen as an (L2, L1) matrix: one tiled TMA load per box of
      // box_rows rows (a whole tile for L2 <= 256) instead of one small bulk copy per row (measured 2.5x slower than plain loads)
      const int rows = 1 << g.l2;
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, 8u * ncolsB * rows);
        const int c0 = (int)(t.idx << g.lntrB), r0 = (int)(t.b << g.l2);
        for (int r = 0; r < rows; r += box_rows) tma_load_2d(dst + (size_t)r * ncolsB, &tmapB, c0, r0 + r, bar);
      }
      __syncwarp();
    }
    return true;
  };
  auto draw = [&]() -> FusedTile {
    __syncthreads();
    if (threadIdx.x == 0) s_t = atomicAdd(&ctl[0], 1u);
    __syncthreads();
    return fused_decode(g, s_t);
  };
  unsigned parity[2] = {0u, 0u};
  int buf = 0;
  FusedTile cur = draw();
  if (cur.valid && threadIdx.x < 32) issue(cur, 0, true);
  while (cur.valid) {
    const FusedTile nxt = draw();
    bool issued = false;
    if (nxt.valid && threadIdx.x < 32) issued = issue(nxt, buf ^ 1, false);
    mbar_wait(&bars[buf], parity[buf]);
    parity[buf] ^= 1u;
    const double* stg = buf ? stg1 : stg0;
    const int64_t item0 = (int64_t)cur.b << (g.l1 + g.l2);
    if (!cur.phaseB) {
      const int64_t g0 = item0 + ((int64_t)cur.idx << (g.l1 + g.lntrA));
      const int l1 = g.l1;
      auto gld = [&](int tr, int e) -> double { return stg[(tr << l1) + e] * scale; };
      auto gst = [&](int tr, int e, double v) { out[g0 + ((int64_t)tr << l1) + e] = v; };
      block_wht_io<false>(sm, g.l1, g.lntrA, g.LPA, wht_sched_coalesced(g.l1), gld, gst);
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&done[cur.b], 1u);
      }
    } else {
      double* base = out + item0 + ((int64_t)cur.idx << g.lntrB);
      const int l1 = g.l1, lc = g.lntrB;
      auto gld = [&](int tr, int e) -> double { return stg[(e << lc) + tr]; };
      auto gst = [&](int tr, int e, double v) { base[((int64_t)e << l1) + tr] = v; };
      block_wht_io<true>(sm, g.l2, g.lntrB, g.LPB, wht_sched_up(g.l2), gld, gst);
    }
    if (nxt.valid && threadIdx.x < 32) {
      // uniform over warp 0: `issued` came out of a warp-wide shuffle
      if (!issued) issue(nxt, buf ^ 1, true);
    }
    buf ^= 1;
    cur = nxt;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl[1], 1u) == gridDim.x - 1) {
      for (unsigned i = 0; i < g.B; ++i) done[i] = 0u;
      ctl[0] = 0u;
      __threadfence();
      ctl[1] = 0u;
    }
  }
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 24 * 1024) {  // static shared memory counts towards the 48 KiB default limit
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(%zu bytes): %s", bytes, cudaGetErrorString(e));
      return FGP_ECUDA;
    }
  }
  return FGP_OK;
}

static int check_transform_args(const void* in, const void* out, int64_t batch, int64_t n, int maxlog, const char* who) {
  if (!in || !out) {
    set_error("%s: null pointer", who);
    return FGP_EINVAL;
  }
  if (batch < 0 || !is_pow2(n) || ilog2(n) > maxlog) {
    set_error("%s: need batch >= 0 and n a power of two <= 2^%d (got batch=%lld n=%lld)", who, maxlog, (long long)batch,
              (long long)n);
    return FGP_EINVAL;
  }
  return FGP_OK;
}

static int fft_forward(const double* in, double* out, int64_t batch, int64_t n, const void* table, bool real_in,
                       cudaStream_t st) {
  const PassGeom g = make_geom(n, true);
  const FftTables T = make_tables(table);
  const int64_t total_blocks = batch * (n >> g.l1);
  const int64_t ctas = (total_blocks + g.ntrA - 1) / g.ntrA;
  const double scale = 1.0 / sqrt((double)n);
  int rc;
  const int64_t ctasB = g.l2 ? (batch << g.l1) >> g.lntrB : 0;
  // 32 KiB tiles: 64 registers, 4 CTAs per SM; larger tiles: 128 registers, 2 per SM (FGP_FFT_MINB overrides for tuning)
  static const int minb_env = env_int("FGP_FFT_MINB", 0);
  const bool small = minb_env ? minb_env == 4 : (g.smemA <= 40 * 1024 && g.smemB <= 40 * 1024);
#define FGP_FFT_FWD(MINB)                                                                                                        \
  do {                                                                                                                           \
    if (real_in) {                                                                                                               \
      if ((rc = set_smem(fft_passA_fwd<true, MINB>, g.smemA))) return rc;                                                        \
      fft_passA_fwd<true, MINB><<<(unsigned)ctas, g.threadsA, g.smemA, st>>>(in, (double2*)out, total_blocks, g.l1, g.l2, g.lntrA, g.LPA, scale, T); \
    } else {                                                                                                                     \
      if ((rc = set_smem(fft_passA_fwd<false, MINB>, g.smemA))) return rc;                                                       \
      fft_passA_fwd<false, MINB><<<(unsigned)ctas, g.threadsA, g.smemA, st>>>(in, (double2*)out, total_blocks, g.l1, g.l2, g.lntrA, g.LPA, scale, T); \
    }                                                                                                                            \
    FGP_LAUNCH_NAMED("fft_passA_fwd", st);                                                                                       \
    if (g.l2) {                                                                                                                  \
      if ((rc = set_smem(fft_passB<false, MINB>, g.smemB))) return rc;                                                           \
      fft_passB<false, MINB><<<(unsigned)ctasB, g.threadsB, g.smemB, st>>>((const double2*)out, (double2*)out, g.l1, g.l2, g.lntrB, g.LPB, T); \
      FGP_LAUNCH_NAMED("fft_passB_fwd", st);                                                                                     \
    }                                                                                                                            \
  } while (0)
  if (small)
    FGP_FFT_FWD(4);
  else
    FGP_FFT_FWD(FGP_LB_BLOCKS);
#undef FGP_FFT_FWD
  return FGP_OK;
}

}  // namespace fgp

extern "C" {

size_t fgp_fft_table_bytes(int64_t n) {
  (void)n;
  return (size_t)3 * fgp::kTabLen * sizeof(double2);
}

int fgp_fft_table_init(int64_t n, void* table_dev, fgp_stream_t stream) {
  FGP_REQUIRE(table_dev, "fft_table_init: null table");
  FGP_REQUIRE(fgp::is_pow2(n) && fgp::ilog2(n) <= FGP_MAX_LOG2N_FFT, "fft_table_init: n=%lld not a power of two <= 2^%d",
              (long long)n, FGP_MAX_LOG2N_FFT);
  double2* s = (double2*)table_dev;
  fgp::fft_table_kernel<<<fgp::kTabLen / 256, 256, 0, (cudaStream_t)stream>>>(s, s + fgp::kTabLen, s + 2 * fgp::kTabLen,
                                                                            (double)n);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

int fgp_fftbr_r2c(const double* in_dev, double* out_dev, int64_t batch, int64_t n, const void* table_dev,
                  fgp_stream_t stream) {
  int rc = fgp::check_transform_args(in_dev, out_dev, batch, n, FGP_MAX_LOG2N_FFT, "fftbr_r2c");
  if (rc) return rc;
  FGP_REQUIRE(table_dev, "fftbr_r2c: null table");
  FGP_REQUIRE((const void*)in_dev != (const void*)out_dev, "fftbr_r2c: in-place real->complex is not supported");
  if (batch == 0) return FGP_OK;
  return fgp::fft_forward(in_dev, out_dev, batch, n, table_dev, true, (cudaStream_t)stream);
}

int fgp_fftbr_c2c(const double* in_dev, double* out_dev, int64_t batch, int64_t n, const void* table_dev,
                  fgp_stream_t stream) {
  int rc = fgp::check_transform_args(in_dev, out_dev, batch, n, FGP_MAX_LOG2N_FFT, "fftbr_c2c");
  if (rc) return rc;
  FGP_REQUIRE(table_dev, "fftbr_c2c: null table");
  if (batch == 0) return FGP_OK;
  return fgp::fft_forward(in_dev, out_dev, batch, n, table_dev, false, (cudaStream_t)stream);
}

int fgp_ifftbr_c2c(const double* in_dev, double* out_dev, int64_t batch, int64_t n, const void* table_dev,
                   fgp_stream_t stream) {
  using namespace fgp;
  int rc = check_transform_args(in_dev, out_dev, batch, n, FGP_MAX_LOG2N_FFT, "ifftbr_c2c");
  if (rc) return rc;
  FGP_REQUIRE(table_dev, "ifftbr_c2c: null table");
  if (batch == 0) return FGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const PassGeom g = make_geom(n, true);
  const FftTables T = make_tables(table_dev);
  const double2* src = (const double2*)in_dev;
  const int64_t ctasB = g.l2 ? (batch << g.l1) >> g.lntrB : 0;
  const int64_t total_blocks = batch * (n >> g.l1);
  const int64_t ctas = (total_blocks + g.ntrA - 1) / g.ntrA;
  const double scale = 1.0 / sqrt((double)n);
  static const int minb_env = env_int("FGP_FFT_MINB", 0);
  const bool small = minb_env ? minb_env == 4 : (g.smemA <= 40 * 1024 && g.smemB <= 40 * 1024);
#define FGP_FFT_INV(MINB)                                                                                                   \
  do {                                                                                                                      \
    if (g.l2) {                                                                                                             \
      if ((rc = set_smem(fft_passB<true, MINB>, g.smemB))) return rc;                                                       \
      fft_passB<true, MINB><<<(unsigned)ctasB, g.threadsB, g.smemB, st>>>(src, (double2*)out_dev, g.l1, g.l2, g.lntrB, g.LPB, T); \
      FGP_LAUNCH_NAMED("fft_passB_inv", st);                                                                                \
      src = (const double2*)out_dev;                                                                                        \
    }                                                                                                                       \
    if ((rc = set_smem(fft_passA_inv<MINB>, g.smemA))) return rc;                                                           \
    fft_passA_inv<MINB><<<(unsigned)ctas, g.threadsA, g.smemA, st>>>(src, (double2*)out_dev, total_blocks, g.l1, g.lntrA, g.LPA, scale, T); \
    FGP_LAUNCH_NAMED("fft_passA_inv", st);                                                                                  \
  } while (0)
  if (small)
    FGP_FFT_INV(4);
  else
    FGP_FFT_INV(FGP_LB_BLOCKS);
#undef FGP_FFT_INV
  return FGP_OK;
}

int fgp_fwht(const double* in_dev, double* out_dev, int64_t batch, int64_t n, fgp_stream_t stream) {
  using namespace fgp;
  int rc = check_transform_args(in_dev, out_dev, batch, n, FGP_MAX_LOG2N_WHT, "fwht");
  if (rc) return rc;
  if (batch == 0) return FGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const PassGeom g = make_geom(n, false, true);
  const int64_t total_blocks = batch * (n >> g.l1);
  const int64_t ctas = (total_blocks + g.ntrA - 1) / g.ntrA;
  const int64_t ctasB = g.l2 ? (batch << g.l1) >> g.lntrB : 0;
  const double scale = 1.0 / sqrt((double)n);
  // 32 KiB tiles: 64 registers, 4 CTAs per SM in different phases; 64 KiB tiles (HBM-sized data): 128 registers, 2 per SM
  if (g.smemA <= 40 * 1024 && g.smemB <= 72 * 1024) {
    if ((rc = set_smem(wht_passA<FGP_LB_BLOCKS_R>, g.smemA))) return rc;
    wht_passA<FGP_LB_BLOCKS_R><<<(unsigned)ctas, g.threadsA, g.smemA, st>>>(in_dev, out_dev, total_blocks, g.l1, g.lntrA, g.LPA, scale);
    FGP_LAUNCH_NAMED("wht_passA", st);
    if (g.l2) {
      if ((rc = set_smem(wht_passB<FGP_LB_BLOCKS_R>, g.smemB))) return rc;
      wht_passB<FGP_LB_BLOCKS_R><<<(unsigned)ctasB, g.threadsB, g.smemB, st>>>(out_dev, g.l1, g.l2, g.lntrB, g.LPB);
      FGP_LAUNCH_NAMED("wht_passB", st);
    }
  } else {
    if ((rc = set_smem(wht_passA<FGP_LB_BLOCKS>, g.smemA))) return rc;
    wht_passA<FGP_LB_BLOCKS><<<(unsigned)ctas, g.threadsA, g.smemA, st>>>(in_dev, out_dev, total_blocks, g.l1, g.lntrA, g.LPA, scale);
    FGP_LAUNCH_NAMED("wht_passA", st);
    if (g.l2) {
      if ((rc = set_smem(wht_passB<FGP_LB_BLOCKS>, g.smemB))) return rc;
      wht_passB<FGP_LB_BLOCKS><<<(unsigned)ctasB, g.threadsB, g.smemB, st>>>(out_dev, g.l1, g.l2, g.lntrB, g.LPB);
      FGP_LAUNCH_NAMED("wht_passB", st);
    }
  }
  return FGP_OK;
}

int fgp_fwht_fused(const double* in_dev, double* out_dev, int64_t batch, int64_t n, unsigned int* ctl_dev, fgp_stream_t stream) {
  using namespace fgp;
  int rc = check_transform_args(in_dev, out_dev, batch, n, FGP_MAX_LOG2N_WHT, "fwht_fused");
  if (rc) return rc;
  if (batch == 0) return FGP_OK;
  const PassGeom g = make_geom(n, false, true);
  // single-pass sizes have no intermediate; fall through to the plain kernels
  if (!g.l2 || !ctl_dev || batch > (int64_t(1) << 24)) return fgp_fwht(in_dev, out_dev, batch, n, stream);
  cudaStream_t st = (cudaStream_t)stream;
  FusedGeom f;
  f.l1 = g.l1, f.l2 = g.l2, f.lntrA = g.lntrA, f.lntrB = g.lntrB, f.LPA = g.LPA, f.LPB = g.LPB;
  f.nA = (unsigned)g.ctasA;
  f.nB = (unsigned)g.ctasB;
  f.B = (unsigned)batch;
  static const int lag_env = env_int("FGP_FUSED_LAG", 0);
  // enough look-ahead to cover the grid: the B tiles of an item are drawn once ~2 grids worth of A tiles have been drawn
  const int sms = sm_count();
  f.lag = lag_env > 0 ? (unsigned)lag_env : (unsigned)((2 * 4 * sms + f.nA - 1) / f.nA);
  if (f.lag < 1) f.lag = 1;
  const size_t smem = g.smemA > g.smemB ? g.smemA : g.smemB;
  const int threads = g.threadsA > g.threadsB ? g.threadsA : g.threadsB;
  const double scale = 1.0 / sqrt((double)n);
  const int64_t tiles = (int64_t)f.B * (f.nA + f.nB);
  // TMA-staged variant: two dense staging tiles + the padded working tile; rows of a pass-B tile must be >= 16 bytes
  static const int no_tma = env_int("FGP_FUSED_NO_TMA", 0);
  const int tile_log = (g.l1 + g.lntrA) > (g.l2 + g.lntrB) ? (g.l1 + g.lntrA) : (g.l2 + g.lntrB);
  const size_t smem_tma = 2 * (sizeof(double) << tile_log) + smem;
  if (!no_tma && g.lntrB >= 1 && (g.l1 + g.lntrA) >= 9 && smem_tma <= 110 * 1024 && (((uintptr_t)in_dev | (uintptr_t)out_dev) & 15) == 0) {  // lntrB >= 1: box rows of >= 16 bytes
    if ((rc = set_smem(wht_fused_tma<2>, smem_tma))) return rc;
    const int64_t grid = tiles < (int64_t)sms * 2 ? tiles : (int64_t)sms * 2;
    CUtensorMap tmapB;
    const int box_rows = (1 << g.l2) < 256 ? (1 << g.l2) : 256;
    if ((rc = make_tmap_2d(out_dev, (uint64_t)1 << g.l1, (uint64_t)batch << g.l2, 1u << g.lntrB, (uint32_t)box_rows, &tmapB))) return rc;
    wht_fused_tma<2><<<(unsigned)grid, threads, smem_tma, st>>>(in_dev, out_dev, f, scale, ctl_dev, 1 << tile_log, tmapB, box_rows);
    FGP_LAUNCH_NAMED("wht_fused_tma", st);
    return FGP_OK;
  }
  if (smem <= 40 * 1024) {
    if ((rc = set_smem(wht_fused<FGP_LB_BLOCKS_R>, smem))) return rc;
    const int64_t grid = tiles < (int64_t)sms * FGP_LB_BLOCKS_R ? tiles : (int64_t)sms * FGP_LB_BLOCKS_R;
    wht_fused<FGP_LB_BLOCKS_R><<<(unsigned)grid, threads, smem, st>>>(in_dev, out_dev, f, scale, ctl_dev);
  } else {
    if ((rc = set_smem(wht_fused<FGP_LB_BLOCKS>, smem))) return rc;
    const int64_t grid = tiles < (int64_t)sms * FGP_LB_BLOCKS ? tiles : (int64_t)sms * FGP_LB_BLOCKS;
    wht_fused<FGP_LB_BLOCKS><<<(unsigned)grid, threads, smem, st>>>(in_dev, out_dev, f, scale, ctl_dev);
  }
  FGP_LAUNCH_NAMED("wht_fused", st);
  return FGP_OK;
}

}  // extern "C"
