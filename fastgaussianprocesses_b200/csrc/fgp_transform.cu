// K3: stand-alone FFT-BRO (forward / inverse) and FWHT over (batch, n) arrays.
// Replaces qmcpy.fftbr_torch / ifftbr_torch / fwht_torch (fast_gp_lattice.py:224-225, fast_gp_digital_net_b2.py:226).
#include "fgp_transform.cuh"

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

// ---- forward -------------------------------------------------------------------------------------------------
template <bool REAL_IN>
__global__ void __launch_bounds__(256, 2) fft_passA_fwd(const double* __restrict__ in, double2* __restrict__ out,
                                                         int64_t total_blocks, int l1, int l2, int ntr, int LP,
                                                         double scale, FftTables T) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = (double2*)smraw;
  const int64_t blk0 = (int64_t)blockIdx.x * ntr;
  const int nb = (int)min((int64_t)ntr, total_blocks - blk0);
  const int cnt = nb << l1;
  const int64_t g0 = blk0 << l1;
  const int qmask = (1 << l1) - 1;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    double2 v;
    if (REAL_IN) {
      v = make_double2(in[g0 + e] * scale, 0.0);
    } else {
      v = ((const double2*)in)[g0 + e];
      v.x *= scale;
      v.y *= scale;
    }
    sm[(e >> l1) * LP + padidx(e & qmask)] = v;
  }
  __syncthreads();
  block_fft_fwd(sm, l1, nb, LP, T.stage);
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int tr = e >> l1, q = e & qmask;
    double2 v = sm[tr * LP + padidx(q)];
    if (l2) {
      const uint32_t b = (uint32_t)((blk0 + tr) & ((1 << l2) - 1));
      v = cmul(v, twiddle_n(T, brev_bits(b, l2) * (uint32_t)q));
    }
    out[g0 + e] = v;
  }
}

template <bool INV>
__global__ void __launch_bounds__(256, 2) fft_passB(const double2* __restrict__ in, double2* __restrict__ out, int l1,
                                                     int l2, int lntr, int LP, FftTables T) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = (double2*)smraw;
  const int ntr = 1 << lntr;
  const int64_t col0 = (int64_t)blockIdx.x << lntr;
  const int64_t item = col0 >> l1;
  const int q0 = (int)(col0 & ((1 << l1) - 1));
  const int64_t base = (item << (l1 + l2)) + q0;
  const int cnt = ntr << l2;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int cc = e & (ntr - 1), b = e >> lntr;
    sm[cc * LP + padidx(b)] = in[base + ((int64_t)b << l1) + cc];
  }
  __syncthreads();
  if (!INV)
    block_fft_fwd(sm, l2, ntr, LP, T.stage);
  else
    block_fft_inv(sm, l2, ntr, LP, T.stage);
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int cc = e & (ntr - 1), b = e >> lntr;
    double2 v = sm[cc * LP + padidx(b)];
    if (INV) {
      const double2 w = twiddle_n(T, brev_bits((uint32_t)b, l2) * (uint32_t)(q0 + cc));
      v = cmulc(w, v);
    }
    out[base + ((int64_t)b << l1) + cc] = v;
  }
}

__global__ void __launch_bounds__(256, 2) fft_passA_inv(const double2* __restrict__ in, double2* __restrict__ out,
                                                         int64_t total_blocks, int l1, int ntr, int LP, double scale,
                                                         FftTables T) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = (double2*)smraw;
  const int64_t blk0 = (int64_t)blockIdx.x * ntr;
  const int nb = (int)min((int64_t)ntr, total_blocks - blk0);
  const int cnt = nb << l1;
  const int64_t g0 = blk0 << l1;
  const int qmask = (1 << l1) - 1;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) sm[(e >> l1) * LP + padidx(e & qmask)] = in[g0 + e];
  __syncthreads();
  block_fft_inv(sm, l1, nb, LP, T.stage);
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    double2 v = sm[(e >> l1) * LP + padidx(e & qmask)];
    v.x *= scale;
    v.y *= scale;
    out[g0 + e] = v;
  }
}

// ---- FWHT ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) wht_passA(const double* __restrict__ in, double* __restrict__ out,
                                                     int64_t total_blocks, int l1, int ntr, int LP, double scale) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* sm = (double*)smraw;
  const int64_t blk0 = (int64_t)blockIdx.x * ntr;
  const int nb = (int)min((int64_t)ntr, total_blocks - blk0);
  const int cnt = nb << l1;
  const int64_t g0 = blk0 << l1;
  const int qmask = (1 << l1) - 1;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) sm[(e >> l1) * LP + padidx(e & qmask)] = in[g0 + e] * scale;
  __syncthreads();
  block_wht(sm, l1, nb, LP);
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) out[g0 + e] = sm[(e >> l1) * LP + padidx(e & qmask)];
}

__global__ void __launch_bounds__(1024, 1) wht_passB(double* __restrict__ data, int l1, int l2, int lntr, int LP) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* sm = (double*)smraw;
  const int ntr = 1 << lntr;
  const int64_t col0 = (int64_t)blockIdx.x << lntr;
  const int64_t item = col0 >> l1;
  const int q0 = (int)(col0 & ((1 << l1) - 1));
  double* base = data + (item << (l1 + l2)) + q0;
  const int cnt = ntr << l2;
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int cc = e & (ntr - 1), b = e >> lntr;
    sm[cc * LP + padidx(b)] = base[((int64_t)b << l1) + cc];
  }
  __syncthreads();
  block_wht(sm, l2, ntr, LP);
  for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
    const int cc = e & (ntr - 1), b = e >> lntr;
    base[((int64_t)b << l1) + cc] = sm[cc * LP + padidx(b)];
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
  if (real_in) {
    if ((rc = set_smem(fft_passA_fwd<true>, g.smemA))) return rc;
    fft_passA_fwd<true><<<(unsigned)ctas, g.threads, g.smemA, st>>>(in, (double2*)out, total_blocks, g.l1, g.l2, g.ntrA,
                                                                  g.LPA, scale, T);
  } else {
    if ((rc = set_smem(fft_passA_fwd<false>, g.smemA))) return rc;
    fft_passA_fwd<false><<<(unsigned)ctas, g.threads, g.smemA, st>>>(in, (double2*)out, total_blocks, g.l1, g.l2, g.ntrA,
                                                                   g.LPA, scale, T);
  }
  FGP_LAUNCH_CHECK();
  if (g.l2) {
    if ((rc = set_smem(fft_passB<false>, g.smemB))) return rc;
    const int64_t ctasB = (batch << g.l1) / g.ntrB;
    fft_passB<false><<<(unsigned)ctasB, g.threads, g.smemB, st>>>((const double2*)out, (double2*)out, g.l1, g.l2,
                                                                ilog2(g.ntrB), g.LPB, T);
    FGP_LAUNCH_CHECK();
  }
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
  if (g.l2) {
    if ((rc = set_smem(fft_passB<true>, g.smemB))) return rc;
    const int64_t ctasB = (batch << g.l1) / g.ntrB;
    fft_passB<true><<<(unsigned)ctasB, g.threads, g.smemB, st>>>(src, (double2*)out_dev, g.l1, g.l2, ilog2(g.ntrB), g.LPB, T);
    FGP_LAUNCH_CHECK();
    src = (const double2*)out_dev;
  }
  if ((rc = set_smem(fft_passA_inv, g.smemA))) return rc;
  const int64_t total_blocks = batch * (n >> g.l1);
  const int64_t ctas = (total_blocks + g.ntrA - 1) / g.ntrA;
  fft_passA_inv<<<(unsigned)ctas, g.threads, g.smemA, st>>>(src, (double2*)out_dev, total_blocks, g.l1, g.ntrA, g.LPA,
                                                          1.0 / sqrt((double)n), T);
  FGP_LAUNCH_CHECK();
  return FGP_OK;
}

int fgp_fwht(const double* in_dev, double* out_dev, int64_t batch, int64_t n, fgp_stream_t stream) {
  using namespace fgp;
  int rc = check_transform_args(in_dev, out_dev, batch, n, FGP_MAX_LOG2N_WHT, "fwht");
  if (rc) return rc;
  if (batch == 0) return FGP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const PassGeom g = make_geom(n, false, 1024);
  if ((rc = set_smem(wht_passA, g.smemA))) return rc;
  const int64_t total_blocks = batch * (n >> g.l1);
  const int64_t ctas = (total_blocks + g.ntrA - 1) / g.ntrA;
  wht_passA<<<(unsigned)ctas, g.threads, g.smemA, st>>>(in_dev, out_dev, total_blocks, g.l1, g.ntrA, g.LPA,
                                                      1.0 / sqrt((double)n));
  FGP_LAUNCH_CHECK();
  if (g.l2) {
    if ((rc = set_smem(wht_passB, g.smemB))) return rc;
    const int64_t ctasB = (batch << g.l1) / g.ntrB;
    wht_passB<<<(unsigned)ctasB, g.threads, g.smemB, st>>>(out_dev, g.l1, g.l2, ilog2(g.ntrB), g.LPB);
    FGP_LAUNCH_CHECK();
  }
  return FGP_OK;
}

}  // extern "C"
