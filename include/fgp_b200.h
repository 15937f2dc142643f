/*
 * fgp_b200.h -- C ABI of libfgp_b200.so: the B200 (sm_100a) structured-covariance hot path of FastGPs.
 *
 * The reference (alegresor/FastGaussianProcesses, `fastgps` 0.0.4.1a) is pure Python and has NO FFI; the seams this
 * library sits behind are Python call sites, cited per entry point as `file:line` into the reference tree.
 * INTEGRATION.md shows the ctypes stubs a maintainer would add at each seam.
 *
 * Conventions
 *   - every function returns 0 on success and a negative FGP_E* code on failure; fgp_last_error() gives the message
 *     (the reference raises AssertionError at the same places; the Python host layer turns codes into those).
 *   - pointers named *_dev are device pointers owned by the caller (PyTorch); pointers named *_host are host arrays
 *     of at most FGP_MAX_D entries that are copied into kernel parameters (no H2D copy, no allocation).
 *   - no function allocates or frees memory; workspaces and twiddle tables are caller-provided
 *     (sizes from the *_bytes functions).  All work is enqueued on `stream` (a cudaStream_t); nothing synchronises.
 *   - all real data are IEEE float64, complex data are interleaved (re,im) float64 pairs, indices int64/uint64.
 *   - "bit-reversed order" (BRO): the natural point order of an extensible lattice/net; transforms take BRO input
 *     and give natural-order output (forward) or the converse (inverse) with no permutation pass.
 */
#ifndef FGP_B200_H
#define FGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGP_VERSION 100
#define FGP_MAX_D 32          /* dimensions handled by the fused kernels */
#define FGP_MAX_ALPHA 10      /* lattice smoothness (Bernoulli order 2*alpha <= 20) */
#define FGP_DERIV_STRIDE 24   /* doubles per (term, dimension) of the derivative-kernel parameter table */
#define FGP_MAX_LOG2N_FFT 24  /* two-pass FFT-BRO: n <= 2^24 */
#define FGP_MAX_LOG2N_WHT 26  /* two-pass FWHT:    n <= 2^26 */

#define FGP_OK 0
#define FGP_EINVAL (-1)   /* bad argument (size not a power of two, d > FGP_MAX_D, null pointer, ...) */
#define FGP_ECUDA (-2)    /* a CUDA runtime call or launch failed */
#define FGP_ENODEV (-3)   /* no sm_100 device */

typedef void* fgp_stream_t; /* cudaStream_t */

int fgp_version(void);
const char* fgp_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t fgp_launch_count(void);
/* per-kernel device timings for bench.py: begin arms event marks after every named launch on `stream`; end synchronises,
 * disarms and returns the number of (name, milliseconds) pairs written (each = time since the previous mark). */
int fgp_profile_begin(fgp_stream_t stream);
int fgp_profile_end(fgp_stream_t stream, int max_entries, const char** names, float* ms);
int fgp_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin);

/* ---------------------------------------------------------------------------------------------------------------
 * K1  point generation      (replaces qmcpy generator calls at abstract_gp.py:307-309 and
 *                            fast_gp_digital_net_b2.py:266-269, cached by util.py:24-38)
 * ------------------------------------------------------------------------------------------------------------- */
/* x[i-i0, j] = ( frac(phi2(i) * z_j) + shift_j ) mod 1 for i in [i0,i1), row-major (i1-i0, d).  Bit-exact integer
 * radical inverse; one IEEE add for the shift. */
int fgp_lattice_points(const uint64_t* z_host, const double* shift_host, int d, uint64_t i0, uint64_t i1,
                       double* x_dev, fgp_stream_t stream);
/* xb[i-i0, j] = XOR_{k: bit k of i} C[j*mmax+k] ^ dshift_j ; x = (double)xb * 2^-t.  C_dev: device (d, mmax). */
int fgp_dnb2_points(const uint64_t* C_dev, int mmax, const uint64_t* dshift_host, int d, int t, uint64_t i0,
                    uint64_t i1, int64_t* xb_dev, double* x_dev, fgp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K2  kernel parts and product kernel   (fast_gp_lattice.py:263-273, fast_gp_digital_net_b2.py:270-301,
 *                                        abstract_fast_gp.py:173-196; first column cached by util.py:50-62)
 * ------------------------------------------------------------------------------------------------------------- */
/* parts[i,j] = c_j * B_{2 alpha_j}((x[i,j]-z[j]) mod 1), c_j = (-1)^(alpha_j+1) (2 pi)^(2 alpha_j)/(2 alpha_j)!  */
int fgp_lattice_kernel_parts(const double* x_dev, int64_t n, int d, const double* z_host, const int* alpha_host,
                             double* parts_dev, fgp_stream_t stream);
/* parts[i,j] = W_{alpha_j}(xb[i,j] ^ zb[j]) - 1  (alpha_j = 1: 1 - 3*2^-beta), beta = t - floor(log2 delta) */
int fgp_dnb2_kernel_parts(const int64_t* xb_dev, int64_t n, int d, const int64_t* zb_host, const int* alpha_host,
                          int t, double* parts_dev, fgp_stream_t stream);
/* k[b,i] = scale[b] * prod_j (1 + ls[b,j] * parts[i,j]);  scale_dev (B), ls_dev (B,d) on the device. */
int fgp_kernel_from_parts(const double* parts_dev, int64_t n, int d, int B, const double* scale_dev,
                          const double* ls_dev, double* k_dev, fgp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K3  fast transforms    (qmcpy.fftbr_torch / ifftbr_torch / fwht_torch injected at fast_gp_lattice.py:224-225 and
 *                         fast_gp_digital_net_b2.py:226; wrapped by abstract_fast_gp.py:197-228)
 *     Orthonormal, along the last dim of a (batch, n) row-major array, n = 2^m.  In-place (in == out) is allowed
 *     when input and output element types match.
 * ------------------------------------------------------------------------------------------------------------- */
size_t fgp_fft_table_bytes(int64_t n);
int fgp_fft_table_init(int64_t n, void* table_dev, fgp_stream_t stream);
/* real BRO input -> complex natural-order output */
int fgp_fftbr_r2c(const double* in_dev, double* out_dev, int64_t batch, int64_t n, const void* table_dev,
                  fgp_stream_t stream);
int fgp_fftbr_c2c(const double* in_dev, double* out_dev, int64_t batch, int64_t n, const void* table_dev,
                  fgp_stream_t stream);
/* complex natural-order input -> complex BRO output (inverse of fgp_fftbr_c2c) */
int fgp_ifftbr_c2c(const double* in_dev, double* out_dev, int64_t batch, int64_t n, const void* table_dev,
                   fgp_stream_t stream);
/* Sylvester-ordered Walsh-Hadamard transform (self-inverse) */
int fgp_fwht(const double* in_dev, double* out_dev, int64_t batch, int64_t n, fgp_stream_t stream);
/* The same transform as ONE persistent kernel (pass-A and pass-B tiles drawn from a ticket counter, the intermediate stays
 * in the L2, HBM traffic = the algorithmic 16n per item instead of 32n).  ctl_dev: at least batch + 2 zeroed uint32; the
 * kernel leaves it zeroed again, so one buffer serves every later call on the same stream.  Falls back to fgp_fwht for
 * single-pass sizes or a null ctl_dev. */
int fgp_fwht_fused(const double* in_dev, double* out_dev, int64_t batch, int64_t n, unsigned int* ctl_dev, fgp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K4  fused eigen-solve + marginal log-likelihood + gradients
 *     (util.py:95-141 _LamCaches, :275-300 __call__ single task, :354-370 norm/logdet; autograd of
 *      abstract_gp.py:253-260,294 replaced by the analytic gradient)
 *
 *     For each of B independent hyperparameter sets b (shared points, per-b data):
 *       k1_i   = scale_b prod_j (1 + ls_bj P_ij)          P from x (lattice) or xb (net), never stored
 *       lam_k  = sum_i T_ki k1_i + noise_b                (unnormalised FFT-BRO / FWHT  ==  sqrt(n) ft(k1) + noise)
 *       norm   = sum_k ysq_bk Re(1/lam_k),  logdet = sum_k log|lam_k|
 *       out[b] = { norm, logdet, dL/dnoise, dL/dscale, dL/dls_0 .. dL/dls_{d-1} },  L = wn_b norm + wl_b logdet
 *     weights_dev (B,2) = (wn_b, wl_b), NULL for (1/2, 1/2): the reference weighs logdet by the number of batch
 *     columns sharing a hyperparameter set (abstract_gp.py:255-256), so the weights are inputs.
 *     ysq_dev (B,n): sum over the reference's batch dims of |ytilde_k|^2 (natural transform order).
 *     lam_dev: optional (B,n) complex (lattice) / real (net) output of lam (may be NULL).
 *     want_grad = 0 skips the backward transform.
 * ------------------------------------------------------------------------------------------------------------- */
size_t fgp_mll_workspace_bytes(int family /*0 lattice, 1 net*/, int64_t n, int d, int B); /* the last 256 bytes: zero once (control words) */
int fgp_lattice_mll_grad(const double* x_dev, int64_t n, int d, const int* alpha_host, int B, const double* ysq_dev,
                         const double* scale_dev, const double* ls_dev, const double* noise_dev,
                         const double* weights_dev, const void* table_dev, void* workspace_dev, double* lam_dev, double* out_dev, int want_grad,
                         fgp_stream_t stream);
/* Generator form for a rank-1 lattice in natural order: the points are never read.  x_i - x_0 = frac(phi2(i) z_j) exactly
 * (the random shift cancels in the first kernel column), so delta is regenerated from the index and z_host[d] inside
 * the first and last pass; this removes the 16*n*d bytes of point traffic per iteration.  n <= 2^24. */
int fgp_lattice_mll_grad_z(const uint64_t* z_host, int64_t n, int d, const int* alpha_host, int B, const double* ysq_dev,
                           const double* scale_dev, const double* ls_dev, const double* noise_dev,
                           const double* weights_dev, const void* table_dev, void* workspace_dev, double* lam_dev, double* out_dev, int want_grad,
                           fgp_stream_t stream);
int fgp_dnb2_mll_grad(const int64_t* xb_dev, int64_t n, int d, const int* alpha_host, int t, int B,
                      const double* ysq_dev, const double* scale_dev, const double* ls_dev, const double* noise_dev,
                      const double* weights_dev, void* workspace_dev, double* lam_dev, double* out_dev, int want_grad, fgp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K4b device-side fit() loop bookkeeping   (abstract_gp.py:236-298 with the default Rprop optimiser of
 *     abstract_fast_gp.py:53-57 and the default (log, exp) transforms of fast_gp_lattice.py:137-139)
 *     One single-CTA kernel per iteration consumes the (B, d+4) output of fgp_*_mll_grad, assembles the loss, runs the
 *     reference's best/save/wait early-stop state machine, snapshots the best iterate, records history rows, applies
 *     torch.optim.Rprop's update to the raw (log) parameters IN PLACE and rewrites the effective hyperparameters
 *     (scale_B, ls_B, noise_B) that the next fgp_*_mll_grad call reads.  Nothing returns to the host, so
 *     [mll_grad, fit_step] x k can be captured in a CUDA graph; the host polls state[4] (stopped) now and then.
 *     state block (doubles): [0] best loss [1] save loss [2] wait [3] next iteration index [4] stopped [5] last
 *     evaluated iteration [6] its loss [7] term1 [8] term2; [9..] options; then prev-grad, step-size, best-raw (P each),
 *     then B+1 completion tickets used by fgp_fit_iteration.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  int B, d;                            /* hyperparameter sets, dimension */
  int n_scale, n_ls_b, n_ls_d, n_noise; /* raw parameter layouts: n_scale,n_ls_b,n_noise in {1,B}; n_ls_d in {1,d} */
  int req_scale, req_ls, req_noise;    /* requires_grad flags */
  double tau;                          /* 1x1 task-kernel value folded into scale and noise (util.py:293,298) */
  double *raw_scale, *raw_ls, *raw_noise; /* device: the nn.Parameter storages, updated in place */
  double *scale_B, *ls_B, *noise_B;    /* device: (B), (B,d), (B) effective values for fgp_*_mll_grad */
  double *state;                       /* device: fgp_fit_state_doubles(P) doubles, P = n_scale+n_ls_b*n_ls_d+n_noise */
  double *loss_hist;                   /* device (hist_capacity,3) rows {loss, term1, term2}, or NULL */
  double *scale_hist, *ls_hist, *noise_hist; /* device (hist_capacity, numel) rows of exp(raw), or NULL */
} fgp_fit_layout;
typedef struct {
  int iterations, stop_wait, hist_capacity;
  double logtol;      /* log(1 + stop_crit_improvement_threshold) */
  double half_const;  /* d_out * n * log(2 pi) / 2 */
  double wn, wl;      /* loss = sum_b (wn norm_b + wl logdet_b) + half_const; MLL: wn = 1/2, wl = d_out/(2B) */
  double lr, etaminus, etaplus, step_min, step_max; /* torch.optim.Rprop: 0.1, 0.5, 1.2, 1e-6, 50 */
} fgp_fit_options;
size_t fgp_fit_state_doubles(int n_raw_params, int B);
int fgp_fit_init(const fgp_fit_layout* layout, const fgp_fit_options* opt, fgp_stream_t stream); /* one launch, nothing synchronises */
int fgp_fit_step(const fgp_fit_layout* layout, const double* mll_out_dev, fgp_stream_t stream);
int fgp_fit_finish(const fgp_fit_layout* layout, fgp_stream_t stream); /* best iterate -> parameters */
/* The same two calls for a layout whose raw_* point at POOLED staging buffers (captured CUDA graphs are tied to their addresses):
 * fgp_fit_init_from first copies the caller's parameter storages (same shapes as the layout's) into the layout's, fgp_fit_finish_to also
 * writes the best iterate out to them -- one launch each instead of three device-to-device copies around the loop (abstract_gp.py:297-298). */
int fgp_fit_init_from(const fgp_fit_layout* layout, const fgp_fit_options* opt, const double* raw_scale_src, const double* raw_ls_src,
                      const double* raw_noise_src, fgp_stream_t stream);
int fgp_fit_finish_to(const fgp_fit_layout* layout, double* raw_scale_dst, double* raw_ls_dst, double* raw_noise_dst, fgp_stream_t stream);
/* One whole fit() iteration in one call: the fused eigen-solve of K4 on layout->scale_B / ls_B / noise_B, whose last
 * CTA reduces the partial sums and runs the fit step in its tail (no separate finalize / fit_step launches).
 * x_dev: points (lattice float64 / net int64 (n,d)); z_host / C_dev: lattice generating vector / net generating matrices
 * for generator mode (x_dev may then be NULL); weights_dev (B,2) as in fgp_*_mll_grad; out_dev (B, d+4) scratch for the reduced terms. */
typedef struct {
  int family;              /* 0 lattice, 1 digital net */
  const void* x_dev;
  const uint64_t* z_host;
  const uint64_t* C_dev;   /* net generating matrices for generator mode (x_dev may then be NULL), (d, mmax) */
  int mmax;
  int64_t n;
  int d;
  const int* alpha_host;
  int t;
  const double* ysq_dev;
  const double* weights_dev;
  const void* table_dev;   /* lattice twiddle table */
  void* workspace_dev;     /* fgp_mll_workspace_bytes */
  double* out_dev;
} fgp_fit_problem;
int fgp_fit_iteration(const fgp_fit_problem* problem, const fgp_fit_layout* layout, fgp_stream_t stream);
/* `iterations` fit() iterations in ONE launch of the persistent cooperative kernel (two-pass sizes): every CTA stays resident,
 * the three passes of an iteration are separated by grid barriers, the fit step runs in the tail of the last pass-C tile and the
 * loop leaves as soon as the early-stop state machine raises state[4] (abstract_gp.py:276-283).  state[4] == 2 reports a
 * device-side failure.  The last 256 bytes of workspace_dev are control words: zero them once after allocation; every launch
 * leaves them zeroed.  fgp_fit_iterations_per_launch: the largest `iterations` this size accepts (1: three launches per iteration). */
int fgp_fit_iterations(const fgp_fit_problem* problem, const fgp_fit_layout* layout, int iterations, fgp_stream_t stream);
int fgp_fit_iterations_per_launch(int family, int64_t n);

/* Generator form for a base-2 digital net in natural order: xb_i ^ xb_0 = XOR_{k in bits(i)} C[j][k] exactly (the digital
 * shift cancels), rebuilt per tile from two shared-memory XOR-fold tables; the (n,d) int64 points are never read.
 * C_dev: device (d, mmax) columns as in fgp_dnb2_points. */
int fgp_dnb2_mll_grad_C(const uint64_t* C_dev, int mmax, int64_t n, int d, const int* alpha_host, int t, int B, const double* ysq_dev,
                        const double* scale_dev, const double* ls_dev, const double* noise_dev, const double* weights_dev,
                        void* workspace_dev, double* lam_dev, double* out_dev, int want_grad, fgp_stream_t stream);

/* K^-1 y for R right-hand sides sharing one spectrum: out = T^-1( T(y) / lam ), util.py:338-344 (single task).
 * lam_dev: (n) complex (family 0) or real (family 1) full eigenvalues sqrt(n) ft(k1)+noise.  y,out: (R,n) real.
 * work_dev: R*n complex (family 0) / unused (family 1, may be NULL). */
int fgp_gram_solve(int family, const double* y_dev, double* out_dev, int64_t R, int64_t n, const double* lam_dev,
                   const void* table_dev, void* work_dev, fgp_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K5  posterior mean / variance as on-the-fly kernel-vector products   (abstract_gp.py:352-380, :381-416)
 * ------------------------------------------------------------------------------------------------------------- */
/* pmean[b,i] = sum_a k_b(xs_i, X_a) coeffs[b,a];  xs (m,d), X (n,d), coeffs (B,n), scale/ls host arrays of one
 * hyperparameter set shared by the B data columns.  partial_dev: workspace of fgp_post_mean_workspace_bytes. */
size_t fgp_post_mean_workspace_bytes(int64_t m, int64_t n, int d, int B);
int fgp_lattice_post_mean(const double* xs_dev, int64_t m, const double* x_dev, int64_t n, int d,
                          const int* alpha_host, double scale, const double* ls_host, const double* coeffs_dev,
                          int B, void* partial_dev, double* pmean_dev, fgp_stream_t stream);
int fgp_dnb2_post_mean(const double* xs_dev, int64_t m, const int64_t* xb_dev, int64_t n, int d,
                       const int* alpha_host, int t, double scale, const double* ls_host, const double* coeffs_dev,
                       int B, void* partial_dev, double* pmean_dev, fgp_stream_t stream);
/* pvar[i] = max(0, k(xs_i,xs_i) - sum_k |T k(xs_i, X)|_k^2 / lam_k)  (abstract_gp.py:381-416).  lam_dev as in fgp_gram_solve.
 * work_dev: fgp_post_var_workspace_bytes.  Lattice: two test points share one complex transform (z = k(x_a,X) + i k(x_b,X),
 * spectra (Z_k +- conj Z_{n-k})/2), the reduction runs over the pairs (k, n-k); net: one real FWHT per test point. */
size_t fgp_post_var_workspace_bytes(int family, int64_t m, int64_t n);
int fgp_lattice_post_var(const double* xs_dev, int64_t m, const double* x_dev, int64_t n, int d,
                         const int* alpha_host, double scale, const double* ls_host, const double* lam_dev,
                         const void* table_dev, void* work_dev, double* pvar_dev, fgp_stream_t stream);
int fgp_dnb2_post_var(const double* xs_dev, int64_t m, const int64_t* xb_dev, int64_t n, int d,
                      const int* alpha_host, int t, double scale, const double* ls_host, const double* lam_dev,
                      void* work_dev, double* pvar_dev, fgp_stream_t stream);
/* Fused generator form of fgp_lattice_post_var for two-pass sizes (n > 2^12): the training points are regenerated from the point
 * index, x_i - shift = frac(phi2(i) z), inside the first transform pass (nothing but the test points is read), and the reduction
 * over the spectral pairs (k, n-k) runs in the epilogue of the second pass (mirror-paired column tiles), so a pair of test points
 * moves 32 n bytes instead of ~96 n and takes two kernels instead of four.  z_host[d] < 2^32, shift_host[d] in [0,1). */
size_t fgp_lattice_post_var_z_workspace_bytes(int64_t m, int64_t n);
int fgp_lattice_post_var_z(const double* xs_dev, int64_t m, const uint64_t* z_host, const double* shift_host, int64_t n, int d,
                           const int* alpha_host, double scale, const double* ls_host, const double* lam_dev, const void* table_dev,
                           void* work_dev, double* pvar_dev, fgp_stream_t stream);
/* The same for a base-2 digital net in generator form: xb_i = XOR_{k in bits(i)} C[j][k] ^ dshift_j is rebuilt per tile from two shared-memory
 * XOR-fold tables, k(x*, X) goes straight into the first pass of the real Walsh-Hadamard transform, sum_k v_k^2 / lam_k is the epilogue of the
 * second.  C_dev: device (d, mmax) columns as in fgp_dnb2_points; dshift_host[d]; lam_dev: (n) real eigenvalues. */
size_t fgp_dnb2_post_var_C_workspace_bytes(int64_t m, int64_t n);
int fgp_dnb2_post_var_C(const double* xs_dev, int64_t m, const uint64_t* C_dev, int mmax, const uint64_t* dshift_host, int t, int64_t n, int d,
                        const int* alpha_host, double scale, const double* ls_host, const double* lam_dev, void* work_dev, double* pvar_dev,
                        fgp_stream_t stream);
/* dense cross-kernel tile K[i,a] = k(xs_i, X_a) (m,n) for post_cov / user-facing kernel() calls */
int fgp_lattice_cross_kernel(const double* xs_dev, int64_t m, const double* x_dev, int64_t n, int d,
                             const int* alpha_host, double scale, const double* ls_host, double* k_dev,
                             fgp_stream_t stream);
int fgp_dnb2_cross_kernel(const double* xs_dev, int64_t m, const int64_t* xb_dev, int64_t n, int d,
                          const int* alpha_host, int t, double scale, const double* ls_host, double* k_dev,
                          fgp_stream_t stream);

/* k[i] = k(x_i, z_i) for N row pairs (the k(x,x) term of abstract_gp.py:407 and elementwise kernel() calls,
 * abstract_gp.py:693-706).  x (N,d) float64; z (N,d) float64, or int64 net integers when z_is_int. */
int fgp_kernel_pairs(int family, const double* x_dev, const void* z_dev, int z_is_int, int64_t N, int d,
                     const int* alpha_host, int t, double scale, const double* ls_host, double* k_dev, fgp_stream_t stream);

/* Derivative-informed kernels (replaces fast_gp_lattice.py:267-273, fast_gp_digital_net_b2.py:289-301 and the term sum of
 * abstract_fast_gp.py:181-191 when `derivatives` is given).  A term is one pair (t0,t1) of derivative multi-indices; per
 * (term, dimension) q = term*d + j the DEVICE tables hold
 *   lattice: ord[q] = degree, par[q*FGP_DERIV_STRIDE + p] = coefficient p of coef*B_order(a), a = (x - z) mod 1
 *   net:     ord[q] = Walsh order 1..4, par[q*S] = (-2)^(beta+kappa), par[q*S+1] = [beta+kappa > 0]
 *   ind[q] = [beta0_j + beta1_j == 0],  w[term] = c0[t0]*c1[t1].
 * fgp_deriv_kernel_parts: parts (n, nterms, d) of n points (float64 lattice / int64 net) against ONE point z_host (d).
 * fgp_deriv_cross_kernel: K[i,a] = scale sum_term w prod_j (ind + ls_j part) for xs (m,d) float64 against x (n,d). */
int fgp_deriv_kernel_parts(int family, const void* x_dev, int64_t n, int d, const void* z_host, int nterms, const int* ord_dev,
                           const double* par_dev, int t, double* parts_dev, fgp_stream_t stream);
int fgp_deriv_cross_kernel(int family, const double* xs_dev, int64_t m, const void* x_dev, int64_t n, int d, int nterms,
                           const int* ord_dev, const double* par_dev, const double* ind_dev, const double* w_dev, int t,
                           double scale, const double* ls_host, double* k_dev, fgp_stream_t stream);

/* K3b: the spectrum of the data in one call (abstract_fast_gp.py:197-212 `ft` applied to y, util.py:164-183 _YtildeCache, and the |ytilde|^2
 * that util.py:364-370 contracts against 1/lam): y (rows, n) float64, rows = lead * B, row l*B + b belongs to hyperparameter set b.
 * ytilde (rows, n) = orthonormal FFT-BRO (family 0: interleaved complex128) / FWHT (family 1: float64) of y, mean-stabilised like the
 * reference (transform y - mean, add mean sqrt(n) back at frequency 0); ysq (B, n) = sum over lead of |ytilde|^2.  work: *_workspace_bytes. */
size_t fgp_data_spectrum_workspace_bytes(int64_t rows, int64_t n);
int fgp_data_spectrum(int family, const double* y_dev, int64_t rows, int64_t B, int64_t n, const void* table_dev, double* ytilde_dev,
                      double* ysq_dev, double* work_dev, fgp_stream_t stream);

/* Per-frequency block systems of the multi-task / derivative-informed eigen-solve (replaces the Schur-complement recursion of
 * util.py:301-323 and the block solve of :354-363): nm independent R x R matrices L[k] (row-major (nm,R,R); interleaved complex when cplx != 0,
 * real otherwise; R <= 16), A[k] = L[k]^-1 and logdet[k] = log|det L[k]| by in-register Gauss-Jordan elimination with partial pivoting. */
int fgp_block_inv_logdet(int cplx, const double* L_dev, int64_t nm, int R, double* A_dev, double* logdet_dev, fgp_stream_t stream);

/* FP64 FMA-chain peak probe used by bench.py for the FP64 roofline denominator: runs `iters` dependent-chain
 * DFMA blocks on every SM and returns the flop count in *flops (time it with events on `stream`). */
int fgp_fp64_peak_probe(int iters, double* sink_dev, double* flops, fgp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FGP_B200_H */
