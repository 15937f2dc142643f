// K4 host side: argument checks, geometry, workspace carving and dispatch to the per-variant translation units.
#include "fgp_mll.cuh"

namespace fgp {

static int mll_common(bool net, const uint64_t* z_host, const uint64_t* C_dev, int mmax, const void* x, int64_t n, int d, const int* alpha_host, int t, int B,
                      const double* ysq, const double* scale, const double* ls, const double* noise, const double* weights,
                      const void* table, void* workspace, double* lam, double* out, int want_grad, fgp_stream_t stream,
                      const fgp_fit_layout* fit = nullptr, int iters = 1) {
  FGP_REQUIRE((x || z_host || C_dev) && alpha_host && ysq && scale && ls && noise && out, "mll_grad: null pointer");
  FGP_REQUIRE(!(net && z_host) && !(!net && C_dev), "mll_grad: generator arguments do not match the family");
  if (C_dev) FGP_REQUIRE(mmax >= 1 && mmax <= 64 && (mmax == 64 || n <= (int64_t(1) << mmax)), "mll_grad: n exceeds 2^mmax generating-matrix columns");
  FGP_REQUIRE(d >= 1 && d <= FGP_MAX_D, "mll_grad: d=%d outside 1..%d", d, FGP_MAX_D);
  FGP_REQUIRE(B >= 1 && B <= 65535, "mll_grad: B=%d outside 1..65535", B);
  FGP_REQUIRE(is_pow2(n) && ilog2(n) <= (net ? FGP_MAX_LOG2N_WHT : FGP_MAX_LOG2N_FFT),
              "mll_grad: n=%lld must be a power of two <= 2^%d", (long long)n, net ? FGP_MAX_LOG2N_WHT : FGP_MAX_LOG2N_FFT);
  MllArgs a;
  memset(&a, 0, sizeof(a));
  a.x = (z_host || C_dev) ? nullptr : x;
  a.C = C_dev;
  a.mmax = mmax;
  if (z_host)
    for (int j = 0; j < d; ++j) a.z.v[j] = z_host[j];
  a.n = n;
  a.d = d;
  a.t = t;
  a.tscale = ldexp(1.0, -t);
  bool all2 = true;
  for (int j = 0; j < d; ++j) all2 = all2 && alpha_host[j] == 2;
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
  const PassGeom g = make_geom(n, !net, false, true, B);
  a.ysq = ysq;
  a.scale = scale;
  a.ls = ls;
  a.noise = noise;
  a.weights = weights;
  a.lam = lam;
  a.out = out;
  a.want_grad = want_grad;
  if (fit) {
    int rc = make_layout(fit, &a.fit);
    if (rc) return rc;
    FGP_REQUIRE(a.fit.B == B && a.fit.d == d, "fit_iteration: layout B/d do not match the problem");
    a.has_fit = 1;
  }
  a.iters = iters;
  a.pdl = pdl_mode();
  a.l1 = g.l1;
  a.l2 = g.l2;
  a.lntrA = g.l2 ? g.lntrA : 0;  // the single-pass kernel runs one transform per CTA
  a.lntrB = g.lntrB;
  a.LPA = g.LPA;
  a.LPB = g.LPB;
  a.tab_off = (int)((g.smemA + 15) & ~(size_t)15);
  a.ctasA = (int)g.ctasA;
  a.ctasB = (int)g.ctasB;
  // half-spectrum mode of the lattice two-pass kernels (fgp_mll.cuh); the full spectrum is computed when the caller wants lam
  static const int no_hs = env_int("FGP_NO_HS", 0);
  if (!net && g.l2 >= 1 && g.lntrA == 0 && g.lntrB < g.l1 && !lam && !no_hs) {
    a.hs = 1;
    a.ctasA = (1 << (g.l2 - 1)) + 1;
    a.ctasB = ((1 << (g.l1 - 1)) >> g.lntrB) + 1;
  }
  if (g.l2) {
    FGP_REQUIRE(workspace, "mll_grad: null workspace");
    const size_t wbytes = align256((size_t)B * n * (net ? sizeof(double) : sizeof(double2)));
    const size_t pb = align256((size_t)B * a.ctasB * 3 * sizeof(double));
    a.W = workspace;
    a.partB = (double*)((char*)workspace + wbytes);
    a.partC = (double*)((char*)workspace + wbytes + pb);
    // control words of the persistent kernel behind the partial sums (fgp_mll_workspace_bytes reserves them)
    const size_t pc = align256((size_t)B * g.ctasA * (d + 1) * sizeof(double));
    a.bar = (unsigned int*)((char*)workspace + wbytes + align256((size_t)B * g.ctasB * 3 * sizeof(double)) + pc);
  }
  FGP_REQUIRE(iters == 1 || (fit && g.l2 && coop_enabled()) , "fit_iterations: several iterations per launch need the persistent two-pass kernel");
  mll_launch_fn fn;
  if (net && C_dev) {
    fn = mll_net_z_gen_alpha;
    if (all2) fn = d == 2 ? mll_net_z_a2_d2 : d == 4 ? mll_net_z_a2_d4 : d == 8 ? mll_net_z_a2_d8 : d == 16 ? mll_net_z_a2_d16 : fn;
  } else if (net) {
    fn = mll_net_x_gen_alpha;
    if (all2) fn = d == 2 ? mll_net_x_a2_d2 : d == 4 ? mll_net_x_a2_d4 : d == 8 ? mll_net_x_a2_d8 : d == 16 ? mll_net_x_a2_d16 : fn;
  } else if (z_host) {
    fn = mll_lat_z_gen_alpha;
    if (all2) fn = d == 2 ? mll_lat_z_a2_d2 : d == 4 ? mll_lat_z_a2_d4 : d == 8 ? mll_lat_z_a2_d8 : d == 16 ? mll_lat_z_a2_d16 : fn;
  } else {
    fn = mll_lat_x_gen_alpha;
    if (all2) fn = d == 2 ? mll_lat_x_a2_d2 : d == 4 ? mll_lat_x_a2_d4 : d == 8 ? mll_lat_x_a2_d8 : d == 16 ? mll_lat_x_a2_d16 : fn;
  }
  return fn(a, g, B, (cudaStream_t)stream);
}

}  // namespace fgp

extern "C" {

size_t fgp_mll_workspace_bytes(int family, int64_t n, int d, int B) {
  using namespace fgp;
  if (!is_pow2(n) || B < 1 || d < 1) return 0;
  const bool net = family != 0;
  const PassGeom g = make_geom(n, !net, false, true, B);
  if (g.l2 == 0) return 256;
  return align256((size_t)B * n * (net ? sizeof(double) : sizeof(double2))) + align256((size_t)B * g.ctasB * 3 * sizeof(double)) +
         align256((size_t)B * g.ctasA * (d + 1) * sizeof(double)) + 256 /* control words of the persistent kernel */;
}

int fgp_lattice_mll_grad(const double* x_dev, int64_t n, int d, const int* alpha_host, int B, const double* ysq_dev, const double* scale_dev,
                         const double* ls_dev, const double* noise_dev, const double* weights_dev, const void* table_dev, void* workspace_dev,
                         double* lam_dev, double* out_dev, int want_grad, fgp_stream_t stream) {
  return fgp::mll_common(false, nullptr, nullptr, 0, x_dev, n, d, alpha_host, 0, B, ysq_dev, scale_dev, ls_dev, noise_dev, weights_dev, table_dev,
                         workspace_dev, lam_dev, out_dev, want_grad, stream);
}

int fgp_lattice_mll_grad_z(const uint64_t* z_host, int64_t n, int d, const int* alpha_host, int B, const double* ysq_dev,
                           const double* scale_dev, const double* ls_dev, const double* noise_dev, const double* weights_dev,
                           const void* table_dev, void* workspace_dev, double* lam_dev, double* out_dev, int want_grad, fgp_stream_t stream) {
  FGP_REQUIRE(z_host, "mll_grad_z: null generating vector");
  return fgp::mll_common(false, z_host, nullptr, 0, nullptr, n, d, alpha_host, 0, B, ysq_dev, scale_dev, ls_dev, noise_dev, weights_dev, table_dev,
                         workspace_dev, lam_dev, out_dev, want_grad, stream);
}

int fgp_dnb2_mll_grad_C(const uint64_t* C_dev, int mmax, int64_t n, int d, const int* alpha_host, int t, int B, const double* ysq_dev,
                        const double* scale_dev, const double* ls_dev, const double* noise_dev, const double* weights_dev, void* workspace_dev,
                        double* lam_dev, double* out_dev, int want_grad, fgp_stream_t stream) {
  FGP_REQUIRE(C_dev, "mll_grad_C: null generating matrices");
  return fgp::mll_common(true, nullptr, C_dev, mmax, nullptr, n, d, alpha_host, t, B, ysq_dev, scale_dev, ls_dev, noise_dev, weights_dev, nullptr,
                         workspace_dev, lam_dev, out_dev, want_grad, stream);
}

int fgp_fit_iteration(const fgp_fit_problem* p, const fgp_fit_layout* layout, fgp_stream_t stream) {
  FGP_REQUIRE(p && layout, "fit_iteration: null argument");
  const bool net = p->family != 0;
  const int want_grad = (layout->req_scale || layout->req_ls || layout->req_noise) ? 1 : 0;
  return fgp::mll_common(net, net ? nullptr : p->z_host, net ? p->C_dev : nullptr, p->mmax, p->x_dev, p->n, p->d, p->alpha_host, p->t, layout->B, p->ysq_dev, layout->scale_B,
                         layout->ls_B, layout->noise_B, p->weights_dev, p->table_dev, p->workspace_dev, nullptr, p->out_dev, want_grad,
                         stream, layout);
}

int fgp_fit_iterations(const fgp_fit_problem* p, const fgp_fit_layout* layout, int iterations, fgp_stream_t stream) {
  FGP_REQUIRE(p && layout && iterations >= 1, "fit_iterations: bad argument");
  const bool net = p->family != 0;
  const int want_grad = (layout->req_scale || layout->req_ls || layout->req_noise) ? 1 : 0;
  return fgp::mll_common(net, net ? nullptr : p->z_host, net ? p->C_dev : nullptr, p->mmax, p->x_dev, p->n, p->d, p->alpha_host, p->t, layout->B, p->ysq_dev, layout->scale_B,
                         layout->ls_B, layout->noise_B, p->weights_dev, p->table_dev, p->workspace_dev, nullptr, p->out_dev, want_grad,
                         stream, layout, iterations);
}

int fgp_fit_iterations_per_launch(int family, int64_t n) {
  using namespace fgp;
  if (!is_pow2(n)) return 0;
  const PassGeom g = make_geom(n, family == 0, false, true);
  return (g.l2 && coop_enabled()) ? 1 << 20 : 1;
}

int fgp_dnb2_mll_grad(const int64_t* xb_dev, int64_t n, int d, const int* alpha_host, int t, int B, const double* ysq_dev,
                      const double* scale_dev, const double* ls_dev, const double* noise_dev, const double* weights_dev, void* workspace_dev,
                      double* lam_dev, double* out_dev, int want_grad, fgp_stream_t stream) {
  return fgp::mll_common(true, nullptr, nullptr, 0, xb_dev, n, d, alpha_host, t, B, ysq_dev, scale_dev, ls_dev, noise_dev, weights_dev, nullptr,
                         workspace_dev, lam_dev, out_dev, want_grad, stream);
}

}  // extern "C"
