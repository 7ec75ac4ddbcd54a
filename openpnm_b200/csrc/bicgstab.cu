// bicgstab.cu -- Jacobi-scaled BiCGStab for the non-symmetric systems of
// AdvectionDiffusion.  Reference path (file:line under
// /root/reference): openpnm/algorithms/_advection_diffusion.py:43-105 builds A
// from an (Nt,2) conductance (network/_network.py:281-285), so A is not
// symmetric and the reference hands it to a direct solver
// (solvers/_scipy.py:8-15).  Here: van der Vorst's BiCGStab on
// A~ = D^-1/2 A D^-1/2, with the SpMV + three dot products of the PCG's K_A
// reused for both products of an iteration:
//   v = A~ p   with  (p.v, rhat.v, v.v)   -> alpha = rho / rhat.v
//   t = A~ s   with  (s.t, rhat.t, t.t)   -> omega = s.t / t.t
// Stopping rule as everywhere: ||b - A x||_2 <= max(rtol ||b||_2, atol) on the
// unscaled residual, confirmed with an explicitly recomputed residual.
#include "common.cuh"

#define BI_T 256

// scalar slots inside ctx->scal
enum { B_G1 = 0, B_RHO = 3, B_RHON = 4, B_G2 = 5, B_RR = 8, B_TRUE = 9 };

// s = r - alpha v
__global__ void __launch_bounds__(BI_T) k_bi_s(const double* __restrict__ r,
                                               const double* __restrict__ v,
                                               double* __restrict__ s, int64_t n,
                                               const double* __restrict__ sc) {
  const double rv = sc[B_G1 + G_RAP];
  const double alpha = (rv > 0.0 || rv < 0.0) ? sc[B_RHO] / rv : 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    s[i] = r[i] - alpha * v[i];
}

// x += alpha p + omega s ; r = s - omega t ; sums rhat.r, r.r, sum d r^2
__global__ void __launch_bounds__(BI_T) k_bi_x(double* __restrict__ x, double* __restrict__ r,
                                               const double* __restrict__ p,
                                               const double* __restrict__ s,
                                               const double* __restrict__ t,
                                               const double* __restrict__ rhat,
                                               const double* __restrict__ d, int64_t n,
                                               const double* __restrict__ sc, double* part,
                                               unsigned* ticket, double* scw) {
  const double rv = sc[B_G1 + G_RAP];
  const double alpha = (rv > 0.0 || rv < 0.0) ? sc[B_RHO] / rv : 0.0;
  const double tt = sc[B_G2 + G_APAP];
  const double omega = (tt > 0.0) ? sc[B_G2 + G_PAP] / tt : 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double a = 0.0, b = 0.0, c = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double si = s[i];
    const double rn = si - omega * t[i];
    x[i] += alpha * p[i] + omega * si;
    r[i] = rn;
    a += rhat[i] * rn;
    b += rn * rn;
    c += d[i] * rn * rn;
  }
  double v[3] = {a, b, c};
  double* o[3] = {scw + B_RHON, scw + B_RR, scw + B_TRUE};
  grid_sum_finalize<3>(v, part, ticket, o);
}

// p = r + beta (p - omega v),  beta = (rho'/rho)(alpha/omega)
__global__ void __launch_bounds__(BI_T) k_bi_p(double* __restrict__ p,
                                               const double* __restrict__ r,
                                               const double* __restrict__ v, int64_t n,
                                               const double* __restrict__ sc) {
  const double rv = sc[B_G1 + G_RAP];
  const double alpha = (rv > 0.0 || rv < 0.0) ? sc[B_RHO] / rv : 0.0;
  const double tt = sc[B_G2 + G_APAP];
  const double omega = (tt > 0.0) ? sc[B_G2 + G_PAP] / tt : 0.0;
  const double rho = sc[B_RHO];
  const double beta = ((rho > 0.0 || rho < 0.0) && (omega > 0.0 || omega < 0.0))
                          ? (sc[B_RHON] / rho) * (alpha / omega) : 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    p[i] = r[i] + beta * (p[i] - omega * v[i]);
}

int b200_bicgstab(b200_ctx* ctx, const double* x0, double rtol, double atol, int64_t maxiter,
                  double* x, int64_t* iters, double* relres, int* info) {
  // slab run (one rank per GPU): the operand of every product gets its halo windows from
  // the neighbour ranks first, every dot product is all-reduced before it is consumed, and
  // every decision is taken on all-reduced numbers (same control flow on all ranks)
  const bool multi = ctx->nranks > 1;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  RET(b200_pcg_prepare(ctx, ctx->force_fmt));
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  const int64_t n = ctx->n, pad = ctx->pad;
  if (maxiter <= 0) maxiter = 10 * ctx->n_global;
  const size_t vlen = (size_t)(n + 2 * pad);
  CK(ctx->vbi.ensure(3 * vlen * sizeof(double)));
  CK(cudaMemsetAsync(ctx->vbi.p, 0, 3 * vlen * sizeof(double), ctx->stream));
  double* sc = ctx->scal.as<double>();
  double* part = ctx->w_part.as<double>();
  unsigned* ticket = ctx->w_ticket.as<unsigned>();
  int* flags = ctx->w_flags.as<int>();
  double* xt = ctx->vx.as<double>() + pad;
  double* r = ctx->vr.as<double>() + pad;
  double* p = ctx->vp.as<double>() + pad;
  double* v = ctx->vap.as<double>() + pad;
  double* bt = ctx->vb.as<double>() + pad;
  double* rhat = ctx->vbi.as<double>() + pad;
  double* s = ctx->vbi.as<double>() + vlen + pad;
  double* t = ctx->vbi.as<double>() + 2 * vlen + pad;
  const double* d = b200_cur_diag(ctx);
  if (!ctx->unit_diag) {
    CK(ctx->w_tmp1.ensure((size_t)n * sizeof(double)));
    std::vector<double> ones((size_t)n, 1.0);
    CK(cudaMemcpyAsync(ctx->w_tmp1.p, ones.data(), (size_t)n * sizeof(double),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    d = ctx->w_tmp1.as<double>();
  }
  const int g = grid_for(n, BI_T);
  CK(cudaMemsetAsync(flags, 0, FLAG_COUNT * sizeof(int), ctx->stream));
  CK(cudaMemsetAsync(sc, 0, SC_COUNT * sizeof(double), ctx->stream));
  // x_owned - pad = base of the padded vector whose halo windows are exchanged
  auto halo = [&](double* owned) { return b200_halo_exchange(ctx, owned - pad); };
  auto resid = [&]() -> int {                 // r = p = b~ - A~ x, rho = r.r, true residual norm
    RET(halo(xt));
    RET(b200_spmv_dots(ctx, xt, bt, v, sc + SC_AUX0));
    RET(b200_launch_resid(ctx, bt, v, d, r, p, sc + B_RHO, sc + B_TRUE));
    RET(b200_allreduce_sum(ctx, sc + B_RHO, 1));
    RET(b200_allreduce_sum(ctx, sc + B_TRUE, 1));
    return B200_OK;
  };
  RET(b200_launch_init(ctx, x0, xt, bt, sc + SC_BNORM));
  RET(b200_allreduce_sum(ctx, sc + SC_BNORM, 1));
  RET(resid());                                                            // rhat = r
  CK(cudaMemcpyAsync(rhat, r, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->scal_h, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaMemcpyAsync(ctx->flags_h, flags, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if ((!multi && ctx->flags_h[FLAG_NONFINITE]) || !isfinite(ctx->scal_h[SC_BNORM]) ||
      !isfinite(ctx->scal_h[B_TRUE]))
    FAIL(B200_ERR_NONFINITE, "b or x0 contains inf/nan values");
  const double bnorm2 = ctx->scal_h[SC_BNORM];
  double tol2 = rtol * rtol * bnorm2;
  if (atol * atol > tol2) tol2 = atol * atol;
  double true2 = ctx->scal_h[B_TRUE];
  int64_t it = 0;
  int inf = 0, restarts = 0;
  if (bnorm2 == 0.0) {
    CK(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (iters) *iters = 0;
    if (relres) *relres = 0.0;
    if (info) *info = 0;
    return B200_OK;
  }
  bool done = !(true2 > tol2);
  while (!done && it < maxiter) {
    RET(halo(p));
    RET(b200_spmv_dots(ctx, p, rhat, v, sc + B_G1));
    RET(b200_allreduce_sum(ctx, sc + B_G1, 3));
    LAUNCH(ctx, k_bi_s, g, BI_T, 0, r, v, s, n, sc);
    RET(halo(s));
    RET(b200_spmv_dots(ctx, s, rhat, t, sc + B_G2));
    RET(b200_allreduce_sum(ctx, sc + B_G2, 3));
    LAUNCH(ctx, k_bi_x, g, BI_T, 0, xt, r, p, s, t, rhat, d, n, sc, part, ticket, sc);
    RET(b200_allreduce_sum(ctx, sc + B_RHON, 1));
    RET(b200_allreduce_sum(ctx, sc + B_RR, 2));
    LAUNCH(ctx, k_bi_p, g, BI_T, 0, p, r, v, n, sc);
    CK(cudaMemcpyAsync(ctx->scal_h, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaMemcpyAsync(sc + B_RHO, sc + B_RHON, sizeof(double), cudaMemcpyDeviceToDevice,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ++it;
    true2 = ctx->scal_h[B_TRUE];
    const double rhat_v = ctx->scal_h[B_G1 + G_RAP], tt = ctx->scal_h[B_G2 + G_APAP],
                 rho_new = ctx->scal_h[B_RHON];
    if (ctx->debug)
      fprintf(stderr, "[b200 bicgstab] it=%lld true2=%.3e tol2=%.3e rho'=%.3e\n", (long long)it,
              true2, tol2, rho_new);
    if (!isfinite(true2) || !isfinite(rho_new)) { inf = -1; break; }
    if (true2 <= tol2) {
      RET(resid());
      CK(cudaMemcpyAsync(rhat, r, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice,
                         ctx->stream));
      CK(cudaMemcpyAsync(ctx->scal_h, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost,
                         ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      true2 = ctx->scal_h[B_TRUE];
      if (true2 <= tol2 * (1.0 + 1e-6) || restarts >= 4) { done = true; break; }
      ++restarts;              // restarted from the true residual (p = rhat = r)
      continue;
    }
    if (rhat_v == 0.0 || tt == 0.0 || rho_new == 0.0) {
      // breakdown: restart from the current residual with a fresh shadow vector
      if (restarts >= 8) { inf = -1; break; }
      ++restarts;
      RET(resid());
      CK(cudaMemcpyAsync(rhat, r, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice,
                         ctx->stream));
    }
  }
  RET(b200_launch_final(ctx, xt, x));
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->ev2, ctx->stream));
  CK(cudaEventSynchronize(ctx->ev2));
  float ms_setup = 0.f, ms_solve = 0.f;
  cudaEventElapsedTime(&ms_setup, ctx->ev0, ctx->ev1);
  cudaEventElapsedTime(&ms_solve, ctx->ev1, ctx->ev2);
  ctx->stats.t_setup_ms = ms_setup;
  ctx->stats.t_solve_ms = ms_solve;
  ctx->stats.cg_iters = it;
  ctx->stats.n = n;
  ctx->stats.nnz_system = ctx->sys_nnz;
  const double rel = sqrt(true2 / bnorm2);
  ctx->stats.relres = rel;
  if (!done && inf == 0) inf = (int)(it > 0x7fffffff ? 0x7fffffff : (it > 0 ? it : 1));
  if (iters) *iters = it;
  if (relres) *relres = rel;
  if (info) *info = inf;
  return B200_OK;
}
