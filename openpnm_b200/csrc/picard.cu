// picard.cu -- K5 (device Picard / Newton-linearised outer loop) and K6 (rate).
//
// Reference behaviour replaced (file:line under /root/reference):
//   openpnm/algorithms/_reactive_transport.py:150-213  _run_special outer loop
//   openpnm/algorithms/_reactive_transport.py:234-244  _get_residual = A x - b
//   openpnm/algorithms/_transport.py:209-221           solve + relaxation
//   openpnm/algorithms/_solution.py:17-20              first-dx aliasing quirk
//   scipy/optimize/_nonlin.py TerminationCondition.check (f_tol = x_tol = inf)
//   openpnm/algorithms/_transport.py:259-330           rate()
#include "common.cuh"

extern "C" int b200_clear_sources(b200_ctx* ctx) {
  if (!ctx) return B200_ERR_ARG;
  ctx->srcs.n = 0;
  ctx->lin_valid = false;
  ctx->op_epoch = ~0ull;
  return B200_OK;
}

extern "C" int b200_add_source(b200_ctx* ctx, int kind, const uint8_t* mask, const double* p1,
                               const double* p2, const double* p3) {
  if (!ctx) return B200_ERR_ARG;
  if (ctx->srcs.n >= B200_MAX_SRC) FAIL(B200_ERR_ARG, "at most %d source terms", B200_MAX_SRC);
  if (kind < B200_SRC_LINEAR || kind > B200_SRC_STANDARD_KINETICS || !mask || !p1 || !p2)
    FAIL(B200_ERR_ARG, "add_source: bad arguments");
  SrcTerm& s = ctx->srcs.s[ctx->srcs.n++];
  s.kind = kind; s.mask = mask; s.p1 = p1; s.p2 = p2; s.p3 = p3;
  ctx->lin_valid = false;
  ctx->op_epoch = ~0ull;
  return B200_OK;
}

// x <- w x_new + (1-w) x ; returns ||x - xold||^2 (xold = x before the update)
__global__ void __launch_bounds__(256) k_relax(double* __restrict__ x,
                                               const double* __restrict__ xnew, double w,
                                               int64_t n, double* part, unsigned* ticket,
                                               double* out_dx2, int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double xo = x[i];
    const double xn = w * xnew[i] + (1.0 - w) * xo;
    x[i] = xn;
    const double dx = xn - xo;
    s += dx * dx;
    if (!isfinite(xn)) bad = true;
  }
  if (bad) flags[FLAG_NONFINITE] = 1;
  double v[1] = {s};
  double* o[1] = {out_dx2};
  grid_sum_finalize<1>(v, part, ticket, o);
}

__global__ void k_check_finite(const double* __restrict__ a, int64_t n, int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (!isfinite(a[i])) bad = true;
  if (bad) flags[FLAG_NONFINITE] = 1;
}

// _update_iterative_props + _update_A_and_b through the caller's hook: the hook
// may re-assemble the system (b200_build_laplacian, b200_apply_bcs) and replace
// the sources, so every workspace pointer is fetched again afterwards.
static int picard_update(b200_ctx* ctx, b200_update_fn update, void* user, const double* x,
                         int64_t n) {
  if (update) {
    CK(cudaStreamSynchronize(ctx->stream));
    if (update(user, x) != 0) {
      if (ctx->err.empty() || ctx->err.find("update hook") == std::string::npos)
        ctx->err = "picard: update hook failed: " + ctx->err;
      return B200_ERR_ARG;
    }
    CK(cudaSetDevice(ctx->device));
    if (!ctx->have_sys || ctx->n != n)
      FAIL(B200_ERR_ARG, "picard: update hook left no system of the same size");
    CK(ctx->w_tmp0.ensure((size_t)n * sizeof(double)));
    CK(ctx->w_part.ensure((size_t)4 * 148 * 16 * sizeof(double)));
  }
  if (ctx->gm.active) {
    // conductance model on the device: g(x) -> K1 -> K2, no host round trip.  The
    // sources registered on the context survive (they live outside the system matrix).
    RET(b200_update_conductance(ctx, x));
    if (ctx->n != n) FAIL(B200_ERR_ARG, "picard: the conductance model changed the system size");
    CK(ctx->w_tmp0.ensure((size_t)n * sizeof(double)));
    CK(ctx->w_part.ensure((size_t)4 * 148 * 16 * sizeof(double)));
  }
  return b200_linearize(ctx, x);
}

extern "C" int b200_picard(b200_ctx* ctx, const double* x0, double relaxation,
                           int64_t newton_maxiter, double f_rtol, double x_rtol, double cg_rtol,
                           int64_t cg_maxiter, double* x, int64_t* num_iter, int* converged,
                           int64_t* cg_iters_total, int* last_info) {
  return b200_picard_cb(ctx, x0, relaxation, newton_maxiter, f_rtol, x_rtol, cg_rtol, cg_maxiter,
                        x, num_iter, converged, cg_iters_total, last_info, nullptr, nullptr);
}

extern "C" int b200_picard_cb(b200_ctx* ctx, const double* x0, double relaxation,
                              int64_t newton_maxiter, double f_rtol, double x_rtol,
                              double cg_rtol, int64_t cg_maxiter, double* x, int64_t* num_iter,
                              int* converged, int64_t* cg_iters_total, int* last_info,
                              b200_update_fn update, void* user) {
  if (!ctx || !ctx->have_sys || !x) {
    if (ctx) ctx->err = "picard: apply_bcs / load a system first";
    return B200_ERR_ARG;
  }
  if (update && ctx->nranks > 1)
    FAIL(B200_ERR_ARG, "picard: the update hook is single-GPU (no slab mode)");
  CK(cudaSetDevice(ctx->device));
  const int64_t n = ctx->n;
  CK(ctx->w_part.ensure((size_t)4 * 148 * 16 * sizeof(double)));
  if (!ctx->w_ticket.p) {
    CK(ctx->w_ticket.ensure(64));
    CK(cudaMemsetAsync(ctx->w_ticket.p, 0, 64, ctx->stream));
  }
  CK(ctx->scal.ensure(SC_COUNT * sizeof(double)));
  CK(ctx->w_tmp0.ensure((size_t)n * sizeof(double)));   // x_new
  RET(b200_zero_flags(ctx));
  int* flags = ctx->w_flags.as<int>();
  if (x0) {
    if (x0 != x)
      CK(cudaMemcpyAsync(x, x0, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    // _validate_x0 (_transport.py:229-233)
    LAUNCH(ctx, k_check_finite, grid_for(n, 256), 256, 0, x, n, flags);
    CK(cudaMemcpyAsync(ctx->flags_h, flags, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    bool bad = ctx->flags_h[FLAG_NONFINITE] != 0;
    if (ctx->nranks > 1) {            // same verdict on every rank
      double* sc = ctx->scal.as<double>();
      ctx->scal_h[SC_AUX2] = bad ? 1.0 : 0.0;
      CK(cudaMemcpyAsync(sc + SC_AUX2, ctx->scal_h + SC_AUX2, sizeof(double),
                         cudaMemcpyHostToDevice, ctx->stream));
      RET(b200_allreduce_sum(ctx, sc + SC_AUX2, 1));
      CK(cudaMemcpyAsync(ctx->scal_h + SC_AUX2, sc + SC_AUX2, sizeof(double),
                         cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      bad = ctx->scal_h[SC_AUX2] > 0.0;
    }
    if (bad) FAIL(B200_ERR_NONFINITE, "x0 contains inf/nan values");
  } else {
    CK(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), ctx->stream));
  }
  RET(picard_update(ctx, update, user, x, n));

  double f0 = -1.0, dx2 = 0.0;
  int64_t iters = 0, cg_total = 0;
  int conv = 0, info = 0;
  bool first = true;
  for (int64_t i = 0; i < newton_maxiter; ++i) {
    double r2 = 0.0, x2 = 0.0;
    RET(b200_residual_norms(ctx, x, &r2, &x2));
    const double fn = sqrt(r2), xn = sqrt(x2), dxn = sqrt(dx2);
    if (f0 < 0.0) f0 = fn;
    if (ctx->debug)
      fprintf(stderr, "[b200 picard] i=%lld |R|=%.6e |R0|=%.6e |dx|=%.6e |x|=%.6e\n",
              (long long)i, fn, f0, dxn, xn);
    // TerminationCondition.check: f_tol = inf, x_tol = inf
    if (fn == 0.0 || (fn / f_rtol <= f0 && dxn / x_rtol <= xn)) { conv = 1; break; }
    // _validate_linear_system (_transport.py:253-257)
    CK(cudaMemcpyAsync(ctx->flags_h, flags, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // (in slab mode a local inf/nan shows up in the all-reduced norms on every rank)
    if ((ctx->nranks == 1 && ctx->flags_h[FLAG_NONFINITE]) || !isfinite(fn))
      FAIL(B200_ERR_NONFINITE, "A or b contains inf/nan values");
    int64_t its = 0;
    double rel = 0.0;
    double* xnew = ctx->w_tmp0.as<double>();
    RET(b200_pcg(ctx, x, cg_rtol, 0.0, cg_maxiter, xnew, &its, &rel, &info));
    cg_total += its;
    double* sc = ctx->scal.as<double>();
    LAUNCH(ctx, k_relax, grid_for(n, 256, 2), 256, 0, x, xnew, relaxation, n,
           ctx->w_part.as<double>(), ctx->w_ticket.as<unsigned>(), sc + SC_AUX0, flags);
    RET(b200_allreduce_sum(ctx, sc + SC_AUX0, 1));     // slab mode: global ||dx||^2
    CK(cudaMemcpyAsync(ctx->scal_h + SC_AUX0, sc + SC_AUX0, sizeof(double), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    dx2 = first ? 0.0 : ctx->scal_h[SC_AUX0];   // aliasing quirk: first dx == 0
    first = false;
    RET(picard_update(ctx, update, user, x, n));
    flags = ctx->w_flags.as<int>();
    iters = i + 1;
  }
  CK(cudaStreamSynchronize(ctx->stream));
  if (num_iter) *num_iter = iters;
  if (converged) *converged = conv;
  if (cg_iters_total) *cg_iters_total = cg_total;
  if (last_info) *last_info = info;
  return B200_OK;
}

// ==================================================================== K6: rate
// Net rate leaving pore i = sum_t g_t (x_i - x_nbr) = (L_pure x)_i: one row of
// the pure CSR per requested pore, 8 lanes per row.
__global__ void __launch_bounds__(256) k_rate_pores(const int* __restrict__ indptr,
                                                    const int* __restrict__ indices,
                                                    const double* __restrict__ data,
                                                    const double* __restrict__ x,
                                                    const long long* __restrict__ pores, int64_t k,
                                                    int64_t rb, int64_t n, double* __restrict__ out,
                                                    int* __restrict__ flags) {
  constexpr int G = 8;
  const int lane = threadIdx.x % G;
  const int64_t gstride = ((int64_t)gridDim.x * blockDim.x) / G;
  const int64_t kround = (k + (32 / G) - 1) / (32 / G) * (32 / G);   // warp-uniform trips
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; q < kround; q += gstride) {
    double s = 0.0;
    if (q < k) {
      const int64_t r = pores[q] - rb;
      if (r < 0 || r >= n) {
        if (lane == 0) flags[FLAG_INDEX] = 1;
      } else {
        for (int e = indptr[r] + lane; e < indptr[r + 1]; e += G) s += data[e] * x[indices[e]];
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, G);
    if (lane == 0 && q < k) out[q] = s;
  }
}

extern "C" int b200_rate_pores(b200_ctx* ctx, const double* x, const int64_t* pores, int64_t k,
                               double* out) {
  if (!ctx || !ctx->have_pure) {
    if (ctx) ctx->err = "rate: build_laplacian first";
    return B200_ERR_ARG;
  }
  if (k <= 0) return B200_OK;
  RET(b200_zero_flags(ctx));
  LAUNCH(ctx, k_rate_pores, grid_for(k * 8, 256), 256, 0, ctx->pure_indptr.as<int>(),
         ctx->pure_indices.as<int>(), ctx->pure_data.as<double>(), x,
         reinterpret_cast<const long long*>(pores), k, ctx->row_begin, ctx->n, out,
         ctx->w_flags.as<int>());
  CK(cudaGetLastError());
  return b200_check_flags(ctx);
}

__global__ void k_rate_throats(const longlong2* __restrict__ conns, const double* __restrict__ g,
                               const double* __restrict__ x, const long long* __restrict__ throats,
                               int64_t k, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < k; q += stride) {
    const long long t = throats[q];
    const longlong2 c = conns[t];
    out[q] = fabs(g[t] * (x[c.y] - x[c.x]));
  }
}

extern "C" int b200_rate_throats(b200_ctx* ctx, const int64_t* conns, const double* g,
                                 const double* x, const int64_t* throats, int64_t k, double* out) {
  if (!ctx) return B200_ERR_ARG;
  if (k <= 0) return B200_OK;
  LAUNCH(ctx, k_rate_throats, grid_for(k, 256), 256, 0, reinterpret_cast<const longlong2*>(conns),
         g, x, reinterpret_cast<const long long*>(throats), k, out);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}


// Two-way (Nt,2) conductance (AdvectionDiffusion): |g[t][0] x[c1] - g[t][1] x[c0]|
// (_transport.py:306-319 with the flipped (Nt,2) g of network/_network.py:281-285)
__global__ void k_rate_throats2(const longlong2* __restrict__ conns, const double* __restrict__ g2,
                                const double* __restrict__ x, const long long* __restrict__ throats,
                                int64_t k, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < k; q += stride) {
    const long long t = throats[q];
    const longlong2 c = conns[t];
    // two rounded products, then the difference -- as numpy evaluates it; a fused
    // multiply-add would round differently where the two fluxes nearly cancel
    out[q] = fabs(__dsub_rn(__dmul_rn(g2[2 * t], x[c.y]), __dmul_rn(g2[2 * t + 1], x[c.x])));
  }
}

extern "C" int b200_rate_throats2(b200_ctx* ctx, const int64_t* conns, const double* g2,
                                  const double* x, const int64_t* throats, int64_t k, double* out) {
  if (!ctx) return B200_ERR_ARG;
  if (k <= 0) return B200_OK;
  LAUNCH(ctx, k_rate_throats2, grid_for(k, 256), 256, 0,
         reinterpret_cast<const longlong2*>(conns), g2, x,
         reinterpret_cast<const long long*>(throats), k, out);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

// S1, S2 and the rate r of one registered source term at x, for the masked pores (0
// elsewhere): what `regenerate_models` leaves on the phase after the run
// (_algorithm.py:89-91; source_terms/_funcs.py:61-69,108-116,158-168).
__global__ void k_source_values(SrcTerm s, const double* __restrict__ x, int64_t n,
                                double* __restrict__ S1o, double* __restrict__ S2o,
                                double* __restrict__ ro) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double X = x[i], p1 = s.p1[i], p2 = s.p2[i], p3 = s.p3 ? s.p3[i] : 0.0;
    double S1, S2, r;
    if (s.kind == B200_SRC_LINEAR) {
      S1 = p1; S2 = p2; r = p1 * X + p2;
    } else if (s.kind == B200_SRC_POWER_LAW) {
      const double xb = pow(X, p2);
      r = p1 * xb + p3;
      S1 = p1 * p2 * pow(X, p2 - 1.0);
      S2 = p1 * xb * (1.0 - p2) + p3;
    } else {
      const double xb = pow(X, p2);
      r = p1 * xb;
      S1 = p1 * p2 * pow(X, p2 - 1.0);
      S2 = p1 * (1.0 - p2) * xb;
    }
    S1o[i] = S1; S2o[i] = S2; ro[i] = r;
  }
}

extern "C" int b200_source_values(b200_ctx* ctx, int index, const double* x, double* S1,
                                  double* S2, double* rate) {
  if (!ctx || !x || !S1 || !S2 || !rate) return B200_ERR_ARG;
  if (index < 0 || index >= ctx->srcs.n) FAIL(B200_ERR_ARG, "source_values: no such source");
  LAUNCH(ctx, k_source_values, grid_for(ctx->n, 256), 256, 0, ctx->srcs.s[index], x, ctx->n, S1, S2,
         rate);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
