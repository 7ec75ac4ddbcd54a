// transient.cu -- explicit time integration of  dy/dt = (-A y + b) / V  on the
// device.  Reference path (file:line under
// /root/reference): openpnm/algorithms/_transient_reactive_transport.py:51-123
// (`run`, `_build_rhs`: A and b are re-assembled at every right-hand-side call
// with the sources linearised at y) and openpnm/integrators/_scipy.py:18-56
// (`ScipyRK45` -> scipy.integrate.solve_ivp(method="RK45", t_eval=saveat)).
//
// The integrator is a restatement of SciPy's RK45 (scipy/integrate/_ivp/rk.py:
// Dormand-Prince 5(4) tableau, `select_initial_step`, the accept/reject step
// controller of `RungeKutta._step_impl`, the quartic dense output of
// `RkDenseOutput`) and of the t_eval bookkeeping of solve_ivp
// (scipy/integrate/_ivp/ivp.py), so that the sequence of steps is the same.
//
// Right-hand side without re-assembly: with A0, b0 the system after BCs only,
//   -A y + b = -A0 y + b0 + sum_src (S1 y + S2)   on the source pores,
// and A0 y = S^-1 A~ (S^-1 y) is applied through the prepared PCG operator
// (K_A: DIA-sym or CSR-stream), so the matrix traffic is the PCG's.
#include <math.h>

#include "common.cuh"

#define TR_T 256

struct KPack {
  const double* k[7];
};
struct Coef7 {
  double c[7];
};

// Dormand-Prince coefficients as used by scipy.integrate.RK45
// (the c_i nodes are not needed: the right-hand side is autonomous)
static const double RK_A[6][5] = {
    {0.0, 0.0, 0.0, 0.0, 0.0},
    {0.2, 0.0, 0.0, 0.0, 0.0},
    {0.075, 0.225, 0.0, 0.0, 0.0},
    {0.9777777777777777, -3.7333333333333334, 3.5555555555555554, 0.0, 0.0},
    {2.9525986892242035, -11.595793324188385, 9.822892851699436, -0.2908093278463649, 0.0},
    {2.8462752525252526, -10.757575757575758, 8.906422717743473, 0.2784090909090909,
     -0.2735313036020583}};
static const double RK_B[6] = {0.09114583333333333, 0.0, 0.44923629829290207, 0.6510416666666666,
                               -0.322376179245283, 0.13095238095238096};
static const double RK_E[7] = {-0.0012326388888888888, 0.0, 0.0042527702905061394,
                               -0.03697916666666667, 0.05086379716981132, -0.0419047619047619,
                               0.025};
static const double RK_P[7][4] = {
    {1.0, -2.8535800653862835, 3.0717434641059005, -1.1270175653862835},
    {0.0, 0.0, 0.0, 0.0},
    {0.0, 4.023133379230305, -6.249321565289, 2.675424484351598},
    {0.0, -3.7324019615885042, 10.068970589843675, -5.685526961588504},
    {0.0, 2.5548038301849423, -6.399112377351017, 3.5219323679207912},
    {0.0, -1.3744241142186024, 3.272657752246729, -1.7672812570757455},
    {0.0, 1.3824689317781436, -3.764937863556287, 2.382468931778144}};

struct P74 {
  double p[7][4];
};

// w = y / s  (into the padded operand vector of K_A)
__global__ void __launch_bounds__(TR_T) k_tr_prescale(const double* __restrict__ y,
                                                      const double* __restrict__ s,
                                                      double* __restrict__ w, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    w[i] = y[i] / s[i];
}

__device__ __forceinline__ void tr_src(int kind, double X, double p1, double p2, double p3,
                                       double& S1, double& S2) {
  if (kind == B200_SRC_LINEAR) {
    S1 = p1; S2 = p2;
  } else if (kind == B200_SRC_POWER_LAW) {
    const double xb = pow(X, p2);
    S1 = p1 * p2 * pow(X, p2 - 1.0);
    S2 = p1 * xb * (1.0 - p2) + p3;
  } else {
    const double xb = pow(X, p2);
    S1 = p1 * p2 * pow(X, p2 - 1.0);
    S2 = p1 * (1.0 - p2) * xb;
  }
}

// f = (-(A~ w)/s + b0 + sum_src (S1 y + S2)) / V
__global__ void __launch_bounds__(TR_T) k_tr_rhs(const double* __restrict__ aw,
                                                 const double* __restrict__ s,
                                                 const double* __restrict__ b0,
                                                 const double* __restrict__ y,
                                                 const double* __restrict__ vol, SrcPack sp,
                                                 double* __restrict__ f, int64_t n,
                                                 int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double yi = y[i];
    double acc = -aw[i] / s[i] + b0[i];
    for (int q = 0; q < sp.n; ++q) {
      if (sp.s[q].mask[i]) {
        double S1, S2;
        tr_src(sp.s[q].kind, yi, sp.s[q].p1[i], sp.s[q].p2[i], sp.s[q].p3 ? sp.s[q].p3[i] : 0.0,
               S1, S2);
        // diag += -S1, b += S2 (_reactive_transport.py:136-146)  =>  + S1 y + S2
        acc += S1 * yi + S2;
      }
    }
    const double v = acc / vol[i];
    if (!isfinite(v)) bad = true;
    f[i] = v;
  }
  if (bad) flags[FLAG_NONFINITE] = 1;
}

// out = y + h * sum_j c_j K_j
__global__ void __launch_bounds__(TR_T) k_tr_combine(double* __restrict__ out,
                                                     const double* __restrict__ y, double h,
                                                     int nk, Coef7 c, KPack K, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double dy = 0.0;
    for (int j = 0; j < nk; ++j) dy += K.k[j][i] * c.c[j];
    out[i] = y[i] + dy * h;
  }
}

// sum ((h * sum_j E_j K_j) / (atol + rtol max(|y|, |ynew|)))^2
__global__ void __launch_bounds__(TR_T) k_tr_error(const double* __restrict__ y,
                                                   const double* __restrict__ ynew, double h,
                                                   Coef7 e, KPack K, double rtol, double atol,
                                                   int64_t n, double* part, unsigned* ticket,
                                                   double* out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double err = 0.0;
#pragma unroll
    for (int j = 0; j < 7; ++j) err += K.k[j][i] * e.c[j];
    err *= h;
    const double scale = atol + fmax(fabs(y[i]), fabs(ynew[i])) * rtol;
    const double q = err / scale;
    s += q * q;
  }
  double v[1] = {s};
  double* o[1] = {out};
  grid_sum_finalize<1>(v, part, ticket, o);
}

// sums for select_initial_step: sum (a/scale)^2, sum (b/scale)^2,
// sum ((c - b)/scale)^2 with scale = atol + |y| rtol (c may be null)
__global__ void __launch_bounds__(TR_T) k_tr_norms(const double* __restrict__ y,
                                                   const double* __restrict__ a,
                                                   const double* __restrict__ b,
                                                   const double* __restrict__ c, double rtol,
                                                   double atol, int64_t n, double* part,
                                                   unsigned* ticket, double* out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double scale = atol + fabs(y[i]) * rtol;
    const double qa = a[i] / scale, qb = b[i] / scale;
    s0 += qa * qa;
    s1 += qb * qb;
    if (c) { const double qc = (c[i] - b[i]) / scale; s2 += qc * qc; }
  }
  double v[3] = {s0, s1, s2};
  double* o[3] = {out, out + 1, out + 2};
  grid_sum_finalize<3>(v, part, ticket, o);
}

// dense output: out = yold + h * sum_p (sum_j K_j P[j][p]) x^(p+1)
__global__ void __launch_bounds__(TR_T) k_tr_dense(double* __restrict__ out,
                                                   const double* __restrict__ yold, double h,
                                                   double x, P74 P, KPack K, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double x1 = x, x2 = x * x, x3 = x2 * x, x4 = x3 * x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const double kj = K.k[j][i];
      q0 += kj * P.p[j][0]; q1 += kj * P.p[j][1]; q2 += kj * P.p[j][2]; q3 += kj * P.p[j][3];
    }
    out[i] = h * (q0 * x1 + q1 * x2 + q2 * x3 + q3 * x4) + yold[i];
  }
}

// defined in solver.cu
int b200_apply_operator(b200_ctx* ctx, const double* w_owned, double* out_owned);

namespace {
struct Tr {
  b200_ctx* ctx;
  int64_t n, pad;
  const double* vol;
  const double* s;
  const double* b0;
  SrcPack srcs;
  int g;
  int64_t nfev = 0;
  int rhs(const double* y, double* f) {
    double* w = ctx->vp.as<double>() + pad;
    double* aw = ctx->vap.as<double>() + pad;
    LAUNCH(ctx, k_tr_prescale, g, TR_T, 0, y, s, w, n);
    RET(b200_halo_exchange(ctx, ctx->vp.as<double>()));     // slab run: the neighbours' edge rows
    RET(b200_apply_operator(ctx, w, aw));
    LAUNCH(ctx, k_tr_rhs, g, TR_T, 0, aw, s, b0, y, vol, srcs, f, n, ctx->w_flags.as<int>());
    ++nfev;
    return B200_OK;
  }
};
}  // namespace

extern "C" int b200_transient_rk45(b200_ctx* ctx, const double* x0, const double* volume,
                                   double t0, double t1, double rtol, double atol,
                                   const double* saveat_h, int64_t nsave, int64_t capacity,
                                   double* out, double* times_h, int64_t* nout,
                                   int64_t* nsteps, int64_t* nfev, int* status) {
  if (!ctx || !ctx->have_sys || !x0 || !volume || !out || !times_h || !nout || capacity < 1)
    return B200_ERR_ARG;
  // slab run (one rank per GPU): every right-hand side exchanges the halo of its operand,
  // the step controller's norms are all-reduced, so every rank takes the same steps
  const bool multi = ctx->nranks > 1;
  if (!(t1 > t0)) FAIL(B200_ERR_ARG, "transient: tspan must be increasing");
  CK(cudaSetDevice(ctx->device));
  const int64_t n = ctx->n;
  // operator of the source-free system (A0, b0); sources enter through their rate
  SrcPack srcs = ctx->srcs;
  ctx->srcs.n = 0;
  ctx->lin_valid = true;
  ctx->op_epoch = ~0ull;
  int rc = b200_pcg_prepare(ctx, ctx->force_fmt);
  ctx->srcs = srcs;
  ctx->lin_valid = false;
  if (rc != B200_OK) return rc;
  const uint64_t epoch_keep = ctx->sys_epoch;
  RET(b200_zero_flags(ctx));

  DevBuf store;
  if (store.ensure((size_t)10 * n * sizeof(double)) != cudaSuccess)
    FAIL(B200_ERR_CUDA, "transient: out of device memory");
  double* base = store.as<double>();
  double* y = base;
  double* ynew = base + n;
  double* ystage = base + 2 * n;
  double* Kbuf[7];
  for (int j = 0; j < 7; ++j) Kbuf[j] = base + (3 + j) * n;
  KPack K;
  for (int j = 0; j < 7; ++j) K.k[j] = Kbuf[j];
  P74 P;
  for (int j = 0; j < 7; ++j)
    for (int p = 0; p < 4; ++p) P.p[j][p] = RK_P[j][p];
  Coef7 E;
  for (int j = 0; j < 7; ++j) E.c[j] = RK_E[j];

  Tr tr;
  tr.ctx = ctx; tr.n = n; tr.pad = ctx->pad; tr.vol = volume;
  tr.s = ctx->sc.as<double>() + ctx->pad;
  tr.b0 = ctx->b.as<double>();
  tr.srcs = srcs;
  tr.g = grid_for(n, TR_T);
  double* sc = ctx->scal.as<double>();
  double* part = ctx->w_part.as<double>();
  unsigned* ticket = ctx->w_ticket.as<unsigned>();
  const int g = tr.g;
  auto fail = [&](int code) { store.release(); ctx->op_epoch = ~0ull; return code; };
#define TRY(call)                                  \
  do {                                             \
    int _r = (call);                               \
    if (_r != B200_OK) return fail(_r);            \
  } while (0)
#define CKT(call)                                                          \
  do {                                                                     \
    cudaError_t _e = (call);                                               \
    if (_e != cudaSuccess) {                                               \
      ctx->err = std::string(#call) + " -> " + cudaGetErrorString(_e);     \
      return fail(B200_ERR_CUDA);                                          \
    }                                                                      \
  } while (0)

  CKT(cudaMemcpyAsync(y, x0, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  TRY(tr.rhs(y, Kbuf[0]));                       // self.f = fun(t0, y0)
  double t = t0;
  const double sqn = sqrt((double)(multi ? ctx->n_global : n));

  // ---- select_initial_step (scipy/integrate/_ivp/common.py)
  double h_abs;
  {
    LAUNCH(ctx, k_tr_norms, g, TR_T, 0, y, y, Kbuf[0], (const double*)nullptr, rtol, atol, n, part,
           ticket, sc + SC_AUX0);
    TRY(b200_allreduce_sum(ctx, sc + SC_AUX0, 3));
    CKT(cudaMemcpyAsync(ctx->scal_h + SC_AUX0, sc + SC_AUX0, 3 * sizeof(double),
                        cudaMemcpyDeviceToHost, ctx->stream));
    CKT(cudaStreamSynchronize(ctx->stream));
    const double d0 = sqrt(ctx->scal_h[SC_AUX0]) / sqn, d1 = sqrt(ctx->scal_h[SC_AUX0 + 1]) / sqn;
    const double interval = fabs(t1 - t0);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    if (h0 > interval) h0 = interval;
    Coef7 c1{};
    c1.c[0] = 1.0;
    LAUNCH(ctx, k_tr_combine, g, TR_T, 0, ystage, y, h0, 1, c1, K, n);     // y1 = y0 + h0 f0
    TRY(tr.rhs(ystage, Kbuf[1]));
    LAUNCH(ctx, k_tr_norms, g, TR_T, 0, y, y, Kbuf[0], Kbuf[1], rtol, atol, n, part, ticket,
           sc + SC_AUX0);
    TRY(b200_allreduce_sum(ctx, sc + SC_AUX0, 3));
    CKT(cudaMemcpyAsync(ctx->scal_h + SC_AUX0, sc + SC_AUX0, 3 * sizeof(double),
                        cudaMemcpyDeviceToHost, ctx->stream));
    CKT(cudaStreamSynchronize(ctx->stream));
    const double d2 = sqrt(ctx->scal_h[SC_AUX0 + 2]) / sqn / h0;
    double h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
    else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);          // order = 4 -> 1/(order+1)
    h_abs = fmin(fmin(100.0 * h0, h1), interval);
  }

  int64_t steps = 0, n_out = 0, i_eval = 0;
  int st = 0;
  const bool every_step = (nsave == 0);
  if (every_step) {                                 // ts = [t0], ys = [y0]
    CKT(cudaMemcpyAsync(out, y, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    times_h[n_out++] = t0;
  }
  const double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0, EXPO = -1.0 / 5.0;
  while (t < t1) {
    // ---- RungeKutta._step_impl
    const double min_step = 10.0 * fabs(nextafter(t, INFINITY) - t);
    if (h_abs < min_step) h_abs = min_step;
    bool accepted = false, rejected = false;
    double h = 0.0, t_new = t;
    while (!accepted) {
      if (h_abs < min_step) { st = -1; break; }
      h = h_abs;
      t_new = t + h;
      if (t_new - t1 > 0) t_new = t1;
      h = t_new - t;
      h_abs = fabs(h);
      // rk_step: K[0] = f already
      for (int sidx = 1; sidx < 6; ++sidx) {
        Coef7 a{};
        for (int j = 0; j < sidx; ++j) a.c[j] = RK_A[sidx][j];
        LAUNCH(ctx, k_tr_combine, g, TR_T, 0, ystage, y, h, sidx, a, K, n);
        TRY(tr.rhs(ystage, Kbuf[sidx]));
      }
      Coef7 bcoef{};
      for (int j = 0; j < 6; ++j) bcoef.c[j] = RK_B[j];
      LAUNCH(ctx, k_tr_combine, g, TR_T, 0, ynew, y, h, 6, bcoef, K, n);
      TRY(tr.rhs(ynew, Kbuf[6]));
      LAUNCH(ctx, k_tr_error, g, TR_T, 0, y, ynew, h, E, K, rtol, atol, n, part, ticket,
             sc + SC_AUX0);
      TRY(b200_allreduce_sum(ctx, sc + SC_AUX0, 1));
      CKT(cudaMemcpyAsync(ctx->scal_h + SC_AUX0, sc + SC_AUX0, sizeof(double),
                          cudaMemcpyDeviceToHost, ctx->stream));
      CKT(cudaMemcpyAsync(ctx->flags_h, ctx->w_flags.p, FLAG_COUNT * sizeof(int),
                          cudaMemcpyDeviceToHost, ctx->stream));
      CKT(cudaStreamSynchronize(ctx->stream));
      // (slab run: a rank-local inf/nan makes the all-reduced norm non-finite on every rank)
      if ((!multi && ctx->flags_h[FLAG_NONFINITE]) || !isfinite(ctx->scal_h[SC_AUX0])) {
        ctx->err = "transient: the right-hand side produced inf/nan";
        return fail(B200_ERR_NONFINITE);
      }
      const double error_norm = sqrt(ctx->scal_h[SC_AUX0]) / sqn;
      if (error_norm < 1.0) {
        double factor = (error_norm == 0.0) ? MAX_FACTOR
                                            : fmin(MAX_FACTOR, SAFETY * pow(error_norm, EXPO));
        if (rejected) factor = fmin(1.0, factor);
        h_abs *= factor;
        accepted = true;
      } else {
        h_abs *= fmax(MIN_FACTOR, SAFETY * pow(error_norm, EXPO));
        rejected = true;
      }
    }
    if (st != 0) break;
    ++steps;
    // ---- solve_ivp bookkeeping: y_old = y, K holds the stages of this step
    if (every_step) {
      if (n_out >= capacity) { st = -2; break; }
      CKT(cudaMemcpyAsync(out + n_out * n, ynew, (size_t)n * sizeof(double),
                          cudaMemcpyDeviceToDevice, ctx->stream));
      times_h[n_out++] = t_new;
    } else {
      // t_eval points in (previous i_eval .. searchsorted(t_eval, t_new, 'right'))
      while (i_eval < nsave && saveat_h[i_eval] <= t_new) {
        if (n_out >= capacity) { st = -2; break; }
        const double x = (saveat_h[i_eval] - t) / h;
        LAUNCH(ctx, k_tr_dense, g, TR_T, 0, out + n_out * n, y, h, x, P, K, n);
        times_h[n_out++] = saveat_h[i_eval];
        ++i_eval;
      }
      if (st != 0) break;
    }
    // advance: y <- ynew, f <- f_new (FSAL)
    double* tmpy = y; y = ynew; ynew = tmpy;
    double* tmpk = Kbuf[0]; Kbuf[0] = Kbuf[6]; Kbuf[6] = tmpk;
    K.k[0] = Kbuf[0]; K.k[6] = Kbuf[6];
    t = t_new;
  }
  CKT(cudaGetLastError());
  CKT(cudaStreamSynchronize(ctx->stream));
  store.release();
  (void)epoch_keep;
  ctx->op_epoch = ~0ull;          // the PCG must rebuild its operator (sources change the diagonal)
  *nout = n_out;
  if (nsteps) *nsteps = steps;
  if (nfev) *nfev = tr.nfev;
  if (status) *status = st;
#undef TRY
#undef CKT
  return B200_OK;
}
