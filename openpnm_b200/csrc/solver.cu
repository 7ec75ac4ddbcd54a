// solver.cu -- K3 (SpMV), K4 (Jacobi-PCG on the symmetrically scaled operator),
// source linearisation (part of K5).  sm_100a, fp64, HBM-bound kernels.
//
// Reference behaviour replaced (file:line under /root/reference):
//   openpnm/solvers/_scipy.py:18-26   ScipyCG.solve: cg(A, b, tol, atol=|b|*tol)
//   openpnm/solvers/_base.py:21-63    IterativeSolver(tol, maxiter), residual
//   openpnm/algorithms/_reactive_transport.py:125-148  _apply_sources
//   openpnm/models/physics/source_terms/_funcs.py:21-168 linearised sources
//
// Operator formats (the PCG never touches the canonical CSR again):
//   DIA  : the matrix has <= 13 distinct super-diagonals (any lattice network in
//          index order).  Symmetric half storage: one fp64 array per
//          super-diagonal, no indices, unit main diagonal implied.
//          y_i = x_i + sum_m v_m[i] x[i+k_m] + v_m[i-k_m] x[i-k_m]
//   CSR  : everything else (Delaunay/Voronoi): off-diagonals only, int32
//          local column + fp64 value, streamed coalesced through shared memory
//          in fixed-size row blocks ("CSR-stream"), unit diagonal implied.
// Vectors live in a padded layout [pad | n owned | pad] so that neighbour
// reads i +/- k never need a bounds check; in slab mode the pads are the
// halo windows filled by the neighbour ranks.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include <algorithm>

#define PCG_T 256

static inline double* owned(const DevBuf& b, int64_t pad) { return b.as<double>() + pad; }

static inline const double* cur_diag(b200_ctx* ctx) {
  return (ctx->srcs.n > 0 && ctx->lin_valid) ? ctx->diag_eff.as<double>() : ctx->sys_diag.as<double>();
}
static inline const double* cur_b(b200_ctx* ctx) {
  return (ctx->srcs.n > 0 && ctx->lin_valid) ? ctx->b_eff.as<double>() : ctx->b.as<double>();
}

static int ensure_ws(b200_ctx* ctx) {
  CK(ctx->w_part.ensure((size_t)4 * 148 * 16 * sizeof(double)));
  if (!ctx->w_ticket.p) {
    CK(ctx->w_ticket.ensure(64));
    CK(cudaMemsetAsync(ctx->w_ticket.p, 0, 64, ctx->stream));
  }
  CK(ctx->scal.ensure(SC_COUNT * sizeof(double)));
  CK(ctx->w_flags.ensure(FLAG_COUNT * sizeof(int)));
  return B200_OK;
}

// ======================================================== raw CSR SpMV (K3)
// y = M x on a canonical CSR (diagonal stored).  8 lanes per row; used for
// b200_spmv, residuals and rate(), not inside the PCG loop.
template <int G>
__global__ void __launch_bounds__(256) k_spmv_csr_vec(const int* __restrict__ indptr,
                                                      const int* __restrict__ indices,
                                                      const double* __restrict__ data, int64_t n,
                                                      int64_t col_shift,
                                                      const double* __restrict__ x,
                                                      double* __restrict__ y) {
  const int lane = threadIdx.x % G;
  const int64_t gstride = ((int64_t)gridDim.x * blockDim.x) / G;
  // the trip count is warp-uniform (rows rounded up to a whole warp of groups):
  // every lane reaches the full-mask shuffles, idle groups just add zeros
  const int64_t nround = (n + (32 / G) - 1) / (32 / G) * (32 / G);
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; r < nround; r += gstride) {
    double s = 0.0;
    if (r < n) {
      const int e1 = indptr[r + 1];
      for (int e = indptr[r] + lane; e < e1; e += G) s += data[e] * x[indices[e] - col_shift];
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, G);
    if (lane == 0 && r < n) y[r] = s;
  }
}

int b200_spmv_csr_raw(b200_ctx* ctx, const int* indptr, const int* indices, const double* data,
                      int64_t n, int64_t col_shift, const double* x, double* y) {
  LAUNCH(ctx, k_spmv_csr_vec<8>, grid_for(n * 8, 256), 256, 0, indptr, indices, data, n,
         col_shift, x, y);
  CK(cudaGetLastError());
  return B200_OK;
}

extern "C" int b200_spmv(b200_ctx* ctx, int which, const double* x, double* y) {
  if (!ctx) return B200_ERR_ARG;
  if (which == B200_MAT_PURE) {
    if (!ctx->have_pure) FAIL(B200_ERR_ARG, "no pure matrix");
    return b200_spmv_csr_raw(ctx, ctx->pure_indptr.as<int>(), ctx->pure_indices.as<int>(),
                             ctx->pure_data.as<double>(), ctx->n, 0, x, y);
  }
  if (!ctx->have_sys) FAIL(B200_ERR_ARG, "no system matrix");
  return b200_spmv_csr_raw(ctx, ctx->sys_indptr.as<int>(), ctx->sys_indices.as<int>(),
                           ctx->sys_data.as<double>(), ctx->n, 0, x, y);
}

// =============================================== source linearisation (K5a)
// diag_eff = diag - sum_s S1_s, b_eff = b + sum_s S2_s on the masked pores,
// S1/S2 evaluated at X = x (source_terms/_funcs.py:61-69,108-116,158-168), and
// the diagonal entry of the system CSR is patched so residuals see it.
__device__ __forceinline__ void src_eval(int kind, double X, double p1, double p2, double p3,
                                         double& S1, double& S2) {
  if (kind == B200_SRC_LINEAR) {            // r = A1 X + A2
    S1 = p1;
    S2 = p2;
  } else if (kind == B200_SRC_POWER_LAW) {  // r = A X^B + C
    const double xb = pow(X, p2);
    S1 = p1 * p2 * pow(X, p2 - 1.0);
    S2 = p1 * xb * (1.0 - p2) + p3;
  } else {                                  // standard kinetics r = A X^b
    const double xb = pow(X, p2);
    S1 = p1 * p2 * pow(X, p2 - 1.0);
    S2 = p1 * (1.0 - p2) * xb;
  }
}

__global__ void k_linearize(SrcPack sp, const double* __restrict__ x,
                            const double* __restrict__ diag0, const double* __restrict__ b0,
                            const int* __restrict__ diagpos, int64_t n,
                            double* __restrict__ diag_eff, double* __restrict__ b_eff,
                            double* __restrict__ sys_data, int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double d = diag0[i], bb = b0[i];
    const double X = x[i];
    for (int s = 0; s < sp.n; ++s) {
      if (sp.s[s].mask[i]) {
        double S1, S2;
        src_eval(sp.s[s].kind, X, sp.s[s].p1[i], sp.s[s].p2[i], sp.s[s].p3 ? sp.s[s].p3[i] : 0.0,
                 S1, S2);
        d += -S1;
        bb += S2;
      }
    }
    if (!isfinite(d) || !isfinite(bb)) bad = true;
    diag_eff[i] = d;
    b_eff[i] = bb;
    const int dp = diagpos[i];
    if (dp >= 0) sys_data[dp] = d;
  }
  if (bad) flags[FLAG_NONFINITE] = 1;
}

int b200_linearize(b200_ctx* ctx, const double* x) {
  if (!ctx->have_sys) FAIL(B200_ERR_ARG, "no system matrix");
  if (ctx->srcs.n == 0) {
    ctx->lin_valid = true;
    return B200_OK;
  }
  const int64_t n = ctx->n;
  RET(ensure_ws(ctx));
  CK(ctx->diag_eff.ensure((size_t)n * sizeof(double)));
  CK(ctx->b_eff.ensure((size_t)n * sizeof(double)));
  LAUNCH(ctx, k_linearize, grid_for(n, 256), 256, 0, ctx->srcs, x, ctx->sys_diag.as<double>(),
         ctx->b.as<double>(), ctx->sys_diagpos.as<int>(), n, ctx->diag_eff.as<double>(),
         ctx->b_eff.as<double>(), ctx->sys_data.as<double>(), ctx->w_flags.as<int>());
  CK(cudaGetLastError());
  ctx->lin_valid = true;
  ctx->lin_epoch++;
  ctx->op_epoch = ~0ull;   // scaling depends on the diagonal
  return B200_OK;
}

// residual R = A x - b with the current linearisation; returns ||R||^2, ||x||^2.
// xcol is indexed by (global) column, xrow by local row -- the same pointer on
// one GPU; in slab mode xcol points into a halo-exchanged padded copy of x.
__global__ void __launch_bounds__(256) k_residual_norms(const int* __restrict__ indptr,
                                                        const int* __restrict__ indices,
                                                        const double* __restrict__ data,
                                                        const double* __restrict__ b,
                                                        const double* __restrict__ xcol,
                                                        const double* __restrict__ xrow, int64_t n,
                                                        double* part, unsigned* ticket,
                                                        double* out_r2, double* out_x2) {
  constexpr int G = 8;
  const int lane = threadIdx.x % G;
  const int64_t gstride = ((int64_t)gridDim.x * blockDim.x) / G;
  double r2 = 0.0, x2 = 0.0;
  const int64_t nround = (n + (32 / G) - 1) / (32 / G) * (32 / G);   // warp-uniform trips
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; r < nround; r += gstride) {
    double s = 0.0;
    if (r < n) {
      const int e1 = indptr[r + 1];
      for (int e = indptr[r] + lane; e < e1; e += G) s += data[e] * xcol[indices[e]];
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, G);
    if (lane == 0 && r < n) {
      const double res = s - b[r];
      r2 += res * res;
      x2 += xrow[r] * xrow[r];
    }
  }
  double v[2] = {r2, x2};
  double* o[2] = {out_r2, out_x2};
  grid_sum_finalize<2>(v, part, ticket, o);
}

int b200_residual_norms(b200_ctx* ctx, const double* x, double* res2, double* x2) {
  RET(ensure_ws(ctx));
  double* sc = ctx->scal.as<double>();
  const double* xcol = x;
  if (ctx->nranks > 1) {
    // slab: columns are global ids; gather from a padded copy of x whose halo
    // windows the neighbour ranks fill (the PCG's x buffer is free between solves)
    RET(b200_pcg_prepare(ctx, ctx->force_fmt));
    double* xo = owned(ctx->vx, ctx->pad);
    CK(cudaMemcpyAsync(xo, x, (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToDevice,
                       ctx->stream));
    RET(b200_halo_exchange(ctx, ctx->vx.as<double>()));
    xcol = xo - ctx->row_begin;
  }
  LAUNCH(ctx, k_residual_norms, grid_for(ctx->n * 8, 256), 256, 0, ctx->sys_indptr.as<int>(),
         ctx->sys_indices.as<int>(), ctx->sys_data.as<double>(), cur_b(ctx), xcol, x, ctx->n,
         ctx->w_part.as<double>(), ctx->w_ticket.as<unsigned>(), sc + SC_AUX0, sc + SC_AUX1);
  RET(b200_allreduce_sum(ctx, sc + SC_AUX0, 2));
  CK(cudaMemcpyAsync(ctx->scal_h + SC_AUX0, sc + SC_AUX0, 2 * sizeof(double),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *res2 = ctx->scal_h[SC_AUX0];
  *x2 = ctx->scal_h[SC_AUX1];
  return B200_OK;
}

// ============================================================ operator setup
struct OffTable {
  int n;                      // number of distinct positive offsets found
  int off[B200_MAX_DIAGS];
  int overflow;
  int maxband;
};

// distinct |col - row| of the off-diagonals of the system matrix; integer CAS only
__global__ void k_detect_offsets(const int* __restrict__ indptr, const int* __restrict__ indices,
                                 int64_t n, int64_t rb, OffTable* tab) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int mymax = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const int64_t gi = r + rb;
    // once the table has overflowed (an irregular network) only the bandwidth is
    // still wanted: skip the slot search, which would hammer one cache line
    const bool full = *reinterpret_cast<volatile int*>(&tab->overflow) != 0;
    for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
      const int64_t o64 = (int64_t)indices[e] - gi;
      const int ao = (int)(o64 < 0 ? -o64 : o64);
      if (ao > mymax) mymax = ao;
      if (o64 == 0 || full) continue;
      // |offset| of lower entries too: in a slab a lower entry may point below
      // the owned rows while no owned row has the mirrored upper entry
      const int o = ao;
      bool found = false;
      for (int k = 0; k < B200_MAX_DIAGS && !found; ++k) {
        int cur = tab->off[k];      // may be stale (L1): the CAS below is the truth
        if (cur == 0) cur = atomicCAS(&tab->off[k], 0, o);
        if (cur == 0 || cur == o) found = true;
      }
      if (!found) tab->overflow = 1;
    }
  }
  if (mymax > 0) atomicMax(&tab->maxband, mymax);
}

// s_i = d_i^-1/2 in the padded layout; singleton[i] = row has no off-diagonal
__global__ void k_scale(const double* __restrict__ diag, int64_t n, int unit,
                        double* __restrict__ s_owned, int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool nonpos = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double d = diag[i];
    if (!(d > 0.0) || !isfinite(d)) nonpos = true;
    s_owned[i] = unit ? ((d > 0.0) ? 1.0 / sqrt(d) : 1.0) : 1.0;
  }
  if (nonpos) flags[FLAG_NONPOS_DIAG] = 1;
}

__device__ __forceinline__ int off_lookup(const OffTable& t, int o) {
#pragma unroll
  for (int k = 0; k < B200_MAX_DIAGS; ++k)
    if (k < t.n && t.off[k] == o) return k;
  return -1;
}

// scatter the scaled upper triangle into the diagonal arrays; lower entries
// that point below the owned range land in the front pad of their diagonal.
__global__ void k_build_dia(const int* __restrict__ indptr, const int* __restrict__ indices,
                            const double* __restrict__ data, int64_t n, int64_t rb,
                            const double* __restrict__ s_owned, OffTable tab, double* vstore,
                            int64_t vstride, int64_t pad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const int64_t gi = r + rb;
    const double sr = s_owned[r];
    for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
      const int64_t j = indices[e];
      const int64_t o = j - gi;
      if (o == 0) continue;
      const int64_t lj = j - rb;
      if (o > 0) {
        const int m = off_lookup(tab, (int)o);
        vstore[(int64_t)m * vstride + pad + r] = data[e] * sr * s_owned[lj];
      } else if (lj < 0) {
        const int m = off_lookup(tab, (int)(-o));
        vstore[(int64_t)m * vstride + pad + lj] = data[e] * sr * s_owned[lj];
      }
    }
  }
}

// generic matrices only: every lower entry must mirror the stored upper one
__global__ void k_check_sym_dia(const int* __restrict__ indptr, const int* __restrict__ indices,
                                const double* __restrict__ data, int64_t n, int64_t rb,
                                const double* __restrict__ s_owned, OffTable tab,
                                const double* vstore, int64_t vstride, int64_t pad,
                                int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const int64_t gi = r + rb;
    const double sr = s_owned[r];
    for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
      const int64_t j = indices[e];
      const int64_t o = j - gi;
      const int64_t lj = j - rb;
      if (o >= 0 || lj < 0) continue;
      const int m = off_lookup(tab, (int)(-o));
      if (m < 0) { bad = true; continue; }
      const double a = data[e] * sr * s_owned[lj];
      const double u = vstore[(int64_t)m * vstride + pad + lj];
      if (fabs(a - u) > 1e-12 * fmax(fabs(a), fabs(u))) bad = true;
    }
  }
  if (bad) flags[FLAG_NONSYM] = 1;
}

// scaled off-diagonal CSR: counts, then values + local signed columns
__global__ void k_offdiag_count(const int* __restrict__ indptr, const int* __restrict__ diagpos,
                                int64_t n, int* __restrict__ cnt, int* __restrict__ maxrow) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int mx = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    int c = indptr[r + 1] - indptr[r] - (diagpos[r] >= 0 ? 1 : 0);
    cnt[r] = c;
    if (c > mx) mx = c;
  }
  atomicMax(maxrow, mx);
}

__global__ void k_offdiag_emit(const int* __restrict__ indptr, const int* __restrict__ indices,
                               const double* __restrict__ data, int64_t n, int64_t rb,
                               const double* __restrict__ s_owned, const int* __restrict__ optr,
                               int* __restrict__ ocols, double* __restrict__ ovals) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const int64_t gi = r + rb;
    const double sr = s_owned[r];
    int o = optr[r];
    for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
      const int64_t j = indices[e];
      if (j == gi) continue;
      const int64_t lj = j - rb;
      B200_ASSERT(o < optr[r + 1]);
      ocols[o] = (int)lj;
      ovals[o] = data[e] * sr * s_owned[lj];
      ++o;
    }
  }
}

// row blocks for the CSR-stream kernel: block q owns the rows whose work
// counter (entries before the row + row index) falls in [q*Q, (q+1)*Q)
__global__ void k_rowstart(const int* __restrict__ optr, int64_t n, int64_t Q, int64_t nblocks,
                           int* __restrict__ rowstart) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q > nblocks) return;
  if (q == nblocks) { rowstart[q] = (int)n; return; }
  const int64_t target = q * Q;
  int64_t lo = 0, hi = n;   // first r with optr[r] + r >= target
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if ((int64_t)optr[mid] + mid >= target) hi = mid; else lo = mid + 1;
  }
  rowstart[q] = (int)lo;
}

// ---- SELL-32-sigma build (irregular networks).  Inside every window of SELL_SIGMA
// consecutive rows the rows are sorted by descending length (bitonic sort of
// (length, index) keys in shared memory: deterministic), 32 sorted rows form a slice
// whose width is its first row's length; entries are stored column-major inside the
// slice, so a warp reads 32 consecutive values / columns per step and no reduction
// across lanes is needed.  Padding entries multiply 0.0 by the row's own p.
#define SELL_SIGMA 1024
#define SELL_C 32
__global__ void __launch_bounds__(SELL_SIGMA) k_sell_sort(const int* __restrict__ optr, int64_t n,
                                                          int* __restrict__ perm,
                                                          int* __restrict__ swidth) {
  __shared__ unsigned key[SELL_SIGMA];
  const int64_t base = (int64_t)blockIdx.x * SELL_SIGMA;
  const int t = threadIdx.x;
  const int64_t r = base + t;
  // ascending sort of (0x1fffff - len) << 10 | t  ==  descending length, ascending index
  unsigned k = 0xffffffffu;
  if (r < n) {
    const int len = optr[r + 1] - optr[r];
    k = ((unsigned)(0x1fffff - len) << 10) | (unsigned)t;
  }
  key[t] = k;
  __syncthreads();
  for (int size = 2; size <= SELL_SIGMA; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int partner = t ^ stride;
      if (partner > t) {
        const unsigned a = key[t], b = key[partner];
        const bool up = (t & size) == 0;
        if ((a > b) == up) { key[t] = b; key[partner] = a; }
      }
      __syncthreads();
    }
  const unsigned kk = key[t];
  const bool valid = kk != 0xffffffffu;
  perm[base + t] = valid ? (int)(base + (kk & 1023u)) : -1;
  if ((t & (SELL_C - 1)) == 0)
    swidth[(base + t) / SELL_C] = valid ? (int)(0x1fffff - (kk >> 10)) * SELL_C : 0;
}

__global__ void __launch_bounds__(256) k_sell_fill(const int* __restrict__ optr,
                                                   const int* __restrict__ ocols,
                                                   const double* __restrict__ ovals,
                                                   const int* __restrict__ perm,
                                                   const int* __restrict__ sptr, int64_t nslices,
                                                   int* __restrict__ scols,
                                                   double* __restrict__ svals) {
  const int lane = threadIdx.x & 31;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nslices; s += wstride) {
    const int b = sptr[s], width = (sptr[s + 1] - b) / SELL_C;
    const int r = perm[s * SELL_C + lane];
    const int e0 = r >= 0 ? optr[r] : 0, len = r >= 0 ? optr[r + 1] - e0 : 0;
    const int self = r >= 0 ? r : 0;
    for (int k = 0; k < width; ++k) {
      const bool in = k < len;
      B200_ASSERT(b + k * SELL_C + lane < sptr[nslices]);
      scols[b + k * SELL_C + lane] = in ? ocols[e0 + k] : self;
      svals[b + k * SELL_C + lane] = in ? ovals[e0 + k] : 0.0;
    }
  }
}

#ifndef CSR_Q
#define CSR_Q 2048          // work units (entries + rows) per CTA
#endif
#define CSR_MAXROW 4096     // longest row the streaming kernel accepts

static int prepare_operator(b200_ctx* ctx) {
  const int64_t n = ctx->n;
  RET(ensure_ws(ctx));
  if (!ctx->lin_valid) {
    if (ctx->srcs.n > 0)
      FAIL(B200_ERR_ARG, "sources are registered but not linearised: call b200_picard");
    RET(b200_linearize(ctx, nullptr));   // no sources: trivial
  }
  CK(cudaMemsetAsync(ctx->w_flags.p, 0, FLAG_COUNT * sizeof(int), ctx->stream));
  int* flags = ctx->w_flags.as<int>();
  const int* indptr = ctx->sys_indptr.as<int>();
  const int* indices = ctx->sys_indices.as<int>();
  const double* data = ctx->sys_data.as<double>();

  // ---- structure: offsets + bandwidth
  CK(ctx->w_tmp2.ensure(sizeof(OffTable)));
  OffTable* dtab = ctx->w_tmp2.as<OffTable>();
  CK(cudaMemsetAsync(dtab, 0, sizeof(OffTable), ctx->stream));
  LAUNCH(ctx, k_detect_offsets, grid_for(n, 128), 128, 0, indptr, indices, n, ctx->row_begin, dtab);
  OffTable tab;
  CK(cudaMemcpyAsync(&tab, dtab, sizeof tab, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // the slots were claimed in race order: sort them, so that the diagonals are
  // always summed in the same order and repeated solves agree bit for bit
  tab.n = 0;
  for (int k = 0; k < B200_MAX_DIAGS; ++k)
    if (tab.off[k] != 0) tab.off[tab.n++] = tab.off[k];
  std::sort(tab.off, tab.off + tab.n);
  for (int k = tab.n; k < B200_MAX_DIAGS; ++k) tab.off[k] = 0;
  const int64_t upper = (ctx->sys_nnz - n) / 2;   // off-diagonals are symmetric pairs
  bool want_dia = !tab.overflow && tab.n > 0 &&
                  (double)upper >= 0.25 * (double)tab.n * (double)n;
  if (ctx->force_fmt == B200_FMT_CSR || ctx->nonsym) want_dia = false;
  if (ctx->force_fmt == B200_FMT_DIA && (tab.overflow || tab.n == 0))
    FAIL(B200_ERR_ARG, "DIA format forced but the matrix has more than %d diagonals",
         B200_MAX_DIAGS);
  if (ctx->force_fmt == B200_FMT_DIA) want_dia = true;
  // slab runs: the format (and, below, the scaling) decides which communication path
  // the iteration takes, so the ranks must agree on it -- one rank on the NCCL path
  // while the others wait on their mailboxes would hang
  {
    int64_t no_dia = want_dia ? 0 : 1;
    RET(b200_allreduce_max_host(ctx, &no_dia));
    if (no_dia) {
      if (ctx->force_fmt == B200_FMT_DIA)
        FAIL(B200_ERR_ARG, "DIA format forced but a rank cannot use it");
      want_dia = false;
    }
  }
  int64_t halo = (want_dia || ctx->nranks > 1) ? tab.maxband : 0;
  RET(b200_allreduce_max_host(ctx, &halo));         // every rank uses the same halo width
  int64_t nmin = n;
  if (ctx->nranks > 1) {
    int64_t neg = -n;
    RET(b200_allreduce_max_host(ctx, &neg));
    nmin = -neg;
  }
  if (ctx->nranks > 1 && halo > nmin)
    FAIL(B200_ERR_ARG, "a slab of %lld rows is thinner than the matrix half-bandwidth %lld",
         (long long)nmin, (long long)halo);
  const int64_t pad = (halo + 31) / 32 * 32;        // allocation keeps 256-byte alignment
  ctx->halo = halo;
  ctx->pad = pad;
  const size_t vbytes = (size_t)(n + 2 * pad) * sizeof(double);

  // ---- scaling
  CK(ctx->sc.ensure(vbytes));
  CK(cudaMemsetAsync(ctx->sc.p, 0, vbytes, ctx->stream));
  double* s_owned = owned(ctx->sc, pad);
  LAUNCH(ctx, k_scale, grid_for(n, 256), 256, 0, cur_diag(ctx), n, 1, s_owned, flags);
  CK(cudaMemcpyAsync(ctx->flags_h, flags, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  {
    int64_t nonpos = ctx->flags_h[FLAG_NONPOS_DIAG] ? 1 : 0;   // agreed on by all ranks
    RET(b200_allreduce_max_host(ctx, &nonpos));
    ctx->unit_diag = !nonpos;
  }
  if (!ctx->unit_diag) {
    // indefinite / semi-definite diagonal: no scaling, explicit diagonal, CSR
    LAUNCH(ctx, k_scale, grid_for(n, 256), 256, 0, cur_diag(ctx), n, 0, s_owned, flags);
    CK(ctx->dexp.ensure((size_t)n * sizeof(double)));
    CK(cudaMemcpyAsync(ctx->dexp.p, cur_diag(ctx), (size_t)n * sizeof(double),
                       cudaMemcpyDeviceToDevice, ctx->stream));
    want_dia = false;
  }
  RET(b200_halo_exchange(ctx, ctx->sc.as<double>()));   // s of the halo columns

  for (DevBuf* v : {&ctx->vx, &ctx->vr, &ctx->vp, &ctx->vap, &ctx->vb}) {
    CK(v->ensure(vbytes));
    CK(cudaMemsetAsync(v->p, 0, vbytes, ctx->stream));
  }

  if (want_dia) {
    const int64_t vstride = n + pad;
    CK(ctx->dia_store.ensure((size_t)tab.n * vstride * sizeof(double)));
    CK(cudaMemsetAsync(ctx->dia_store.p, 0, (size_t)tab.n * vstride * sizeof(double), ctx->stream));
    LAUNCH(ctx, k_build_dia, grid_for(n, 128), 128, 0, indptr, indices, data, n, ctx->row_begin,
           s_owned, tab, ctx->dia_store.as<double>(), vstride, pad);
    if (ctx->sys_generic) {
      LAUNCH(ctx, k_check_sym_dia, grid_for(n, 128), 128, 0, indptr, indices, data, n,
             ctx->row_begin, s_owned, tab, ctx->dia_store.as<double>(), vstride, pad, flags);
      CK(cudaMemcpyAsync(ctx->flags_h, flags, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost,
                         ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      int64_t nonsym = ctx->flags_h[FLAG_NONSYM] ? 1 : 0;
      RET(b200_allreduce_max_host(ctx, &nonsym));
      if (nonsym) want_dia = false;
    }
    if (want_dia) {
      ctx->dia.nd = tab.n;
      for (int k = 0; k < tab.n; ++k) {
        ctx->dia.off[k] = tab.off[k];
        ctx->dia.v[k] = ctx->dia_store.as<double>() + (int64_t)k * vstride + pad;
      }
      ctx->fmt = B200_FMT_DIA;
    }
  }
  if (!want_dia) {
    CK(ctx->w_cnt.ensure((size_t)(n + 1) * sizeof(int)));
    CK(ctx->o_indptr.ensure((size_t)(n + 1) * sizeof(int)));
    int* mx = flags + FLAG_COUNT - 1;
    CK(cudaMemsetAsync(mx, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, k_offdiag_count, grid_for(n, 256), 256, 0, indptr, ctx->sys_diagpos.as<int>(), n,
           ctx->w_cnt.as<int>(), mx);
    int64_t onnz = 0;
    RET(b200_scan_exclusive_i32(ctx, ctx->w_cnt.as<int>(), ctx->o_indptr.as<int>(), n, &onnz));
    int maxrow = 0;
    CK(cudaMemcpy(&maxrow, mx, sizeof(int), cudaMemcpyDeviceToHost));
    CK(ctx->o_cols.ensure((size_t)(onnz + 2) * sizeof(int)));
    CK(ctx->o_vals.ensure((size_t)(onnz + 2) * sizeof(double)));
    LAUNCH(ctx, k_offdiag_emit, grid_for(n, 128), 128, 0, indptr, indices, data, n,
           ctx->row_begin, s_owned, ctx->o_indptr.as<int>(), ctx->o_cols.as<int>(),
           ctx->o_vals.as<double>());
    ctx->o_nnz = onnz;
    ctx->o_maxrow = maxrow;
    const int64_t work = onnz + n;
    ctx->o_nblocks = (work + CSR_Q - 1) / CSR_Q;
    CK(ctx->o_rowstart.ensure((size_t)(ctx->o_nblocks + 1) * sizeof(int)));
    LAUNCH(ctx, k_rowstart, (int)((ctx->o_nblocks + 1 + 255) / 256), 256, 0,
           ctx->o_indptr.as<int>(), n, (int64_t)CSR_Q, ctx->o_nblocks, ctx->o_rowstart.as<int>());
    ctx->fmt = B200_FMT_CSR;
    // sliced-ELL copy of the same operator (B200_CSR_KERNEL=stream keeps the
    // shared-memory CSR-stream kernel for A/B runs)
    const char* ek = getenv("B200_CSR_KERNEL");
    ctx->use_sell = !(ek && ek[0] == 's' && ek[1] == 't') && maxrow < 0x1fffff && n > 0;
    ctx->sell_staged = !(ek && strcmp(ek, "sell0") == 0);   // sell0: plain SELL, no staging
    if (ctx->use_sell) {
      const int64_t nwin = (n + SELL_SIGMA - 1) / SELL_SIGMA;
      const int64_t nsl = nwin * (SELL_SIGMA / SELL_C);
      CK(ctx->s_perm.ensure((size_t)(nwin * SELL_SIGMA) * sizeof(int)));
      CK(ctx->w_cnt.ensure((size_t)(nsl + 1) * sizeof(int)));
      CK(ctx->s_ptr.ensure((size_t)(nsl + 1) * sizeof(int)));
      LAUNCH(ctx, k_sell_sort, (int)nwin, SELL_SIGMA, 0, ctx->o_indptr.as<int>(), n,
             ctx->s_perm.as<int>(), ctx->w_cnt.as<int>());
      int64_t sent = 0;
      RET(b200_scan_exclusive_i32(ctx, ctx->w_cnt.as<int>(), ctx->s_ptr.as<int>(), nsl, &sent));
      CK(ctx->s_cols.ensure((size_t)(sent + 32) * sizeof(int)));
      CK(ctx->s_vals.ensure((size_t)(sent + 32) * sizeof(double)));
      LAUNCH(ctx, k_sell_fill, grid_for(nsl * 32, 256), 256, 0, ctx->o_indptr.as<int>(),
             ctx->o_cols.as<int>(), ctx->o_vals.as<double>(), ctx->s_perm.as<int>(),
             ctx->s_ptr.as<int>(), nsl, ctx->s_cols.as<int>(), ctx->s_vals.as<double>());
      ctx->s_nslices = nsl;
      ctx->s_entries = sent;
    }
  }
  CK(cudaGetLastError());
  ctx->op_epoch = ctx->sys_epoch;
  ctx->stats.fmt = ctx->fmt;
  ctx->stats.ndiags = ctx->fmt == B200_FMT_DIA ? ctx->dia.nd : 0;
  return B200_OK;
}

extern "C" int b200_pcg_prepare(b200_ctx* ctx, int force_format) {
  if (!ctx || !ctx->have_sys) {
    if (ctx) ctx->err = "pcg_prepare: no system matrix";
    return B200_ERR_ARG;
  }
  CK(cudaSetDevice(ctx->device));
  if (force_format != ctx->force_fmt) ctx->op_epoch = ~0ull;
  ctx->force_fmt = force_format;
  if (ctx->op_epoch == ctx->sys_epoch && ctx->lin_valid) return B200_OK;
  return prepare_operator(ctx);
}

extern "C" int b200_set_format(b200_ctx* ctx, int force_format) {
  if (!ctx || force_format < 0 || force_format > B200_FMT_DIA) return B200_ERR_ARG;
  if (force_format != ctx->force_fmt) ctx->op_epoch = ~0ull;
  ctx->force_fmt = force_format;
  return B200_OK;
}

extern "C" int b200_pcg_format(b200_ctx* ctx, int* fmt, int* ndiags) {
  if (!ctx) return B200_ERR_ARG;
  if (fmt) *fmt = ctx->fmt;
  if (ndiags) *ndiags = ctx->fmt == B200_FMT_DIA ? ctx->dia.nd : 0;
  return B200_OK;
}

// ======================================================= K4: the PCG kernels
// All kernels read their scalars (alpha/beta ingredients) from device memory;
// the host never sits between two iterations.

// ---- K_A (DIA): Ap = A~ p, partial p.Ap.  One row per thread, every load
// coalesced; the 2*ND neighbour reads of p and the second read of each v_m hit
// L1/L2 (re-use distance <= k_max rows).
template <int ND, bool UNIT>
__global__ void __launch_bounds__(PCG_T) k_dia_spmv_dot(DiaPack A, const double* __restrict__ p,
                                                        const double* __restrict__ rv,
                                                        const double* __restrict__ dexp,
                                                        double* __restrict__ ap, int64_t n,
                                                        int64_t lo0, int64_t hi0, int64_t lo1,
                                                        int accumulate, int chunk, double* part,
                                                        unsigned* ticket, double* out_pap) {
  // rows [lo0, hi0) and [lo1, n): the whole range (lo1 = n) or the two halo-
  // reading edges of a slab (hi0 = halo, lo1 = n - halo)
  // a CTA takes CHUNK consecutive 256-row tiles at a time, so that the +-k_y
  // neighbour reads of a tile fall into rows the same CTA (same L1) touches
  const int64_t len0 = hi0 - lo0, total = len0 + (n - lo1);
  const int64_t tile = (int64_t)blockDim.x * chunk;
  double dot = 0.0, dot2 = 0.0, dotr = 0.0;
  for (int64_t base = (int64_t)blockIdx.x * tile; base < total; base += (int64_t)gridDim.x * tile)
  for (int c = 0; c < chunk; ++c) {
    const int64_t q = base + (int64_t)c * blockDim.x + threadIdx.x;
    if (q >= total) break;
    const int64_t i = q < len0 ? lo0 + q : lo1 + (q - len0);
    const double pi = p[i];
    double acc = UNIT ? pi : dexp[i] * pi;
#pragma unroll
    for (int m = 0; m < ND; ++m) {
      const int64_t k = A.off[m];
      acc += A.v[m][i] * p[i + k];
      acc += A.v[m][i - k] * p[i - k];
    }
    ap[i] = acc;
    dot += pi * acc;
    dotr += rv[i] * acc;
    dot2 += acc * acc;
  }
  double v[3] = {dot, dotr, dot2};
  double* o[3] = {out_pap + G_PAP, out_pap + G_RAP, out_pap + G_APAP};
  grid_sum_finalize<3>(v, part, ticket, o, accumulate != 0);
}

// ---- fine-level multigrid smoothing on the DIA operator (unit diagonal): the
// product A~ p and the Chebyshev update / weighted residual that consumes it in
// one pass, written out of place (p is read with its stencil).  amg.cu, level 0.
//   mode 1  Chebyshev(2) from a zero guess on the right-hand side p:
//             x1 = c0 p, dd = c1 x1 + c2 (p - c0 y), d = dd, out = x1 + dd
//   mode 2  first step from x = p:   dd = c0 (b - y),          d = dd, out = p + dd
//   mode 3  next step from x = p:    dd = c1 d + c2 (b - y),   d = dd, out = p + dd,
//             and sum aux[i] out[i] (aux may be null)
//   mode 4  weighted residual:       out = aux[i] (b - y)
// (an fp32 copy of the diagonals for this kernel was tried: 301 instead of 278 ms per solve at
// 400^3 -- slower, the kernel is not limited by those bytes -- and dropped)
template <int ND>
__global__ void __launch_bounds__(PCG_T) k_dia_smooth(DiaPack A, int mode,
                                                      const double* __restrict__ p,
                                                      const double* __restrict__ b,
                                                      const double* __restrict__ aux,
                                                      double* __restrict__ d,
                                                      double* __restrict__ out, int64_t n,
                                                      double c0, double c1, double c2, int chunk,
                                                      double* part, unsigned* ticket,
                                                      double* dot_out) {
  const int64_t tile = (int64_t)blockDim.x * chunk;
  double dot = 0.0;
  for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile)
  for (int c = 0; c < chunk; ++c) {
    const int64_t i = base + (int64_t)c * blockDim.x + threadIdx.x;
    if (i >= n) break;
    const double pi = p[i];
    double y = pi;
#pragma unroll
    for (int m = 0; m < ND; ++m) {
      const int64_t k = A.off[m];
      y += A.v[m][i] * p[i + k];
      y += A.v[m][i - k] * p[i - k];
    }
    if (mode == 1) {
      const double x1 = c0 * pi;
      const double dd = c1 * x1 + c2 * (pi - c0 * y);
      d[i] = dd;
      out[i] = x1 + dd;
    } else if (mode == 2) {
      const double dd = c0 * (b[i] - y);
      d[i] = dd;
      out[i] = pi + dd;
    } else if (mode == 3) {
      const double dd = c1 * d[i] + c2 * (b[i] - y);
      const double o = pi + dd;
      d[i] = dd;
      out[i] = o;
      if (aux) dot += aux[i] * o;
    } else {
      out[i] = aux[i] * (b[i] - y);
    }
  }
  double v[1] = {dot};
  double* o[1] = {dot_out};
  grid_sum_finalize<1>(v, part, ticket, o);
}

bool b200_dia_smooth_available(b200_ctx* ctx) {
  return ctx->fmt == B200_FMT_DIA && ctx->unit_diag;
}

int b200_dia_smooth(b200_ctx* ctx, int mode, const double* p_owned, const double* b_owned,
                    const double* aux, double* d_owned, double* out_owned, double c0, double c1,
                    double c2, double* dot_out) {
  const int64_t n = ctx->n;
  double* part = ctx->w_part.as<double>();
  unsigned* ticket = ctx->w_ticket.as<unsigned>();
#define SM_CASE(ND)                                                                        \
  case ND: {                                                                               \
    static int occ = 0;                                                                    \
    if (!occ) {                                                                            \
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dia_smooth<ND>, PCG_T, 0)); \
      if (occ < 1) occ = 1;                                                                \
    }                                                                                      \
    int64_t g = (n + PCG_T - 1) / PCG_T;                                                   \
    if (g > (int64_t)occ * ctx->num_sms) g = (int64_t)occ * ctx->num_sms;                  \
    if (g < 1) g = 1;                                                                      \
    LAUNCH(ctx, (k_dia_smooth<ND>), (int)g, PCG_T, 0, ctx->dia, mode, p_owned, b_owned,    \
           aux, d_owned, out_owned, n, c0, c1, c2, ctx->ka_chunk, part, ticket, dot_out);  \
  } break;
  switch (ctx->dia.nd) {
    SM_CASE(1) SM_CASE(2) SM_CASE(3) SM_CASE(4) SM_CASE(5) SM_CASE(6) SM_CASE(7) SM_CASE(8)
    SM_CASE(9) SM_CASE(10) SM_CASE(11) SM_CASE(12) SM_CASE(13)
    default: FAIL(B200_ERR_ARG, "bad diagonal count");
  }
#undef SM_CASE
  CK(cudaGetLastError());
  return B200_OK;
}

// ---- K_A (SELL-32-sigma): one warp per slice, one row per lane; every step of the
// slice loop reads 32 consecutive values (256 B) and columns (128 B) and gathers p
// through L1/L2; four independent chains per lane keep the loads in flight.  The row's
// own p, r and Ap are scattered inside the sorting window (<= SELL_SIGMA rows: 8 KB).
template <bool UNIT>
__global__ void __launch_bounds__(PCG_T) k_sell_spmv_dot(
    const int* __restrict__ sptr, const int* __restrict__ scols, const double* __restrict__ svals,
    const int* __restrict__ perm, int64_t nslices, const double* __restrict__ p,
    const double* __restrict__ rv, const double* __restrict__ dexp, double* __restrict__ ap,
    double* part, unsigned* ticket, double* out_pap) {
  const int lane = threadIdx.x & 31;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double dot = 0.0, dot2 = 0.0, dotr = 0.0;
  for (int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nslices; s += wstride) {
    const int b = sptr[s], width = (sptr[s + 1] - b) >> 5;
    const int r = perm[s * SELL_C + lane];
    const int* c = scols + b + lane;
    const double* v = svals + b + lane;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int k = 0;
    for (; k + 4 <= width; k += 4) {
      const int c0 = c[(k + 0) * SELL_C], c1 = c[(k + 1) * SELL_C], c2 = c[(k + 2) * SELL_C],
                c3 = c[(k + 3) * SELL_C];
      const double v0 = v[(k + 0) * SELL_C], v1 = v[(k + 1) * SELL_C], v2 = v[(k + 2) * SELL_C],
                   v3 = v[(k + 3) * SELL_C];
      a0 += v0 * p[c0];
      a1 += v1 * p[c1];
      a2 += v2 * p[c2];
      a3 += v3 * p[c3];
    }
    for (; k < width; ++k) a0 += v[k * SELL_C] * p[c[k * SELL_C]];
    if (r >= 0) {
      const double pi = p[r];
      const double acc = (UNIT ? pi : dexp[r] * pi) + ((a0 + a1) + (a2 + a3));
      ap[r] = acc;
      dot += pi * acc;
      dotr += rv[r] * acc;
      dot2 += acc * acc;
    }
  }
  double vv[3] = {dot, dotr, dot2};
  double* o[3] = {out_pap + G_PAP, out_pap + G_RAP, out_pap + G_APAP};
  grid_sum_finalize<3>(vv, part, ticket, o);
}

// ---- K_A (SELL-32-sigma, staged operand): one CTA per sorting window of SELL_SIGMA rows.
// The slice of p the window's rows mostly read -- their own rows +- SELL_W -- is loaded
// coalesced into shared memory first, so a gather costs a shared-memory read instead of
// a 32-byte sector from L2 (columns outside the staged range fall back to global
// memory; with the pores numbered along a space-filling curve inside coordinate bins,
// dist.spatial_order, ~85 % of the columns are inside).  Row results go through
// shared memory too, so Ap is written and r is read coalesced for the three dot
// products.  North star: "shared-memory staging of x".
template <bool UNIT, int W, int UNR, int T>
__global__ void __launch_bounds__(T, (UNR >= 16 ? 2 : UNR >= 8 ? 4 : 5) * (256 / T)) k_sell_spmv_dot_w(
    const int* __restrict__ sptr, const int* __restrict__ scols, const double* __restrict__ svals,
    const int* __restrict__ perm, int64_t nwin, int64_t n, int64_t pad,
    const double* __restrict__ p, const double* __restrict__ rv, const double* __restrict__ dexp,
    double* __restrict__ ap, double* part, unsigned* ticket, double* out_pap) {
  extern __shared__ double sm[];
  double* pw = sm;                                  // p[lo .. lo + SIGMA + 2W)
  double* accw = sm + SELL_SIGMA + 2 * W;           // row sums of the window
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = T / 32, SPW = SELL_SIGMA / SELL_C;
  double dot = 0.0, dot2 = 0.0, dotr = 0.0;
  for (int64_t win = blockIdx.x; win < nwin; win += gridDim.x) {
    const int64_t r0 = win * SELL_SIGMA, lo = r0 - W;
    for (int i = threadIdx.x; i < SELL_SIGMA + 2 * W; i += T) {
      const int64_t idx = lo + i;
      pw[i] = (idx >= -pad && idx < n + pad) ? p[idx] : 0.0;
    }
    __syncthreads();
    const int ilo = (int)lo;
#define SELL_GET(cc) (((unsigned)((cc) - ilo) < (unsigned)(SELL_SIGMA + 2 * W)) ? pw[(cc) - ilo] : p[cc])
#define SELL_CHECK(cc) B200_ASSERT((cc) >= -pad && (cc) < n + pad)
    for (int sl = warp; sl < SPW; sl += NW) {
      const int64_t s = win * SPW + sl;
      const int b = sptr[s], width = (sptr[s + 1] - b) >> 5;
      const int r = perm[s * SELL_C + lane];
      const int* c = scols + b + lane;
      const double* v = svals + b + lane;
      double a[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) a[u] = 0.0;
      int k = 0;
      // tiers of UNR, then 4, then single steps: every tier issues all its column and
      // value loads before the first gather (memory-level parallelism is what this
      // kernel lives on: ncu shows it latency-bound, not bandwidth-bound)
#define SELL_TIER(T)                                                          \
      for (; k + (T) <= width; k += (T)) {                                    \
        int cc[T];                                                            \
        double vv[T];                                                         \
        _Pragma("unroll") for (int u = 0; u < (T); ++u) cc[u] = c[(k + u) * SELL_C];   \
        _Pragma("unroll") for (int u = 0; u < (T); ++u) vv[u] = v[(k + u) * SELL_C];   \
        _Pragma("unroll") for (int u = 0; u < (T); ++u) { SELL_CHECK(cc[u]); a[u] += vv[u] * SELL_GET(cc[u]); } \
      }
      SELL_TIER(UNR)
      if (UNR > 4) { SELL_TIER(4) }
      SELL_TIER(1)
#undef SELL_TIER
      double tot = 0.0;
#pragma unroll
      for (int u = UNR - 1; u >= 0; --u) tot += a[u];
      B200_ASSERT(r < 0 || (r >= r0 && r < r0 + SELL_SIGMA && r < n));
      if (r >= 0) accw[r - r0] = tot;
    }
#undef SELL_GET
#undef SELL_CHECK
    __syncthreads();
    for (int i = threadIdx.x; i < SELL_SIGMA; i += T) {
      const int64_t r = r0 + i;
      if (r < n) {
        const double pi = pw[i + W];
        const double acc = (UNIT ? pi : dexp[r] * pi) + accw[i];
        ap[r] = acc;
        dot += pi * acc;
        dotr += rv[r] * acc;
        dot2 += acc * acc;
      }
    }
    __syncthreads();
  }
  double vv3[3] = {dot, dotr, dot2};
  double* o[3] = {out_pap + G_PAP, out_pap + G_RAP, out_pap + G_APAP};
  grid_sum_finalize<3>(vv3, part, ticket, o);
}

// ---- K_A (CSR-stream): vals/cols streamed coalesced into shared memory as
// products, then one thread per row adds its segment.
template <bool UNIT>
__global__ void __launch_bounds__(PCG_T) k_csr_stream_spmv_dot(
    const int* __restrict__ optr, const int* __restrict__ ocols, const double* __restrict__ ovals,
    const int* __restrict__ rowstart, int nblocks, const double* __restrict__ p,
    const double* __restrict__ rv, const double* __restrict__ dexp, double* __restrict__ ap,
    double* part, unsigned* ticket, double* out_pap) {
  extern __shared__ double prod[];
  double dot = 0.0, dot2 = 0.0, dotr = 0.0;
  // persistent CTAs: row blocks are dealt round-robin
  for (int q = blockIdx.x; q < nblocks; q += gridDim.x) {
    const int r0 = rowstart[q], r1 = rowstart[q + 1];
    const int e0 = optr[r0], e1 = optr[r1];
    int e = e0 + threadIdx.x;
    // four independent (val, col, gather) chains in flight per thread
    for (; e + 3 * PCG_T < e1; e += 4 * PCG_T) {
      const int c0 = ocols[e], c1 = ocols[e + PCG_T], c2 = ocols[e + 2 * PCG_T],
                c3 = ocols[e + 3 * PCG_T];
      const double v0 = ovals[e], v1 = ovals[e + PCG_T], v2 = ovals[e + 2 * PCG_T],
                   v3 = ovals[e + 3 * PCG_T];
      const double x0 = p[c0], x1 = p[c1], x2 = p[c2], x3 = p[c3];
      prod[e - e0] = v0 * x0;
      prod[e + PCG_T - e0] = v1 * x1;
      prod[e + 2 * PCG_T - e0] = v2 * x2;
      prod[e + 3 * PCG_T - e0] = v3 * x3;
    }
    for (; e < e1; e += PCG_T) prod[e - e0] = ovals[e] * p[ocols[e]];
    __syncthreads();
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
      const double pi = p[r];
      double acc = UNIT ? pi : dexp[r] * pi;
      const int b = optr[r] - e0, e = optr[r + 1] - e0;
      for (int k = b; k < e; ++k) acc += prod[k];
      ap[r] = acc;
      dot += pi * acc;
      dotr += rv[r] * acc;
      dot2 += acc * acc;
    }
    __syncthreads();
  }
  double v[3] = {dot, dotr, dot2};
  double* o[3] = {out_pap + G_PAP, out_pap + G_RAP, out_pap + G_APAP};
  grid_sum_finalize<3>(v, part, ticket, o);
}

// fallback for matrices with very long rows: one warp per row
template <bool UNIT>
__global__ void __launch_bounds__(PCG_T) k_csr_warp_spmv_dot(
    const int* __restrict__ optr, const int* __restrict__ ocols, const double* __restrict__ ovals,
    const double* __restrict__ p, const double* __restrict__ rv, const double* __restrict__ dexp,
    double* __restrict__ ap, int64_t n, double* part, unsigned* ticket, double* out_pap) {
  const int lane = threadIdx.x & 31;
  const int64_t wstride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double dot = 0.0, dot2 = 0.0, dotr = 0.0;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += wstride) {
    double s = 0.0;
    for (int e = optr[r] + lane; e < optr[r + 1]; e += 32) s += ovals[e] * p[ocols[e]];
    s = warp_sum(s);
    if (lane == 0) {
      const double pi = p[r];
      s += UNIT ? pi : dexp[r] * pi;
      ap[r] = s;
      dot += pi * s;
      dotr += rv[r] * s;
      dot2 += s * s;
    }
  }
  double v[3] = {dot, dotr, dot2};
  double* o[3] = {out_pap + G_PAP, out_pap + G_RAP, out_pap + G_APAP};
  grid_sum_finalize<3>(v, part, ticket, o);
}

// ---- K_BC: x += alpha p ; r -= alpha Ap ; p = r + beta p in one pass (56 B/row).
// One all-reduce per iteration: K_A delivers p.Ap, r.Ap and Ap.Ap together, so
//   alpha = rr / p.Ap        rr' = |r - alpha Ap|^2 = rr - 2 alpha r.Ap + alpha^2 Ap.Ap
// gives beta = rr'/rr before r is touched.  rr' is an algebraic identity on the
// actual vectors (no conjugacy assumption), its cancellation error is
// eps*rr in absolute terms, i.e. eps in beta.  It is used for beta only: the
// kernel also sums the new r.r explicitly and that exact value (all-reduced
// with the next iteration's K_A sums) is the rr of the next step, so nothing
// accumulates.  TRUE_NORM adds sum d_i r_i^2 for the host's convergence test.
template <bool TRUE_NORM>
__global__ void __launch_bounds__(PCG_T) k_update_xrp(double* __restrict__ x,
                                                      double* __restrict__ r,
                                                      double* __restrict__ p,
                                                      const double* __restrict__ ap,
                                                      const double* __restrict__ diag, int64_t n,
                                                      const double* __restrict__ gin,
                                                      double* __restrict__ gout, double* part,
                                                      unsigned* ticket) {
  const double rr = gin[G_RR], pap = gin[G_PAP], rap = gin[G_RAP], apap = gin[G_APAP];
  const double alpha = (pap > 0.0 || pap < 0.0) ? rr / pap : 0.0;
  double rr_new = rr - 2.0 * alpha * rap + alpha * alpha * apap;
  if (!(rr_new > 0.0)) rr_new = 0.0;
  const double beta = (rr > 0.0) ? rr_new / rr : 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0, t = 0.0;
  const int64_t n2 = n >> 1;
  double2* x2 = reinterpret_cast<double2*>(x);
  double2* r2 = reinterpret_cast<double2*>(r);
  double2* p2 = reinterpret_cast<double2*>(p);
  const double2* a2 = reinterpret_cast<const double2*>(ap);
  const double2* d2 = reinterpret_cast<const double2*>(diag);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 xv = x2[i], rv = r2[i], pv = p2[i];
    const double2 av = a2[i];
    xv.x += alpha * pv.x; xv.y += alpha * pv.y;
    rv.x -= alpha * av.x; rv.y -= alpha * av.y;
    pv.x = rv.x + beta * pv.x; pv.y = rv.y + beta * pv.y;
    x2[i] = xv; r2[i] = rv; p2[i] = pv;
    s += rv.x * rv.x + rv.y * rv.y;
    if (TRUE_NORM) { const double2 dv = d2[i]; t += dv.x * rv.x * rv.x + dv.y * rv.y * rv.y; }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t i = n - 1;
    const double po = p[i];
    const double rn = r[i] - alpha * ap[i];
    x[i] += alpha * po;
    r[i] = rn;
    p[i] = rn + beta * po;
    s += rn * rn;
    if (TRUE_NORM) t += diag[i] * rn * rn;
  }
  if (TRUE_NORM) {
    double v[2] = {s, t};
    double* o[2] = {gout + G_RR, gout + G_TRUE};
    grid_sum_finalize<2>(v, part, ticket, o);
  } else {
    double v[1] = {s};
    double* o[1] = {gout + G_RR};
    grid_sum_finalize<1>(v, part, ticket, o);
  }
}

// ---- setup / teardown kernels
// xt = sqrt(d) x0 (or the exact value on singleton rows), bt = s b, |b|^2
__global__ void __launch_bounds__(PCG_T) k_pcg_init(const double* __restrict__ x0,
                                                    const double* __restrict__ b,
                                                    const double* __restrict__ s,
                                                    const int* __restrict__ sys_indptr,
                                                    const int* __restrict__ diagpos, int unit,
                                                    const double* __restrict__ dexp,
                                                    double* __restrict__ xt, double* __restrict__ bt,
                                                    int64_t n, double* part, unsigned* ticket,
                                                    double* out_b2, int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double b2 = 0.0;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double bi = b[i], si = s[i];
    const double bti = si * bi;
    bt[i] = bti;
    b2 += bi * bi;
    const bool singleton = (sys_indptr[i + 1] - sys_indptr[i] == 1) && diagpos[i] >= 0;
    double x = x0 ? x0[i] / si : 0.0;
    if (!isfinite(x) || !isfinite(bi)) bad = true;
    if (singleton) x = unit ? bti : ((dexp[i] != 0.0) ? bti / dexp[i] : x);
    xt[i] = x;
  }
  if (bad) flags[FLAG_NONFINITE] = 1;
  double v[1] = {b2};
  double* o[1] = {out_b2};
  grid_sum_finalize<1>(v, part, ticket, o);
}

// r = bt - ap ; p = r ; rr = r.r ; true = sum d r^2
__global__ void __launch_bounds__(PCG_T) k_pcg_resid(const double* __restrict__ bt,
                                                     const double* __restrict__ ap,
                                                     const double* __restrict__ diag,
                                                     double* __restrict__ r, double* __restrict__ p,
                                                     int64_t n, double* part, unsigned* ticket,
                                                     double* out_rr, double* out_true) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0, t = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double ri = bt[i] - ap[i];
    r[i] = ri;
    p[i] = ri;
    s += ri * ri;
    t += diag[i] * ri * ri;
  }
  double v[2] = {s, t};
  double* o[2] = {out_rr, out_true};
  grid_sum_finalize<2>(v, part, ticket, o);
}

__global__ void __launch_bounds__(PCG_T) k_pcg_final(const double* __restrict__ xt,
                                                     const double* __restrict__ s,
                                                     double* __restrict__ x, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] = s[i] * xt[i];
}

__global__ void k_fill_const(double* a, int64_t n, double v) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = v;
}

// launch helpers shared with bicgstab.cu
int b200_launch_init(b200_ctx* ctx, const double* x0, double* xt, double* bt, double* out_b2) {
  const int g1 = grid_for(ctx->n, PCG_T);
  LAUNCH(ctx, k_pcg_init, g1, PCG_T, 0, x0, cur_b(ctx), owned(ctx->sc, ctx->pad),
         ctx->sys_indptr.as<int>(), ctx->sys_diagpos.as<int>(), ctx->unit_diag ? 1 : 0,
         ctx->dexp.as<double>(), xt, bt, ctx->n, ctx->w_part.as<double>(),
         ctx->w_ticket.as<unsigned>(), out_b2, ctx->w_flags.as<int>());
  return B200_OK;
}
int b200_launch_resid(b200_ctx* ctx, const double* bt, const double* ap, const double* d,
                      double* r, double* p, double* out_rr, double* out_true) {
  const int g1 = grid_for(ctx->n, PCG_T);
  LAUNCH(ctx, k_pcg_resid, g1, PCG_T, 0, bt, ap, d, r, p, ctx->n, ctx->w_part.as<double>(),
         ctx->w_ticket.as<unsigned>(), out_rr, out_true);
  return B200_OK;
}
int b200_launch_final(b200_ctx* ctx, const double* xt, double* x) {
  LAUNCH(ctx, k_pcg_final, grid_for(ctx->n, PCG_T), PCG_T, 0, xt, owned(ctx->sc, ctx->pad), x,
         ctx->n);
  return B200_OK;
}
const double* b200_cur_diag(b200_ctx* ctx) { return cur_diag(ctx); }

// --------------------------------------------------------------- launchers
static int pcg_grid(b200_ctx* ctx, int per_thread) {
  // persistent-style grid: a multiple of the SM count, grid-stride inside
  int64_t want = (ctx->n + (int64_t)PCG_T * per_thread - 1) / ((int64_t)PCG_T * per_thread);
  const int64_t cap = 148 * 8;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

// part = 0: all rows; 1: interior rows only (no halo read); 2: the two edges,
// accumulated onto the sums part 1 left (DIA only; CSR always runs whole)
static int launch_spmv_dot(b200_ctx* ctx, const double* p_owned, const double* r_owned,
                           double* ap_owned, double* out, int part_sel = 0) {
  double* part = ctx->w_part.as<double>();
  unsigned* ticket = ctx->w_ticket.as<unsigned>();
  const double* dexp = ctx->dexp.as<double>();
  const int64_t n = ctx->n;
  if (ctx->fmt == B200_FMT_DIA) {
    int64_t lo0 = 0, hi0 = n, lo1 = n;
    if (part_sel == 1) { lo0 = ctx->halo; hi0 = n - ctx->halo; }
    if (part_sel == 2) { lo0 = 0; hi0 = ctx->halo; lo1 = n - ctx->halo; }
    const int64_t rows = (hi0 - lo0) + (n - lo1);
    const int acc = part_sel == 2 ? 1 : 0;
    // one full wave: resident CTAs per SM (from the occupancy calculator, the
    // register count differs per ND) x SM count, grid-stride inside
#define DIA_CASE(ND)                                                                       \
  case ND: {                                                                               \
    static int occ = 0;                                                                    \
    if (!occ) {                                                                            \
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dia_spmv_dot<ND, true>,     \
                                                       PCG_T, 0));                         \
      if (occ < 1) occ = 1;                                                                \
    }                                                                                      \
    int64_t g = (rows + PCG_T - 1) / PCG_T;                                                \
    if (g > (int64_t)occ * ctx->num_sms) g = (int64_t)occ * ctx->num_sms;                  \
    if (g < 1) g = 1;                                                                      \
    LAUNCH(ctx, (k_dia_spmv_dot<ND, true>), (int)g, PCG_T, 0, ctx->dia, p_owned, r_owned,  \
           dexp, ap_owned, n, lo0, hi0, lo1, acc, ctx->ka_chunk, part, ticket, out);                      \
  } break;
    switch (ctx->dia.nd) {
      DIA_CASE(1) DIA_CASE(2) DIA_CASE(3) DIA_CASE(4) DIA_CASE(5) DIA_CASE(6) DIA_CASE(7)
      DIA_CASE(8) DIA_CASE(9) DIA_CASE(10) DIA_CASE(11) DIA_CASE(12) DIA_CASE(13)
      default: FAIL(B200_ERR_ARG, "bad diagonal count");
    }
#undef DIA_CASE
  } else if (ctx->use_sell && ctx->sell_staged && n + 2 * ctx->pad < (1ll << 31) - 8192) {
    const int64_t nwin = ctx->s_nslices / (SELL_SIGMA / SELL_C);
    int64_t g = nwin < 148 * 16 ? nwin : 148 * 16;
    if (g < 1) g = 1;
    static int variant = -1;      // tuning knob: B200_SELL_VARIANT = 0..5 (default 4: staged window +-256,
                                  // unroll 8, 128-thread CTAs: best of the measured set, profiles/README.md)
    if (variant < 0) {
      const char* e = getenv("B200_SELL_VARIANT");
      variant = (e && e[0] >= '0' && e[0] <= '5') ? e[0] - '0' : 4;
    }
#define SELL_LAUNCH(W, UNR, T)                                                                \
  {                                                                                           \
    const size_t smem = (size_t)(2 * SELL_SIGMA + 2 * (W)) * sizeof(double);                   \
    static bool attr = false;                                                                 \
    if (!attr) {                                                                              \
      CK(cudaFuncSetAttribute(k_sell_spmv_dot_w<true, W, UNR, T>,                             \
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
      CK(cudaFuncSetAttribute(k_sell_spmv_dot_w<false, W, UNR, T>,                            \
                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
      attr = true;                                                                            \
    }                                                                                         \
    if (ctx->unit_diag)                                                                       \
      LAUNCH(ctx, (k_sell_spmv_dot_w<true, W, UNR, T>), (int)g, T, smem, ctx->s_ptr.as<int>(), \
             ctx->s_cols.as<int>(), ctx->s_vals.as<double>(), ctx->s_perm.as<int>(), nwin, n, \
             ctx->pad, p_owned, r_owned, dexp, ap_owned, part, ticket, out);                  \
    else                                                                                      \
      LAUNCH(ctx, (k_sell_spmv_dot_w<false, W, UNR, T>), (int)g, T, smem, ctx->s_ptr.as<int>(), \
             ctx->s_cols.as<int>(), ctx->s_vals.as<double>(), ctx->s_perm.as<int>(), nwin, n, \
             ctx->pad, p_owned, r_owned, dexp, ap_owned, part, ticket, out);                  \
  }
    if (variant == 1) SELL_LAUNCH(1024, 8, 256)
    else if (variant == 2) SELL_LAUNCH(512, 8, 128)
    else if (variant == 4) SELL_LAUNCH(256, 8, 128)
    else if (variant == 5) SELL_LAUNCH(512, 8, 64)
    else if (variant == 0) SELL_LAUNCH(1024, 4, 256)
    else SELL_LAUNCH(512, 8, 256)
#undef SELL_LAUNCH
  } else if (ctx->use_sell) {
    static int occ = 0;
    if (!occ) {
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sell_spmv_dot<true>, PCG_T, 0));
      if (occ < 1) occ = 1;
    }
    int64_t g = (ctx->s_nslices * 32 + PCG_T - 1) / PCG_T;
    if (g > (int64_t)occ * ctx->num_sms) g = (int64_t)occ * ctx->num_sms;
    if (g < 1) g = 1;
    if (ctx->unit_diag)
      LAUNCH(ctx, k_sell_spmv_dot<true>, (int)g, PCG_T, 0, ctx->s_ptr.as<int>(),
             ctx->s_cols.as<int>(), ctx->s_vals.as<double>(), ctx->s_perm.as<int>(),
             ctx->s_nslices, p_owned, r_owned, dexp, ap_owned, part, ticket, out);
    else
      LAUNCH(ctx, k_sell_spmv_dot<false>, (int)g, PCG_T, 0, ctx->s_ptr.as<int>(),
             ctx->s_cols.as<int>(), ctx->s_vals.as<double>(), ctx->s_perm.as<int>(),
             ctx->s_nslices, p_owned, r_owned, dexp, ap_owned, part, ticket, out);
  } else if (ctx->o_maxrow <= CSR_MAXROW) {
    const size_t smem = (size_t)(CSR_Q + ctx->o_maxrow + 1) * sizeof(double);
    const int nb = (int)ctx->o_nblocks;
    const int g = nb < 148 * 8 ? nb : 148 * 8;
    if (ctx->unit_diag)
      LAUNCH(ctx, k_csr_stream_spmv_dot<true>, g, PCG_T, smem, ctx->o_indptr.as<int>(),
             ctx->o_cols.as<int>(), ctx->o_vals.as<double>(), ctx->o_rowstart.as<int>(), nb,
             p_owned, r_owned, dexp, ap_owned, part, ticket, out);
    else
      LAUNCH(ctx, k_csr_stream_spmv_dot<false>, g, PCG_T, smem, ctx->o_indptr.as<int>(),
             ctx->o_cols.as<int>(), ctx->o_vals.as<double>(), ctx->o_rowstart.as<int>(), nb,
             p_owned, r_owned, dexp, ap_owned, part, ticket, out);
  } else {
    const int g = grid_for(n * 32, PCG_T);
    if (ctx->unit_diag)
      LAUNCH(ctx, k_csr_warp_spmv_dot<true>, g, PCG_T, 0, ctx->o_indptr.as<int>(),
             ctx->o_cols.as<int>(), ctx->o_vals.as<double>(), p_owned, r_owned, dexp, ap_owned, n,
             part, ticket, out);
    else
      LAUNCH(ctx, k_csr_warp_spmv_dot<false>, g, PCG_T, 0, ctx->o_indptr.as<int>(),
             ctx->o_cols.as<int>(), ctx->o_vals.as<double>(), p_owned, r_owned, dexp, ap_owned, n,
             part, ticket, out);
  }
  return B200_OK;
}

// out = A~ w on the prepared operator (w, out: owned pointers of padded vectors);
// the dot products K_A makes on the side land in a scratch group
// K_A on arbitrary operands: out = A~ w and the three sums w.out, rv.out, out.out
// into group[G_PAP], group[G_RAP], group[G_APAP]
int b200_spmv_dots(b200_ctx* ctx, const double* w_owned, const double* rv_owned,
                   double* out_owned, double* group) {
  return launch_spmv_dot(ctx, w_owned, rv_owned, out_owned, group);
}

// DIA operator: rows [hi0, lo1) only (the rows between the two halo-reading edges of a
// slab); sums overwrite group[G_PAP..G_APAP].  p2p.cu adds the edges.
int b200_spmv_dots_rows(b200_ctx* ctx, const double* p_owned, const double* r_owned,
                        double* ap_owned, double* group, int64_t hi0, int64_t lo1) {
  if (ctx->fmt != B200_FMT_DIA) FAIL(B200_ERR_ARG, "spmv_dots_rows: DIA operator only");
  const int64_t n = ctx->n;
  const int64_t rows = lo1 - hi0;
  double* part = ctx->w_part.as<double>();
  unsigned* ticket = ctx->w_ticket.as<unsigned>();
  const double* dexp = ctx->dexp.as<double>();
#define ROWS_CASE(ND)                                                                      \
  case ND: {                                                                               \
    static int occ = 0;                                                                    \
    if (!occ) {                                                                            \
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dia_spmv_dot<ND, true>,     \
                                                       PCG_T, 0));                         \
      if (occ < 1) occ = 1;                                                                \
    }                                                                                      \
    int64_t g = (rows + PCG_T - 1) / PCG_T;                                                \
    if (g > (int64_t)occ * ctx->num_sms) g = (int64_t)occ * ctx->num_sms;                  \
    if (g < 1) g = 1;                                                                      \
    LAUNCH(ctx, (k_dia_spmv_dot<ND, true>), (int)g, PCG_T, 0, ctx->dia, p_owned, r_owned,  \
           dexp, ap_owned, n, hi0, lo1, n, 0, ctx->ka_chunk, part, ticket, group);         \
  } break;
  switch (ctx->dia.nd) {
    ROWS_CASE(1) ROWS_CASE(2) ROWS_CASE(3) ROWS_CASE(4) ROWS_CASE(5) ROWS_CASE(6) ROWS_CASE(7)
    ROWS_CASE(8) ROWS_CASE(9) ROWS_CASE(10) ROWS_CASE(11) ROWS_CASE(12) ROWS_CASE(13)
    default: FAIL(B200_ERR_ARG, "bad diagonal count");
  }
#undef ROWS_CASE
  return B200_OK;
}

int b200_apply_operator(b200_ctx* ctx, const double* w_owned, double* out_owned) {
  return launch_spmv_dot(ctx, w_owned, ctx->vr.as<double>() + ctx->pad, out_owned,
                         ctx->scal.as<double>() + SC_AUX0);
}

extern "C" int b200_pcg(b200_ctx* ctx, const double* x0, double rtol, double atol,
                        int64_t maxiter, double* x, int64_t* iters, double* relres, int* info) {
  if (!ctx || !ctx->have_sys || !x) {
    if (ctx) ctx->err = "pcg: no system matrix / null output";
    return B200_ERR_ARG;
  }
  CK(cudaSetDevice(ctx->device));
  if (ctx->nonsym) return b200_bicgstab(ctx, x0, rtol, atol, maxiter, x, iters, relres, info);
  if (b200_amg_applicable(ctx)) {
    int used = 0;
    int amg_info = 0;
    // the multigrid-preconditioned solve gets a bounded number of iterations: it
    // needs ~30 when it works; a matrix it cannot handle (not an M-matrix, say)
    // is handed to Jacobi-PCG below, starting from the last iterate
    int64_t amg_cap = 300;
    if (const char* e = getenv("B200_AMG_MAXITER")) {      // test / tuning knob
      const long v = strtol(e, nullptr, 10);
      if (v > 0) amg_cap = v;
    }
    if (maxiter > 0 && maxiter < amg_cap) amg_cap = maxiter;
    RET(b200_amg_pcg(ctx, x0, rtol, atol, amg_cap, x, iters, relres, &amg_info, &used));
    if (used && amg_info == 0) {
      if (info) *info = 0;
      return B200_OK;
    }
    if (used) {
      if (ctx->debug)
        fprintf(stderr, "[b200 amg] not converged (info %d): continuing with Jacobi-PCG\n", amg_info);
      if (amg_info > 0) x0 = x;    // finite but slow: keep what was gained; breakdown: start over
    }
  }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  RET(b200_pcg_prepare(ctx, ctx->force_fmt));
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  const int64_t n = ctx->n, pad = ctx->pad;
  if (maxiter <= 0) maxiter = 10 * ctx->n_global;
  double* sc = ctx->scal.as<double>();
  double* part = ctx->w_part.as<double>();
  unsigned* ticket = ctx->w_ticket.as<unsigned>();
  int* flags = ctx->w_flags.as<int>();
  double* xt = owned(ctx->vx, pad);
  double* r = owned(ctx->vr, pad);
  double* p = owned(ctx->vp, pad);
  double* ap = owned(ctx->vap, pad);
  double* bt = owned(ctx->vb, pad);
  const double* s = owned(ctx->sc, pad);
  const double* d = cur_diag(ctx);
  const int g2 = pcg_grid(ctx, 2);
  const int g1 = pcg_grid(ctx, 1);
  const bool multi = ctx->nranks > 1;

  CK(cudaMemsetAsync(flags, 0, FLAG_COUNT * sizeof(int), ctx->stream));
  CK(cudaMemsetAsync(sc, 0, SC_COUNT * sizeof(double), ctx->stream));
  LAUNCH(ctx, k_pcg_init, g1, PCG_T, 0, x0, cur_b(ctx), s, ctx->sys_indptr.as<int>(),
         ctx->sys_diagpos.as<int>(), ctx->unit_diag ? 1 : 0, ctx->dexp.as<double>(), xt, bt, n,
         part, ticket, sc + SC_BNORM, flags);
  if (!ctx->unit_diag) {
    // no scaling was applied: the recurrence residual already is the true one
    CK(ctx->w_tmp1.ensure((size_t)n * sizeof(double)));
    LAUNCH(ctx, k_fill_const, g1, PCG_T, 0, ctx->w_tmp1.as<double>(), n, 1.0);
    d = ctx->w_tmp1.as<double>();
  }
  if (multi) {
    RET(b200_allreduce_sum(ctx, sc + SC_BNORM, 1));
    RET(b200_halo_exchange(ctx, ctx->vx.as<double>()));
  }
  // scratch for the sums of SpMVs whose dot products nobody reads: r is passed
  // as bt (any valid vector); group 0 receives rr0 / true0 of the start residual
  RET(launch_spmv_dot(ctx, xt, bt, ap, sc + SC_AUX0));
  LAUNCH(ctx, k_pcg_resid, g1, PCG_T, 0, bt, ap, d, r, p, n, part, ticket, sc + G_RR,
         sc + G_TRUE);
  if (multi) RET(b200_allreduce_sum(ctx, sc + G_RR, 2));
  CK(cudaMemcpyAsync(ctx->scal_h, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaMemcpyAsync(ctx->flags_h, flags, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // every rank must take the same decision: in slab mode judge by the all-reduced
  // norms (an inf/nan anywhere makes them non-finite everywhere)
  if ((!multi && ctx->flags_h[FLAG_NONFINITE]) || !isfinite(ctx->scal_h[SC_BNORM]) ||
      !isfinite(ctx->scal_h[G_TRUE]))
    FAIL(B200_ERR_NONFINITE, "b or x0 contains inf/nan values");

  const double bnorm2 = ctx->scal_h[SC_BNORM];
  double tol2 = rtol * rtol * bnorm2;
  if (atol * atol > tol2) tol2 = atol * atol;
  double true2 = ctx->scal_h[G_TRUE];
  int64_t it = 0;
  int inf = 0;
  int restarts = 0;
  if (bnorm2 == 0.0) {
    // scipy.sparse.linalg.cg returns b (zeros) at once for a zero rhs
    CK(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.cg_iters = 0;
    ctx->stats.relres = 0.0;
    if (iters) *iters = 0;
    if (relres) *relres = 0.0;
    if (info) *info = 0;
    return B200_OK;
  }

  int batch = 8;
  double prev_true2 = true2;
  bool done = !(true2 > tol2);
  // rr / true2 of the group about to be consumed are already global (they were
  // all-reduced on their own for the host); otherwise they ride along with the
  // K_A sums in the single per-iteration all-reduce
  bool rr_is_global = true;
  const bool p2p = multi && b200_p2p_usable(ctx);
  bool p_halo_stale = true;      // p = r from k_pcg_resid: its halo was not pushed by K_BC
  while (!done) {
    if (it >= maxiter) break;
    int64_t todo = batch;
    if (it + todo > maxiter) todo = maxiter - it;
    for (int64_t k = 0; k < todo; ++k, ++it) {
      double* gin = sc + SC_GROUP * (int)(it & 1);
      double* gout = sc + SC_GROUP * (int)((it + 1) & 1);
      if (p2p) {
        // fused communication over NVLink peer memory (p2p.cu): the only NCCL
        // call left is the halo of a p that did not come out of K_BC
        if (p_halo_stale) RET(b200_halo_exchange(ctx, ctx->vp.as<double>()));
        RET(b200_p2p_iteration(ctx, gin, gout, k == todo - 1, !p_halo_stale, rr_is_global, d));
        p_halo_stale = false;
        rr_is_global = false;
        continue;
      }
      if (multi && ctx->fmt == B200_FMT_DIA && n > 4 * ctx->halo) {
        // interior rows while the halo windows of p travel; then the two edges
        RET(b200_halo_exchange_begin(ctx, ctx->vp.as<double>()));
        RET(launch_spmv_dot(ctx, p, r, ap, gin, 1));
        RET(b200_halo_exchange_end(ctx));
        RET(launch_spmv_dot(ctx, p, r, ap, gin, 2));
      } else {
        if (multi) RET(b200_halo_exchange(ctx, ctx->vp.as<double>()));
        RET(launch_spmv_dot(ctx, p, r, ap, gin));
      }
      if (multi) RET(b200_allreduce_sum(ctx, gin, rr_is_global ? 3 : SC_GROUP));
      rr_is_global = false;
      if (k == todo - 1)
        LAUNCH(ctx, k_update_xrp<true>, g2, PCG_T, 0, xt, r, p, ap, d, n, gin, gout, part, ticket);
      else
        LAUNCH(ctx, k_update_xrp<false>, g2, PCG_T, 0, xt, r, p, ap, d, n, gin, gout, part, ticket);
    }
    double* gcur = sc + SC_GROUP * (int)(it & 1);   // rr / true2 after the last update
    const int cur = SC_GROUP * (int)(it & 1);
    const int prv = SC_GROUP * (int)((it + 1) & 1);
    if (multi) RET(b200_allreduce_sum(ctx, gcur + G_RR, 2));
    rr_is_global = true;
    CK(cudaMemcpyAsync(ctx->scal_h, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (p2p) RET(b200_p2p_check(ctx));
    true2 = ctx->scal_h[cur + G_TRUE];
    const double pap = ctx->scal_h[prv + G_PAP];
    const double rrn = ctx->scal_h[cur + G_RR];
    if (ctx->debug)
      fprintf(stderr, "[b200 pcg] it=%lld batch=%d true2=%.3e tol2=%.3e rr=%.3e pAp=%.3e\n",
              (long long)it, batch, true2, tol2, rrn, pap);
    if (!isfinite(true2) || !isfinite(pap) || !isfinite(rrn)) { inf = -1; break; }
    if (rrn > 0.0 && !(pap > 0.0)) { inf = -1; break; }   // not SPD: breakdown
    if (true2 <= tol2) {
      // confirm with the explicitly recomputed residual (guards against drift
      // of the recurrence); r and p restart from it if it disagrees
      if (multi) RET(b200_halo_exchange(ctx, ctx->vx.as<double>()));
      RET(launch_spmv_dot(ctx, xt, bt, ap, sc + SC_AUX0));
      LAUNCH(ctx, k_pcg_resid, g1, PCG_T, 0, bt, ap, d, r, p, n, part, ticket, gcur + G_RR,
             gcur + G_TRUE);
      if (multi) RET(b200_allreduce_sum(ctx, gcur + G_RR, 2));
      CK(cudaMemcpyAsync(ctx->scal_h, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost,
                         ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      true2 = ctx->scal_h[cur + G_TRUE];
      if (ctx->debug)
        fprintf(stderr, "[b200 pcg] verify it=%lld true2=%.3e tol2=%.3e restarts=%d\n",
                (long long)it, true2, tol2, restarts);
      if (true2 <= tol2 * (1.0 + 1e-6) || restarts >= 4) { done = true; break; }
      ++restarts;
      prev_true2 = true2;
      batch = 8;
      p_halo_stale = true;
      continue;
    }
    // choose the next batch from the observed convergence rate
    if (true2 < prev_true2 && true2 > 0.0) {
      const double rate = log(prev_true2 / true2) / (double)todo;   // per iteration
      const double need = log(true2 / tol2) / rate;
      int nb = (int)(need * 0.7);
      if (nb < 4) nb = 4;
      if (nb > 64) nb = 64;
      batch = nb;
    } else {
      batch = batch < 64 ? batch * 2 : 64;
    }
    prev_true2 = true2;
  }
  LAUNCH(ctx, k_pcg_final, g1, PCG_T, 0, xt, s, x, n);
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

// bytes one K_A launch has to move in the storage format in use (what the roofline
// fraction "on format bytes" divides by): operator + p, r read + Ap written
extern "C" int b200_operator_bytes(b200_ctx* ctx, int64_t* bytes, int64_t* entries) {
  if (!ctx || !bytes) return B200_ERR_ARG;
  if (ctx->fmt == B200_FMT_NONE) FAIL(B200_ERR_ARG, "operator_bytes: run b200_pcg_prepare first");
  const int64_t n = ctx->n;
  int64_t b = 24 * n, e = 0;
  if (ctx->fmt == B200_FMT_DIA) {
    b += (int64_t)ctx->dia.nd * 8 * n;
    e = (int64_t)ctx->dia.nd * n;
  } else if (ctx->use_sell) {
    b += 12 * ctx->s_entries + 4 * ctx->s_nslices * SELL_C + 4 * ctx->s_nslices;
    e = ctx->s_entries;
  } else {
    b += 12 * ctx->o_nnz + 4 * n;
    e = ctx->o_nnz;
  }
  *bytes = b;
  if (entries) *entries = e;
  return B200_OK;
}

extern "C" int b200_time_kernel(b200_ctx* ctx, int which, int reps, double* ms_avg) {
  if (!ctx || !ctx->have_sys || reps < 1 || !ms_avg) return B200_ERR_ARG;
  if (ctx->fmt == B200_FMT_NONE) FAIL(B200_ERR_ARG, "time_kernel: run a solve first");
  CK(cudaSetDevice(ctx->device));
  const int64_t n = ctx->n, pad = ctx->pad;
  double* sc = ctx->scal.as<double>();
  double* part = ctx->w_part.as<double>();
  unsigned* ticket = ctx->w_ticket.as<unsigned>();
  double* xt = owned(ctx->vx, pad);
  double* r = owned(ctx->vr, pad);
  double* p = owned(ctx->vp, pad);
  double* ap = owned(ctx->vap, pad);
  const double* d = cur_diag(ctx);
  const int g2 = pcg_grid(ctx, 2);
  // warm-up launch, then the timed train.  Scalars are parked in a spare group
  // (alpha = 0 there, so the vectors stay finite).
  double* gin = sc + SC_AUX0;
  double* gout = sc + SC_AUX0;
  for (int rep = -1; rep < reps; ++rep) {
    if (rep == 0) CK(cudaEventRecord(ctx->ev0, ctx->stream));
    switch (which) {
      case 0: RET(launch_spmv_dot(ctx, p, r, ap, gin)); break;
      case 1:
        LAUNCH(ctx, k_update_xrp<false>, g2, PCG_T, 0, xt, r, p, ap, d, n, gin, gout, part, ticket);
        break;
      case 2:
        LAUNCH(ctx, k_update_xrp<true>, g2, PCG_T, 0, xt, r, p, ap, d, n, gin, gout, part, ticket);
        break;
      case 3:
        RET(b200_spmv_csr_raw(ctx, ctx->sys_indptr.as<int>(), ctx->sys_indices.as<int>(),
                              ctx->sys_data.as<double>(), n, ctx->row_begin, p, ap));
        break;
      default: FAIL(B200_ERR_ARG, "time_kernel: which must be 0..3");
    }
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaEventSynchronize(ctx->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ms_avg = (double)ms / reps;
  return B200_OK;
}
