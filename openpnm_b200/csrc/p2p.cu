// p2p.cu -- PCG iteration with the collectives fused into the kernels over
// NVLink peer memory (CUDA IPC): no NCCL call, no extra launch per iteration.
//
//   K_A  (k_dia_spmv_dot_p2p):  waits for the neighbours' halo flags, Ap = A~ p,
//        and its LAST block stores {p.Ap, r.Ap, Ap.Ap, r.r, sum d r^2} of this
//        rank into every peer's mailbox (remote st + st.release.sys of the
//        sequence number).
//   K_BC (k_update_xrp_p2p):    every block reads the N mailboxes (ld.acquire.sys
//        spin on the sequence number), adds them in rank order -- bit-identical
//        on all ranks, no floating-point atomics -- updates x, r, p and stores
//        the edge entries of the new p directly into the neighbours' halo
//        windows; its last block publishes the halo flags.
//
// This is the slab algorithm of csrc/dist.cu (reference: the row-block CG of
// openpnm/solvers/_petsc.py:117-128,160) with the all-reduce and the halo
// exchange expressed as peer stores.  Kernels of different ranks run on
// different GPUs and only ever wait for data a *previous* kernel of the peer
// produced, so there is no circular wait; every spin is bounded and raises an
// error flag instead of hanging.
#include <stdlib.h>

#include "common.cuh"

#define P2P_T 256
#define P2P_MAXR 8
#define P2P_SPIN_LIMIT (1ll << 24)

struct P2PArgs {
  int nranks, rank;
  unsigned long long seq;        // sequence number of this iteration (>= 1)
  unsigned long long wait_halo;  // K_A: neighbours' halo flags must reach this (0: no wait)
  int rr_global;                 // rr / true2 of the group are already global sums
  double* mbox;                  // local mailbox [2][8][8]
  unsigned long long* flags;     // local flags
  double* peer_mbox[P2P_MAXR];
  unsigned long long* peer_flags_lo;   // flags of rank-1 / rank+1 (null at the ends)
  unsigned long long* peer_flags_hi;
  double* peer_vp_lo;            // base of the neighbours' padded p vectors
  double* peer_vp_hi;
  long long n_lo;                // rows owned by rank-1
  long long pad, halo;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// bounded spin; returns false (and raises the error flag) on timeout
__device__ __forceinline__ bool spin_until_ge(const unsigned long long* p, unsigned long long want,
                                              unsigned long long* flags) {
  for (long long t = 0; t < P2P_SPIN_LIMIT; ++t) {
    if (ld_acquire_sys(p) >= want) return true;
    __nanosleep(64);
  }
  atomicExch(&flags[2], 1ull);
  return false;
}

// grid-wide deterministic sums; returns true in the last block, whose shared
// `tot` then holds the totals (also written to out[])
template <int NV>
__device__ __forceinline__ bool grid_sum_last(const double (&v)[NV], double* part,
                                              unsigned* ticket, double* const (&out)[NV],
                                              double* tot) {
  __shared__ bool is_last;
  double bs[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) bs[k] = block_sum(v[k]);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) part[(size_t)k * gridDim.x + blockIdx.x] = bs[k];
    __threadfence();
    unsigned t = atomicInc(ticket, gridDim.x - 1);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double s = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x)
        s += __ldcg(&part[(size_t)k * gridDim.x + i]);
      s = block_sum(s);
      if (threadIdx.x == 0) { *out[k] = s; tot[k] = s; }
    }
    __syncthreads();
  }
  return is_last;
}

// K_A of the fused iteration is two launches: the rows that read no halo entry go
// through the ordinary K_A (solver.cu: k_dia_spmv_dot, same code and speed as on one
// GPU) without waiting for anybody; this kernel then waits for the neighbours' halo
// flags -- usually long since set, the interior is ~98 % of the rows -- does the
// 2 * halo edge rows, adds its sums to the interior's and its last block publishes the
// rank's totals to every peer's mailbox.  The halo windows are written by the peers'
// K_BC while this rank may already be computing: the edge rows read p with
// ld.global.cg (from L2, where peer stores land) after the acquire on the flag, never
// through the read-only path.  rows: [0, hi0) and [lo1, n); empty ranges = publish only.
template <int ND>
__global__ void __launch_bounds__(P2P_T) k_dia_edges_p2p(DiaPack A, const double* p,
                                                         const double* __restrict__ rv,
                                                         double* __restrict__ ap, int64_t n,
                                                         int64_t hi0, int64_t lo1, double* part,
                                                         unsigned* ticket, double* gin, P2PArgs a) {
  __shared__ double tot[3];
  if (a.wait_halo) {
    if (threadIdx.x == 0) {
      if (a.rank > 0) spin_until_ge(&a.flags[0], a.wait_halo, a.flags);
      if (a.rank < a.nranks - 1) spin_until_ge(&a.flags[1], a.wait_halo, a.flags);
    }
    __syncthreads();
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t cnt = hi0 + (n - lo1);
  double dot = 0.0, dot2 = 0.0, dotr = 0.0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += stride) {
    const int64_t i = q < hi0 ? q : lo1 + (q - hi0);
    const double pi = __ldcg(p + i);
    double acc = pi;
#pragma unroll
    for (int m = 0; m < ND; ++m) {
      const int64_t k = A.off[m];
      acc += A.v[m][i] * __ldcg(p + i + k);
      acc += A.v[m][i - k] * __ldcg(p + i - k);
    }
    ap[i] = acc;
    dot += pi * acc;
    dotr += rv[i] * acc;
    dot2 += acc * acc;
  }
  // add the interior's sums (written by the previous launch) in a fixed order
  double v[3] = {dot, dotr, dot2};
  double part_int[3] = {gin[G_PAP], gin[G_RAP], gin[G_APAP]};
  double* o[3] = {gin + G_PAP, gin + G_RAP, gin + G_APAP};
  if (grid_sum_last<3>(v, part, ticket, o, tot)) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { tot[k] += part_int[k]; *o[k] = tot[k]; }
    }
    __syncthreads();
    if (threadIdx.x < a.nranks) {
      double rrc = gin[G_RR], trc = gin[G_TRUE];
      if (a.rr_global && a.rank != 0) { rrc = 0.0; trc = 0.0; }
      double* base = (threadIdx.x == a.rank) ? a.mbox : a.peer_mbox[threadIdx.x];
      double* dst = base + (((a.seq & 1ull) * P2P_MAXR) + a.rank) * 8;
      dst[0] = tot[0]; dst[1] = tot[1]; dst[2] = tot[2]; dst[3] = rrc; dst[4] = trc;
      st_release_sys(reinterpret_cast<unsigned long long*>(dst + 7), a.seq);
    }
  }
}

template <bool TRUE_NORM>
__global__ void __launch_bounds__(P2P_T) k_update_xrp_p2p(double* __restrict__ x,
                                                          double* __restrict__ r,
                                                          double* __restrict__ p,
                                                          const double* __restrict__ ap,
                                                          const double* __restrict__ diag,
                                                          int64_t n, double* gin, double* gout,
                                                          double* part, unsigned* ticket,
                                                          P2PArgs a) {
  __shared__ double pay[P2P_MAXR][5];
  __shared__ double g[5];
  __shared__ double tot[2];
  if (threadIdx.x < a.nranks) {
    const double* src = a.mbox + (((a.seq & 1ull) * P2P_MAXR) + threadIdx.x) * 8;
    spin_until_ge(reinterpret_cast<const unsigned long long*>(src + 7), a.seq, a.flags);
#pragma unroll
    for (int k = 0; k < 5; ++k) pay[threadIdx.x][k] = __ldcg(src + k);
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0.0;
    for (int j = 0; j < a.nranks; ++j) s += pay[j][threadIdx.x];   // fixed rank order
    g[threadIdx.x] = s;
    if (blockIdx.x == 0) gin[threadIdx.x] = s;                     // global sums, for the host
  }
  __syncthreads();
  const double rr = g[G_RR], pap = g[G_PAP], rap = g[G_RAP], apap = g[G_APAP];
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
  const int64_t hi_begin = n - a.halo;
  bool remote = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    double2 xv = x2[i], rv = r2[i], pv = p2[i];
    const double2 av = a2[i];
    xv.x += alpha * pv.x; xv.y += alpha * pv.y;
    rv.x -= alpha * av.x; rv.y -= alpha * av.y;
    pv.x = rv.x + beta * pv.x; pv.y = rv.y + beta * pv.y;
    x2[i] = xv; r2[i] = rv; p2[i] = pv;
    s += rv.x * rv.x + rv.y * rv.y;
    if (TRUE_NORM) { const double2 dv = d2[i]; t += dv.x * rv.x * rv.x + dv.y * rv.y * rv.y; }
    // edge entries of the new p go straight into the neighbours' halo windows
    const int64_t row = 2 * i;
    if (a.peer_vp_lo && row < a.halo) {
      double* dst = a.peer_vp_lo + a.pad + a.n_lo;
      dst[row] = pv.x;
      if (row + 1 < a.halo) dst[row + 1] = pv.y;
      remote = true;
    }
    if (a.peer_vp_hi && row + 1 >= hi_begin) {
      double* dst = a.peer_vp_hi + a.pad - a.halo;
      if (row >= hi_begin) dst[row - hi_begin] = pv.x;
      dst[row + 1 - hi_begin] = pv.y;
      remote = true;
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t i = n - 1;
    const double po = p[i];
    const double rn = r[i] - alpha * ap[i];
    const double pn = rn + beta * po;
    x[i] += alpha * po;
    r[i] = rn;
    p[i] = pn;
    s += rn * rn;
    if (TRUE_NORM) t += diag[i] * rn * rn;
    if (a.peer_vp_lo && i < a.halo) { a.peer_vp_lo[a.pad + a.n_lo + i] = pn; remote = true; }
    if (a.peer_vp_hi && i >= hi_begin) { a.peer_vp_hi[a.pad - a.halo + (i - hi_begin)] = pn; remote = true; }
  }
  if (remote) __threadfence_system();    // peer stores visible before this block's ticket
  bool last;
  if (TRUE_NORM) {
    double v[2] = {s, t};
    double* o[2] = {gout + G_RR, gout + G_TRUE};
    last = grid_sum_last<2>(v, part, ticket, o, tot);
  } else {
    double v[1] = {s};
    double* o[1] = {gout + G_RR};
    last = grid_sum_last<1>(v, part, ticket, o, tot);
  }
  if (last && threadIdx.x == 0) {
    __threadfence_system();
    if (a.peer_flags_lo) st_release_sys(a.peer_flags_lo + 1, a.seq);   // I am their rank+1
    if (a.peer_flags_hi) st_release_sys(a.peer_flags_hi + 0, a.seq);   // I am their rank-1
  }
}

// ------------------------------------------------------------------ host side
static const size_t kP2PBytes = 4096;

extern "C" int b200_p2p_export(b200_ctx* ctx, void* handles128_h) {
  if (!ctx || !handles128_h) return B200_ERR_ARG;
  if (ctx->nranks < 2 || ctx->nranks > P2P_MAXR)
    FAIL(B200_ERR_ARG, "p2p needs 2..%d ranks", P2P_MAXR);
  if (!ctx->vp.p || ctx->fmt != B200_FMT_DIA)
    FAIL(B200_ERR_ARG, "p2p: call b200_pcg_prepare first (DIA operator only)");
  CK(cudaSetDevice(ctx->device));
  P2PState& s = ctx->p2p;
  if (!s.mem) {
    CK(cudaMalloc(&s.mem, kP2PBytes));
    CK(cudaMemset(s.mem, 0, kP2PBytes));
    s.mbox = reinterpret_cast<double*>(s.mem);
    s.flags = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(s.mem) + 2048);
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  cudaIpcMemHandle_t h[2];
  CK(cudaIpcGetMemHandle(&h[0], s.mem));
  CK(cudaIpcGetMemHandle(&h[1], ctx->vp.p));
  memcpy(handles128_h, h, sizeof h);
  return B200_OK;
}

void b200_p2p_release(b200_ctx* ctx) {
  P2PState& s = ctx->p2p;
  for (int i = 0; i < s.n_opened; ++i)
    if (s.opened[i]) cudaIpcCloseMemHandle(s.opened[i]);
  s.n_opened = 0;
  if (s.mem) cudaFree(s.mem);
  s = P2PState();
}

extern "C" int b200_p2p_disable(b200_ctx* ctx) {
  if (!ctx) return B200_ERR_ARG;
  ctx->p2p.ready = false;
  return B200_OK;
}

extern "C" int b200_p2p_attach(b200_ctx* ctx, const void* all_handles_h, const int64_t* all_rows_h) {
  if (!ctx || !all_handles_h || !all_rows_h) return B200_ERR_ARG;
  P2PState& s = ctx->p2p;
  if (!s.mem) FAIL(B200_ERR_ARG, "p2p_attach: export first");
  CK(cudaSetDevice(ctx->device));
  for (int i = 0; i < s.n_opened; ++i)
    if (s.opened[i]) cudaIpcCloseMemHandle(s.opened[i]);
  s.n_opened = 0;
  const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles_h);
  for (int r = 0; r < ctx->nranks; ++r) {
    s.n_of[r] = all_rows_h[r];
    s.peer_mbox[r] = nullptr;
    s.peer_flags[r] = nullptr;
    s.peer_vp[r] = nullptr;
    if (r == ctx->rank) continue;
    void* m = nullptr;
    CK(cudaIpcOpenMemHandle(&m, h[2 * r], cudaIpcMemLazyEnablePeerAccess));
    s.opened[s.n_opened++] = m;
    s.peer_mbox[r] = reinterpret_cast<double*>(m);
    s.peer_flags[r] = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(m) + 2048);
    if (r == ctx->rank - 1 || r == ctx->rank + 1) {
      void* v = nullptr;
      CK(cudaIpcOpenMemHandle(&v, h[2 * r + 1], cudaIpcMemLazyEnablePeerAccess));
      s.opened[s.n_opened++] = v;
      s.peer_vp[r] = reinterpret_cast<double*>(v);
    }
  }
  s.vp_at_attach = ctx->vp.p;
  s.pad_at_attach = ctx->pad;
  s.halo_at_attach = ctx->halo;
  s.seq = 0;
  CK(cudaMemset(s.mem, 0, kP2PBytes));
  s.ready = true;
  return B200_OK;
}

bool b200_p2p_usable(b200_ctx* ctx) {
  const P2PState& s = ctx->p2p;
  return s.ready && ctx->nranks > 1 && ctx->fmt == B200_FMT_DIA && ctx->unit_diag &&
         s.vp_at_attach == ctx->vp.p && s.pad_at_attach == ctx->pad &&
         s.halo_at_attach == ctx->halo;
}

int b200_p2p_check(b200_ctx* ctx) {
  unsigned long long e = 0;
  CK(cudaMemcpyAsync(&e, ctx->p2p.flags + 2, sizeof e, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (e) {
    // a timed-out exchange leaves mailboxes and sequence numbers in an unknown state:
    // clear the flag and stop using the fused path on this context (the NCCL back end
    // takes over at the next solve) instead of failing every later solve as well
    CK(cudaMemsetAsync(ctx->p2p.flags + 2, 0, sizeof(unsigned long long), ctx->stream));
    ctx->p2p.ready = false;
    FAIL(B200_ERR_NCCL, "p2p: a peer did not answer in time (spin limit reached); "
                        "fused communication disabled for this context");
  }
  return B200_OK;
}

int b200_p2p_iteration(b200_ctx* ctx, double* gin, double* gout, bool true_norm, bool wait_halo,
                       bool rr_is_global, const double* d) {
  P2PState& s = ctx->p2p;
  const int64_t n = ctx->n, pad = ctx->pad;
  P2PArgs a;
  a.nranks = ctx->nranks;
  a.rank = ctx->rank;
  a.seq = ++s.seq;
  a.wait_halo = wait_halo ? a.seq - 1 : 0;
  a.rr_global = rr_is_global ? 1 : 0;
  a.mbox = s.mbox;
  a.flags = s.flags;
  for (int r = 0; r < P2P_MAXR; ++r) a.peer_mbox[r] = r < ctx->nranks ? s.peer_mbox[r] : nullptr;
  a.peer_flags_lo = ctx->rank > 0 ? s.peer_flags[ctx->rank - 1] : nullptr;
  a.peer_flags_hi = ctx->rank < ctx->nranks - 1 ? s.peer_flags[ctx->rank + 1] : nullptr;
  a.peer_vp_lo = ctx->rank > 0 ? s.peer_vp[ctx->rank - 1] : nullptr;
  a.peer_vp_hi = ctx->rank < ctx->nranks - 1 ? s.peer_vp[ctx->rank + 1] : nullptr;
  a.n_lo = ctx->rank > 0 ? s.n_of[ctx->rank - 1] : 0;
  a.pad = pad;
  a.halo = ctx->halo;
  double* part = ctx->w_part.as<double>();
  unsigned* ticket = ctx->w_ticket.as<unsigned>();
  double* xt = ctx->vx.as<double>() + pad;
  double* r = ctx->vr.as<double>() + pad;
  double* p = ctx->vp.as<double>() + pad;
  double* ap = ctx->vap.as<double>() + pad;
  // interior rows: the ordinary K_A, nobody waits; with no halo to wait for (first
  // iteration after an NCCL exchange) it takes all rows and the edge kernel only publishes
  int64_t hi0 = 0, lo1 = n;
  if (a.wait_halo && n > 2 * ctx->halo) {
    hi0 = ctx->rank > 0 ? ctx->halo : 0;
    lo1 = ctx->rank < ctx->nranks - 1 ? n - ctx->halo : n;
    RET(b200_spmv_dots_rows(ctx, p, r, ap, gin, hi0, lo1));
  } else if (!a.wait_halo) {
    RET(b200_spmv_dots_rows(ctx, p, r, ap, gin, 0, n));
  } else {
    hi0 = n;                      // slab thinner than two halos: every row waits
    CK(cudaMemsetAsync(gin, 0, 3 * sizeof(double), ctx->stream));
  }
#define P2P_CASE(ND)                                                                        \
  case ND: {                                                                                \
    int64_t g = (hi0 + (n - lo1) + P2P_T - 1) / P2P_T;                                      \
    if (g > (int64_t)8 * ctx->num_sms) g = (int64_t)8 * ctx->num_sms;                       \
    if (g < 1) g = 1;                                                                       \
    LAUNCH(ctx, k_dia_edges_p2p<ND>, (int)g, P2P_T, 0, ctx->dia, p, r, ap, n, hi0, lo1,     \
           part, ticket, gin, a);                                                           \
  } break;
  switch (ctx->dia.nd) {
    P2P_CASE(1) P2P_CASE(2) P2P_CASE(3) P2P_CASE(4) P2P_CASE(5) P2P_CASE(6) P2P_CASE(7)
    P2P_CASE(8) P2P_CASE(9) P2P_CASE(10) P2P_CASE(11) P2P_CASE(12) P2P_CASE(13)
    default: FAIL(B200_ERR_ARG, "bad diagonal count");
  }
#undef P2P_CASE
  int64_t g2 = (n + (int64_t)P2P_T * 2 - 1) / ((int64_t)P2P_T * 2);
  if (g2 > (int64_t)ctx->num_sms * 8) g2 = (int64_t)ctx->num_sms * 8;
  if (g2 < 1) g2 = 1;
  if (true_norm)
    LAUNCH(ctx, k_update_xrp_p2p<true>, (int)g2, P2P_T, 0, xt, r, p, ap, d, n, gin, gout, part,
           ticket, a);
  else
    LAUNCH(ctx, k_update_xrp_p2p<false>, (int)g2, P2P_T, 0, xt, r, p, ap, d, n, gin, gout, part,
           ticket, a);
  return B200_OK;
}
