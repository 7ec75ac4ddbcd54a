// common.cuh -- context, buffers, scan and reduction primitives shared by the
// kernels of libb200pnm.so (sm_100a).  No reference counterpart: the reference
// (pure Python) delegates all of this to numpy/scipy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "b200pnm.h"

// Debug build (python -m openpnm_b200.build --bounds -> libb200pnm_dbg.so, selected with
// B200_LIB=dbg): in-kernel index checks on the gathers and scatters of the hot kernels.
// compute-sanitizer is not available on the GPU pool, so the small golden tests are run
// under this build instead (profiles/README.md).
#ifdef B200_BOUNDS
#include <assert.h>
#define B200_ASSERT(c) assert(c)
#else
#define B200_ASSERT(c) ((void)0)
#endif

#define B200_MAX_SRC 8
#define B200_MAX_DIAGS 13     // Cubic(connectivity=26) has 13 super-diagonals

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct SrcTerm {
  int kind;
  const uint8_t* mask;
  const double *p1, *p2, *p3;
};
struct SrcPack {
  int n;
  SrcTerm s[B200_MAX_SRC];
};

struct DiaPack {
  int nd;
  int off[B200_MAX_DIAGS];
  const double* v[B200_MAX_DIAGS];  // each points at row 0 of the owned range
};

// device-side scalar slots of the PCG (all doubles, in ctx->scal).
// Two ping-pong groups of five, indexed by iteration parity; a group is the
// unit of the single all-reduce an iteration makes:
//   [0] p.Ap  [1] r.Ap  [2] Ap.Ap      written by K_A of iteration k
//   [3] r.r   [4] sum d_i r_i^2        written by K_BC of iteration k-1
enum {
  SC_GROUP = 5,
  G_PAP = 0, G_RAP = 1, G_APAP = 2, G_RR = 3, G_TRUE = 4,
  SC_BNORM = 10, // ||b||^2
  SC_AUX0 = 11,
  SC_AUX1 = 12,
  SC_AUX2 = 13,
  SC_AUX3 = 14,
  SC_AUX4 = 15,
  SC_COUNT = 16
};

// Peer-to-peer state of the fused-communication PCG (p2p.cu): mailboxes and
// halo windows of the neighbour ranks mapped through CUDA IPC over NVLink.
struct P2PState {
  bool ready = false;
  void* mem = nullptr;                 // local: mailboxes + flags (one cudaMalloc, IPC-exported)
  double* mbox = nullptr;              // [2 parities][8 ranks][8 doubles]
  unsigned long long* flags = nullptr; // [0] halo from rank-1, [1] halo from rank+1, [2] error
  double* peer_mbox[8] = {};
  unsigned long long* peer_flags[8] = {};
  double* peer_vp[8] = {};
  void* opened[16] = {};
  int n_opened = 0;
  void* vp_at_attach = nullptr;
  int64_t pad_at_attach = 0, halo_at_attach = 0;
  int64_t n_of[8] = {};
  unsigned long long seq = 0;
};

// Conductance that follows the quantity, evaluated on the device inside the Picard loop
// (assembly.cu: b200_set_conductance_model).  All pointers are caller-owned device arrays.
struct GModel {
  bool active = false;
  int kind = 0, interp = 0, npoly = 0, sf_cols = 1, keep_zero_diag = 0;
  double poly[8] = {};
  const int64_t* conns = nullptr;
  int64_t Nt = 0, Np = 0;
  const double* throat_prop = nullptr;
  const double* sf = nullptr;
  const double* bcv = nullptr;
  const double* bcr = nullptr;
};

struct b200_ctx {
  int device = 0;
  int num_sms = 148;
  int ka_chunk = 2;           // 256-row tiles a K_A CTA takes at a time (B200_KA_CHUNK)
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaStream_t comm_stream = nullptr;   // halo exchange overlapping the interior SpMV
  cudaEvent_t ev_vec_ready = nullptr, ev_halo_done = nullptr;
  std::string err;
  bool debug = false;         // B200_DEBUG=1: trace the host loops on stderr
  int64_t launches = 0;

  // ownership (slab); single GPU: row_begin = 0, row_end = n_global
  int64_t n_global = 0, row_begin = 0, row_end = 0, n = 0;
  int nranks = 1, rank = 0;
  void* nccl = nullptr;  // ncclComm_t
  P2PState p2p;

  // pure Laplacian, canonical CSR of the owned rows (global column ids)
  DevBuf pure_indptr, pure_indices, pure_data, pure_diag;
  int64_t pure_nnz = 0;
  bool have_pure = false;
  bool nonsym = false;        // two-way (Nt,2) weights: A is not symmetric -> BiCGStab, no DIA-sym

  // system matrix (after BCs, or loaded), b, diagonal bookkeeping
  DevBuf sys_indptr, sys_indices, sys_data, sys_diag, sys_diagpos, b;
  int64_t sys_nnz = 0;
  bool have_sys = false;
  bool sys_generic = false;   // came from load_coo/load_csr: symmetry unknown
  double f = 0.0;
  bool f_override = false;
  uint64_t sys_epoch = 0;     // bumped whenever structure/values change

  // sources
  SrcPack srcs{};
  DevBuf diag_eff, b_eff;     // linearised diagonal and rhs
  bool lin_valid = false;
  uint64_t lin_epoch = 0;     // bumped by every re-linearisation of the sources

  // PCG operator (Jacobi-scaled, unit diagonal)
  uint64_t op_epoch = ~0ull;
  int fmt = B200_FMT_NONE;
  int force_fmt = 0;
  bool unit_diag = true;
  int64_t pad = 0;            // padding entries on each side of vectors (>= halo)
  int64_t halo = 0;           // entries actually exchanged with each neighbour rank
  DevBuf sc;                  // s_i = d_i^-1/2, padded layout
  DevBuf dexp;                // explicit diagonal when !unit_diag
  DiaPack dia{};
  DevBuf dia_store;
  DevBuf o_indptr, o_cols, o_vals, o_rowstart;  // scaled off-diagonal CSR
  int64_t o_nnz = 0, o_nblocks = 0, o_maxrow = 0;
  // the same operator as sliced ELLPACK (SELL-32-sigma): rows sorted by length inside
  // windows of SELL_SIGMA rows, 32 rows per slice stored column-major (solver.cu)
  DevBuf s_perm, s_ptr, s_cols, s_vals;
  int64_t s_nslices = 0, s_entries = 0;
  bool use_sell = false, sell_staged = true;
  DevBuf vx, vr, vp, vap, vb;  // PCG vectors (padded layout)
  DevBuf vbi;                  // BiCGStab extras: rhat, s (padded), t
  DevBuf scal;                 // SC_COUNT doubles
  double* scal_h = nullptr;    // pinned mirror

  // workspaces
  DevBuf w_cnt, w_keys, w_scan, w_part, w_ticket, w_flags, w_tmp0, w_tmp1, w_tmp2;
  int* flags_h = nullptr;      // pinned
  DevBuf h_row, h_col, h_dat, h_vec;   // device staging of b200_solve_coo_h
  DevBuf w_coll;                       // staging of the small host-visible collectives (dist.cu)

  GModel gm;
  DevBuf gm_g;                 // conductances of the last evaluation of the model

  // preconditioner: B200_PRECOND_JACOBI or B200_PRECOND_AMG (amg.cu owns `amg`)
  int precond = 0;
  void* amg = nullptr;

  // statistics
  b200_stats stats{};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
};

// ---------------------------------------------------------------- error glue
#define CK(call)                                                              \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) {                                                  \
      char _buf[512];                                                         \
      snprintf(_buf, sizeof _buf, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,\
               cudaGetErrorString(_e));                                       \
      ctx->err = _buf;                                                        \
      return B200_ERR_CUDA;                                                   \
    }                                                                         \
  } while (0)

#define FAIL(code, ...)                                                       \
  do {                                                                        \
    char _buf[512];                                                           \
    snprintf(_buf, sizeof _buf, __VA_ARGS__);                                 \
    ctx->err = _buf;                                                          \
    return (code);                                                            \
  } while (0)

#define RET(call)                                                             \
  do {                                                                        \
    int _r = (call);                                                          \
    if (_r != B200_OK) return _r;                                             \
  } while (0)

#define LAUNCH(ctx, kern, grid, block, smem, ...)                             \
  do {                                                                        \
    kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);            \
    (ctx)->launches++;                                                        \
  } while (0)

static inline int grid_for(int64_t n, int block, int per_thread = 1) {
  int64_t g = (n + (int64_t)block * per_thread - 1) / ((int64_t)block * per_thread);
  if (g < 1) g = 1;
  const int64_t cap = 148 * 16;   // 148 SMs x 16 resident CTAs: grid-stride beyond
  return (int)(g > cap ? cap : g);
}

// --------------------------------------------------------- device reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  Fixed order -> deterministic.
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();           // protect sh against a previous call
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  return v;
}

// Grid-wide deterministic sum of NV values per block: every block stores its
// partials, the last block to arrive (integer ticket) adds them in a fixed
// order and writes out[k].  No floating-point atomics anywhere.
template <int NV>
__device__ __forceinline__ void grid_sum_finalize(const double (&v)[NV], double* part,
                                                  unsigned* ticket, double* const (&out)[NV],
                                                  bool accumulate = false) {
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
      if (threadIdx.x == 0) *out[k] = accumulate ? *out[k] + s : s;
    }
  }
}

// ------------------------------------------------------------ host helpers
int b200_scan_exclusive_i32(b200_ctx* ctx, const int* in, int* out, int64_t n,
                            int64_t* total_h);
int b200_check_flags(b200_ctx* ctx);   // reads w_flags, maps to an error code
int b200_zero_flags(b200_ctx* ctx);

enum { FLAG_INDEX = 0, FLAG_NONFINITE = 1, FLAG_NONPOS_DIAG = 2, FLAG_OVERFLOW = 3,
       FLAG_NONSYM = 4, FLAG_ZERO_DIAG = 5, FLAG_COUNT = 8 };

// dist.cu (no-ops when nranks == 1)
int b200_halo_exchange(b200_ctx* ctx, double* base);           // padded vector
int b200_halo_exchange_begin(b200_ctx* ctx, double* base);     // on comm_stream
int b200_halo_exchange_end(b200_ctx* ctx);                     // stream waits for it
int b200_allreduce_sum(b200_ctx* ctx, double* dev, int count); // in place, fp64
int b200_allreduce_max_host(b200_ctx* ctx, int64_t* v);
int b200_allreduce_sum_n(b200_ctx* ctx, double* dev, size_t count);
int b200_allreduce_sum_i32(b200_ctx* ctx, int* dev, size_t count);
int b200_halo_exchange_w(b200_ctx* ctx, void* owned, int64_t n, int elem_bytes, int64_t wlo,
                         int64_t whi, int64_t slo, int64_t shi);
int b200_allgather_i64_host(b200_ctx* ctx, const int64_t* mine, int k, int64_t* all);
void b200_comm_release(b200_ctx* ctx);
// solver.cu helpers used by transient.cu / bicgstab.cu
int b200_spmv_dots(b200_ctx* ctx, const double* w_owned, const double* rv_owned,
                   double* out_owned, double* group);
int b200_spmv_dots_rows(b200_ctx* ctx, const double* p_owned, const double* r_owned,
                        double* ap_owned, double* group, int64_t hi0, int64_t lo1);
int b200_launch_init(b200_ctx* ctx, const double* x0, double* xt, double* bt, double* out_b2);
int b200_launch_resid(b200_ctx* ctx, const double* bt, const double* ap, const double* d,
                      double* r, double* p, double* out_rr, double* out_true);
int b200_launch_final(b200_ctx* ctx, const double* xt, double* x);
const double* b200_cur_diag(b200_ctx* ctx);
bool b200_dia_smooth_available(b200_ctx* ctx);
int b200_dia_smooth(b200_ctx* ctx, int mode, const double* p_owned, const double* b_owned,
                    const double* aux, double* d_owned, double* out_owned, double c0, double c1,
                    double c2, double* dot_out);
int b200_bicgstab(b200_ctx* ctx, const double* x0, double rtol, double atol, int64_t maxiter,
                  double* x, int64_t* iters, double* relres, int* info);
// amg.cu
void b200_amg_release(b200_ctx* ctx);
bool b200_amg_applicable(b200_ctx* ctx);
int b200_amg_pcg(b200_ctx* ctx, const double* x0, double rtol, double atol, int64_t maxiter,
                 double* x, int64_t* iters, double* relres, int* info, int* used);
// p2p.cu
void b200_p2p_release(b200_ctx* ctx);
bool b200_p2p_usable(b200_ctx* ctx);
int b200_p2p_iteration(b200_ctx* ctx, double* gin, double* gout, bool true_norm, bool wait_halo,
                       bool rr_is_global, const double* d);
int b200_p2p_check(b200_ctx* ctx);
// assembly.cu
int b200_update_conductance(b200_ctx* ctx, const double* x);
// solver.cu
int b200_linearize(b200_ctx* ctx, const double* x);
int b200_residual_norms(b200_ctx* ctx, const double* x, double* res2, double* x2);
int b200_spmv_csr_raw(b200_ctx* ctx, const int* indptr, const int* indices,
                      const double* data, int64_t n, int64_t col_shift,
                      const double* x, double* y);
