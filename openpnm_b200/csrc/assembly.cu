// assembly.cu -- K1 (conns + g -> canonical CSR Laplacian), K2 (boundary
// conditions), generic COO -> CSR, export.  sm_100a, no floating-point atomics.
//
// Reference behaviour restated here (file:line under /root/reference):
//   openpnm/network/_network.py:274-293   row=[c0;c1] col=[c1;c0] data=[g;g]
//   openpnm/algorithms/_transport.py:113-114  spgr.laplacian(am).astype(float)
//   scipy/sparse/csgraph/_laplacian.py:449-484 + _coo.py:511-547 (setdiag)
//   openpnm/algorithms/_transport.py:145-169  _apply_BCs
//   openpnm/solvers/_scipy.py:13-14           A.tocsr()  (sum duplicates, sort)
#include "common.cuh"

// ===================================================================== scan
#define SCAN_T 256
#define SCAN_I 8
#define SCAN_TILE (SCAN_T * SCAN_I)

__device__ __forceinline__ int warp_incl_scan(int v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of one int per thread across the block; returns exclusive
// prefix, *total gets the block total (valid in all threads)
__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int wsum[32];
  __shared__ int tot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = warp_incl_scan(v);
  __syncthreads();
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    int s = (lane < nw) ? wsum[lane] : 0;
    int si = warp_incl_scan(s);
    wsum[lane] = si - s;
    if (lane == 31) tot = si;
  }
  __syncthreads();
  *total = tot;
  return inc - v + wsum[w];
}

__global__ void __launch_bounds__(SCAN_T) k_scan_tilesum(const int* __restrict__ in, int64_t n,
                                                        long long* __restrict__ tilesum) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_I; ++k) {
    int64_t i = base + (int64_t)k * SCAN_T + threadIdx.x;
    if (i < n) s += in[i];
  }
  int tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) tilesum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_top(long long* tilesum, int64_t ntiles) {
  // single block; sequential over chunks of 1024 tiles
  __shared__ long long carry;
  __shared__ long long wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t c = 0; c < ntiles; c += 1024) {
    int64_t i = c + threadIdx.x;
    long long v = (i < ntiles) ? tilesum[i] : 0;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
      long long s = wsum[lane];
      long long si = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, si, o);
        if (lane >= o) si += t;
      }
      wsum[lane] = si - s;
    }
    __syncthreads();
    long long excl = inc - v + wsum[w] + carry;
    if (i < ntiles) tilesum[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) tilesum[ntiles] = carry;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_apply(const int* __restrict__ in, int* __restrict__ out,
                                                      int64_t n,
                                                      const long long* __restrict__ tilesum,
                                                      int64_t ntiles) {
  // thread owns SCAN_I consecutive items (blocked arrangement)
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_I;
  int v[SCAN_I];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_I; ++k) {
    int64_t i = base + k;
    v[k] = (i < n) ? in[i] : 0;
    s += v[k];
  }
  int tot;
  int excl = block_excl_scan(s, &tot);
  long long run = tilesum[blockIdx.x] + excl;
#pragma unroll
  for (int k = 0; k < SCAN_I; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = (int)run;
    run += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (int)tilesum[ntiles];
}

// out has n+1 entries; out[n] = total.  in may alias out.
int b200_scan_exclusive_i32(b200_ctx* ctx, const int* in, int* out, int64_t n,
                            int64_t* total_h) {
  int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (ntiles < 1) ntiles = 1;
  CK(ctx->w_scan.ensure((size_t)(ntiles + 1) * sizeof(long long)));
  long long* ts = ctx->w_scan.as<long long>();
  LAUNCH(ctx, k_scan_tilesum, (int)ntiles, SCAN_T, 0, in, n, ts);
  LAUNCH(ctx, k_scan_top, 1, 1024, 0, ts, ntiles);
  LAUNCH(ctx, k_scan_apply, (int)ntiles, SCAN_T, 0, in, out, n, ts, ntiles);
  CK(cudaGetLastError());
  if (total_h) {
    long long t = 0;
    CK(cudaMemcpyAsync(&t, ts + ntiles, sizeof t, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *total_h = t;
    if (t >= (1ll << 31)) FAIL(B200_ERR_TOO_LARGE, "nnz %lld exceeds int32 index range", t);
  }
  return B200_OK;
}

// ==================================================================== flags
int b200_zero_flags(b200_ctx* ctx) {
  CK(ctx->w_flags.ensure(FLAG_COUNT * sizeof(int)));
  CK(cudaMemsetAsync(ctx->w_flags.p, 0, FLAG_COUNT * sizeof(int), ctx->stream));
  return B200_OK;
}

int b200_check_flags(b200_ctx* ctx) {
  CK(cudaMemcpyAsync(ctx->flags_h, ctx->w_flags.p, FLAG_COUNT * sizeof(int),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->flags_h[FLAG_INDEX]) FAIL(B200_ERR_INDEX, "pore index out of range");
  if (ctx->flags_h[FLAG_NONFINITE]) FAIL(B200_ERR_NONFINITE, "A or b contains inf/nan values");
  if (ctx->flags_h[FLAG_OVERFLOW]) FAIL(B200_ERR_TOO_LARGE, "index overflow");
  return B200_OK;
}

// ===================================================== K1: entry enumeration
// An "entry" is one off-diagonal COO triple.  For the Laplacian there are 2*Nt
// of them, implied by conns: e < Nt -> (c0, c1), e >= Nt -> (c1, c0); the value
// is -g[e mod Nt].  For a generic COO the triples are explicit.
struct ConnsSrc {
  const longlong2* conns;   // (Nt) pairs
  const double* g;          // (Nt,) symmetric weights, or (Nt,2) row-major when twoway
  int64_t Nt;
  int twoway;               // 1: entry (c0,c1) carries g[t][0], entry (c1,c0) carries g[t][1]
};

__global__ void k_count_conns(ConnsSrc src, int64_t Np, int64_t rb, int64_t re,
                              int* __restrict__ cnt, int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad_idx = false, bad_val = false;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < src.Nt; t += stride) {
    longlong2 c = src.conns[t];
    if (src.twoway) {
      if (!isfinite(src.g[2 * t]) || !isfinite(src.g[2 * t + 1])) bad_val = true;
    } else if (!isfinite(src.g[t])) {
      bad_val = true;
    }
    if (c.x < 0 || c.x >= Np || c.y < 0 || c.y >= Np) { bad_idx = true; continue; }
    if (c.x == c.y) continue;   // self-loop: COO setdiag drops it (_coo.py:520)
    if (c.x >= rb && c.x < re) atomicAdd(&cnt[c.x - rb], 1);
    if (c.y >= rb && c.y < re) atomicAdd(&cnt[c.y - rb], 1);
  }
  if (bad_idx) flags[FLAG_INDEX] = 1;
  if (bad_val) flags[FLAG_NONFINITE] = 1;
}

// cursor[] holds the remaining count per row (copy of cnt); slots are claimed
// from the back.  The claim order is arbitrary (integer atomics) but the
// per-row sort on the 64-bit key (col, throat) below makes the final layout and
// every floating-point sum independent of it.
__global__ void k_fill_conns(ConnsSrc src, int64_t rb, int64_t re,
                             const int* __restrict__ rowofs, int* __restrict__ cursor,
                             unsigned long long* __restrict__ keys) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < src.Nt; t += stride) {
    longlong2 c = src.conns[t];
    if (c.x == c.y) continue;
    if (c.x >= rb && c.x < re) {
      int r = (int)(c.x - rb);
      int slot = atomicSub(&cursor[r], 1) - 1;
      B200_ASSERT(slot >= 0 && rowofs[r] + slot < rowofs[r + 1]);
      keys[(int64_t)rowofs[r] + slot] = ((unsigned long long)c.y << 32) | (unsigned)t;
    }
    if (c.y >= rb && c.y < re) {
      int r = (int)(c.y - rb);
      int slot = atomicSub(&cursor[r], 1) - 1;
      B200_ASSERT(slot >= 0 && rowofs[r] + slot < rowofs[r + 1]);
      // bit 31 of the low word marks the (c1 -> c0) direction of the throat
      keys[(int64_t)rowofs[r] + slot] =
          ((unsigned long long)c.x << 32) | (unsigned)t | 0x80000000u;
    }
  }
}

struct CooSrc {
  const int* row;
  const int* col;
  const double* data;
  int64_t nnz;
};

__global__ void k_count_coo(CooSrc src, int64_t n, int* __restrict__ cnt, int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad_idx = false, bad_val = false;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < src.nnz; e += stride) {
    int r = src.row[e], c = src.col[e];
    if (!isfinite(src.data[e])) bad_val = true;
    if (r < 0 || r >= n || c < 0 || c >= n) { bad_idx = true; continue; }
    atomicAdd(&cnt[r], 1);
  }
  if (bad_idx) flags[FLAG_INDEX] = 1;
  if (bad_val) flags[FLAG_NONFINITE] = 1;
}

__global__ void k_fill_coo(CooSrc src, int64_t n, const int* __restrict__ rowofs,
                           int* __restrict__ cursor, unsigned long long* __restrict__ keys) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < src.nnz; e += stride) {
    int r = src.row[e], c = src.col[e];
    if (r < 0 || r >= n || c < 0 || c >= n) continue;
    int slot = atomicSub(&cursor[r], 1) - 1;
    keys[(int64_t)rowofs[r] + slot] = ((unsigned long long)(unsigned)c << 32) | (unsigned)e;
  }
}

// ------------------------------------------------ per-row sort + unique count
__device__ __forceinline__ void sift_down(unsigned long long* a, int start, int end) {
  int root = start;
  while (2 * root + 1 <= end) {
    int child = 2 * root + 1, sw = root;
    if (a[sw] < a[child]) sw = child;
    if (child + 1 <= end && a[sw] < a[child + 1]) sw = child + 1;
    if (sw == root) return;
    unsigned long long t = a[root]; a[root] = a[sw]; a[sw] = t;
    root = sw;
  }
}

// One thread per row.  Rows of a pore network are short (<= 6 on a cubic
// lattice, ~15 on a Delaunay graph): insertion sort; long rows fall back to an
// in-place heapsort so a hub pore cannot go quadratic.
// extra = 1 reserves a slot for the Laplacian diagonal.
__global__ void k_row_sort_count(const int* __restrict__ rowofs, unsigned long long* __restrict__ keys,
                                 int64_t nrows, int extra, int* __restrict__ ucnt) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const int b = rowofs[r], e = rowofs[r + 1];
    unsigned long long* a = keys + b;
    const int m = e - b;
    if (m <= 32) {
      for (int i = 1; i < m; ++i) {
        unsigned long long k = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > k) { a[j + 1] = a[j]; --j; }
        a[j + 1] = k;
      }
    } else {
      for (int s = (m - 2) / 2; s >= 0; --s) sift_down(a, s, m - 1);
      for (int end = m - 1; end > 0; --end) {
        unsigned long long t = a[end]; a[end] = a[0]; a[0] = t;
        sift_down(a, 0, end - 1);
      }
    }
    int u = 0;
    unsigned prev = 0xffffffffu;
    for (int i = 0; i < m; ++i) {
      unsigned c = (unsigned)(a[i] >> 32);
      if (i == 0 || c != prev) ++u;
      prev = c;
    }
    ucnt[r] = u + extra;
  }
}

// Laplacian emit: off-diagonals -sum(g) per distinct neighbour (duplicates
// added in throat order), diagonal = +column sum (scipy's in-degree Laplacian:
// the weight of the mirrored entry), inserted at its sorted position.
// Weights: w(t, dir) = g[t*gs + dir*gd]; symmetric (gs,gd) = (1,0), two-way
// (Nt,2) weights (2,1).  The entry of row c0 is direction 0, of row c1 direction 1.
__global__ void k_emit_laplacian(const int* __restrict__ rowofs,
                                 const unsigned long long* __restrict__ keys,
                                 const double* __restrict__ g, int gs, int gd, int64_t nrows,
                                 int64_t rb, const int* __restrict__ indptr,
                                 int* __restrict__ indices, double* __restrict__ data,
                                 double* __restrict__ diag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const int b = rowofs[r], e = rowofs[r + 1];
    int o = indptr[r];
    const unsigned self = (unsigned)(r + rb);
    double dsum = 0.0;
    bool diag_done = false;
    int dpos = -1;
    int i = b;
    while (i < e) {
      unsigned c = (unsigned)(keys[i] >> 32);
      double s = 0.0;
      while (i < e && (unsigned)(keys[i] >> 32) == c) {
        const unsigned lo = (unsigned)(keys[i] & 0xffffffffull);
        const int64_t t = lo & 0x7fffffffu;
        const int dir = lo >> 31;
        s += g[t * gs + dir * gd];
        dsum += g[t * gs + (1 - dir) * gd];
        ++i;
      }
      if (!diag_done && c > self) { dpos = o++; diag_done = true; }
      indices[o] = (int)c;
      data[o] = -s;
      ++o;
    }
    if (!diag_done) dpos = o++;
    indices[dpos] = (int)self;
    data[dpos] = dsum;
    diag[r] = dsum;
  }
}

// Generic emit (tocsr semantics): duplicates summed in entry order, explicit
// zeros kept, diagonal is an ordinary entry.  diagpos = -1 when absent.
__global__ void k_emit_coo(const int* __restrict__ rowofs,
                           const unsigned long long* __restrict__ keys,
                           const double* __restrict__ vals, int64_t nrows,
                           const int* __restrict__ indptr, int* __restrict__ indices,
                           double* __restrict__ data, double* __restrict__ diag,
                           int* __restrict__ diagpos) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += stride) {
    const int b = rowofs[r], e = rowofs[r + 1];
    int o = indptr[r];
    double d = 0.0;
    int dp = -1;
    int i = b;
    while (i < e) {
      unsigned c = (unsigned)(keys[i] >> 32);
      double s = 0.0;
      while (i < e && (unsigned)(keys[i] >> 32) == c) {
        s += vals[(unsigned)(keys[i] & 0xffffffffull)];
        ++i;
      }
      if ((int64_t)c == r) { d = s; dp = o; }
      indices[o] = (int)c;
      data[o] = s;
      ++o;
    }
    diag[r] = d;
    diagpos[r] = dp;
  }
}

static int build_laplacian_impl(b200_ctx* ctx, const int64_t* conns, const double* g, int twoway,
                                int64_t Np, int64_t Nt, int64_t row_begin, int64_t row_end,
                                int64_t* nnz) {
  if (!ctx) return B200_ERR_ARG;
  if (Np <= 0 || Nt < 0 || row_begin < 0 || row_end > Np || row_begin >= row_end)
    FAIL(B200_ERR_ARG, "build_laplacian: bad sizes Np=%lld Nt=%lld rows=[%lld,%lld)",
         (long long)Np, (long long)Nt, (long long)row_begin, (long long)row_end);
  if (Np >= (1ll << 31) || 2 * Nt >= (1ll << 31))
    FAIL(B200_ERR_TOO_LARGE, "Np or 2*Nt exceeds the int32 index range");
  CK(cudaSetDevice(ctx->device));
  const int64_t nr = row_end - row_begin;
  ctx->n_global = Np;
  ctx->row_begin = row_begin;
  ctx->row_end = row_end;
  ctx->n = nr;
  ctx->have_pure = false;
  ctx->have_sys = false;
  ctx->lin_valid = false;
  RET(b200_zero_flags(ctx));
  int* flags = ctx->w_flags.as<int>();

  CK(ctx->w_cnt.ensure((size_t)(nr + 1) * sizeof(int)));
  CK(ctx->w_tmp0.ensure((size_t)(nr + 1) * sizeof(int)));   // raw row offsets
  CK(ctx->w_tmp1.ensure((size_t)(nr + 1) * sizeof(int)));   // cursor / unique counts
  int* cnt = ctx->w_cnt.as<int>();
  int* rowofs = ctx->w_tmp0.as<int>();
  int* cur = ctx->w_tmp1.as<int>();
  CK(cudaMemsetAsync(cnt, 0, (size_t)(nr + 1) * sizeof(int), ctx->stream));

  ConnsSrc src{reinterpret_cast<const longlong2*>(conns), g, Nt, twoway};
  const int B = 256;
  LAUNCH(ctx, k_count_conns, grid_for(Nt, B), B, 0, src, Np, row_begin, row_end, cnt, flags);
  CK(cudaMemcpyAsync(cur, cnt, (size_t)nr * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  int64_t raw = 0;
  RET(b200_scan_exclusive_i32(ctx, cnt, rowofs, nr, &raw));
  RET(b200_check_flags(ctx));

  CK(ctx->w_keys.ensure((size_t)(raw > 0 ? raw : 1) * sizeof(unsigned long long)));
  unsigned long long* keys = ctx->w_keys.as<unsigned long long>();
  LAUNCH(ctx, k_fill_conns, grid_for(Nt, B), B, 0, src, row_begin, row_end, rowofs, cur, keys);
  LAUNCH(ctx, k_row_sort_count, grid_for(nr, 128), 128, 0, rowofs, keys, nr, 1, cur);

  CK(ctx->pure_indptr.ensure((size_t)(nr + 1) * sizeof(int)));
  int64_t total = 0;
  RET(b200_scan_exclusive_i32(ctx, cur, ctx->pure_indptr.as<int>(), nr, &total));
  CK(ctx->pure_indices.ensure((size_t)total * sizeof(int)));
  CK(ctx->pure_data.ensure((size_t)total * sizeof(double)));
  CK(ctx->pure_diag.ensure((size_t)nr * sizeof(double)));
  LAUNCH(ctx, k_emit_laplacian, grid_for(nr, 128), 128, 0, rowofs, keys, g, twoway ? 2 : 1,
         twoway ? 1 : 0, nr, row_begin, ctx->pure_indptr.as<int>(), ctx->pure_indices.as<int>(),
         ctx->pure_data.as<double>(), ctx->pure_diag.as<double>());
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->pure_nnz = total;
  ctx->have_pure = true;
  ctx->f_override = false;
  ctx->nonsym = twoway != 0;
  if (nnz) *nnz = total;
  return B200_OK;
}

extern "C" int b200_build_laplacian(b200_ctx* ctx, const int64_t* conns, const double* g,
                                    int64_t Np, int64_t Nt, int64_t row_begin, int64_t row_end,
                                    int64_t* nnz) {
  return build_laplacian_impl(ctx, conns, g, 0, Np, Nt, row_begin, row_end, nnz);
}

extern "C" int b200_build_laplacian2(b200_ctx* ctx, const int64_t* conns, const double* g2,
                                     int64_t Np, int64_t Nt, int64_t* nnz) {
  return build_laplacian_impl(ctx, conns, g2, 1, Np, Nt, 0, Np, nnz);
}

extern "C" int b200_build_laplacian2_rows(b200_ctx* ctx, const int64_t* conns, const double* g2,
                                          int64_t Np, int64_t Nt, int64_t row_begin,
                                          int64_t row_end, int64_t* nnz) {
  return build_laplacian_impl(ctx, conns, g2, 1, Np, Nt, row_begin, row_end, nnz);
}

// diag[i] += add[i] where add is finite (AdvectionDiffusion outflow BC,
// openpnm/algorithms/_advection_diffusion.py:86-105); the diagonal entry of the
// system CSR is patched, so apply_bcs must have kept every diagonal entry
__global__ void k_add_diag(const double* __restrict__ add, int64_t n, int64_t rb,
                           double* __restrict__ diag, const int* __restrict__ diagpos,
                           double* __restrict__ data, int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double a = add[i + rb];
    if (isnan(a)) continue;
    if (!isfinite(a)) { flags[FLAG_NONFINITE] = 1; continue; }
    const double d = diag[i] + a;
    diag[i] = d;
    const int dp = diagpos[i];
    if (dp >= 0) data[dp] = d; else flags[FLAG_ZERO_DIAG] = 1;
  }
}

extern "C" int b200_add_to_diagonal(b200_ctx* ctx, const double* add) {
  if (!ctx || !ctx->have_sys || !add) return B200_ERR_ARG;
  RET(b200_zero_flags(ctx));
  LAUNCH(ctx, k_add_diag, grid_for(ctx->n, 256), 256, 0, add, ctx->n, ctx->row_begin,
         ctx->sys_diag.as<double>(), ctx->sys_diagpos.as<int>(), ctx->sys_data.as<double>(),
         ctx->w_flags.as<int>());
  CK(cudaGetLastError());
  RET(b200_check_flags(ctx));
  if (ctx->flags_h[FLAG_ZERO_DIAG])
    FAIL(B200_ERR_ARG, "add_to_diagonal: a diagonal entry is missing (apply_bcs with keep_zero_diag)");
  ctx->lin_valid = false;
  ctx->sys_epoch++;
  return B200_OK;
}

// ============================================================ generic COO/CSR
static int alloc_system(b200_ctx* ctx, int64_t n, int64_t nnz) {
  CK(ctx->sys_indptr.ensure((size_t)(n + 1) * sizeof(int)));
  CK(ctx->sys_indices.ensure((size_t)(nnz > 0 ? nnz : 1) * sizeof(int)));
  CK(ctx->sys_data.ensure((size_t)(nnz > 0 ? nnz : 1) * sizeof(double)));
  CK(ctx->sys_diag.ensure((size_t)n * sizeof(double)));
  CK(ctx->sys_diagpos.ensure((size_t)n * sizeof(int)));
  CK(ctx->b.ensure((size_t)n * sizeof(double)));
  return B200_OK;
}

// every off-diagonal entry must have an equal mirror entry, else the system is
// flagged non-symmetric (-> BiCGStab instead of CG); binary search in the
// column-sorted mirror row
__global__ void k_check_symmetric(const int* __restrict__ indptr, const int* __restrict__ indices,
                                  const double* __restrict__ data, int64_t n,
                                  int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
      const int j = indices[e];
      if ((int64_t)j == r) continue;
      int lo = indptr[j], hi = indptr[j + 1];
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)indices[mid] < r) lo = mid + 1; else hi = mid;
      }
      const double a = data[e];
      if (lo >= indptr[j + 1] || (int64_t)indices[lo] != r) {
        if (a != 0.0) bad = true;
      } else {
        const double b = data[lo];
        if (fabs(a - b) > 1e-12 * fmax(fabs(a), fabs(b))) bad = true;
      }
    }
  }
  if (bad) flags[FLAG_NONSYM] = 1;
}

static int detect_nonsymmetric(b200_ctx* ctx) {
  RET(b200_zero_flags(ctx));
  LAUNCH(ctx, k_check_symmetric, grid_for(ctx->n, 128), 128, 0, ctx->sys_indptr.as<int>(),
         ctx->sys_indices.as<int>(), ctx->sys_data.as<double>(), ctx->n, ctx->w_flags.as<int>());
  RET(b200_check_flags(ctx));
  ctx->nonsym = ctx->flags_h[FLAG_NONSYM] != 0;
  return B200_OK;
}

extern "C" int b200_load_coo(b200_ctx* ctx, const int32_t* row, const int32_t* col,
                             const double* data, int64_t nnz_in, int64_t n, int64_t* nnz) {
  if (!ctx) return B200_ERR_ARG;
  if (n <= 0 || nnz_in < 0) FAIL(B200_ERR_ARG, "load_coo: bad sizes");
  if (n >= (1ll << 31) || nnz_in >= (1ll << 31)) FAIL(B200_ERR_TOO_LARGE, "load_coo: too large");
  CK(cudaSetDevice(ctx->device));
  if (ctx->nranks == 1) { ctx->n_global = n; ctx->row_begin = 0; ctx->row_end = n; }
  ctx->n = n;
  ctx->have_sys = false;
  ctx->lin_valid = false;
  RET(b200_zero_flags(ctx));
  int* flags = ctx->w_flags.as<int>();
  CK(ctx->w_cnt.ensure((size_t)(n + 1) * sizeof(int)));
  CK(ctx->w_tmp0.ensure((size_t)(n + 1) * sizeof(int)));
  CK(ctx->w_tmp1.ensure((size_t)(n + 1) * sizeof(int)));
  int* cnt = ctx->w_cnt.as<int>();
  int* rowofs = ctx->w_tmp0.as<int>();
  int* cur = ctx->w_tmp1.as<int>();
  CK(cudaMemsetAsync(cnt, 0, (size_t)(n + 1) * sizeof(int), ctx->stream));
  CooSrc src{row, col, data, nnz_in};
  const int B = 256;
  LAUNCH(ctx, k_count_coo, grid_for(nnz_in, B), B, 0, src, n, cnt, flags);
  CK(cudaMemcpyAsync(cur, cnt, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  int64_t raw = 0;
  RET(b200_scan_exclusive_i32(ctx, cnt, rowofs, n, &raw));
  RET(b200_check_flags(ctx));
  CK(ctx->w_keys.ensure((size_t)(raw > 0 ? raw : 1) * sizeof(unsigned long long)));
  unsigned long long* keys = ctx->w_keys.as<unsigned long long>();
  LAUNCH(ctx, k_fill_coo, grid_for(nnz_in, B), B, 0, src, n, rowofs, cur, keys);
  LAUNCH(ctx, k_row_sort_count, grid_for(n, 128), 128, 0, rowofs, keys, n, 0, cur);
  CK(ctx->sys_indptr.ensure((size_t)(n + 1) * sizeof(int)));
  int64_t total = 0;
  RET(b200_scan_exclusive_i32(ctx, cur, ctx->sys_indptr.as<int>(), n, &total));
  RET(alloc_system(ctx, n, total));
  LAUNCH(ctx, k_emit_coo, grid_for(n, 128), 128, 0, rowofs, keys, data, n,
         ctx->sys_indptr.as<int>(), ctx->sys_indices.as<int>(), ctx->sys_data.as<double>(),
         ctx->sys_diag.as<double>(), ctx->sys_diagpos.as<int>());
  CK(cudaMemsetAsync(ctx->b.p, 0, (size_t)n * sizeof(double), ctx->stream));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->sys_nnz = total;
  ctx->have_sys = true;
  ctx->sys_generic = true;
  ctx->sys_epoch++;
  RET(detect_nonsymmetric(ctx));
  if (nnz) *nnz = total;
  return B200_OK;
}

__global__ void k_csr_diag(const int* __restrict__ indptr, const int* __restrict__ indices,
                           const double* __restrict__ data, int64_t n, int64_t rb,
                           double* __restrict__ diag, int* __restrict__ diagpos,
                           int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false, badidx = false;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    double d = 0.0;
    int dp = -1;
    for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
      if (!isfinite(data[e])) bad = true;
      if (indices[e] < 0) badidx = true;
      if ((int64_t)indices[e] == r + rb) { d += data[e]; if (dp < 0) dp = e; }
    }
    diag[r] = d;
    diagpos[r] = dp;
  }
  if (bad) flags[FLAG_NONFINITE] = 1;
  if (badidx) flags[FLAG_INDEX] = 1;
}

extern "C" int b200_load_csr(b200_ctx* ctx, const int32_t* indptr, const int32_t* indices,
                             const double* data, int64_t n, int64_t nnz) {
  if (!ctx) return B200_ERR_ARG;
  if (n <= 0 || nnz < 0 || nnz >= (1ll << 31)) FAIL(B200_ERR_ARG, "load_csr: bad sizes");
  CK(cudaSetDevice(ctx->device));
  if (ctx->nranks == 1) { ctx->n_global = n; ctx->row_begin = 0; ctx->row_end = n; }
  ctx->n = n;
  RET(b200_zero_flags(ctx));
  RET(alloc_system(ctx, n, nnz));
  CK(cudaMemcpyAsync(ctx->sys_indptr.p, indptr, (size_t)(n + 1) * sizeof(int),
                     cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->sys_indices.p, indices, (size_t)nnz * sizeof(int),
                     cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->sys_data.p, data, (size_t)nnz * sizeof(double),
                     cudaMemcpyDeviceToDevice, ctx->stream));
  LAUNCH(ctx, k_csr_diag, grid_for(n, 128), 128, 0, ctx->sys_indptr.as<int>(),
         ctx->sys_indices.as<int>(), ctx->sys_data.as<double>(), n, ctx->row_begin,
         ctx->sys_diag.as<double>(), ctx->sys_diagpos.as<int>(), ctx->w_flags.as<int>());
  CK(cudaMemsetAsync(ctx->b.p, 0, (size_t)n * sizeof(double), ctx->stream));
  RET(b200_check_flags(ctx));
  ctx->sys_nnz = nnz;
  ctx->have_sys = true;
  ctx->sys_generic = true;
  ctx->lin_valid = false;
  ctx->sys_epoch++;
  return detect_nonsymmetric(ctx);
}

extern "C" int b200_set_b(b200_ctx* ctx, const double* b) {
  if (!ctx || !ctx->have_sys) return B200_ERR_ARG;
  CK(cudaMemcpyAsync(ctx->b.p, b, (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToDevice,
                     ctx->stream));
  ctx->lin_valid = false;
  return B200_OK;
}

extern "C" int b200_get_b(b200_ctx* ctx, double* b) {
  if (!ctx || !ctx->have_sys) return B200_ERR_ARG;
  const double* src = (ctx->lin_valid && ctx->srcs.n > 0) ? ctx->b_eff.as<double>()
                                                          : ctx->b.as<double>();
  CK(cudaMemcpyAsync(b, src, (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToDevice,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

// ======================================================= K2: boundary values
// deterministic sum of an array (for f = mean of the pure diagonal)
__global__ void __launch_bounds__(256) k_sum(const double* __restrict__ a, int64_t n,
                                             double* part, unsigned* ticket, double* out) {
  double s = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) s += a[i];
  double v[1] = {s};
  double* o[1] = {out};
  grid_sum_finalize<1>(v, part, ticket, o);
}

static int ensure_reduce_ws(b200_ctx* ctx) {
  CK(ctx->w_part.ensure((size_t)4 * 148 * 16 * sizeof(double)));
  if (!ctx->w_ticket.p) {
    CK(ctx->w_ticket.ensure(64));
    CK(cudaMemsetAsync(ctx->w_ticket.p, 0, 64, ctx->stream));
  }
  CK(ctx->scal.ensure(SC_COUNT * sizeof(double)));
  return B200_OK;
}

extern "C" int b200_pure_diag_sum(b200_ctx* ctx, double* sum, int64_t* count) {
  if (!ctx || !ctx->have_pure) return B200_ERR_ARG;
  RET(ensure_reduce_ws(ctx));
  double* out = ctx->scal.as<double>() + SC_AUX0;
  LAUNCH(ctx, k_sum, grid_for(ctx->n, 256, 4), 256, 0, ctx->pure_diag.as<double>(), ctx->n,
         ctx->w_part.as<double>(), ctx->w_ticket.as<unsigned>(), out);
  CK(cudaMemcpyAsync(ctx->scal_h + SC_AUX0, out, sizeof(double), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (sum) *sum = ctx->scal_h[SC_AUX0];
  if (count) *count = ctx->n;
  return B200_OK;
}

extern "C" int b200_set_diag_mean(b200_ctx* ctx, double f) {
  if (!ctx) return B200_ERR_ARG;
  ctx->f = f;
  ctx->f_override = true;
  return B200_OK;
}

// Row pass 1: count surviving entries and form b.
//   value-BC row i : only the diagonal f survives (if f != 0); b_i = v_i f
//   other row i    : keep (i,j) when j has no value BC and a_ij != 0
//                    b_i = (rate_i or 0) - sum_{j in BC} a_ij v_j
// (_transport.py:149-169; eliminate_zeros() drops explicit zeros, which also
//  removes zero-conductance throats.)
__global__ void k_bc_count(const int* __restrict__ indptr, const int* __restrict__ indices,
                           const double* __restrict__ data, int64_t n, int64_t rb,
                           const double* __restrict__ bcv, const double* __restrict__ bcr,
                           double f, int keep_zero_diag, int* __restrict__ cnt,
                           double* __restrict__ b, int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const int64_t gi = r + rb;
    const double vi = bcv ? bcv[gi] : nan("");
    int c = 0;
    double bb;
    if (isfinite(vi)) {
      c = (f != 0.0 || keep_zero_diag) ? 1 : 0;
      bb = vi * f;
    } else {
      bb = 0.0;
      if (bcr) { double q = bcr[gi]; if (isfinite(q)) bb = q; }
      double acc = 0.0;
      for (int e = indptr[r]; e < indptr[r + 1]; ++e) {
        const int j = indices[e];
        const double a = data[e];
        const double vj = bcv ? bcv[j] : nan("");
        if (isfinite(vj)) acc += a * vj;
        else if (a != 0.0 || (keep_zero_diag && (int64_t)j == gi)) ++c;
      }
      bb -= acc;
    }
    if (!isfinite(bb)) bad = true;
    cnt[r] = c;
    b[r] = bb;
  }
  if (bad) flags[FLAG_NONFINITE] = 1;
}

// Row pass 2, warp-cooperative: a warp takes 32 consecutive rows, streams their
// entries coalesced (32 lanes = 32 consecutive entries), finds each entry's row
// by a 5-step binary search over the 32 row pointers held one per lane
// (shuffles), evaluates the keep rule and compacts order-preservingly with
// ballot/popc -- the output segment of the chunk is contiguous, so the writes
// are coalesced as well.  Integer work only; same result as a per-row walk.
__global__ void __launch_bounds__(256) k_bc_emit(const int* __restrict__ indptr,
                                                 const int* __restrict__ indices,
                                                 const double* __restrict__ data, int64_t n,
                                                 int64_t rb, const double* __restrict__ bcv,
                                                 double f, int keep_zero_diag,
                                                 const int* __restrict__ optr,
                                                 int* __restrict__ oidx, double* __restrict__ odat,
                                                 double* __restrict__ odiag,
                                                 int* __restrict__ odiagpos,
                                                 int* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t nchunks = (n + 31) >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  bool bad = false;
  for (int64_t ch = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ch < nchunks;
       ch += nwarps) {
    const int64_t r0 = ch << 5;
    const int64_t r = r0 + lane;
    const int my_ptr = indptr[r < n ? r : n];
    const int e_begin = __shfl_sync(0xffffffffu, my_ptr, 0);
    const int64_t rlast = r0 + 32 < n ? r0 + 32 : n;
    const int e_end = indptr[rlast];
    int out = optr[r0];
    bool row_bc = false;
    if (r < n) {
      const double vi = bcv ? bcv[r + rb] : nan("");
      row_bc = isfinite(vi);
      odiag[r] = row_bc ? f : 0.0;      // overwritten below when the diagonal entry survives
      odiagpos[r] = -1;
    }
    const unsigned bc_mask = __ballot_sync(0xffffffffu, row_bc);
    __syncwarp();
    for (int e0 = e_begin; e0 < e_end; e0 += 32) {
      const int e = e0 + lane;
      const bool valid = e < e_end;
      int j = 0;
      double a = 0.0;
      if (valid) { j = indices[e]; a = data[e]; }
      // row of e: the last lane whose row pointer is <= e
      int lo = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int cand = lo + step;
        const int pc = __shfl_sync(0xffffffffu, my_ptr, cand & 31);
        if (cand < 32 && pc <= e) lo = cand;
      }
      const int64_t row = r0 + lo;
      const bool rbc = (bc_mask >> lo) & 1u;
      const bool isdiag = ((int64_t)j == row + rb);
      bool keep = false;
      double val = a;
      if (valid) {
        if (rbc) {
          keep = isdiag && (f != 0.0 || keep_zero_diag);
          val = f;
        } else {
          const double vj = bcv ? bcv[j] : nan("");
          keep = !isfinite(vj) && (a != 0.0 || (keep_zero_diag && isdiag));
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int pos = out + __popc(m & lt_mask);
        if (!isfinite(val)) bad = true;
        oidx[pos] = j;
        odat[pos] = val;
        if (isdiag) { odiag[row] = val; odiagpos[row] = pos; }
      }
      out += __popc(m);
    }
    __syncwarp();
  }
  if (bad) flags[FLAG_NONFINITE] = 1;
}

extern "C" int b200_apply_bcs(b200_ctx* ctx, const double* bc_value, const double* bc_rate,
                              int keep_zero_diag, double* f_out, int64_t* nnz) {
  if (!ctx || !ctx->have_pure) {
    if (ctx) ctx->err = "apply_bcs: build_laplacian first";
    return B200_ERR_ARG;
  }
  CK(cudaSetDevice(ctx->device));
  const int64_t n = ctx->n;
  RET(b200_zero_flags(ctx));
  if (!ctx->f_override) {
    double s = 0;
    RET(b200_pure_diag_sum(ctx, &s, nullptr));
    ctx->f = s / (double)n;
  }
  const double f = ctx->f;
  if (!isfinite(f)) FAIL(B200_ERR_NONFINITE, "A or b contains inf/nan values");
  CK(ctx->w_cnt.ensure((size_t)(n + 1) * sizeof(int)));
  CK(ctx->b.ensure((size_t)n * sizeof(double)));
  CK(ctx->sys_indptr.ensure((size_t)(n + 1) * sizeof(int)));
  int* cnt = ctx->w_cnt.as<int>();
  LAUNCH(ctx, k_bc_count, grid_for(n, 128), 128, 0, ctx->pure_indptr.as<int>(),
         ctx->pure_indices.as<int>(), ctx->pure_data.as<double>(), n, ctx->row_begin, bc_value,
         bc_rate, f, keep_zero_diag, cnt, ctx->b.as<double>(), ctx->w_flags.as<int>());
  int64_t total = 0;
  RET(b200_scan_exclusive_i32(ctx, cnt, ctx->sys_indptr.as<int>(), n, &total));
  RET(alloc_system(ctx, n, total));
  LAUNCH(ctx, k_bc_emit, grid_for(n, 256 / 32), 256, 0, ctx->pure_indptr.as<int>(),
         ctx->pure_indices.as<int>(), ctx->pure_data.as<double>(), n, ctx->row_begin, bc_value, f,
         keep_zero_diag, ctx->sys_indptr.as<int>(), ctx->sys_indices.as<int>(),
         ctx->sys_data.as<double>(), ctx->sys_diag.as<double>(), ctx->sys_diagpos.as<int>(),
         ctx->w_flags.as<int>());
  CK(cudaGetLastError());
  RET(b200_check_flags(ctx));
  ctx->sys_nnz = total;
  ctx->have_sys = true;
  ctx->sys_generic = false;
  ctx->lin_valid = false;
  ctx->sys_epoch++;
  if (f_out) *f_out = f;
  if (nnz) *nnz = total;
  return B200_OK;
}

// ================================================================== export
extern "C" int b200_csr_shape(b200_ctx* ctx, int which, int64_t* nrows, int64_t* nnz) {
  if (!ctx) return B200_ERR_ARG;
  if (which == B200_MAT_PURE) {
    if (!ctx->have_pure) FAIL(B200_ERR_ARG, "no pure matrix");
    if (nrows) *nrows = ctx->n;
    if (nnz) *nnz = ctx->pure_nnz;
  } else {
    if (!ctx->have_sys) FAIL(B200_ERR_ARG, "no system matrix");
    if (nrows) *nrows = ctx->n;
    if (nnz) *nnz = ctx->sys_nnz;
  }
  return B200_OK;
}

extern "C" int b200_export_csr(b200_ctx* ctx, int which, int32_t* indptr, int32_t* indices,
                               double* data) {
  if (!ctx) return B200_ERR_ARG;
  const DevBuf *ip, *ix, *dt;
  int64_t nnz;
  if (which == B200_MAT_PURE) {
    if (!ctx->have_pure) FAIL(B200_ERR_ARG, "no pure matrix");
    ip = &ctx->pure_indptr; ix = &ctx->pure_indices; dt = &ctx->pure_data; nnz = ctx->pure_nnz;
  } else {
    if (!ctx->have_sys) FAIL(B200_ERR_ARG, "no system matrix");
    ip = &ctx->sys_indptr; ix = &ctx->sys_indices; dt = &ctx->sys_data; nnz = ctx->sys_nnz;
  }
  if (indptr)
    CK(cudaMemcpyAsync(indptr, ip->p, (size_t)(ctx->n + 1) * sizeof(int), cudaMemcpyDeviceToDevice,
                       ctx->stream));
  if (indices)
    CK(cudaMemcpyAsync(indices, ix->p, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToDevice,
                       ctx->stream));
  if (data)
    CK(cudaMemcpyAsync(data, dt->p, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToDevice,
                       ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}


// ===================================== conductance models on the device (f-3)
// Replaces, inside the Picard loop, the host round trip of `_update_iterative_props`
// (openpnm/algorithms/_algorithm.py:69-91) for the common chain
//   quantity X -> pore transport property  P = sum_k a_k X^k   (models/misc: polynomial)
//              -> throat value of it        Pt = given array, or mean / min / max of the
//                                           two pores (models/misc: from_neighbor_pores)
//              -> conduit conductance       generic_hydraulic
//                                           (models/physics/hydraulic_conductance/_funcs.py:37-53:
//                                            g_i = F_i / mu_i, series sum) or _poisson_conductance
//                                           (models/physics/_utils.py:46-61: g_i = D_i F_i;
//                                            generic_diffusive / thermal / electrical)
// followed by the re-assembly (K1) and the BCs (K2).  Nothing leaves the device.
struct PolyArg { double a[8]; int n; };

__global__ void k_conduit_conductance(const longlong2* __restrict__ conns, int64_t Nt,
                                      const double* __restrict__ x, PolyArg poly, int kind,
                                      int interp, const double* __restrict__ throat_prop,
                                      const double* __restrict__ sf, int sf_cols,
                                      double* __restrict__ g) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < Nt; t += stride) {
    const longlong2 c = conns[t];
    double P[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const double X = x[s == 0 ? c.x : c.y];
      double v = poly.a[poly.n - 1];
      for (int k = poly.n - 2; k >= 0; --k) v = v * X + poly.a[k];     // Horner
      P[s] = v;
    }
    double Pt;
    if (throat_prop) Pt = throat_prop[t];
    else if (interp == 1) Pt = fmin(P[0], P[1]);
    else if (interp == 2) Pt = fmax(P[0], P[1]);
    else Pt = 0.5 * (P[0] + P[1]);
    double out;
    if (sf_cols == 3) {
      const double F1 = sf[3 * t], Ft = sf[3 * t + 1], F2 = sf[3 * t + 2];
      double g1, gt, g2;
      if (kind == B200_COND_HYDRAULIC) { g1 = F1 / P[0]; gt = Ft / Pt; g2 = F2 / P[1]; }
      else { g1 = P[0] * F1; gt = Pt * Ft; g2 = P[1] * F2; }
      out = 1.0 / (1.0 / g1 + 1.0 / gt + 1.0 / g2);
    } else {
      // one factor for the whole conduit: F1 = F2 = inf in generic_hydraulic, Dt F in
      // _poisson_conductance
      out = (kind == B200_COND_HYDRAULIC) ? 1.0 / (1.0 / (sf[t] / Pt)) : Pt * sf[t];
    }
    g[t] = out;
  }
}

extern "C" int b200_set_conductance_model(b200_ctx* ctx, int kind, const int64_t* conns,
                                          int64_t Nt, int64_t Np, const double* poly_h, int npoly,
                                          int interp, const double* throat_prop,
                                          const double* size_factors, int sf_cols,
                                          const double* bc_value, const double* bc_rate,
                                          int keep_zero_diag) {
  if (!ctx) return B200_ERR_ARG;
  if ((kind != B200_COND_HYDRAULIC && kind != B200_COND_POISSON) || !conns || !poly_h ||
      npoly < 1 || npoly > 8 || interp < 0 || interp > 2 || !size_factors ||
      (sf_cols != 1 && sf_cols != 3) || Nt <= 0 || Np <= 0)
    FAIL(B200_ERR_ARG, "set_conductance_model: bad arguments");
  if (ctx->nranks > 1) FAIL(B200_ERR_ARG, "set_conductance_model: single GPU only");
  GModel& m = ctx->gm;
  m.active = true;
  m.kind = kind;
  m.interp = interp;
  m.npoly = npoly;
  for (int k = 0; k < 8; ++k) m.poly[k] = k < npoly ? poly_h[k] : 0.0;
  m.conns = conns;
  m.Nt = Nt;
  m.Np = Np;
  m.throat_prop = throat_prop;
  m.sf = size_factors;
  m.sf_cols = sf_cols;
  m.bcv = bc_value;
  m.bcr = bc_rate;
  m.keep_zero_diag = keep_zero_diag;
  return B200_OK;
}

extern "C" int b200_clear_conductance_model(b200_ctx* ctx) {
  if (!ctx) return B200_ERR_ARG;
  ctx->gm.active = false;
  return B200_OK;
}

// g(x) -> K1 -> K2 (x: device, Np entries).  Also usable on its own: the Picard loop
// calls it before every linearisation while a model is set.
extern "C" int b200_update_conductance(b200_ctx* ctx, const double* x) {
  if (!ctx || !x) return B200_ERR_ARG;
  GModel& m = ctx->gm;
  if (!m.active) FAIL(B200_ERR_ARG, "update_conductance: no model set");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->gm_g.ensure((size_t)m.Nt * sizeof(double)));
  PolyArg pa;
  for (int k = 0; k < 8; ++k) pa.a[k] = m.poly[k];
  pa.n = m.npoly;
  LAUNCH(ctx, k_conduit_conductance, grid_for(m.Nt, 256), 256, 0,
         reinterpret_cast<const longlong2*>(m.conns), m.Nt, x, pa, m.kind, m.interp,
         m.throat_prop, m.sf, m.sf_cols, ctx->gm_g.as<double>());
  CK(cudaGetLastError());
  int64_t nnz = 0;
  RET(build_laplacian_impl(ctx, m.conns, ctx->gm_g.as<double>(), 0, m.Np, m.Nt, 0, m.Np, &nnz));
  double f = 0.0;
  return b200_apply_bcs(ctx, m.bcv, m.bcr, m.keep_zero_diag, &f, &nnz);
}

extern "C" int b200_conductance_values(b200_ctx* ctx, double* g_out) {
  if (!ctx || !g_out) return B200_ERR_ARG;
  if (!ctx->gm.active || !ctx->gm_g.p) FAIL(B200_ERR_ARG, "conductance_values: no model evaluated");
  CK(cudaMemcpyAsync(g_out, ctx->gm_g.p, (size_t)ctx->gm.Nt * sizeof(double),
                     cudaMemcpyDeviceToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
