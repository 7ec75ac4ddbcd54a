// amg.cu -- aggregation multigrid preconditioner behind the same solver entry
// point ("better preconditioner behind the same solver
// API"; the reference's analogue is PyamgRugeStubenSolver,
// openpnm/solvers/_pyamg.py:10-17).  Jacobi-PCG needs O(N) iterations on an N^3
// lattice (3256 at 400^3); this needs ~30 at any size.
//
// Algorithm (plain aggregation + K-cycle, after Notay's AGMG, restated for the GPU):
//   setup, per level: three passes (four on the fine level) of pairwise matching by handshaking -- every
//     unmatched row proposes to its strongest unmatched neighbour (largest -a_ij,
//     smaller index on ties), mutual proposals match; leftovers join the
//     aggregate of their strongest matched neighbour; rows without negative
//     off-diagonals (boundary pores after _apply_BCs) stay off the coarse grids.
//     Each pass is followed by the Galerkin product A_c = P^T A P with the
//     piecewise-constant P, built per coarse row (gather, insertion sort by
//     column, duplicates summed in sorted order -> no floating-point atomics,
//     bit-repeatable).  Aggregates of ~8-10 rows per level.
//   cycle: Chebyshev(2) pre/post smoothing on D^-1 A over [lmax/8, lmax], lmax
//     from the row-sum bound; two flexible-CG inner iterations on the two finest
//     coarse problems (K-cycle), V-cycle below; the coarsest level (<= 2048 rows)
//     is solved by CG inside one thread block.
//   outer: flexible CG(1) on the Jacobi-scaled fine operator (the DIA / CSR-stream
//     kernel of solver.cu does every fine-level product), stopping rule as
//     everywhere: ||b - A x||_2 <= max(rtol ||b||_2, atol), confirmed with an
//     explicitly recomputed residual.
// Symmetric systems.  On slabs (one rank per GPU, row-block ownership as in
// openpnm/solvers/_petsc.py:117-128) the hierarchy is ONE multigrid for the whole
// problem: the matching only sees the columns a rank owns, so aggregates never cross a
// rank boundary and every level stays row-partitioned by the same slabs, while the
// Galerkin products keep the couplings between coarse rows of neighbouring ranks (the
// aggregation map of the halo columns is exchanged with the two neighbours).  Every
// product on such a level is preceded by a halo exchange of its operand (contiguous
// windows: coarse ids follow the fine row order), dot products are all-reduced.  Once a
// level has at most B200_AMG_REPL rows globally (default 2^19) it is gathered and it and
// everything below it are built and applied redundantly on every rank: no communication
// below that level except the one gather of the restricted residual per visit.
// tests/models/amg_slabs.py + amg_dist_model.py are the executable specification.
// Environment knobs for experiments: B200_AMG_KLEVELS (0..3, default 2),
// B200_AMG_PASSES0 (pairwise passes on the fine level, default 4),
// B200_AMG_F32=0 (fp64 coarse products), B200_AMG_MAXITER (solver.cu: iteration cap
// before the hand-over to Jacobi-PCG, default 300), B200_AMG_REPL (rows below which a
// level of a slab run is replicated), B200_AMG_SLABS=0 (slab runs stay with Jacobi-PCG).
#include <math.h>
#include <stdlib.h>
#include <time.h>

#include "common.cuh"

#define AMG_T 256
#define AMG_ROUNDS 5
#define AMG_PASSES 3
#define AMG_COARSEST 2048
#define AMG_MAX_LEVELS 12
#define AMG_KLEVELS_DEFAULT 2
#define AMG_KLEVELS_MAX 3
#define AMG_CHEB_RATIO 8.0
#define AMG_COARSE_ITERS 30

struct CsrOut {
  DevBuf indptr, indices, data, dinv;
  int64_t n = 0, nnz = 0;
  // row-partitioned output of a pass in a slab run (columns are local signed ids:
  // < 0 rank-1's rows, >= n rank+1's rows)
  int64_t ng = 0, goff = 0, nnzg = 0;
  int wlo = 0, whi = 0, slo = 0, shi = 0;
};

struct AmgLevel {
  int64_t n = 0, nnz = 0, nc = 0;       // rows / entries held by this rank; nc: rows of the next level
  CsrOut own;                           // levels >= 1: the rows this rank built (last pass's output)
  CsrOut gath;                          // replicated level of a slab run: the gathered matrix
  const double* dv = nullptr;           // 1 / diagonal of whichever of the two is live
  DevBuf dataf;                         // fp32 copy of the values for the cycle's products
  const int* ip = nullptr;
  const int* ix = nullptr;
  const double* av = nullptr;
  DevBuf agg, cptr, cmem;               // aggregation to the next level + member lists
  DevBuf x, b, ax, dir;                 // cycle vectors (levels >= 1); x is [wlo | n | whi]
  DevBuf kb, kc1, kv1, kv2, ks;         // K-cycle buffers and scalars (levels 1..AMG_KLEVELS); kc1 padded like x
  DevBuf cgw;                           // coarsest: r, p, q
  double lmax = 2.0;
  double c0 = 0, c1 = 0, c2 = 0;        // Chebyshev(2) coefficients
  // slab runs: a level is either row-partitioned (dist) or replicated on every rank
  bool dist = false;
  int64_t ng = 0, goff = 0;             // global rows, global index of the first owned row
  int64_t nnzg = 0;                     // global entries
  int wlo = 0, whi = 0, slo = 0, shi = 0;   // halo entries received from / sent to rank-1, rank+1
  bool next_repl = false;               // dist level whose coarse level is replicated
  int64_t coff = 0;                     // offset of this rank's coarse rows in the next level
  double* xo() const { return x.as<double>() + wlo; }      // first owned entry of the padded vectors
  double* kc1o() const { return kc1.as<double>() + wlo; }
};

// the matrix a pairwise pass works on
struct PassMat {
  int n = 0;
  const int* ip = nullptr;
  const int* ix = nullptr;
  const double* av = nullptr;
  const double* dg = nullptr;           // separate diagonal (level 0: the linearised one)
  int shift = 0;                        // column - shift = local row index
  bool dist = false;
  int64_t ng = 0;
  int wlo = 0, whi = 0, slo = 0, shi = 0;
};

struct AmgState {
  std::vector<AmgLevel> lv;      // buffers persist across setups; nlev of them are live
  size_t nlev = 0;
  uint64_t epoch = ~0ull, lin_epoch = ~0ull;
  bool valid = false;
  // setup temporaries, kept so that a rebuild of the same size allocates nothing
  DevBuf t_match, t_prop, t_ids, t_cptr, t_cmem, t_tc, t_tv, t_aggpass, t_cnt, t_cur;
  DevBuf t_aggg, t_minmax, t_gcnt;      // slab runs: global ids of the halo'd map, column range, gather
  CsrOut t_csr[2];
  // level 0 lives in the Jacobi-scaled, padded layout of the PCG operator
  DevBuf z0, z1, ax0, dir0, winv, fd, fq;
  DevBuf lmax_dev;
};

// tuning knob for experiments (B200_AMG_KLEVELS=0..3): levels whose coarse problem
// gets two inner FCG iterations instead of one cycle
static int amg_klevels() {
  static int k = -1;
  if (k < 0) {
    const char* e = getenv("B200_AMG_KLEVELS");
    k = (e && e[0] >= '0' && e[0] <= '0' + AMG_KLEVELS_MAX) ? e[0] - '0' : AMG_KLEVELS_DEFAULT;
  }
  return k;
}
#define AMG_KLEVELS amg_klevels()

// pairwise passes on the fine level (B200_AMG_PASSES0=1..6): the fine level runs
// at 85 % of the HBM peak on its DIA operator, the coarse CSR levels at ~40 %, so
// the first coarsening is one pass more aggressive than the others -- a few more
// iterations, half the coarse-level work
#define AMG_PASSES0_DEFAULT 4
static int amg_passes0() {
  static int k = -1;
  if (k < 0) {
    const char* e = getenv("B200_AMG_PASSES0");
    k = (e && e[0] >= '1' && e[0] <= '6') ? e[0] - '0' : AMG_PASSES0_DEFAULT;
  }
  return k;
}

// B200_AMG_F32=0 keeps the coarse operators in fp64 (A/B experiments)
static bool amg_f32() {
  static int k = -1;
  if (k < 0) {
    const char* e = getenv("B200_AMG_F32");
    k = (e && e[0] == '0') ? 0 : 1;
  }
  return k != 0;
}

// slab runs: a level with at most this many rows (globally) is replicated on every rank
// (read at every setup, so a test can switch it inside one process)
static int64_t amg_repl_rows() {
  const char* e = getenv("B200_AMG_REPL");
  const long long v = e ? atoll(e) : 0;
  return v > 0 ? v : (1ll << 19);
}

static AmgState* amg_of(b200_ctx* ctx) {
  if (!ctx->amg) ctx->amg = new AmgState();
  return static_cast<AmgState*>(ctx->amg);
}

void b200_amg_release(b200_ctx* ctx) {
  if (!ctx->amg) return;
  AmgState* A = static_cast<AmgState*>(ctx->amg);
  for (AmgLevel& L : A->lv)
    for (DevBuf* b : {&L.own.indptr, &L.own.indices, &L.own.data, &L.own.dinv, &L.gath.indptr,
                      &L.gath.indices, &L.gath.data, &L.gath.dinv, &L.agg, &L.cptr, &L.cmem, &L.x,
                      &L.b, &L.ax, &L.dir, &L.kb, &L.kc1, &L.kv1, &L.kv2, &L.ks, &L.cgw, &L.dataf})
      b->release();
  for (DevBuf* b : {&A->z0, &A->z1, &A->ax0, &A->dir0, &A->winv, &A->fd, &A->fq, &A->lmax_dev,
                    &A->t_match, &A->t_prop, &A->t_ids, &A->t_cptr, &A->t_cmem, &A->t_tc, &A->t_tv,
                    &A->t_aggpass, &A->t_cnt, &A->t_cur, &A->t_aggg, &A->t_minmax, &A->t_gcnt})
    b->release();
  for (CsrOut* c : {&A->t_csr[0], &A->t_csr[1]})
    for (DevBuf* b : {&c->indptr, &c->indices, &c->data, &c->dinv}) b->release();
  delete A;
  ctx->amg = nullptr;
}

// ================================================================== matching
// one thread per row; CSR columns ascending, so the first maximum is the
// smallest index among equal weights
// (shift: columns are global ids in a slab; entries that point outside the owned
// rows [shift, shift + n) are not part of this rank's hierarchy)
__global__ void __launch_bounds__(AMG_T) k_amg_propose(const int* __restrict__ ip,
                                                       const int* __restrict__ ix,
                                                       const double* __restrict__ av, int n,
                                                       int shift,
                                                       const int* __restrict__ match,
                                                       int* __restrict__ prop) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int best = -1;
    if (match[i] < 0) {
      double bw = 0.0;
      for (int e = ip[i]; e < ip[i + 1]; ++e) {
        const int j = ix[e] - shift;
        const double w = -av[e];
        if (j < 0 || j >= n) continue;
        if (j == i || !(w > bw) || match[j] >= 0) continue;
        bw = w;
        best = j;
      }
    }
    prop[i] = best;
  }
}

__global__ void __launch_bounds__(AMG_T) k_amg_accept(int n, const int* __restrict__ prop,
                                                      int* __restrict__ match) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int j = prop[i];
    if (j >= 0 && prop[j] == i) match[i] = j;     // each thread writes its own entry only
  }
}

// leader of every row's aggregate: min of a matched pair; a leftover joins its
// strongest matched neighbour; coupled only to leftovers -> singleton; no
// negative off-diagonal at all -> -1 (not represented on the coarse grid)
__global__ void __launch_bounds__(AMG_T) k_amg_leader(const int* __restrict__ ip,
                                                      const int* __restrict__ ix,
                                                      const double* __restrict__ av, int n,
                                                      int shift,
                                                      const int* __restrict__ match,
                                                      int* __restrict__ leader,
                                                      int* __restrict__ flag) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int L;
    const int m = match[i];
    if (m >= 0) {
      L = i < m ? i : m;
    } else {
      int best = -1, deg = 0;
      double bw = 0.0;
      for (int e = ip[i]; e < ip[i + 1]; ++e) {
        const int j = ix[e] - shift;
        const double w = -av[e];
        if (j == i || !(w > 0.0)) continue;
        ++deg;                    // couplings to a neighbour rank's rows count: such a row is a
        if (j < 0 || j >= n) continue;   // singleton aggregate, not a boundary row
        if (match[j] < 0 || !(w > bw)) continue;
        bw = w;
        best = j;
      }
      if (best >= 0) {
        const int mb = match[best];
        L = best < mb ? best : mb;
      } else {
        L = deg > 0 ? i : -1;
      }
    }
    leader[i] = L;
    flag[i] = (L == i) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(AMG_T) k_amg_assign(int n, const int* __restrict__ leader,
                                                      const int* __restrict__ ids,
                                                      int* __restrict__ agg) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int L = leader[i];
    agg[i] = L >= 0 ? ids[L] : -1;
  }
}

// total[i] <- pass[total[i]]
__global__ void __launch_bounds__(AMG_T) k_amg_compose(int n, int* __restrict__ total,
                                                       const int* __restrict__ pass) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int a = total[i];
    if (a >= 0) total[i] = pass[a];
  }
}

// ============================================================== member lists
__global__ void __launch_bounds__(AMG_T) k_amg_count_members(int n, const int* __restrict__ agg,
                                                             int* __restrict__ cnt) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int a = agg[i];
    if (a >= 0) atomicAdd(&cnt[a], 1);
  }
}

__global__ void __launch_bounds__(AMG_T) k_amg_fill_members(int n, const int* __restrict__ agg,
                                                            const int* __restrict__ cptr,
                                                            int* __restrict__ cur,
                                                            int* __restrict__ cmem) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int a = agg[i];
    if (a >= 0) {
      const int slot = atomicAdd(&cur[a], 1);
      B200_ASSERT(cptr[a] + slot < cptr[a + 1]);
      cmem[cptr[a] + slot] = i;
    }
  }
}

// slots were claimed in race order: sort every (short) list ascending
__global__ void __launch_bounds__(AMG_T) k_amg_sort_members(int nc, const int* __restrict__ cptr,
                                                            int* __restrict__ cmem) {
  const int stride = gridDim.x * blockDim.x;
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += stride) {
    const int lo = cptr[I], hi = cptr[I + 1];
    for (int a = lo + 1; a < hi; ++a) {
      const int v = cmem[a];
      int b = a - 1;
      while (b >= lo && cmem[b] > v) { cmem[b + 1] = cmem[b]; --b; }
      cmem[b + 1] = v;
    }
  }
}

// ================================================================== Galerkin
__global__ void __launch_bounds__(AMG_T) k_amg_row_ub(int nc, const int* __restrict__ cptr,
                                                      const int* __restrict__ cmem,
                                                      const int* __restrict__ ip,
                                                      int* __restrict__ ub) {
  const int stride = gridDim.x * blockDim.x;
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += stride) {
    int s = 0;
    for (int m = cptr[I]; m < cptr[I + 1]; ++m) {
      const int i = cmem[m];
      s += ip[i + 1] - ip[i];
    }
    ub[I] = s;
  }
}

// gather the mapped entries of the member rows (members ascending, entries in CSR
// order) and add the ones that land in the same coarse column, in that order (a
// fixed order: bit-repeatable).  `agg` is indexed by local column j in [jlo, jhi): the
// owned rows, plus -- in a slab run -- the halo windows of the neighbours' map.  The table of a row lives in thread-local memory
// (coalesced, L1-resident) while it has at most AMG_ROWCAP distinct columns and
// in the row's global segment otherwise.  Columns stay in first-arrival order:
// nothing downstream needs them sorted.
#define AMG_ROWCAP 48
__global__ void __launch_bounds__(128) k_amg_row_build(int nc, const int* __restrict__ cptr,
                                                       const int* __restrict__ cmem,
                                                       const int* __restrict__ ip,
                                                       const int* __restrict__ ix,
                                                       const double* __restrict__ av,
                                                       const double* __restrict__ diag,
                                                       const int* __restrict__ agg, int shift,
                                                       int jlo, int jhi, int coff,
                                                       const int* __restrict__ tofs,
                                                       int* __restrict__ tc,
                                                       double* __restrict__ tv,
                                                       int* __restrict__ len) {
  const int stride = gridDim.x * blockDim.x;
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += stride) {
    const int base = tofs[I];
    int lc[AMG_ROWCAP];
    double lv[AMG_ROWCAP];
    int m = 0;
    bool spilled = false;
    for (int q = cptr[I]; q < cptr[I + 1]; ++q) {
      const int i = cmem[q];
      for (int e = ip[i]; e < ip[i + 1]; ++e) {
        const int j = ix[e] - shift;
        if (j < jlo || j >= jhi) continue;
        int a = agg[j];           // slab runs: global coarse id (halo entries included)
        if (a < 0) continue;
        a -= coff;                // -> local signed id: < 0 rank-1's rows, >= nc rank+1's
        B200_ASSERT(m <= AMG_ROWCAP || spilled);
        const double v = (diag != nullptr && j == i) ? diag[i] : av[e];
        int k = 0;
        if (!spilled) {
          while (k < m && lc[k] != a) ++k;
          if (k < m) { lv[k] += v; continue; }
          if (m < AMG_ROWCAP) { lc[m] = a; lv[m] = v; ++m; continue; }
          for (int t = 0; t < m; ++t) { tc[base + t] = lc[t]; tv[base + t] = lv[t]; }
          spilled = true;
        }
        k = 0;
        while (k < m && tc[base + k] != a) ++k;
        if (k < m) { tv[base + k] += v; } else { tc[base + m] = a; tv[base + m] = v; ++m; }
      }
    }
    if (!spilled)
      for (int t = 0; t < m; ++t) { tc[base + t] = lc[t]; tv[base + t] = lv[t]; }
    len[I] = m;
  }
}

__global__ void __launch_bounds__(128) k_amg_row_emit(int nc, const int* __restrict__ tofs,
                                                      const int* __restrict__ tc,
                                                      const double* __restrict__ tv,
                                                      const int* __restrict__ cip,
                                                      int* __restrict__ cix,
                                                      double* __restrict__ cav,
                                                      double* __restrict__ dinv,
                                                      int* __restrict__ flags) {
  const int stride = gridDim.x * blockDim.x;
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += stride) {
    const int src = tofs[I], dst = cip[I], m = cip[I + 1] - cip[I];
    double d = 0.0;
    for (int r = 0; r < m; ++r) {
      const int c = tc[src + r];
      const double v = tv[src + r];
      cix[dst + r] = c;
      cav[dst + r] = v;
      if (c == I) d = v;
    }
    if (!(d > 0.0) || !isfinite(d)) flags[FLAG_NONPOS_DIAG] = 1;
    dinv[I] = d > 0.0 ? 1.0 / d : 1.0;
  }
}

// max_i sum_j |a_ij| / a_ii >= lambda_max(D^-1 A); positive doubles order like
// their bit patterns, so an integer atomicMax is exact and order-independent
__global__ void __launch_bounds__(AMG_T) k_amg_rowbound(const int* __restrict__ ip,
                                                        const int* __restrict__ ix,
                                                        const double* __restrict__ av,
                                                        const double* __restrict__ diag, int n,
                                                        int shift,
                                                        unsigned long long* __restrict__ out) {
  const int stride = gridDim.x * blockDim.x;
  double mx = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double s = 0.0, d = diag ? diag[i] : 0.0;
    for (int e = ip[i]; e < ip[i + 1]; ++e) {
      const double v = av[e];
      if (ix[e] - shift == i) {
        if (!diag) d = v;
        s += fabs(diag ? diag[i] : v);
      } else {
        s += fabs(v);
      }
    }
    if (d > 0.0 && isfinite(s)) mx = fmax(mx, s / d);
  }
  if (mx > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(mx));
}

// ============================================================ slab-run helpers
// owned part of the halo'd aggregation map, in global coarse ids
__global__ void __launch_bounds__(AMG_T) k_amg_global_ids(int n, const int* __restrict__ agg,
                                                          int coff, int* __restrict__ out) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int a = agg[i];
    out[i] = a >= 0 ? a + coff : -1;
  }
}

// mm[0] = min column, mm[1] = max column of a pass's output (integer atomics)
__global__ void __launch_bounds__(AMG_T) k_amg_col_range(int64_t nnz, const int* __restrict__ ix,
                                                         int* __restrict__ mm) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int lo = 0x7fffffff, hi = -0x7fffffff;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    const int c = ix[e];
    lo = min(lo, c);
    hi = max(hi, c);
  }
  if (lo <= hi) {
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
  }
}

__global__ void __launch_bounds__(AMG_T) k_amg_row_counts(int n, const int* __restrict__ ip,
                                                          int* __restrict__ cnt) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) cnt[i] = ip[i + 1] - ip[i];
}

// this rank's entries into their place in the gathered (replicated) level, columns global
__global__ void __launch_bounds__(AMG_T) k_amg_scatter_entries(int64_t nnz,
                                                               const int* __restrict__ ix,
                                                               const double* __restrict__ av,
                                                               int coff, int* __restrict__ oix,
                                                               double* __restrict__ oav) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    oix[e] = ix[e] + coff;
    oav[e] = av[e];
  }
}

// export helper: local signed columns -> global ids
__global__ void __launch_bounds__(AMG_T) k_amg_shift_ids(int64_t n, const int* __restrict__ in,
                                                         int add, int keep_negative,
                                                         int* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int v = in[i];
    out[i] = (keep_negative && v < 0) ? v : v + add;
  }
}

__global__ void k_amg_count_nonfinite(int64_t n, const double* __restrict__ a, int* cnt) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (!isfinite(a[i])) ++c;
  if (c) atomicAdd(cnt, c);
}

// ============================================================== cycle kernels
// Coarse-level product y = A x with the values held in fp32 (the coarse operators
// only precondition: 8 instead of 12 bytes per entry); x, y and the sums stay fp64.
// 8 lanes per row, warp-uniform trip count (see k_spmv_csr_vec).
__global__ void __launch_bounds__(256) k_amg_spmv_f32(const int* __restrict__ indptr,
                                                      const int* __restrict__ indices,
                                                      const float* __restrict__ data, int64_t n,
                                                      const double* __restrict__ x,
                                                      double* __restrict__ y) {
  const int lane = threadIdx.x % 8;
  const int64_t gstride = ((int64_t)gridDim.x * blockDim.x) / 8;
  const int64_t nround = (n + 3) / 4 * 4;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8; r < nround; r += gstride) {
    double s = 0.0;
    if (r < n) {
      const int e1 = indptr[r + 1];
      for (int e = indptr[r] + lane; e < e1; e += 8) s += (double)data[e] * x[indices[e]];
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, 8);
    if (lane == 0 && r < n) y[r] = s;
  }
}

__global__ void __launch_bounds__(AMG_T) k_amg_to_f32(int64_t n, const double* __restrict__ a,
                                                      float* __restrict__ o) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    o[i] = (float)a[i];
}

// dinv == nullptr: unit diagonal (level 0, Jacobi-scaled operator)
__global__ void __launch_bounds__(AMG_T) k_amg_cheb_first0(int64_t n,
                                                           const double* __restrict__ dinv,
                                                           const double* __restrict__ b, double c0,
                                                           double* __restrict__ d,
                                                           double* __restrict__ x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = c0 * (dinv ? dinv[i] : 1.0) * b[i];
    d[i] = v;
    x[i] = v;
  }
}

__global__ void __launch_bounds__(AMG_T) k_amg_cheb_first(int64_t n,
                                                          const double* __restrict__ dinv,
                                                          const double* __restrict__ b,
                                                          const double* __restrict__ ax, double c0,
                                                          double* __restrict__ d,
                                                          double* __restrict__ x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = c0 * (dinv ? dinv[i] : 1.0) * (b[i] - ax[i]);
    d[i] = v;
    x[i] += v;
  }
}

__global__ void __launch_bounds__(AMG_T) k_amg_cheb_next(int64_t n,
                                                         const double* __restrict__ dinv,
                                                         const double* __restrict__ b,
                                                         const double* __restrict__ ax, double c1,
                                                         double c2, double* __restrict__ d,
                                                         double* __restrict__ x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = c1 * d[i] + c2 * (dinv ? dinv[i] : 1.0) * (b[i] - ax[i]);
    d[i] = v;
    x[i] += v;
  }
}

// bc[I] = sum over members of w (b - ax); w == nullptr: 1; ax == nullptr: b is the residual
__global__ void __launch_bounds__(AMG_T) k_amg_restrict(int nc, const int* __restrict__ cptr,
                                                        const int* __restrict__ cmem,
                                                        const double* __restrict__ b,
                                                        const double* __restrict__ ax,
                                                        const double* __restrict__ w,
                                                        double* __restrict__ bc) {
  const int stride = gridDim.x * blockDim.x;
  for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += stride) {
    double s = 0.0;
    for (int m = cptr[I]; m < cptr[I + 1]; ++m) {
      const int i = cmem[m];
      B200_ASSERT(i >= 0);
      const double r = ax ? b[i] - ax[i] : b[i];
      s += w ? w[i] * r : r;
    }
    bc[I] = s;
  }
}

__global__ void __launch_bounds__(AMG_T) k_amg_prolong(int64_t n, const int* __restrict__ agg,
                                                       const double* __restrict__ ec,
                                                       const double* __restrict__ w,
                                                       double* __restrict__ x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int a = agg[i];
    B200_ASSERT(a >= -1);
    if (a >= 0) x[i] += w ? w[i] * ec[a] : ec[a];
  }
}

// K-cycle scalars: ks[0] = c1.v1 (rho1), ks[1] = c1.rc (alpha1), ks[2] = c2.v1 (gamma),
// ks[3] = c2.v2 (beta), ks[4] = c2.r1 (alpha2)
__global__ void __launch_bounds__(AMG_T) k_amg_dot2(int64_t n, const double* __restrict__ a,
                                                    const double* __restrict__ b,
                                                    const double* __restrict__ c, double* part,
                                                    unsigned* ticket, double* out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double ai = a[i];
    s0 += ai * b[i];
    s1 += ai * c[i];
  }
  double v[2] = {s0, s1};
  double* o[2] = {out, out + 1};
  grid_sum_finalize<2>(v, part, ticket, o);
}

__global__ void __launch_bounds__(AMG_T) k_amg_dot3(int64_t n, const double* __restrict__ a,
                                                    const double* __restrict__ b,
                                                    const double* __restrict__ c,
                                                    const double* __restrict__ e, double* part,
                                                    unsigned* ticket, double* out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double ai = a[i];
    s0 += ai * b[i];
    s1 += ai * c[i];
    s2 += ai * e[i];
  }
  double v[3] = {s0, s1, s2};
  double* o[3] = {out, out + 1, out + 2};
  grid_sum_finalize<3>(v, part, ticket, o);
}

// r1 = rc - (alpha1 / rho1) v1
__global__ void __launch_bounds__(AMG_T) k_amg_kres(int64_t n, const double* __restrict__ rc,
                                                    const double* __restrict__ v1,
                                                    const double* __restrict__ ks,
                                                    double* __restrict__ r1) {
  const double rho1 = ks[0];
  const double t = (rho1 > 0.0) ? ks[1] / rho1 : 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    r1[i] = rc[i] - t * v1[i];
}

// x (= c2) <- (alpha1/rho1 - gamma alpha2/(rho1 rho2)) c1 + (alpha2/rho2) c2
__global__ void __launch_bounds__(AMG_T) k_amg_kcomb(int64_t n, const double* __restrict__ c1,
                                                     const double* __restrict__ ks,
                                                     double* __restrict__ x) {
  const double rho1 = ks[0], a1 = ks[1], gam = ks[2], bet = ks[3], a2 = ks[4];
  double f1 = 0.0, f2 = 0.0;
  if (rho1 > 0.0) {
    const double rho2 = bet - gam * gam / rho1;
    f1 = a1 / rho1;
    if (rho2 > 0.0 && isfinite(rho2)) {
      f2 = a2 / rho2;
      f1 -= gam * a2 / (rho1 * rho2);
    }
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] = f1 * c1[i] + f2 * x[i];
}

// Jacobi-CG on the coarsest level inside one thread block (n <= a few thousand).
// Two block reductions per iteration, each one barrier: warp sums go to one of
// two alternating shared slots and every thread adds the 32 of them itself.
template <int NV>
__device__ __forceinline__ void cg_block_sums(double (&v)[NV], double (*slot)[32], int& flip) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double (*sl)[32] = slot + flip * NV;
  flip ^= 1;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) sl[k][w] = s;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s += sl[k][i];
    v[k] = s;
  }
}

__global__ void __launch_bounds__(1024) k_amg_coarse_cg(int n, const int* __restrict__ ip,
                                                        const int* __restrict__ ix,
                                                        const double* __restrict__ av,
                                                        const double* __restrict__ dinv,
                                                        const double* __restrict__ b,
                                                        double* __restrict__ x,
                                                        double* __restrict__ r,
                                                        double* __restrict__ p,
                                                        double* __restrict__ q, int maxit,
                                                        double rtol2) {
  __shared__ double slot[2 * 2][32];
  int flip = 0;
  double v2[2] = {0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double bi = b[i], z = dinv[i] * bi;
    x[i] = 0.0;
    r[i] = bi;
    p[i] = z;
    v2[0] += bi * z;
    v2[1] += bi * bi;
  }
  cg_block_sums<2>(v2, slot, flip);      // the barrier inside also publishes p
  double rz = v2[0];
  const double b2 = v2[1];
  if (!(b2 > 0.0) || !(rz > 0.0)) return;
  for (int it = 0; it < maxit; ++it) {
    double v1[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double s = 0.0;
      for (int e = ip[i]; e < ip[i + 1]; ++e) s += av[e] * p[ix[e]];
      q[i] = s;
      v1[0] += p[i] * s;
    }
    cg_block_sums<2>(v1, slot, flip);    // every thread is past its reads of p
    const double pq = v1[0];
    if (!(pq > 0.0)) return;
    const double al = rz / pq;
    double v[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      x[i] += al * p[i];
      const double ri = r[i] - al * q[i];
      r[i] = ri;
      v[0] += ri * ri;
      v[1] += ri * ri * dinv[i];
    }
    cg_block_sums<2>(v, slot, flip);
    const double rr = v[0], rzn = v[1];
    if (rr <= rtol2 * b2 || !(rzn > 0.0)) return;
    const double beta = rzn / rz;
    rz = rzn;
    for (int i = threadIdx.x; i < n; i += blockDim.x) p[i] = dinv[i] * r[i] + beta * p[i];
    __syncthreads();                     // p complete before the next product gathers it
  }
}

// ============================================================= outer FCG kernels
enum { F_ZQ = 0, F_DR = 1, F_RR = 2, F_TRUE = 3, F_DQ = 4 /* +1,+2 written by K_A */, F_COUNT = 8 };

__global__ void __launch_bounds__(AMG_T) k_amg_dot1(int64_t n, const double* __restrict__ a,
                                                    const double* __restrict__ b, double* part,
                                                    unsigned* ticket, double* out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    s += a[i] * b[i];
  double v[1] = {s};
  double* o[1] = {out};
  grid_sum_finalize<1>(v, part, ticket, o);
}

// d = z - (z.q_old / d_old.q_old) d   (first: d = z);  accumulates d.r
__global__ void __launch_bounds__(AMG_T) k_amg_fcg_dir(int64_t n, const double* __restrict__ z,
                                                       const double* __restrict__ r,
                                                       double* __restrict__ d, int first,
                                                       const double* __restrict__ fs,
                                                       double* part, unsigned* ticket,
                                                       double* out_dr) {
  double beta = 0.0;
  if (!first) {
    const double dq = fs[F_DQ];
    beta = (dq > 0.0) ? fs[F_ZQ] / dq : 0.0;
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double v = first ? z[i] : z[i] - beta * d[i];
    d[i] = v;
    s += v * r[i];
  }
  double v[1] = {s};
  double* o[1] = {out_dr};
  grid_sum_finalize<1>(v, part, ticket, o);
}

// x += alpha d ; r -= alpha q ; sums r.r and sum diag r^2 (unscaled residual norm)
__global__ void __launch_bounds__(AMG_T) k_amg_fcg_update(int64_t n, double* __restrict__ x,
                                                          double* __restrict__ r,
                                                          const double* __restrict__ d,
                                                          const double* __restrict__ q,
                                                          const double* __restrict__ diag,
                                                          const double* __restrict__ fs,
                                                          double* part, unsigned* ticket,
                                                          double* out) {
  const double dq = fs[F_DQ];
  const double alpha = (dq > 0.0) ? fs[F_DR] / dq : 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    x[i] += alpha * d[i];
    const double ri = r[i] - alpha * q[i];
    r[i] = ri;
    s0 += ri * ri;
    s1 += diag[i] * ri * ri;
  }
  double v[2] = {s0, s1};
  double* o[2] = {out + F_RR, out + F_TRUE};
  grid_sum_finalize<2>(v, part, ticket, o);
}

// w_i = sqrt(d_i): P~ = D^1/2 P maps coarse corrections into the scaled space
__global__ void __launch_bounds__(AMG_T) k_amg_winv(int64_t n, const double* __restrict__ s,
                                                    double* __restrict__ w) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    w[i] = 1.0 / s[i];
}

// ======================================================================= setup
static void cheb_coeffs(AmgLevel& L) {
  const double lmax = 1.05 * L.lmax, lmin = lmax / AMG_CHEB_RATIO;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin);
  const double sigma = theta / delta, rho0 = 1.0 / sigma, rho1 = 1.0 / (2.0 * sigma - rho0);
  L.c0 = 1.0 / theta;
  L.c1 = rho1 * rho0;
  L.c2 = 2.0 * rho1 / delta;
}

static int amg_members(b200_ctx* ctx, AmgState* A, int n, int nc, const int* agg, DevBuf& cptr,
                       DevBuf& cmem) {
  CK(A->t_cnt.ensure((size_t)(nc + 1) * sizeof(int)));
  CK(A->t_cur.ensure((size_t)(nc + 1) * sizeof(int)));
  int* cnt = A->t_cnt.as<int>();
  int* cur = A->t_cur.as<int>();
  CK(cudaMemsetAsync(cnt, 0, (size_t)(nc + 1) * sizeof(int), ctx->stream));
  CK(cudaMemsetAsync(cur, 0, (size_t)(nc + 1) * sizeof(int), ctx->stream));
  LAUNCH(ctx, k_amg_count_members, grid_for(n, AMG_T), AMG_T, 0, n, agg, cnt);
  CK(cptr.ensure((size_t)(nc + 1) * sizeof(int)));
  int64_t tot = 0;
  RET(b200_scan_exclusive_i32(ctx, cnt, cptr.as<int>(), nc, &tot));
  CK(cmem.ensure((size_t)(tot > 0 ? tot : 1) * sizeof(int)));
  LAUNCH(ctx, k_amg_fill_members, grid_for(n, AMG_T), AMG_T, 0, n, agg, cptr.as<int>(), cur,
         cmem.as<int>());
  LAUNCH(ctx, k_amg_sort_members, grid_for(nc, AMG_T), AMG_T, 0, nc, cptr.as<int>(),
         cmem.as<int>());
  CK(cudaGetLastError());
  return B200_OK;
}

// One pairwise pass on M -> agg (M.n local coarse ids, -1 = off the coarse grid) and the
// Galerkin matrix of the rows this rank owns.  Slab run (M.dist): the matching ignores
// the halo columns, the coarse numbering is local id + (coarse rows of the lower ranks),
// the neighbours' part of the map arrives by one integer halo exchange, the product is
// then row-local; the column range of the result gives the halo widths of the next
// matrix, which the neighbours learn through one small all-gather.  Every decision that
// changes the control flow is taken on all-gathered numbers.
static int amg_pair_pass(b200_ctx* ctx, AmgState* A, const PassMat& M, DevBuf& agg_out,
                         CsrOut& out) {
  const int n = M.n;
  DevBuf& match = A->t_match;
  DevBuf& prop = A->t_prop;
  DevBuf& ids = A->t_ids;
  CK(match.ensure((size_t)(n + 1) * sizeof(int)));
  CK(prop.ensure((size_t)(n + 1) * sizeof(int)));
  CK(ids.ensure((size_t)(n + 1) * sizeof(int)));
  CK(agg_out.ensure((size_t)(n + 1) * sizeof(int)));
  CK(cudaMemsetAsync(match.p, 0xff, (size_t)(n + 1) * sizeof(int), ctx->stream));
  const int g = grid_for(n, AMG_T);
  for (int r = 0; r < AMG_ROUNDS; ++r) {
    LAUNCH(ctx, k_amg_propose, g, AMG_T, 0, M.ip, M.ix, M.av, n, M.shift, match.as<int>(),
           prop.as<int>());
    LAUNCH(ctx, k_amg_accept, g, AMG_T, 0, n, prop.as<int>(), match.as<int>());
  }
  // prop <- leader, ids <- flags -> scanned ids
  LAUNCH(ctx, k_amg_leader, g, AMG_T, 0, M.ip, M.ix, M.av, n, M.shift, match.as<int>(),
         prop.as<int>(), ids.as<int>());
  int64_t nc64 = 0;
  RET(b200_scan_exclusive_i32(ctx, ids.as<int>(), ids.as<int>(), n, &nc64));
  const int nc = (int)nc64;
  LAUNCH(ctx, k_amg_assign, g, AMG_T, 0, n, prop.as<int>(), ids.as<int>(), agg_out.as<int>());
  int64_t coff = 0, ncg = nc;
  std::vector<int64_t> all((size_t)ctx->nranks * 4);
  if (M.dist) {
    const int64_t mine = nc;
    RET(b200_allgather_i64_host(ctx, &mine, 1, all.data()));
    ncg = 0;
    for (int r = 0; r < ctx->nranks; ++r) {
      if (r < ctx->rank) coff += all[r];
      ncg += all[r];
    }
    if (ncg >= (1ll << 31)) FAIL(B200_ERR_TOO_LARGE, "amg: coarse level exceeds int32 ids");
  }
  out.n = nc;
  out.ng = ncg;
  out.goff = coff;
  out.nnz = out.nnzg = 0;
  out.wlo = out.whi = out.slo = out.shi = 0;
  if (ncg == 0) return B200_OK;
  RET(amg_members(ctx, A, n, nc, agg_out.as<int>(), A->t_cptr, A->t_cmem));
  const int* cptr = A->t_cptr.as<int>();
  const int* cmem = A->t_cmem.as<int>();
  const int* aggmap = agg_out.as<int>();
  int jlo = 0, jhi = n;
  if (M.dist) {
    CK(A->t_aggg.ensure((size_t)(M.wlo + n + M.whi + 1) * sizeof(int)));
    int* go = A->t_aggg.as<int>() + M.wlo;
    LAUNCH(ctx, k_amg_global_ids, g, AMG_T, 0, n, agg_out.as<int>(), (int)coff, go);
    RET(b200_halo_exchange_w(ctx, go, n, 4, M.wlo, M.whi, M.slo, M.shi));
    aggmap = go;
    jlo = -M.wlo;
    jhi = n + M.whi;
  }
  // upper bounds -> segment offsets (match is reused as ub / len, ids as tofs)
  const int gc = grid_for(nc, 128);
  LAUNCH(ctx, k_amg_row_ub, grid_for(nc, AMG_T), AMG_T, 0, nc, cptr, cmem, M.ip, match.as<int>());
  int64_t T = 0;
  RET(b200_scan_exclusive_i32(ctx, match.as<int>(), ids.as<int>(), nc, &T));
  CK(A->t_tc.ensure((size_t)(T > 0 ? T : 1) * sizeof(int)));
  CK(A->t_tv.ensure((size_t)(T > 0 ? T : 1) * sizeof(double)));
  // (a register-resident table for the short rows of the first passes was tried and was
  // slower: 119 vs 98 ms of setup at 400^3 -- the gathers, not the table, bound this kernel)
  LAUNCH(ctx, k_amg_row_build, gc, 128, 0, nc, cptr, cmem, M.ip, M.ix, M.av, M.dg, aggmap,
         M.shift, jlo, jhi, (int)coff, ids.as<int>(), A->t_tc.as<int>(), A->t_tv.as<double>(),
         match.as<int>());
  CK(out.indptr.ensure((size_t)(nc + 1) * sizeof(int)));
  int64_t cnnz = 0;
  RET(b200_scan_exclusive_i32(ctx, match.as<int>(), out.indptr.as<int>(), nc, &cnnz));
  CK(out.indices.ensure((size_t)(cnnz > 0 ? cnnz : 1) * sizeof(int)));
  CK(out.data.ensure((size_t)(cnnz > 0 ? cnnz : 1) * sizeof(double)));
  CK(out.dinv.ensure((size_t)(nc > 0 ? nc : 1) * sizeof(double)));
  LAUNCH(ctx, k_amg_row_emit, gc, 128, 0, nc, ids.as<int>(), A->t_tc.as<int>(),
         A->t_tv.as<double>(), out.indptr.as<int>(), out.indices.as<int>(), out.data.as<double>(),
         out.dinv.as<double>(), ctx->w_flags.as<int>());
  CK(cudaGetLastError());
  out.nnz = out.nnzg = cnnz;
  if (M.dist) {
    CK(A->t_minmax.ensure(2 * sizeof(int)));
    int mm[2] = {0x7fffffff, -0x7fffffff};
    CK(cudaMemcpyAsync(A->t_minmax.p, mm, sizeof mm, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_amg_col_range, grid_for(cnnz, AMG_T), AMG_T, 0, cnnz, out.indices.as<int>(),
           A->t_minmax.as<int>());
    CK(cudaMemcpyAsync(mm, A->t_minmax.p, sizeof mm, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int64_t mine[3] = {0, 0, cnnz};
    if (mm[0] <= mm[1]) {
      mine[0] = mm[0] < 0 ? -(int64_t)mm[0] : 0;
      mine[1] = mm[1] >= nc ? (int64_t)mm[1] - nc + 1 : 0;
    }
    RET(b200_allgather_i64_host(ctx, mine, 3, all.data()));
    out.wlo = (int)mine[0];
    out.whi = (int)mine[1];
    out.slo = ctx->rank > 0 ? (int)all[(size_t)(ctx->rank - 1) * 3 + 1] : 0;
    out.shi = ctx->rank < ctx->nranks - 1 ? (int)all[(size_t)(ctx->rank + 1) * 3 + 0] : 0;
    out.nnzg = 0;
    for (int r = 0; r < ctx->nranks; ++r) out.nnzg += all[(size_t)r * 3 + 2];
    // aggregates never cross a rank boundary and the fine couplings join neighbouring
    // ranks only, so the windows lie inside the neighbours' rows
    const int64_t nlo = ctx->rank > 0 ? 1 : 0, nhi = ctx->rank < ctx->nranks - 1 ? 1 : 0;
    if ((out.wlo > 0 && !nlo) || (out.whi > 0 && !nhi) || out.slo > nc || out.shi > nc)
      FAIL(B200_ERR_ARG, "amg: a coarse coupling reaches past the neighbouring rank");
  }
  return B200_OK;
}

static int amg_level_bound(b200_ctx* ctx, AmgState* A, AmgLevel& L, const double* diag,
                           int shift) {
  CK(A->lmax_dev.ensure(sizeof(unsigned long long)));
  CK(cudaMemsetAsync(A->lmax_dev.p, 0, sizeof(unsigned long long), ctx->stream));
  LAUNCH(ctx, k_amg_rowbound, grid_for(L.n, AMG_T), AMG_T, 0, L.ip, L.ix, L.av, diag, (int)L.n,
         shift, A->lmax_dev.as<unsigned long long>());
  double lm = 0.0;
  CK(cudaMemcpyAsync(&lm, A->lmax_dev.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (L.dist) {
    // positive doubles order like their bit patterns: the all-gathered maximum is exact
    int64_t bits = 0;
    if (lm > 0.0 && isfinite(lm)) memcpy(&bits, &lm, sizeof bits);
    std::vector<int64_t> all((size_t)ctx->nranks);
    RET(b200_allgather_i64_host(ctx, &bits, 1, all.data()));
    for (int r = 0; r < ctx->nranks; ++r) bits = all[r] > bits ? all[r] : bits;
    memcpy(&lm, &bits, sizeof lm);
  }
  L.lmax = (lm > 1.0 && isfinite(lm)) ? lm : 2.0;
  cheb_coeffs(L);
  return B200_OK;
}

// Row-partitioned output of the last pass -> the same matrix held in full by every
// rank (columns global): every rank writes its rows into zeroed global arrays and the
// sum over the ranks is the concatenation.
static int amg_gather_level(b200_ctx* ctx, AmgState* A, const CsrOut& loc, CsrOut& full) {
  const int64_t ncg = loc.ng, nnzg = loc.nnzg;
  std::vector<int64_t> all((size_t)ctx->nranks);
  const int64_t mine = loc.nnz;
  RET(b200_allgather_i64_host(ctx, &mine, 1, all.data()));
  int64_t eoff = 0;
  for (int r = 0; r < ctx->rank; ++r) eoff += all[r];
  CK(A->t_gcnt.ensure((size_t)(ncg + 1) * sizeof(int)));
  CK(cudaMemsetAsync(A->t_gcnt.p, 0, (size_t)(ncg + 1) * sizeof(int), ctx->stream));
  LAUNCH(ctx, k_amg_row_counts, grid_for(loc.n, AMG_T), AMG_T, 0, (int)loc.n,
         loc.indptr.as<int>(), A->t_gcnt.as<int>() + loc.goff);
  RET(b200_allreduce_sum_i32(ctx, A->t_gcnt.as<int>(), (size_t)ncg));
  CK(full.indptr.ensure((size_t)(ncg + 1) * sizeof(int)));
  int64_t tot = 0;
  RET(b200_scan_exclusive_i32(ctx, A->t_gcnt.as<int>(), full.indptr.as<int>(), ncg, &tot));
  if (tot != nnzg) FAIL(B200_ERR_ARG, "amg: gathered level is inconsistent (%lld vs %lld entries)",
                        (long long)tot, (long long)nnzg);
  const size_t ne = (size_t)(nnzg > 0 ? nnzg : 1);
  CK(full.indices.ensure(ne * sizeof(int)));
  CK(full.data.ensure(ne * sizeof(double)));
  CK(full.dinv.ensure((size_t)ncg * sizeof(double)));
  CK(cudaMemsetAsync(full.indices.p, 0, ne * sizeof(int), ctx->stream));
  CK(cudaMemsetAsync(full.data.p, 0, ne * sizeof(double), ctx->stream));
  CK(cudaMemsetAsync(full.dinv.p, 0, (size_t)ncg * sizeof(double), ctx->stream));
  LAUNCH(ctx, k_amg_scatter_entries, grid_for(loc.nnz, AMG_T), AMG_T, 0, loc.nnz,
         loc.indices.as<int>(), loc.data.as<double>(), (int)loc.goff,
         full.indices.as<int>() + eoff, full.data.as<double>() + eoff);
  if (loc.n > 0)
    CK(cudaMemcpyAsync(full.dinv.as<double>() + loc.goff, loc.dinv.p,
                       (size_t)loc.n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  RET(b200_allreduce_sum_i32(ctx, full.indices.as<int>(), (size_t)nnzg));
  RET(b200_allreduce_sum_n(ctx, full.data.as<double>(), (size_t)nnzg));
  RET(b200_allreduce_sum_n(ctx, full.dinv.as<double>(), (size_t)ncg));
  CK(cudaGetLastError());
  full.n = full.ng = ncg;
  full.nnz = full.nnzg = nnzg;
  full.goff = 0;
  full.wlo = full.whi = full.slo = full.shi = 0;
  return B200_OK;
}

static int amg_setup(b200_ctx* ctx) {
  AmgState* A = amg_of(ctx);
  A->valid = false;
  A->nlev = 0;
  RET(b200_zero_flags(ctx));
  const bool multi = ctx->nranks > 1;
  const int64_t repl_rows = amg_repl_rows();
  const double* diag0 = b200_cur_diag(ctx);
  if (A->lv.empty()) A->lv.emplace_back();
  {
    AmgLevel& L = A->lv[0];
    L.n = ctx->n;
    L.nnz = ctx->sys_nnz;
    L.ip = ctx->sys_indptr.as<int>();
    L.ix = ctx->sys_indices.as<int>();
    L.av = ctx->sys_data.as<double>();
    L.dist = multi;
    L.ng = multi ? ctx->n_global : ctx->n;
    L.goff = ctx->row_begin;
    L.wlo = L.slo = (multi && ctx->rank > 0) ? (int)ctx->halo : 0;
    L.whi = L.shi = (multi && ctx->rank < ctx->nranks - 1) ? (int)ctx->halo : 0;
    L.next_repl = false;
    L.coff = 0;
    RET(amg_level_bound(ctx, A, L, diag0, (int)ctx->row_begin));
  }
  size_t l = 0;
  double t_prev = 0.0;
  auto now_ms = [&]() {
    cudaStreamSynchronize(ctx->stream);
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
  };
  if (ctx->debug) t_prev = now_ms();
  bool broken = false;       // slab run whose row-partitioned levels cannot reach the coarsest size
  while (l + 1 < AMG_MAX_LEVELS && A->lv[l].ng > AMG_COARSEST) {
    if (A->lv.size() <= l + 1) A->lv.emplace_back();     // may move the vector: index, never keep refs
    const int n = (int)A->lv[l].n;
    const int64_t ng = A->lv[l].ng;
    PassMat M;
    M.n = n;
    M.ip = A->lv[l].ip;
    M.ix = A->lv[l].ix;
    M.av = A->lv[l].av;
    M.dg = (l == 0) ? diag0 : nullptr;
    M.shift = (l == 0) ? (int)ctx->row_begin : 0;
    M.dist = A->lv[l].dist;
    M.ng = ng;
    M.wlo = A->lv[l].wlo; M.whi = A->lv[l].whi; M.slo = A->lv[l].slo; M.shi = A->lv[l].shi;
    bool ok = true;
    CsrOut* last = nullptr;
    const int npass = (l == 0) ? amg_passes0() : AMG_PASSES;
    for (int pass = 0; pass < npass; ++pass) {
      // the last pass writes the next level's own buffers: nothing is reallocated
      // when a hierarchy of the same shape is rebuilt
      CsrOut& dst = (pass == npass - 1) ? A->lv[l + 1].own : A->t_csr[pass & 1];
      DevBuf& aggp = (pass == 0) ? A->lv[l].agg : A->t_aggpass;
      RET(amg_pair_pass(ctx, A, M, aggp, dst));
      if (dst.ng == 0) { ok = false; break; }
      if (pass > 0)
        LAUNCH(ctx, k_amg_compose, grid_for(n, AMG_T), AMG_T, 0, n, A->lv[l].agg.as<int>(),
               aggp.as<int>());
      if (ctx->debug) {
        const double t = now_ms();
        fprintf(stderr, "[b200 amg r%d] level %d pass %d: %lld -> %lld rows (global %lld -> %lld), "
                "%lld nnz, halo %d/%d, %.2f ms\n", ctx->rank, (int)l, pass, (long long)M.n,
                (long long)dst.n, (long long)M.ng, (long long)dst.ng, (long long)dst.nnz, dst.wlo,
                dst.whi, t - t_prev);
        t_prev = t;
      }
      M.n = (int)dst.n;
      M.ip = dst.indptr.as<int>();
      M.ix = dst.indices.as<int>();
      M.av = dst.data.as<double>();
      M.dg = nullptr;
      M.shift = 0;
      M.ng = dst.ng;
      M.wlo = dst.wlo; M.whi = dst.whi; M.slo = dst.slo; M.shi = dst.shi;
      last = &dst;
    }
    if (!ok || last == nullptr || (double)last->ng > 0.7 * (double)ng) {
      // coarsening stalled: level l is the last one (a row-partitioned level cannot be:
      // the coarsest problem is solved inside one thread block)
      if (A->lv[l].dist) broken = true;
      break;
    }
    AmgLevel& F = A->lv[l];
    AmgLevel& C = A->lv[l + 1];
    F.nc = last->n;
    F.coff = last->goff;
    F.next_repl = false;
    RET(amg_members(ctx, A, n, (int)last->n, F.agg.as<int>(), F.cptr, F.cmem));
    const CsrOut* live = last;
    if (F.dist && last->ng > repl_rows && last->ng > AMG_COARSEST) {
      // the next level stays row-partitioned
      C.dist = true;
      C.n = last->n; C.nnz = last->nnz; C.ng = last->ng; C.nnzg = last->nnzg;
      C.goff = last->goff;
      C.wlo = last->wlo; C.whi = last->whi; C.slo = last->slo; C.shi = last->shi;
    } else if (F.dist) {
      // small enough: every rank holds the next level (and everything below it) in full
      RET(amg_gather_level(ctx, A, *last, C.gath));
      live = &C.gath;
      F.next_repl = true;
      C.dist = false;
      C.n = C.ng = C.gath.n; C.nnz = C.nnzg = C.gath.nnz;
      C.goff = 0;
      C.wlo = C.whi = C.slo = C.shi = 0;
    } else {
      C.dist = false;
      C.n = C.ng = last->n; C.nnz = C.nnzg = last->nnz;
      C.goff = 0;
      C.wlo = C.whi = C.slo = C.shi = 0;
    }
    C.next_repl = false;
    C.coff = 0;
    C.ip = live->indptr.as<int>();
    C.ix = live->indices.as<int>();
    C.av = live->data.as<double>();
    C.dv = live->dinv.as<double>();
    RET(amg_level_bound(ctx, A, C, nullptr, 0));
    if (amg_f32()) {
      CK(C.dataf.ensure((size_t)(C.nnz > 0 ? C.nnz : 1) * sizeof(float)));
      LAUNCH(ctx, k_amg_to_f32, grid_for(C.nnz, AMG_T), AMG_T, 0, C.nnz, C.av, C.dataf.as<float>());
    }
    const size_t vlen = (size_t)(C.n > 0 ? C.n : 1), plen = vlen + C.wlo + C.whi;
    CK(C.x.ensure(plen * sizeof(double)));
    for (DevBuf* b : {&C.b, &C.ax, &C.dir}) CK(b->ensure(vlen * sizeof(double)));
    if ((int)(l + 1) <= AMG_KLEVELS_MAX) {
      CK(C.kc1.ensure(plen * sizeof(double)));
      for (DevBuf* b : {&C.kb, &C.kv1, &C.kv2}) CK(b->ensure(vlen * sizeof(double)));
      CK(C.ks.ensure(8 * sizeof(double)));
    }
    ++l;
  }
  A->nlev = l + 1;
  if (A->lv[l].dist && A->nlev > 1) broken = true;       // e.g. the level cap was reached
  if (A->nlev > 1 && !A->lv[l].dist)
    CK(A->lv[l].cgw.ensure((size_t)3 * A->lv[l].n * sizeof(double)));
  CK(cudaMemcpyAsync(ctx->flags_h, ctx->w_flags.p, FLAG_COUNT * sizeof(int),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  int64_t bad = ctx->flags_h[FLAG_NONPOS_DIAG] ? 1 : 0;
  RET(b200_allreduce_max_host(ctx, &bad));               // same verdict on every rank
  if (bad) FAIL(B200_ERR_SINGULAR, "amg: a coarse matrix has a non-positive diagonal");
  A->valid = A->nlev > 1 && !broken;
  A->epoch = ctx->sys_epoch;
  A->lin_epoch = ctx->lin_epoch;
  if (ctx->debug) {
    // how many non-finite numbers each level's matrix and inverse diagonal hold
    for (size_t k = 1; k < A->nlev; ++k) {
      int* cnt = ctx->w_flags.as<int>() + FLAG_COUNT - 2;
      cudaMemsetAsync(cnt, 0, sizeof(int), ctx->stream);
      k_amg_count_nonfinite<<<grid_for(A->lv[k].nnz, AMG_T), AMG_T, 0, ctx->stream>>>(
          A->lv[k].nnz, A->lv[k].av, cnt);
      k_amg_count_nonfinite<<<grid_for(A->lv[k].n, AMG_T), AMG_T, 0, ctx->stream>>>(
          A->lv[k].n, A->lv[k].dv, cnt);
      int h = 0;
      cudaMemcpyAsync(&h, cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      fprintf(stderr, "[b200 amg r%d] level %d: %d non-finite values, halo %d/%d send %d/%d\n",
              ctx->rank, (int)k, h, A->lv[k].wlo, A->lv[k].whi, A->lv[k].slo, A->lv[k].shi);
    }
    fprintf(stderr, "[b200 amg r%d] levels:", ctx->rank);
    for (size_t k = 0; k < A->nlev; ++k)
      fprintf(stderr, " n=%lld/%lld nnz=%lld %s lmax=%.3f |", (long long)A->lv[k].n,
              (long long)A->lv[k].ng, (long long)A->lv[k].nnz,
              A->lv[k].dist ? "rows" : "full", A->lv[k].lmax);
    fprintf(stderr, "%s\n", broken ? " (unusable: coarsening stalled on a row-partitioned level)" : "");
  }
  return B200_OK;
}

// ======================================================================= cycle
// halo windows of a level's product operand (x: first owned entry)
static int amg_halo(b200_ctx* ctx, AmgState* A, size_t l, double* x) {
  AmgLevel& L = A->lv[l];
  if (!L.dist) return B200_OK;
  if (l == 0) return b200_halo_exchange(ctx, x - ctx->pad);
  return b200_halo_exchange_w(ctx, x, L.n, 8, L.wlo, L.whi, L.slo, L.shi);
}

// y = A_l x; on a row-partitioned level x must be a padded vector (level 0: the PCG
// layout; below: [wlo | n | whi]) and its halo windows are filled first
static int amg_spmv(b200_ctx* ctx, AmgState* A, size_t l, double* x, double* y) {
  RET(amg_halo(ctx, A, l, x));
  if (l == 0) return b200_spmv_dots(ctx, x, x, y, ctx->scal.as<double>() + SC_AUX0);
  AmgLevel& L = A->lv[l];
  if (amg_f32()) {
    LAUNCH(ctx, k_amg_spmv_f32, grid_for(L.n * 8, 256), 256, 0, L.ip, L.ix, L.dataf.as<float>(),
           L.n, x, y);
    return B200_OK;
  }
  return b200_spmv_csr_raw(ctx, L.ip, L.ix, L.av, L.n, 0, x, y);
}

// solves A_l x ~ b (x from zero); level 0: b, x, ax, dir are owned pointers of the
// padded layout, everything in the Jacobi-scaled space
static int amg_child(b200_ctx* ctx, AmgState* A, size_t l);

// restriction of level l's residual into the right-hand side of level l+1 / the
// correction back; when the next level is replicated the ranks' parts are gathered
static int amg_restrict(b200_ctx* ctx, AmgState* A, size_t l, const double* b, const double* ax,
                        const double* w) {
  AmgLevel& L = A->lv[l];
  AmgLevel& C = A->lv[l + 1];
  double* cb = C.b.as<double>();
  if (L.next_repl) {
    CK(cudaMemsetAsync(cb, 0, (size_t)C.n * sizeof(double), ctx->stream));
    cb += L.coff;
  }
  LAUNCH(ctx, k_amg_restrict, grid_for(L.nc, AMG_T), AMG_T, 0, (int)L.nc, L.cptr.as<int>(),
         L.cmem.as<int>(), b, ax, w, cb);
  if (L.next_repl) RET(b200_allreduce_sum_n(ctx, C.b.as<double>(), (size_t)C.n));
  return B200_OK;
}

static int amg_prolong(b200_ctx* ctx, AmgState* A, size_t l, const double* w, double* x) {
  AmgLevel& L = A->lv[l];
  AmgLevel& C = A->lv[l + 1];
  const double* ec = C.xo() + (L.next_repl ? L.coff : 0);
  LAUNCH(ctx, k_amg_prolong, grid_for(L.n, AMG_T), AMG_T, 0, L.n, L.agg.as<int>(), ec, w, x);
  return B200_OK;
}

// level 0 on the DIA operator: every product is fused with the update that
// consumes it (k_dia_smooth); zq_aux/zq_out: optional sum zq_aux . x of the result
static int amg_cycle_fine_dia(b200_ctx* ctx, AmgState* A, double* b, double* x, double* xalt,
                              double* res, double* dir, const double* zq_aux, double* zq_out) {
  AmgLevel& L = A->lv[0];
  double* scratch = ctx->scal.as<double>() + SC_AUX0;
  const double* w = A->winv.as<double>();
  RET(amg_halo(ctx, A, 0, b));
  RET(b200_dia_smooth(ctx, 1, b, nullptr, nullptr, dir, x, L.c0, L.c1, L.c2, scratch));
  RET(amg_halo(ctx, A, 0, x));
  RET(b200_dia_smooth(ctx, 4, x, b, w, nullptr, res, 0.0, 0.0, 0.0, scratch));
  RET(amg_restrict(ctx, A, 0, res, nullptr, nullptr));
  RET(amg_child(ctx, A, 0));
  RET(amg_prolong(ctx, A, 0, w, x));
  RET(amg_halo(ctx, A, 0, x));
  RET(b200_dia_smooth(ctx, 2, x, b, nullptr, dir, xalt, L.c0, L.c1, L.c2, scratch));
  RET(amg_halo(ctx, A, 0, xalt));
  RET(b200_dia_smooth(ctx, 3, xalt, b, zq_aux, dir, x, L.c0, L.c1, L.c2,
                      zq_aux ? zq_out : scratch));
  return B200_OK;
}

static int amg_cycle(b200_ctx* ctx, AmgState* A, size_t l, const double* b, double* x, double* ax,
                     double* dir) {
  AmgLevel& L = A->lv[l];
  const int64_t n = L.n;
  if (l + 1 == A->nlev) {
    double* w = L.cgw.as<double>();
    LAUNCH(ctx, k_amg_coarse_cg, 1, 1024, 0, (int)n, L.ip, L.ix, L.av, L.dv, b, x,
           w, w + n, w + 2 * n, AMG_COARSE_ITERS, 1e-2);
    return B200_OK;
  }
  const double* dinv = (l == 0) ? nullptr : L.dv;
  const double* wgt = (l == 0) ? A->winv.as<double>() : nullptr;
  const int g = grid_for(n, AMG_T);
  // pre-smoothing, Chebyshev(2) from zero
  LAUNCH(ctx, k_amg_cheb_first0, g, AMG_T, 0, n, dinv, b, L.c0, dir, x);
  RET(amg_spmv(ctx, A, l, x, ax));
  LAUNCH(ctx, k_amg_cheb_next, g, AMG_T, 0, n, dinv, b, ax, L.c1, L.c2, dir, x);
  // coarse-grid correction
  RET(amg_spmv(ctx, A, l, x, ax));
  RET(amg_restrict(ctx, A, l, b, ax, wgt));
  RET(amg_child(ctx, A, l));
  RET(amg_prolong(ctx, A, l, wgt, x));
  // post-smoothing, Chebyshev(2)
  RET(amg_spmv(ctx, A, l, x, ax));
  LAUNCH(ctx, k_amg_cheb_first, g, AMG_T, 0, n, dinv, b, ax, L.c0, dir, x);
  RET(amg_spmv(ctx, A, l, x, ax));
  LAUNCH(ctx, k_amg_cheb_next, g, AMG_T, 0, n, dinv, b, ax, L.c1, L.c2, dir, x);
  CK(cudaGetLastError());
  return B200_OK;
}

// coarse-grid correction below level l: C.b holds the restricted residual, the
// correction comes back in C.x (K-cycle on the finest coarse levels, V below)
static int amg_child(b200_ctx* ctx, AmgState* A, size_t l) {
  AmgLevel& C = A->lv[l + 1];
  const int64_t nc = C.n;
  const int gc = grid_for(nc, AMG_T);
  double* part = ctx->w_part.as<double>();
  unsigned* ticket = ctx->w_ticket.as<unsigned>();
  const bool kcyc = (int)l < AMG_KLEVELS && l + 2 < A->nlev;
  if (kcyc) {
    double* ks = C.ks.as<double>();
    CK(cudaMemcpyAsync(C.kb.p, C.b.p, (size_t)nc * sizeof(double), cudaMemcpyDeviceToDevice,
                       ctx->stream));
    RET(amg_cycle(ctx, A, l + 1, C.b.as<double>(), C.xo(), C.ax.as<double>(), C.dir.as<double>()));
    CK(cudaMemcpyAsync(C.kc1o(), C.xo(), (size_t)nc * sizeof(double), cudaMemcpyDeviceToDevice,
                       ctx->stream));
    RET(amg_spmv(ctx, A, l + 1, C.kc1o(), C.kv1.as<double>()));
    LAUNCH(ctx, k_amg_dot2, gc, AMG_T, 0, nc, C.kc1o(), C.kv1.as<double>(), C.kb.as<double>(),
           part, ticket, ks);
    if (C.dist) RET(b200_allreduce_sum(ctx, ks, 2));
    LAUNCH(ctx, k_amg_kres, gc, AMG_T, 0, nc, C.kb.as<double>(), C.kv1.as<double>(), ks,
           C.b.as<double>());
    RET(amg_cycle(ctx, A, l + 1, C.b.as<double>(), C.xo(), C.ax.as<double>(), C.dir.as<double>()));
    RET(amg_spmv(ctx, A, l + 1, C.xo(), C.kv2.as<double>()));
    LAUNCH(ctx, k_amg_dot3, gc, AMG_T, 0, nc, C.xo(), C.kv1.as<double>(), C.kv2.as<double>(),
           C.b.as<double>(), part, ticket, ks + 2);
    if (C.dist) RET(b200_allreduce_sum(ctx, ks + 2, 3));
    LAUNCH(ctx, k_amg_kcomb, gc, AMG_T, 0, nc, C.kc1o(), ks, C.xo());
  } else {
    RET(amg_cycle(ctx, A, l + 1, C.b.as<double>(), C.xo(), C.ax.as<double>(), C.dir.as<double>()));
  }
  return B200_OK;
}

// ================================================================ outer solver
// (both settings are the same on every rank of a slab run; everything that can
// differ between ranks is agreed on collectively inside b200_amg_pcg / amg_setup)
bool b200_amg_applicable(b200_ctx* ctx) {
  if (ctx->precond != B200_PRECOND_AMG || ctx->nonsym) return false;
  if (ctx->nranks == 1) return true;
  const char* e = getenv("B200_AMG_SLABS");      // =0: slab runs stay with Jacobi-PCG
  return !(e && e[0] == '0');
}

// returns B200_OK with *used = 0 when the hierarchy could not be built (coarsening
// stalled at once, non-unit diagonal): the caller then runs Jacobi-PCG
int b200_amg_pcg(b200_ctx* ctx, const double* x0, double rtol, double atol, int64_t maxiter,
                 double* x, int64_t* iters, double* relres, int* info, int* used) {
  *used = 0;
  // Slab mode (one rank per GPU): the outer flexible CG runs on the global operator
  // (halo exchange + all-reduced dot products, as in the PCG) and the preconditioner is
  // one multigrid cycle of the row-partitioned hierarchy (see the header).
  const bool multi = ctx->nranks > 1;
  CK(cudaEventRecord(ctx->ev0, ctx->stream));
  RET(b200_pcg_prepare(ctx, ctx->force_fmt));
  // every rank must take the same path: agree on the local verdicts
  const int64_t n_all = multi ? ctx->n_global : ctx->n;
  int64_t bad = (!ctx->unit_diag || n_all <= AMG_COARSEST || n_all >= (1ll << 31)) ? 1 : 0;
  RET(b200_allreduce_max_host(ctx, &bad));
  if (bad) return B200_OK;
  AmgState* A = amg_of(ctx);
  // a re-linearised source term moves the diagonal: coarse operators built from the
  // old one precondition badly (config 5: 250 instead of 31 iterations), so rebuild
  if (A->epoch != ctx->sys_epoch || A->lin_epoch != ctx->lin_epoch || A->nlev == 0) {
    const int rc = amg_setup(ctx);
    if (rc != B200_OK) {
      if (!multi) return rc;
      A->valid = false;          // reported below as "no hierarchy": all ranks use Jacobi
      A->nlev = 0;
    }
  }
  bad = A->valid ? 0 : 1;
  RET(b200_allreduce_max_host(ctx, &bad));
  if (bad) return B200_OK;
  *used = 1;
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  const int64_t n = ctx->n, pad = ctx->pad;
  if (maxiter <= 0) maxiter = 10 * ctx->n_global;
  const size_t vlen = (size_t)(n + 2 * pad);
  for (DevBuf* b : {&A->z0, &A->z1, &A->ax0, &A->dir0, &A->fd, &A->fq}) {
    CK(b->ensure(vlen * sizeof(double)));
    CK(cudaMemsetAsync(b->p, 0, vlen * sizeof(double), ctx->stream));
  }
  CK(A->winv.ensure((size_t)n * sizeof(double)));
  double* sc = ctx->scal.as<double>();
  double* part = ctx->w_part.as<double>();
  unsigned* ticket = ctx->w_ticket.as<unsigned>();
  int* flags = ctx->w_flags.as<int>();
  double* xt = ctx->vx.as<double>() + pad;
  double* r = ctx->vr.as<double>() + pad;
  double* tmp = ctx->vp.as<double>() + pad;
  double* v = ctx->vap.as<double>() + pad;
  double* bt = ctx->vb.as<double>() + pad;
  double* z = A->z0.as<double>() + pad;
  double* z1 = A->z1.as<double>() + pad;
  const bool fused = b200_dia_smooth_available(ctx);
  double* ax0 = A->ax0.as<double>() + pad;
  double* dir0 = A->dir0.as<double>() + pad;
  double* d = A->fd.as<double>() + pad;
  double* q = A->fq.as<double>() + pad;
  const double* diag = b200_cur_diag(ctx);
  double* fs = sc;                 // F_* slots at the start of ctx->scal
  const int g = grid_for(n, AMG_T);
  LAUNCH(ctx, k_amg_winv, g, AMG_T, 0, n, ctx->sc.as<double>() + pad, A->winv.as<double>());
  CK(cudaMemsetAsync(flags, 0, FLAG_COUNT * sizeof(int), ctx->stream));
  CK(cudaMemsetAsync(sc, 0, SC_COUNT * sizeof(double), ctx->stream));
  RET(b200_launch_init(ctx, x0, xt, bt, sc + SC_BNORM));
  if (multi) {
    RET(b200_allreduce_sum(ctx, sc + SC_BNORM, 1));
    RET(b200_halo_exchange(ctx, ctx->vx.as<double>()));
  }
  RET(b200_spmv_dots(ctx, xt, bt, v, sc + SC_AUX0));
  RET(b200_launch_resid(ctx, bt, v, diag, r, tmp, sc + F_RR, sc + F_TRUE));
  if (multi) RET(b200_allreduce_sum(ctx, sc + F_RR, 2));
  CK(cudaMemcpyAsync(ctx->scal_h, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaMemcpyAsync(ctx->flags_h, flags, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost,
                     ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // in slab mode judge by the all-reduced norms only (same verdict on every rank)
  if ((!multi && ctx->flags_h[FLAG_NONFINITE]) || !isfinite(ctx->scal_h[SC_BNORM]) ||
      !isfinite(ctx->scal_h[F_TRUE]))
    FAIL(B200_ERR_NONFINITE, "b or x0 contains inf/nan values");
  const double bnorm2 = ctx->scal_h[SC_BNORM];
  double tol2 = rtol * rtol * bnorm2;
  if (atol * atol > tol2) tol2 = atol * atol;
  double true2 = ctx->scal_h[F_TRUE];
  int64_t it = 0;
  int inf = 0, restarts = 0;
  if (bnorm2 == 0.0) {
    CK(cudaMemsetAsync(x, 0, (size_t)n * sizeof(double), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.cg_iters = 0;
    ctx->stats.relres = 0.0;
    if (iters) *iters = 0;
    if (relres) *relres = 0.0;
    if (info) *info = 0;
    return B200_OK;
  }
  bool done = !(true2 > tol2);
  bool first = true;
  while (!done && it < maxiter) {
    if (fused) {
      RET(amg_cycle_fine_dia(ctx, A, r, z, z1, ax0, dir0, first ? nullptr : q, fs + F_ZQ));
    } else {
      RET(amg_cycle(ctx, A, 0, r, z, ax0, dir0));
      if (!first) LAUNCH(ctx, k_amg_dot1, g, AMG_T, 0, n, z, q, part, ticket, fs + F_ZQ);
    }
    if (multi && !first) RET(b200_allreduce_sum(ctx, fs + F_ZQ, 1));
    LAUNCH(ctx, k_amg_fcg_dir, g, AMG_T, 0, n, z, r, d, first ? 1 : 0, fs, part, ticket,
           fs + F_DR);
    if (multi) {
      RET(b200_allreduce_sum(ctx, fs + F_DR, 1));
      RET(b200_halo_exchange(ctx, A->fd.as<double>()));
    }
    RET(b200_spmv_dots(ctx, d, r, q, fs + F_DQ));          // fs[F_DQ] = d.q
    if (multi) RET(b200_allreduce_sum(ctx, fs + F_DQ, 3));
    LAUNCH(ctx, k_amg_fcg_update, g, AMG_T, 0, n, xt, r, d, q, diag, fs, part, ticket, fs);
    if (multi) RET(b200_allreduce_sum(ctx, fs + F_RR, 2));
    CK(cudaMemcpyAsync(ctx->scal_h, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ++it;
    first = false;
    true2 = ctx->scal_h[F_TRUE];
    const double dq = ctx->scal_h[F_DQ];
    if (ctx->debug)
      fprintf(stderr, "[b200 amg r%d] it=%lld true2=%.3e tol2=%.3e d.q=%.3e d.r=%.3e z.q=%.3e\n",
              ctx->rank, (long long)it, true2, tol2, dq, ctx->scal_h[F_DR], ctx->scal_h[F_ZQ]);
    if (!isfinite(true2) || !isfinite(dq)) { inf = -1; break; }
    if (!(dq > 0.0)) { inf = -1; break; }
    if (true2 <= tol2) {
      if (multi) RET(b200_halo_exchange(ctx, ctx->vx.as<double>()));
      RET(b200_spmv_dots(ctx, xt, bt, v, sc + SC_AUX0));
      RET(b200_launch_resid(ctx, bt, v, diag, r, tmp, sc + F_RR, sc + F_TRUE));
      if (multi) RET(b200_allreduce_sum(ctx, sc + F_RR, 2));
      CK(cudaMemcpyAsync(ctx->scal_h, sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost,
                         ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      true2 = ctx->scal_h[F_TRUE];
      if (true2 <= tol2 * (1.0 + 1e-6) || restarts >= 4) { done = true; break; }
      ++restarts;
      first = true;                // restart the direction from the true residual
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

// ================================================================== C-ABI glue
extern "C" int b200_set_precond(b200_ctx* ctx, int kind) {
  if (!ctx || (kind != B200_PRECOND_JACOBI && kind != B200_PRECOND_AMG)) return B200_ERR_ARG;
  ctx->precond = kind;
  return B200_OK;
}

extern "C" int b200_amg_setup(b200_ctx* ctx, int* nlevels) {
  if (!ctx || !ctx->have_sys) {
    if (ctx) ctx->err = "amg_setup: no system matrix";
    return B200_ERR_ARG;
  }
  if (ctx->nonsym) FAIL(B200_ERR_ARG, "amg: symmetric systems only");
  CK(cudaSetDevice(ctx->device));
  RET(b200_pcg_prepare(ctx, ctx->force_fmt));
  RET(amg_setup(ctx));
  if (nlevels) *nlevels = (int)amg_of(ctx)->nlev;
  return B200_OK;
}

extern "C" int b200_amg_level_shape(b200_ctx* ctx, int level, int64_t* n, int64_t* nnz) {
  if (!ctx || !ctx->amg) return B200_ERR_ARG;
  AmgState* A = amg_of(ctx);
  if (level < 0 || level >= (int)A->nlev) FAIL(B200_ERR_ARG, "amg: no such level");
  if (n) *n = A->lv[level].n;
  if (nnz) *nnz = A->lv[level].nnz;
  return B200_OK;
}

extern "C" int b200_amg_level_info(b200_ctx* ctx, int level, int64_t* out8) {
  if (!ctx || !ctx->amg || !out8) return B200_ERR_ARG;
  AmgState* A = amg_of(ctx);
  if (level < 0 || level >= (int)A->nlev) FAIL(B200_ERR_ARG, "amg: no such level");
  const AmgLevel& L = A->lv[level];
  out8[0] = L.n;
  out8[1] = L.nnz;
  out8[2] = L.ng;
  out8[3] = L.goff;
  out8[4] = L.dist ? 1 : 0;
  out8[5] = L.wlo;
  out8[6] = L.whi;
  out8[7] = (level + 1 < (int)A->nlev) ? L.coff : 0;
  return B200_OK;
}

// Slab runs: rows are the ones this rank owns; column ids and the aggregation map are
// exported in GLOBAL numbering, so that the ranks' parts stack to the global level.
extern "C" int b200_amg_export_level(b200_ctx* ctx, int level, int32_t* indptr, int32_t* indices,
                                     double* data, int32_t* agg) {
  if (!ctx || !ctx->amg) return B200_ERR_ARG;
  AmgState* A = amg_of(ctx);
  if (level < 0 || level >= (int)A->nlev) FAIL(B200_ERR_ARG, "amg: no such level");
  if (A->epoch != ctx->sys_epoch)
    FAIL(B200_ERR_ARG, "amg: the hierarchy is older than the system matrix (b200_amg_setup)");
  AmgLevel& L = A->lv[level];
  if (indptr)
    CK(cudaMemcpyAsync(indptr, L.ip, (size_t)(L.n + 1) * sizeof(int), cudaMemcpyDeviceToDevice,
                       ctx->stream));
  if (indices) {
    // level 0 already holds global ids; row-partitioned coarse levels local signed ones
    const int add = (L.dist && level > 0) ? (int)L.goff : 0;
    LAUNCH(ctx, k_amg_shift_ids, grid_for(L.nnz, AMG_T), AMG_T, 0, L.nnz, L.ix, add, 0, indices);
  }
  if (data)
    CK(cudaMemcpyAsync(data, L.av, (size_t)L.nnz * sizeof(double), cudaMemcpyDeviceToDevice,
                       ctx->stream));
  if (agg) {
    if (level + 1 >= (int)A->nlev) FAIL(B200_ERR_ARG, "amg: the last level has no aggregation");
    LAUNCH(ctx, k_amg_shift_ids, grid_for(L.n, AMG_T), AMG_T, 0, L.n, L.agg.as<int>(),
           L.dist ? (int)L.coff : 0, 1, agg);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
