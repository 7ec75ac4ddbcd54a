/*
 * b200pnm.h -- C-ABI of libb200pnm.so: the B200-native transport-solve path
 * of OpenPNM (assembly -> boundary conditions -> sources -> Jacobi-PCG ->
 * Picard loop -> rate), hand-written CUDA for sm_100a.
 *
 * This is the drop-in boundary.  The reference (pure Python, /root/reference,
 * citations below are file:line under it) has no FFI of its own; the entry
 * points here are what a ctypes/cffi binding inside `openpnm.solvers` and
 * `openpnm.algorithms` would bind -- see INTEGRATION.md for that stub.
 *
 * Conventions
 *   - every function returns an int status: 0 = B200_OK, <0 = error
 *     (b200_last_error(ctx) gives the text).  No C++ exception crosses this
 *     boundary.  Solver outcome (converged / maxiter / breakdown) is NOT an
 *     error: it is reported through an `info` out-parameter with SciPy's
 *     convention (0 converged, >0 = iterations done when maxiter hit,
 *     <0 breakdown), which is what `Transport._run_special` stores into
 *     `soln.is_converged = not bool(exit_code)` (algorithms/_transport.py:221).
 *   - pointers are DEVICE pointers unless the name ends in `_h` (host).
 *     Scalar out-parameters (`int64_t* nnz`, `double* f`, ...) are host.
 *   - float data is float64; `conns`/pore lists are int64 as in the reference
 *     (network['throat.conns']); CSR/COO indices are int32 as in SciPy.
 *   - a context owns one device, one stream and all workspace; it is not
 *     thread-safe (the reference calls the path from one Python thread,
 *     GIL held: algorithms/_reactive_transport.py:196-210).
 */
#ifndef B200PNM_H
#define B200PNM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_ctx b200_ctx;

enum {
  B200_OK = 0,
  B200_ERR_CUDA = -1,          /* a CUDA runtime call failed                  */
  B200_ERR_ARG = -2,           /* bad argument / wrong call order            */
  B200_ERR_INDEX = -3,         /* pore index out of range in conns / COO     */
  B200_ERR_NONFINITE = -4,     /* inf/nan in A, b or x0 (ref: "A or b contains
                                  inf/nan values", _transport.py:253-257)     */
  B200_ERR_TOO_LARGE = -5,     /* nnz >= 2^31 (SciPy int32 index limit)      */
  B200_ERR_NCCL = -6,
  B200_ERR_SINGULAR = -7       /* a row has a zero diagonal and no BC        */
};

/* which matrix an export / spmv call refers to */
enum { B200_MAT_PURE = 0, B200_MAT_SYSTEM = 1 };

/* source-term models: openpnm/models/physics/source_terms/_funcs.py
 *   LINEAR            :73-116   r = p1*X + p2
 *   POWER_LAW         :120-168  r = p1*X^p2 + p3
 *   STANDARD_KINETICS :21-69    r = p1*X^p2                                 */
enum { B200_SRC_LINEAR = 0, B200_SRC_POWER_LAW = 1, B200_SRC_STANDARD_KINETICS = 2 };

/* storage the PCG picked for the Jacobi-scaled operator */
enum { B200_FMT_NONE = 0, B200_FMT_CSR = 1, B200_FMT_DIA = 2 };

/* ---------------------------------------------------------------- lifecycle */
int b200_version(void);
int b200_create(int device, b200_ctx** out);
int b200_destroy(b200_ctx* ctx);
const char* b200_last_error(b200_ctx* ctx);
/* stream all work of this context is enqueued on (a cudaStream_t).  The host
 * may hand in its own (e.g. torch.cuda.current_stream().cuda_stream) so that
 * its copies and this library's kernels are ordered without extra syncs.    */
void* b200_stream(b200_ctx* ctx);
int b200_set_stream(b200_ctx* ctx, void* cuda_stream);
int b200_synchronize(b200_ctx* ctx);

/* ----------------------------------------------------- K1: Laplacian assembly
 * Replaces Network.create_adjacency_matrix(weights=g, fmt='coo')
 * (network/_network.py:265-293) + scipy.sparse.csgraph.laplacian
 * (algorithms/_transport.py:113-114) + A.tocsr() (solvers/_scipy.py:13-14):
 * conns (Nt,2) int64 C-order, g (Nt,) f64 -> canonical CSR of L = D - W
 * (column-sorted, duplicate throats summed, self-loops dropped, explicit
 * zeros kept), built without floating-point atomics: integer histogram ->
 * exclusive scan -> fill -> per-row sort on (col, throat) -> fixed-order
 * segmented sums.  Rows [row_begin,row_end) only (slab ownership); pass
 * 0,Np for the whole network.  Column indices stay global.                  */
int b200_build_laplacian(b200_ctx* ctx, const int64_t* conns, const double* g,
                         int64_t Np, int64_t Nt, int64_t row_begin,
                         int64_t row_end, int64_t* nnz);

/* Non-symmetric variant for AdvectionDiffusion
 * (openpnm/algorithms/_advection_diffusion.py:43-105; (Nt,2) weights,
 * network/_network.py:281-285): g2 is (Nt,2) row-major, entry (c0,c1) carries
 * g2[t][0], entry (c1,c0) carries g2[t][1]; the diagonal is the column sum as
 * in scipy's in-degree Laplacian.  Systems built from it are solved by
 * Jacobi-scaled BiCGStab inside b200_pcg / b200_picard (slab runs included).
 * b200_add_to_diagonal adds the outflow BC values (NaN = none) to the system
 * diagonal after b200_apply_bcs(keep_zero_diag = 1).                        */
int b200_build_laplacian2(b200_ctx* ctx, const int64_t* conns, const double* g2,
                          int64_t Np, int64_t Nt, int64_t* nnz);
/* rows [row_begin,row_end) only, from the throats touching them (slab runs)  */
int b200_build_laplacian2_rows(b200_ctx* ctx, const int64_t* conns,
                               const double* g2, int64_t Np, int64_t Nt,
                               int64_t row_begin, int64_t row_end, int64_t* nnz);
int b200_add_to_diagonal(b200_ctx* ctx, const double* add);

/* ------------------------------------------------- topology health check
 * Replaces Transport._validate_topology_health (_transport.py:243-251) ->
 * topotools.is_fully_connected (topotools/_topotools.py:993-1024): counts the
 * connected clusters of the network and those that contain no pore with a
 * boundary condition (NaN = no BC in bc_value / bc_rate, either may be NULL).
 * The reference raises "Your network is clustered ..." when the second count
 * is non-zero.                                                               */
int b200_check_topology(b200_ctx* ctx, const int64_t* conns, int64_t Nt,
                        int64_t Np, const double* bc_value,
                        const double* bc_rate, int64_t* n_clusters,
                        int64_t* n_without_bc);

/* Cluster labels (labels[i] = smallest pore index of pore i's cluster; device, Np
 * int32) -- `csgraph.connected_components` in topotools/_topotools.py:1014 and
 * models/network/_health.py:55-62.                                           */
int b200_cluster_labels(b200_ctx* ctx, const int64_t* conns, int64_t Nt,
                        int64_t Np, int32_t* labels, int64_t* n_clusters);

/* The same health check in a slab run: `conns` are the throats touching this
 * rank's rows [row_begin,row_end) (global pore ids), `halo` the network's index
 * bandwidth max|c1-c0|.  Clusters that stay away from the two cut regions
 * [row_begin-halo,row_begin+halo) and [row_end-halo,row_end+halo) are complete
 * and judged here (*interior_without_bc); for the pores of the cut regions the
 * local cluster label and its has-a-BC-pore mark are exported (device,
 * 4*halo entries each: low region then high region, -1 outside the network) so
 * that the host can join clusters across ranks (openpnm_b200/dist.py).       */
int b200_slab_clusters(b200_ctx* ctx, const int64_t* conns, int64_t Nt,
                       int64_t Np, int64_t row_begin, int64_t row_end,
                       int64_t halo, const double* bc_value,
                       const double* bc_rate, int32_t* win_labels,
                       uint8_t* win_has_bc, int64_t* interior_without_bc);

/* ------------------------------------------ network health + trim (f-2)
 * Replaces utils.check_network_health (openpnm/utils/_health.py:40-98) with the
 * models it calls (openpnm/models/network/_health.py:11-76): per pore the size
 * of its cluster and whether it is isolated (no throat touches it), per throat
 * whether it loops back onto its own pore; any output may be NULL.
 * counts6_h (host) = {clusters, isolated pores, pores outside the largest
 * cluster ('disconnected_pores'), looped throats, 0, size of the largest
 * cluster}.  Throats pointing at non-existent pores are B200_ERR_INDEX.
 * b200_trim replaces topotools.trim (openpnm/topotools/_topotools.py:157-229):
 * pores with pore_keep == 0 go together with every throat touching them,
 * throats with throat_keep == 0 (NULL: keep all) too; survivors keep their
 * order.  pore_map / throat_map (device, int64): new index or -1; conns_out
 * (device, room for Nt pairs): the remapped connections of the kept throats. */
int b200_network_health(b200_ctx* ctx, const int64_t* conns, int64_t Nt,
                        int64_t Np, int32_t* cluster_size, uint8_t* isolated,
                        uint8_t* looped, int32_t* labels, int64_t* counts6_h);
int b200_trim(b200_ctx* ctx, const int64_t* conns, int64_t Nt, int64_t Np,
              const uint8_t* pore_keep, const uint8_t* throat_keep,
              int64_t* pore_map, int64_t* throat_map, int64_t* conns_out,
              int64_t* Np_new, int64_t* Nt_new);

/* ------------------------------------------------ K2: boundary conditions
 * Replaces Transport._build_b + _apply_BCs (_transport.py:117-121,145-169).
 * bc_value / bc_rate: (Np,) f64 for ALL pores (global indexing), NaN = no BC;
 * either may be NULL.  Produces the system matrix (BC rows/cols eliminated,
 * f on BC diagonals, zeros compacted away = eliminate_zeros) and b.
 * f = mean of the pure diagonal; for a slab, pass the global sum and count
 * through b200_set_diag_mean() first, otherwise the local mean is used.     */
int b200_apply_bcs(b200_ctx* ctx, const double* bc_value, const double* bc_rate,
                   int keep_zero_diag, double* f, int64_t* nnz);
int b200_pure_diag_sum(b200_ctx* ctx, double* sum, int64_t* count);
int b200_set_diag_mean(b200_ctx* ctx, double f);

/* ------------------------------------------------------------ generic system
 * For solver.solve(A, b, x0) where A was assembled by the reference itself
 * (solvers/_base.py:6-14, _transport.py:213): COO triples (unsorted, possibly
 * duplicated; exactly what `alg.A` is) -> canonical system CSR like
 * `A.tocsr()`.                                                              */
int b200_load_coo(b200_ctx* ctx, const int32_t* row, const int32_t* col,
                  const double* data, int64_t nnz_in, int64_t n, int64_t* nnz);
int b200_load_csr(b200_ctx* ctx, const int32_t* indptr, const int32_t* indices,
                  const double* data, int64_t n, int64_t nnz);
int b200_set_b(b200_ctx* ctx, const double* b);
int b200_get_b(b200_ctx* ctx, double* b);

/* sizes + export (device destinations) for the bit-exact index parity tests */
int b200_csr_shape(b200_ctx* ctx, int which, int64_t* nrows, int64_t* nnz);
int b200_export_csr(b200_ctx* ctx, int which, int32_t* indptr, int32_t* indices,
                    double* data);

/* ----------------------------------------------------------- K3: SpMV
 * y = M x for the pure or the system matrix (CSR, fp64/int32).  x, y (n,).
 * `A * x` at _transport.py:159 and _reactive_transport.py:244.              */
int b200_spmv(b200_ctx* ctx, int which, const double* x, double* y);

/* ------------------------------------------------- K4: Jacobi-PCG
 * Replaces solver.solve(A, b, x0) -> (x, exit_code) (solvers/_scipy.py:8-26).
 * Solves system * x = b by CG on the symmetrically Jacobi-scaled operator
 * D^-1/2 A D^-1/2 (unit diagonal; identical iterates to Jacobi-PCG in exact
 * arithmetic).  Stops when ||b - A x||_2 <= max(rtol*||b||_2, atol) -- the
 * rule ScipyCG uses (tol=self.tol, atol=norm(b)*tol).  x0 may be NULL
 * (zeros) and is never written.  b200_pcg_prepare() (re)builds the scaled
 * operator after the system matrix changed; b200_pcg() calls it if needed. */
int b200_set_format(b200_ctx* ctx, int force_format);   /* 0 = automatic */
int b200_pcg_prepare(b200_ctx* ctx, int force_format);
int b200_pcg(b200_ctx* ctx, const double* x0, double rtol, double atol,
             int64_t maxiter, double* x, int64_t* iters, double* relres,
             int* info);
int b200_pcg_format(b200_ctx* ctx, int* fmt, int* ndiags);

/* ------------------------------- preconditioner behind the same solve call
 * SURVEY section 8 f-1; the reference's analogue is PyamgRugeStubenSolver
 * (openpnm/solvers/_pyamg.py:10-17).  B200_PRECOND_AMG makes b200_pcg (and the
 * solves inside b200_picard) run flexible CG preconditioned by an aggregation
 * multigrid K-cycle built on the device from the system matrix (csrc/amg.cu):
 * ~30 iterations at any size instead of O(N) for an N^3 lattice.  Same
 * stopping rule, same outputs (`iters` counts outer FCG iterations).  Applies
 * to symmetric systems with more than 2048 rows; otherwise, and for matrices
 * that cannot be coarsened, the call silently stays with Jacobi.
 * Slab runs (b200_comm_init with several ranks; the row-block ownership of the
 * reference's PETSc wrapper, solvers/_petsc.py:117-128): ONE hierarchy for the
 * whole problem, every level row-partitioned by the same slabs (aggregates do
 * not cross rank boundaries, the Galerkin products keep the couplings between
 * neighbouring ranks) until a level is small enough to be replicated on every
 * rank.  Collective: every rank must make the same calls.
 * b200_amg_setup builds the hierarchy now (it is otherwise built by the first
 * solve after the matrix changed); the accessors export a level's CSR (level
 * 0 = the system matrix; rows this rank holds, GLOBAL column ids) and its
 * aggregation map (row -> global coarse row, -1 for rows left off the coarse
 * grid) for the Galerkin-identity tests.  b200_amg_level_info fills
 * {rows held, entries held, global rows, global index of the first held row,
 *  1 if row-partitioned / 0 if held in full, halo entries from rank-1, from
 *  rank+1, offset of this rank's coarse rows in the next level}.            */
enum { B200_PRECOND_JACOBI = 0, B200_PRECOND_AMG = 1 };
int b200_set_precond(b200_ctx* ctx, int kind);
int b200_amg_setup(b200_ctx* ctx, int* nlevels);
int b200_amg_level_shape(b200_ctx* ctx, int level, int64_t* n, int64_t* nnz);
int b200_amg_level_info(b200_ctx* ctx, int level, int64_t* out8);
int b200_amg_export_level(b200_ctx* ctx, int level, int32_t* indptr,
                          int32_t* indices, double* data, int32_t* agg);

/* ------------------------------------------- K5: sources + Picard loop
 * Replaces ReactiveTransport.set_source bookkeeping hand-off,
 * _update_iterative_props (_algorithm.py:69-91), _apply_sources
 * (_reactive_transport.py:125-148), _get_residual (:234-244) and the outer
 * loop _run_special (:150-213) including SciPy's TerminationCondition
 * (2-norms, f_rtol AND x_rtol) and the first-iteration dx == 0 quirk
 * (_solution.py:17-20).  mask: (n,) uint8; p1..p3: (n,) f64 (p3 may be NULL).
 * With zero sources the loop degenerates to one linear solve (num_iter 1).  */
int b200_clear_sources(b200_ctx* ctx);
int b200_add_source(b200_ctx* ctx, int kind, const uint8_t* mask,
                    const double* p1, const double* p2, const double* p3);
int b200_picard(b200_ctx* ctx, const double* x0, double relaxation,
                int64_t newton_maxiter, double f_rtol, double x_rtol,
                double cg_rtol, int64_t cg_maxiter, double* x,
                int64_t* num_iter, int* converged, int64_t* cg_iters_total,
                int* last_info);

/* Same loop with the reference's `_update_iterative_props` hook
 * (_algorithm.py:69-91, called from _update_A_and_b, _reactive_transport.py:
 * 226-232): `update(user, x_dev)` runs before the first residual and after
 * every relaxed update, with the stream idle and x (n,) on the device.  It may
 * regenerate host-side pore-scale models at x and hand the result back with
 * b200_build_laplacian / b200_apply_bcs (variable conductance) and
 * b200_clear_sources / b200_add_source (B200_SRC_LINEAR with S1, S2 arrays for
 * a model that has no device kernel).  Non-zero return aborts the loop with
 * B200_ERR_ARG.  Single GPU only.                                           */
typedef int (*b200_update_fn)(void* user, const double* x_dev);
int b200_picard_cb(b200_ctx* ctx, const double* x0, double relaxation,
                   int64_t newton_maxiter, double f_rtol, double x_rtol,
                   double cg_rtol, int64_t cg_maxiter, double* x,
                   int64_t* num_iter, int* converged, int64_t* cg_iters_total,
                   int* last_info, b200_update_fn update, void* user);

/* ------------------- conductance that follows the quantity, on the device (f-3)
 * For the usual chain  X -> pore property P(X) = sum_k poly[k] X^k
 * (models/misc polynomial) -> throat value (given array; or mean / min / max of
 * the two pore values, models/misc from_neighbor_pores: interp 0 / 1 / 2) ->
 * conduit conductance
 *   B200_COND_HYDRAULIC  generic_hydraulic, g_i = F_i / mu_i in series
 *                        (models/physics/hydraulic_conductance/_funcs.py:37-53)
 *   B200_COND_POISSON    _poisson_conductance, g_i = D_i F_i in series
 *                        (models/physics/_utils.py:46-61; generic_diffusive,
 *                        generic_thermal, generic_electrical)
 * with size factors (Nt,3) row-major (sf_cols = 3) or (Nt,) (sf_cols = 1).
 * While a model is set, b200_picard re-evaluates g at the current iterate,
 * re-assembles (K1) and re-applies the BCs (K2, the arrays given here) before
 * every linearisation: `_update_iterative_props` + `_update_A_and_b`
 * (_algorithm.py:69-91, _reactive_transport.py:226-232) without leaving the
 * device.  All arrays are caller-owned device memory and must outlive the
 * model; poly_h is host.  Single GPU.                                        */
enum { B200_COND_HYDRAULIC = 0, B200_COND_POISSON = 1 };
int b200_set_conductance_model(b200_ctx* ctx, int kind, const int64_t* conns,
                               int64_t Nt, int64_t Np, const double* poly_h,
                               int npoly, int interp, const double* throat_prop,
                               const double* size_factors, int sf_cols,
                               const double* bc_value, const double* bc_rate,
                               int keep_zero_diag);
int b200_clear_conductance_model(b200_ctx* ctx);
int b200_update_conductance(b200_ctx* ctx, const double* x);
int b200_conductance_values(b200_ctx* ctx, double* g_out);

/* ------------------------------------------------ transient (explicit RK45)
 * Replaces TransientReactiveTransport.run + _build_rhs
 * (openpnm/algorithms/_transient_reactive_transport.py:51-123) with
 * integrators.ScipyRK45 (openpnm/integrators/_scipy.py:18-56):
 *   dy/dt = (-A y + b) / V,  A, b = system after BCs with the registered
 * sources linearised at y, integrated from t0 to t1 by SciPy's RK45 scheme
 * (Dormand-Prince 5(4), same step controller and dense output).  x0 must
 * already carry the BC values on the BC pores.  saveat_h (host, ascending,
 * within [t0, t1]) are the output times (t_eval); nsave = 0 stores every
 * accepted step like t_eval=None.  out: device, `capacity` rows of n;
 * times_h: host, `capacity` doubles.  status: 0 finished, -1 step size too
 * small, -2 capacity exhausted.  Slab runs (b200_comm_init): x0, volume and
 * out hold the rank's rows; the halo of every right-hand side's operand is
 * exchanged and the controller's norms are all-reduced, so all ranks take the
 * same steps (symmetric systems; collective).                                */
int b200_transient_rk45(b200_ctx* ctx, const double* x0, const double* volume,
                        double t0, double t1, double rtol, double atol,
                        const double* saveat_h, int64_t nsave, int64_t capacity,
                        double* out, double* times_h, int64_t* nout,
                        int64_t* nsteps, int64_t* nfev, int* status);

/* ---------------------------------------------------------------- K6: rate
 * Replaces Transport.rate (_transport.py:259-330).  Pores: net rate leaving
 * each pore = (L_pure x)[pore]; out has k entries (mode 'single') and the
 * caller sums for 'group'.  Throats: |g (x[c1]-x[c0])|.                     */
int b200_rate_pores(b200_ctx* ctx, const double* x, const int64_t* pores,
                    int64_t k, double* out);
int b200_rate_throats(b200_ctx* ctx, const int64_t* conns, const double* g,
                      const double* x, const int64_t* throats, int64_t k,
                      double* out);

/* (Nt,2) conductance of AdvectionDiffusion: |g2[t][0] x[c1] - g2[t][1] x[c0]|          */
int b200_rate_throats2(b200_ctx* ctx, const int64_t* conns, const double* g2,
                       const double* x, const int64_t* throats, int64_t k,
                       double* out);
/* S1, S2 and rate of registered source number `index` at x for every owned pore
 * (device, n each): what `regenerate_models` leaves on the phase after a run
 * (_algorithm.py:89-91, source_terms/_funcs.py:61-69,108-116,158-168).        */
int b200_source_values(b200_ctx* ctx, int index, const double* x, double* S1,
                       double* S2, double* rate);

/* ------------------------------------------------------------- statistics */
typedef struct b200_stats {
  int64_t n, nnz_system, cg_iters;
  int fmt, ndiags;
  double t_setup_ms, t_solve_ms;      /* CUDA-event times of the last pcg     */
  double relres;
  int64_t kernel_launches;            /* kernels launched by this ctx so far  */
} b200_stats;
int b200_get_stats(b200_ctx* ctx, b200_stats* out);

/* Measurement hook for bench.py: average CUDA-event time of `reps` back-to-back
 * launches of one kernel of the prepared PCG on its resident data (call after
 * a solve; it scribbles on the PCG work vectors).  which: 0 = K_A SpMV + dot
 * on the scaled operator, 1 = K_B x/r update + norms, 2 = K_C p update,
 * 3 = K3 plain CSR SpMV of the system matrix.                                */
int b200_time_kernel(b200_ctx* ctx, int which, int reps, double* ms_avg);
/* Bytes one K_A launch must move in the operator's storage format (DIA-sym: nd+3
 * doubles per row; irregular: 12 B per stored entry of the sliced-ELL copy, padding
 * included, + permutation + the three vectors) and the stored entries.          */
int b200_operator_bytes(b200_ctx* ctx, int64_t* bytes, int64_t* entries);

/* --------------------------------------------------- host-buffer entry points
 * What a ctypes binder holding numpy arrays calls; H2D/D2H copies inside.
 * b200_solve_coo_h is the whole of `B200PCG.solve(A, b, x0)`.               */
int b200_solve_coo_h(b200_ctx* ctx, const int32_t* row_h, const int32_t* col_h,
                     const double* data_h, int64_t nnz_in, int64_t n,
                     const double* b_h, const double* x0_h, double rtol,
                     double atol, int64_t maxiter, double* x_h, int64_t* iters,
                     double* relres, int* info);

/* ------------------------------------------------------ multi-GPU (slabs)
 * One context per rank.  The unique id is produced on rank 0 and broadcast by
 * the host (torch.distributed); after b200_comm_init the PCG exchanges halo
 * windows with rank-1/rank+1 and all-reduces its dot products over NCCL.
 * The reference's only statement of this algorithm is its PETSc wrapper
 * (row-block ownership, CG + Jacobi: solvers/_petsc.py:117-128,160).        */
int b200_nccl_unique_id(void* id128_h);
int b200_comm_init(b200_ctx* ctx, int nranks, int rank, const void* id128_h);
int b200_set_slab(b200_ctx* ctx, int64_t n_global, int64_t row_begin,
                  int64_t row_end);
/* Fused communication over NVLink peer memory (optional, lattice/DIA operator):
 * after b200_pcg_prepare every rank exports two 64-byte CUDA-IPC handles
 * (its mailbox block and its direction vector p); the host all-gathers them
 * together with every rank's row count and attaches.  From then on the PCG
 * iteration makes no NCCL call: K_A's last block stores its dot products into
 * every peer's mailbox, K_BC sums the mailboxes in rank order and stores the
 * edge entries of the new p straight into the neighbours' halo windows.     */
int b200_p2p_export(b200_ctx* ctx, void* handles128_h);
int b200_p2p_attach(b200_ctx* ctx, const void* all_handles_h, const int64_t* all_rows_h);
int b200_p2p_disable(b200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200PNM_H */
