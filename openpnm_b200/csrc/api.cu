// api.cu -- context lifecycle, statistics and the host-buffer entry points of
// the C-ABI declared in include/b200pnm.h.
#include <stdlib.h>

#include "common.cuh"

extern "C" int b200_version(void) { return 100; }

extern "C" int b200_create(int device, b200_ctx** out) {
  if (!out) return B200_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return B200_ERR_CUDA;   // no CPU fallback: the product path needs the GPU
  b200_ctx* ctx = new b200_ctx();
  ctx->device = device;
  const char* dbg = getenv("B200_DEBUG");
  ctx->debug = dbg && dbg[0] && dbg[0] != '0';
  const char* ch = getenv("B200_KA_CHUNK");
  if (ch && atoi(ch) >= 1 && atoi(ch) <= 64) ctx->ka_chunk = atoi(ch);
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreate(&ctx->ev2) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_vec_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_halo_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaMallocHost(&ctx->scal_h, SC_COUNT * sizeof(double)) != cudaSuccess ||
      cudaMallocHost(&ctx->flags_h, FLAG_COUNT * sizeof(int)) != cudaSuccess) {
    delete ctx;
    return B200_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (ctx->num_sms < 1) ctx->num_sms = 148;
  *out = ctx;
  return B200_OK;
}

extern "C" int b200_destroy(b200_ctx* ctx) {
  if (!ctx) return B200_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  b200_amg_release(ctx);
  b200_p2p_release(ctx);
  b200_comm_release(ctx);
  for (DevBuf* b : {&ctx->pure_indptr, &ctx->pure_indices, &ctx->pure_data, &ctx->pure_diag,
                    &ctx->sys_indptr, &ctx->sys_indices, &ctx->sys_data, &ctx->sys_diag,
                    &ctx->sys_diagpos, &ctx->b, &ctx->diag_eff, &ctx->b_eff, &ctx->sc, &ctx->dexp,
                    &ctx->dia_store, &ctx->o_indptr, &ctx->o_cols, &ctx->o_vals, &ctx->o_rowstart,
                    &ctx->vx, &ctx->vr, &ctx->vp, &ctx->vap, &ctx->vb, &ctx->scal, &ctx->w_cnt,
                    &ctx->w_keys, &ctx->w_scan, &ctx->w_part, &ctx->w_ticket, &ctx->w_flags,
                    &ctx->w_tmp0, &ctx->w_tmp1, &ctx->w_tmp2, &ctx->vbi, &ctx->h_row, &ctx->h_col,
                    &ctx->h_dat, &ctx->h_vec, &ctx->w_coll, &ctx->s_perm, &ctx->s_ptr,
                    &ctx->s_cols, &ctx->s_vals, &ctx->gm_g})
    b->release();
  if (ctx->scal_h) cudaFreeHost(ctx->scal_h);
  if (ctx->flags_h) cudaFreeHost(ctx->flags_h);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->ev2) cudaEventDestroy(ctx->ev2);
  if (ctx->ev_vec_ready) cudaEventDestroy(ctx->ev_vec_ready);
  if (ctx->ev_halo_done) cudaEventDestroy(ctx->ev_halo_done);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return B200_OK;
}

extern "C" const char* b200_last_error(b200_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" void* b200_stream(b200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int b200_set_stream(b200_ctx* ctx, void* cuda_stream) {
  if (!ctx) return B200_ERR_ARG;
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return B200_OK;
}

extern "C" int b200_synchronize(b200_ctx* ctx) {
  if (!ctx) return B200_ERR_ARG;
  CK(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

extern "C" int b200_get_stats(b200_ctx* ctx, b200_stats* out) {
  if (!ctx || !out) return B200_ERR_ARG;
  ctx->stats.kernel_launches = ctx->launches;
  *out = ctx->stats;
  return B200_OK;
}

// Whole of `solver.solve(A, b, x0)` for a caller that holds host arrays:
// H2D of the COO triples and vectors, COO -> CSR, scaling, PCG, D2H of x.
extern "C" int b200_solve_coo_h(b200_ctx* ctx, const int32_t* row_h, const int32_t* col_h,
                                const double* data_h, int64_t nnz_in, int64_t n,
                                const double* b_h, const double* x0_h, double rtol, double atol,
                                int64_t maxiter, double* x_h, int64_t* iters, double* relres,
                                int* info) {
  if (!ctx || !row_h || !col_h || !data_h || !b_h || !x_h || n <= 0 || nnz_in < 0)
    return B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  // staging buffers live in the context: a Picard loop of the reference calls
  // this thousands of times with the same sizes (_reactive_transport.py:196-210)
  DevBuf& drow = ctx->h_row;
  DevBuf& dcol = ctx->h_col;
  DevBuf& ddat = ctx->h_dat;
  DevBuf& dvec = ctx->h_vec;
  int rc = B200_OK;
#define CKX(call)                                                           \
  do {                                                                      \
    cudaError_t _e = (call);                                                \
    if (_e != cudaSuccess) {                                                \
      ctx->err = std::string(#call) + " -> " + cudaGetErrorString(_e);      \
      return B200_ERR_CUDA;                                                 \
    }                                                                       \
  } while (0)
  CKX(drow.ensure((size_t)nnz_in * sizeof(int)));
  CKX(dcol.ensure((size_t)nnz_in * sizeof(int)));
  CKX(ddat.ensure((size_t)nnz_in * sizeof(double)));
  CKX(dvec.ensure((size_t)3 * n * sizeof(double)));
  double* db = dvec.as<double>();
  double* dx0 = db + n;
  double* dx = dx0 + n;
  CKX(cudaMemcpyAsync(drow.p, row_h, (size_t)nnz_in * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CKX(cudaMemcpyAsync(dcol.p, col_h, (size_t)nnz_in * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CKX(cudaMemcpyAsync(ddat.p, data_h, (size_t)nnz_in * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CKX(cudaMemcpyAsync(db, b_h, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (x0_h)
    CKX(cudaMemcpyAsync(dx0, x0_h, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int64_t nnz = 0;
  b200_clear_sources(ctx);
  rc = b200_load_coo(ctx, drow.as<int>(), dcol.as<int>(), ddat.as<double>(), nnz_in, n, &nnz);
  if (rc == B200_OK) rc = b200_set_b(ctx, db);
  if (rc == B200_OK)
    rc = b200_pcg(ctx, x0_h ? dx0 : nullptr, rtol, atol, maxiter, dx, iters, relres, info);
  if (rc == B200_OK) {
    CKX(cudaMemcpyAsync(x_h, dx, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CKX(cudaStreamSynchronize(ctx->stream));
  }
#undef CKX
  return rc;
}
