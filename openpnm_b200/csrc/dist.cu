// dist.cu -- slab (row-block) partition plumbing: halo windows with the two
// neighbour ranks and the fp64 all-reduce of the PCG dot products, over NCCL.
//
// The reference's only statement of a parallel algorithm is its PETSc wrapper
// (openpnm/solvers/_petsc.py:117-128 contiguous row ownership, :160 KSP cg +
// PC jacobi, :225 gather).  Here: rank r owns rows [row_begin, row_end) of the
// index-ordered network; because the matrix is banded in that order (a cubic
// lattice has half-bandwidth Ny*Nz; a coordinate-sorted Delaunay graph a few
// planes' worth of pores) the halo of a vector is two contiguous windows of
// `pad` entries, sent without packing.
//
// NCCL is resolved at run time from the libnccl.so.2 already loaded by the
// host process (torch's), so the library has no link-time NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace {
struct NcclApi {
  bool ok = false;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool load_nccl() {
  if (g_nccl.ok) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
#define SYM(field, name)                                        \
  g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);        \
  if (!g_nccl.field) return false;
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(AllReduce, "ncclAllReduce")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.ok = true;
  return true;
}
}  // namespace

#define NCK(call)                                                             \
  do {                                                                        \
    ncclResult_t _r = (call);                                                 \
    if (_r != ncclSuccess) {                                                  \
      ctx->err = std::string(#call) + " -> " + g_nccl.GetErrorString(_r);     \
      return B200_ERR_NCCL;                                                   \
    }                                                                         \
  } while (0)

extern "C" int b200_nccl_unique_id(void* id128_h) {
  if (!id128_h || !load_nccl()) return B200_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return B200_ERR_NCCL;
  memcpy(id128_h, &id, sizeof id);
  return B200_OK;
}

extern "C" int b200_comm_init(b200_ctx* ctx, int nranks, int rank, const void* id128_h) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return B200_ERR_ARG;
  if (nranks == 1) { ctx->nranks = 1; ctx->rank = 0; return B200_OK; }
  if (!id128_h) FAIL(B200_ERR_ARG, "comm_init: null unique id");
  if (!load_nccl()) FAIL(B200_ERR_NCCL, "libnccl.so.2 not found");
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128_h, sizeof id);
  ncclComm_t comm;
  NCK(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->nccl = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  return B200_OK;
}

void b200_comm_release(b200_ctx* ctx) {
  if (ctx->nccl && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)ctx->nccl);
  ctx->nccl = nullptr;
}

extern "C" int b200_set_slab(b200_ctx* ctx, int64_t n_global, int64_t row_begin, int64_t row_end) {
  if (!ctx || row_begin < 0 || row_end > n_global || row_begin >= row_end) return B200_ERR_ARG;
  ctx->n_global = n_global;
  ctx->row_begin = row_begin;
  ctx->row_end = row_end;
  ctx->n = row_end - row_begin;
  return B200_OK;
}

// base points at the start of a padded vector [pad | n | pad]
static int halo_on(b200_ctx* ctx, double* base, cudaStream_t st) {
  ncclComm_t comm = (ncclComm_t)ctx->nccl;
  const int64_t pad = ctx->pad, h = ctx->halo, n = ctx->n;
  double* own = base + pad;
  NCK(g_nccl.GroupStart());
  if (ctx->rank > 0) {
    NCK(g_nccl.Send(own, (size_t)h, ncclDouble, ctx->rank - 1, comm, st));
    NCK(g_nccl.Recv(own - h, (size_t)h, ncclDouble, ctx->rank - 1, comm, st));
  }
  if (ctx->rank < ctx->nranks - 1) {
    NCK(g_nccl.Send(own + n - h, (size_t)h, ncclDouble, ctx->rank + 1, comm, st));
    NCK(g_nccl.Recv(own + n, (size_t)h, ncclDouble, ctx->rank + 1, comm, st));
  }
  NCK(g_nccl.GroupEnd());
  return B200_OK;
}

int b200_halo_exchange(b200_ctx* ctx, double* base) {
  if (ctx->nranks == 1 || ctx->halo == 0) return B200_OK;
  return halo_on(ctx, base, ctx->stream);
}

// Split form: the exchange runs on comm_stream while the caller's stream
// computes the rows that do not read the halo; _end makes the stream wait.
int b200_halo_exchange_begin(b200_ctx* ctx, double* base) {
  if (ctx->nranks == 1 || ctx->halo == 0) return B200_OK;
  CK(cudaEventRecord(ctx->ev_vec_ready, ctx->stream));
  CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_vec_ready, 0));
  RET(halo_on(ctx, base, ctx->comm_stream));
  CK(cudaEventRecord(ctx->ev_halo_done, ctx->comm_stream));
  return B200_OK;
}

int b200_halo_exchange_end(b200_ctx* ctx) {
  if (ctx->nranks == 1 || ctx->halo == 0) return B200_OK;
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_halo_done, 0));
  return B200_OK;
}

int b200_allreduce_sum(b200_ctx* ctx, double* dev, int count) {
  if (ctx->nranks == 1) return B200_OK;
  NCK(g_nccl.AllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl,
                       ctx->stream));
  return B200_OK;
}

// Vectors that every rank fills on its own segment and leaves zero elsewhere: the sum
// over ranks IS the concatenation (x + 0 is exact), i.e. an all-gather with uneven
// counts that needs no displacement bookkeeping.  amg.cu uses these to replicate the
// small coarse levels of the multigrid on every rank.
int b200_allreduce_sum_n(b200_ctx* ctx, double* dev, size_t count) {
  if (ctx->nranks == 1 || count == 0) return B200_OK;
  NCK(g_nccl.AllReduce(dev, dev, count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
  return B200_OK;
}

int b200_allreduce_sum_i32(b200_ctx* ctx, int* dev, size_t count) {
  if (ctx->nranks == 1 || count == 0) return B200_OK;
  NCK(g_nccl.AllReduce(dev, dev, count, ncclInt32, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
  return B200_OK;
}

// Halo windows of unequal widths (coarse multigrid levels): `owned` points at the
// first owned entry of a vector laid out [wlo | n | whi]; this rank receives wlo
// entries from rank-1 and whi from rank+1 and sends its first slo / last shi owned
// entries to them (slo == rank-1's whi, shi == rank+1's wlo: agreed on at setup).
int b200_halo_exchange_w(b200_ctx* ctx, void* owned, int64_t n, int elem_bytes, int64_t wlo,
                         int64_t whi, int64_t slo, int64_t shi) {
  if (ctx->nranks == 1) return B200_OK;
  if (wlo + whi + slo + shi == 0) return B200_OK;
  ncclComm_t comm = (ncclComm_t)ctx->nccl;
  const ncclDataType_t ty = elem_bytes == 8 ? ncclDouble : ncclInt32;
  char* o = static_cast<char*>(owned);
  const int64_t eb = elem_bytes;
  NCK(g_nccl.GroupStart());
  if (ctx->rank > 0) {
    if (slo > 0) NCK(g_nccl.Send(o, (size_t)slo, ty, ctx->rank - 1, comm, ctx->stream));
    if (wlo > 0) NCK(g_nccl.Recv(o - wlo * eb, (size_t)wlo, ty, ctx->rank - 1, comm, ctx->stream));
  }
  if (ctx->rank < ctx->nranks - 1) {
    if (shi > 0)
      NCK(g_nccl.Send(o + (n - shi) * eb, (size_t)shi, ty, ctx->rank + 1, comm, ctx->stream));
    if (whi > 0) NCK(g_nccl.Recv(o + n * eb, (size_t)whi, ty, ctx->rank + 1, comm, ctx->stream));
  }
  NCK(g_nccl.GroupEnd());
  return B200_OK;
}

// k int64 per rank -> all[nranks * k] on the host of every rank (one synchronisation)
int b200_allgather_i64_host(b200_ctx* ctx, const int64_t* mine, int k, int64_t* all) {
  if (ctx->nranks == 1) {
    for (int i = 0; i < k; ++i) all[i] = mine[i];
    return B200_OK;
  }
  const size_t tot = (size_t)ctx->nranks * k;
  CK(ctx->w_coll.ensure(tot * sizeof(long long)));
  long long* d = ctx->w_coll.as<long long>();
  CK(cudaMemsetAsync(d, 0, tot * sizeof(long long), ctx->stream));
  CK(cudaMemcpyAsync(d + (size_t)ctx->rank * k, mine, (size_t)k * sizeof(long long),
                     cudaMemcpyHostToDevice, ctx->stream));
  NCK(g_nccl.AllReduce(d, d, tot, ncclInt64, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
  CK(cudaMemcpyAsync(all, d, tot * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

int b200_allreduce_max_host(b200_ctx* ctx, int64_t* v) {
  if (ctx->nranks == 1) return B200_OK;
  CK(ctx->w_tmp2.ensure(64));
  long long* d = ctx->w_tmp2.as<long long>();
  long long h = *v;
  CK(cudaMemcpyAsync(d, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
  NCK(g_nccl.AllReduce(d, d, 1, ncclInt64, ncclMax, (ncclComm_t)ctx->nccl, ctx->stream));
  CK(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *v = h;
  return B200_OK;
}
