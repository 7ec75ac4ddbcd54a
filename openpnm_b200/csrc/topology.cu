// topology.cu -- device version of the pre-solve health check
// `Transport._validate_topology_health` (openpnm/algorithms/_transport.py:243-251)
// -> `topotools.is_fully_connected(network, pores_BC)`
// (openpnm/topotools/_topotools.py:993-1024): every connected cluster of the
// network must contain at least one pore that carries a boundary condition.
// The reference builds a LIL adjacency (Python lists; 5 s at 1 M pores) and calls
// scipy's connected_components; here: lock-free union-find over the throat
// list (integer CAS hooking of the larger root under the smaller, path
// halving), so the root of a cluster is its smallest pore index -- a
// deterministic labelling -- followed by one pass that marks the roots
// reached by a BC pore.
#include "common.cuh"

__device__ __forceinline__ int uf_find(int* __restrict__ parent, int x) {
  int p = parent[x];
  while (p != x) {
    const int gp = parent[p];
    if (gp != p) parent[x] = gp;     // path halving; benign race (only ever moves up)
    x = p;
    p = gp;
  }
  return x;
}

__global__ void k_uf_init(int* __restrict__ parent, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    parent[i] = (int)i;
}

__global__ void k_uf_union(const longlong2* __restrict__ conns, int64_t Nt, int64_t Np,
                           int* __restrict__ parent, int* __restrict__ flags) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < Nt; t += stride) {
    const longlong2 c = conns[t];
    if (c.x < 0 || c.x >= Np || c.y < 0 || c.y >= Np) { flags[FLAG_INDEX] = 1; continue; }
    int u = (int)c.x, v = (int)c.y;
    while (true) {
      u = uf_find(parent, u);
      v = uf_find(parent, v);
      if (u == v) break;
      if (u < v) { const int tmp = u; u = v; v = tmp; }     // u is the larger root
      if (atomicCAS(&parent[u], u, v) == u) break;          // hook; retry if u stopped being a root
    }
  }
}

__global__ void k_uf_compress(int* __restrict__ parent, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int r = (int)i;
    while (parent[r] != r) r = parent[r];
    parent[i] = r;
  }
}

__global__ void k_mark_bc_roots(const int* __restrict__ parent, const double* __restrict__ bcv,
                                const double* __restrict__ bcr, int64_t n,
                                unsigned char* __restrict__ has_bc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const bool bc = (bcv && !isnan(bcv[i])) || (bcr && !isnan(bcr[i]));
    if (bc) has_bc[parent[i]] = 1;
  }
}

__global__ void k_count_clusters(const int* __restrict__ parent,
                                 const unsigned char* __restrict__ has_bc, int64_t n,
                                 unsigned long long* __restrict__ counts) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long roots = 0, bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (parent[i] == (int)i) { ++roots; if (!has_bc[i]) ++bad; }
  }
  // integer counts: atomics are exact and order-independent
  for (int o = 16; o > 0; o >>= 1) {
    roots += __shfl_down_sync(0xffffffffu, roots, o);
    bad += __shfl_down_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (roots) atomicAdd(&counts[0], roots);
    if (bad) atomicAdd(&counts[1], bad);
  }
}

extern "C" int b200_check_topology(b200_ctx* ctx, const int64_t* conns, int64_t Nt, int64_t Np,
                                   const double* bc_value, const double* bc_rate,
                                   int64_t* n_clusters, int64_t* n_without_bc) {
  if (!ctx || Np <= 0 || Nt < 0 || Np >= (1ll << 31)) return B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  RET(b200_zero_flags(ctx));
  CK(ctx->w_cnt.ensure((size_t)(Np + 1) * sizeof(int)));
  CK(ctx->w_tmp1.ensure((size_t)Np + 64));
  CK(ctx->w_tmp2.ensure(256));
  int* parent = ctx->w_cnt.as<int>();
  unsigned char* has_bc = ctx->w_tmp1.as<unsigned char>();
  unsigned long long* counts = ctx->w_tmp2.as<unsigned long long>();
  CK(cudaMemsetAsync(has_bc, 0, (size_t)Np, ctx->stream));
  CK(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), ctx->stream));
  const int B = 256;
  LAUNCH(ctx, k_uf_init, grid_for(Np, B), B, 0, parent, Np);
  if (Nt > 0)
    LAUNCH(ctx, k_uf_union, grid_for(Nt, B), B, 0, reinterpret_cast<const longlong2*>(conns), Nt,
           Np, parent, ctx->w_flags.as<int>());
  LAUNCH(ctx, k_uf_compress, grid_for(Np, B), B, 0, parent, Np);
  LAUNCH(ctx, k_mark_bc_roots, grid_for(Np, B), B, 0, parent, bc_value, bc_rate, Np, has_bc);
  LAUNCH(ctx, k_count_clusters, grid_for(Np, B), B, 0, parent, has_bc, Np, counts);
  CK(cudaGetLastError());
  unsigned long long h[2] = {0, 0};
  CK(cudaMemcpyAsync(h, counts, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  RET(b200_check_flags(ctx));
  if (n_clusters) *n_clusters = (int64_t)h[0];
  if (n_without_bc) *n_without_bc = (int64_t)h[1];
  return B200_OK;
}
