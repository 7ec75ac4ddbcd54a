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

// clusters(ctx, ...) leaves parent[] (= labels) in w_cnt and has_bc[] in w_tmp1
static int clusters(b200_ctx* ctx, const int64_t* conns, int64_t Nt, int64_t Np,
                    const double* bc_value, const double* bc_rate, int64_t* n_clusters,
                    int64_t* n_without_bc) {
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

extern "C" int b200_check_topology(b200_ctx* ctx, const int64_t* conns, int64_t Nt, int64_t Np,
                                   const double* bc_value, const double* bc_rate,
                                   int64_t* n_clusters, int64_t* n_without_bc) {
  return clusters(ctx, conns, Nt, Np, bc_value, bc_rate, n_clusters, n_without_bc);
}

extern "C" int b200_cluster_labels(b200_ctx* ctx, const int64_t* conns, int64_t Nt, int64_t Np,
                                   int32_t* labels, int64_t* n_clusters) {
  RET(clusters(ctx, conns, Nt, Np, nullptr, nullptr, n_clusters, nullptr));
  if (labels)
    CK(cudaMemcpyAsync(labels, ctx->w_cnt.p, (size_t)Np * sizeof(int), cudaMemcpyDeviceToDevice,
                       ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

// ------------------------------------------------ slab run: the check per rank
// A rank sees the throats that touch its rows [rb, re).  Clusters of that sub-network
// which stay away from the two cut regions [rb-halo, rb+halo) and [re-halo, re+halo)
// are complete and are judged here; the others may continue on a neighbour rank, so
// their labels and BC marks on the cut regions are handed to the host, which joins
// them across the ranks (openpnm_b200/dist.py: slab_topology_check).
__global__ void k_mark_owned_window(const int* __restrict__ parent, int64_t rb, int64_t re,
                                    int64_t w0a, int64_t w0b, int64_t w1a, int64_t w1b,
                                    unsigned char* __restrict__ marks) {
  // marks[root]: bit 1 = contains an owned pore, bit 2 = touches a cut region
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t lo = w0a < rb ? w0a : rb, hi = w1b > re ? w1b : re;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    unsigned char m = 0;
    if (i >= rb && i < re) m |= 1;
    if ((i >= w0a && i < w0b) || (i >= w1a && i < w1b)) m |= 2;
    if (m) {
      unsigned char* dst = &marks[parent[i]];
      // byte-wide OR through the containing 32-bit word (integer atomic)
      unsigned int* word = reinterpret_cast<unsigned int*>(reinterpret_cast<size_t>(dst) & ~(size_t)3);
      atomicOr(word, (unsigned int)m << (8 * (reinterpret_cast<size_t>(dst) & 3)));
    }
  }
}

__global__ void k_count_interior_bad(const int* __restrict__ parent,
                                     const unsigned char* __restrict__ has_bc,
                                     const unsigned char* __restrict__ marks, int64_t lo,
                                     int64_t hi, unsigned long long* __restrict__ count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride)
    if (parent[i] == (int)i && marks[i] == 1 && !has_bc[i]) ++bad;
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_down_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(count, bad);
}

__global__ void k_window_export(const int* __restrict__ parent,
                                const unsigned char* __restrict__ has_bc, int64_t a, int64_t b,
                                int* __restrict__ lab, unsigned char* __restrict__ bc) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = a + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += stride) {
    const int r = parent[i];
    lab[i - a] = r;
    bc[i - a] = has_bc[r];
  }
}

extern "C" int b200_slab_clusters(b200_ctx* ctx, const int64_t* conns, int64_t Nt, int64_t Np,
                                  int64_t row_begin, int64_t row_end, int64_t halo,
                                  const double* bc_value, const double* bc_rate,
                                  int32_t* win_labels, uint8_t* win_has_bc,
                                  int64_t* interior_without_bc) {
  if (!ctx || row_begin < 0 || row_end > Np || row_begin >= row_end || halo < 0 || !win_labels ||
      !win_has_bc)
    return B200_ERR_ARG;
  RET(clusters(ctx, conns, Nt, Np, bc_value, bc_rate, nullptr, nullptr));
  const int* parent = ctx->w_cnt.as<int>();
  const unsigned char* has_bc = ctx->w_tmp1.as<unsigned char>();
  auto clip = [&](int64_t v) { return v < 0 ? 0 : (v > Np ? Np : v); };
  const int64_t w0a = clip(row_begin - halo), w0b = clip(row_begin + halo);
  const int64_t w1a = clip(row_end - halo), w1b = clip(row_end + halo);
  CK(ctx->w_tmp0.ensure((size_t)Np + 64));
  unsigned char* marks = ctx->w_tmp0.as<unsigned char>();
  CK(cudaMemsetAsync(marks, 0, (size_t)Np + 4, ctx->stream));
  unsigned long long* count = ctx->w_tmp2.as<unsigned long long>();
  CK(cudaMemsetAsync(count, 0, sizeof(unsigned long long), ctx->stream));
  const int B = 256;
  const int64_t lo = w0a < row_begin ? w0a : row_begin, hi = w1b > row_end ? w1b : row_end;
  LAUNCH(ctx, k_mark_owned_window, grid_for(hi - lo, B), B, 0, parent, row_begin, row_end, w0a,
         w0b, w1a, w1b, marks);
  LAUNCH(ctx, k_count_interior_bad, grid_for(Np, B), B, 0, parent, has_bc, marks, (int64_t)0, Np,
         count);
  // layout of the export: [w0a, w0b) then [w1a, w1b), each padded to 2*halo entries
  CK(cudaMemsetAsync(win_labels, 0xff, (size_t)(4 * halo) * sizeof(int), ctx->stream));
  CK(cudaMemsetAsync(win_has_bc, 0, (size_t)(4 * halo), ctx->stream));
  if (w0b > w0a)
    LAUNCH(ctx, k_window_export, grid_for(w0b - w0a, B), B, 0, parent, has_bc, w0a, w0b,
           win_labels + (w0a - (row_begin - halo)), win_has_bc + (w0a - (row_begin - halo)));
  if (w1b > w1a)
    LAUNCH(ctx, k_window_export, grid_for(w1b - w1a, B), B, 0, parent, has_bc, w1a, w1b,
           win_labels + 2 * halo + (w1a - (row_end - halo)),
           win_has_bc + 2 * halo + (w1a - (row_end - halo)));
  CK(cudaGetLastError());
  unsigned long long h = 0;
  CK(cudaMemcpyAsync(&h, count, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (interior_without_bc) *interior_without_bc = (int64_t)h;
  return B200_OK;
}

// ---------------------------------------------------- network health and trimming
// check_network_health (openpnm/utils/_health.py:40-98) with the pore-scale models it
// calls (openpnm/models/network/_health.py:11-76: cluster_size, isolated_pores,
// looped_throats) and topotools.trim (openpnm/topotools/_topotools.py:157-229), on the
// device: config 4 (a Delaunay network whose over-long hull throats were cut) drops its
// isolated pores and keeps the largest cluster before K1 assembles it.
__global__ void k_health_throats(const longlong2* __restrict__ conns, int64_t Nt,
                                 int* __restrict__ deg, unsigned char* __restrict__ tflags,
                                 unsigned long long* __restrict__ counts) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long looped = 0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < Nt; t += stride) {
    const longlong2 c = conns[t];
    atomicAdd(&deg[c.x], 1);
    atomicAdd(&deg[c.y], 1);
    const bool lp = c.x == c.y;
    if (tflags) tflags[t] = lp ? 1 : 0;
    if (lp) ++looped;
  }
  for (int o = 16; o > 0; o >>= 1) looped += __shfl_down_sync(0xffffffffu, looped, o);
  if ((threadIdx.x & 31) == 0 && looped) atomicAdd(&counts[3], looped);
}

__global__ void k_health_sizes(const int* __restrict__ parent, int64_t Np, int* __restrict__ size) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += stride)
    atomicAdd(&size[parent[i]], 1);
}

__global__ void k_health_max(const int* __restrict__ parent, const int* __restrict__ size,
                             int64_t Np, unsigned long long* __restrict__ counts) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long roots = 0, mx = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += stride)
    if (parent[i] == (int)i) {
      ++roots;
      if ((unsigned long long)size[i] > mx) mx = (unsigned long long)size[i];
    }
  for (int o = 16; o > 0; o >>= 1) {
    roots += __shfl_down_sync(0xffffffffu, roots, o);
    const unsigned long long other = __shfl_down_sync(0xffffffffu, mx, o);
    mx = other > mx ? other : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    if (roots) atomicAdd(&counts[0], roots);
    if (mx) atomicMax(&counts[5], mx);
  }
}

__global__ void k_health_pores(const int* __restrict__ parent, const int* __restrict__ size,
                               const int* __restrict__ deg, int64_t Np,
                               int* __restrict__ cluster_size, unsigned char* __restrict__ isolated,
                               unsigned long long* __restrict__ counts) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int mx = (int)counts[5];
  unsigned long long iso = 0, disc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Np; i += stride) {
    const int s = size[parent[i]];
    const bool is = deg[i] == 0;
    if (cluster_size) cluster_size[i] = s;
    if (isolated) isolated[i] = is ? 1 : 0;
    if (is) ++iso;
    if (s < mx) ++disc;
  }
  for (int o = 16; o > 0; o >>= 1) {
    iso += __shfl_down_sync(0xffffffffu, iso, o);
    disc += __shfl_down_sync(0xffffffffu, disc, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (iso) atomicAdd(&counts[1], iso);
    if (disc) atomicAdd(&counts[2], disc);
  }
}

extern "C" int b200_network_health(b200_ctx* ctx, const int64_t* conns, int64_t Nt, int64_t Np,
                                   int32_t* cluster_size, uint8_t* isolated, uint8_t* looped,
                                   int32_t* labels, int64_t* counts6_h) {
  if (!ctx || !counts6_h) return B200_ERR_ARG;
  // throats that point at non-existent pores ("headless") are an error here
  // (B200_ERR_INDEX), as they are for the assembly
  RET(clusters(ctx, conns, Nt, Np, nullptr, nullptr, nullptr, nullptr));
  const int* parent = ctx->w_cnt.as<int>();
  CK(ctx->w_keys.ensure((size_t)(2 * Np + 2) * sizeof(int)));
  int* size = ctx->w_keys.as<int>();
  int* deg = size + Np;
  CK(cudaMemsetAsync(size, 0, (size_t)(2 * Np) * sizeof(int), ctx->stream));
  CK(ctx->w_tmp2.ensure(256));
  unsigned long long* counts = ctx->w_tmp2.as<unsigned long long>();
  CK(cudaMemsetAsync(counts, 0, 8 * sizeof(unsigned long long), ctx->stream));
  const int B = 256;
  if (Nt > 0)
    LAUNCH(ctx, k_health_throats, grid_for(Nt, B), B, 0,
           reinterpret_cast<const longlong2*>(conns), Nt, deg, looped, counts);
  LAUNCH(ctx, k_health_sizes, grid_for(Np, B), B, 0, parent, Np, size);
  LAUNCH(ctx, k_health_max, grid_for(Np, B), B, 0, parent, size, Np, counts);
  LAUNCH(ctx, k_health_pores, grid_for(Np, B), B, 0, parent, size, deg, Np, cluster_size,
         isolated, counts);
  if (labels)
    CK(cudaMemcpyAsync(labels, parent, (size_t)Np * sizeof(int), cudaMemcpyDeviceToDevice,
                       ctx->stream));
  CK(cudaGetLastError());
  unsigned long long h[6];
  CK(cudaMemcpyAsync(h, counts, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 6; ++k) counts6_h[k] = (int64_t)h[k];
  return B200_OK;
}

__global__ void k_trim_flags(const longlong2* __restrict__ conns, int64_t Nt,
                             const unsigned char* __restrict__ pkeep,
                             const unsigned char* __restrict__ tkeep, int* __restrict__ tflag) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < Nt; t += stride) {
    const longlong2 c = conns[t];
    tflag[t] = ((!tkeep || tkeep[t]) && pkeep[c.x] && pkeep[c.y]) ? 1 : 0;
  }
}

__global__ void k_trim_u8_to_i32(const unsigned char* __restrict__ in, int64_t n,
                                 int* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = in[i] ? 1 : 0;
}

__global__ void k_trim_map(const int* __restrict__ flag, const int* __restrict__ pos, int64_t n,
                           long long* __restrict__ map) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    map[i] = flag[i] ? (long long)pos[i] : -1ll;
}

__global__ void k_trim_conns(const longlong2* __restrict__ conns, int64_t Nt,
                             const long long* __restrict__ pmap,
                             const long long* __restrict__ tmap, longlong2* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < Nt; t += stride) {
    const long long k = tmap[t];
    if (k < 0) continue;
    const longlong2 c = conns[t];
    out[k] = make_longlong2(pmap[c.x], pmap[c.y]);
  }
}

extern "C" int b200_trim(b200_ctx* ctx, const int64_t* conns, int64_t Nt, int64_t Np,
                         const uint8_t* pore_keep, const uint8_t* throat_keep, int64_t* pore_map,
                         int64_t* throat_map, int64_t* conns_out, int64_t* Np_new,
                         int64_t* Nt_new) {
  if (!ctx || !conns || !pore_keep || !pore_map || !throat_map || !conns_out || Np <= 0 ||
      Nt < 0 || Np >= (1ll << 31) || Nt >= (1ll << 31))
    return B200_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const int B = 256;
  const int64_t m = Np > Nt ? Np : Nt;
  CK(ctx->w_cnt.ensure((size_t)(m + 1) * sizeof(int)));
  CK(ctx->w_keys.ensure((size_t)(m + 2) * sizeof(int)));
  int* flag = ctx->w_cnt.as<int>();
  int* pos = ctx->w_keys.as<int>();
  int64_t np_new = 0, nt_new = 0;
  LAUNCH(ctx, k_trim_u8_to_i32, grid_for(Np, B), B, 0, pore_keep, Np, flag);
  RET(b200_scan_exclusive_i32(ctx, flag, pos, Np, &np_new));
  if (np_new == 0) FAIL(B200_ERR_ARG, "Cannot delete ALL pores");
  LAUNCH(ctx, k_trim_map, grid_for(Np, B), B, 0, flag, pos, Np,
         reinterpret_cast<long long*>(pore_map));
  if (Nt > 0) {
    LAUNCH(ctx, k_trim_flags, grid_for(Nt, B), B, 0, reinterpret_cast<const longlong2*>(conns),
           Nt, pore_keep, throat_keep, flag);
    RET(b200_scan_exclusive_i32(ctx, flag, pos, Nt, &nt_new));
    LAUNCH(ctx, k_trim_map, grid_for(Nt, B), B, 0, flag, pos, Nt,
           reinterpret_cast<long long*>(throat_map));
    LAUNCH(ctx, k_trim_conns, grid_for(Nt, B), B, 0, reinterpret_cast<const longlong2*>(conns),
           Nt, reinterpret_cast<const long long*>(pore_map),
           reinterpret_cast<const long long*>(throat_map),
           reinterpret_cast<longlong2*>(conns_out));
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  if (Np_new) *Np_new = np_new;
  if (Nt_new) *Nt_new = nt_new;
  return B200_OK;
}
