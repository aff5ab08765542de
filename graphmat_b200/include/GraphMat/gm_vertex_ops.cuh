// gm_vertex_ops.cuh -- device-side Graph::applyToAllVertices / applyReduceAllVertices / applyToAllEdges
// for callers that pass a GM_HD FUNCTOR instead of the reference's host function pointer.
//
// Replaces, on the device,
//   Apply       include/GMDP/singlenode/apply.h      (via Graph.h:371-375)
//   MapReduce   include/GMDP/singlenode/reduce.h, multinode/reduce.h:55-72  (via Graph.h:377-381)
//   ApplyEdges  include/GMDP/singlenode/applyedges.h:38-95, multinode/applyedges.h:45-161 (via Graph.h:389-402)
// of narayanan2004/GraphMat.  The reference's signatures take C function pointers that only exist on the host;
// those overloads keep the host path of Graph.h.  A functor travels to the kernels by value (its `param`
// becomes a member), nothing crosses PCIe but the reduced value.
//   map functor      void operator()(const V& in, V* out) const
//   reduce pair      void map(V* v, T* out) const   +   void reduce(const T& a, const T& b, T* c) const
//                    (reduce must be associative and commutative: the device folds a tree, not the vertex order)
//   edge functor     void operator()(E* edge, const V& src, const V& dst) const
#ifndef GRAPHMAT_B200_VERTEX_OPS_CUH
#define GRAPHMAT_B200_VERTEX_OPS_CUH
#include <cuda_runtime.h>

#include "graphmat_b200.h"

namespace gm {

template <class V, class F>
__global__ void __launch_bounds__(256) k_map_vertices(V* __restrict__ vp, int n, F f) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V in = vp[i];
  V out = in;
  f(in, &out);
  vp[i] = out;
}

// block tree over 256 mapped values; partial[b] per block, valid[b] = the block saw at least one vertex
template <class V, class T, class M, class R>
__global__ void __launch_bounds__(256) k_map_reduce(V* __restrict__ vp, int n, M map, R reduce, T* __restrict__ partial) {
  __shared__ __align__(16) unsigned char sm_raw[256 * sizeof(T)];  // raw storage: T may have a constructor
  T* sm = reinterpret_cast<T*>(sm_raw);
  __shared__ unsigned char have[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  have[threadIdx.x] = i < n;
  if (i < n) map(vp + i, &sm[threadIdx.x]);
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d && have[threadIdx.x + d]) {
      if (have[threadIdx.x]) {
        T c;
        reduce(sm[threadIdx.x], sm[threadIdx.x + d], &c);
        sm[threadIdx.x] = c;
      } else {
        sm[threadIdx.x] = sm[threadIdx.x + d];
        have[threadIdx.x] = 1;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];  // every launched block holds at least one vertex
}
template <class T, class R>
__global__ void __launch_bounds__(256) k_reduce_partials(const T* __restrict__ in, int n, R reduce, T* __restrict__ out) {
  __shared__ __align__(16) unsigned char sm_raw[256 * sizeof(T)];
  T* sm = reinterpret_cast<T*>(sm_raw);
  __shared__ unsigned char have[256];
  T acc;
  bool h = false;
  for (int i = threadIdx.x; i < n; i += 256) {
    if (h) {
      T c;
      reduce(acc, in[i], &c);
      acc = c;
    } else {
      acc = in[i];
      h = true;
    }
  }
  have[threadIdx.x] = h;
  if (h) sm[threadIdx.x] = acc;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (threadIdx.x < d && have[threadIdx.x + d]) {
      if (have[threadIdx.x]) {
        T c;
        reduce(sm[threadIdx.x], sm[threadIdx.x + d], &c);
        sm[threadIdx.x] = c;
      } else {
        sm[threadIdx.x] = sm[threadIdx.x + d];
        have[threadIdx.x] = 1;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sm[0];
}

// every stored entry of one operand matrix: ROW_IS_DST (AT: rows = edge destinations) or not (A: rows = sources)
template <class V, class E, class F, bool ROW_IS_DST>
__global__ void __launch_bounds__(256) k_apply_edges_sell(gm_matrix_view M, const V* __restrict__ vp, F f) {
  const int lane = threadIdx.x & 31;
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= M.n_slices) return;
  const int slot = M.n_heavy + s * 32 + lane;
  const int len = M.row_len[slot];
  if (len == 0) return;
  const int rv = M.identity ? slot : M.slot_vertex[slot];
  const V rowp = vp[rv];
  E* vals = reinterpret_cast<E*>(const_cast<void*>(M.s_val));
  const long long base = M.slice_ptr[s] + lane;
  for (int i = 0; i < len; i++) {
    const long long pos = base + 32ll * i;
    const V colp = vp[M.s_col[pos]];  // one rank: an x index is a local vertex
    if (ROW_IS_DST) f(vals + pos, colp, rowp);
    else f(vals + pos, rowp, colp);
  }
}
template <class V, class E, class F, bool ROW_IS_DST>
__global__ void __launch_bounds__(256) k_apply_edges_heavy(gm_matrix_view M, const V* __restrict__ vp, F f) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slot >= M.n_heavy) return;
  const long long beg = M.h_ptr[slot], end = M.h_ptr[slot + 1];
  if (beg == end) return;
  const int rv = M.identity ? slot : M.slot_vertex[slot];
  const V rowp = vp[rv];
  E* vals = reinterpret_cast<E*>(const_cast<void*>(M.h_val));
  for (long long k = beg + lane; k < end; k += 32) {
    const V colp = vp[M.h_col[k]];
    if (ROW_IS_DST) f(vals + k, colp, rowp);
    else f(vals + k, rowp, colp);
  }
}

template <class V, class F>
int map_vertices(gm_graph* g, F f) {
  gm_graph_view gv;
  if (gm_graph_view_get(g, &gv)) return 1;
  if (gv.n_local > 0)
    k_map_vertices<V, F><<<(gv.n_local + 255) / 256, 256, 0, (cudaStream_t)gv.stream>>>((V*)gv.vertexproperty, gv.n_local, f);
  return cudaGetLastError() != cudaSuccess || gm_graph_synchronize(g);
}

template <class V, class T, class M, class R>
int map_reduce_vertices(gm_graph* g, T* val, M map, R reduce) {
  gm_graph_view gv;
  if (gm_graph_view_get(g, &gv)) return 1;
  if (gv.world != 1) return 2;  // partial results of the other ranks are the host language's business
  if (gv.n_local <= 0) return 0;
  const int blocks = (gv.n_local + 255) / 256;
  T* d = nullptr;
  if (cudaMalloc((void**)&d, ((size_t)blocks + 1) * sizeof(T)) != cudaSuccess) return 1;
  cudaStream_t st = (cudaStream_t)gv.stream;
  k_map_reduce<V, T, M, R><<<blocks, 256, 0, st>>>((V*)gv.vertexproperty, gv.n_local, map, reduce, d + 1);
  k_reduce_partials<T, R><<<1, 256, 0, st>>>(d + 1, blocks, reduce, d);
  int rc = cudaGetLastError() != cudaSuccess;
  rc |= cudaMemcpyAsync(val, d, sizeof(T), cudaMemcpyDeviceToHost, st) != cudaSuccess;
  rc |= cudaStreamSynchronize(st) != cudaSuccess;
  cudaFree(d);
  return rc;
}

template <class V, class E, class F>
int apply_edges(gm_graph* g, F f) {
  gm_graph_view gv;
  if (gm_graph_view_get(g, &gv)) return 1;
  if (gv.world != 1) return 2;
  cudaStream_t st = (cudaStream_t)gv.stream;
  const V* vp = (const V*)gv.vertexproperty;
  if (gv.AT.n_slots > 0) {  // rows = destinations, columns = sources
    if (gv.AT.n_slices > 0) k_apply_edges_sell<V, E, F, true><<<(gv.AT.n_slices + 7) / 8, 256, 0, st>>>(gv.AT, vp, f);
    if (gv.AT.n_heavy > 0) k_apply_edges_heavy<V, E, F, true><<<(gv.AT.n_heavy + 7) / 8, 256, 0, st>>>(gv.AT, vp, f);
  }
  if (gv.A.n_slots > 0) {  // rows = sources, columns = destinations
    if (gv.A.n_slices > 0) k_apply_edges_sell<V, E, F, false><<<(gv.A.n_slices + 7) / 8, 256, 0, st>>>(gv.A, vp, f);
    if (gv.A.n_heavy > 0) k_apply_edges_heavy<V, E, F, false><<<(gv.A.n_heavy + 7) / 8, 256, 0, st>>>(gv.A, vp, f);
  }
  if (cudaGetLastError() != cudaSuccess) return 1;
  return gm_graph_edges_changed(g);  // the column-major companion of the sparse-frontier path holds edge values too
}

}  // namespace gm
#endif
