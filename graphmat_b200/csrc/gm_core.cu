// gm_core.cu -- graph container + device-side matrix construction behind the C ABI
// (include/graphmat_b200.h).
//
// Replaces the reference's Graph::ReadEdgelist (include/Graph.h:210-246),
// SpMat::ingestEdgelist + DCSCTile ctor + Transpose (include/GMDP/matrices/SpMat.h:97-278,
// 422-443, DCSCTile.h:241-381) and the Graph accessors (include/Graph.h:263-364) of
// narayanan2004/GraphMat -- NOT by translating them: everything is built on the device
// with radix sorts, into the row-sorted CSR + sliced-ELL layout the kernels in
// gm_engine.cuh stream ("Data layout in HBM").
//
// What is kept from the reference is the LOGICAL layout that fixes results:
//   * public id -> native id permutation (Graph.h:111-130) with ref_threads
//   * per row, entries in ascending native column id (the fold order of
//     GMDP/singlenode/spmspv.h:55-77)
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gm_internal.h"

static thread_local std::string g_err;
void gm_set_error(const std::string& s) { g_err = s; }
extern "C" const char* gm_last_error(void) { return g_err.c_str(); }
extern "C" int gm_abi_struct_sizes(int out[6]) {
  out[0] = (int)sizeof(gm_graph_opts);
  out[1] = (int)sizeof(gm_matrix_view);
  out[2] = (int)sizeof(gm_graph_view);
  out[3] = (int)sizeof(gm_vectors_view);
  out[4] = (int)sizeof(gm_run_stats);
  out[5] = (int)sizeof(gm_push_plan);
  return 0;
}

#define CK(call)                                                                             \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      char b_[512];                                                                          \
      snprintf(b_, sizeof b_, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      gm_set_error(b_);                                                                      \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

template <class T>
static int dalloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  CK(cudaMalloc((void**)p, n * sizeof(T)));
  return 0;
}

// --------------------------------------------------------------- id mapping --
// include/Graph.h:111-130 with nsegments = 1 (parity target is the 1-rank reference)
__host__ __device__ static inline int to_native0(int pub1, int n, int npart) {
  int v = pub1 - 1;
  int height = n / npart;
  int vmax = height * npart;
  if (v >= vmax) return v;
  int col = v % npart;
  int row = v / npart;
  return row + col * height;
}

__global__ void k_to_native(int* ids, long long nnz, int n, int npart) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nnz) ids[e] = to_native0(ids[e], n, npart);
}

// public ids must lie in [1, n] (edgelist.h trusts the file header; a bad id would index out of bounds below)
__global__ void k_check_ids(const int* a, const int* b, long long nnz, int n, int* bad) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  bool x = e < nnz && (a[e] < 1 || a[e] > n || b[e] < 1 || b[e] > n);
  if (__ballot_sync(0xffffffffu, x) && (threadIdx.x & 31) == 0) atomicExch(bad, 1);
}
static int check_ids(const int* d_src, const int* d_dst, long long nnz, int n, cudaStream_t st, const char* who);

__global__ void k_hist(const int* ids, long long nnz, int* cnt) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nnz) atomicAdd(cnt + ids[e], 1);
}

__global__ void k_vertex_keys(const int* deg, int n, unsigned long long* keys) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) keys[v] = ((unsigned long long)(0xffffffffu - (unsigned)deg[v]) << 32) | (unsigned)v;
}

// placement p (hot first) -> owner p % world, local p / world, x index owner * n_pad + local
__global__ void k_place(const unsigned long long* keys, int n, int world, int n_pad, int* xidx_of_native) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) xidx_of_native[(unsigned)(keys[p] & 0xffffffffu)] = (p % world) * n_pad + p / world;
}

__global__ void k_count_owned(const int* rows, long long nnz, const int* xidx, int n_pad, int rank, int* len_local) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int xr = xidx[rows[e]];
  if (xr / n_pad == rank) atomicAdd(len_local + (xr % n_pad), 1);
}

__global__ void k_slot_from_keys(const unsigned long long* keys, int n_pad, const int* len_local, int* slot_vertex,
                                 int* slot_of_local, int* row_len) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_pad) return;
  int local = (int)(keys[s] & 0xffffffffu);
  slot_vertex[s] = local;
  slot_of_local[local] = s;
  row_len[s] = len_local[local];
}

__global__ void k_edge_keys(const int* rows, const int* cols, long long nnz, const int* xidx, int n_pad, int rank,
                            const int* slot_of_local, unsigned long long* keys, unsigned* payload) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int xr = xidx[rows[e]];
  unsigned long long k = ~0ull;
  if (xr / n_pad == rank) {
    int local = xr % n_pad;
    int slot = slot_of_local ? slot_of_local[local] : local;
    k = ((unsigned long long)(unsigned)slot << 32) | (unsigned)cols[e];
  }
  keys[e] = k;
  payload[e] = (unsigned)e;
}

__global__ void k_len_to_ll(const int* len, int n, long long* out, int mul) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (long long)len[i] * mul;
}

__global__ void k_count_heavy(const int* row_len, int n, int thr, int* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool h = i < n && row_len[i] > thr;
  unsigned m = __ballot_sync(0xffffffffu, h);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, __popc(m));
}

__global__ void k_count_nonzero(const int* row_len, int n, int* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool h = i < n && row_len[i] > 0;
  unsigned m = __ballot_sync(0xffffffffu, h);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, __popc(m));
}

__global__ void k_slice_width(const int* row_len, int n_heavy, int n_slices, long long* out) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_slices) out[s] = 32ll * row_len[n_heavy + s * 32];
}

__global__ void k_seg_count(const long long* h_ptr, int n_heavy, int seg_len, int* cnt) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_heavy) cnt[r] = (int)((h_ptr[r + 1] - h_ptr[r] + seg_len - 1) / seg_len);
}
__global__ void k_seg_rows(const int* seg_ptr, int n_heavy, int* seg_row) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_heavy) return;
  for (int k = seg_ptr[r]; k < seg_ptr[r + 1]; k++) seg_row[k] = r;
}

template <class E>
__global__ void k_fill_heavy(const unsigned long long* keys, const unsigned* payload, long long cnt, const int* xidx,
                             const E* val, int* h_col, E* h_val) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  h_col[i] = xidx[(unsigned)(keys[i] & 0xffffffffu)];
  h_val[i] = val ? val[payload[i]] : E(1);
}

template <class E>
__global__ void k_fill_sell(const unsigned long long* keys, const unsigned* payload, long long first, long long cnt,
                            const int* xidx, const E* val, const long long* row_ptr, const long long* slice_ptr,
                            int n_heavy, int* s_col, E* s_val) {
  long long i = first + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= cnt) return;
  unsigned long long k = keys[i];
  int slot = (int)(k >> 32);
  long long j = i - row_ptr[slot];
  int rel = slot - n_heavy;
  long long pos = slice_ptr[rel >> 5] + j * 32 + (rel & 31);
  s_col[pos] = xidx[(unsigned)(k & 0xffffffffu)];
  s_val[pos] = val ? val[payload[i]] : E(1);
}

// ------------------------------------------------------------ RMAT generator --
__host__ __device__ static inline unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// Graph500 quadrant probabilities a,b,c,d = .57,.19,.19,.05 as 32-bit thresholds
#define GM_RMAT_TA 2448131358u  /* floor(0.57 * 2^32) */
#define GM_RMAT_TAB 3264175145u /* floor(0.76 * 2^32) */
#define GM_RMAT_TABC 4080218931u /* floor(0.95 * 2^32) */
__host__ __device__ static inline void rmat_edge(int scale, unsigned long long seed, unsigned long long e, int* s,
                                                 int* d) {
  unsigned si = 0, di = 0;
  unsigned long long h = 0;
  for (int l = 0; l < scale; l++) {
    if ((l & 1) == 0) h = splitmix64(seed * 0x100000001B3ull + e * 32ull + (unsigned long long)(l >> 1));
    unsigned r = (l & 1) ? (unsigned)(h >> 32) : (unsigned)h;
    unsigned sb = r >= GM_RMAT_TAB;                                  // quadrants c, d
    unsigned db = (r >= GM_RMAT_TA && r < GM_RMAT_TAB) || r >= GM_RMAT_TABC;  // quadrants b, d
    si = (si << 1) | sb;
    di = (di << 1) | db;
  }
  *s = (int)si + 1;
  *d = (int)di + 1;
}
__host__ __device__ static inline int rmat_weight(unsigned long long wseed, unsigned long long e, int weight_max) {
  if (weight_max <= 0) return 1;
  return 1 + (int)(splitmix64(wseed * 0x9E3779B1ull + e + 0x5555555555ull) % (unsigned long long)weight_max);
}
__global__ void k_rmat(int scale, unsigned long long seed, int weight_max, unsigned long long wseed, long long nnz,
                       int* src, int* dst, int* val) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int s, d;
  rmat_edge(scale, seed, (unsigned long long)e, &s, &d);
  src[e] = s;
  dst[e] = d;
  if (val) val[e] = rmat_weight(wseed, (unsigned long long)e, weight_max);
}

extern "C" int gm_rmat_edges_host(int scale, int edge_factor, unsigned long long seed, int weight_max,
                                  unsigned long long weight_seed, int* src, int* dst, int* val) {
  long long nnz = (long long)edge_factor << scale;
#pragma omp parallel for
  for (long long e = 0; e < nnz; e++) {
    rmat_edge(scale, seed, (unsigned long long)e, &src[e], &dst[e]);
    if (val) val[e] = rmat_weight(weight_seed, (unsigned long long)e, weight_max);
  }
  return 0;
}

// ------------------------------------------------------------- device utils --
extern "C" int gm_set_device(int device) {
  CK(cudaSetDevice(device));
  return 0;
}
extern "C" int gm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// blocking device->host read ordered on the graph's (non-blocking) stream
static int d2h(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

static inline unsigned nblk(long long n, int b = 256) { return (unsigned)((n + b - 1) / b); }

static int ceil_log2(long long v) {
  int b = 0;
  while ((1ll << b) < v) b++;
  return b;
}

template <class K, class V>
static int sort_pairs(K* k0, K* k1, V* v0, V* v1, long long n, int end_bit, cudaStream_t st, K** kout, V** vout) {
  cub::DoubleBuffer<K> kb(k0, k1);
  cub::DoubleBuffer<V> vb(v0, v1);
  size_t tb = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, kb, vb, n, 0, end_bit, st));
  void* tmp = nullptr;
  CK(cudaMalloc(&tmp, tb ? tb : 1));
  CK(cub::DeviceRadixSort::SortPairs(tmp, tb, kb, vb, n, 0, end_bit, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaFree(tmp));
  *kout = kb.Current();
  *vout = vb.Current();
  return 0;
}

template <class K>
static int sort_keys(K* k0, K* k1, long long n, int end_bit, cudaStream_t st, K** kout) {
  cub::DoubleBuffer<K> kb(k0, k1);
  size_t tb = 0;
  CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, kb, n, 0, end_bit, st));
  void* tmp = nullptr;
  CK(cudaMalloc(&tmp, tb ? tb : 1));
  CK(cub::DeviceRadixSort::SortKeys(tmp, tb, kb, n, 0, end_bit, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaFree(tmp));
  *kout = kb.Current();
  return 0;
}

static int exclusive_scan_ll(const long long* in, long long* out, int n, cudaStream_t st) {
  // out has n + 1 entries; out[n] = total
  size_t tb = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n + 1, st));
  void* tmp = nullptr;
  CK(cudaMalloc(&tmp, tb ? tb : 1));
  CK(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n + 1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaFree(tmp));
  return 0;
}

static int check_ids(const int* d_src, const int* d_dst, long long nnz, int n, cudaStream_t st, const char* who) {
  if (nnz == 0) return 0;
  int* bad = nullptr;
  if (dalloc(&bad, 1)) return 1;
  CK(cudaMemsetAsync(bad, 0, 4, st));
  k_check_ids<<<nblk(nnz), 256, 0, st>>>(d_src, d_dst, nnz, n, bad);
  int h = 0;
  int rc = d2h(&h, bad, 4, st);
  cudaFree(bad);
  if (rc) return 1;
  if (h) {
    gm_set_error(std::string(who) + ": an edge endpoint is outside [1, nvertices] (ids are public and 1-based)");
    return 1;
  }
  return 0;
}

// ------------------------------------------------------------- matrix build --
static void matrix_free(gm_matrix& M) {
  cudaFree(M.c_ptr);
  cudaFree(M.c_row);
  cudaFree(M.c_rank);
  cudaFree(M.c_val);
  cudaFree(M.big_cols);
  cudaFree(M.slot_vertex);
  cudaFree(M.row_len);
  cudaFree(M.h_ptr);
  cudaFree(M.h_col);
  cudaFree(M.h_val);
  cudaFree(M.slice_ptr);
  cudaFree(M.s_col);
  cudaFree(M.s_val);
  cudaFree(M.seg_ptr);
  cudaFree(M.seg_row);
  M = gm_matrix();
}

// rows/cols: native 0-based ids per edge (device); val: edge values or NULL (all ones)
template <class E>
static int build_matrix(gm_graph* g, gm_matrix& M, const int* rows, const int* cols, const E* val, long long nnz,
                        bool identity) {
  cudaStream_t st = g->stream;
  const int n_pad = g->n_pad;
  int* len_local = nullptr;
  if (dalloc(&len_local, n_pad)) return 1;
  CK(cudaMemsetAsync(len_local, 0, (size_t)n_pad * 4, st));
  if (nnz) k_count_owned<<<nblk(nnz), 256, 0, st>>>(rows, nnz, g->d_xidx, n_pad, g->rank, len_local);

  int* slot_of_local = nullptr;
  if (dalloc(&M.row_len, n_pad)) return 1;
  if (identity) {
    CK(cudaMemcpyAsync(M.row_len, len_local, (size_t)n_pad * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    unsigned long long *k0, *k1, *ks;
    if (dalloc(&k0, n_pad) || dalloc(&k1, n_pad)) return 1;
    k_vertex_keys<<<nblk(n_pad), 256, 0, st>>>(len_local, n_pad, k0);
    if (sort_keys(k0, k1, n_pad, 64, st, &ks)) return 1;
    if (dalloc(&M.slot_vertex, n_pad) || dalloc(&slot_of_local, n_pad)) return 1;
    k_slot_from_keys<<<nblk(n_pad), 256, 0, st>>>(ks, n_pad, len_local, M.slot_vertex, slot_of_local, M.row_len);
    CK(cudaStreamSynchronize(st));
    cudaFree(k0);
    cudaFree(k1);
  }

  // row_ptr by slot
  long long *len_ll = nullptr, *row_ptr = nullptr;
  if (dalloc(&len_ll, (size_t)n_pad + 1) || dalloc(&row_ptr, (size_t)n_pad + 1)) return 1;
  CK(cudaMemsetAsync(len_ll, 0, ((size_t)n_pad + 1) * 8, st));
  k_len_to_ll<<<nblk(n_pad), 256, 0, st>>>(M.row_len, n_pad, len_ll, 1);
  if (exclusive_scan_ll(len_ll, row_ptr, n_pad, st)) return 1;
  long long owned = 0;
  if (d2h(&owned, row_ptr + n_pad, 8, st)) return 1;
  M.nnz = owned;

  // heavy prefix / non-empty prefix
  int* cnt = nullptr;
  if (dalloc(&cnt, 4)) return 1;
  CK(cudaMemsetAsync(cnt, 0, 16, st));
  k_count_heavy<<<nblk(n_pad), 256, 0, st>>>(M.row_len, n_pad, 31, cnt + 3);  // rows with >= 32 entries
  k_count_heavy<<<nblk(n_pad), 256, 0, st>>>(M.row_len, n_pad, g->heavy_threshold, cnt);
  k_count_nonzero<<<nblk(n_pad), 256, 0, st>>>(M.row_len, n_pad, cnt + 1);
  k_count_heavy<<<nblk(n_pad), 256, 0, st>>>(M.row_len, n_pad, std::max(g->coop_threshold, g->heavy_threshold), cnt + 2);
  int hc[4];
  if (d2h(hc, cnt, 16, st)) return 1;
  cudaFree(cnt);
  int n_heavy = std::min(n_pad, (hc[0] + 31) / 32 * 32);
  int n_nonzero = hc[1];
  int n_slices = n_nonzero > n_heavy ? (n_nonzero - n_heavy + 31) / 32 : 0;
  M.n_slots = n_pad;
  M.n_heavy = n_heavy;
  M.n_slices = n_slices;
  M.n_coop = std::min(hc[2], n_heavy);
  {
    int* cl = nullptr;
    if (dalloc(&cl, 1)) return 1;
    CK(cudaMemsetAsync(cl, 0, 4, st));
    k_count_heavy<<<nblk(n_pad), 256, 0, st>>>(M.row_len, n_pad, std::max(g->long_threshold, g->coop_threshold), cl);
    int nl = 0;
    if (d2h(&nl, cl, 4, st)) return 1;
    cudaFree(cl);
    M.n_long = std::min(nl, M.n_coop);
  }
  M.n_slices_wide = hc[3] > n_heavy ? std::min(n_slices, (hc[3] - n_heavy + 31) / 32) : 0;
  M.identity = identity ? 1 : 0;

  // sort owned edges by (slot, native column)
  unsigned long long *k0 = nullptr, *k1 = nullptr, *ks = nullptr;
  unsigned *p0 = nullptr, *p1 = nullptr, *ps = nullptr;
  if (dalloc(&k0, nnz) || dalloc(&k1, nnz) || dalloc(&p0, nnz) || dalloc(&p1, nnz)) return 1;
  if (nnz) {
    k_edge_keys<<<nblk(nnz), 256, 0, st>>>(rows, cols, nnz, g->d_xidx, n_pad, g->rank, slot_of_local, k0, p0);
    // the unowned sentinel ~0 needs all 64 bits only when some edge is unowned
    int end_bit = (g->world > 1) ? 64 : 32 + std::max(1, ceil_log2(n_pad));
    if (sort_pairs(k0, k1, p0, p1, nnz, end_bit, st, &ks, &ps)) return 1;
  }

  // heavy rows: the sorted prefix is already row-contiguous
  long long nh = 0;
  if (dalloc(&M.h_ptr, (size_t)n_heavy + 1)) return 1;
  CK(cudaMemcpyAsync(M.h_ptr, row_ptr, ((size_t)n_heavy + 1) * 8, cudaMemcpyDeviceToDevice, st));
  if (d2h(&nh, row_ptr + n_heavy, 8, st)) return 1;
  M.long_entries = 0;
  if (M.n_long > 0 && d2h(&M.long_entries, row_ptr + M.n_long, 8, st)) return 1;
  E* hv = nullptr;
  if (dalloc(&M.h_col, nh + 16) || dalloc(&hv, nh + 16)) return 1;  // +16: kernels read aligned groups of 8
  CK(cudaMemsetAsync(M.h_col + nh, 0, 16 * 4, st));
  CK(cudaMemsetAsync(hv + nh, 0, 16 * sizeof(E), st));
  M.h_val = hv;
  if (nh) k_fill_heavy<E><<<nblk(nh), 256, 0, st>>>(ks, ps, nh, g->d_xidx, val, M.h_col, hv);

  // segments of the heavy rows (two-phase fold for associative programs)
  M.seg_len = GM_SEG_LEN;
  M.n_segs = 0;
  if (dalloc(&M.seg_ptr, (size_t)n_heavy + 1)) return 1;
  CK(cudaMemsetAsync(M.seg_ptr, 0, ((size_t)n_heavy + 1) * 4, st));
  if (n_heavy > 0) {
    int* segcnt = nullptr;
    if (dalloc(&segcnt, (size_t)n_heavy + 1)) return 1;
    CK(cudaMemsetAsync(segcnt, 0, ((size_t)n_heavy + 1) * 4, st));
    k_seg_count<<<nblk(n_heavy), 256, 0, st>>>(M.h_ptr, n_heavy, M.seg_len, segcnt);
    size_t tb = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, segcnt, M.seg_ptr, n_heavy + 1, st));
    void* tmp = nullptr;
    CK(cudaMalloc(&tmp, tb ? tb : 1));
    CK(cub::DeviceScan::ExclusiveSum(tmp, tb, segcnt, M.seg_ptr, n_heavy + 1, st));
    if (d2h(&M.n_segs, M.seg_ptr + n_heavy, 4, st)) return 1;
    cudaFree(tmp);
    cudaFree(segcnt);
  }
  if (dalloc(&M.seg_row, (size_t)M.n_segs)) return 1;
  if (M.n_segs) k_seg_rows<<<nblk(n_heavy), 256, 0, st>>>(M.seg_ptr, n_heavy, M.seg_row);

  // sliced ELL for the rest
  long long* widths = nullptr;
  if (dalloc(&widths, (size_t)n_slices + 1) || dalloc(&M.slice_ptr, (size_t)n_slices + 1)) return 1;
  CK(cudaMemsetAsync(widths, 0, ((size_t)n_slices + 1) * 8, st));
  if (n_slices) k_slice_width<<<nblk(n_slices), 256, 0, st>>>(M.row_len, n_heavy, n_slices, widths);
  if (exclusive_scan_ll(widths, M.slice_ptr, n_slices, st)) return 1;
  long long total = 0;
  if (d2h(&total, M.slice_ptr + n_slices, 8, st)) return 1;
  E* sv = nullptr;
  if (dalloc(&M.s_col, total) || dalloc(&sv, total)) return 1;
  M.s_val = sv;
  M.s_entries = total;
  CK(cudaMemsetAsync(M.s_col, 0, (size_t)total * 4, st));
  CK(cudaMemsetAsync(sv, 0, (size_t)total * sizeof(E), st));
  if (owned > nh)
    k_fill_sell<E><<<nblk(owned - nh), 256, 0, st>>>(ks, ps, nh, owned, g->d_xidx, val, row_ptr, M.slice_ptr, n_heavy,
                                                    M.s_col, sv);
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  cudaFree(widths);
  cudaFree(k0);
  cudaFree(k1);
  cudaFree(p0);
  cudaFree(p1);
  cudaFree(len_ll);
  cudaFree(row_ptr);
  cudaFree(len_local);
  cudaFree(slot_of_local);
  return 0;
}

static void fill_view(const gm_matrix& M, gm_matrix_view* v) {
  v->n_slots = M.n_slots;
  v->n_heavy = M.n_heavy;
  v->n_slices = M.n_slices;
  v->identity = M.identity;
  v->n_coop = M.n_coop;
  v->n_slices_wide = M.n_slices_wide;
  v->slot_vertex = M.slot_vertex;
  v->row_len = M.row_len;
  v->h_ptr = M.h_ptr;
  v->h_col = M.h_col;
  v->h_val = M.h_val;
  v->slice_ptr = M.slice_ptr;
  v->s_col = M.s_col;
  v->s_val = M.s_val;
  v->nnz = M.nnz;
  v->n_segs = M.n_segs;
  v->seg_len = M.seg_len;
  v->seg_ptr = M.seg_ptr;
  v->seg_row = M.seg_row;
  v->c_ptr = M.c_ptr;
  v->c_row = M.c_row;
  v->c_rank = M.c_rank;
  v->c_val = M.c_val;
  v->rank_bits = M.rank_bits;
  v->n_big_cols = M.n_big_cols;
  v->big_cols = M.big_cols;
  v->n_long = M.n_long;
  v->long_entries = M.long_entries;
}

// d_src/d_dst: device copies owned by this call (public ids, overwritten with native ids)
template <class E>
static int build_graph(gm_graph* g, int* d_src, int* d_dst, const E* d_val, long long nnz, const gm_graph* like,
                       int build_mask) {
  cudaStream_t st = g->stream;
  const int n = g->n;
  const int npart = g->ref_threads * 16;
  if (g->heavy_auto) {
    // A sliced-ELL lane folds its row as one serial chain (~1.5 us per 16 entries): keep the longest
    // chain well under the time the whole pass needs (~nnz / 3e11 s), i.e. threshold ~ nnz * 1e-5.
    long long per_rank = nnz / g->world;
    int t = 256;
    while (t < GM_DEFAULT_HEAVY_THRESHOLD && (long long)t * 2 <= per_rank / 100000) t *= 2;
    g->heavy_threshold = t;
  }
  // first public vertex with an out-edge (the benchmark's BFS/SSSP source)
  g->first_source = 0;
  if (check_ids(d_src, d_dst, nnz, n, st, "gm_graph_create")) return 1;
  if (nnz) {
    int* dmin;
    if (dalloc(&dmin, 1)) return 1;
    size_t tb = 0;
    CK(cub::DeviceReduce::Min(nullptr, tb, d_src, dmin, nnz, st));
    void* tmp;
    CK(cudaMalloc(&tmp, tb ? tb : 1));
    CK(cub::DeviceReduce::Min(tmp, tb, d_src, dmin, nnz, st));
    if (d2h(&g->first_source, dmin, 4, st)) return 1;
    cudaFree(tmp);
    cudaFree(dmin);
    k_to_native<<<nblk(nnz), 256, 0, st>>>(d_src, nnz, n, npart);
    k_to_native<<<nblk(nnz), 256, 0, st>>>(d_dst, nnz, n, npart);
  }
  if (dalloc(&g->d_xidx, n)) return 1;
  if (like) {
    if (like->n != n || like->world != g->world || like->ref_threads != g->ref_threads) {
      gm_set_error("order_like graph has a different shape");
      return 1;
    }
    CK(cudaMemcpyAsync(g->d_xidx, like->d_xidx, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    // placement: vertices by decreasing in-degree (= AT row length), ties by native id
    int* indeg;
    unsigned long long *k0, *k1, *ks;
    if (dalloc(&indeg, n) || dalloc(&k0, n) || dalloc(&k1, n)) return 1;
    CK(cudaMemsetAsync(indeg, 0, (size_t)n * 4, st));
    if (nnz) k_hist<<<nblk(nnz), 256, 0, st>>>(d_dst, nnz, indeg);
    k_vertex_keys<<<nblk(n), 256, 0, st>>>(indeg, n, k0);
    if (sort_keys(k0, k1, n, 64, st, &ks)) return 1;
    k_place<<<nblk(n), 256, 0, st>>>(ks, n, g->world, g->n_pad, g->d_xidx);
    CK(cudaStreamSynchronize(st));
    cudaFree(indeg);
    cudaFree(k0);
    cudaFree(k1);
  }
  if (build_mask & 2)
    if (build_matrix<E>(g, g->AT, d_dst, d_src, d_val, nnz, like == nullptr)) return 1;
  if (build_mask & 1)
    if (build_matrix<E>(g, g->A, d_src, d_dst, d_val, nnz, false)) return 1;
  return 0;
}

static int graph_common_alloc(gm_graph* g) {
  cudaStream_t st = g->stream;
  CK(cudaMalloc(&g->vp, (size_t)g->n_pad * g->sizeof_V));
  CK(cudaMemsetAsync(g->vp, 0, (size_t)g->n_pad * g->sizeof_V, st));
  g->vp_owner = true;
  if (dalloc(&g->active, (size_t)(g->n_pad >> 5))) return 1;
  CK(cudaMemsetAsync(g->active, 0, (size_t)(g->n_pad >> 5) * 4, st));  // active->setAll(false), Graph.h:236-237
  if (dalloc(&g->d_flags, 16)) return 1;
  CK(cudaMemsetAsync(g->d_flags, 0, 64, st));
  {
    // the auxiliary stream carries the latency-bound kernels of a pass (long rows): highest priority, so that their
    // few large blocks are placed before the thousands of sliced-ELL blocks that would otherwise fill every SM
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    CK(cudaStreamCreateWithPriority(&g->aux_stream, cudaStreamNonBlocking, getenv("GM_NO_PRIORITY") ? lo : hi));
  }
  CK(cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming));
  if (const char* e = getenv("GM_HOT_LIMIT")) g->hot_limit = atoi(e);
  if (const char* e = getenv("GM_PUSH_DIVISOR")) g->push_divisor = atoi(e);
  if (const char* e = getenv("GM_PUSH_MIN_NNZ")) g->push_min_nnz = atoll(e);
  CK(cudaStreamCreateWithFlags(&g->aux_stream2, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&g->aux_stream3, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&g->ev_join2, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&g->ev_join3, cudaEventDisableTiming));
  {
    // One auxiliary stream by default: two or three (warp-per-row heavy rows / the narrow tail on streams of their
    // own, GM_AUX_STREAMS) measured no faster -- rank 0 of 8: 0.875 vs 0.812 ms per pass, 1 GPU: 3.78 vs 3.78
    // (profiles/r2_multistream_rejected.txt, which also records the kernel race those experiments uncovered).
    int n_aux = 1;
    if (const char* e = getenv("GM_AUX_STREAMS")) n_aux = atoi(e);
    if (getenv("GM_NO_AUX_STREAM")) n_aux = 0;
    if (n_aux < 3) { cudaStreamDestroy(g->aux_stream3); g->aux_stream3 = nullptr; }
    if (n_aux < 2) { cudaStreamDestroy(g->aux_stream2); g->aux_stream2 = nullptr; }
    if (n_aux < 1) { cudaStreamDestroy(g->aux_stream); g->aux_stream = nullptr; }
  }
  CK(cudaMallocHost((void**)&g->h_flags, 64));
  memset(g->h_flags, 0, 64);
  CK(cudaStreamSynchronize(st));
  return 0;
}

static gm_graph* graph_new(int nvertices, int sizeof_E, int sizeof_V, const gm_graph_opts* opts) {
  gm_graph* g = new gm_graph();
  g->n = nvertices;
  g->sizeof_E = sizeof_E;
  g->sizeof_V = sizeof_V;
  g->ref_threads = (opts && opts->ref_threads > 0) ? opts->ref_threads : 4;
  g->rank = opts ? opts->rank : 0;
  g->world = (opts && opts->world > 0) ? opts->world : 1;
  g->heavy_threshold = (opts && opts->heavy_threshold > 0) ? opts->heavy_threshold : GM_DEFAULT_HEAVY_THRESHOLD;
  g->coop_threshold = (opts && opts->coop_threshold > 0) ? opts->coop_threshold : GM_DEFAULT_COOP_THRESHOLD;
  g->heavy_auto = !(opts && opts->heavy_threshold > 0);
  if (const char* e = getenv("GM_HEAVY_THRESHOLD")) if (g->heavy_auto) { g->heavy_threshold = atoi(e); g->heavy_auto = false; }
  if (const char* e = getenv("GM_COOP_THRESHOLD")) if (!(opts && opts->coop_threshold > 0)) g->coop_threshold = atoi(e);
  // staging the longest rows pays when a rank's pass is short enough for one row to be its critical path
  // (sharded runs); on one GPU the extra gather kernel only competes with the sliced-ELL rows (measured)
  g->long_threshold = g->world > 1 ? GM_DEFAULT_LONG_THRESHOLD : 0x7fffffff;
  if (const char* e = getenv("GM_LONG_ROW")) g->long_threshold = atoi(e);
  int per = (nvertices + g->world - 1) / g->world;
  g->n_pad = std::max(32, (per + 31) / 32 * 32);
  g->n_local = nvertices > g->rank ? (nvertices - g->rank + g->world - 1) / g->world : 0;
  g->n_full = g->n_pad * g->world;
  return g;
}

extern "C" int gm_graph_create(gm_graph** out, int nvertices, long long nnz, const int* src, const int* dst,
                               const void* val, int sizeof_E, int sizeof_V, const gm_graph_opts* opts) {
  *out = nullptr;
  if (nvertices <= 0 || nnz < 0 || nnz >= (1ll << 32)) {
    gm_set_error("gm_graph_create: bad nvertices/nnz");
    return 1;
  }
  if (sizeof_E != 4) {
    gm_set_error("gm_graph_create: only 4-byte edge values are supported (the five programs use E = int)");
    return 1;
  }
  if (sizeof_V <= 0 || sizeof_V % 4) {
    gm_set_error("gm_graph_create: sizeof_V must be a positive multiple of 4");
    return 1;
  }
  gm_graph* g = graph_new(nvertices, sizeof_E, sizeof_V, opts);
  if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) {
    gm_set_error("gm_graph_create: cannot create a CUDA stream (no device?)");
    delete g;
    return 1;
  }
  g->nnz = nnz;
  int *d_src = nullptr, *d_dst = nullptr, *d_val = nullptr;
  cudaMemcpyKind kind = (opts && opts->edges_on_device) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  int rc = dalloc(&d_src, nnz) || dalloc(&d_dst, nnz) || (val && dalloc(&d_val, nnz));
  if (!rc && nnz) {
    bool ok = cudaMemcpyAsync(d_src, src, (size_t)nnz * 4, kind, g->stream) == cudaSuccess &&
              cudaMemcpyAsync(d_dst, dst, (size_t)nnz * 4, kind, g->stream) == cudaSuccess &&
              (!val || cudaMemcpyAsync(d_val, val, (size_t)nnz * 4, kind, g->stream) == cudaSuccess);
    if (!ok) {
      gm_set_error("gm_graph_create: copying the edge list to the device failed");
      rc = 1;
    }
  }
  int mask = (opts && opts->build_mask) ? opts->build_mask : 3;
  if (!rc) rc = build_graph<int>(g, d_src, d_dst, d_val, nnz, opts ? opts->order_like : nullptr, mask);
  cudaFree(d_src);
  cudaFree(d_dst);
  cudaFree(d_val);
  if (rc || graph_common_alloc(g)) {
    gm_graph_destroy(g);
    return 1;
  }
  *out = g;
  return 0;
}

// Graph::applyToAllEdges (include/Graph.h:389-402) rewrites every stored edge value in both operand
// matrices.  The per-edge function is host code in the reference's signature, so the host side
// evaluates it and hands the new values over with the edge list; the matrices are rebuilt in place
// (same placement, same row order, same fold order) while vertex properties and the active set stay.
extern "C" int gm_graph_set_edge_values(gm_graph* g, long long nnz, const int* src, const int* dst, const void* val) {
  if (nnz != g->nnz) {
    gm_set_error("gm_graph_set_edge_values: nnz differs from the graph's");
    return 1;
  }
  cudaStream_t st = g->stream;
  CK(cudaStreamSynchronize(st));
  if (g->aux_stream) CK(cudaStreamSynchronize(g->aux_stream));
  if (g->aux_stream2) CK(cudaStreamSynchronize(g->aux_stream2));
  if (g->aux_stream3) CK(cudaStreamSynchronize(g->aux_stream3));
  const bool hasA = g->A.n_slots > 0, hasAT = g->AT.n_slots > 0;
  const bool identA = g->A.identity != 0, identAT = g->AT.identity != 0;
  int *d_src = nullptr, *d_dst = nullptr, *d_val = nullptr;
  if (dalloc(&d_src, nnz) || dalloc(&d_dst, nnz) || dalloc(&d_val, nnz)) return 1;
  if (nnz) {
    CK(cudaMemcpyAsync(d_src, src, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_dst, dst, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_val, val, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    if (check_ids(d_src, d_dst, nnz, g->n, st, "gm_graph_set_edge_values")) {
      cudaFree(d_src); cudaFree(d_dst); cudaFree(d_val);
      return 1;
    }
    const int npart = g->ref_threads * 16;
    k_to_native<<<nblk(nnz), 256, 0, st>>>(d_src, nnz, g->n, npart);
    k_to_native<<<nblk(nnz), 256, 0, st>>>(d_dst, nnz, g->n, npart);
  }
  int rc = 0;
  if (hasAT) {
    matrix_free(g->AT);
    g->AT = gm_matrix();
    rc = build_matrix<int>(g, g->AT, d_dst, d_src, d_val, nnz, identAT);
  }
  if (!rc && hasA) {
    matrix_free(g->A);
    g->A = gm_matrix();
    rc = build_matrix<int>(g, g->A, d_src, d_dst, d_val, nnz, identA);
  }
  cudaFree(d_src);
  cudaFree(d_dst);
  cudaFree(d_val);
  return rc;
}

extern "C" int gm_graph_create_rmat(gm_graph** out, int scale, int edge_factor, unsigned long long seed, int weight_max,
                                    unsigned long long weight_seed, int sizeof_V, const gm_graph_opts* opts) {
  *out = nullptr;
  if (scale < 1 || scale > 30 || edge_factor < 1 || ((long long)edge_factor << scale) >= (1ll << 32)) {
    gm_set_error("gm_graph_create_rmat: bad scale/edge_factor");
    return 1;
  }
  long long nnz = (long long)edge_factor << scale;
  gm_graph* g = graph_new(1 << scale, 4, sizeof_V, opts);
  if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) {
    gm_set_error("gm_graph_create_rmat: cannot create a CUDA stream (no device?)");
    delete g;
    return 1;
  }
  g->nnz = nnz;
  int *d_src = nullptr, *d_dst = nullptr, *d_val = nullptr;
  int rc = dalloc(&d_src, nnz) || dalloc(&d_dst, nnz) || (weight_max > 0 && dalloc(&d_val, nnz));
  if (!rc) {
    k_rmat<<<nblk(nnz), 256, 0, g->stream>>>(scale, seed, weight_max, weight_seed, nnz, d_src, d_dst, d_val);
    if (cudaGetLastError() != cudaSuccess) {
      gm_set_error("gm_graph_create_rmat: generator launch failed");
      rc = 1;
    }
  }
  int mask = (opts && opts->build_mask) ? opts->build_mask : 3;
  if (!rc) rc = build_graph<int>(g, d_src, d_dst, d_val, nnz, opts ? opts->order_like : nullptr, mask);
  cudaFree(d_src);
  cudaFree(d_dst);
  cudaFree(d_val);
  if (rc || graph_common_alloc(g)) {
    gm_graph_destroy(g);
    return 1;
  }
  *out = g;
  return 0;
}

extern "C" int gm_graph_destroy(gm_graph* g) {
  if (!g) return 0;
  if (g->stream) cudaStreamSynchronize(g->stream);
  matrix_free(g->A);
  matrix_free(g->AT);
  cudaFree(g->d_xidx);
  cudaFree(g->push_scratch);
  if (g->vp_owner) cudaFree(g->vp);
  cudaFree(g->active);
  cudaFree(g->d_flags);
  if (g->h_flags) cudaFreeHost(g->h_flags);
  cudaFree(g->staging);
  gm_sym_free(g, &g->sym_staging);
  gm_sym_free(g, &g->sync);
  for (void* p : g->retired) cudaFree(p);
  if (g->aux_stream) cudaStreamDestroy(g->aux_stream);
  if (g->aux_stream2) cudaStreamDestroy(g->aux_stream2);
  if (g->aux_stream3) cudaStreamDestroy(g->aux_stream3);
  if (g->ev_fork) cudaEventDestroy(g->ev_fork);
  if (g->ev_join) cudaEventDestroy(g->ev_join);
  if (g->ev_join2) cudaEventDestroy(g->ev_join2);
  if (g->ev_join3) cudaEventDestroy(g->ev_join3);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
  return 0;
}

extern "C" int gm_graph_view_get(const gm_graph* g, gm_graph_view* v) {
  v->nvertices = g->n;
  v->n_local = g->n_local;
  v->n_local_pad = g->n_pad;
  v->n_full = g->n_full;
  v->rank = g->rank;
  v->world = g->world;
  v->ref_threads = g->ref_threads;
  v->sizeof_V = g->sizeof_V;
  v->sizeof_E = g->sizeof_E;
  v->nnz = g->nnz;
  v->vertexproperty = g->vp;
  v->active_bits = g->active;
  fill_view(g->A, &v->A);
  fill_view(g->AT, &v->AT);
  v->d_flags = g->d_flags;
  v->h_flags = g->h_flags;
  v->stream = (void*)g->stream;
  v->aux_stream = (void*)g->aux_stream;
  v->ev_fork = (void*)g->ev_fork;
  v->ev_join = (void*)g->ev_join;
  v->aux_stream2 = (void*)g->aux_stream2;
  v->aux_stream3 = (void*)g->aux_stream3;
  v->ev_join2 = (void*)g->ev_join2;
  v->ev_join3 = (void*)g->ev_join3;
  v->hot_limit = g->hot_limit;
  v->owner = const_cast<gm_graph*>(g);
  v->push_divisor = g->push_divisor;
  v->push_min_nnz = g->push_min_nnz;
  return 0;
}

extern "C" int gm_graph_synchronize(const gm_graph* g) {
  CK(cudaStreamSynchronize(g->stream));
  return 0;
}

static int staging(gm_graph* g, size_t bytes);
// ----------------------------------------------------------------- activity --
__global__ void k_fill_words(unsigned* bits, int n_valid, int n_pad, unsigned fill) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (n_pad >> 5)) return;
  int lo = w << 5;
  unsigned v = 0;
  if (lo + 32 <= n_valid) v = 0xffffffffu;
  else if (lo < n_valid) v = (1u << (n_valid - lo)) - 1u;
  bits[w] = v & fill;
}
__global__ void k_set_bit(unsigned* bits, int i, int on) {
  if (on) atomicOr(bits + (i >> 5), 1u << (i & 31));
  else atomicAnd(bits + (i >> 5), ~(1u << (i & 31)));
}

static int host_xidx(const gm_graph* g) {
  gm_graph* m = const_cast<gm_graph*>(g);
  if (m->h_xidx.empty()) {
    m->h_xidx.resize(g->n);
    if (d2h(m->h_xidx.data(), g->d_xidx, (size_t)g->n * 4, g->stream)) return 1;
  }
  return 0;
}
// public 1-based id -> (owner, local); returns 1 on a bad id
static int locate(const gm_graph* g, int v, int* owner, int* local) {
  if (v < 1 || v > g->n) {
    gm_set_error("vertex id out of range (ids are 1-based)");
    return 1;
  }
  if (host_xidx(g)) return 1;
  int xi = g->h_xidx[to_native0(v, g->n, g->ref_threads * 16)];
  *owner = xi / g->n_pad;
  *local = xi % g->n_pad;
  return 0;
}

extern "C" int gm_graph_set_all_active(gm_graph* g) {
  k_fill_words<<<nblk(g->n_pad >> 5), 256, 0, g->stream>>>(g->active, g->n_local, g->n_pad, 0xffffffffu);
  CK(cudaGetLastError());
  return 0;
}
extern "C" int gm_graph_set_all_inactive(gm_graph* g) {
  CK(cudaMemsetAsync(g->active, 0, (size_t)(g->n_pad >> 5) * 4, g->stream));
  return 0;
}
static int set_active(gm_graph* g, int v, int on) {
  int owner, local;
  if (locate(g, v, &owner, &local)) return 1;
  if (owner != g->rank) return 0;
  k_set_bit<<<1, 1, 0, g->stream>>>(g->active, local, on);
  CK(cudaGetLastError());
  return 0;
}
__global__ void k_set_active_array(unsigned* bits, const unsigned char* flags, int n, int npart, const int* xidx, int n_pad,
                                   int rank) {
  int pub = blockIdx.x * blockDim.x + threadIdx.x;
  if (pub >= n || !flags[pub]) return;
  int xi = xidx[to_native0(pub + 1, n, npart)];
  if (xi / n_pad != rank) return;
  int local = xi % n_pad;
  atomicOr(bits + (local >> 5), 1u << (local & 31));
}
// the active set from a host array of n flags in public-id order (flags[v-1] != 0 <=> setActive(v), all others inactive):
// what the apps' "for every vertex: if (...) G.setActive(i)" loops amount to (src/TopologicalSort.cpp:157-167)
extern "C" int gm_graph_set_active_array(gm_graph* g, const unsigned char* flags) {
  if (staging(g, (size_t)g->n)) return 1;
  CK(cudaMemcpyAsync(g->staging, flags, (size_t)g->n, cudaMemcpyHostToDevice, g->stream));
  CK(cudaMemsetAsync(g->active, 0, (size_t)(g->n_pad >> 5) * 4, g->stream));
  k_set_active_array<<<nblk(g->n), 256, 0, g->stream>>>(g->active, (const unsigned char*)g->staging, g->n, g->ref_threads * 16,
                                                       g->d_xidx, g->n_pad, g->rank);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g->stream));
  return 0;
}
extern "C" int gm_graph_set_active(gm_graph* g, int v) { return set_active(g, v, 1); }
extern "C" int gm_graph_set_inactive(gm_graph* g, int v) { return set_active(g, v, 0); }

extern "C" int gm_graph_vertex_owner(const gm_graph* g, int v) {
  int owner, local;
  if (locate(g, v, &owner, &local)) return -1;
  return owner;
}
extern "C" int gm_graph_out_degree_source(const gm_graph* g, int* v) {
  *v = g->first_source;
  return g->first_source > 0 ? 0 : 1;
}

// ---------------------------------------------------------- vertex property --
__global__ void k_broadcast_words(unsigned* dst, const unsigned* one, int words_per, long long total_words) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < total_words) dst[i] = one[i % words_per];
}
// public-order array <-> placement-order storage
__global__ void k_permute_in(unsigned* vp, const unsigned* in, int n, int npart, const int* xidx, int n_pad, int rank,
                             int words_per) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long pub = t / words_per;
  int w = (int)(t % words_per);
  if (pub >= n) return;
  int xi = xidx[to_native0((int)pub + 1, n, npart)];
  if (xi / n_pad != rank) return;
  vp[(long long)(xi % n_pad) * words_per + w] = in[pub * words_per + w];
}
__global__ void k_permute_out(const unsigned* vp, unsigned* out, int n, int npart, const int* xidx, int n_pad, int rank,
                              int words_per) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long pub = t / words_per;
  int w = (int)(t % words_per);
  if (pub >= n) return;
  int xi = xidx[to_native0((int)pub + 1, n, npart)];
  if (xi / n_pad != rank) return;
  out[pub * words_per + w] = vp[(long long)(xi % n_pad) * words_per + w];
}

static int staging(gm_graph* g, size_t bytes) {
  if (g->staging_bytes < bytes) {
    cudaFree(g->staging);
    g->staging = nullptr;
    g->staging_bytes = 0;
    CK(cudaMalloc(&g->staging, bytes));
    g->staging_bytes = bytes;
  }
  return 0;
}

extern "C" int gm_graph_set_all_vertexproperty(gm_graph* g, const void* value) {
  if (staging(g, g->sizeof_V)) return 1;
  CK(cudaMemcpyAsync(g->staging, value, g->sizeof_V, cudaMemcpyHostToDevice, g->stream));
  int wp = g->sizeof_V / 4;
  long long tw = (long long)g->n_pad * wp;
  k_broadcast_words<<<nblk(tw), 256, 0, g->stream>>>((unsigned*)g->vp, (const unsigned*)g->staging, wp, tw);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g->stream));
  return 0;
}
extern "C" int gm_graph_set_vertexproperty(gm_graph* g, int v, const void* value) {
  int owner, local;
  if (locate(g, v, &owner, &local)) return 1;
  if (owner != g->rank) return 0;
  CK(cudaMemcpyAsync((char*)g->vp + (size_t)local * g->sizeof_V, value, g->sizeof_V, cudaMemcpyHostToDevice, g->stream));
  CK(cudaStreamSynchronize(g->stream));
  return 0;
}
extern "C" int gm_graph_get_vertexproperty(const gm_graph* g, int v, void* value) {
  int owner, local;
  if (locate(g, v, &owner, &local)) return 1;
  if (owner != g->rank) return 2;
  CK(cudaMemcpyAsync(value, (const char*)g->vp + (size_t)local * g->sizeof_V, g->sizeof_V, cudaMemcpyDeviceToHost,
                     g->stream));
  CK(cudaStreamSynchronize(g->stream));
  return 0;
}
extern "C" int gm_graph_set_vertexproperties(gm_graph* g, const void* values) {
  size_t bytes = (size_t)g->n * g->sizeof_V;
  if (staging(g, bytes)) return 1;
  CK(cudaMemcpyAsync(g->staging, values, bytes, cudaMemcpyHostToDevice, g->stream));
  int wp = g->sizeof_V / 4;
  long long t = (long long)g->n * wp;
  k_permute_in<<<nblk(t), 256, 0, g->stream>>>((unsigned*)g->vp, (const unsigned*)g->staging, g->n, g->ref_threads * 16,
                                              g->d_xidx, g->n_pad, g->rank, wp);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g->stream));
  return 0;
}
extern "C" int gm_graph_get_vertexproperties(const gm_graph* g, void* values) {
  gm_graph* m = const_cast<gm_graph*>(g);
  size_t bytes = (size_t)g->n * g->sizeof_V;
  if (staging(m, bytes)) return 1;
  int wp = g->sizeof_V / 4;
  long long t = (long long)g->n * wp;
  if (g->world > 1) CK(cudaMemcpyAsync(m->staging, values, bytes, cudaMemcpyHostToDevice, g->stream));
  k_permute_out<<<nblk(t), 256, 0, g->stream>>>((const unsigned*)g->vp, (unsigned*)m->staging, g->n, g->ref_threads * 16,
                                               g->d_xidx, g->n_pad, g->rank, wp);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(values, m->staging, bytes, cudaMemcpyDeviceToHost, g->stream));
  CK(cudaStreamSynchronize(g->stream));
  return 0;
}
extern "C" int gm_graph_share_vertexproperty(gm_graph* g, gm_graph* owner) {
  if (g->n != owner->n || g->n_pad != owner->n_pad || g->sizeof_V != owner->sizeof_V) {
    gm_set_error("gm_graph_share_vertexproperty: graphs differ in shape");
    return 1;
  }
  // the two graphs must agree on vertex placement: check a sample of the map
  if (host_xidx(g) || host_xidx(owner)) return 1;
  if (g->h_xidx != owner->h_xidx) {
    gm_set_error("gm_graph_share_vertexproperty: create the graph with opts.order_like = owner");
    return 1;
  }
  if (g->vp_owner) cudaFree(g->vp);
  g->vp = owner->vp;
  g->vp_owner = false;
  return 0;
}


// ------------------------------------------------- sparse frontiers (push) --
// Column-major companion of one operand matrix, derived on the device from the row-major arrays:
// every entry becomes (x index; row slot, position in the row's fold order, edge value), sorted by
// x index.  The position ("rank") lets the push path fold a row's contributions in exactly the
// order the row-major kernels (and the reference, spmspv.h:55-77) use.
__global__ void k_csc_emit_heavy(const long long* h_ptr, const int* h_col, const unsigned* h_val, int n_heavy,
                                 unsigned* key, unsigned* eidx, int* row, int* rank, unsigned* val) {
  int r = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;  // warp per row
  if (r >= n_heavy) return;
  long long b = h_ptr[r], e = h_ptr[r + 1];
  for (long long i = b + lane; i < e; i += 32) {
    key[i] = (unsigned)h_col[i];
    eidx[i] = (unsigned)i;
    row[i] = r;
    rank[i] = (int)(i - b);
    val[i] = h_val[i];
  }
}
__global__ void k_csc_emit_sell(const long long* slice_ptr, const int* s_col, const unsigned* s_val, const int* row_len,
                                const long long* row_off, int n_heavy, int n_rows, long long nh, unsigned* key,
                                unsigned* eidx, int* row, int* rank, unsigned* val) {
  int rel = blockIdx.x * blockDim.x + threadIdx.x;  // thread per sliced-ELL row: a warp reads one slice step coalesced
  if (rel >= n_rows) return;
  int slot = n_heavy + rel;
  int len = row_len[slot];
  long long base = slice_ptr[rel >> 5] + (rel & 31);
  long long out = nh + row_off[rel];
  for (int i = 0; i < len; i++) {
    long long pos = base + 32ll * i;
    key[out + i] = (unsigned)s_col[pos];
    eidx[out + i] = (unsigned)(out + i);
    row[out + i] = slot;
    rank[out + i] = i;
    val[out + i] = s_val[pos];
  }
}
__global__ void k_csc_fill(const unsigned* key, const unsigned* eidx, const int* row, const int* rank, const unsigned* val,
                           long long total, int* c_row, int* c_rank, unsigned* c_val, long long* c_cnt) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  unsigned e = eidx[i];
  c_row[i] = row[e];
  c_rank[i] = rank[e];
  c_val[i] = val[e];
  atomicAdd((unsigned long long*)(c_cnt + key[i]), 1ull);
}
__global__ void k_len_slice_to_ll(const int* row_len, int first, int n, long long* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = row_len[first + i];
}

__global__ void k_big_cols(const long long* c_ptr, int n_full, int thr, int* n_big, int* big_cols) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_full && c_ptr[c + 1] - c_ptr[c] > thr) big_cols[atomicAdd(n_big, 1)] = c;
}

static int build_push(gm_graph* g, gm_matrix& M) {
  if (M.push_built) return 0;
  cudaStream_t st = g->stream;
  if (g->sizeof_E != 4) {
    gm_set_error("push companion: only 4-byte edge values are supported");
    return 1;
  }
  const long long total = M.nnz;
  if (total >= (1ll << 32)) {
    gm_set_error("push companion: more than 2^32 entries per rank");
    return 1;
  }
  const int n_rows = M.n_slices > 0 ? std::min(M.n_slices * 32, M.n_slots - M.n_heavy) : 0;
  long long nh = 0;
  if (M.n_heavy > 0 && d2h(&nh, M.h_ptr + M.n_heavy, 8, st)) return 1;
  int maxlen = 0;
  if (M.n_slots > 0 && d2h(&maxlen, M.row_len, 4, st)) return 1;  // rows are stored longest first
  M.rank_bits = std::max(1, ceil_log2((long long)maxlen + 1));
  unsigned *k0 = nullptr, *k1 = nullptr, *e0 = nullptr, *e1 = nullptr, *val = nullptr, *ks = nullptr, *es = nullptr;
  int *row = nullptr, *rank = nullptr;
  long long *len_ll = nullptr, *row_off = nullptr, *cnt = nullptr;
  if (dalloc(&k0, total) || dalloc(&k1, total) || dalloc(&e0, total) || dalloc(&e1, total) || dalloc(&val, total) ||
      dalloc(&row, total) || dalloc(&rank, total))
    return 1;
  if (nh > 0)
    k_csc_emit_heavy<<<nblk((long long)M.n_heavy * 32), 256, 0, st>>>(M.h_ptr, M.h_col, (const unsigned*)M.h_val, M.n_heavy,
                                                                     k0, e0, row, rank, val);
  if (n_rows > 0) {
    if (dalloc(&len_ll, (size_t)n_rows + 1) || dalloc(&row_off, (size_t)n_rows + 1)) return 1;
    CK(cudaMemsetAsync(len_ll, 0, ((size_t)n_rows + 1) * 8, st));
    k_len_slice_to_ll<<<nblk(n_rows), 256, 0, st>>>(M.row_len, M.n_heavy, n_rows, len_ll);
    if (exclusive_scan_ll(len_ll, row_off, n_rows, st)) return 1;
    k_csc_emit_sell<<<nblk(n_rows), 256, 0, st>>>(M.slice_ptr, M.s_col, (const unsigned*)M.s_val, M.row_len, row_off,
                                                 M.n_heavy, n_rows, nh, k0, e0, row, rank, val);
  }
  CK(cudaGetLastError());
  if (total > 0 && sort_pairs(k0, k1, e0, e1, total, std::max(1, ceil_log2(g->n_full)), st, &ks, &es)) return 1;
  unsigned* cv = nullptr;
  if (dalloc(&M.c_row, total) || dalloc(&M.c_rank, total) || dalloc(&cv, total) || dalloc(&cnt, (size_t)g->n_full + 1) ||
      dalloc(&M.c_ptr, (size_t)g->n_full + 1))
    return 1;
  M.c_val = cv;
  CK(cudaMemsetAsync(cnt, 0, ((size_t)g->n_full + 1) * 8, st));
  if (total > 0) k_csc_fill<<<nblk(total), 256, 0, st>>>(ks, es, row, rank, val, total, M.c_row, M.c_rank, cv, cnt);
  CK(cudaGetLastError());
  if (exclusive_scan_ll(cnt, M.c_ptr, g->n_full, st)) return 1;
  {  // the columns above GM_PUSH_BIG_COL entries (hubs): a short list the atomic push walks with many blocks
    int* nb = nullptr;
    if (dalloc(&nb, 1) || dalloc(&M.big_cols, (size_t)(total / GM_PUSH_BIG_COL + 1))) return 1;
    CK(cudaMemsetAsync(nb, 0, 4, st));
    k_big_cols<<<nblk(g->n_full), 256, 0, st>>>(M.c_ptr, g->n_full, GM_PUSH_BIG_COL, nb, M.big_cols);
    CK(cudaGetLastError());
    if (d2h(&M.n_big_cols, nb, 4, st)) return 1;
    cudaFree(nb);
  }
  cudaFree(k0); cudaFree(k1); cudaFree(e0); cudaFree(e1); cudaFree(val); cudaFree(row); cudaFree(rank);
  cudaFree(len_ll); cudaFree(row_off); cudaFree(cnt);
  M.push_built = true;
  return 0;
}

static gm_matrix* which_matrix(gm_graph* g, int which) { return which == 0 ? &g->A : &g->AT; }

static void push_free(gm_matrix& M) {
  cudaFree(M.c_ptr); cudaFree(M.c_row); cudaFree(M.c_rank); cudaFree(M.c_val); cudaFree(M.big_cols);
  M.c_ptr = nullptr; M.c_row = nullptr; M.c_rank = nullptr; M.c_val = nullptr; M.big_cols = nullptr;
  M.n_big_cols = 0;
  M.push_built = false;
}
extern "C" int gm_graph_edges_changed(gm_graph* g) {
  CK(cudaStreamSynchronize(g->stream));
  push_free(g->A);
  push_free(g->AT);
  return 0;
}
extern "C" int gm_graph_set_push_policy(gm_graph* g, int divisor, long long min_nnz) {
  g->push_divisor = divisor < 0 ? 0 : divisor;
  g->push_min_nnz = min_nnz < 0 ? 0 : min_nnz;
  return 0;
}
extern "C" int gm_graph_push_ready(gm_graph* g, int which) {
  gm_matrix* M = which_matrix(g, which);
  if (M->n_slots == 0) {
    gm_set_error("push companion: this operand matrix was not built (build_mask)");
    return 1;
  }
  return build_push(g, *M);
}

// frontier statistics: active columns that own entries here, and how many entries they have
__global__ void k_push_count(const unsigned* __restrict__ xbits, int n_words, const long long* __restrict__ c_ptr,
                             unsigned long long* out) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long cols = 0, ents = 0;
  if (w < n_words) {
    unsigned m = xbits[w];
    while (m) {
      int b = __ffs(m) - 1;
      m &= m - 1;
      long long d = c_ptr[w * 32 + b + 1] - c_ptr[w * 32 + b];
      if (d > 0) { cols++; ents += (unsigned long long)d; }
    }
  }
  for (int o = 16; o; o >>= 1) {
    cols += __shfl_down_sync(0xffffffffu, cols, o);
    ents += __shfl_down_sync(0xffffffffu, ents, o);
  }
  if ((threadIdx.x & 31) == 0 && cols) {
    atomicAdd(out, cols);
    atomicAdd(out + 1, ents);
  }
}
__global__ void k_push_compact(const unsigned* __restrict__ xbits, int n_words, const long long* __restrict__ c_ptr,
                               int* counter, int* f_col, long long* f_deg) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  unsigned m = xbits[w];
  while (m) {
    int b = __ffs(m) - 1;
    m &= m - 1;
    long long d = c_ptr[w * 32 + b + 1] - c_ptr[w * 32 + b];
    if (d > 0) {
      int p = atomicAdd(counter, 1);
      f_col[p] = w * 32 + b;
      f_deg[p] = d;
    }
  }
}

extern "C" int gm_push_count(gm_graph* g, int which, const gm_vectors* v, int* n_active, long long* n_entries) {
  gm_matrix* M = which_matrix(g, which);
  if (build_push(g, *M)) return 1;
  cudaStream_t st = g->stream;
  unsigned long long* d = reinterpret_cast<unsigned long long*>(g->d_flags + 8);
  CK(cudaMemsetAsync(d, 0, 16, st));
  const int n_words = g->n_full >> 5;
  k_push_count<<<nblk(n_words), 256, 0, st>>>(v->x_bits, n_words, M->c_ptr, d);
  CK(cudaGetLastError());
  unsigned long long* h = reinterpret_cast<unsigned long long*>(g->h_flags + 8);  // pinned: no staging copy
  if (d2h(h, d, 16, st)) return 1;
  *n_active = (int)h[0];
  *n_entries = (long long)h[1];
  return 0;
}

extern "C" int gm_push_prepare(gm_graph* g, int which, gm_vectors* v, int n_active, long long n_entries,
                               gm_push_plan* plan) {
  gm_matrix* M = which_matrix(g, which);
  if (!M->push_built) {
    gm_set_error("gm_push_prepare before gm_push_count");
    return 1;
  }
  cudaStream_t st = g->stream;
  memset(plan, 0, sizeof *plan);
  plan->n_active = n_active;
  plan->n_entries = n_entries;
  int slot_bits = std::max(1, ceil_log2(M->n_slots));
  plan->key_bits = std::min(64, slot_bits + M->rank_bits);
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  size_t scan_tb = 0, sort_tb = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, scan_tb, (long long*)nullptr, (long long*)nullptr, n_active + 1, st));
  {
    cub::DoubleBuffer<unsigned long long> kb(nullptr, nullptr);
    cub::DoubleBuffer<unsigned> vb(nullptr, nullptr);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, sort_tb, kb, vb, n_entries, 0, plan->key_bits, st));
  }
  const size_t tmp_b = up(std::max(scan_tb, sort_tb));
  const size_t b_col = up((size_t)n_active * 4), b_off = up(((size_t)n_active + 1) * 8);
  const size_t b_key = up((size_t)n_entries * 8), b_ord = up((size_t)n_entries * 4), b_val = up((size_t)n_entries * v->sizeof_U);
  const size_t need = b_col + b_off + 2 * b_key + 2 * b_ord + b_val + tmp_b;
  if (need > g->push_scratch_bytes) {  // kept in the graph: run_graph_program may create its vectors per call
    if (g->push_scratch) g->retired.push_back(g->push_scratch);  // freed with the graph (cudaFree waits for the device)
    g->push_scratch = nullptr;
    g->push_scratch_bytes = 0;
    const size_t want = std::max(need + need / 2, (size_t)1 << 22);
    CK(cudaMalloc(&g->push_scratch, want));
    g->push_scratch_bytes = want;
  }
  unsigned char* p = (unsigned char*)g->push_scratch;
  plan->f_col = (int*)p; p += b_col;
  plan->f_off = (long long*)p; p += b_off;
  plan->keys = (unsigned long long*)p; p += b_key;
  plan->keys_alt = (unsigned long long*)p; p += b_key;
  plan->order = (unsigned*)p; p += b_ord;
  plan->order_alt = (unsigned*)p; p += b_ord;
  plan->vals = p; p += b_val;
  plan->sort_tmp = p;
  plan->sort_tmp_bytes = (long long)tmp_b;
  int* counter = g->d_flags + 12;
  CK(cudaMemsetAsync(counter, 0, 4, st));
  CK(cudaMemsetAsync(plan->f_off + n_active, 0, 8, st));
  const int n_words = g->n_full >> 5;
  k_push_compact<<<nblk(n_words), 256, 0, st>>>(v->x_bits, n_words, M->c_ptr, counter, plan->f_col, plan->f_off);
  CK(cudaGetLastError());
  size_t tb = tmp_b;
  CK(cub::DeviceScan::ExclusiveSum(plan->sort_tmp, tb, plan->f_off, plan->f_off, n_active + 1, st));
  return 0;
}

extern "C" int gm_push_sort(gm_graph* g, gm_push_plan* plan) {
  cub::DoubleBuffer<unsigned long long> kb(plan->keys, plan->keys_alt);
  cub::DoubleBuffer<unsigned> vb(plan->order, plan->order_alt);
  size_t tb = (size_t)plan->sort_tmp_bytes;
  CK(cub::DeviceRadixSort::SortPairs(plan->sort_tmp, tb, kb, vb, plan->n_entries, 0, plan->key_bits, g->stream));
  if (kb.Current() != plan->keys) std::swap(plan->keys, plan->keys_alt);
  if (vb.Current() != plan->order) std::swap(plan->order, plan->order_alt);
  return 0;
}

// ------------------------------------------------------------------ vectors --
extern "C" int gm_vectors_create(gm_vectors** out, const gm_graph* g, int sizeof_T, int sizeof_U) {
  *out = nullptr;
  if (sizeof_T <= 0 || sizeof_U <= 0) {
    gm_set_error("gm_vectors_create: message sizes must be positive");
    return 1;
  }
  gm_vectors* v = new gm_vectors();
  gm_graph* gm = const_cast<gm_graph*>(g);
  v->g = gm;
  v->sizeof_T = sizeof_T;
  v->sizeof_U = sizeof_U;
  v->n_full = g->n_full;
  v->n_pad = g->n_pad;
  const size_t xb = (size_t)g->n_full * sizeof_T, bb = (size_t)(g->n_full >> 5) * 4;
  if (g->peers_on) {  // the message buffers are mapped on every rank (gm_peer.cu); zero-filled by gm_sym_alloc
    v->sym = true;  // before the allocations: a failure of the second one releases the first through gm_sym_free
    if (gm_sym_alloc(gm, xb, &v->s_val) || gm_sym_alloc(gm, bb, &v->s_bits)) {
      gm_vectors_destroy(v);
      return 1;
    }
    v->x_val = v->s_val.local;
    v->x_bits = (unsigned*)v->s_bits.local;
  } else {
    if (cudaMalloc(&v->x_val, xb) != cudaSuccess || cudaMalloc((void**)&v->x_bits, bb) != cudaSuccess) {
      gm_set_error("gm_vectors_create: out of device memory");
      gm_vectors_destroy(v);
      return 1;
    }
    cudaMemsetAsync(v->x_val, 0, xb, g->stream);
    cudaMemsetAsync(v->x_bits, 0, bb, g->stream);
  }
  if (cudaMalloc(&v->y_val, (size_t)g->n_pad * sizeof_U) != cudaSuccess ||
      cudaMalloc((void**)&v->y_bits, (size_t)(g->n_pad >> 5) * 4) != cudaSuccess) {
    gm_set_error("gm_vectors_create: out of device memory");
    gm_vectors_destroy(v);
    return 1;
  }
  cudaMemsetAsync(v->y_val, 0, (size_t)g->n_pad * sizeof_U, g->stream);
  cudaMemsetAsync(v->y_bits, 0, (size_t)(g->n_pad >> 5) * 4, g->stream);
  if (cudaStreamSynchronize(g->stream) != cudaSuccess) {
    gm_set_error("gm_vectors_create: device error");
    gm_vectors_destroy(v);
    return 1;
  }
  *out = v;
  return 0;
}
extern "C" int gm_vectors_need_alt(gm_vectors* v) {
  if (v->x_alt) return 0;
  const size_t xb = (size_t)v->n_full * v->sizeof_T;
  if (v->sym) {
    if (gm_sym_alloc(v->g, xb, &v->s_alt)) return 1;
    v->x_alt = v->s_alt.local;
  } else {
    CK(cudaMalloc(&v->x_alt, xb));
    CK(cudaMemsetAsync(v->x_alt, 0, xb, v->g->stream));
  }
  return 0;
}
extern "C" int gm_vectors_destroy(gm_vectors* v) {
  if (!v) return 0;
  if (v->sym) {
    gm_sym_free(v->g, &v->s_val);
    gm_sym_free(v->g, &v->s_bits);
    gm_sym_free(v->g, &v->s_alt);
  } else {
    cudaFree(v->x_val);
    cudaFree(v->x_bits);
    cudaFree(v->x_alt);
  }
  cudaFree(v->y_val);
  cudaFree(v->y_bits);
  cudaFree(v->scratch);
  cudaFree(v->aux);
  for (void* p : v->retired) cudaFree(p);
  delete v;
  return 0;
}
extern "C" int gm_vectors_view_get(const gm_vectors* v, gm_vectors_view* o) {
  memset(o, 0, sizeof *o);
  o->sizeof_T = v->sizeof_T;
  o->sizeof_U = v->sizeof_U;
  o->x_val = v->x_val;
  o->x_bits = v->x_bits;
  o->y_val = v->y_val;
  o->y_bits = v->y_bits;
  o->x_alt = v->x_alt;
  if (v->sym && v->g->world > 1) {
    int np = 0;
    for (int q = 0; q < v->g->world; q++)
      if (q != v->g->rank) {
        o->peer_x_val[np] = v->s_val.peer[q];
        o->peer_x_bits[np] = (unsigned*)v->s_bits.peer[q];
        o->peer_x_alt[np] = v->s_alt.peer[q];
        np++;
      }
    o->n_peers = np;
  }
  return 0;
}

extern "C" int gm_vectors_scratch(gm_vectors* v, long long bytes, void** out) {
  if ((size_t)bytes > v->scratch_bytes) {
    // cudaFree waits for the whole device: with several ranks in one process a peer may be spinning in the
    // barrier kernel for THIS rank, so the old block is retired and freed with the vectors
    if (v->scratch) v->retired.push_back(v->scratch);
    v->scratch = nullptr;
    v->scratch_bytes = 0;
    const size_t want = (size_t)bytes + (size_t)bytes / 4;
    CK(cudaMalloc(&v->scratch, want));
    v->scratch_bytes = want;
  }
  *out = v->scratch;
  return 0;
}

extern "C" int gm_vectors_aux(gm_vectors* v, long long bytes, void** out) {
  if ((size_t)bytes > v->aux_bytes) {
    if (v->aux) v->retired.push_back(v->aux);
    v->aux = nullptr;
    v->aux_bytes = 0;
    CK(cudaMalloc(&v->aux, (size_t)bytes));
    CK(cudaMemsetAsync(v->aux, 0xff, (size_t)bytes, v->g->stream));
    v->aux_bytes = (size_t)bytes;
  }
  *out = v->aux;
  return 0;
}

// ----------------------------------------------------------------- exchange --
extern "C" int gm_graph_set_exchange(gm_graph* g, gm_allgather_fn allgather, gm_allreduce_or_fn allreduce_or, void* ctx) {
  g->allgather = allgather;
  g->allreduce_or = allreduce_or;
  g->xctx = ctx;
  return 0;
}
extern "C" int gm_graph_exchange_x_parts(gm_graph* g, gm_vectors* v, int values, int bits) {
  if (g->world == 1) return 0;
  if (!g->allgather) {
    gm_set_error("world > 1 but no exchange functions were registered (gm_graph_set_exchange)");
    return 1;
  }
  if (values && g->allgather(g->xctx, v->x_val, (long long)g->n_pad * v->sizeof_T, (void*)g->stream)) return 1;
  if (bits && g->allgather(g->xctx, v->x_bits, (long long)(g->n_pad >> 5) * 4, (void*)g->stream)) return 1;
  return 0;
}
extern "C" int gm_graph_exchange_x(gm_graph* g, gm_vectors* v) { return gm_graph_exchange_x_parts(g, v, 1, 1); }
extern "C" int gm_graph_exchange_buffer(gm_graph* g, void* buf, long long bytes_per_rank) {
  if (g->world == 1) return 0;
  if (!g->allgather) {
    gm_set_error("world > 1 but no exchange functions were registered (gm_graph_set_exchange)");
    return 1;
  }
  return g->allgather(g->xctx, buf, bytes_per_rank, (void*)g->stream);
}
extern "C" int gm_graph_allreduce_or(gm_graph* g, int* flag) {
  if (g->world == 1) return 0;
  if (!g->allreduce_or) {
    gm_set_error("world > 1 but no exchange functions were registered (gm_graph_set_exchange)");
    return 1;
  }
  return g->allreduce_or(g->xctx, flag);
}

// ------------------------------------------------------------------- reduce --
__global__ void k_reduce(const unsigned* vp, int n_local, int words_per, int what, int param, double* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < n_local) {
    const unsigned* p = vp + (long long)i * words_per;
    if (what == GM_REDUCE_REACHABLE) v = p[0] < 0xffffffffu ? 1.0 : 0.0;
    else if (what == GM_REDUCE_BUCKET_NOT_EMPTY) {
      int b = (int)p[1];
      v = (b >= param && b < 0x7fffffff) ? 1.0 : 0.0;
    } else {
      v = *reinterpret_cast<const double*>(p + words_per - 2);
    }
  }
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(out, v);
}
extern "C" int gm_graph_reduce(const gm_graph* g, int what, int param, double* result) {
  gm_graph* m = const_cast<gm_graph*>(g);
  if (staging(m, 8)) return 1;
  CK(cudaMemsetAsync(m->staging, 0, 8, g->stream));
  k_reduce<<<nblk(g->n_local), 256, 0, g->stream>>>((const unsigned*)g->vp, g->n_local, g->sizeof_V / 4, what, param,
                                                   (double*)m->staging);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(result, m->staging, 8, cudaMemcpyDeviceToHost, g->stream));
  CK(cudaStreamSynchronize(g->stream));
  return 0;
}
