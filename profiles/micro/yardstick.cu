// yardstick.cu -- external yardsticks for the PageRank SpMSpV pass on the IDENTICAL RMAT-26 matrix
// (VERDICT r1, item 5a): how fast do kernels run that are NOT bound to the reference's fold order?
//
//   cusparse   cusparseSpMV, CSR fp32, y = A^T-operand * x (rows = edge destinations, as the engine's AT),
//              algorithms DEFAULT / ALG1 / ALG2, native-id columns
//   cusparse_p the same with rows and columns renumbered by decreasing in-degree (the engine's placement)
//   coo_atomic one thread per edge in GENERATION order: atomicAdd(&y[dst], x[src])   (unordered scatter)
//   coo_sorted the same with the edges sorted by destination (consecutive threads hit the same y)
//   csr_warp   plain CSR-vector: a warp per 32 rows, one row per lane (no ordering guarantee needed, no
//              sliced-ELL layout): what the gather costs without any of the engine's bookkeeping
//
// All of them move at least nnz gathers of 4 bytes from a 256 MB vector; none reproduces the reference's
// per-row left fold (atomics and cuSPARSE reorder it).  Output: ms per pass and GTEPS, to set beside the
// engine's pass time (bench.py `roofline.ms_per_launch`).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o yardstick yardstick.cu -lcusparse
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <cusparse.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#define CHECKS(x) do { cusparseStatus_t s_ = (x); if (s_ != CUSPARSE_STATUS_SUCCESS) { printf("%s: cusparse status %d\n", #x, (int)s_); exit(1); } } while (0)

__host__ __device__ static inline unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// the generator of the engine (gm_core.cu: rmat_edge), seed 1: 0-based ids
__global__ void k_rmat(int scale, unsigned long long seed, long long nnz, int* src, int* dst) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  unsigned si = 0, di = 0;
  unsigned long long h = 0;
  for (int l = 0; l < scale; l++) {
    if ((l & 1) == 0) h = splitmix64(seed * 0x100000001B3ull + (unsigned long long)e * 32ull + (unsigned long long)(l >> 1));
    unsigned r = (l & 1) ? (unsigned)(h >> 32) : (unsigned)h;
    unsigned sb = r >= 3264175145u;
    unsigned db = (r >= 2448131358u && r < 3264175145u) || r >= 4080218931u;
    si = (si << 1) | sb;
    di = (di << 1) | db;
  }
  src[e] = (int)si;
  dst[e] = (int)di;
}
__global__ void k_hist(const int* ids, long long nnz, int* cnt) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nnz) atomicAdd(cnt + ids[e], 1);
}
__global__ void k_keys(const int* deg, int n, unsigned long long* keys) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) keys[v] = ((unsigned long long)(0xffffffffu - (unsigned)deg[v]) << 32) | (unsigned)v;
}
__global__ void k_rank(const unsigned long long* keys, int n, int* rank_of) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) rank_of[(unsigned)(keys[p] & 0xffffffffu)] = p;
}
__global__ void k_relabel(int* ids, long long nnz, const int* rank_of) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nnz) ids[e] = rank_of[ids[e]];
}
__global__ void k_pack(const int* dst, const int* src, long long nnz, unsigned long long* keys) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nnz) keys[e] = ((unsigned long long)(unsigned)dst[e] << 32) | (unsigned)src[e];
}
__global__ void k_unpack(const unsigned long long* keys, long long nnz, int* row, int* col) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nnz) { row[e] = (int)(keys[e] >> 32); col[e] = (int)(keys[e] & 0xffffffffu); }
}
__global__ void k_fill(float* x, long long n, float v) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) x[i] = v * (1.0f + (float)(i & 7) * 0.125f);
}
__global__ void __launch_bounds__(256) k_coo_atomic(const int* __restrict__ row, const int* __restrict__ col, long long nnz,
                                                    const float* __restrict__ x, float* y) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < nnz) atomicAdd(y + row[e], __ldg(x + col[e]));
}
__global__ void __launch_bounds__(256) k_csr_lane(const long long* __restrict__ ptr, const int* __restrict__ col, int n,
                                                  const float* __restrict__ x, float* __restrict__ y) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  float acc = 0.f;
  for (long long k = ptr[r]; k < ptr[r + 1]; k++) acc += __ldg(x + __ldg(col + k));
  y[r] = acc;
}
__global__ void __launch_bounds__(256) k_csr_warp(const long long* __restrict__ ptr, const int* __restrict__ col, int n,
                                                  const float* __restrict__ x, float* __restrict__ y) {
  int r = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (r >= n) return;
  float acc = 0.f;
  for (long long k = ptr[r] + lane; k < ptr[r + 1]; k += 32) acc += __ldg(x + __ldg(col + k));
  for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) y[r] = acc;
}

__global__ void k_narrow(const long long* in, long long n, int* out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)in[i];
}

static inline unsigned nb(long long n) { return (unsigned)((n + 255) / 256); }

template <class F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  f();
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(b);
  CHECK(cudaEventSynchronize(b));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const int scale = argc > 1 ? atoi(argv[1]) : 26;
  const int n = 1 << scale;
  const long long nnz = 16ll << scale;
  printf("RMAT scale-%d, n = %d, nnz = %lld (seed 1, duplicates kept)\n", scale, n, nnz);
  int *src, *dst, *row, *col;
  CHECK(cudaMalloc(&src, nnz * 4));
  CHECK(cudaMalloc(&dst, nnz * 4));
  CHECK(cudaMalloc(&row, nnz * 4));
  CHECK(cudaMalloc(&col, nnz * 4));
  k_rmat<<<nb(nnz), 256>>>(scale, 1ull, nnz, src, dst);
  float *x, *y;
  CHECK(cudaMalloc(&x, (size_t)n * 4));
  CHECK(cudaMalloc(&y, (size_t)n * 4));
  k_fill<<<nb(n), 256>>>(x, n, 1e-3f);
  CHECK(cudaDeviceSynchronize());
  auto report = [&](const char* name, float ms) {
    printf("%-44s %8.3f ms  %7.1f GTEPS\n", name, ms, nnz / (ms * 1e-3) / 1e9);
    fflush(stdout);
  };

  // ---- unordered scatter in generation order ----
  report("coo_atomic (generation order)", time_ms([&] { cudaMemsetAsync(y, 0, (size_t)n * 4); k_coo_atomic<<<nb(nnz), 256>>>(dst, src, nnz, x, y); }));

  unsigned long long *k0, *k1;
  CHECK(cudaMalloc(&k0, nnz * 8));
  CHECK(cudaMalloc(&k1, nnz * 8));
  long long* ptr;
  CHECK(cudaMalloc(&ptr, ((size_t)n + 1) * 8));
  int* deg;
  CHECK(cudaMalloc(&deg, (size_t)n * 4));
  long long* deg_ll;
  CHECK(cudaMalloc(&deg_ll, ((size_t)n + 1) * 8));

  cusparseHandle_t h;
  CHECKS(cusparseCreate(&h));
  for (int placed = 0; placed < 2; placed++) {
    if (placed) {
      // renumber rows and columns by decreasing in-degree: the engine's hot-first placement
      CHECK(cudaMemset(deg, 0, (size_t)n * 4));
      k_hist<<<nb(nnz), 256>>>(dst, nnz, deg);
      unsigned long long *v0 = k0, *v1 = k1;
      k_keys<<<nb(n), 256>>>(deg, n, v0);
      cub::DoubleBuffer<unsigned long long> kb(v0, v1);
      size_t tb = 0;
      cub::DeviceRadixSort::SortKeys(nullptr, tb, kb, n);
      void* tmp;
      CHECK(cudaMalloc(&tmp, tb));
      cub::DeviceRadixSort::SortKeys(tmp, tb, kb, n);
      int* rank_of;
      CHECK(cudaMalloc(&rank_of, (size_t)n * 4));
      k_rank<<<nb(n), 256>>>(kb.Current(), n, rank_of);
      k_relabel<<<nb(nnz), 256>>>(src, nnz, rank_of);
      k_relabel<<<nb(nnz), 256>>>(dst, nnz, rank_of);
      CHECK(cudaDeviceSynchronize());
      cudaFree(tmp);
      cudaFree(rank_of);
    }
    // CSR by destination, columns ascending
    k_pack<<<nb(nnz), 256>>>(dst, src, nnz, k0);
    {
      cub::DoubleBuffer<unsigned long long> kb(k0, k1);
      size_t tb = 0;
      cub::DeviceRadixSort::SortKeys(nullptr, tb, kb, nnz, 0, 32 + scale);
      void* tmp;
      CHECK(cudaMalloc(&tmp, tb));
      cub::DeviceRadixSort::SortKeys(tmp, tb, kb, nnz, 0, 32 + scale);
      k_unpack<<<nb(nnz), 256>>>(kb.Current(), nnz, row, col);
      CHECK(cudaDeviceSynchronize());
      cudaFree(tmp);
    }
    CHECK(cudaMemset(deg, 0, (size_t)n * 4));
    k_hist<<<nb(nnz), 256>>>(row, nnz, deg);
    {
      // 64-bit row pointers for our kernels; cuSPARSE wants offsets and column indices of one width: nnz = 2^30
      // still fits a signed 32-bit offset, so it gets a 32-bit copy (k_narrow below)
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, deg, ptr, n + 1);
      void* tmp;
      CHECK(cudaMalloc(&tmp, tb));
      // deg has n entries; scan n+1 with a trailing zero: copy into a padded buffer
      int* degp;
      CHECK(cudaMalloc(&degp, ((size_t)n + 1) * 4));
      CHECK(cudaMemset(degp, 0, ((size_t)n + 1) * 4));
      CHECK(cudaMemcpy(degp, deg, (size_t)n * 4, cudaMemcpyDeviceToDevice));
      cub::DeviceScan::ExclusiveSum(tmp, tb, degp, ptr, n + 1);
      CHECK(cudaDeviceSynchronize());
      cudaFree(tmp);
      cudaFree(degp);
    }
    const char* tag = placed ? "placement order" : "native order";
    char name[128];
    snprintf(name, sizeof name, "coo_atomic (sorted by row, %s)", tag);
    report(name, time_ms([&] { cudaMemsetAsync(y, 0, (size_t)n * 4); k_coo_atomic<<<nb(nnz), 256>>>(row, col, nnz, x, y); }));
    snprintf(name, sizeof name, "csr, one row per lane (%s)", tag);
    report(name, time_ms([&] { k_csr_lane<<<nb(n), 256>>>(ptr, col, n, x, y); }));
    snprintf(name, sizeof name, "csr, one row per warp (%s)", tag);
    report(name, time_ms([&] { k_csr_warp<<<nb((long long)n * 32), 256>>>(ptr, col, n, x, y); }, 2));

    float* vals;
    CHECK(cudaMalloc(&vals, nnz * 4));
    k_fill<<<nb(nnz), 256>>>(vals, nnz, 1.0f);
    cusparseSpMatDescr_t A;
    cusparseDnVecDescr_t vx, vy;
    int* ptr32;
    CHECK(cudaMalloc(&ptr32, ((size_t)n + 1) * 4));
    k_narrow<<<nb((long long)n + 1), 256>>>(ptr, (long long)n + 1, ptr32);
    CHECKS(cusparseCreateCsr(&A, n, n, nnz, ptr32, col, vals, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F));
    CHECKS(cusparseCreateDnVec(&vx, n, x, CUDA_R_32F));
    CHECKS(cusparseCreateDnVec(&vy, n, y, CUDA_R_32F));
    const float one = 1.f, zero = 0.f;
    cusparseSpMVAlg_t algs[3] = {CUSPARSE_SPMV_ALG_DEFAULT, CUSPARSE_SPMV_CSR_ALG1, CUSPARSE_SPMV_CSR_ALG2};
    const char* an[3] = {"DEFAULT", "CSR_ALG1", "CSR_ALG2"};
    for (int a = 0; a < 3; a++) {
      size_t bs = 0;
      cusparseStatus_t st = cusparseSpMV_bufferSize(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vx, &zero, vy, CUDA_R_32F, algs[a], &bs);
      if (st != CUSPARSE_STATUS_SUCCESS) {
        printf("cusparseSpMV %s (%s): bufferSize status %d\n", an[a], tag, (int)st);
        continue;
      }
      void* buf;
      CHECK(cudaMalloc(&buf, bs ? bs : 4));
      snprintf(name, sizeof name, "cusparseSpMV %s (%s)", an[a], tag);
      st = cusparseSpMV(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vx, &zero, vy, CUDA_R_32F, algs[a], buf);
      if (st != CUSPARSE_STATUS_SUCCESS) {
        printf("%s: status %d\n", name, (int)st);
      } else {
        report(name, time_ms([&] { cusparseSpMV(h, CUSPARSE_OPERATION_NON_TRANSPOSE, &one, A, vx, &zero, vy, CUDA_R_32F, algs[a], buf); }));
      }
      cudaFree(buf);
    }
    cusparseDestroySpMat(A);
    cusparseDestroyDnVec(vx);
    cusparseDestroyDnVec(vy);
    cudaFree(vals);
    cudaFree(ptr32);
  }
  return 0;
}
