// gather_skew.cu -- upper bound for the SpMSpV gather pattern of PageRank on RMAT-26, without any
// of the row bookkeeping: a coalesced index stream (4 B per entry) drives 4-byte gathers into a
// 256 MB vector whose entries are stored hottest-first, the indices drawn with RMAT popularity
// (each of 26 id bits set with probability 1/4; rank = position in the popularity order).
//   ldg        every gather is ld.global.nc
//   hot(K)     gathers with index < K come from a shared-memory copy of x[0..K), the rest ld.global.nc
// Reports gathers per cycle per SM and the time 1.07 G gathers (one RMAT-26 pass) would take.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_skew gather_skew.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int SCALE = 26;

__device__ __forceinline__ unsigned mix(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return (unsigned)((z ^ (z >> 31)) >> 16);
}

__global__ void k_make_idx(const int* __restrict__ rank, int* idx, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned v = mix(2 * i) & mix(2 * i + 1) & ((1u << SCALE) - 1);
  idx[i] = rank[v];
}

__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <int POL>
__device__ __forceinline__ float ld_pol(const float* p) {
  float v;
  if (POL == 1) asm volatile("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (POL == 2) asm volatile("ld.global.nc.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (POL == 3) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (POL == 4) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (POL == 5) asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (POL == 6) asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (POL == 8) v = 0.f;  // lane predicated off: is an inactive lane free?
  else if (POL == 7) asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else v = __ldg(p);
  return v;
}

// split policy: indices below hot_n use HP, the rest CP (no shared memory)
template <int HP, int CP, int U>
__global__ void __launch_bounds__(1024) k_policy(const int* __restrict__ idx, long long n_chunks, const float* __restrict__ x,
                                                 int hot_n, int chunks_per_grab, unsigned long long* counter, float* out) {
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  while (true) {
    unsigned long long c0 = 0;
    if (lane == 0) c0 = atomicAdd(counter, (unsigned long long)chunks_per_grab);
    c0 = __shfl_sync(0xffffffffu, c0, 0);
    if ((long long)c0 >= n_chunks) break;
    long long c1 = c0 + chunks_per_grab;
    if (c1 > n_chunks) c1 = n_chunks;
    const int* p = idx + c0 * (32 * U) + lane;
    for (long long ch = c0; ch < c1; ch++) {
      int c[U];
#pragma unroll
      for (int u = 0; u < U; u++) c[u] = ld_stream(p + u * 32);
      p += 32 * U;
      float v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (c[u] < hot_n) v[u] = ld_pol<HP>(x + c[u]);
        else v[u] = ld_pol<CP>(x + c[u]);
      }
#pragma unroll
      for (int u = 0; u < U; u++) acc += v[u];
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int HP, int CP, int U>
void runp(const char* name, int threads, int blocks_per_sm, const int* idx, long long n_idx, const float* x, int hot_n,
          unsigned long long* counter, float* out) {
  int dev, sms, mhz;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, dev);
  long long n_chunks = n_idx / (32 * U);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; rep++) {
    CHECK(cudaMemset(counter, 0, 8));
    cudaEventRecord(e0);
    k_policy<HP, CP, U><<<sms * blocks_per_sm, threads>>>(idx, n_chunks, x, hot_n, 4, counter, out);
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    CHECK(cudaGetLastError());
    cudaEventElapsedTime(&ms, e0, e1);
  }
  double g = (double)n_chunks * 32 * U;
  printf("%-34s hot=%6d U=%2d threads=%4d x%d/SM  %.3f ms  %.2f gathers/cycle/SM  -> RMAT-26 pass %.2f ms\n", name, hot_n, U,
         threads, blocks_per_sm, ms, g / (ms * 1e-3) / sms / (mhz * 1e3), 1.0737e9 / (g / ms));
  fflush(stdout);
}

// each warp takes chunks of 32*U consecutive indices (lane-strided), handed out by an atomic counter
template <bool HOT, int U>
__global__ void __launch_bounds__(1024) k_gather(const int* __restrict__ idx, long long n_chunks, const float* __restrict__ x,
                                                 int hot_n, int chunks_per_grab, unsigned long long* counter, float* out) {
  extern __shared__ __align__(16) float hx[];
  if (HOT) {
    for (int i = threadIdx.x; i < hot_n / 4; i += blockDim.x)
      reinterpret_cast<float4*>(hx)[i] = __ldg(reinterpret_cast<const float4*>(x) + i);
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  while (true) {
    unsigned long long c0 = 0;
    if (lane == 0) c0 = atomicAdd(counter, (unsigned long long)chunks_per_grab);
    c0 = __shfl_sync(0xffffffffu, c0, 0);
    if ((long long)c0 >= n_chunks) break;
    long long c1 = c0 + chunks_per_grab;
    if (c1 > n_chunks) c1 = n_chunks;
    const int* p = idx + c0 * (32 * U) + lane;
    int c[U], cn[U];
#pragma unroll
    for (int u = 0; u < U; u++) cn[u] = ld_stream(p + u * 32);
    for (long long ch = c0; ch < c1; ch++) {
#pragma unroll
      for (int u = 0; u < U; u++) c[u] = cn[u];
      p += 32 * U;
      float v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (HOT && c[u] < hot_n) v[u] = hx[c[u]];
        else v[u] = __ldg(x + c[u]);
      }
      if (ch + 1 < c1) {
#pragma unroll
        for (int u = 0; u < U; u++) cn[u] = ld_stream(p + u * 32);
      }
#pragma unroll
      for (int u = 0; u < U; u++) acc += v[u];
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <bool HOT, int U>
void run(const char* name, int threads, int blocks_per_sm, const int* idx, long long n_idx, const float* x, int hot_n,
         unsigned long long* counter, float* out) {
  int dev, sms, mhz;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, dev);
  size_t smem = HOT ? (size_t)hot_n * 4 : 0;
  CHECK(cudaFuncSetAttribute(k_gather<HOT, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long n_chunks = n_idx / (32 * U);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; rep++) {
    CHECK(cudaMemset(counter, 0, 8));
    cudaEventRecord(e0);
    k_gather<HOT, U><<<sms * blocks_per_sm, threads, smem>>>(idx, n_chunks, x, hot_n, 4, counter, out);
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    CHECK(cudaGetLastError());
    cudaEventElapsedTime(&ms, e0, e1);
  }
  double g = (double)n_chunks * 32 * U;
  printf("%-10s hot=%6d U=%2d threads=%4d x%d/SM  %.3f ms  %.2f gathers/cycle/SM  chip %.1f G/s  -> RMAT-26 pass (1.074 G) %.2f ms\n",
         name, HOT ? hot_n : 0, U, threads, blocks_per_sm, ms, g / (ms * 1e-3) / sms / (mhz * 1e3), g / (ms * 1e6),
         1.0737e9 / (g / ms));
  fflush(stdout);
}

int main() {
  const unsigned n = 1u << SCALE;
  std::vector<int> rank(n);
  {
    long long off[SCALE + 2] = {0};
    for (unsigned v = 0; v < n; v++) off[__builtin_popcount(v) + 1]++;
    for (int k = 0; k <= SCALE; k++) off[k + 1] += off[k];
    for (unsigned v = 0; v < n; v++) rank[v] = (int)off[__builtin_popcount(v)]++;
  }
  int* d_rank;
  CHECK(cudaMalloc(&d_rank, (size_t)n * 4));
  CHECK(cudaMemcpy(d_rank, rank.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  const long long n_idx = 1ll << 29;  // 2 GB of indices: larger than L2, half an RMAT-26 pass
  int* idx;
  CHECK(cudaMalloc(&idx, (size_t)n_idx * 4));
  k_make_idx<<<(unsigned)((n_idx + 255) / 256), 256>>>(d_rank, idx, n_idx);
  CHECK(cudaDeviceSynchronize());
  float *x, *out;
  unsigned long long* counter;
  CHECK(cudaMalloc(&x, (size_t)n * 4));
  CHECK(cudaMemset(x, 0, (size_t)n * 4));
  CHECK(cudaMalloc(&out, 4));
  CHECK(cudaMalloc(&counter, 8));
  run<false, 8>("ldg", 1024, 1, idx, n_idx, x, 0, counter, out);
  run<false, 8>("ldg", 1024, 2, idx, n_idx, x, 0, counter, out);
  run<false, 16>("ldg", 1024, 1, idx, n_idx, x, 0, counter, out);
  run<false, 16>("ldg", 512, 2, idx, n_idx, x, 0, counter, out);
  run<false, 16>("ldg", 256, 4, idx, n_idx, x, 0, counter, out);
  if (getenv("GS_POLICY")) {
    for (int K : {49152, 262144, 4194304, 33554432})
      runp<8, 0, 16>("hot lanes OFF / cold ldg", 1024, 2, idx, n_idx, x, K, counter, out);
    for (int K : {16384, 57344, 262144}) {
      runp<0, 0, 16>("hot ldg / cold ldg", 1024, 2, idx, n_idx, x, K, counter, out);
      runp<1, 2, 16>("hot evict_last / cold evict_first", 1024, 2, idx, n_idx, x, K, counter, out);
      runp<0, 2, 16>("hot ldg / cold evict_first", 1024, 2, idx, n_idx, x, K, counter, out);
      runp<0, 3, 16>("hot ldg / cold nc.no_allocate", 1024, 2, idx, n_idx, x, K, counter, out);
      runp<0, 4, 16>("hot ldg / cold cg", 1024, 2, idx, n_idx, x, K, counter, out);
      runp<0, 5, 16>("hot ldg / cold relaxed.gpu", 1024, 2, idx, n_idx, x, K, counter, out);
      runp<0, 6, 16>("hot ldg / cold ld.no_allocate", 1024, 2, idx, n_idx, x, K, counter, out);
      runp<7, 4, 16>("hot ca / cold cg", 1024, 2, idx, n_idx, x, K, counter, out);
    }
    return 0;
  }
  run<true, 8>("hot", 1024, 1, idx, n_idx, x, 48 * 1024, counter, out);
  run<true, 16>("hot", 1024, 1, idx, n_idx, x, 48 * 1024, counter, out);
  run<true, 16>("hot", 1024, 1, idx, n_idx, x, 32 * 1024, counter, out);
  run<true, 16>("hot", 1024, 1, idx, n_idx, x, 16 * 1024, counter, out);
  run<true, 16>("hot", 1024, 2, idx, n_idx, x, 24 * 1024, counter, out);
  run<true, 8>("hot", 1024, 2, idx, n_idx, x, 24 * 1024, counter, out);
  run<true, 16>("hot", 512, 1, idx, n_idx, x, 48 * 1024, counter, out);
  return 0;
}
