// dsmem_gather.cu -- microbenchmark behind DESIGN.md's "hot cache" decision: how many random
// 4-byte gathers per cycle per SM can be served from (a) the SM's own shared memory, (b) the
// distributed shared memory of a thread-block cluster, (c) an L2-resident table in global memory.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dsmem_gather dsmem_gather.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 4; }

// mode 0: own smem, 1: cluster dsmem, 2: global table
template <int MODE>
__global__ void __launch_bounds__(1024) k(const float* __restrict__ table, int table_n, int per_cta, int iters, float* out) {
  extern __shared__ float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int csize = MODE == 1 ? cluster.num_blocks() : 1;
  for (int i = threadIdx.x; i < per_cta; i += blockDim.x) sm[i] = (float)(i & 1023);
  if (MODE == 1) cluster.sync(); else __syncthreads();
  unsigned s = blockIdx.x * 9781u + threadIdx.x * 6271u + 1u;
  float acc = 0.f;
  for (int it = 0; it < iters; it++) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      unsigned r = lcg(s);
      if (MODE == 0) v[u] = sm[r % per_cta];
      else if (MODE == 1) {
        unsigned idx = r % (unsigned)(per_cta * csize);
        const float* remote = cluster.map_shared_rank(sm, idx / per_cta);
        v[u] = remote[idx % per_cta];
      } else v[u] = __ldg(table + (r % (unsigned)table_n));
    }
#pragma unroll
    for (int u = 0; u < 8; u++) acc += v[u];
  }
  if (MODE == 1) cluster.sync();
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
void run(const char* name, int csize, int threads, int per_cta, const float* table, int table_n, float* out) {
  int dev, sms, mhz;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, dev);
  const int iters = 2000;
  size_t smem = (size_t)per_cta * 4;
  CHECK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (csize > 8) CHECK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  int grid = (sms / csize) * csize;
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<MODE>, table, table_n, per_cta, iters, out);
    if (e != cudaSuccess) { printf("%-28s launch failed: %s\n", name, cudaGetErrorString(e)); cudaGetLastError(); return; }
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double gathers = (double)grid * threads * iters * 8;
  double per_ns_sm = gathers / (ms * 1e6) / grid;
  printf("%-28s cluster=%2d threads=%4d table/cta=%6d KB  %.3f ms  %.2f gathers/ns/SM  (%.2f per cycle at %.2f GHz)  chip %.1f G gathers/s\n",
         name, csize, threads, per_cta * 4 / 1024, ms, per_ns_sm, per_ns_sm / (mhz * 1e-6), mhz * 1e-6, gathers / (ms * 1e6));
}

int main() {
  float *table, *out;
  const int table_n = 1 << 20;  // 4 MB: L2-resident
  CHECK(cudaMalloc(&table, (size_t)table_n * 4));
  CHECK(cudaMemset(table, 0, (size_t)table_n * 4));
  CHECK(cudaMalloc(&out, 4));
  run<0>("own shared memory", 1, 1024, 48 * 1024, table, table_n, out);
  for (int c : {2, 4, 8, 16}) run<1>("cluster distributed smem", c, 1024, 48 * 1024, table, table_n, out);
  run<1>("cluster distributed smem", 8, 512, 48 * 1024, table, table_n, out);
  run<2>("global (L2-resident 4 MB)", 1, 1024, 1024, table, table_n, out);
  float* big;
  const int big_n = 1 << 26;  // 256 MB: mostly DRAM
  CHECK(cudaMalloc(&big, (size_t)big_n * 4));
  CHECK(cudaMemset(big, 0, (size_t)big_n * 4));
  run<2>("global (256 MB table)", 1, 1024, 1024, big, big_n, out);
  return 0;
}
