// gather_paths.cu -- which hardware path serves random 4-byte gathers fastest on B200?
// The SpMSpV pass is limited by the L1TEX wavefront rate (one distinct 128-byte line per cycle per
// SM for a divergent LDG, hit or miss) and behind it by the L2 sector rate.  This measures, per SM
// and chip-wide, the gather rate of
//   ldg      ld.global.nc (what the engine uses)
//   tex      tex1Dfetch on a linear texture object over the same table
//   tma      one 16-byte cp.async.bulk global->shared per gather, completion on an mbarrier
//   ldg+tex  half the gathers through each path (are the two paths additive?)
//   ldg+tma  same for LDG and the bulk-copy engine
// for an L2-resident table (4 MB) and a mostly-DRAM table (256 MB).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_paths gather_paths.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 4; }

constexpr int U = 8;

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk16(void* dst, const void* src, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
               ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// MODE 0 ldg, 1 tex, 2 tma, 3 ldg+tex, 4 ldg+tma
template <int MODE>
__global__ void __launch_bounds__(1024) k(const float* __restrict__ table, cudaTextureObject_t tex, unsigned mask, int iters, float* out) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smraw);                       // one per warp
  float4* stage = reinterpret_cast<float4*>(smraw + 8 * 32) + (size_t)warp * 32 * U;  // 32 lanes x U slots x 16 B
  if (MODE == 2 || MODE == 4) {
    if (lane == 0) mbar_init(&bars[warp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
  }
  (void)nwarps;
  unsigned s = blockIdx.x * 9781u + threadIdx.x * 6271u + 1u;
  float acc = 0.f;
  unsigned parity = 0;
  for (int it = 0; it < iters; it++) {
    float v[U];
    unsigned idx[U];
#pragma unroll
    for (int u = 0; u < U; u++) idx[u] = lcg(s) & mask;
    if (MODE == 0) {
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = __ldg(table + idx[u]);
    } else if (MODE == 1) {
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = tex1Dfetch<float>(tex, (int)idx[u]);
    } else if (MODE == 3) {
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = (u & 1) ? tex1Dfetch<float>(tex, (int)idx[u]) : __ldg(table + idx[u]);
    } else {
      constexpr int NT = MODE == 2 ? U : U / 2;  // gathers through the bulk-copy engine
      if (lane == 0) mbar_expect(&bars[warp], 32 * NT * 16);
      __syncwarp();
#pragma unroll
      for (int u = 0; u < NT; u++) bulk16(&stage[u * 32 + lane], table + (idx[u] & ~3u), &bars[warp]);
      if (MODE == 4) {
#pragma unroll
        for (int u = NT; u < U; u++) v[u] = __ldg(table + idx[u]);
      }
      mbar_wait(&bars[warp], parity);
      parity ^= 1;
#pragma unroll
      for (int u = 0; u < NT; u++) v[u] = reinterpret_cast<const float*>(&stage[u * 32 + lane])[idx[u] & 3];
      __syncwarp();
    }
#pragma unroll
    for (int u = 0; u < U; u++) acc += v[u];
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
void run(const char* name, int threads, int ctas_per_sm, const float* table, cudaTextureObject_t tex, unsigned n, float* out) {
  int dev, sms, mhz;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, dev);
  const int iters = 1000;
  size_t smem = 8 * 32 + ((MODE == 2 || MODE == 4) ? (size_t)(threads / 32) * 32 * U * 16 : 0);
  CHECK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = sms * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    k<MODE><<<grid, threads, smem>>>(table, tex, n - 1, iters, out);
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    CHECK(cudaGetLastError());
    cudaEventElapsedTime(&ms, e0, e1);
  }
  double gathers = (double)grid * threads * iters * U;
  double per_ns_sm = gathers / (ms * 1e6) / sms;
  printf("%-10s table=%4u MB threads=%4d x%d/SM  %.3f ms  %.2f gathers/cycle/SM (%.2f GHz)  chip %.1f G gathers/s\n", name,
         n / (1u << 18), threads, ctas_per_sm, ms, per_ns_sm / (mhz * 1e-6), mhz * 1e-6, gathers / (ms * 1e6));
  fflush(stdout);
}

int main() {
  float* out;
  CHECK(cudaMalloc(&out, 4));
  for (unsigned n : {1u << 20, 1u << 26}) {
    float* table;
    CHECK(cudaMalloc(&table, (size_t)n * 4));
    CHECK(cudaMemset(table, 0, (size_t)n * 4));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = table;
    rd.res.linear.desc = cudaCreateChannelDesc<float>();
    rd.res.linear.sizeInBytes = (size_t)n * 4;
    cudaTextureDesc td = {};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    cudaError_t te = cudaCreateTextureObject(&tex, &rd, &td, nullptr);
    if (te != cudaSuccess) { printf("texture object over %u elements: %s\n", n, cudaGetErrorString(te)); cudaGetLastError(); tex = 0; }
    for (int cps : {1, 2}) {
      run<0>("ldg", 1024, cps, table, tex, n, out);
      if (tex) run<1>("tex", 1024, cps, table, tex, n, out);
      run<2>("tma16", 1024, cps, table, tex, n, out);
      if (tex) run<3>("ldg+tex", 1024, cps, table, tex, n, out);
      run<4>("ldg+tma16", 1024, cps, table, tex, n, out);
    }
    if (tex) cudaDestroyTextureObject(tex);
    cudaFree(table);
  }
  return 0;
}
