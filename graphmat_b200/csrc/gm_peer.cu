// gm_peer.cu -- the multi-GPU exchange over PEER MEMORY (NVLink / NVSwitch), without a collective
// library call per iteration.
//
// Replaces, for a graph sharded one tile-row per GPU, the reference's MPI traffic of
//   include/GMDP/multinode/spmspv.h:61-116   (column broadcast of the message segments)
//   include/GMDP/vectors/DenseSegment.h:532-538,665-700 (dense / (index,value) wire formats)
//   include/GraphMatRuntime.h:226            (MPI_Allreduce on "converged")
// of narayanan2004/GraphMat.  Every rank maps the other ranks' message buffers into its own address
// space (CUDA IPC between processes; plain pointers between the ranks of one process, which is how the
// tests run several ranks on one GPU) and the kernels store into them directly:
//   * the fused apply+send epilogue of the SpMSpV kernels (gm_engine.cuh) writes each new message to
//     every rank as the row finishes -- the transfer overlaps the pass;
//   * k_push_x stores a rank's slice (bit words + values; only ACTIVE values for sparse frontiers);
//   * k_peer_barrier is a one-block kernel: release-store of (round, changed) into every peer's flag
//     array, acquire-spin on the own array; it orders iterations and ORs the "changed" flag.
// The host language supplies only a blocking all-gather of small host blobs for the handle exchange.
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "gm_internal.h"

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

namespace {

struct peer_blob {
  long long pid;
  unsigned long long ptr;
  int device;
  int ok;
  cudaIpcMemHandle_t handle;
};

inline unsigned nblk(long long n, int b = 256) { return (unsigned)((n + b - 1) / b); }

struct ptr_table {
  void* p[GM_MAX_WORLD];
};

// ---- barrier ---------------------------------------------------------------------------------
// flags[q] = rank q's array of 2 * GM_MAX_WORLD words; word [parity * GM_MAX_WORLD + r] is written by rank r.
// A rank cannot enter round k + 2 before everybody has left round k + 1, i.e. has read round k: two
// alternating sets of words are enough.
__global__ void k_peer_barrier(const __grid_constant__ ptr_table flags, int rank, int world, unsigned round, int* changed, int* err) {
  __shared__ int any;
  if (threadIdx.x == 0) any = 0;
  __syncthreads();
  const int q = threadIdx.x;
  const int par = (round & 1u) * GM_MAX_WORLD;
  if (q < world) {
    const unsigned long long mine = ((unsigned long long)round << 1) | ((changed && *changed) ? 1ull : 0ull);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(flags.p[q]) + par + rank;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(mine) : "memory");
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(flags.p[rank]) + par + q;
    unsigned long long t0 = 0, now = 0, w = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
      if ((w >> 1) >= round) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > 30000000000ull) {  // 30 s: a peer died; report instead of hanging the GPU
        *err = 1;
        break;
      }
      __nanosleep(100);
    }
    if (w & 1ull) atomicOr(&any, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0 && changed) *changed = any;
}

// ---- slice of x -> every peer -------------------------------------------------------------------
// words: the slice as 4-byte words, wp per vertex; dense: every vertex, else only where the bit is set
// W: unsigned (messages that are a multiple of 4 bytes) or unsigned char (e.g. TopSort's bool)
template <class W>
__global__ void __launch_bounds__(256)
    k_push_x(const __grid_constant__ ptr_table val, const __grid_constant__ ptr_table bits, int n_peers, const W* __restrict__ lval,
             const unsigned* __restrict__ lbits, long long word_off, int bit_off, int n_pad, int wp, int dense) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)n_pad * wp;
  if (t < (n_pad >> 5)) {
    const unsigned w = lbits[bit_off + t];
    for (int q = 0; q < n_peers; q++) reinterpret_cast<unsigned*>(bits.p[q])[bit_off + t] = w;
  }
  if (t >= total) return;
  const int i = (int)(t / wp);
  if (!dense && !((lbits[bit_off + (i >> 5)] >> (i & 31)) & 1u)) return;
  const W v = lval[word_off + t];
  for (int q = 0; q < n_peers; q++) reinterpret_cast<W*>(val.p[q])[word_off + t] = v;
}

__host__ __device__ inline int to_native0(int pub1, int n, int npart) {
  int v = pub1 - 1;
  int height = n / npart;
  int vmax = height * npart;
  if (v >= vmax) return v;
  int col = v % npart;
  int row = v / npart;
  return row + col * height;
}

// owned vertex properties -> the staging area of the rank whose public slice holds the vertex
__global__ void k_permute_out_slices(const unsigned* vp, const __grid_constant__ ptr_table staging, int n, int npart, const int* xidx, int n_pad,
                                     int rank, int words_per, long long per) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long pub = t / words_per;
  int w = (int)(t % words_per);
  if (pub >= n) return;
  int xi = xidx[to_native0((int)pub + 1, n, npart)];
  if (xi / n_pad != rank) return;
  unsigned* dst = reinterpret_cast<unsigned*>(staging.p[pub / per]);
  dst[pub * words_per + w] = vp[(long long)(xi % n_pad) * words_per + w];
}
__global__ void k_permute_in_all(unsigned* vp, const unsigned* in, int n, int npart, const int* xidx, int n_pad, int rank,
                                 int words_per) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long pub = t / words_per;
  int w = (int)(t % words_per);
  if (pub >= n) return;
  int xi = xidx[to_native0((int)pub + 1, n, npart)];
  if (xi / n_pad != rank) return;
  vp[(long long)(xi % n_pad) * words_per + w] = in[pub * words_per + w];
}

int host_barrier(gm_graph* g) {
  if (!g->host_gather) return 0;
  std::vector<int> all(g->world, 0);
  int mine = 1;
  if (g->host_gather(g->host_ctx, &mine, all.data(), (int)sizeof(int))) {
    gm_set_error("peer memory: the host all-gather callback failed");
    return 1;
  }
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------- symmetric buffers --
int gm_sym_alloc(gm_graph* g, size_t bytes, gm_sym* out) {
  *out = gm_sym();
  if (bytes == 0) bytes = 256;
  CK(cudaMalloc(&out->local, bytes));
  CK(cudaMemsetAsync(out->local, 0, bytes, g->stream));
  CK(cudaStreamSynchronize(g->stream));
  out->bytes = bytes;
  out->peer[g->rank] = out->local;
  if (g->world == 1) return 0;
  if (!g->host_gather) {
    cudaFree(out->local);
    *out = gm_sym();
    gm_set_error("peer memory: gm_graph_enable_peers was not called");
    return 1;
  }
  peer_blob mine;
  memset(&mine, 0, sizeof mine);
  mine.pid = (long long)getpid();
  mine.ptr = (unsigned long long)out->local;
  cudaGetDevice(&mine.device);
  std::string why;
  {
    cudaError_t e = cudaIpcGetMemHandle(&mine.handle, out->local);
    mine.ok = e == cudaSuccess ? 1 : 0;
    if (e != cudaSuccess) why = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e);
    cudaGetLastError();
  }
  std::vector<peer_blob> all(g->world);
  if (g->host_gather(g->host_ctx, &mine, all.data(), (int)sizeof(peer_blob))) {
    cudaFree(out->local);
    *out = gm_sym();
    gm_set_error("peer memory: the host all-gather callback failed");
    return 1;
  }
  int bad = 0;
  for (int q = 0; q < g->world && !bad; q++) {
    if (q == g->rank) continue;
    if (all[q].pid == mine.pid) {  // a rank of this process: its pointer is valid here
      g->peers_same_process = true;
      if (all[q].device != mine.device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(all[q].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) bad = 1;
        cudaGetLastError();
      }
      out->peer[q] = (void*)all[q].ptr;
    } else if (!all[q].ok) {
      if (why.empty()) why = "the peer could not export its buffer (cudaIpcGetMemHandle failed there)";
      bad = 1;
    } else {
      cudaError_t e = cudaIpcOpenMemHandle(&out->peer[q], all[q].handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        why = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
        cudaGetLastError();
        bad = 1;
      } else {
        out->opened[q] = true;
      }
    }
  }
  // agree on the outcome: a rank that could not map a peer must not leave the others storing into it
  std::vector<int> oks(g->world, 0);
  int ok = bad ? 0 : 1;
  const bool gathered = g->host_gather(g->host_ctx, &ok, oks.data(), (int)sizeof(int)) == 0;
  if (!gathered) why = "the host all-gather callback failed";
  for (int q = 0; q < g->world; q++)
    if (!gathered || !oks[q]) bad = 1;
  if (bad) {
    for (int q = 0; q < g->world; q++)
      if (out->opened[q]) cudaIpcCloseMemHandle(out->peer[q]);
    if (gathered) host_barrier(g);  // the importers have closed before anybody frees
    cudaFree(out->local);
    *out = gm_sym();
    gm_set_error("peer memory: a peer buffer could not be mapped (" + (why.empty() ? std::string("another rank failed") : why) + ")");
    return 1;
  }
  return 0;
}

int gm_sym_free(gm_graph* g, gm_sym* s) {
  if (!s->local) return 0;
  if (g && g->stream) cudaStreamSynchronize(g->stream);
  bool any = false;
  for (int q = 0; q < GM_MAX_WORLD; q++)
    if (s->opened[q]) {
      cudaIpcCloseMemHandle(s->peer[q]);
      any = true;
    }
  // nobody frees a buffer another PROCESS still has mapped (cudaFree before the importer's close is undefined):
  // collective there.  Ranks of one process hold plain pointers; their caller destroys them after all work is done.
  if (g && g->world > 1 && g->host_gather && any) host_barrier(g);
  cudaFree(s->local);
  *s = gm_sym();
  return 0;
}

extern "C" int gm_graph_enable_peers(gm_graph* g, gm_allgather_host_fn allgather_host, void* ctx) {
  if (g->world == 1) return 0;
  if (g->world > GM_MAX_WORLD) {
    gm_set_error("peer memory: world > GM_MAX_WORLD");
    return 1;
  }
  if (g->peers_on) return 0;
  g->host_gather = allgather_host;
  g->host_ctx = ctx;
  if (gm_sym_alloc(g, 2 * GM_MAX_WORLD * sizeof(unsigned long long), &g->sync)) {
    g->host_gather = nullptr;
    return 1;
  }
  g->barrier_round = 0;
  g->peers_on = true;
  return 0;
}
extern "C" int gm_graph_peers_enabled(const gm_graph* g) { return g->peers_on ? 1 : 0; }
// Destroying symmetric buffers is collective (nobody frees what another process still has mapped).  A process that is
// going down alone (an exception, interpreter shutdown) calls this first: its frees then skip the rendezvous instead of
// waiting for ranks that will never come.
extern "C" int gm_graph_detach_host(gm_graph* g) {
  g->host_gather = nullptr;
  g->host_ctx = nullptr;
  return 0;
}

extern "C" int gm_graph_peer_barrier(gm_graph* g, int or_changed_flag) {
  if (g->world == 1) return 0;
  if (!g->peers_on) {
    gm_set_error("gm_graph_peer_barrier: peers are not enabled");
    return 1;
  }
  if (g->peers_same_process) {
    // Ranks of ONE process share one device context: a device memory allocation or a shared hardware queue
    // between the streams makes "kernel of rank B waits behind the spinning kernel of rank A" possible (implicit
    // synchronization, CUDA programming guide 3.2.8.5.3), which would deadlock the device-side barrier.  The
    // single-process harness (tests) therefore meets on the host; one process per GPU uses the kernel below.
    CK(cudaStreamSynchronize(g->stream));
    if (g->aux_stream) CK(cudaStreamSynchronize(g->aux_stream));
    int mine = 0;
    if (or_changed_flag) {
      CK(cudaMemcpyAsync(g->h_flags + 14, g->d_flags, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
      CK(cudaStreamSynchronize(g->stream));
      mine = g->h_flags[14];
    }
    std::vector<int> all(g->world, 0);
    if (g->host_gather(g->host_ctx, &mine, all.data(), (int)sizeof(int))) {
      gm_set_error("peer memory: the host all-gather callback failed");
      return 1;
    }
    if (or_changed_flag) {
      int any = 0;
      for (int q = 0; q < g->world; q++) any |= all[q];
      g->h_flags[14] = any;
      CK(cudaMemcpyAsync(g->d_flags, g->h_flags + 14, sizeof(int), cudaMemcpyHostToDevice, g->stream));
      CK(cudaStreamSynchronize(g->stream));
    }
    return 0;
  }
  ptr_table t;
  for (int q = 0; q < GM_MAX_WORLD; q++) t.p[q] = q < g->world ? g->sync.peer[q] : nullptr;
  g->barrier_round++;
  k_peer_barrier<<<1, 32, 0, g->stream>>>(t, g->rank, g->world, g->barrier_round, or_changed_flag ? g->d_flags : nullptr,
                                          g->h_flags + 15);
  CK(cudaGetLastError());
  return 0;
}

extern "C" int gm_graph_push_x(gm_graph* g, gm_vectors* v, int dense) {
  if (g->world == 1) return 0;
  if (!v->sym) {
    gm_set_error("gm_graph_push_x: these vectors were created before gm_graph_enable_peers");
    return 1;
  }
  ptr_table pv, pb;
  int np = 0;
  for (int q = 0; q < g->world; q++)
    if (q != g->rank) {
      pv.p[np] = v->s_val.peer[q];
      pb.p[np] = v->s_bits.peer[q];
      np++;
    }
  if (v->sizeof_T % 4 == 0) {
    const int wp = v->sizeof_T / 4;
    const long long total = (long long)g->n_pad * wp;
    k_push_x<unsigned><<<nblk(total), 256, 0, g->stream>>>(pv, pb, np, (const unsigned*)v->x_val, v->x_bits,
                                                           (long long)g->rank * g->n_pad * wp, g->rank * (g->n_pad >> 5),
                                                           g->n_pad, wp, dense);
  } else {
    const int wp = v->sizeof_T;
    const long long total = (long long)g->n_pad * wp;
    k_push_x<unsigned char><<<nblk(total), 256, 0, g->stream>>>(pv, pb, np, (const unsigned char*)v->x_val, v->x_bits,
                                                                (long long)g->rank * g->n_pad * wp, g->rank * (g->n_pad >> 5),
                                                                g->n_pad, wp, dense);
  }
  CK(cudaGetLastError());
  return 0;
}

int gm_peer_copy_slice(gm_graph* g, gm_sym* s, size_t offset, size_t bytes) {
  if (bytes == 0) return 0;
  for (int q = 0; q < g->world; q++)
    if (q != g->rank)
      CK(cudaMemcpyAsync((char*)s->peer[q] + offset, (const char*)s->local + offset, bytes, cudaMemcpyDeviceToDevice, g->stream));
  return 0;
}

// ------------------------------------------------- distributed vertex-property accessors --
extern "C" long long gm_graph_slice_begin(const gm_graph* g, int rank) {
  const long long per = ((long long)g->n + g->world - 1) / g->world;
  const long long b = per * rank;
  return b < g->n ? b : g->n;
}

static int slice_staging(gm_graph* g) {
  const size_t bytes = (size_t)g->n * g->sizeof_V;
  if (g->sym_staging.local && g->sym_staging.bytes >= bytes) return 0;
  if (g->sym_staging.local && gm_sym_free(g, &g->sym_staging)) return 1;
  return gm_sym_alloc(g, bytes, &g->sym_staging);
}

extern "C" int gm_graph_set_vertexproperties_slice(gm_graph* g, const void* slice_values) {
  if (g->world == 1) return gm_graph_set_vertexproperties(g, slice_values);
  if (!g->peers_on) {
    gm_set_error("gm_graph_set_vertexproperties_slice needs gm_graph_enable_peers");
    return 1;
  }
  if (slice_staging(g)) return 1;
  const long long lo = gm_graph_slice_begin(g, g->rank), hi = gm_graph_slice_begin(g, g->rank + 1);
  const size_t off = (size_t)lo * g->sizeof_V, bytes = (size_t)(hi - lo) * g->sizeof_V;
  if (gm_graph_peer_barrier(g, 0)) return 1;  // nobody still reads the staging area of an earlier call
  if (bytes) CK(cudaMemcpyAsync((char*)g->sym_staging.local + off, slice_values, bytes, cudaMemcpyHostToDevice, g->stream));
  if (gm_peer_copy_slice(g, &g->sym_staging, off, bytes)) return 1;
  if (gm_graph_peer_barrier(g, 0)) return 1;
  const int wp = g->sizeof_V / 4;
  const long long t = (long long)g->n * wp;
  k_permute_in_all<<<nblk(t), 256, 0, g->stream>>>((unsigned*)g->vp, (const unsigned*)g->sym_staging.local, g->n,
                                                  g->ref_threads * 16, g->d_xidx, g->n_pad, g->rank, wp);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(g->stream));
  return 0;
}

extern "C" int gm_graph_get_vertexproperties_slice(gm_graph* g, void* slice_values) {
  if (g->world == 1) return gm_graph_get_vertexproperties(g, slice_values);
  if (!g->peers_on) {
    gm_set_error("gm_graph_get_vertexproperties_slice needs gm_graph_enable_peers");
    return 1;
  }
  if (slice_staging(g)) return 1;
  const long long lo = gm_graph_slice_begin(g, g->rank), hi = gm_graph_slice_begin(g, g->rank + 1);
  const size_t off = (size_t)lo * g->sizeof_V, bytes = (size_t)(hi - lo) * g->sizeof_V;
  if (gm_graph_peer_barrier(g, 0)) return 1;
  ptr_table st;
  for (int q = 0; q < GM_MAX_WORLD; q++) st.p[q] = q < g->world ? g->sym_staging.peer[q] : nullptr;
  const int wp = g->sizeof_V / 4;
  const long long t = (long long)g->n * wp;
  const long long per = ((long long)g->n + g->world - 1) / g->world;
  k_permute_out_slices<<<nblk(t), 256, 0, g->stream>>>((const unsigned*)g->vp, st, g->n, g->ref_threads * 16, g->d_xidx,
                                                      g->n_pad, g->rank, wp, per);
  CK(cudaGetLastError());
  if (gm_graph_peer_barrier(g, 0)) return 1;
  if (bytes) CK(cudaMemcpyAsync(slice_values, (const char*)g->sym_staging.local + off, bytes, cudaMemcpyDeviceToHost, g->stream));
  CK(cudaStreamSynchronize(g->stream));
  if (g->h_flags[15]) {
    gm_set_error("peer memory: a peer did not reach the barrier (timeout)");
    return 1;
  }
  return 0;
}
