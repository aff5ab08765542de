// gm_builtin.cu -- the five vertex programs BASELINE.json names, instantiated over the
// device engine and exported through the C ABI (gm_run_program / gm_step_*).
// The programs themselves are in graphmat_b200/include/GraphMat/programs/ and cite the
// reference app lines they restate (narayanan2004/GraphMat src/{PageRank,BFS,SSSP,
// DeltaStepping,SGD}.cpp).
#include <cstring>

#include "GraphMat/gm_engine.cuh"
#include "GraphMat/programs/BFS.h"
#include "GraphMat/gm_vertex_ops.cuh"
#include "GraphMat/programs/IncrementalPageRank.h"
#include "GraphMat/programs/LDA.h"
#include "GraphMat/programs/TopologicalSort.h"
#include "GraphMat/programs/PageRank.h"
#include "GraphMat/programs/SGD.h"
#include "GraphMat/programs/SSSP.h"
#include "gm_internal.h"

namespace {

// state block <-> program object
template <class P> struct binder;
template <> struct binder<Degree<PR, int> > {
  static int bytes() { return 0; }
  static void in(Degree<PR, int>&, const void*) {}
  static void out(const Degree<PR, int>&, void*) {}
};
template <> struct binder<PageRank<int> > {
  static int bytes() { return sizeof(gm_pagerank_state); }
  static void in(PageRank<int>& p, const void* s) { if (s) p.alpha = ((const gm_pagerank_state*)s)->alpha; }
  static void out(const PageRank<int>&, void*) {}
};
template <> struct binder<BFS2> {
  static int bytes() { return sizeof(gm_bfs_state); }
  static void in(BFS2& p, const void* s) { if (s) p.current_depth = ((const gm_bfs_state*)s)->current_depth; }
  static void out(const BFS2& p, void* s) { if (s) ((gm_bfs_state*)s)->current_depth = p.current_depth; }
};
template <> struct binder<SSSP<int> > {
  static int bytes() { return 0; }
  static void in(SSSP<int>&, const void*) {}
  static void out(const SSSP<int>&, void*) {}
};
template <> struct binder<DeltaStepping> {
  static int bytes() { return sizeof(gm_deltastepping_state); }
  static void in(DeltaStepping& p, const void* s) {
    if (s) { p.delta = ((const gm_deltastepping_state*)s)->delta; p.bid = ((const gm_deltastepping_state*)s)->bid; }
  }
  static void out(const DeltaStepping& p, void* s) {
    if (s) { ((gm_deltastepping_state*)s)->delta = p.delta; ((gm_deltastepping_state*)s)->bid = p.bid; }
  }
};
template <unsigned K> struct binder<SGDProgram<K> > {
  static int bytes() { return sizeof(gm_sgd_state); }
  static void in(SGDProgram<K>& p, const void* s) {
    if (s) { p.lambda = ((const gm_sgd_state*)s)->lambda; p.step = ((const gm_sgd_state*)s)->step; }
  }
  static void out(const SGDProgram<K>&, void*) {}
};
template <unsigned K> struct binder<RMSEProgram<K> > {
  static int bytes() { return 0; }
  static void in(RMSEProgram<K>&, const void*) {}
  static void out(const RMSEProgram<K>&, void*) {}
};

template <> struct binder<Degree<dPR, int> > {
  static int bytes() { return 0; }
  static void in(Degree<dPR, int>&, const void*) {}
  static void out(const Degree<dPR, int>&, void*) {}
};
template <> struct binder<DeltaPageRank> {
  static int bytes() { return sizeof(gm_deltapagerank_state); }
  static void in(DeltaPageRank& p, const void* s) {
    if (s) { p.alpha = ((const gm_deltapagerank_state*)s)->alpha; p.iter = ((const gm_deltapagerank_state*)s)->iter; }
  }
  static void out(const DeltaPageRank& p, void* s) { if (s) ((gm_deltapagerank_state*)s)->iter = p.iter; }
};
template <> struct binder<InDegree<TopSortVertex, int> > {
  static int bytes() { return 0; }
  static void in(InDegree<TopSortVertex, int>&, const void*) {}
  static void out(const InDegree<TopSortVertex, int>&, void*) {}
};
template <> struct binder<TopSort> {
  static int bytes() { return sizeof(gm_topsort_state); }
  static void in(TopSort& p, const void* s) { if (s) p.current_topsort_order = ((const gm_topsort_state*)s)->current_topsort_order; }
  static void out(const TopSort& p, void* s) { if (s) ((gm_topsort_state*)s)->current_topsort_order = p.current_topsort_order; }
};

// LDA, K = 20 (src/LDA.cpp:278): global_N is recomputed on the device through the functor reduce
void lda_recalc20(void* ctx, LDAVector<20>* out) {
  gm::map_reduce_vertices<LDAVector<20>, LDAVector<20> >((gm_graph*)ctx, out, LDAIfTerm<20>(), LDAAdd<20>());
}
template <> struct binder<LDAInitProgram<20> > {
  static int bytes() { return 0; }
  static void in(LDAInitProgram<20>&, const void*) {}
  static void out(const LDAInitProgram<20>&, void*) {}
};
template <> struct binder<LDAProgram<20> > {
  static int bytes() { return sizeof(gm_lda_state); }
  static void in(LDAProgram<20>& p, const void* s) {
    if (!s) return;
    const gm_lda_state* t = (const gm_lda_state*)s;
    p.alpha = t->alpha; p.eta = t->eta; p.vocab_size = t->vocab_size;
    for (int i = 0; i < 20; i++) p.global_N.N[i] = t->global_N[i];
  }
  static void out(const LDAProgram<20>& p, void* s) {
    if (s) for (int i = 0; i < 20; i++) ((gm_lda_state*)s)->global_N[i] = p.global_N.N[i];
  }
};
template <> struct binder<LDALLProgram<20> > {
  static int bytes() { return sizeof(gm_ldall_state); }
  static void in(LDALLProgram<20>& p, const void* s) {
    if (!s) return;
    const gm_ldall_state* t = (const gm_ldall_state*)s;
    p.eta = t->eta; p.nterms = t->nterms;
    for (int i = 0; i < 20; i++) p.N_k.N[i] = t->N_k[i] + t->nterms * (t->eta - 1.0);  // the constructor's smoothing, :209-212
  }
  static void out(const LDALLProgram<20>&, void*) {}
};
// programs that need the graph itself on the host side (LDAProgram's do_every_iteration reduces over all vertices)
template <class P> void bind_graph(P&, gm_graph*) {}
template <> void bind_graph<LDAProgram<20> >(LDAProgram<20>& p, gm_graph* g) {
  p.recalc = lda_recalc20;
  p.recalc_ctx = g;
  p.calcGlobalN();  // ldap.calcGlobalN() right after construction, src/LDA.cpp:300
}

enum { OP_RUN, OP_SEND, OP_SPMSPV, OP_APPLY, OP_SIZES };
struct call {
  int op;
  gm_graph* g;
  void* state;
  int iterations;
  gm_vectors* tmp;
  gm_run_stats* stats;
  int* changed;
  int *sT, *sU, *sV, *sS;
};

template <class P>
int dispatch(const call& c) {
  typedef gm::engine<P> EN;
  if (c.op == OP_SIZES) {
    *c.sT = (int)sizeof(typename EN::T);
    *c.sU = (int)sizeof(typename EN::U);
    *c.sV = (int)sizeof(typename EN::V);
    *c.sS = binder<P>::bytes();
    return 0;
  }
  P prog;
  binder<P>::in(prog, c.state);
  if (c.op == OP_RUN) {
    bind_graph(prog, c.g);
    int rc = EN::run(prog, c.g, c.iterations, c.tmp, c.stats);
    binder<P>::out(prog, c.state);
    return rc;
  }
  gm_graph_view gv;
  gm_vectors_view vv;
  if (gm_graph_view_get(c.g, &gv) || gm_vectors_view_get(c.tmp, &vv) || EN::check(gv, vv)) return 1;
  cudaStream_t st = (cudaStream_t)gv.stream;
  if (c.op == OP_SEND) {
    if (EN::send(prog, gv, vv, nullptr)) return 1;
    if (gv.world > 1 && gm_graph_exchange_x(c.g, c.tmp)) return 1;
  } else if (c.op == OP_SPMSPV) {
    if (EN::spmspv(prog, gv, vv, false, nullptr, c.tmp)) return 1;
  } else {
    if (cudaMemsetAsync(gv.d_flags, 0, sizeof(int), st) != cudaSuccess) return 1;
    if (EN::apply(prog, gv, vv, nullptr)) return 1;
    if (cudaMemcpyAsync(gv.h_flags, gv.d_flags, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return 1;
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) {
    gm_set_error("kernel failed");
    return 1;
  }
  if (c.op == OP_APPLY && c.changed) *c.changed = gv.h_flags[0];
  return 0;
}

int route(int program, const call& c) {
  switch (program) {
    case GM_PROG_DEGREE: return dispatch<Degree<PR, int> >(c);
    case GM_PROG_PAGERANK: return dispatch<PageRank<int> >(c);
    case GM_PROG_BFS: return dispatch<BFS2>(c);
    case GM_PROG_SSSP: return dispatch<SSSP<int> >(c);
    case GM_PROG_DELTASTEPPING: return dispatch<DeltaStepping>(c);
    case GM_PROG_SGD20: return dispatch<SGDProgram<20> >(c);
    case GM_PROG_RMSE20: return dispatch<RMSEProgram<20> >(c);
    case GM_PROG_SGD32: return dispatch<SGDProgram<32> >(c);
    case GM_PROG_RMSE32: return dispatch<RMSEProgram<32> >(c);
    case GM_PROG_SGD4: return dispatch<SGDProgram<4> >(c);
    case GM_PROG_RMSE4: return dispatch<RMSEProgram<4> >(c);
    case GM_PROG_DEGREE_DPR: return dispatch<Degree<dPR, int> >(c);
    case GM_PROG_DELTAPAGERANK: return dispatch<DeltaPageRank>(c);
    case GM_PROG_INDEGREE: return dispatch<InDegree<TopSortVertex, int> >(c);
    case GM_PROG_TOPSORT: return dispatch<TopSort>(c);
    case GM_PROG_LDAINIT20: return dispatch<LDAInitProgram<20> >(c);
    case GM_PROG_LDA20: return dispatch<LDAProgram<20> >(c);
    case GM_PROG_LDALL20: return dispatch<LDALLProgram<20> >(c);
  }
  gm_set_error("unknown program id");
  return 1;
}

}  // namespace

extern "C" int gm_program_sizes(int program, int* sizeof_T, int* sizeof_U, int* sizeof_V, int* sizeof_state) {
  int a, b, c2, d;
  call c = {OP_SIZES, nullptr, nullptr, 0, nullptr, nullptr, nullptr, sizeof_T ? sizeof_T : &a, sizeof_U ? sizeof_U : &b,
            sizeof_V ? sizeof_V : &c2, sizeof_state ? sizeof_state : &d};
  return route(program, c);
}
extern "C" int gm_run_program(gm_graph* g, int program, void* state, int iterations, gm_vectors* tmp, gm_run_stats* stats) {
  call c = {OP_RUN, g, state, iterations, tmp, stats, nullptr, nullptr, nullptr, nullptr, nullptr};
  return route(program, c);
}
extern "C" int gm_step_send(gm_graph* g, int program, const void* state, gm_vectors* tmp) {
  call c = {OP_SEND, g, const_cast<void*>(state), 0, tmp, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  return route(program, c);
}
extern "C" int gm_step_spmspv(gm_graph* g, int program, const void* state, gm_vectors* tmp) {
  call c = {OP_SPMSPV, g, const_cast<void*>(state), 0, tmp, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  return route(program, c);
}
extern "C" int gm_step_apply(gm_graph* g, int program, void* state, gm_vectors* tmp, int* changed) {
  call c = {OP_APPLY, g, state, 0, tmp, nullptr, changed, nullptr, nullptr, nullptr, nullptr};
  return route(program, c);
}

// ---- test hooks for the exact fp32 fold (gm_fadd32.cuh) ----
namespace {
// host walk through the same block structure and the same fx:: arithmetic as fx::warp_fold
void host_warp_fold(const float* v /*256*/, const unsigned* vmask /*32*/, float& s, bool& have) {
  using namespace gm::fx;
  unsigned pending = 0;
  for (int l = 0; l < 32; l++) if (vmask[l]) pending |= 1u << l;
  while (pending) {
    binade b;
    bool hot = have && binade_of(s, b);
    if (!hot) {
      int f = __builtin_ctz(pending);
      float vv[8];
      for (int k = 0; k < 8; k++) vv[k] = v[f * 8 + k];
      serial8(vv, vmask[f], s, have);
      pending &= ~(1u << f);
      continue;
    }
    qmap incl[32];
    bool over[32];
    qmap run = identity();
    for (int l = 0; l < 32; l++) {
      bool bad = false;
      qmap mine = identity();
      if ((pending >> l) & 1u)
        for (int k = 0; k < 8; k++) mine = compose(mine, quantize(v[l * 8 + k], b, bad));
      run = compose(run, mine);
      incl[l] = run;
      over[l] = bad || apply(run, b.m) >= (1u << 24);
    }
    int f = -1;
    for (int l = 0; l < 32; l++) if (over[l] && ((pending >> l) & 1u)) { f = l; break; }
    if (f < 0) {
      s = (float)apply(incl[31], b.m) * b.u;
      pending = 0;
    } else {
      unsigned m_prev = f == 0 ? b.m : apply(incl[f - 1], b.m);
      float sf = (float)m_prev * b.u;
      bool hf = true;
      float vv[8];
      for (int k = 0; k < 8; k++) vv[k] = v[f * 8 + k];
      serial8(vv, vmask[f], sf, hf);
      s = sf;
      pending &= ~((2u << f) - 1u);
    }
  }
}
}  // namespace

extern "C" int gm_debug_fold_f32_host(const float* a, long long n, float* out) {
  float s = 0.f;
  bool have = false;
  for (long long k0 = 0; k0 < n; k0 += 256) {
    float v[256];
    unsigned vm[32];
    for (int l = 0; l < 32; l++) {
      vm[l] = 0;
      for (int k = 0; k < 8; k++) {
        long long i = k0 + l * 8 + k;
        v[l * 8 + k] = i < n ? a[i] : 0.f;
        if (i < n) vm[l] |= 1u << k;
      }
    }
    host_warp_fold(v, vm, s, have);
  }
  *out = s;
  return have ? 0 : 2;
}

__global__ void k_iota(int* p, long long n) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}

// one row holding a[0..n): warps = 1 -> warp-per-row kernel, 16 -> block-per-row kernel, 32 -> TMA-streamed kernel; offset shifts the
// row start inside the index array (exercises the aligned-group masking)
extern "C" int gm_debug_fold_f32_device(const float* a, long long n, int warps, int offset, float* out) {
  typedef PageRank<int> P;
  if (n <= 0 || n >= (1ll << 31) || offset < 0 || offset > 64) { gm_set_error("bad n/offset"); return 1; }
  float *dx = nullptr, *dy = nullptr;
  int *dcol = nullptr, *dval = nullptr;
  long long* dptr = nullptr;
  unsigned* dbits = nullptr;
  long long tot = n + offset;
  cudaMalloc(&dx, (tot + (gm::GM_TMA_BUFS + 1) * gm::GM_TMA_ROUND) * 4); cudaMalloc(&dcol, (tot + 16) * 4); cudaMalloc(&dval, (tot + 16) * 4);
  cudaMalloc(&dptr, 16); cudaMalloc(&dy, 4 * 32); cudaMalloc(&dbits, 4);
  cudaMemset(dx, 0, (tot + (gm::GM_TMA_BUFS + 1) * gm::GM_TMA_ROUND) * 4);
  cudaMemcpy(dx + offset, a, n * 4, cudaMemcpyHostToDevice);
  cudaMemset(dval, 0, (tot + 16) * 4);
  k_iota<<<(unsigned)((tot + 16 + 255) / 256), 256>>>(dcol, tot + 16);
  long long hp[2] = {offset, tot};
  cudaMemcpy(dptr, hp, 16, cudaMemcpyHostToDevice);
  cudaMemset(dbits, 0, 4);
  cudaMemset(dy, 0, 4 * 32);
  gm_matrix_view M;
  memset(&M, 0, sizeof M);
  M.n_slots = 32; M.n_heavy = 1; M.identity = 1; M.h_ptr = dptr; M.h_col = dcol; M.h_val = dval;
  P prog;
  gm::prog_bytes<P> pb = gm::pack(prog);
  if (warps == 32) {  // the TMA-streamed fold of the staged longest rows: dx is its own staging array
    auto kt = gm::k_heavy_fadd32_tma<P, float, PR, true, false>;
    cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, gm::GM_TMA_BUFS * gm::GM_TMA_ROUND * 4);
    kt<<<1, gm::GM_TMA_W * 32, gm::GM_TMA_BUFS * gm::GM_TMA_ROUND * 4>>>(pb, M, 0, 1, dx, dy, dbits, gm::epilogue<float, PR>());
  } else if (warps == 1) gm::k_heavy_fadd32<P, float, PR, int, true, true, 1><<<1, 128>>>(pb, M, 0, 1, 1 << 15, dx, nullptr, dy, dbits);
  else gm::k_heavy_fadd32<P, float, PR, int, true, true, 16><<<1, 512>>>(pb, M, 0, 1, 1 << 15, dx, nullptr, dy, dbits);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned bits = 0;
  cudaMemcpy(out, dy, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&bits, dbits, 4, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dy); cudaFree(dcol); cudaFree(dval); cudaFree(dptr); cudaFree(dbits);
  if (e != cudaSuccess) { gm_set_error(cudaGetErrorString(e)); return 1; }
  return (bits & 1u) ? 0 : 2;
}
