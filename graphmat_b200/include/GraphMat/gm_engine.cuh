// gm_engine.cuh -- the device engine behind run_graph_program: send -> generalized
// SpMSpV -> apply, as sm_100a kernels templated on the vertex program.
//
// Replaces, for one rank's tile-row, the reference's
//   IntersectReduce      include/GMDP/singlenode/intersectreduce.h:39-66   (k_send)
//   my_spmspv / my_spmspv3  include/GMDP/singlenode/spmspv.h:39-86, spmspv3.h:38-90  (k_sell, k_heavy*)
//   the apply loop       include/GraphMatRuntime.h:184-226                 (k_apply)
// of narayanan2004/GraphMat.  Semantics kept bit for bit:
//   * x bit = active bit; send_message's bool result is ignored (GraphMatRuntime.h:79-85)
//   * per destination row the contributions are left-folded, accumulator first
//     (SPMV.h:54-59), in ascending NATIVE column id (spmspv.h:55-77) -- the matrix
//     build sorts every row that way, the kernels never reorder a row unless the
//     program opts in (gm_reorderable / gm_fadd32_exact below)
//   * apply runs only where a message arrived; "changed" is the program's operator!=
//
// Data layout: rows are stored by decreasing length.  The longest
// n_heavy rows are row-contiguous and folded by one warp each (coalesced index
// stream, gathers 32 wide); the rest are 32-row sliced-ELL, one row per lane, so a
// lane's fold is a private sequential chain and every index load is one 128-byte
// line per warp.  x lives in "placement" order (hot columns first) so gathers hit
// L1/L2; none of this changes the order in which a row is folded.
#ifndef GRAPHMAT_B200_ENGINE_CUH
#define GRAPHMAT_B200_ENGINE_CUH

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <utility>
#include <vector>
#include <algorithm>

#include "GraphProgram.h"
#include "gm_fadd32.cuh"
#include "graphmat_b200.h"

namespace gm {

// ------------------------------------------------------------------ traits --
template <class T, class U, class V, class E>
struct prog_types {
  typedef T Tm;
  typedef U Um;
  typedef V Vp;
  typedef E Ev;
};
template <class T, class U, class V, class E>
prog_types<T, U, V, E> deduce_types(const GraphMat::GraphProgram<T, U, V, E>*);

// A program may declare `static const bool gm_reorderable = true;` when its
// reduce_function is associative (min, integer +, "last writer": a = b); long rows
// are then folded as an ORDER-PRESERVING tree instead of a serial chain.
template <class P, class = void>
struct is_reorderable : std::false_type {};
template <class P>
struct is_reorderable<P, typename std::enable_if<P::gm_reorderable>::type> : std::true_type {};
// `static const bool gm_last_writer = true;`: reduce_function(a, b) is `a = b` (BFS, src/BFS.cpp:72-74).
// The left fold in ascending column order then keeps the contribution of the LAST active entry of
// the row, and process_message is const (GraphProgram.h:79), so a sparse-x pass may scan each row
// from its end and stop at the first active entry: same bits, far fewer gathers on a wide frontier.
template <class P, class = void>
struct is_last_writer : std::false_type {};
template <class P>
struct is_last_writer<P, typename std::enable_if<P::gm_last_writer>::type> : std::true_type {};
// `static const bool gm_atomic_min = true;`: U is a 32-bit unsigned integer and reduce_function(a, b) is
// a = min(a, b) (SSSP, DeltaStepping): commutative and associative, so a sparse-frontier pass may fold with
// atomicMin straight into y instead of sorting its contributions.
template <class P, class = void>
struct is_atomic_min : std::false_type {};
template <class P>
struct is_atomic_min<P, typename std::enable_if<P::gm_atomic_min>::type> : std::true_type {};
// `static bool gm_null_message(const T&)`: true for a message that cannot change any receiver (process_message
// maps it to the identity of reduce_function AND apply ignores that identity) -- DeltaStepping's MAX_DIST from the
// vertices outside the current bucket (src/DeltaStepping.cpp:79-84).  k_send then leaves the sender's x bit clear:
// the frontier shrinks to the vertices that matter and the pass can take the sparse-frontier path.  Vertex
// properties, "changed" flags and iteration counts are the same as with the message sent.
template <class P, class T, class = void>
struct has_null_message : std::false_type {};
template <class P, class T>
struct has_null_message<P, T, decltype(void(P::gm_null_message(std::declval<const T&>())))> : std::true_type {};
// `static const bool gm_fadd32_exact = true;`: T = U = float, process_message is
// res = message, reduce is a += b and messages are >= 0.  Long rows then use the
// bit-exact parallel emulation of the serial fp32 fold (k_heavy_fadd32).
template <class P, class = void>
struct is_fadd32 : std::false_type {};
template <class P>
struct is_fadd32<P, typename std::enable_if<P::gm_fadd32_exact>::type> : std::true_type {};

// The program travels to the device as raw bytes: it has a vptr, and every call
// below is qualified (P::f), so no virtual dispatch happens on the device.
template <class P>
struct prog_bytes {
  alignas(16) unsigned char b[sizeof(P)];
  __device__ __forceinline__ const P& get() const { return *reinterpret_cast<const P*>(b); }
};
template <class P>
inline prog_bytes<P> pack(const P& p) {
  prog_bytes<P> r;
  memcpy(r.b, (const void*)&p, sizeof(P));
  return r;
}

__device__ __forceinline__ bool test_bit(const unsigned* __restrict__ bits, int i) {
  return (__ldg(bits + (i >> 5)) >> (i & 31)) & 1u;
}

// ------------------------------------------------------- fused apply + send --
// ALL_VERTICES programs that do not override do_every_iteration and sweep ONE operand matrix
// (PageRank): the thread that finishes a row's fold holds the reduced message in a register, so it
// runs the apply loop body (GraphMatRuntime.h:201-213) and the next iteration's send_message
// (GraphMatRuntime.h:79-85) right there.  y never goes to memory and there is no apply kernel.
// The next message vector is a SECOND buffer (the pass still gathers from the current one); on a
// sharded graph the same store also goes to every peer's copy of that buffer over NVLink (peer
// memory), so the exchange of SURVEY 8e overlaps the pass row by row instead of following it.
constexpr int GM_MAX_PEERS = GM_MAX_WORLD - 1;
template <class T, class V>
struct epilogue {
  V* vp;                    // local vertex properties (placement order)
  T* x_next;                // next message vector, local copy (n_full entries)
  T* x_peer[GM_MAX_PEERS];  // the same buffer on the other ranks (n_peers valid entries)
  int n_peers;
  int x_off;                // rank * n_local_pad: this rank's slice of x
  int n_valid;              // local vertices that exist (the rest is padding)
  int* flag;                // "some vertex changed"
};
// returns "the vertex property changed" (the caller raises the flag: warp-aggregated where it can)
template <class P, class T, class U, class V>
__device__ __forceinline__ bool fused_apply_send(const prog_bytes<P>& pb, const epilogue<T, V>& ep, int vtx, bool have,
                                                 const U& acc) {
  if (vtx >= ep.n_valid) return false;
  alignas(16) unsigned char pbuf[sizeof(P)];
  memcpy(pbuf, pb.b, sizeof(P));
  P& prog = *reinterpret_cast<P*>(pbuf);  // apply is non-const in the reference
  V cur = ep.vp[vtx];
  bool changed = false;
  if (have) {  // apply runs only where a message arrived
    V old = cur;
    prog.P::apply(acc, cur);
    changed = (old != cur);
    ep.vp[vtx] = cur;
  }
  if (ep.x_next) {  // NULL: single-iteration run, nobody reads the next message vector
    T t;
    (void)prog.P::send_message(cur, t);  // the bool is ignored (GraphMatRuntime.h:79-85)
    const size_t i = (size_t)ep.x_off + (size_t)vtx;
    ep.x_next[i] = t;
    for (int p = 0; p < ep.n_peers; p++) ep.x_peer[p][i] = t;
  }
  return changed;
}
__device__ __forceinline__ void raise_flag(int* flag) {
  if (*((volatile int*)flag) == 0) atomicExch(flag, 1);
}


// ---- cache-hinted loads ------------------------------------------------------
// The index/edge streams are read once per pass: keep them out of L1.  The message
// vector is stored hot-columns-first, so a gather of x[c] with c < hot_limit is worth
// an L1 line (evict_last); a cold gather is not (no_allocate) -- otherwise the cold
// 90 % of the columns, which carry a third of the gathers, thrash the hot set.
template <class X>
__device__ __forceinline__ X ld_stream(const X* p) {
#if defined(GM_STREAM_NO_ALLOCATE)
  if constexpr (sizeof(X) == 4) {
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return *reinterpret_cast<X*>(&v);
  } else if constexpr (sizeof(X) == 8) {
    unsigned long long v;
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return *reinterpret_cast<X*>(&v);
  } else {
    return *p;
  }
#else
  if constexpr (sizeof(X) == 4 || sizeof(X) == 8) return __ldg(p);
  else return *p;
#endif
}
__device__ __forceinline__ int4 ld_stream4(const int* p) {
#if defined(GM_STREAM_NO_ALLOCATE)
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#else
  return __ldg(reinterpret_cast<const int4*>(p));
#endif
}
// hot_limit < 0: plain read-only load (default); >= 0: L1 hint by hotness (experimental)
template <class X>
__device__ __forceinline__ X ld_gather(const X* x, int c, int hot_limit) {
  if constexpr (sizeof(X) == 4) {
    if (hot_limit == -1) return __ldg(x + c);
#ifdef GM_EXPERIMENTS
    if (hot_limit < -1) {  // GM_HOT_LIMIT=-K: what a perfect on-SM cache of K columns would buy (WRONG results)
      if (c < -hot_limit) return __ldg(x + (c & 63));  // always an L1 hit
      return __ldg(x + c);
    }
#else
    if (hot_limit < -1) return __ldg(x + c);
#endif
    unsigned v;
    if (c < hot_limit) return __ldg(x + c);
    else asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(x + c));
    return *reinterpret_cast<X*>(&v);
  } else if constexpr (sizeof(X) == 8) {
    if (hot_limit < 0) return __ldg(x + c);
    unsigned long long v;
    if (c < hot_limit) return __ldg(x + c);
    else asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(x + c));
    return *reinterpret_cast<X*>(&v);
  } else {
    return x[c];
  }
}

// -------------------------------------------------------------------- send --
// x[i] = send_message(vp[i]) where active; x bits = active bits.
// ybits != NULL: the same sweep clears y's bit words (Clear(&y), GraphMatRuntime.h:139) -- one launch less per iteration.
template <class P, class T, class V>
__global__ void __launch_bounds__(256) k_send(prog_bytes<P> pb, int n_pad, const V* __restrict__ vp,
                                              const unsigned* __restrict__ active, T* __restrict__ x,
                                              unsigned* __restrict__ xbits, unsigned* __restrict__ ybits) {
  const P& prog = pb.get();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  unsigned w = active[i >> 5];
  bool on = (w >> (i & 31)) & 1u;
  if (on) {
    T t;
    (void)prog.P::send_message(vp[i], t);
    if constexpr (has_null_message<P, T>::value) on = !P::gm_null_message(t);
    if (on) x[i] = t;
  }
  if constexpr (has_null_message<P, T>::value) w = __ballot_sync(0xffffffffu, on);  // n_pad is a multiple of 32
  if ((i & 31) == 0) {
    xbits[i >> 5] = w;
    if (ybits) ybits[i >> 5] = 0u;
  }
}

// ------------------------------------------------------------------- apply --
// FUSE (ALL_VERTICES programs that do not override do_every_iteration): every vertex is active
// again in the next iteration, so the same pass also produces the next message vector
// (send_message on the fresh property) and re-arms the active set: one sweep over the vertex
// properties per iteration instead of three kernels.
constexpr int GM_APPLY_VPT = 4;  // vertices per thread: all loads of the 4 are issued before any is consumed
// c_ptr != NULL (ACTIVE_ONLY programs with the column-major companion built, one GPU): the vertices that change
// here ARE the next frontier, so the number of matrix entries their columns hold -- what decides between the
// sparse-frontier and the row-major pass -- is summed into next_entries and travels to the host with the
// "changed" flag: no separate counting kernel and no second host round trip per iteration.
// RESET (gm_atomic_min programs): y is handed back holding the identity of min wherever a message was
// consumed, so the next sparse pass can atomicMin into it without clearing the whole vector.
template <class P, class T, class U, class V, bool FUSE, bool RESET = false>
__global__ void __launch_bounds__(256) k_apply(prog_bytes<P> pb, int n_valid, int n_pad, U* __restrict__ y,
                                               const unsigned* __restrict__ ybits, V* __restrict__ vp,
                                               unsigned* __restrict__ active, int* __restrict__ flags,
                                               T* __restrict__ x, unsigned* __restrict__ xbits,
                                               const long long* __restrict__ c_ptr = nullptr,
                                               unsigned long long* __restrict__ next_entries = nullptr, int x_off = 0) {
  alignas(16) unsigned char pbuf[sizeof(P)];
  memcpy(pbuf, pb.b, sizeof(P));
  P& prog = *reinterpret_cast<P*>(pbuf);  // apply is non-const in the reference
  constexpr int VPT = (sizeof(V) + sizeof(U) <= 32) ? GM_APPLY_VPT : 1;
  const int base = blockIdx.x * (blockDim.x * VPT) + threadIdx.x;
  bool got[VPT], touch[VPT];
  V cur[VPT];
  U msg[VPT];
#pragma unroll
  for (int k = 0; k < VPT; k++) {
    const int i = base + k * 256;
    got[k] = touch[k] = false;
    if (i < n_pad) {
      got[k] = (__ldg(ybits + (i >> 5)) >> (i & 31)) & 1u;
      touch[k] = got[k] || (FUSE && i < n_valid);
      if (touch[k]) cur[k] = vp[i];
      if (got[k]) msg[k] = y[i];
    }
  }
  bool any = false;
  unsigned long long ents = 0;
  int n_changed = 0;  // warp-uniform
#pragma unroll
  for (int k = 0; k < VPT; k++) {
    const int i = base + k * 256;
    bool changed = false;
    if (got[k]) {
      V old = cur[k];
      prog.P::apply(msg[k], cur[k]);
      changed = (old != cur[k]);
      vp[i] = cur[k];
      if constexpr (RESET) {
        static_assert(sizeof(U) == 4, "gm_atomic_min programs reduce 32-bit unsigned messages");
        reinterpret_cast<unsigned*>(y)[i] = 0xffffffffu;
      }
      if (!FUSE && changed && c_ptr) ents += (unsigned long long)(__ldg(c_ptr + x_off + i + 1) - __ldg(c_ptr + x_off + i));
    }
    if (FUSE && touch[k]) {
      T t;
      (void)prog.P::send_message(cur[k], t);
      x[i] = t;
    }
    const unsigned m = __ballot_sync(0xffffffffu, changed);
    n_changed += __popc(m);
    if ((threadIdx.x & 31) == 0 && i < n_pad) {
      if (FUSE) {
        const unsigned all = i + 32 <= n_valid ? 0xffffffffu : (i < n_valid ? (1u << (n_valid - i)) - 1u : 0u);
        active[i >> 5] = all;  // setAllActive of the next iteration (GraphMatRuntime.h:250-252)
        xbits[i >> 5] = all;
      } else {
        active[i >> 5] = m;  // setAllInactive + set where changed
      }
      any |= (m != 0);
    }
  }
  if (any && *((volatile int*)flags) == 0) atomicExch(flags, 1);
  if (!FUSE && c_ptr) {
    for (int o = 16; o; o >>= 1) ents += __shfl_down_sync(0xffffffffu, ents, o);
    if ((threadIdx.x & 31) == 0 && n_changed) {
      atomicAdd(next_entries, ents);
      atomicAdd(next_entries - 1, (unsigned long long)n_changed);
    }
  }
}

// ---- ACTIVE_ONLY programs: the same two steps as sweeps over BIT WORDS ----
// A warp loads 32 words of the active set (y bits); only the non-empty words are visited, lane j taking bit j, so
// a step touches 32 consecutive vertices exactly like the per-vertex kernels do, but a sparse frontier costs a few
// hundred blocks instead of one thread per vertex (BFS on RMAT-22: 13 us -> ~3 us per sweep).
constexpr int GM_WORD_SPARSE = 4;  // words with at most this many bits are walked by their own lane
template <class P, class T, class V>
__global__ void __launch_bounds__(256) k_send_words(prog_bytes<P> pb, int n_words, const V* __restrict__ vp,
                                                    const unsigned* __restrict__ active, T* __restrict__ x,
                                                    unsigned* __restrict__ xbits, unsigned* __restrict__ ybits) {
  const P& prog = pb.get();
  const int lane = threadIdx.x & 31;
  const int w0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
  const int w = w0 + lane;
  unsigned word = w < n_words ? active[w] : 0u;
  unsigned out = word;
  const int pc = __popc(word);
  // a word with a few bits (the usual case on a thin frontier: about one active vertex per word) is walked by its own
  // lane -- 32 words in flight per warp instead of 32 one-lane steps; dense words go through the whole warp, lane j = bit j
  if (pc > 0 && pc <= GM_WORD_SPARSE) {
    unsigned m = word;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const int i = w * 32 + b;
      T t;
      (void)prog.P::send_message(vp[i], t);
      bool on = true;
      if constexpr (has_null_message<P, T>::value) on = !P::gm_null_message(t);
      if (on) x[i] = t;
      else out &= ~(1u << b);
    }
  }
  unsigned lanes = __ballot_sync(0xffffffffu, pc > GM_WORD_SPARSE);
  while (lanes) {
    const int src = __ffs(lanes) - 1;
    lanes &= lanes - 1;
    const unsigned m = __shfl_sync(0xffffffffu, word, src);
    bool on = (m >> lane) & 1u;
    if (on) {
      const int i = (w0 + src) * 32 + lane;
      T t;
      (void)prog.P::send_message(vp[i], t);
      if constexpr (has_null_message<P, T>::value) on = !P::gm_null_message(t);
      if (on) x[i] = t;
    }
    if constexpr (has_null_message<P, T>::value) {
      const unsigned kept = __ballot_sync(0xffffffffu, on);
      if (lane == src) out = kept;
    }
  }
  if (w < n_words) {
    xbits[w] = out;
    if (ybits) ybits[w] = 0u;
  }
}
template <class P, class T, class U, class V, bool RESET>
__global__ void __launch_bounds__(256) k_apply_words(prog_bytes<P> pb, int n_words, U* __restrict__ y,
                                                     const unsigned* __restrict__ ybits, V* __restrict__ vp,
                                                     unsigned* __restrict__ active, int* __restrict__ flags,
                                                     const long long* __restrict__ c_ptr,
                                                     unsigned long long* __restrict__ next_entries, int x_off) {
  alignas(16) unsigned char pbuf[sizeof(P)];
  memcpy(pbuf, pb.b, sizeof(P));
  P& prog = *reinterpret_cast<P*>(pbuf);  // apply is non-const in the reference
  const int lane = threadIdx.x & 31;
  const int w0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
  const int w = w0 + lane;
  const unsigned word = w < n_words ? __ldg(ybits + w) : 0u;
  unsigned out = 0u;
  unsigned long long ents = 0;
  auto one = [&](int i) -> bool {  // the apply loop body for vertex i (GraphMatRuntime.h:201-213); returns "changed"
    V cur = vp[i];
    const V old = cur;
    const U msg = y[i];
    prog.P::apply(msg, cur);
    const bool changed = (old != cur);
    vp[i] = cur;
    if constexpr (RESET) reinterpret_cast<unsigned*>(y)[i] = 0xffffffffu;
    if (changed && c_ptr) ents += (unsigned long long)(__ldg(c_ptr + x_off + i + 1) - __ldg(c_ptr + x_off + i));
    return changed;
  };
  const int pc = __popc(word);
  if (pc > 0 && pc <= GM_WORD_SPARSE) {  // few messages in this word: its own lane walks them (see k_send_words)
    unsigned m = word;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      if (one(w * 32 + b)) out |= 1u << b;
    }
  }
  unsigned lanes = __ballot_sync(0xffffffffu, pc > GM_WORD_SPARSE);
  while (lanes) {
    const int src = __ffs(lanes) - 1;
    lanes &= lanes - 1;
    const unsigned m = __shfl_sync(0xffffffffu, word, src);
    bool changed = false;
    if ((m >> lane) & 1u) changed = one((w0 + src) * 32 + lane);
    const unsigned cm = __ballot_sync(0xffffffffu, changed);
    if (lane == src) out = cm;
  }
  if (w < n_words) active[w] = out;  // setAllInactive + set where changed
  if (__ballot_sync(0xffffffffu, out != 0) != 0 && lane == 0) raise_flag(flags);
  if (c_ptr) {
    unsigned long long verts = __popc(out);
    for (int o = 16; o; o >>= 1) {
      ents += __shfl_down_sync(0xffffffffu, ents, o);
      verts += __shfl_down_sync(0xffffffffu, verts, o);
    }
    if (lane == 0 && verts) {
      atomicAdd(next_entries, ents);
      atomicAdd(next_entries - 1, verts);  // [-1]: vertices of the next frontier, [0]: entries of their columns
    }
  }
}

__global__ void k_fill_bits(unsigned* bits, int n_valid, int n_pad) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (n_pad >> 5)) return;
  int lo = w << 5;
  unsigned v = 0;
  if (lo + 32 <= n_valid) v = 0xffffffffu;
  else if (lo < n_valid) v = (1u << (n_valid - lo)) - 1u;
  bits[w] = v;
}

// ------------------------------------------------------- SpMSpV: sliced ELL --
// One row per lane.  The fold is a private left-to-right chain per lane; loads of
// UNROLL steps are issued before they are consumed.
// Launch shape: block b, warp w folds slices [slice_begin + (8b + w) * spw, + spw).  Slices are
// stored longest first and blocks are dispatched in order, so the hardware block scheduler gives
// longest-processing-time-first load balance for free; spw > 1 only for the short-row tail.
template <class P, class T, class U, class V, class E, bool ALLACT, bool NEEDVP, bool IDENT, bool ACCUM, int UNROLL,
          bool EPI = false>
__global__ void __launch_bounds__(256)
    k_sell(prog_bytes<P> pb, gm_matrix_view M, int slice_begin, int slice_end, int spw, int hot_limit,
           const T* __restrict__ x, const unsigned* __restrict__ xbits, const V* __restrict__ vp, U* __restrict__ y,
           unsigned* __restrict__ ybits, epilogue<T, V> ep = epilogue<T, V>()) {
  const P& prog = pb.get();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* __restrict__ cols = M.s_col;
  const E* __restrict__ vals = reinterpret_cast<const E*>(M.s_val);
  int s = slice_begin + warp * spw;
  const int s_end = min(s + spw, slice_end);
  for (; s < s_end; s++) {
    const int slot = M.n_heavy + s * 32 + lane;
    const int len = __ldg(M.row_len + slot);
    const long long base = __ldg(M.slice_ptr + s);
    const int width = __shfl_sync(0xffffffffu, len, 0);  // rows are sorted: lane 0 is the longest
    const int vtx = IDENT ? slot : __ldg(M.slot_vertex + slot);
    V vprop;  // V(): the dummy of SPMV.h:44 when the program does not read it
    if (NEEDVP && len > 0) vprop = vp[vtx];
    U acc;
    bool have = false;
    if (ACCUM && (IDENT || len > 0)) {  // second operand of ALL_EDGES: continue the fold held in y
      have = test_bit(ybits, vtx);
      if (have && len > 0) acc = y[vtx];
    }
    const int* cp = cols + base + lane;
    const E* evp = vals + base + lane;
    for (int i = 0; i < width; i += UNROLL) {
      int c[UNROLL];
      E ev[UNROLL];
      bool on[UNROLL];
      T xv[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        on[u] = (i + u) < len;
        if (on[u]) {
          c[u] = ld_stream(cp + (long long)(i + u) * 32);
          ev[u] = ld_stream(evp + (long long)(i + u) * 32);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        if (on[u]) {
          if (!ALLACT) on[u] = test_bit(xbits, c[u]);
          if (on[u]) xv[u] = ld_gather(x, c[u], hot_limit);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        if (on[u]) {
          if (have) {
            U tmp;
            prog.P::process_message(xv[u], ev[u], vprop, tmp);
            prog.P::reduce_function(acc, tmp);
          } else {
            prog.P::process_message(xv[u], ev[u], vprop, acc);
            have = true;
          }
        }
      }
    }
    if constexpr (EPI) {
      const bool ch = fused_apply_send<P, T, U, V>(pb, ep, vtx, have && len > 0, acc);
      if (__ballot_sync(0xffffffffu, ch) != 0 && lane == 0) raise_flag(ep.flag);
    } else {
      if (have && len > 0) y[vtx] = acc;
      unsigned m = __ballot_sync(0xffffffffu, have);
      if (IDENT) {
        if (lane == 0) ybits[slot >> 5] = m;  // heavy and ELL slots never share a word
      } else if (have && len > 0) {
        atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
      }
    }
  }
}

// ---------------------------------- SpMSpV: sliced ELL, last-writer programs --
// Same launch shape as k_sell; every lane walks its row from the END and stops at the first active
// entry (is_last_writer).  The warp leaves a slice when all its rows are decided.
template <class P, class T, class U, class V, class E, bool NEEDVP, bool IDENT, bool ACCUM, int UNROLL>
__global__ void __launch_bounds__(256)
    k_sell_last(prog_bytes<P> pb, gm_matrix_view M, int slice_begin, int slice_end, int spw, int hot_limit,
                const T* __restrict__ x, const unsigned* __restrict__ xbits, const V* __restrict__ vp, U* __restrict__ y,
                unsigned* __restrict__ ybits) {
  const P& prog = pb.get();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* __restrict__ cols = M.s_col;
  const E* __restrict__ vals = reinterpret_cast<const E*>(M.s_val);
  int s = slice_begin + warp * spw;
  const int s_end = min(s + spw, slice_end);
  for (; s < s_end; s++) {
    const int slot = M.n_heavy + s * 32 + lane;
    const int len = __ldg(M.row_len + slot);
    const long long base = __ldg(M.slice_ptr + s);
    const int width = __shfl_sync(0xffffffffu, len, 0);
    const int vtx = IDENT ? slot : __ldg(M.slot_vertex + slot);
    bool have = false;
    if (ACCUM && (IDENT || len > 0)) have = test_bit(ybits, vtx);
    bool found = false;
    U acc;
    const int* cp = cols + base + lane;
    const E* ep = vals + base + lane;
    for (int top = width; top > 0; top -= UNROLL) {
      int c[UNROLL];
      bool on[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {  // u = 0 is the highest index of the batch
        const int i = top - 1 - u;
        on[u] = !found && i >= 0 && i < len;
        if (on[u]) c[u] = ld_stream(cp + (long long)i * 32);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; u++)
        if (on[u]) on[u] = test_bit(xbits, c[u]);
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        if (on[u] && !found) {
          const int i = top - 1 - u;
          V vprop;
          if (NEEDVP) vprop = vp[vtx];
          prog.P::process_message(ld_gather(x, c[u], hot_limit), ld_stream(ep + (long long)i * 32), vprop, acc);
          found = true;
        }
      }
      if (__ballot_sync(0xffffffffu, !found && len > 0) == 0) break;
    }
    if (found) y[vtx] = acc;
    have = have || found;
    const unsigned m = __ballot_sync(0xffffffffu, have);
    if (IDENT) {
      if (lane == 0) ybits[slot >> 5] = m;
    } else if (found) {
      atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
    }
  }
}

// ------------------------------------------------- SpMSpV: heavy rows (CSR) --
// One warp per row: the index stream is read 32 wide (coalesced), the gathers are
// 32 in flight, process_message runs on all lanes.  The fold of each batch is
//   REORDER = false: lane 0 walks the batch left to right (exact serial order)
//   REORDER = true : order-preserving pairwise tree (program declared associative)
template <class P, class T, class U, class V, class E, bool ALLACT, bool NEEDVP, bool IDENT, bool ACCUM, bool REORDER,
          bool EPI = false>
__global__ void __launch_bounds__(128)
    k_heavy(prog_bytes<P> pb, gm_matrix_view M, int row_begin, int hot_limit, const T* __restrict__ x,
            const unsigned* __restrict__ xbits, const V* __restrict__ vp, U* __restrict__ y, unsigned* __restrict__ ybits,
            epilogue<T, V> ep = epilogue<T, V>()) {
  const P& prog = pb.get();
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  U* buf = reinterpret_cast<U*>(smem) + wib * 32;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int* __restrict__ cols = M.h_col;
  const E* __restrict__ vals = reinterpret_cast<const E*>(M.h_val);

  for (int slot = row_begin + warp; slot < M.n_heavy; slot += nwarps) {
    const long long beg = __ldg(M.h_ptr + slot), end = __ldg(M.h_ptr + slot + 1);
    const int vtx = IDENT ? slot : __ldg(M.slot_vertex + slot);
    if (beg == end) {  // padding slot of the heavy prefix
      if constexpr (EPI) {
        U none;
        if (lane == 0 && fused_apply_send<P, T, U, V>(pb, ep, vtx, false, none)) raise_flag(ep.flag);
      }
      continue;
    }
    V vprop;
    if (NEEDVP) vprop = vp[vtx];
    U acc;
    bool have = false;  // meaningful on lane 0
    if (ACCUM && lane == 0) {
      have = test_bit(ybits, vtx);
      if (have) acc = y[vtx];
    }
    for (long long k = beg; k < end; k += 32) {
      const long long idx = k + lane;
      bool on = idx < end;
      int c = 0;
      if (on) {
        c = ld_stream(cols + idx);
        if (!ALLACT) on = test_bit(xbits, c);
      }
      if (on) {
        T xv = ld_gather(x, c, hot_limit);
        E ev = ld_stream(vals + idx);
        U tmp;
        prog.P::process_message(xv, ev, vprop, tmp);
        buf[lane] = tmp;
      }
      unsigned m = __ballot_sync(0xffffffffu, on);
      if (m == 0) continue;
      __syncwarp();
      if (REORDER) {
        // order-preserving pairwise tree: at stride d the lane with lane % 2d == 0
        // absorbs lane + d (left operand first); validity rides along in `v`
        bool v = on;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const unsigned vm = __ballot_sync(0xffffffffu, v);
          if ((lane & (2 * d - 1)) == 0 && ((vm >> (lane + d)) & 1u)) {
            if (v) {
              U a = buf[lane];
              prog.P::reduce_function(a, buf[lane + d]);
              buf[lane] = a;
            } else {
              buf[lane] = buf[lane + d];
              v = true;
            }
          }
          __syncwarp();
        }
        if (lane == 0) {
          if (have) prog.P::reduce_function(acc, buf[0]);
          else { acc = buf[0]; have = true; }
        }
      } else {
        if (lane == 0) {
          unsigned mm = m;
          while (mm) {
            int b = __ffs(mm) - 1;
            mm &= mm - 1;
            if (have) prog.P::reduce_function(acc, buf[b]);
            else { acc = buf[b]; have = true; }
          }
        }
      }
      __syncwarp();
    }
    if constexpr (EPI) {
      if (lane == 0 && fused_apply_send<P, T, U, V>(pb, ep, vtx, have, acc)) raise_flag(ep.flag);
    } else if (lane == 0 && have) {
      y[vtx] = acc;
      atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
    }
  }
}


// ------------------- SpMSpV: heavy rows in two phases (associative programs) --
// Phase 1, k_heavy_seg: heavy rows are cut into segments of seg_len consecutive entries; a warp
// folds one segment (lane L the L-th run of seg_len/32 consecutive entries, left to right; then an
// order-preserving tree over the 32 lane partials) into partial[seg].  Phase 2, k_heavy_combine:
// a warp folds the partials of one row, again as ordered lane runs + ordered tree, and appends the
// result to y.  Order is preserved everywhere, only the association changes, so this is exact for
// min / integer + / "last writer" and within rounding for fp64 sums.  Any row length is spread over
// as many warps as it has segments: no tail behind the longest row.
template <class U, class P>
__device__ __forceinline__ void warp_ordered_tree(const P& prog, U* wb, U& part, bool& pv, int lane) {
  if (pv) wb[lane] = part;
  __syncwarp();
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned vm = __ballot_sync(0xffffffffu, pv);
    if ((lane & (2 * d - 1)) == 0 && ((vm >> (lane + d)) & 1u)) {
      if (pv) {
        U a = wb[lane];
        prog.P::reduce_function(a, wb[lane + d]);
        wb[lane] = a;
      } else {
        wb[lane] = wb[lane + d];
        pv = true;
      }
    }
    __syncwarp();
  }
}

template <class P, class T, class U, class V, class E, bool ALLACT, bool NEEDVP, bool IDENT, bool LAST = false>
__global__ void __launch_bounds__(128)
    k_heavy_seg(prog_bytes<P> pb, gm_matrix_view M, int hot_limit, const T* __restrict__ x,
                const unsigned* __restrict__ xbits, const V* __restrict__ vp, U* __restrict__ partial,
                unsigned char* __restrict__ pvalid) {
  const P& prog = pb.get();
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  U* wb = reinterpret_cast<U*>(smem) + wib * 32;
  const int seg = blockIdx.x * 4 + wib;
  if (seg >= M.n_segs) return;
  const int row = __ldg(M.seg_row + seg);
  const long long rbeg = __ldg(M.h_ptr + row), rend = __ldg(M.h_ptr + row + 1);
  const long long sbeg = rbeg + (long long)(seg - __ldg(M.seg_ptr + row)) * M.seg_len;
  const long long send = min(sbeg + M.seg_len, rend);
  const int run = M.seg_len >> 5;
  long long i = sbeg + (long long)lane * run;
  const long long iend = min(i + run, send);
  const int* __restrict__ cols = M.h_col;
  const E* __restrict__ vals = reinterpret_cast<const E*>(M.h_val);
  V vprop;
  if (NEEDVP) vprop = vp[IDENT ? row : __ldg(M.slot_vertex + row)];
  U part;
  bool pv = false;
  constexpr int UNROLL = sizeof(T) <= 8 ? 8 : 1;
  if constexpr (LAST) {
    // last-writer programs: this lane's run from its end, stop at the first active entry
    for (long long top = iend; top > i && !pv; top -= UNROLL) {
      int c[UNROLL];
      bool on[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        on[u] = top - 1 - u >= i;
        if (on[u]) c[u] = __ldg(cols + top - 1 - u);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; u++)
        if (on[u] && !ALLACT) on[u] = test_bit(xbits, c[u]);
#pragma unroll
      for (int u = 0; u < UNROLL; u++) {
        if (on[u] && !pv) {
          prog.P::process_message(ld_gather(x, c[u], hot_limit), __ldg(vals + top - 1 - u), vprop, part);
          pv = true;
        }
      }
    }
    i = iend;
  }
  // a lane reads 8 consecutive entries: let L1 keep the sector between the 8 loads
  for (; i < iend; i += UNROLL) {
    int c[UNROLL];
    E ev[UNROLL];
    bool on[UNROLL];
    T xv[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      on[u] = i + u < iend;
      if (on[u]) {
        c[u] = __ldg(cols + i + u);
        ev[u] = __ldg(vals + i + u);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (on[u]) {
        if (!ALLACT) on[u] = test_bit(xbits, c[u]);
        if (on[u]) xv[u] = ld_gather(x, c[u], hot_limit);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (on[u]) {
        if (pv) {
          U tmp;
          prog.P::process_message(xv[u], ev[u], vprop, tmp);
          prog.P::reduce_function(part, tmp);
        } else {
          prog.P::process_message(xv[u], ev[u], vprop, part);
          pv = true;
        }
      }
    }
  }
  warp_ordered_tree<U, P>(prog, wb, part, pv, lane);
  if (lane == 0) {
    pvalid[seg] = pv ? 1 : 0;
    if (pv) partial[seg] = wb[0];
  }
}

template <class P, class T, class U, class V, bool IDENT, bool ACCUM, bool EPI = false>
__global__ void __launch_bounds__(128)
    k_heavy_combine(prog_bytes<P> pb, gm_matrix_view M, const U* __restrict__ partial,
                    const unsigned char* __restrict__ pvalid, U* __restrict__ y, unsigned* __restrict__ ybits,
                    epilogue<T, V> ep = epilogue<T, V>()) {
  const P& prog = pb.get();
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  U* wb = reinterpret_cast<U*>(smem) + wib * 32;
  const int row = blockIdx.x * 4 + wib;
  if (row >= M.n_heavy) return;
  const int s0 = __ldg(M.seg_ptr + row), s1 = __ldg(M.seg_ptr + row + 1);
  const int vtx = IDENT ? row : __ldg(M.slot_vertex + row);
  if (s0 == s1) {  // empty row inside the heavy prefix
    if constexpr (EPI) {
      U none;
      if (lane == 0 && fused_apply_send<P, T, U, V>(pb, ep, vtx, false, none)) raise_flag(ep.flag);
    }
    return;
  }
  const int run = (s1 - s0 + 31) >> 5;
  int k = s0 + lane * run;
  const int kend = min(k + run, s1);
  U part;
  bool pv = false;
  for (; k < kend; k++) {
    if (pvalid[k]) {
      if (pv) prog.P::reduce_function(part, partial[k]);
      else { part = partial[k]; pv = true; }
    }
  }
  warp_ordered_tree<U, P>(prog, wb, part, pv, lane);
  if constexpr (EPI) {
    if (lane == 0) {
      U r;
      if (pv) r = wb[0];
      if (fused_apply_send<P, T, U, V>(pb, ep, vtx, pv, r)) raise_flag(ep.flag);
    }
  } else if (lane == 0 && pv) {
    if (ACCUM && test_bit(ybits, vtx)) {
      U a = y[vtx];
      prog.P::reduce_function(a, wb[0]);
      y[vtx] = a;
    } else {
      y[vtx] = wb[0];
    }
    atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
  }
}

// ------------------------------ SpMSpV: heavy rows, exact fp32 + (gm_fadd32_exact) --
// A lane owns 8 CONSECUTIVE entries (two 128-bit index loads); the fold is the bit-exact parallel evaluation of
// the serial fp32 sum (gm_fadd32.cuh).  W = 1: one warp per row, several rows per block.
// W > 1: the block folds W*256 addends per round.  Every warp scans its 256 under the
// binade of the block's entry value; warp 0 scans the W warp totals and finds the first
// warp whose exit value would leave the binade (or that saw an addend the scan cannot
// prove).  Warps before it are applied, that warp alone runs the warp-level fold from
// its exact entry value, and the remaining warps are re-scanned under the new binade.
// The few rows above gm_matrix_view::n_long entries are the critical path of a pass (one block walks the row
// round by round, each round an index load, then a gather, then the scan): their gathers are hoisted out of
// the block.  k_stage_rows lets the whole GPU write x[h_col[k]] for those rows into a dense staging array;
// k_heavy_fadd32_tma (below) then folds from that array, the same addends in the same order.
template <class T>
__global__ void __launch_bounds__(256)
    k_stage_rows(const int* __restrict__ h_col, long long n_entries, const T* __restrict__ x, int hot_limit,
                 T* __restrict__ staged) {
  const long long i0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8;
  if (i0 >= n_entries) return;
  int c[8];
  *reinterpret_cast<int4*>(&c[0]) = ld_stream4(h_col + i0);
  *reinterpret_cast<int4*>(&c[4]) = ld_stream4(h_col + i0 + 4);
  T v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) v[j] = ld_gather(x, c[j], hot_limit);  // the pad behind the last row holds column 0
#pragma unroll
  for (int j = 0; j < 8; j++) staged[i0 + j] = v[j];
}

template <class P, class T, class V, class E, bool ALLACT, bool IDENT, int W, bool EPI = false>
__global__ void __launch_bounds__(W == 1 ? 128 : W * 32)
    k_heavy_fadd32(prog_bytes<P> pb, gm_matrix_view M, int row_begin, int row_end, int hot_limit,
                   const T* __restrict__ x, const unsigned* __restrict__ xbits, float* __restrict__ y,
                   unsigned* __restrict__ ybits, epilogue<T, V> ep = epilogue<T, V>()) {
  const P& prog = pb.get();
  constexpr int WPB = (W == 1) ? 4 : W;  // warps per block
  __shared__ float sm_s;
  __shared__ int sm_have, sm_fail;
  __shared__ unsigned sm_d0[WPB], sm_d1[WPB], sm_bad[WPB];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int* __restrict__ cols = M.h_col;
  const E* __restrict__ vals = reinterpret_cast<const E*>(M.h_val);
  V vdummy;
  const int stride = (W == 1) ? gridDim.x * WPB : gridDim.x;
  for (int slot = row_begin + ((W == 1) ? blockIdx.x * WPB + w : blockIdx.x); slot < row_end; slot += stride) {
    const long long beg = __ldg(M.h_ptr + slot), end = __ldg(M.h_ptr + slot + 1);
    const int vtx = IDENT ? slot : __ldg(M.slot_vertex + slot);
    float s = 0.f;
    bool have = false;
    if (W > 1) {
      if (threadIdx.x == 0) { sm_s = 0.f; sm_have = 0; }
      __syncthreads();
    }
    for (long long k0 = beg & ~7ll; k0 < end; k0 += W * 256) {
      const long long i0 = k0 + (W == 1 ? 0 : w * 256) + lane * 8;
      float v[8];
      unsigned vmask = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) v[j] = 0.f;
      if (i0 < end && i0 + 8 > beg) {
        int c[8];
        E ev[8];
        *reinterpret_cast<int4*>(&c[0]) = ld_stream4(cols + i0);
        *reinterpret_cast<int4*>(&c[4]) = ld_stream4(cols + i0 + 4);
#pragma unroll
        for (int j = 0; j < 8; j++) ev[j] = ld_stream(vals + i0 + j);
        T xv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          bool on = (i0 + j >= beg) && (i0 + j < end);
          if (on && !ALLACT) on = test_bit(xbits, c[j]);
          if (on) { xv[j] = ld_gather(x, c[j], hot_limit); vmask |= 1u << j; }
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
          if ((vmask >> j) & 1u) prog.P::process_message(xv[j], ev[j], vdummy, v[j]);
      }
      if (W == 1) {
        fx::warp_fold(v, vmask, s, have, lane);
      } else {
        // warps of this round that hold at least one addend (block-uniform by construction)
        const long long wfirst = k0, wspan = end - k0;
        int nw = (int)((wspan + 255) / 256);
        if (nw > W) nw = W;
        (void)wfirst;
        int first = 0;  // warps [first, nw) are still to be applied
        while (first < nw) {
          s = sm_s;
          have = sm_have != 0;
          fx::binade b;
          const bool hot = have && fx::binade_of(s, b);  // block-uniform
          if (!hot) {
            // no usable accumulator yet: warp `first` folds its 256 alone (exact for any input).  It publishes the
            // result in sm_s / sm_have, which every warp has just read to decide `hot`: the barrier keeps a warp that
            // is still waiting for its gathers from reading the NEW value and taking the other branch (found in
            // round 2: rows went wrong once other kernels ran beside this one, profiles/r2_multistream_rejected.txt)
            __syncthreads();
            if (w == first) {
              float sq = s;
              bool hq = have;
              fx::warp_fold(v, vmask, sq, hq, lane);
              if (lane == 0) { sm_s = sq; sm_have = hq ? 1 : 0; }
            }
            first++;
            __syncthreads();
            continue;
          }
          if (w >= first && w < nw) {
            bool bad = false;
            fx::qmap mine = fx::identity();
#pragma unroll
            for (int j = 0; j < 8; j++) mine = fx::compose(mine, fx::quantize(v[j], b, bad));
            const fx::qmap incl = fx::warp_scan(mine, lane);
            const unsigned anybad = __ballot_sync(0xffffffffu, bad);
            if (lane == 31) { sm_d0[w] = incl.d0; sm_d1[w] = incl.d1; sm_bad[w] = anybad; }
          }
          __syncthreads();
          if (w == 0) {
            fx::qmap t = fx::identity();
            bool bd = false;
            if (lane >= first && lane < nw) { t.d0 = sm_d0[lane]; t.d1 = sm_d1[lane]; bd = sm_bad[lane] != 0; }
            t = fx::warp_scan(t, lane);
            const unsigned m_after = fx::apply(t, b.m);
            const bool over = (lane >= first && lane < nw) && (bd || m_after >= (1u << 24));
            const unsigned fail = __ballot_sync(0xffffffffu, over);
            unsigned m_prev = __shfl_up_sync(0xffffffffu, m_after, 1);
            if (lane == 0) m_prev = b.m;
            if (fail == 0) {
              if (lane == nw - 1) { sm_s = __fmul_rn(__uint2float_rn(m_after), b.u); sm_fail = -1; }
            } else {
              const int f = __ffs(fail) - 1;
              if (lane == f) { sm_s = __fmul_rn(__uint2float_rn(m_prev), b.u); sm_fail = f; }
            }
          }
          __syncthreads();
          const int f = sm_fail;
          if (f < 0) break;
          if (w == f) {
            float sq = sm_s;
            bool hq = true;
            fx::warp_fold(v, vmask, sq, hq, lane);
            if (lane == 0) sm_s = sq;
          }
          first = f + 1;
          __syncthreads();
        }
      }
    }
    if (W == 1) {
      if constexpr (EPI) {
        if (lane == 0 && fused_apply_send<P, T, float, V>(pb, ep, vtx, have, s)) raise_flag(ep.flag);
      } else if (lane == 0 && have) {
        y[vtx] = s;
        atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
      }
    } else {
      __syncthreads();
      if constexpr (EPI) {
        if (threadIdx.x == 0) {
          const float r = sm_s;
          if (fused_apply_send<P, T, float, V>(pb, ep, vtx, sm_have != 0, r)) raise_flag(ep.flag);
        }
      } else if (threadIdx.x == 0 && sm_have) {
        y[vtx] = sm_s;
        atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
      }
      __syncthreads();
    }
  }
}

// ---------- SpMSpV: the longest rows, exact fp32 fold from the staging array through TMA --
// The rows above gm_matrix_view::n_long entries are the critical path of a sharded pass: one block walks a row
// round by round.  Their gathers are already hoisted out (k_stage_rows); this kernel streams the dense staging
// array with BULK ASYNC COPIES (cp.async.bulk, the TMA engine) into a double-buffered shared-memory ring, one
// mbarrier per buffer, so the copy of round r+1 runs under the fold of round r and a round costs shared-memory
// reads instead of an L2 round trip.  A round is GM_TMA_ROUND = 32 warps x GM_TMA_CH chunks x 256 addends; every warp
// scans its chunks under the binade of the block's entry value and publishes ONE composed map; after a single
// block barrier EVERY warp scans the 32 published maps itself and so knows the new running value (held in
// registers by all threads) -- a round that stays inside its binade costs one barrier (k_heavy_fadd32: three, with
// warp 0 scanning while the others wait).  Same addends, same order, same exactness argument as k_heavy_fadd32
// (all q >= 0: the exit value decides).
constexpr int GM_TMA_W = 32;                                   // warps per block
#ifndef GM_TMA_CHUNKS
#define GM_TMA_CHUNKS 1
#endif
#ifndef GM_TMA_NBUF
#define GM_TMA_NBUF 2
#endif
constexpr int GM_TMA_CH = GM_TMA_CHUNKS;                       // 256-addend chunks per warp and round
constexpr int GM_TMA_ROUND = GM_TMA_W * 256 * GM_TMA_CH;       // addends per round (8192 -> 32 KB per buffer)
// The ring is kept SMALL (2 x 32 KB): the SMs that hosted a long-row block keep its shared-memory carve-out, and the
// sliced-ELL blocks that follow there lose that much L1.  Measured, pass of rank 0 of 2 / 4 / 8 (176 / 88 / 44
// long-row blocks): 3 x 64 KB 3.08 / - / 0.725 ms, 2 x 64 KB 2.28 / 1.22 / 0.70 ms, 2 x 32 KB 2.17 / 1.19 / 0.69 ms.
constexpr int GM_TMA_BUFS = GM_TMA_NBUF;
constexpr int GM_TMA_PARTS = 8;                                // bulk copies per round (8 KB each)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  }
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class P, class T, class V, bool IDENT, bool EPI>
__global__ void __launch_bounds__(GM_TMA_W * 32)
    k_heavy_fadd32_tma(prog_bytes<P> pb, gm_matrix_view M, int row_begin, int row_end, const float* __restrict__ staged,
                       float* __restrict__ y, unsigned* __restrict__ ybits, epilogue<T, V> ep) {
  static_assert(sizeof(T) == 4, "staged rows hold 4-byte messages");
  constexpr int W = GM_TMA_W, CH = GM_TMA_CH, NB = GM_TMA_BUFS;
  extern __shared__ __align__(128) unsigned char tma_smem[];
  float* ring = reinterpret_cast<float*>(tma_smem);  // NB buffers of GM_TMA_ROUND floats
  __shared__ __align__(8) unsigned long long bar[NB];
  // published values are double-buffered by generation: a warp may already write generation g + 1 while a slower
  // warp still reads generation g (a writer of g + 2 has passed the barrier of g + 1, which every reader of g reaches
  // only after its read)
  __shared__ float sm_s[2];
  __shared__ int sm_have[2];
  __shared__ unsigned sm_d0[2][W], sm_d1[2][W], sm_bad[2][W];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NB; i++) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned phases = 0u;  // bit i = the parity buffer i's mbarrier completes next
  unsigned gen = 0;
  // one round = GM_TMA_ROUND floats, fetched as GM_TMA_PARTS bulk copies (lanes of warp 0) on one mbarrier, so the
  // copy engine has several requests in flight; rounds r+1 .. r+NB-1 are in flight while round r is folded
  auto issue = [&](int b, const float* src) {
    if (w == 0) {
      if (lane == 0) mbar_expect_tx(&bar[b], GM_TMA_ROUND * 4);
      __syncwarp();
      if (lane < GM_TMA_PARTS) {
        constexpr int part = GM_TMA_ROUND / GM_TMA_PARTS;
        tma_load_1d(ring + (size_t)b * GM_TMA_ROUND + lane * part, src + lane * part, part * 4, &bar[b]);
      }
    }
  };
  for (int slot = row_begin + blockIdx.x; slot < row_end; slot += gridDim.x) {
    const long long beg = __ldg(M.h_ptr + slot), end = __ldg(M.h_ptr + slot + 1);
    const int vtx = IDENT ? slot : __ldg(M.slot_vertex + slot);
    // the running value of the fold, held by EVERY thread (block-uniform): the fast path never goes through memory
    float s = 0.f;
    bool have = false;
    const long long k_first = beg & ~7ll;  // 32-byte aligned start: the staging array is 256-byte aligned
    const long long n_rounds = (end - k_first + GM_TMA_ROUND - 1) / GM_TMA_ROUND;
    __syncthreads();  // everybody is done reading the ring (previous row) before the async proxy overwrites it
    for (int q = 0; q < NB - 1 && q < n_rounds; q++) issue(q, staged + k_first + (long long)q * GM_TMA_ROUND);
    int cur = 0;
    for (long long r = 0; r < n_rounds; r++) {
      const long long k0 = k_first + r * GM_TMA_ROUND;
      // the buffer of round r + NB - 1 is the one round r - 1 used, and every path of a round ends with a block barrier
      // behind its last read
      if (r + NB - 1 < n_rounds) issue((cur + NB - 1) % NB, staged + k0 + (long long)(NB - 1) * GM_TMA_ROUND);
      mbar_wait(&bar[cur], (phases >> cur) & 1u);
      phases ^= 1u << cur;
      const float* buf = ring + (size_t)cur * GM_TMA_ROUND;
      cur = (cur + 1) % NB;
      // chunk c of warp w: addends [k0 + (w*CH + c)*256 + lane*8, +8)
      auto load = [&](int c, float (&v)[8], unsigned& vmask) {
        const int off = (w * CH + c) * 256 + lane * 8;
        const float4 a = *reinterpret_cast<const float4*>(buf + off);
        const float4 b4 = *reinterpret_cast<const float4*>(buf + off + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b4.x; v[5] = b4.y; v[6] = b4.z; v[7] = b4.w;
        const long long i0 = k0 + off;
        if (i0 >= beg && i0 + 8 <= end) {
          vmask = 0xffu;
        } else {
          vmask = 0;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (i0 + j >= beg && i0 + j < end) vmask |= 1u << j;
            else v[j] = 0.f;  // outside the row: +0, the identity of the fold
          }
        }
      };
      const long long span = end - k0;
      int nw = (int)((span + 256 * CH - 1) / (256 * CH));  // warps of this round that hold at least one addend
      if (nw > W) nw = W;
      int first = 0;  // warps [first, nw) are still to be applied
      while (first < nw) {
        fx::binade b;
        const bool hot = have && fx::binade_of(s, b);  // block-uniform
        const unsigned g = gen & 1u;
        if (!hot) {
          // no usable accumulator yet: warp `first` folds its chunks alone (exact for any input) and publishes
          if (w == first) {
            float sq = s;
            bool hq = have;
            for (int c = 0; c < CH; c++) {
              float v[8];
              unsigned vmask;
              load(c, v, vmask);
              fx::warp_fold(v, vmask, sq, hq, lane);
            }
            if (lane == 0) { sm_s[g] = sq; sm_have[g] = hq ? 1 : 0; }
          }
          __syncthreads();
          s = sm_s[g];
          have = sm_have[g] != 0;
          gen++;
          first++;
          continue;
        }
        if (w >= first && w < nw) {
          bool bad = false;
          fx::qmap wt = fx::identity();
#pragma unroll
          for (int c = 0; c < CH; c++) {
            float v[8];
            unsigned vmask;
            load(c, v, vmask);
            // fast path (no addend of the chunk is a tie): the chunk is the translation by the sum of its q0,
            // one integer add per addend and one warp reduction -- only the warp TOTAL is needed here
            bool tie = false;
            unsigned sum = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) sum += fx::quantize_q0(v[j], b, bad, tie);  // < 2^26 per lane
            fx::qmap tot;
            if (!__any_sync(0xffffffffu, tie)) {
              const unsigned total = __reduce_add_sync(0xffffffffu, sum);  // 32 lanes x < 2^26: no overflow
              tot.d0 = tot.d1 = total > fx::kSat ? fx::kSat : total;
            } else {
              fx::qmap mine = fx::identity();
#pragma unroll
              for (int j = 0; j < 8; j++) mine = fx::compose(mine, fx::quantize(v[j], b, bad));
              const fx::qmap incl = fx::warp_scan(mine, lane);
              tot.d0 = __shfl_sync(0xffffffffu, incl.d0, 31);
              tot.d1 = __shfl_sync(0xffffffffu, incl.d1, 31);
            }
            wt = fx::compose(wt, tot);
          }
          const unsigned anybad = __ballot_sync(0xffffffffu, bad);
          if (lane == 31) { sm_d0[g][w] = wt.d0; sm_d1[g][w] = wt.d1; sm_bad[g][w] = anybad; }
        }
        __syncthreads();  // the ONLY block barrier of a round that stays inside its binade
        // every warp scans the W totals itself (same inputs, same result): no second barrier, no broadcast
        fx::qmap t = fx::identity();
        bool bd = false;
        if (lane >= first && lane < nw) { t.d0 = sm_d0[g][lane]; t.d1 = sm_d1[g][lane]; bd = sm_bad[g][lane] != 0; }
        t = fx::warp_scan(t, lane);
        const unsigned m_after = fx::apply(t, b.m);
        const bool over = (lane >= first && lane < nw) && (bd || m_after >= (1u << 24));
        const unsigned fail = __ballot_sync(0xffffffffu, over);
        gen++;
        if (fail == 0) {
          const unsigned m_end = __shfl_sync(0xffffffffu, m_after, nw - 1);
          s = __fmul_rn(__uint2float_rn(m_end), b.u);
          break;
        }
        const int f = __ffs(fail) - 1;
        unsigned m_prev = __shfl_up_sync(0xffffffffu, m_after, 1);
        if (lane == 0) m_prev = b.m;
        m_prev = __shfl_sync(0xffffffffu, m_prev, f);
        const unsigned g2 = gen & 1u;
        if (w == f) {  // the warp whose exit would leave the binade: exact warp-level fold from its exact entry value
          float sq = __fmul_rn(__uint2float_rn(m_prev), b.u);
          bool hq = true;
          for (int c = 0; c < CH; c++) {
            float v[8];
            unsigned vmask;
            load(c, v, vmask);
            fx::warp_fold(v, vmask, sq, hq, lane);
          }
          if (lane == 0) sm_s[g2] = sq;
        }
        __syncthreads();
        s = sm_s[g2];
        gen++;
        first = f + 1;
      }
    }
    if constexpr (EPI) {
      if (threadIdx.x == 0 && fused_apply_send<P, T, float, V>(pb, ep, vtx, have, s)) raise_flag(ep.flag);
    } else if (threadIdx.x == 0 && have) {
      y[vtx] = s;
      atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
    }
  }
}

// ------------------------------------------------------------------ driver --
#define GM_CUDA_OK(call)                                                                      \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      fprintf(stderr, "graphmat_b200: %s failed: %s (%s:%d)\n", #call, cudaGetErrorString(e_), \
              __FILE__, __LINE__);                                                            \
      return 1;                                                                               \
    }                                                                                         \
  } while (0)

// ---------------------------------------------- SpMSpV: sparse frontier (push) --
// The reference's my_spmspv walks only the columns whose x bit is set (spmspv.h:55-63): its work is
// proportional to the frontier.  These kernels do the same over the column-major companion of the
// matrix: every entry of an active column becomes a triple (row slot, position in the row's fold
// order, process_message result); the triples are radix-sorted by (slot, position) and each row's
// run is folded left to right -- the same order, hence the same bits, as the row-major kernels.
template <class P, class T, class U, class V, class E, bool NEEDVP, bool IDENT>
__global__ void __launch_bounds__(256)
    k_push_expand(prog_bytes<P> pb, gm_matrix_view M, int n_active, long long n_entries, const int* __restrict__ f_col,
                  const long long* __restrict__ f_off, const T* __restrict__ x, const V* __restrict__ vp,
                  unsigned long long* __restrict__ keys, unsigned* __restrict__ order, U* __restrict__ vals) {
  const P& prog = pb.get();
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_entries) return;
  int lo = 0, hi = n_active;  // last k with f_off[k] <= t
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(f_off + mid) <= t) lo = mid; else hi = mid;
  }
  const int c = __ldg(f_col + lo);
  const long long e = __ldg(M.c_ptr + c) + (t - __ldg(f_off + lo));
  const int slot = __ldg(M.c_row + e);
  const int rank = __ldg(M.c_rank + e);
  const E ev = __ldg(reinterpret_cast<const E*>(M.c_val) + e);
  V vprop;
  if (NEEDVP) vprop = vp[IDENT ? slot : __ldg(M.slot_vertex + slot)];
  U out;
  prog.P::process_message(x[c], ev, vprop, out);
  vals[t] = out;
  keys[t] = ((unsigned long long)(unsigned)slot << M.rank_bits) | (unsigned)rank;
  order[t] = (unsigned)t;
}

constexpr int GM_PUSH_SHORT_RUN = 16;
// one thread per sorted triple; the first triple of a row's run folds the run if it is short,
// else queues it for k_push_fold_long
template <class P, class U, bool IDENT, bool ACCUM>
__global__ void __launch_bounds__(256)
    k_push_fold(prog_bytes<P> pb, gm_matrix_view M, long long n_entries, const unsigned long long* __restrict__ keys,
                const unsigned* __restrict__ order, const U* __restrict__ vals, U* __restrict__ y,
                unsigned* __restrict__ ybits, int* __restrict__ n_long, long long* __restrict__ long_runs) {
  const P& prog = pb.get();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  const int rb = M.rank_bits;
  const unsigned long long slot = __ldg(keys + i) >> rb;
  if (i > 0 && (__ldg(keys + i - 1) >> rb) == slot) return;
  // end of the run: first index whose slot is larger
  long long lo = i, hi = n_entries;  // keys[lo] has this slot, keys[hi] (if any) does not
  if (i + GM_PUSH_SHORT_RUN < n_entries && (__ldg(keys + i + GM_PUSH_SHORT_RUN) >> rb) == slot) {
    while (hi - lo > 1) {
      const long long mid = (lo + hi) >> 1;
      if ((__ldg(keys + mid) >> rb) == slot) lo = mid; else hi = mid;
    }
    const int q = atomicAdd(n_long, 1);
    long_runs[2 * q] = i;
    long_runs[2 * q + 1] = hi;
    return;
  }
  const int vtx = IDENT ? (int)slot : __ldg(M.slot_vertex + (int)slot);
  U acc;
  bool have = false;
  if (ACCUM && ((__ldg(ybits + (vtx >> 5)) >> (vtx & 31)) & 1u)) {
    acc = y[vtx];
    have = true;
  }
  for (long long j = i; j < n_entries && j <= i + GM_PUSH_SHORT_RUN; j++) {
    if (j > i && (__ldg(keys + j) >> rb) != slot) break;
    const U v = vals[__ldg(order + j)];
    if (have) prog.P::reduce_function(acc, v);
    else { acc = v; have = true; }
  }
  y[vtx] = acc;
  atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
}

// one warp per long run.  REORDER (associative reduce): every lane folds a contiguous chunk in
// order, the 32 partials are folded in lane order.  Otherwise 32 values are loaded at a time and
// folded one by one (exact for any reduce_function).
template <class P, class U, bool IDENT, bool ACCUM, bool REORDER>
__global__ void __launch_bounds__(128)
    k_push_fold_long(prog_bytes<P> pb, gm_matrix_view M, const int* __restrict__ n_long,
                     const long long* __restrict__ long_runs, const unsigned long long* __restrict__ keys,
                     const unsigned* __restrict__ order, const U* __restrict__ vals, U* __restrict__ y,
                     unsigned* __restrict__ ybits) {
  const P& prog = pb.get();
  extern __shared__ __align__(16) unsigned char push_sm[];
  U* stage = reinterpret_cast<U*>(push_sm) + (threadIdx.x >> 5) * 32;
  const int lane = threadIdx.x & 31;
  const int nq = *n_long;
  for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nq; q += (gridDim.x * blockDim.x) >> 5) {
    const long long b = long_runs[2 * q], e = long_runs[2 * q + 1];
    const int slot = (int)(__ldg(keys + b) >> M.rank_bits);
    const int vtx = IDENT ? slot : __ldg(M.slot_vertex + slot);
    U acc;
    bool have = false;
    if (ACCUM && ((__ldg(ybits + (vtx >> 5)) >> (vtx & 31)) & 1u)) {
      acc = y[vtx];
      have = true;
    }
    if (REORDER) {
      const long long chunk = (e - b + 31) / 32;
      const long long cb = b + lane * chunk, ce = cb + chunk < e ? cb + chunk : e;
      U part;
      bool hp = false;
      for (long long j = cb; j < ce; j++) {
        const U v = vals[__ldg(order + j)];
        if (hp) prog.P::reduce_function(part, v);
        else { part = v; hp = true; }
      }
      stage[lane] = part;
      const unsigned hm = __ballot_sync(0xffffffffu, hp);
      __syncwarp();
      if (lane == 0) {
        for (int k = 0; k < 32; k++) {
          if (!((hm >> k) & 1u)) continue;
          if (have) prog.P::reduce_function(acc, stage[k]);
          else { acc = stage[k]; have = true; }
        }
      }
      __syncwarp();
    } else {
      for (long long j0 = b; j0 < e; j0 += 32) {
        const bool on = j0 + lane < e;
        if (on) stage[lane] = vals[__ldg(order + j0 + lane)];
        __syncwarp();
        if (lane == 0) {
          const int cnt = e - j0 < 32 ? (int)(e - j0) : 32;
          for (int k = 0; k < cnt; k++) {
            if (have) prog.P::reduce_function(acc, stage[k]);
            else { acc = stage[k]; have = true; }
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0 && have) {
      y[vtx] = acc;
      atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
    }
  }
}

// ------------------------- SpMSpV: sparse frontier without a sort (atomic push) --
// For programs whose reduce is min (gm_atomic_min) or "last writer" (gm_last_writer) the sorted triples of the
// path above are unnecessary:
//   MODE 2 (min): every entry of an active column does atomicMin(y[row], process_message(...)); y holds the
//     identity 0xffffffff wherever its bit is clear (k_apply<RESET> keeps that invariant).
//   MODE 0 + MODE 1 (last writer): the left fold in ascending column order keeps the contribution with the
//     LARGEST fold position among the active entries of the row.  Pass 0 takes atomicMax(win[row], position);
//     pass 1 walks the same entries again and the one whose position won writes y[row] (and hands win[row]
//     back as -1).  Positions are unique within a row, so exactly one entry writes: the same bits as the fold.
// Work split: a warp takes 32 bit words of x; every active column with at most GM_PUSH_BIG entries is walked
// by the whole warp; the longer ones (a short list fixed at build time, gm_matrix_view::big_cols) are walked by
// extra blocks of the same launch, all together.
constexpr int GM_PUSH_BIG = 2048;
constexpr int GM_PUSH_LANE = 16;
template <class P, class T, class U, class V, class E, bool NEEDVP, bool IDENT, int MODE>
__device__ __forceinline__ void push_entry(const P& prog, const gm_matrix_view& M, long long e, const T& xv,
                                           const V* __restrict__ vp, U* __restrict__ y, unsigned* __restrict__ ybits,
                                           int* __restrict__ win) {
  const int slot = __ldg(M.c_row + e);
  if constexpr (MODE == 0) {
    atomicMax(win + slot, __ldg(M.c_rank + e));
  } else {
    if constexpr (MODE == 1) {
      if (win[slot] != __ldg(M.c_rank + e)) return;
      win[slot] = -1;
    }
    const int vtx = IDENT ? slot : __ldg(M.slot_vertex + slot);
    V vprop;
    if (NEEDVP) vprop = vp[vtx];
    U out;
    prog.P::process_message(xv, __ldg(reinterpret_cast<const E*>(M.c_val) + e), vprop, out);
    if constexpr (MODE == 1) y[vtx] = out;
    else atomicMin(reinterpret_cast<unsigned*>(y) + vtx, *reinterpret_cast<const unsigned*>(&out));
    const unsigned bit = 1u << (vtx & 31);
    if (!(ybits[vtx >> 5] & bit)) atomicOr(ybits + (vtx >> 5), bit);
  }
}
template <class P, class T, class U, class V, class E, bool NEEDVP, bool IDENT, int MODE>
__global__ void __launch_bounds__(256)
    k_push_atomic(prog_bytes<P> pb, gm_matrix_view M, int n_words, int word_blocks, const unsigned* __restrict__ xbits,
                  const T* __restrict__ x, const V* __restrict__ vp, U* __restrict__ y, unsigned* __restrict__ ybits,
                  int* __restrict__ win) {
  const P& prog = pb.get();
  if ((int)blockIdx.x >= word_blocks) {
    // the blocks behind the word sweep: the (few, precomputed) columns above GM_PUSH_BIG entries, each walked
    // by all of these blocks together when its x bit is set
    const long long g = (blockIdx.x - word_blocks) * (long long)blockDim.x + threadIdx.x;
    const long long stride = (gridDim.x - word_blocks) * (long long)blockDim.x;
    __shared__ int act[1024];
    __shared__ int n_act;
    for (int base = 0; base < M.n_big_cols; base += 1024) {  // every block compacts the ACTIVE big columns itself
      if (threadIdx.x == 0) n_act = 0;
      __syncthreads();
      for (int k = base + threadIdx.x; k < M.n_big_cols && k < base + 1024; k += blockDim.x) {
        const int c = __ldg(M.big_cols + k);
        if (test_bit(xbits, c)) act[atomicAdd(&n_act, 1)] = c;
      }
      __syncthreads();
      const int na = n_act;
      for (int j = 0; j < na; j++) {
        const int c = act[j];
        const long long beg = __ldg(M.c_ptr + c), end = __ldg(M.c_ptr + c + 1);
        const T xv = x[c];
        for (long long e = beg + g; e < end; e += stride) push_entry<P, T, U, V, E, NEEDVP, IDENT, MODE>(prog, M, e, xv, vp, y, ybits, win);
      }
      __syncthreads();
    }
    return;
  }
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int w = warp * 32 + lane;
  unsigned word = w < n_words ? __ldg(xbits + w) : 0u;
  // a lane first walks the SHORT columns of its own word by itself (a sparse frontier has about one active column
  // per word: 32 words = 32 independent chains per warp); columns above GM_PUSH_LANE entries are left for the
  // whole warp, 32 entries per step
  unsigned longer = 0u;
  {
    unsigned m = word;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const int c = w * 32 + b;
      const long long beg = __ldg(M.c_ptr + c), end = __ldg(M.c_ptr + c + 1);
      if (end - beg > GM_PUSH_LANE) {
        if (end - beg <= GM_PUSH_BIG) longer |= 1u << b;
        continue;
      }
      if (beg == end) continue;
      const T xv = x[c];
      for (long long e = beg; e < end; e++) push_entry<P, T, U, V, E, NEEDVP, IDENT, MODE>(prog, M, e, xv, vp, y, ybits, win);
    }
  }
  unsigned lanes = __ballot_sync(0xffffffffu, longer != 0);
  while (lanes) {
    const int src = __ffs(lanes) - 1;
    lanes &= lanes - 1;
    unsigned m = __shfl_sync(0xffffffffu, longer, src);
    const int cbase = (warp * 32 + src) * 32;
    while (m) {
      const int c = cbase + __ffs(m) - 1;
      m &= m - 1;
      const long long beg = __ldg(M.c_ptr + c), end = __ldg(M.c_ptr + c + 1);
      const T xv = x[c];
      for (long long e = beg + lane; e < end; e += 32) push_entry<P, T, U, V, E, NEEDVP, IDENT, MODE>(prog, M, e, xv, vp, y, ybits, win);
    }
  }
}

inline int gm_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

struct step_counters {
  long long launches = 0;
  long long edges = 0;
  long long push_passes = 0;
  long long last_frontier_cols = -1, last_frontier_entries = -1;  // of the latest sparse pass (trace only)
  long long next_entries = -1;  // entries of the coming pass's frontier when k_apply counted them (-1: unknown)
  long long next_vertices = -1; // ... and its vertices
  bool last_pass_pushed = false;  // the latest SpMSpV pass took the sparse-frontier path (y is sparse)
  bool ybits_clean = false;     // k_send already cleared y's bit words for the coming pass
  bool counted_by_kernel = false;  // gm_push_count ran in this iteration (its scratch words overlap k_apply's counter)
};

template <class P>
struct engine {
  typedef decltype(deduce_types((P*)nullptr)) types;
  typedef typename types::Tm T;
  typedef typename types::Um U;
  typedef typename types::Vp V;
  typedef typename types::Ev E;
  typedef epilogue<T, V> EP;
  static constexpr bool REORDER = is_reorderable<P>::value;

  static int check(const gm_graph_view& gv, const gm_vectors_view& vv) {
    if ((int)sizeof(V) != gv.sizeof_V || (int)sizeof(E) != gv.sizeof_E || (int)sizeof(T) != vv.sizeof_T ||
        (int)sizeof(U) != vv.sizeof_U) {
      fprintf(stderr, "graphmat_b200: type sizes do not match the graph/vectors (V %zu/%d E %zu/%d T %zu/%d U %zu/%d)\n",
              sizeof(V), gv.sizeof_V, sizeof(E), gv.sizeof_E, sizeof(T), vv.sizeof_T, sizeof(U), vv.sizeof_U);
      return 1;
    }
    return 0;
  }

  // IntersectReduce(active, vertexproperty, &x, send_message)   GraphMatRuntime.h:145
  static int send(const P& prog, const gm_graph_view& gv, const gm_vectors_view& vv, step_counters* sc) {
    cudaStream_t st = (cudaStream_t)gv.stream;
    const int n = gv.n_local_pad;
    T* xloc = reinterpret_cast<T*>(vv.x_val) + (size_t)gv.rank * n;
    unsigned* xb = vv.x_bits + (size_t)gv.rank * (n >> 5);
    // ACTIVE_ONLY programs sweep the bit words.  (Switching to one thread per vertex for dense frontiers / after
    // row-major passes was measured: BFS RMAT-22 0.885 -> 0.85 ms, but RMAT-26 4.22 -> 5.25 ms; not kept.)
    if (prog.getActivity() != GraphMat::ALL_VERTICES) {
      const int nw = n >> 5;
      k_send_words<P, T, V><<<((nw + 31) / 32 + 7) / 8, 256, 0, st>>>(pack(prog), nw, (const V*)gv.vertexproperty, gv.active_bits,
                                                                      xloc, xb, sc ? vv.y_bits : nullptr);
    } else {
      k_send<P, T, V><<<(n + 255) / 256, 256, 0, st>>>(pack(prog), n, (const V*)gv.vertexproperty, gv.active_bits, xloc, xb,
                                                       sc ? vv.y_bits : nullptr);
    }
    if (sc) { sc->launches++; sc->ybits_clean = true; }
    GM_CUDA_OK(cudaGetLastError());
    return 0;
  }

  // heavy rows.  fp32-sum programs: exact parallel fold, [0, n_coop) one thread block per row,
  // [n_coop, n_heavy) one warp per row.  Associative programs: two-phase segmented fold.
  // Anything else: one warp per row, batches folded serially (exact for any reduce_function).
  // EPI: the kernel that finishes a row also applies and sends (fused_apply_send); y is not written.
  // hs[0]: the latency-bound chains (staged long rows, block-per-row), hs[1]: the warp-per-row kernel -- on separate
  // streams when the graph has them, so that their blocks fill the SMs together with the sliced-ELL kernels
  template <bool ALLACT, bool NEEDVP, bool IDENT, bool ACCUM, bool EPI>
  static int heavy_rows(const prog_bytes<P>& pb, const gm_matrix_view& M, int hot, const T* x, const unsigned* xbits,
                        const V* vp, U* y, unsigned* ybits, cudaStream_t const (&hs)[2], step_counters* sc, gm_vectors* vecs,
                        const EP& ep, cudaEvent_t staged_ev = nullptr, bool* staged_recorded = nullptr) {
    constexpr bool FADD = is_fadd32<P>::value && std::is_same<U, float>::value && !NEEDVP && !ACCUM;
    cudaStream_t st = hs[0];
    if constexpr (FADD) {
      const int n_coop = M.n_coop;
      int coop_begin = 0;
      if constexpr (ALLACT && sizeof(T) == 4) {
        // the longest rows: gathers by the whole GPU into a staging array, then one block per row folds from it,
        // streaming the array through TMA into shared memory (k_heavy_fadd32_tma)
        static const bool no_stage = getenv("GM_NO_STAGE") != nullptr;
        if (M.n_long > 0 && !no_stage) {
          void* scratch = nullptr;
          if (gm_vectors_scratch(vecs, (M.long_entries + 64 + (GM_TMA_BUFS + 1) * GM_TMA_ROUND) * (long long)sizeof(T), &scratch)) return 1;
          T* staged = (T*)scratch;
          const long long groups = (M.long_entries + 7) / 8;
          k_stage_rows<T><<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(M.h_col, M.long_entries, x, hot, staged);
          if (staged_ev) {  // the caller holds the sliced-ELL kernels back until here: see mult_t
            GM_CUDA_OK(cudaEventRecord(staged_ev, st));
            *staged_recorded = true;
          }
          auto kt = k_heavy_fadd32_tma<P, T, V, IDENT, EPI>;
          constexpr int ring_bytes = GM_TMA_BUFS * GM_TMA_ROUND * 4;
          static bool attr = false;
          if (!attr) {
            GM_CUDA_OK(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes));
            attr = true;
          }
          kt<<<M.n_long, GM_TMA_W * 32, ring_bytes, st>>>(pb, M, 0, M.n_long, (const float*)staged, (float*)y, ybits, ep);
          if (sc) sc->launches += 2;
          coop_begin = M.n_long;
        }
      }
      if (n_coop > coop_begin) {
        const int cap = gm_sm_count() * 64;
        int blocks = n_coop - coop_begin < cap ? n_coop - coop_begin : cap;
        k_heavy_fadd32<P, T, V, E, ALLACT, IDENT, 16, EPI><<<blocks, 16 * 32, 0, st>>>(
            pb, M, coop_begin, n_coop, hot, x, xbits, (float*)y, ybits, ep);
        if (sc) sc->launches++;
      }
      if (M.n_heavy > n_coop) {
        int rows = M.n_heavy - n_coop;
#ifdef GM_EXPERIMENTS
        if (getenv("GM_DBG_SERIALIZE_H1") && hs[1] != hs[0]) {  // heavy1 on its own stream but after the block-per-row kernel
          static cudaEvent_t dbg_ev = nullptr;
          if (!dbg_ev) cudaEventCreateWithFlags(&dbg_ev, cudaEventDisableTiming);
          cudaEventRecord(dbg_ev, hs[0]);
          cudaStreamWaitEvent(hs[1], dbg_ev, 0);
        }
#endif
        k_heavy_fadd32<P, T, V, E, ALLACT, IDENT, 1, EPI><<<(rows + 3) / 4, 128, 0, hs[1]>>>(
            pb, M, n_coop, M.n_heavy, hot, x, xbits, (float*)y, ybits, ep);
        if (sc) sc->launches++;
      }
    } else if constexpr (REORDER) {
      if (M.n_segs > 0 || EPI) {
        void* scratch = nullptr;
        const size_t pbytes = ((size_t)M.n_segs * sizeof(U) + 255) & ~(size_t)255;
        if (gm_vectors_scratch(vecs, (long long)(pbytes + M.n_segs + 256), &scratch)) return 1;
        U* partial = (U*)scratch;
        unsigned char* pvalid = (unsigned char*)scratch + pbytes;
        const size_t sh = 4 * 32 * sizeof(U);
        constexpr bool LASTW = is_last_writer<P>::value && !ALLACT;
        auto kseg = k_heavy_seg<P, T, U, V, E, ALLACT, NEEDVP, IDENT, LASTW>;
        auto kcmb = k_heavy_combine<P, T, U, V, IDENT, ACCUM, EPI>;
        if (big_smem(kseg, sh) || big_smem(kcmb, sh)) return 1;
        if (M.n_segs > 0) kseg<<<(M.n_segs + 3) / 4, 128, sh, st>>>(pb, M, hot, x, xbits, vp, partial, pvalid);
        kcmb<<<(M.n_heavy + 3) / 4, 128, sh, st>>>(pb, M, partial, pvalid, y, ybits, ep);
        if (sc) sc->launches += 2;
      }
    } else {
      size_t sh = 4 * 32 * sizeof(U);
      auto kh = k_heavy<P, T, U, V, E, ALLACT, NEEDVP, IDENT, ACCUM, false, EPI>;
      if (big_smem(kh, sh)) return 1;
      kh<<<(M.n_heavy + 3) / 4, 128, sh, st>>>(pb, M, 0, hot, x, xbits, vp, y, ybits, ep);
      if (sc) sc->launches++;
    }
    return 0;
  }

  // kernels whose dynamic shared memory is 128 reduced messages: above the 48 KB default (sizeof(U) > 384,
  // e.g. LatentVector<K> with K >= 48) the limit is raised explicitly, and refused beyond the hardware's 227 KB
  template <class K>
  static int big_smem(K kern, size_t bytes) {
    if (bytes <= 48 * 1024) return 0;
    if (bytes > 227 * 1024) {
      fprintf(stderr, "graphmat_b200: reduced message type of %zu bytes is too large for the long-row kernels\n", sizeof(U));
      return 1;
    }
    GM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
  }

  template <bool ALLACT, bool NEEDVP, bool IDENT, bool ACCUM, bool EPI>
  static int mult_t(const P& prog, const gm_graph_view& gv, const gm_matrix_view& M, const gm_vectors_view& vv,
                    step_counters* sc, gm_vectors* vecs, const EP& ep) {
    cudaStream_t st = (cudaStream_t)gv.stream;
    const T* x = (const T*)vv.x_val;
    const V* vp = (const V*)gv.vertexproperty;
    U* y = (U*)vv.y_val;
    prog_bytes<P> pb = pack(prog);
    const int hot = gv.hot_limit;
    // sparse frontier: walk only the active columns when they hold few entries (push)
    if constexpr (!ALLACT && sizeof(U) <= 16 && sizeof(E) == 4) {
      const int push_div = gv.push_divisor;  // 0: never
      const long long push_min = gv.push_min_nnz;
      if (push_div > 0 && M.nnz >= push_min && gv.owner) {
        const int which = (&M == &gv.A) ? 0 : 1;
        int n_act = -1;
        long long n_ent = sc ? sc->next_entries : -1;  // counted by the previous k_apply (one GPU)
        if (n_ent < 0 || !M.c_ptr) {
          if (gm_push_count(gv.owner, which, vecs, &n_act, &n_ent)) return 1;
          if (sc) sc->counted_by_kernel = true;
        }
        if (sc) { sc->last_frontier_cols = n_act; sc->last_frontier_entries = n_ent; }
        if (n_ent * push_div <= M.nnz) {
          if (n_ent == 0) {  // nothing arrives: y stays cleared
            if (sc) { sc->push_passes++; sc->last_pass_pushed = true; }
            return 0;
          }
          gm_graph_view gv2;
          if (gm_graph_view_get(gv.owner, &gv2)) return 1;  // the companion may just have been built
          const gm_matrix_view& MP = which == 0 ? gv2.A : gv2.AT;
          constexpr bool LASTW = is_last_writer<P>::value;
          constexpr bool AMIN = is_atomic_min<P>::value && sizeof(U) == 4;
          bool sorted_path = true;
          if constexpr (!ACCUM && (LASTW || AMIN)) sorted_path = getenv("GM_NO_ATOMIC_PUSH") != nullptr;  // tests
          if constexpr (!ACCUM && (LASTW || AMIN)) if (!sorted_path) {
            // no sort, no host round trip: atomicMin into y, or "largest fold position wins" in two sweeps
            void* aux = nullptr;
            if (LASTW && gm_vectors_aux(vecs, (long long)MP.n_slots * 4, &aux)) return 1;
            int* win = (int*)aux;
            const int n_words = gv.n_full >> 5;
            const int wblocks = ((n_words + 31) / 32 + 7) / 8;
            const unsigned blocks = (unsigned)(wblocks + (MP.n_big_cols > 0 ? gm_sm_count() * 4 : 0));
            if constexpr (LASTW) {
              k_push_atomic<P, T, U, V, E, NEEDVP, IDENT, 0><<<blocks, 256, 0, st>>>(pb, MP, n_words, wblocks, vv.x_bits, x, vp, y, vv.y_bits, win);
              k_push_atomic<P, T, U, V, E, NEEDVP, IDENT, 1><<<blocks, 256, 0, st>>>(pb, MP, n_words, wblocks, vv.x_bits, x, vp, y, vv.y_bits, win);
              if (sc) sc->launches += 2;
            } else {
              k_push_atomic<P, T, U, V, E, NEEDVP, IDENT, 2><<<blocks, 256, 0, st>>>(pb, MP, n_words, wblocks, vv.x_bits, x, vp, y, vv.y_bits, nullptr);
              if (sc) sc->launches += 1;
            }
            if (sc) { sc->edges += n_ent; sc->push_passes++; sc->last_pass_pushed = true; }
            GM_CUDA_OK(cudaGetLastError());
            return 0;
          }
          if (sorted_path) {
            // any other reduce_function: (row, fold position, value) triples, radix-sorted, folded in the reference's order
            if (n_act < 0) {
              if (gm_push_count(gv.owner, which, vecs, &n_act, &n_ent)) return 1;
              if (sc) sc->counted_by_kernel = true;
            }
            if (n_ent == 0) {
              if (sc) { sc->push_passes++; sc->last_pass_pushed = true; }
              return 0;
            }
            gm_push_plan plan;
            if (gm_push_prepare(gv.owner, which, vecs, n_act, n_ent, &plan)) return 1;
            const unsigned blocks = (unsigned)((n_ent + 255) / 256);
            k_push_expand<P, T, U, V, E, NEEDVP, IDENT><<<blocks, 256, 0, st>>>(pb, MP, n_act, n_ent, plan.f_col, plan.f_off, x, vp,
                                                                                plan.keys, plan.order, (U*)plan.vals);
            if (gm_push_sort(gv.owner, &plan)) return 1;
            // the long-run queue reuses the spare key buffer (at most n_ent / 17 runs of 2 entries each)
            int* n_long = gv.d_flags + 13;
            long long* long_runs = reinterpret_cast<long long*>(plan.keys_alt);
            GM_CUDA_OK(cudaMemsetAsync(n_long, 0, sizeof(int), st));
            k_push_fold<P, U, IDENT, ACCUM><<<blocks, 256, 0, st>>>(pb, MP, n_ent, plan.keys, plan.order, (const U*)plan.vals, y,
                                                                    vv.y_bits, n_long, long_runs);
            const int lb = gm_sm_count() * 4;
            auto kfl = k_push_fold_long<P, U, IDENT, ACCUM, REORDER>;
            if (big_smem(kfl, 4 * 32 * sizeof(U))) return 1;
            kfl<<<lb, 128, 4 * 32 * sizeof(U), st>>>(pb, MP, n_long, long_runs, plan.keys, plan.order, (const U*)plan.vals, y,
                                                     vv.y_bits);
            if (sc) { sc->launches += 5; sc->edges += n_ent; sc->push_passes++; sc->last_pass_pushed = true; }
            GM_CUDA_OK(cudaGetLastError());
            return 0;
          }
        }
      }
    }
    // heavy rows and sliced-ELL rows are disjoint (rows and y words): their kernels run concurrently, on up to four
    // streams -- every one of them alone leaves the L1->L2 request port of the SMs partly idle (latency-bound tails)
    cudaStream_t aux[3] = {(cudaStream_t)gv.aux_stream, (cudaStream_t)gv.aux_stream2, (cudaStream_t)gv.aux_stream3};
    cudaEvent_t joins[3] = {(cudaEvent_t)gv.ev_join, (cudaEvent_t)gv.ev_join2, (cudaEvent_t)gv.ev_join3};
    bool used[3] = {false, false, false};
    auto on_aux = [&](int k) -> cudaStream_t {  // stream k if the graph has it, else the nearest lower one
      for (int j = k; j >= 0; j--)
        if (aux[j]) { used[j] = true; return aux[j]; }
      return st;
    };
    const bool fork = aux[0] != nullptr && (M.n_heavy > 0) + (M.n_slices > 0) >= 1;
    if (fork) GM_CUDA_OK(cudaEventRecord((cudaEvent_t)gv.ev_fork, st));
    auto enter = [&](cudaStream_t s2) -> int {
      if (s2 != st) GM_CUDA_OK(cudaStreamWaitEvent(s2, (cudaEvent_t)gv.ev_fork, 0));
      return 0;
    };
    if (M.n_heavy > 0) {
      const bool both = M.n_slices > 0;
      cudaStream_t hs[2] = {both ? on_aux(0) : st, both ? on_aux(1) : (aux[0] ? on_aux(0) : st)};
      if (enter(hs[0]) || (hs[1] != hs[0] && enter(hs[1]))) return 1;
      // The 1024-thread blocks of the staged long-row fold each need a whole SM.  Enqueued behind thousands of
      // sliced-ELL blocks they would only start when those drain (measured: the pass of rank 0 of 8 then lasts
      // the SUM of the two chains).  So the sliced-ELL kernels wait for the staging gather (which fills the GPU by
      // itself) and the high-priority long-row blocks are placed first, on empty SMs.
      bool staged_recorded = false;
      cudaEvent_t mid = (hs[0] != st && !getenv("GM_NO_HOLD")) ? (cudaEvent_t)gv.ev_join2 : nullptr;
      if (heavy_rows<ALLACT, NEEDVP, IDENT, ACCUM, EPI>(pb, M, hot, x, vv.x_bits, vp, y, vv.y_bits, hs, sc, vecs, ep, mid,
                                                         &staged_recorded))
        return 1;
      if (staged_recorded) GM_CUDA_OK(cudaStreamWaitEvent(st, mid, 0));
    }
    if (M.n_slices > 0) {
      // wide slices: one per warp, deep unroll (a lane's chain waits for loads once per UNROLL
      // steps); narrow tail: several slices per warp so blocks stay worth their launch
      constexpr int UW = (sizeof(T) <= 8 && sizeof(U) <= 8) ? 16 : 1;
      constexpr int UN = (sizeof(T) <= 8 && sizeof(U) <= 8) ? 8 : 1;
      static const int spw_tail = getenv("GM_SELL_SPW") ? atoi(getenv("GM_SELL_SPW")) : 8;
      const int nw = M.n_slices_wide < M.n_slices ? M.n_slices_wide : M.n_slices;
      constexpr bool LASTW = is_last_writer<P>::value && !ALLACT && sizeof(T) <= 8 && sizeof(U) <= 8;
      auto launch_wide = [&]() -> int {
        if (nw <= 0) return 0;
        if constexpr (LASTW)
          k_sell_last<P, T, U, V, E, NEEDVP, IDENT, ACCUM, 8><<<(nw + 7) / 8, 256, 0, st>>>(
              pb, M, 0, nw, 1, hot, x, vv.x_bits, vp, y, vv.y_bits);
        else
          k_sell<P, T, U, V, E, ALLACT, NEEDVP, IDENT, ACCUM, UW, EPI><<<(nw + 7) / 8, 256, 0, st>>>(
              pb, M, 0, nw, 1, hot, x, vv.x_bits, vp, y, vv.y_bits, ep);
        if (sc) sc->launches++;
        return 0;
      };
      auto launch_tail = [&]() -> int {
        if (M.n_slices <= nw) return 0;
        const int warps = (M.n_slices - nw + spw_tail - 1) / spw_tail;
        cudaStream_t ts = (nw > 0 && aux[2]) ? on_aux(2) : st;  // the narrow tail beside the wide slices
        if (enter(ts)) return 1;
        if constexpr (LASTW)
          k_sell_last<P, T, U, V, E, NEEDVP, IDENT, ACCUM, 8><<<(warps + 7) / 8, 256, 0, ts>>>(
              pb, M, nw, M.n_slices, spw_tail, hot, x, vv.x_bits, vp, y, vv.y_bits);
        else
          k_sell<P, T, U, V, E, ALLACT, NEEDVP, IDENT, ACCUM, UN, EPI><<<(warps + 7) / 8, 256, 0, ts>>>(
              pb, M, nw, M.n_slices, spw_tail, hot, x, vv.x_bits, vp, y, vv.y_bits, ep);
        if (sc) sc->launches++;
        return 0;
      };
      // The narrow tail holds most of the ROWS, i.e. most of the messages the fused epilogue stores into the peers.
      // Last in the pass, those stores bunch up behind it (8 GPUs: 224 MB per rank in the final ~0.15 ms, more than
      // NVLink takes); first, they travel under the wide slices.  Measured: 8 GPUs 1179 -> 1247 GTEPS, 4 GPUs 694 -> 786,
      // but 2 GPUs 461 -> 417 (one peer: the stores are no bottleneck and the wide slices lose their head start), so
      // from four ranks on.
      static const int tail_first_env = getenv("GM_TAIL_FIRST") ? atoi(getenv("GM_TAIL_FIRST")) : -1;
      const bool tail_first = EPI && (tail_first_env < 0 ? ep.n_peers >= 3 : (tail_first_env != 0 && ep.n_peers > 0));
      if (tail_first) {
        if (launch_tail() || launch_wide()) return 1;
      } else {
        if (launch_wide() || launch_tail()) return 1;
      }
    }
    for (int k = 0; k < 3; k++)
      if (used[k]) {
        GM_CUDA_OK(cudaEventRecord(joins[k], aux[k]));
        GM_CUDA_OK(cudaStreamWaitEvent(st, joins[k], 0));
      }
    if (sc) { sc->edges += M.nnz; sc->last_pass_pushed = false; }
    GM_CUDA_OK(cudaGetLastError());
    return 0;
  }

  // mult_segment / mult_segment3 for one operand matrix   SPMV.h:62-95
  // ep != nullptr: fused apply+send epilogue (only all-active, single-operand passes)
  static int mult(const P& prog, const gm_graph_view& gv, const gm_matrix_view& M, const gm_vectors_view& vv,
                  bool allact, bool accum, step_counters* sc, gm_vectors* vecs, const EP* ep = nullptr) {
    const bool needvp = prog.getProcessMessageRequiresVertexprop();
    const bool ident = M.identity != 0;
    const EP none = EP();
    if (ep) {
      if (!allact || accum) {
        fprintf(stderr, "graphmat_b200: fused epilogue on a pass that is not all-active / single-operand\n");
        return 1;
      }
      if (needvp) {
        if (ident) return mult_t<true, true, true, false, true>(prog, gv, M, vv, sc, vecs, *ep);
        return mult_t<true, true, false, false, true>(prog, gv, M, vv, sc, vecs, *ep);
      }
      if (ident) return mult_t<true, false, true, false, true>(prog, gv, M, vv, sc, vecs, *ep);
      return mult_t<true, false, false, false, true>(prog, gv, M, vv, sc, vecs, *ep);
    }
#define GM_MULT(A_, N_, I_, C_) return mult_t<A_, N_, I_, C_, false>(prog, gv, M, vv, sc, vecs, none)
#define GM_MULT_C(A_, N_, I_) \
  do { if (accum) GM_MULT(A_, N_, I_, true); else GM_MULT(A_, N_, I_, false); } while (0)
#define GM_MULT_I(A_, N_) \
  do { if (ident) GM_MULT_C(A_, N_, true); else GM_MULT_C(A_, N_, false); } while (0)
#define GM_MULT_N(A_) \
  do { if (needvp) GM_MULT_I(A_, true); else GM_MULT_I(A_, false); } while (0)
    if (allact) GM_MULT_N(true);
    else GM_MULT_N(false);
#undef GM_MULT
#undef GM_MULT_C
#undef GM_MULT_I
#undef GM_MULT_N
    return 0;
  }

  // SpMTSpV / SpMSpV selection   GraphMatRuntime.h:160-176
  static int spmspv(const P& prog, const gm_graph_view& gv, const gm_vectors_view& vv, bool allact, step_counters* sc,
                    gm_vectors* vecs, const EP* ep = nullptr) {
    cudaStream_t st = (cudaStream_t)gv.stream;
    const int order = (int)prog.getOrder();
    if (!ep && !(sc && sc->ybits_clean))
      GM_CUDA_OK(cudaMemsetAsync(vv.y_bits, 0, (size_t)(gv.n_local_pad >> 5) * 4, st));  // Clear(&y)
    if (sc) sc->ybits_clean = false;
    if (order == GraphMat::OUT_EDGES) return mult(prog, gv, gv.AT, vv, allact, false, sc, vecs, ep);
    if (order == GraphMat::IN_EDGES) return mult(prog, gv, gv.A, vv, allact, false, sc, vecs, ep);
    if (order == GraphMat::ALL_EDGES) {
      if (mult(prog, gv, gv.AT, vv, allact, false, sc, vecs)) return 1;
      return mult(prog, gv, gv.A, vv, allact, true, sc, vecs);
    }
    printf("Unrecognized option \n");
    exit(1);
  }

  // true when P inherits GraphProgram's empty do_every_iteration: program state cannot change
  // between apply and the next send, so the two may share a kernel
  static constexpr bool NO_HOOK =
      std::is_same<decltype(&P::do_every_iteration), void (GraphMat::GraphProgram<T, U, V, E>::*)(int)>::value;

  static int apply_blocks(int n) {
    constexpr int per = 256 * ((sizeof(V) + sizeof(U) <= 32) ? GM_APPLY_VPT : 1);
    return (n + per - 1) / per;
  }
  // the apply loop   GraphMatRuntime.h:184-226 (flag = !converged)
  static int apply(P& prog, const gm_graph_view& gv, const gm_vectors_view& vv, step_counters* sc, bool fuse_send = false,
                   bool count_next = false, bool* counted = nullptr) {
    cudaStream_t st = (cudaStream_t)gv.stream;
    const int n = gv.n_local_pad;
    T* xloc = reinterpret_cast<T*>(vv.x_val) + (size_t)gv.rank * n;
    unsigned* xb = vv.x_bits + (size_t)gv.rank * (n >> 5);
    constexpr bool RESET = is_atomic_min<P>::value && sizeof(U) == 4;
    if (fuse_send) {
      k_apply<P, T, U, V, true, RESET><<<apply_blocks(n), 256, 0, st>>>(pack(prog), gv.n_local, n, (U*)vv.y_val, vv.y_bits,
                                                                       (V*)gv.vertexproperty, gv.active_bits, gv.d_flags, xloc, xb);
    } else {
      // the vertices that change are the next frontier: count the entries of their columns on the way (one GPU,
      // single-operand ACTIVE_ONLY programs with the column-major companion built)
      const long long* c_ptr = nullptr;
      const int order = (int)prog.getOrder();
      if (count_next && gv.world == 1 && order != GraphMat::ALL_EDGES)
        c_ptr = order == GraphMat::OUT_EDGES ? gv.AT.c_ptr : gv.A.c_ptr;
      unsigned long long* next = reinterpret_cast<unsigned long long*>(gv.d_flags + 10);
      if (c_ptr && (!sc || sc->counted_by_kernel)) GM_CUDA_OK(cudaMemsetAsync(next - 1, 0, 2 * sizeof(unsigned long long), st));
      if (sc) sc->counted_by_kernel = false;
      if (prog.getActivity() != GraphMat::ALL_VERTICES) {
        const int nw = n >> 5;
        k_apply_words<P, T, U, V, RESET><<<((nw + 31) / 32 + 7) / 8, 256, 0, st>>>(
            pack(prog), nw, (U*)vv.y_val, vv.y_bits, (V*)gv.vertexproperty, gv.active_bits, gv.d_flags, c_ptr, next, gv.rank * n);
      } else {
        k_apply<P, T, U, V, false, RESET><<<apply_blocks(n), 256, 0, st>>>(pack(prog), gv.n_local, n, (U*)vv.y_val, vv.y_bits,
                                                                          (V*)gv.vertexproperty, gv.active_bits, gv.d_flags, xloc, xb,
                                                                          c_ptr, next, gv.rank * n);
      }
      if (counted) *counted = c_ptr != nullptr;
    }
    if (sc) sc->launches++;
    GM_CUDA_OK(cudaGetLastError());
    return 0;
  }

  static int set_all_active(const gm_graph_view& gv, step_counters* sc) {
    cudaStream_t st = (cudaStream_t)gv.stream;
    int words = gv.n_local_pad >> 5;
    k_fill_bits<<<(words + 255) / 256, 256, 0, st>>>(gv.active_bits, gv.n_local, gv.n_local_pad);
    if (sc) sc->launches++;
    GM_CUDA_OK(cudaGetLastError());
    return 0;
  }

  // The exchange of SURVEY 8e: every rank's slice of x (values and/or bit words) reaches every rank.
  //   peers mapped (gm_graph_enable_peers): one kernel stores the slice into every peer's x -- dense for
  //     all-active programs, bit words + ACTIVE values only otherwise -- then the barrier kernel;
  //   else the host-supplied all-gather callbacks (NCCL through torch.distributed).
  // Every iteration of a sharded run ends with a barrier, so nobody still gathers from x when this overwrites it.
  static int exchange(gm_graph* g, gm_vectors* tmp, const gm_vectors_view& vv, bool dense, bool bits, step_counters* sc) {
    if (vv.n_peers > 0) {
      if (gm_graph_push_x(g, tmp, dense ? 1 : 0)) return 1;
      if (gm_graph_peer_barrier(g, 0)) return 1;
      if (sc) sc->launches += 2;
      return 0;
    }
    return gm_graph_exchange_x_parts(g, tmp, 1, bits ? 1 : 0);
  }

  // events / vectors the run owns: released on every exit path
  struct run_guard {
    cudaEvent_t e0 = nullptr, e1 = nullptr, s0 = nullptr, s1 = nullptr;
    std::vector<cudaEvent_t> evs;
    gm_vectors* own = nullptr;
    ~run_guard() {
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
      if (s0) cudaEventDestroy(s0);
      if (s1) cudaEventDestroy(s1);
      for (auto& e : evs) cudaEventDestroy(e);
      if (own) gm_vectors_destroy(own);
    }
  };

  // run_graph_program   GraphMatRuntime.h:93-279
  static int run(P& prog, gm_graph* g, int iterations, gm_vectors* tmp, gm_run_stats* stats) {
    gm_graph_view gv;
    if (gm_graph_view_get(g, &gv)) return 1;
    run_guard rg;
    if (!tmp) {
      if (gm_vectors_create(&rg.own, g, (int)sizeof(T), (int)sizeof(U))) return 1;
      tmp = rg.own;
    }
    gm_vectors_view vv;
    if (gm_vectors_view_get(tmp, &vv)) return 1;
    if (check(gv, vv)) return 1;
    cudaStream_t st = (cudaStream_t)gv.stream;
    step_counters sc;
    GM_CUDA_OK(cudaEventCreate(&rg.e0));
    GM_CUDA_OK(cudaEventCreate(&rg.e1));
    GM_CUDA_OK(cudaEventCreate(&rg.s0));
    GM_CUDA_OK(cudaEventCreate(&rg.s1));
    float ms_spmv = 0.f;
    const bool all = prog.getActivity() == GraphMat::ALL_VERTICES;
    const bool fuse = all && NO_HOOK && !getenv("GM_NO_FUSE");
    // fused apply+send epilogue: one operand matrix per pass, so the thread that ends a row's fold owns its message
    const bool epi = fuse && (int)prog.getOrder() != GraphMat::ALL_EDGES && !getenv("GM_NO_EPILOGUE");
    const bool peers = vv.n_peers > 0;
    if (epi && iterations != 1) {
      if (gm_vectors_need_alt(tmp)) return 1;
      if (gm_vectors_view_get(tmp, &vv)) return 1;
    }
    const int order = (int)prog.getOrder();
    const bool may_push = !all && gv.push_divisor > 0 && sizeof(U) <= 16 && sizeof(E) == 4;
    if (may_push && gv.world == 1 && order != GraphMat::ALL_EDGES) {
      // one GPU: k_apply counts the next frontier's entries, which needs the column-major companion's c_ptr now
      const gm_matrix_view& M0 = order == GraphMat::OUT_EDGES ? gv.AT : gv.A;
      if (M0.nnz >= gv.push_min_nnz && !M0.c_ptr) {
        if (gm_graph_push_ready(g, order == GraphMat::OUT_EDGES ? 1 : 0)) return 1;
        if (gm_graph_view_get(g, &gv)) return 1;
      }
    }
    GM_CUDA_OK(cudaEventRecord(rg.e0, st));
    if (may_push && is_atomic_min<P>::value && sizeof(U) == 4)  // identity of min everywhere: see k_apply<RESET>
      GM_CUDA_OK(cudaMemsetAsync(vv.y_val, 0xff, (size_t)gv.n_local_pad * sizeof(U), st));
    if (all && set_all_active(gv, &sc)) return 1;
    int it = 0, converged = 1;
    // Fixed iteration count and no do_every_iteration hook: nothing on the host depends on the
    // "changed" flag between iterations (GraphMatRuntime.h:254-260 tests it only when
    // iterations <= 0), so the whole run is enqueued without a host round trip per iteration and
    // the flag of the last iteration is read once at the end.
    const bool async = iterations > 0 && NO_HOOK && !getenv("GM_SYNC_LOOP");
    constexpr int EV_BATCH = 64;
    std::vector<cudaEvent_t>& evs = rg.evs;
    const bool timing = stats != nullptr;
    auto harvest = [&](int count) -> int {
      if (count <= 0) return 0;
      GM_CUDA_OK(cudaEventSynchronize(evs[2 * (count - 1) + 1]));
      for (int k = 0; k < count; k++) {
        float t;
        GM_CUDA_OK(cudaEventElapsedTime(&t, evs[2 * k], evs[2 * k + 1]));
        ms_spmv += t;
      }
      return 0;
    };
    if (async && timing) {
      evs.resize(2 * (size_t)std::min(iterations, EV_BATCH), nullptr);
      for (auto& e : evs) GM_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDefault));
    }
    int pending = 0;
    bool counted = false;
    void* xbuf[2] = {vv.x_val, vv.x_alt};
    int cur = 0;
    while (1) {
      GM_CUDA_OK(cudaMemsetAsync(gv.d_flags, 0, 12 * sizeof(int), st));  // [0] changed, [8..11] frontier counters
      gm_vectors_view vc = vv;  // this iteration's message vector
      vc.x_val = xbuf[cur];
      EP ep;
      if (epi) {
        if (it == 0) {
          if (send(prog, gv, vv, &sc)) return 1;
          if (gv.world > 1 && exchange(g, tmp, vv, true, true, &sc)) return 1;
          // rows without entries are never visited by the pass: their (constant) message is copied once
          if (xbuf[1]) GM_CUDA_OK(cudaMemcpyAsync(xbuf[1], xbuf[0], (size_t)gv.n_full * sizeof(T), cudaMemcpyDeviceToDevice, st));
        }
        void* const* peer_next = cur == 0 ? vv.peer_x_alt : vv.peer_x_val;
        ep.vp = (V*)gv.vertexproperty;
        ep.x_next = xbuf[1] ? (T*)xbuf[cur ^ 1] : nullptr;  // single iteration: the next message is never read
        ep.n_peers = (peers && xbuf[1]) ? vv.n_peers : 0;
        for (int q = 0; q < ep.n_peers; q++) ep.x_peer[q] = (T*)peer_next[q];
        ep.x_off = gv.rank * gv.n_local_pad;
        ep.n_valid = gv.n_local;
        ep.flag = gv.d_flags;
      } else {
        if (!(fuse && it > 0) && send(prog, gv, vv, &sc)) return 1;  // fused: the previous apply already sent
        // fused ALL_VERTICES programs re-arm every x bit in every iteration: the bit words are exchanged once
        if (gv.world > 1 && exchange(g, tmp, vv, all, !(fuse && it > 0), &sc)) return 1;
      }
      cudaEvent_t es0 = rg.s0, es1 = rg.s1;
      if (async && timing) {
        es0 = evs[2 * pending];
        es1 = evs[2 * pending + 1];
      }
      if (timing) GM_CUDA_OK(cudaEventRecord(es0, st));
      if (spmspv(prog, gv, vc, all, &sc, tmp, epi ? &ep : nullptr)) return 1;
      if (timing) GM_CUDA_OK(cudaEventRecord(es1, st));
      if (epi) {
        if (xbuf[1]) {
          if (gv.world > 1 && !peers && gm_graph_exchange_buffer(g, xbuf[cur ^ 1], (long long)gv.n_local_pad * sizeof(T))) return 1;
          cur ^= 1;
        }
      } else {
        if (apply(prog, gv, vv, &sc, fuse, may_push, &counted)) return 1;
      }
      // every iteration of a run over mapped peers ends with the barrier: it orders the stores into the
      // peers' buffers before their next pass, and it ORs the "changed" flag on the device
      if (peers && gv.world > 1) {
        if (gm_graph_peer_barrier(g, 1)) return 1;
        sc.launches++;
      }
      if (async) {
        if (timing && ++pending == EV_BATCH) {
          if (harvest(pending)) return 1;
          pending = 0;
        }
      } else {
        GM_CUDA_OK(cudaMemcpyAsync(gv.h_flags, gv.d_flags, 12 * sizeof(int), cudaMemcpyDeviceToHost, st));
        GM_CUDA_OK(cudaStreamSynchronize(st));
        int changed = gv.h_flags[0];
        sc.next_entries = counted ? (long long)*reinterpret_cast<unsigned long long*>(gv.h_flags + 10) : -1;
        sc.next_vertices = counted ? (long long)*reinterpret_cast<unsigned long long*>(gv.h_flags + 8) : -1;
        if (gv.world > 1 && !peers && gm_graph_allreduce_or(g, &changed)) return 1;
        converged = !changed;
        if (timing) {
          float t;
          GM_CUDA_OK(cudaEventElapsedTime(&t, rg.s0, rg.s1));
          ms_spmv += t;
          static const bool trace = getenv("GM_TRACE_ITERS") != nullptr;
          if (trace)
            fprintf(stderr, "graphmat_b200: iteration %d: SpMSpV %.3f ms, frontier %lld columns / %lld entries, push passes so far %lld\n",
                    it, t, sc.last_frontier_cols, sc.last_frontier_entries, sc.push_passes);
        }
        prog.do_every_iteration(it);
      }
      if (all && !fuse && set_all_active(gv, &sc)) return 1;
      it++;
      if (it == iterations) break;
      if (iterations <= 0 && converged) break;
    }
    if (epi && cur == 1) {
      // leave the latest message vector where the separate steps (gm_step_*) and the next run expect it
      GM_CUDA_OK(cudaMemcpyAsync(xbuf[0], xbuf[1], (size_t)gv.n_full * sizeof(T), cudaMemcpyDeviceToDevice, st));
    }
    if (async) {
      GM_CUDA_OK(cudaMemcpyAsync(gv.h_flags, gv.d_flags, sizeof(int), cudaMemcpyDeviceToHost, st));
      if (timing && harvest(pending)) return 1;
      GM_CUDA_OK(cudaStreamSynchronize(st));
      int changed = gv.h_flags[0];
      if (gv.world > 1 && !peers && gm_graph_allreduce_or(g, &changed)) return 1;
      converged = !changed;
    }
    GM_CUDA_OK(cudaEventRecord(rg.e1, st));
    GM_CUDA_OK(cudaEventSynchronize(rg.e1));
    if (peers && gv.h_flags[15]) {
      fprintf(stderr, "graphmat_b200: a peer did not reach the barrier (timeout)\n");
      return 1;
    }
    if (stats) {
      float t;
      GM_CUDA_OK(cudaEventElapsedTime(&t, rg.e0, rg.e1));
      stats->iterations = it;
      stats->converged = converged;
      stats->ms_total = t;
      stats->ms_spmv = ms_spmv;
      stats->kernel_launches = sc.launches;
      stats->edges_processed = sc.edges;
      stats->push_passes = sc.push_passes;
    }
    return 0;
  }
};

}  // namespace gm
#endif
