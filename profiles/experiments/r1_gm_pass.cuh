// gm_pass.cuh -- one persistent kernel for a whole SpMSpV pass of an fp32-sum program
// (PageRank), with the hottest columns of the message vector held in shared memory.
//
// Why: the pass is limited by how many random sectors an SM can look up per cycle (L1/L2 path:
// ~1 per cycle per SM = 296 G gathers/s chip-wide, hit or miss), not by HBM
// (profiles/r1_micro_gather.txt).  Shared memory serves ~7.6 random 4-byte reads per cycle per
// SM, and x is stored hottest-columns-first, so a copy of x[0 .. hot_n) per SM takes the most
// frequent gathers off the L1/L2 path.  A shared-memory cache only pays in a kernel that stays
// resident, so the three row classes of gm_engine.cuh (block-cooperative exact fold for the
// longest rows, warp-per-row exact fold, one-row-per-lane sliced ELL) become work items of one
// persistent grid (one 1024-thread block per SM), handed out longest first through two atomic
// counters -- the same longest-processing-time-first balance the block scheduler gave the separate
// launches.  The arithmetic is the same code path as k_heavy_fadd32 / k_sell: results are
// bit-identical (GM_PASS_HOT=65536 GM_PASS_MIN_SLICES=0 pytest -m gpu passes).
//
// STATUS (round 1): correct but NOT yet faster -- 4.7-4.9 ms per RMAT-26 pass against 3.5 ms for the
// separate kernels -- so it is off by default (GM_PASS_HOT=0).  With one 1024-thread block per SM at
// 64 registers the warps are latency bound (each slice is a chain of dependent loads and only 32
// warps per SM hide it); ncu shows l1tex 72 %, lts 56 %, issue 29 %.  Next step: more slices in
// flight per warp (or 2x512-thread blocks with half the cache each) before the cache can pay.
#ifndef GRAPHMAT_B200_PASS_CUH
#define GRAPHMAT_B200_PASS_CUH

namespace gm {

struct pass_group_state {
  float s;
  int have, fail, row;
  unsigned d0[16], d1[16], bad[16];
};

__device__ __forceinline__ void group_sync(int group) {  // 16 warps = 512 threads; id 0 is __syncthreads
  asm volatile("bar.sync %0, 512;" ::"r"(group + 1) : "memory");
}

template <class X>
__device__ __forceinline__ X hot_gather(const X* __restrict__ x, const X* hx, int c, int hot_n) {
  return c < hot_n ? hx[c] : __ldg(x + c);
}

// ---- sliced-ELL slice, one row per lane (same fold as k_sell) ----
template <class P, class T, class V, class E, bool ALLACT, bool IDENT, int UNROLL>
__device__ __forceinline__ void pass_sell_slice(const P& prog, const gm_matrix_view& M, int s, const T* hx, int hot_n,
                                                const T* __restrict__ x, const unsigned* __restrict__ xbits,
                                                float* __restrict__ y, unsigned* __restrict__ ybits, int lane) {
  const int* __restrict__ cols = M.s_col;
  const E* __restrict__ vals = reinterpret_cast<const E*>(M.s_val);
  const int slot = M.n_heavy + s * 32 + lane;
  const int len = __ldg(M.row_len + slot);
  const long long base = __ldg(M.slice_ptr + s);
  const int width = __shfl_sync(0xffffffffu, len, 0);
  const int vtx = IDENT ? slot : __ldg(M.slot_vertex + slot);
  V vdummy;
  float acc = 0.f;
  bool have = false;
  const int* cp = cols + base + lane;
  const E* ep = vals + base + lane;
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
        ev[u] = ld_stream(ep + (long long)(i + u) * 32);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (on[u]) {
        if (!ALLACT) on[u] = test_bit(xbits, c[u]);
        if (on[u]) xv[u] = hot_gather(x, hx, c[u], hot_n);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (on[u]) {
        if (have) {
          float tmp;
          prog.P::process_message(xv[u], ev[u], vdummy, tmp);
          prog.P::reduce_function(acc, tmp);
        } else {
          prog.P::process_message(xv[u], ev[u], vdummy, acc);
          have = true;
        }
      }
    }
  }
  if (have && len > 0) y[vtx] = acc;
  const unsigned m = __ballot_sync(0xffffffffu, have);
  if (IDENT) {
    if (lane == 0) ybits[slot >> 5] = m;
  } else if (have && len > 0) {
    atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
  }
}

// ---- 8 consecutive addends of a heavy row per lane ----
template <class P, class T, class V, class E, bool ALLACT>
__device__ __forceinline__ void pass_load8(const P& prog, const int* __restrict__ cols, const E* __restrict__ vals,
                                           long long i0, long long beg, long long end, const T* hx, int hot_n,
                                           const T* __restrict__ x, const unsigned* __restrict__ xbits, float (&v)[8],
                                           unsigned& vmask) {
  V vdummy;
  vmask = 0;
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
      if (on) {
        xv[j] = hot_gather(x, hx, c[j], hot_n);
        vmask |= 1u << j;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
      if ((vmask >> j) & 1u) prog.P::process_message(xv[j], ev[j], vdummy, v[j]);
  }
}

// ---- heavy row folded by one warp (same fold as k_heavy_fadd32<1>) ----
template <class P, class T, class V, class E, bool ALLACT, bool IDENT>
__device__ __forceinline__ void pass_row_warp(const P& prog, const gm_matrix_view& M, int slot, const T* hx, int hot_n,
                                              const T* __restrict__ x, const unsigned* __restrict__ xbits,
                                              float* __restrict__ y, unsigned* __restrict__ ybits, int lane) {
  const long long beg = __ldg(M.h_ptr + slot), end = __ldg(M.h_ptr + slot + 1);
  const int vtx = IDENT ? slot : __ldg(M.slot_vertex + slot);
  const E* __restrict__ vals = reinterpret_cast<const E*>(M.h_val);
  float s = 0.f;
  bool have = false;
  for (long long k0 = beg & ~7ll; k0 < end; k0 += 256) {
    float v[8];
    unsigned vmask;
    pass_load8<P, T, V, E, ALLACT>(prog, M.h_col, vals, k0 + lane * 8, beg, end, hx, hot_n, x, xbits, v, vmask);
    fx::warp_fold(v, vmask, s, have, lane);
  }
  if (lane == 0 && have) {
    y[vtx] = s;
    atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
  }
}

// ---- heavy row folded by a group of 16 warps (same fold as k_heavy_fadd32<16>) ----
template <class P, class T, class V, class E, bool ALLACT, bool IDENT>
__device__ __forceinline__ void pass_row_group(const P& prog, const gm_matrix_view& M, int slot, const T* hx, int hot_n,
                                               const T* __restrict__ x, const unsigned* __restrict__ xbits,
                                               float* __restrict__ y, unsigned* __restrict__ ybits,
                                               pass_group_state& gs, int group, int w, int lane) {
  constexpr int W = 16;
  const long long beg = __ldg(M.h_ptr + slot), end = __ldg(M.h_ptr + slot + 1);
  const int vtx = IDENT ? slot : __ldg(M.slot_vertex + slot);
  const E* __restrict__ vals = reinterpret_cast<const E*>(M.h_val);
  if (w == 0 && lane == 0) {
    gs.s = 0.f;
    gs.have = 0;
  }
  group_sync(group);
  for (long long k0 = beg & ~7ll; k0 < end; k0 += W * 256) {
    float v[8];
    unsigned vmask;
    pass_load8<P, T, V, E, ALLACT>(prog, M.h_col, vals, k0 + w * 256 + lane * 8, beg, end, hx, hot_n, x, xbits, v, vmask);
    int nw = (int)((end - k0 + 255) / 256);
    if (nw > W) nw = W;
    int first = 0;
    while (first < nw) {
      const float s = gs.s;
      const bool have = gs.have != 0;
      fx::binade b;
      const bool hot = have && fx::binade_of(s, b);  // group-uniform
      if (!hot) {
        if (w == first) {
          float sq = s;
          bool hq = have;
          fx::warp_fold(v, vmask, sq, hq, lane);
          if (lane == 0) {
            gs.s = sq;
            gs.have = hq ? 1 : 0;
          }
        }
        first++;
        group_sync(group);
        continue;
      }
      if (w >= first && w < nw) {
        bool bad = false;
        fx::qmap mine = fx::identity();
#pragma unroll
        for (int j = 0; j < 8; j++) mine = fx::compose(mine, fx::quantize(v[j], b, bad));
        const fx::qmap incl = fx::warp_scan(mine, lane);
        const unsigned anybad = __ballot_sync(0xffffffffu, bad);
        if (lane == 31) {
          gs.d0[w] = incl.d0;
          gs.d1[w] = incl.d1;
          gs.bad[w] = anybad;
        }
      }
      group_sync(group);
      if (w == 0) {
        fx::qmap t = fx::identity();
        bool bd = false;
        if (lane >= first && lane < nw) {
          t.d0 = gs.d0[lane];
          t.d1 = gs.d1[lane];
          bd = gs.bad[lane] != 0;
        }
        t = fx::warp_scan(t, lane);
        const unsigned m_after = fx::apply(t, b.m);
        const bool over = (lane >= first && lane < nw) && (bd || m_after >= (1u << 24));
        const unsigned fail = __ballot_sync(0xffffffffu, over);
        unsigned m_prev = __shfl_up_sync(0xffffffffu, m_after, 1);
        if (lane == 0) m_prev = b.m;
        if (fail == 0) {
          if (lane == nw - 1) {
            gs.s = __fmul_rn(__uint2float_rn(m_after), b.u);
            gs.fail = -1;
          }
        } else {
          const int f = __ffs(fail) - 1;
          if (lane == f) {
            gs.s = __fmul_rn(__uint2float_rn(m_prev), b.u);
            gs.fail = f;
          }
        }
      }
      group_sync(group);
      const int f = gs.fail;
      if (f < 0) break;
      if (w == f) {
        float sq = gs.s;
        bool hq = true;
        fx::warp_fold(v, vmask, sq, hq, lane);
        if (lane == 0) gs.s = sq;
      }
      first = f + 1;
      group_sync(group);
    }
  }
  group_sync(group);
  if (w == 0 && lane == 0 && gs.have) {
    y[vtx] = gs.s;
    atomicOr(ybits + (vtx >> 5), 1u << (vtx & 31));
  }
  group_sync(group);
}

// counters[0]: next block-cooperative row, counters[1]: next warp item
template <class P, class T, class V, class E, bool ALLACT, bool IDENT>
__global__ void __launch_bounds__(1024, 1)
    k_pass_fadd32(prog_bytes<P> pb, gm_matrix_view M, int hot_n, int narrow_spw, int* __restrict__ counters,
                  const T* __restrict__ x, const unsigned* __restrict__ xbits, float* __restrict__ y,
                  unsigned* __restrict__ ybits) {
  const P& prog = pb.get();
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ pass_group_state gstate[2];
  T* hx = reinterpret_cast<T*>(dsm);
  {
    const int n16 = (int)(((size_t)hot_n * sizeof(T)) >> 4);
    const int4* src = reinterpret_cast<const int4*>(x);
    int4* dst = reinterpret_cast<int4*>(dsm);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int group = warp >> 4, w = warp & 15;
  pass_group_state& gs = gstate[group];

  // phase A: the longest rows, one per 16-warp group, longest first.  They are latency bound
  // (serial dependence between 4096-entry rounds), so only group 0 of every block works on them
  // while group 1 already streams warp items: each SM overlaps the two kinds of work.
  while (group == 0) {
    if (w == 0 && lane == 0) gs.row = atomicAdd(&counters[0], 1);
    group_sync(group);
    const int row = gs.row;
    group_sync(group);
    if (row >= M.n_coop) break;
    pass_row_group<P, T, V, E, ALLACT, IDENT>(prog, M, row, hx, hot_n, x, xbits, y, ybits, gs, group, w, lane);
  }
  // phase B: warp items, longest first: remaining heavy rows, wide slices, chunks of narrow slices
  const int n_w1 = M.n_heavy - M.n_coop;
  const int n_wide = M.n_slices_wide < M.n_slices ? M.n_slices_wide : M.n_slices;
  constexpr int WIDE_PER_ITEM = 1;  // (4 per item measured slower: the tail dominates, not the counter)
  const int n_wide_items = (n_wide + WIDE_PER_ITEM - 1) / WIDE_PER_ITEM;
  const int n_narrow = (M.n_slices - n_wide + narrow_spw - 1) / narrow_spw;
  const int n_items = n_w1 + n_wide_items + n_narrow;
  while (true) {
    int it = 0;
    if (lane == 0) it = atomicAdd(&counters[1], 1);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= n_items) break;
    if (it < n_w1) {
      pass_row_warp<P, T, V, E, ALLACT, IDENT>(prog, M, M.n_coop + it, hx, hot_n, x, xbits, y, ybits, lane);
    } else if (it < n_w1 + n_wide_items) {
      int s = (it - n_w1) * WIDE_PER_ITEM;
      const int s_end = min(s + WIDE_PER_ITEM, n_wide);
      for (; s < s_end; s++)
        pass_sell_slice<P, T, V, E, ALLACT, IDENT, 16>(prog, M, s, hx, hot_n, x, xbits, y, ybits, lane);
    } else {
      int s = n_wide + (it - n_w1 - n_wide_items) * narrow_spw;
      const int s_end = min(s + narrow_spw, M.n_slices);
      for (; s < s_end; s++)
        pass_sell_slice<P, T, V, E, ALLACT, IDENT, 8>(prog, M, s, hx, hot_n, x, xbits, y, ybits, lane);
    }
  }
}

}  // namespace gm
#endif
