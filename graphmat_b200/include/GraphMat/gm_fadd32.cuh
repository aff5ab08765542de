// gm_fadd32.cuh -- bit-exact PARALLEL evaluation of a serial fp32 left fold
//     s = a[0]; for k = 1..n-1: s = fl(s + a[k])           (a[k] >= 0)
// which is what the reference computes for PageRank's reduce_function (a += b,
// src/PageRank.cpp:93-95) through my_spmspv's per-row loop
// (include/GMDP/singlenode/spmspv.h:61-77 of narayanan2004/GraphMat).  A tree sum is
// NOT within the 1e-6 parity bound on high-degree rows (measured: 1e-5 relative at
// RMAT-20), and a serial chain of ~10^6 dependent FADDs is ~2 ms, so long rows need
// this.
//
// Idea: while the accumulator s stays inside one binade [2^e, 2^(e+1)), every
// representable value is an integer multiple m*u of u = ulp(s) = 2^(e-23), and
//     fl(m*u + a) = (m + q)*u,   q = RN_even(a/u) -- with the tie broken by the parity of m.
// So one addend acts on m as a map  m -> m + (m even ? q0 : q1), where
//     q0 = (fl(B0 + a) - B0)/u with B0 = 2^e      (an even m),
//     q1 = (fl(B1 + a) - B1)/u with B1 = 2^e + u  (an odd m),
// both computed with two ordinary fp32 adds.  Such maps compose associatively (the exit
// parity of one decides which branch of the next is taken), so a block of addends is
// folded with an ORDER-PRESERVING parallel scan of (q0, q1) pairs.  The block is exact
// as long as m never leaves [2^23, 2^24): since all q >= 0 it suffices to check the
// exit value.  Where the accumulator crosses into the next binade (<= ~30 times per
// row), or an addend is negative / NaN / larger than the accumulator, the affected
// 8-element sub-block is folded with plain serial FADDs and the scan restarts behind it.
#ifndef GRAPHMAT_B200_FADD32_CUH
#define GRAPHMAT_B200_FADD32_CUH
#include <cstring>

#include "gm_hd.h"

namespace gm {
namespace fx {

static const unsigned kSat = 1u << 26;  // anything >= 2^24 means "left the binade"; keep sums inside 32 bits
static const int kPerLane = 8;          // consecutive addends folded by one lane

GM_HD inline unsigned f2u(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
GM_HD inline float u2f(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
GM_HD inline unsigned umin_(unsigned a, unsigned b) { return a < b ? a : b; }

// the action of a run of addends on m, as (delta when m enters even, delta when m enters odd)
struct qmap {
  unsigned d0, d1;
};
GM_HD inline qmap identity() {
  qmap r;
  r.d0 = 0;
  r.d1 = 0;
  return r;
}
// f first, then g
GM_HD inline qmap compose(qmap f, qmap g) {
  qmap h;
  h.d0 = umin_(f.d0 + ((f.d0 & 1u) ? g.d1 : g.d0), kSat);
  h.d1 = umin_(f.d1 + ((f.d1 & 1u) ? g.d0 : g.d1), kSat);
  return h;
}
GM_HD inline unsigned apply(qmap f, unsigned m) { return m + ((m & 1u) ? f.d1 : f.d0); }

struct binade {
  float B0, B1, u, inv_u;
  unsigned m;
};
// false: s is not a positive normal number with room for the scaling trick -> serial path
GM_HD inline bool binade_of(float s, binade& b) {
  unsigned bits = f2u(s);
  unsigned e = (bits >> 23) & 0xffu;
  if ((bits >> 31) || e < 25u || e == 255u) return false;
  b.B0 = u2f(e << 23);
  b.u = u2f((e - 23u) << 23);
  b.B1 = b.B0 + b.u;
  b.inv_u = u2f((277u - e) << 23);
  b.m = (bits & 0x7fffffu) | 0x800000u;
  return true;
}
// (q0, q1) of one addend; bad = the addend cannot be handled inside this binade
GM_HD inline qmap quantize(float a, const binade& b, bool& bad) {
  qmap q;
  if (!(a >= 0.0f && a < b.B0)) {  // negative, NaN, or certain to leave the binade
    bad = true;
    return identity();
  }
#if defined(__CUDA_ARCH__)
  float t0 = __fadd_rn(b.B0, a), t1 = __fadd_rn(b.B1, a);
  q.d0 = (unsigned)__float2int_rn(__fmul_rn(__fsub_rn(t0, b.B0), b.inv_u));
  q.d1 = (unsigned)__float2int_rn(__fmul_rn(__fsub_rn(t1, b.B1), b.inv_u));
#else
  volatile float t0 = b.B0 + a, t1 = b.B1 + a;
  volatile float r0 = t0 - b.B0, r1 = t1 - b.B1;
  q.d0 = (unsigned)(int)(r0 * b.inv_u);
  q.d1 = (unsigned)(int)(r1 * b.inv_u);
#endif
  return q;
}

#if defined(__CUDACC__)
// q0 alone, and whether the addend is a TIE.  a/u is exact (a power-of-two scaling), q0 = RN_even(a/u) is what
// fl(B0 + a) computes (B0's integer 2^23 is even), and q1 differs from q0 only when a/u lies exactly half way
// between two integers (then the odd neighbour is taken).  Without ties a run of addends is the plain translation
// m -> m + sum(q0): no (q0, q1) pairs, no parity selects.
__device__ __forceinline__ unsigned quantize_q0(float a, const binade& b, bool& bad, bool& tie) {
  if (!(a >= 0.0f && a < b.B0)) {  // negative, NaN, or certain to leave the binade
    bad = true;
    return 0u;
  }
  const float t = __fmul_rn(a, b.inv_u);
  const float r = rintf(t);  // round half to even
  tie |= fabsf(__fsub_rn(t, r)) == 0.5f;
  return (unsigned)__float2int_rn(r);
}
#endif

// plain serial fold of up to 8 addends held by one lane; bit k of vmask = addend k exists
GM_HD inline void serial8(const float (&v)[kPerLane], unsigned vmask, float& s, bool& have) {
#pragma unroll
  for (int k = 0; k < kPerLane; k++) {
    if ((vmask >> k) & 1u) {
#if defined(__CUDA_ARCH__)
      s = have ? __fadd_rn(s, v[k]) : v[k];
#else
      volatile float t = s + v[k];
      s = have ? (float)t : v[k];
#endif
      have = true;
    }
  }
}

#if defined(__CUDACC__)
// inclusive, order-preserving warp scan of maps (lane i ends with the composition of lanes 0..i)
__device__ __forceinline__ qmap warp_scan(qmap own, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    qmap prev;
    prev.d0 = __shfl_up_sync(0xffffffffu, own.d0, o);
    prev.d1 = __shfl_up_sync(0xffffffffu, own.d1, o);
    if (lane >= o) own = compose(prev, own);
  }
  return own;
}

// One warp folds its 256 addends (lane L holds addends 8L..8L+7 of the block, in fold
// order; missing ones have their vmask bit clear and value 0) into the warp-uniform
// running state (s, have).  Exact for ANY input: everything the scan cannot prove is
// folded serially.
__device__ __forceinline__ void warp_fold(const float (&v)[kPerLane], unsigned vmask, float& s, bool& have, int lane) {
  unsigned pending = __ballot_sync(0xffffffffu, vmask != 0);  // lanes whose sub-block is not applied yet
  while (pending) {
    binade b;
    const bool hot = have && binade_of(s, b);  // warp-uniform
    if (!hot) {
      // cold: no accumulator yet, or it is tiny/zero/negative/non-finite: fold one sub-block serially
      const int f = __ffs(pending) - 1;
      float sf = s;
      bool hf = have;
      if (lane == f) serial8(v, vmask, sf, hf);
      s = __shfl_sync(0xffffffffu, sf, f);
      have = __shfl_sync(0xffffffffu, (int)hf, f) != 0;
      pending &= ~(1u << f);
      continue;
    }
    bool bad = false;
    qmap mine = identity();
    if ((pending >> lane) & 1u) {
#pragma unroll
      for (int k = 0; k < kPerLane; k++) {
        qmap q = quantize(v[k], b, bad);  // missing addends are +0: the identity
        mine = compose(mine, q);
      }
    }
    const qmap incl = warp_scan(mine, lane);
    const unsigned m_after = apply(incl, b.m);
    const bool over = bad || m_after >= (1u << 24);
    const unsigned fail = __ballot_sync(0xffffffffu, over) & pending;
    if (fail == 0) {
      const unsigned m_end = __shfl_sync(0xffffffffu, m_after, 31);
      s = __fmul_rn(__uint2float_rn(m_end), b.u);
      pending = 0;
    } else {
      // lanes before f are proven exact; lane f folds its own 8 serially from its exact entry value
      const int f = __ffs(fail) - 1;
      unsigned m_prev = __shfl_up_sync(0xffffffffu, m_after, 1);
      if (lane == 0) m_prev = b.m;
      float sf = __fmul_rn(__uint2float_rn(m_prev), b.u);
      bool hf = true;
      if (lane == f) serial8(v, vmask, sf, hf);
      s = __shfl_sync(0xffffffffu, sf, f);
      pending &= ~((2u << f) - 1u);
    }
  }
}
#endif

}  // namespace fx
}  // namespace gm
#endif
