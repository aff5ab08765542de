/* oracle/rmat_gen.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Host-side generator of the synthetic inputs of SURVEY.md 8(d): Graph500-style RMAT
 * (a,b,c,d = .57,.19,.19,.05, edge factor 16, duplicates and self loops kept, public
 * 1-based ids) and uniform integer edge weights.  It restates the device generator of
 * the product library (graphmat_b200/csrc/gm_core.cu: rmat_edge / rmat_weight) so that
 * bench.py's reference arm and the tests can build the SAME edge list without loading
 * the product's .so; tests/test_oracle.py checks the two agree.
 */
#include <stdint.h>

static inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
#define TA 2448131358u   /* floor(0.57 * 2^32) */
#define TAB 3264175145u  /* floor(0.76 * 2^32) */
#define TABC 4080218931u /* floor(0.95 * 2^32) */

int gmo_rmat_edges(int scale, int edge_factor, unsigned long long seed, int weight_max, unsigned long long weight_seed,
                   int* src, int* dst, int* val) {
  const long long nnz = (long long)edge_factor << scale;
#pragma omp parallel for schedule(static)
  for (long long e = 0; e < nnz; e++) {
    unsigned si = 0, di = 0;
    uint64_t h = 0;
    for (int l = 0; l < scale; l++) {
      if ((l & 1) == 0) h = splitmix64(seed * 0x100000001B3ull + (uint64_t)e * 32ull + (uint64_t)(l >> 1));
      const unsigned r = (l & 1) ? (unsigned)(h >> 32) : (unsigned)h;
      const unsigned sb = r >= TAB;
      const unsigned db = (r >= TA && r < TAB) || r >= TABC;
      si = (si << 1) | sb;
      di = (di << 1) | db;
    }
    src[e] = (int)si + 1;
    dst[e] = (int)di + 1;
    if (val)
      val[e] = weight_max <= 0 ? 1
                               : 1 + (int)(splitmix64(weight_seed * 0x9E3779B1ull + (uint64_t)e + 0x5555555555ull) %
                                           (uint64_t)weight_max);
  }
  return 0;
}
