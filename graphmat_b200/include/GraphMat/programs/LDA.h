// LDA.h -- the topic-model programs of the reference app (narayanan2004/GraphMat src/LDA.cpp:36-275),
// annotated GM_HD.  Vertices are documents ('d') and terms ('w') of a bipartite graph whose edge value is a term
// count; LDAInitProgram spreads every count over K topics with a per-edge pseudo-random split, LDAProgram iterates
// the collapsed variational update, LDALLProgram evaluates the token log-likelihood.
//
// The vertex type is called LatentVector<K> in the reference too (as in SGD.cpp); it is LDAVector<K> here so that
// both apps can live in one library.  LDAProgram's do_every_iteration recomputes global_N (the per-topic sum over
// all term vertices) through Graph::applyReduceAllVertices in the reference (:131-134, :177-179); here it calls
// `recalc`, which the host side points at the device reduction (apps/LDA.cu, gm_builtin.cu).
#ifndef GRAPHMAT_B200_PROGRAMS_LDA_H
#define GRAPHMAT_B200_PROGRAMS_LDA_H
#include <cmath>
#include "../GraphProgram.h"

template <unsigned int K>
class LDAVector {  // LatentVector<K>, src/LDA.cpp:36-68
 public:
  double N[K];
  char type;
  double token_loglik;
  GM_HD LDAVector() { token_loglik = 0.0; }
  GM_HD bool operator!=(const LDAVector<K>& p) const {
    bool result = false;
    for (unsigned int i = 0; i < K; i++)
      if (fabs(p.N[i] - N[i]) > 1e-3) result = true;
    return result;
  }
};

namespace gm_lda {
// glibc's rand_r (stdlib/rand_r.c), which process_message of LDAInitProgram calls per edge (:93-96): restated so
// that the device draws the same numbers
GM_HD inline int rand_r_glibc(unsigned int* seed) {
  unsigned int next = *seed;
  int result;
  next *= 1103515245u;
  next += 12345u;
  result = (unsigned int)(next / 65536u) % 2048u;
  next *= 1103515245u;
  next += 12345u;
  result <<= 10;
  result ^= (unsigned int)(next / 65536u) % 1024u;
  next *= 1103515245u;
  next += 12345u;
  result <<= 10;
  result ^= (unsigned int)(next / 65536u) % 1024u;
  *seed = next;
  return result;
}
}  // namespace gm_lda

template <unsigned int K>
class LDAInitProgram : public GraphMat::GraphProgram<LDAVector<K>, LDAVector<K>, LDAVector<K> > {  // :69-111
 public:
  static const bool gm_reorderable = true;  // fp64 vector +, as SGD
  GM_HD LDAInitProgram() {
    this->order = GraphMat::ALL_EDGES;
    this->activity = GraphMat::ALL_VERTICES;
    this->process_message_requires_vertexprop = false;
  }
  GM_HD void reduce_function(LDAVector<K>& v, const LDAVector<K>& w) const {
    for (unsigned int i = 0; i < K; i++) v.N[i] += w.N[i];
  }
  GM_HD void process_message(const LDAVector<K>& message, const int edge_value, const LDAVector<K>& vertexprop,
                             LDAVector<K>& res) const {
    double gamma_wjk[K];
    double sum = 0;
    // the same random split for both directions of the edge: the seed is the edge value
    unsigned int rstart = edge_value;
    for (unsigned int i = 0; i < K; i++) {
      gamma_wjk[i] = (double)gm_lda::rand_r_glibc(&rstart) / 2147483647;  // RAND_MAX
      sum += gamma_wjk[i];
    }
    for (unsigned int i = 0; i < K; i++) res.N[i] = gamma_wjk[i] / sum * (double)edge_value;
  }
  GM_HD bool send_message(const LDAVector<K>& vertexprop, LDAVector<K>& message) const {
    message = vertexprop;
    return true;
  }
  GM_HD void apply(const LDAVector<K>& message_out, LDAVector<K>& vertexprop) {
    for (unsigned int i = 0; i < K; i++) vertexprop.N[i] = message_out.N[i];
  }
};

template <unsigned int K>
class LDAProgram : public GraphMat::GraphProgram<LDAVector<K>, LDAVector<K>, LDAVector<K> > {  // :127-196
 public:
  double alpha;
  double eta;
  double vocab_size;
  LDAVector<K> global_N;
  // host side only: recompute global_N = sum over the term vertices of N (the reference holds a Graph& and calls
  // applyReduceAllVertices(&global_N, IfTerm, Add), :131-134)
  void (*recalc)(void* ctx, LDAVector<K>* out);
  void* recalc_ctx;
  static const bool gm_reorderable = true;

  GM_HD LDAProgram(double a = 1.0, double e = 5.0, double V = 1.0) : alpha(a), eta(e), vocab_size(V) {
    recalc = nullptr;
    recalc_ctx = nullptr;
    for (unsigned int i = 0; i < K; i++) global_N.N[i] = 0;
    this->order = GraphMat::ALL_EDGES;
    this->activity = GraphMat::ALL_VERTICES;
  }
  void calcGlobalN() {
    for (unsigned int i = 0; i < K; i++) global_N.N[i] = 0;
    if (recalc) recalc(recalc_ctx, &global_N);
  }
  GM_HD void reduce_function(LDAVector<K>& v, const LDAVector<K>& w) const {
    for (unsigned int i = 0; i < K; i++) v.N[i] += w.N[i];
  }
  GM_HD void process_message(const LDAVector<K>& message, const int edge_value, const LDAVector<K>& vertexprop,
                             LDAVector<K>& res) const {
    double gamma_wjk[K];
    double my_offset, other_offset;
    if (vertexprop.type == 'd') {
      my_offset = alpha;
      other_offset = eta;
    } else {
      my_offset = eta;
      other_offset = alpha;
    }
    double sum = 0;
    for (unsigned int i = 0; i < K; i++) {
      gamma_wjk[i] = (vertexprop.N[i] + my_offset - 1.0) * (message.N[i] + other_offset - 1.0) /
                     (global_N.N[i] + vocab_size * (eta - 1.0));
      sum += gamma_wjk[i];
    }
    for (unsigned int i = 0; i < K; i++) res.N[i] = gamma_wjk[i] / sum * (double)edge_value;
  }
  GM_HD bool send_message(const LDAVector<K>& vertexprop, LDAVector<K>& message) const {
    message = vertexprop;
    return true;
  }
  GM_HD void apply(const LDAVector<K>& message_out, LDAVector<K>& vertexprop) {
    for (unsigned int i = 0; i < K; i++) vertexprop.N[i] = message_out.N[i];
  }
  void do_every_iteration(int iteration_number) { calcGlobalN(); }
};

template <unsigned int K>
class LDALLProgram : public GraphMat::GraphProgram<LDAVector<K>, double, LDAVector<K> > {  // :198-250
 public:
  LDAVector<K> N_k;
  double eta;
  int nterms;
  static const bool gm_reorderable = true;
  GM_HD LDALLProgram() : eta(5.0), nterms(0) {
    this->activity = GraphMat::ALL_VERTICES;
    this->order = GraphMat::OUT_EDGES;
  }
  GM_HD LDALLProgram(LDAVector<K> _N_k, double _eta, int _nterms) : N_k(_N_k), eta(_eta), nterms(_nterms) {
    this->activity = GraphMat::ALL_VERTICES;
    this->order = GraphMat::OUT_EDGES;
    for (unsigned int i = 0; i < K; i++) N_k.N[i] = N_k.N[i] + nterms * (eta - 1.0);  // smoothed N_k
  }
  GM_HD void reduce_function(double& v, const double& w) const { v += w; }
  GM_HD void process_message(const LDAVector<K>& message, const int edge_value, const LDAVector<K>& vertexprop,
                             double& res) const {
    double phi_wk[K];
    double theta_kj[K];
    double sum = 0;
    for (unsigned int i = 0; i < K; i++) {
      phi_wk[i] = (vertexprop.N[i] + (eta - 1.0)) / (N_k.N[i]);
      theta_kj[i] = (message.N[i] + (eta - 1.0));
      sum += theta_kj[i];
    }
    for (unsigned int i = 0; i < K; i++) theta_kj[i] /= sum;
    double dot = 0.0;
    for (unsigned int i = 0; i < K; i++) dot += phi_wk[i] * theta_kj[i];
    res = edge_value * log(dot);
  }
  GM_HD bool send_message(const LDAVector<K>& vertexprop, LDAVector<K>& message) const {
    message = vertexprop;
    return true;
  }
  GM_HD void apply(const double& message_out, LDAVector<K>& vertexprop) { vertexprop.token_loglik = message_out; }
};

// the map / reduce pair of calcGlobalN (IfTerm / Add, :113-125) and of the final log-likelihood (return_ll, :268-271)
// as GM_HD functors for the device overload of applyReduceAllVertices
template <unsigned int K>
struct LDAIfTerm {
  GM_HD void operator()(LDAVector<K>* v, LDAVector<K>* out) const {
    if (v->type == 'w') {
      for (unsigned int i = 0; i < K; i++) out->N[i] = v->N[i];
    } else {
      for (unsigned int i = 0; i < K; i++) out->N[i] = 0;
    }
  }
};
template <unsigned int K>
struct LDAAdd {
  GM_HD void operator()(const LDAVector<K>& v1, const LDAVector<K>& v2, LDAVector<K>* out) const {
    for (unsigned int i = 0; i < K; i++) out->N[i] = v1.N[i] + v2.N[i];
  }
};
template <unsigned int K>
struct LDAReturnLL {
  GM_HD void operator()(LDAVector<K>* v, double* out) const { *out = v->token_loglik; }
};
struct LDAAddDouble {
  GM_HD void operator()(const double& a, const double& b, double* c) const { *c = a + b; }
};
#endif
