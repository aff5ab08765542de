// LDA app: the driver of the reference's src/LDA.cpp:274-395 on the device engine (K = 20 topics).
// usage: LDA <binary mtx prefix> <#DOC> <#TERMS> [#iterations (default 10)] [--dump out.txt]
#include <iostream>

#include "GraphMatRuntime.h"
#include "GraphMat/programs/LDA.h"
#include "common.h"

template <unsigned int K>
using LatentVector = LDAVector<K>;  // the reference's name for the vertex type

// LDAProgram::calcGlobalN of the reference calls graph_ref.applyReduceAllVertices(&global_N, IfTerm, Add) (:131-134);
// here the same reduction runs on the device through the functor overload
template <unsigned int K>
static void recalc_global_N(void* ctx, LatentVector<K>* out) {
  auto* G = static_cast<GraphMat::Graph<LatentVector<K> >*>(ctx);
  G->applyReduceAllVertices(out, LDAIfTerm<K>(), LDAAdd<K>());
}

void run_lda(const char* filename, int ndoc, int nterms, int niterations, const char* dump) {
  const int k = 20;
  GraphMat::Graph<LatentVector<k> > G;
  G.ReadMTX(filename);
  if (ndoc + nterms != G.getNumberOfVertices()) {
    std::cout << "Number of vertices in graph != NDOC + NTERMS" << std::endl;
    exit(1);
  }
  for (int i = 1; i <= G.getNumberOfVertices(); i++) {
    LatentVector<k> v;
    for (int j = 0; j < k; j++) v.N[j] = 0;  // uninitialised in the reference (:44-46); overwritten by LDAInitProgram
    v.type = (i <= ndoc) ? 'd' : 'w';
    G.setVertexproperty(i, v);
  }
  LDAInitProgram<k> ldainit_program;
  G.setAllActive();
  GraphMat::run_graph_program(&ldainit_program, G, 1);

  double alpha = 1.0;
  double eta = 5.0;
  LDAProgram<k> ldap(alpha, eta, nterms);
  ldap.recalc = recalc_global_N<k>;
  ldap.recalc_ctx = &G;
  ldap.calcGlobalN();
  auto ldap_tmp = GraphMat::graph_program_init(ldap, G);
  printf("LDA Init over\n");

  double t0 = now_ms();
  G.setAllActive();
  GraphMat::run_graph_program(&ldap, G, niterations, &ldap_tmp);
  printf("Time = %.3f ms \n", now_ms() - t0);
  GraphMat::graph_program_clear(ldap_tmp);

  auto Nk = ldap.global_N;
  LDALLProgram<k> ldall(Nk, eta, nterms);
  G.setAllActive();
  GraphMat::run_graph_program(&ldall, G, 1);
  double total_ll = 0.0;
  G.applyReduceAllVertices(&total_ll, LDAReturnLL<k>(), LDAAddDouble());
  printf("Total Loglikelihood = %lf \n", total_ll);
  if (dump) {
    FILE* f = fopen(dump, "w");
    for (int i = 1; i <= G.getNumberOfVertices(); i++) {
      fprintf(f, "%d", i);
      for (int j = 0; j < k; j++) fprintf(f, " %.17g", G.getVertexproperty(i).N[j]);
      fprintf(f, "\n");
    }
    fclose(f);
  }
}

int main(int argc, char* argv[]) {
  if (argc < 4) {
    printf("Correct format: %s A.mtx #DOC #TERMS {#iterations (default 10)}\n", argv[0]);
    return 0;
  }
  int niterations = (argc >= 5 && argv[4][0] != '-') ? atoi(argv[4]) : 10;
  run_lda(argv[1], atoi(argv[2]), atoi(argv[3]), niterations, dump_path(argc, argv));
  return 0;
}
