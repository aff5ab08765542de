// SGD app: the driver of the reference's src/SGD.cpp:163-236 on the device engine
// (collaborative filtering, K latent factors; the reference hard-codes K = 20 at :164).
// usage: SGD <binary mtx prefix> [--k 20|32] [--dump out.txt]
#include <cmath>

#include "GraphMatRuntime.h"
#include "GraphMat/programs/SGD.h"
#include "common.h"

template <class V>
void return_sqerr(V* vertexprop, double* out, void* params) { *out = vertexprop->sqerr; }

template <unsigned K>
void run_sgd(const char* filename, const char* dump) {
  GraphMat::Graph<LatentVector<K> > G;
  G.ReadMTX(filename);
  SGDProgram<K> sgdp(0.001, 0.00000035);
  RMSEProgram<K> rmsep;
  auto sgdp_tmp = GraphMat::graph_program_init(sgdp, G);
  auto rmsep_tmp = GraphMat::graph_program_init(rmsep, G);

  for (int i = 1; i <= G.getNumberOfVertices(); i++) {
    LatentVector<K> v;
    v.sqerr = 0.0;
    unsigned int r = i;
    for (unsigned j = 0; j < K; j++) v.lv[j] = ((double)rand_r(&r) / (double)RAND_MAX);
    G.setVertexproperty(i, v);
  }
  double err = 0.0;
  G.setAllActive();
  GraphMat::run_graph_program(&rmsep, G, 1, &rmsep_tmp);
  G.applyReduceAllVertices(&err, return_sqerr<LatentVector<K> >, GraphMat::AddFn<double>);
  printf("RMSE error = %lf per edge \n", sqrt(err / (double)G.nnz));

  double t0 = now_ms();
  G.setAllActive();
  GraphMat::run_graph_program(&sgdp, G, 10, &sgdp_tmp);
  printf("SGD Time = %.3f ms \n", now_ms() - t0);

  G.setAllActive();
  GraphMat::run_graph_program(&rmsep, G, 1, &rmsep_tmp);
  err = 0.0;
  G.applyReduceAllVertices(&err, return_sqerr<LatentVector<K> >, GraphMat::AddFn<double>);
  printf("RMSE error = %lf per edge \n", sqrt(err / (double)G.nnz));
  GraphMat::graph_program_clear(sgdp_tmp);
  GraphMat::graph_program_clear(rmsep_tmp);
  if (dump) {
    FILE* f = fopen(dump, "w");
    for (int i = 1; i <= G.getNumberOfVertices(); i++) {
      LatentVector<K> v = G.getVertexproperty(i);
      fprintf(f, "%d", i);
      for (unsigned j = 0; j < K; j++) fprintf(f, " %.17g", v.lv[j]);
      fprintf(f, "\n");
    }
    fclose(f);
  }
}

int main(int argc, char* argv[]) {
  if (argc < 2) {
    printf("Correct format: %s A.mtx [--k 20|32]\n", argv[0]);
    return 0;
  }
  int k = 20;
  for (int i = 2; i + 1 < argc; i++)
    if (!strcmp(argv[i], "--k")) k = atoi(argv[i + 1]);
  if (k == 32) run_sgd<32>(argv[1], dump_path(argc, argv));
  else run_sgd<20>(argv[1], dump_path(argc, argv));
  return 0;
}
