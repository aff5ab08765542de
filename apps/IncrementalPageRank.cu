// IncrementalPageRank app: the driver of the reference's src/IncrementalPageRank.cpp:128-190 on the device engine.
// usage: IncrementalPageRank <binary mtx prefix> [--dump out.txt]
#include "GraphMatRuntime.h"
#include "GraphMat/programs/IncrementalPageRank.h"
#include "GraphMat/programs/PageRank.h"
#include "common.h"

void run_pagerank(const char* filename, const char* dump) {
  GraphMat::Graph<dPR> G;
  DeltaPageRank dpr;
  Degree<dPR, int> dg;
  G.ReadMTX(filename);

  auto dg_tmp = GraphMat::graph_program_init(dg, G);
  double t0 = now_ms();
  G.setAllActive();
  GraphMat::run_graph_program(&dg, G, 1, &dg_tmp);
  printf("Degree Time = %.3f ms \n", now_ms() - t0);
  GraphMat::graph_program_clear(dg_tmp);

  auto dpr_tmp = GraphMat::graph_program_init(dpr, G);
  t0 = now_ms();
  G.setAllActive();
  GraphMat::run_graph_program(&dpr, G, GraphMat::UNTIL_CONVERGENCE, &dpr_tmp);
  printf("PR Time = %.3f ms \n", now_ms() - t0);
  GraphMat::graph_program_clear(dpr_tmp);

  for (int i = 1; i <= std::min(25, G.getNumberOfVertices()); i++)
    if (G.vertexNodeOwner(i)) printf("%d : %d %f\n", i, G.getVertexproperty(i).degree, G.getVertexproperty(i).pagerank);
  if (dump) {
    FILE* f = fopen(dump, "w");
    for (int i = 1; i <= G.getNumberOfVertices(); i++)
      fprintf(f, "%d %d %.17g %.17g\n", i, G.getVertexproperty(i).degree, G.getVertexproperty(i).pagerank,
              G.getVertexproperty(i).delta);
    fclose(f);
  }
}

int main(int argc, char* argv[]) {
  if (argc < 2) {
    printf("Correct format: %s A.mtx\n", argv[0]);
    return 0;
  }
  run_pagerank(argv[1], dump_path(argc, argv));
  return 0;
}
