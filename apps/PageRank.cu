// PageRank app: the driver of the reference's src/PageRank.cpp:115-176 on the device engine.
// usage: PageRank <binary mtx prefix> [--dump out.txt]     (reads <prefix>0)
#include "GraphMatRuntime.h"
#include "GraphMat/programs/PageRank.h"
#include "common.h"

void run_pagerank(const char* filename, const char* dump) {
  GraphMat::Graph<PR, int> G;
  PageRank<int> pr;
  Degree<PR, int> dg;

  G.ReadMTX(filename);

  auto dg_tmp = GraphMat::graph_program_init(dg, G);
  double t0 = now_ms();
  G.setAllActive();
  GraphMat::run_graph_program(&dg, G, 1, &dg_tmp);
  printf("Degree Time = %.3f ms \n", now_ms() - t0);
  GraphMat::graph_program_clear(dg_tmp);

  auto pr_tmp = GraphMat::graph_program_init(pr, G);
  t0 = now_ms();
  G.setAllActive();
  GraphMat::run_graph_program(&pr, G, GraphMat::UNTIL_CONVERGENCE, &pr_tmp);
  printf("PR Time = %.3f ms \n", now_ms() - t0);
  GraphMat::graph_program_clear(pr_tmp);

  for (int i = 1; i <= std::min(25, G.getNumberOfVertices()); i++)
    if (G.vertexNodeOwner(i)) printf("%d : %d %f\n", i, G.getVertexproperty(i).degree, G.getVertexproperty(i).pagerank);
  if (dump) {
    FILE* f = fopen(dump, "w");
    for (int i = 1; i <= G.getNumberOfVertices(); i++)
      fprintf(f, "%d %d %.9g\n", i, G.getVertexproperty(i).degree, G.getVertexproperty(i).pagerank);
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
