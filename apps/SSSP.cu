// SSSP app: the driver of the reference's src/SSSP.cpp:102-157 on the device engine.
// usage: SSSP <binary mtx prefix> <source vertex> [--dump out.txt]
#include "GraphMatRuntime.h"
#include "GraphMat/programs/SSSP.h"
#include "common.h"

void reachable_or_not(SSSP_vertex_type* v, int* result, void* params = nullptr) {
  *result = v->distance < gm_sssp::kMaxDist ? 1 : 0;
}

void run_sssp(const char* filename, int v, const char* dump) {
  GraphMat::Graph<SSSP_vertex_type, int> G;
  G.ReadMTX(filename);
  SSSP<int> b;
  auto tmp = GraphMat::graph_program_init(b, G);
  SSSP_vertex_type init;
  init.distance = 0;
  SSSP_vertex_type inf;
  G.setAllVertexproperty(inf);
  G.setAllInactive();
  G.setVertexproperty(v, init);
  G.setActive(v);

  double t0 = now_ms();
  GraphMat::run_graph_program(&b, G, GraphMat::UNTIL_CONVERGENCE, &tmp);
  printf("Time = %.3f ms \n", now_ms() - t0);

  int reachable_vertices = 0;
  G.applyReduceAllVertices(&reachable_vertices, reachable_or_not);
  printf("Reachable vertices = %d \n", reachable_vertices);
  GraphMat::graph_program_clear(tmp);
  for (int i = 1; i <= std::min(10, G.getNumberOfVertices()); i++)
    if (G.vertexNodeOwner(i)) {
      if (G.getVertexproperty(i).distance < gm_sssp::kMaxDist) printf("%d : distance = %u\n", i, G.getVertexproperty(i).distance);
      else printf("%d : distance = INF\n", i);
    }
  if (dump) {
    FILE* f = fopen(dump, "w");
    for (int i = 1; i <= G.getNumberOfVertices(); i++) fprintf(f, "%d %u\n", i, G.getVertexproperty(i).distance);
    fclose(f);
  }
}

int main(int argc, char* argv[]) {
  if (argc < 3) {
    printf("Correct format: %s A.mtx source_vertex (1-based index)\n", argv[0]);
    return 0;
  }
  run_sssp(argv[1], atoi(argv[2]), dump_path(argc, argv));
  return 0;
}
