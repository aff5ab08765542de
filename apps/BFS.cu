// BFS app: the driver of the reference's src/BFS.cpp:110-171 on the device engine.
// usage: BFS <binary mtx prefix> <source vertex> [--dump out.txt]
#include "GraphMatRuntime.h"
#include "GraphMat/programs/BFS.h"
#include "common.h"

void reachable_or_not(BFSD2* v, int* result, void* params = nullptr) { *result = v->depth < gm_bfs::kMaxDist ? 1 : 0; }

void run_bfs(const char* filename, int v, const char* dump) {
  GraphMat::Graph<BFSD2> G;
  G.ReadMTX(filename);
  for (int i = 0; i < G.getNumberOfVertices(); i++) {
    BFSD2 vp = G.getVertexproperty(i + 1);
    vp.id = i + 1;
    G.setVertexproperty(i + 1, vp);
  }
  BFS2 b;
  auto b_tmp = GraphMat::graph_program_init(b, G);
  G.setAllInactive();
  auto source = G.getVertexproperty(v);
  source.depth = 0;
  G.setVertexproperty(v, source);
  G.setActive(v);

  double t0 = now_ms();
  GraphMat::run_graph_program(&b, G, GraphMat::UNTIL_CONVERGENCE, &b_tmp);
  printf("Time = %.3f ms \n", now_ms() - t0);
  GraphMat::graph_program_clear(b_tmp);

  int reachable_vertices = 0;
  G.applyReduceAllVertices(&reachable_vertices, reachable_or_not);
  printf("Reachable vertices = %d \n", reachable_vertices);
  for (int i = 1; i <= std::min(10, G.getNumberOfVertices()); i++)
    if (G.vertexNodeOwner(i)) {
      if (G.getVertexproperty(i).depth < gm_bfs::kMaxDist)
        printf("Depth %d : %u parent: %llu\n", i, G.getVertexproperty(i).depth, G.getVertexproperty(i).parent);
      else
        printf("Depth %d : INF \n", i);
    }
  if (dump) {
    FILE* f = fopen(dump, "w");
    for (int i = 1; i <= G.getNumberOfVertices(); i++)
      fprintf(f, "%d %u %llu\n", i, G.getVertexproperty(i).depth, G.getVertexproperty(i).parent);
    fclose(f);
  }
}

int main(int argc, char* argv[]) {
  if (argc < 3) {
    printf("Correct format: %s A.mtx source_vertex (1-based index)\n", argv[0]);
    return 0;
  }
  run_bfs(argv[1], atoi(argv[2]), dump_path(argc, argv));
  return 0;
}
