// TopologicalSort app: the driver of the reference's src/TopologicalSort.cpp:141-205 on the device engine.
// usage: TopologicalSort <binary mtx prefix> [--dump out.txt]
#include "GraphMatRuntime.h"
#include "GraphMat/programs/TopologicalSort.h"
#include "common.h"

typedef TopSortVertex Vertex_type;

void unreachable(Vertex_type* v, int* result, void* params = nullptr) {
  *result = v->topsort_order == gm_topsort::kMaxDist ? 1 : 0;
}

void run_topsort(const char* filename, const char* dump) {
  GraphMat::Graph<Vertex_type> G;
  G.ReadMTX(filename);

  InDegree<Vertex_type> indeg;
  TopSort topsort;
  auto d_tmp = GraphMat::graph_program_init(indeg, G);
  auto b_tmp = GraphMat::graph_program_init(topsort, G);

  double t0 = now_ms();
  GraphMat::run_graph_program(&indeg, G, 1, &d_tmp);

  G.setAllInactive();
  for (int i = 1; i <= G.getNumberOfVertices(); i++) {
    if (G.vertexNodeOwner(i)) {
      auto v = G.getVertexproperty(i);
      if (v.in_degree == 0) {
        G.setActive(i);
        v.topsort_order = 0;
        G.setVertexproperty(i, v);
      }
    }
  }
  GraphMat::run_graph_program(&topsort, G, GraphMat::UNTIL_CONVERGENCE, &b_tmp);
  printf("Time = %.3f ms \n", now_ms() - t0);
  GraphMat::graph_program_clear(d_tmp);
  GraphMat::graph_program_clear(b_tmp);

  int unreachable_vertices = 0;
  G.applyReduceAllVertices(&unreachable_vertices, unreachable);  // default reduction = sum
  if (dump) {
    FILE* f = fopen(dump, "w");
    for (int i = 1; i <= G.getNumberOfVertices(); i++)
      fprintf(f, "%d %u %d\n", i, G.getVertexproperty(i).topsort_order, G.getVertexproperty(i).in_degree);
    fclose(f);
  }
  if (unreachable_vertices > 0) {
    printf("Topological Sort not possible. Graph has cycles.\n");
    return;
  }
  for (int i = 1; i <= std::min(10, G.getNumberOfVertices()); i++)
    if (G.vertexNodeOwner(i)) printf("Top Sort order %d : %d\n", i, G.getVertexproperty(i).topsort_order);
}

int main(int argc, char* argv[]) {
  if (argc < 2) {
    printf("Correct format: %s A.mtx\n", argv[0]);
    return 0;
  }
  run_topsort(argv[1], dump_path(argc, argv));
  return 0;
}
