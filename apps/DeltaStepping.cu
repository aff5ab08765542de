// DeltaStepping app: the driver of the reference's src/DeltaStepping.cpp:124-214 on the device
// engine: light edges (val <= delta) to convergence, heavy edges once, per bucket.
// usage: DeltaStepping <binary mtx prefix> <delta> <source vertex> [--dump out.txt]
#include <limits>

#include "GraphMatRuntime.h"
#include "GraphMat/programs/SSSP.h"
#include "common.h"

void reachable_or_not(DeltaSteppingDS* v, int* result, void* params = nullptr) {
  *result = v->distance < gm_sssp::kMaxDist ? 1 : 0;
}
void CheckBucketNotEmpty(DeltaSteppingDS* v, int* result, void* param) {
  *result = (v->bucket >= *(int*)param && v->bucket < std::numeric_limits<int>::max()) ? 1 : 0;
}
template <typename T>
void Add(const T& a, const T& b, T* c, void* param) { *c = a + b; }
// The same map / reduce as GM_HD functors (`param` becomes a member): applyReduceAllVertices then runs on the
// device, one kernel and four bytes back per bucket instead of a pull of the whole vertex array (GM_HOST_REDUCE=1
// keeps the reference's function-pointer call, for comparison).
struct CheckBucketNotEmptyFn {
  int bid;
  GM_HD void operator()(DeltaSteppingDS* v, int* result) const { *result = (v->bucket >= bid && v->bucket < 0x7fffffff) ? 1 : 0; }
};
struct AddIntFn {
  GM_HD void operator()(const int& a, const int& b, int* c) const { *c = a + b; }
};
bool less_than_delta(GraphMat::edge_t<int> e, void* param) { return e.val <= *(int*)param; }
bool greater_than_delta(GraphMat::edge_t<int> e, void* param) { return e.val > *(int*)param; }

void run_deltastepping(const char* filename, int delta, int v, const char* dump) {
  GraphMat::edgelist_t<int> E;
  GraphMat::load_edgelist(filename, &E, true, true, true);
  if (E.m != E.n) E.m = E.n = std::max(E.m, E.n);
  auto light_edges = GraphMat::filter_edges(&E, less_than_delta, &delta);
  auto heavy_edges = GraphMat::filter_edges(&E, greater_than_delta, &delta);
  E.clear();

  GraphMat::Graph<DeltaSteppingDS> G;
  G.ReadEdgelist(light_edges);
  GraphMat::Graph<DeltaSteppingDS> G2;
  G2.ReadEdgelist(heavy_edges);
  light_edges.clear();
  heavy_edges.clear();
  G2.shareVertexProperty(G);

  DeltaStepping deltastep(delta);
  auto ds_ts = GraphMat::graph_program_init(deltastep, G);
  G.setAllInactive();
  DeltaSteppingDS src;
  src.distance = 0;
  src.bucket = 0;
  G.setVertexproperty(v, src);
  G.setActive(v);

  double t0 = now_ms();
  int bucket_not_empty = 1;
  do {
    G.setAllActive();
    GraphMat::run_graph_program(&deltastep, G, GraphMat::UNTIL_CONVERGENCE, &ds_ts);
    G2.setAllActive();
    GraphMat::run_graph_program(&deltastep, G2, 1, &ds_ts);
    deltastep.bid++;
    bucket_not_empty = 0;
    if (getenv("GM_HOST_REDUCE")) {
      G.applyReduceAllVertices(&bucket_not_empty, CheckBucketNotEmpty, Add<int>, (void*)&deltastep.bid);
    } else {
      CheckBucketNotEmptyFn map;
      map.bid = deltastep.bid;
      G.applyReduceAllVertices(&bucket_not_empty, map, AddIntFn());
    }
  } while (bucket_not_empty != 0);
  printf("Time = %.3f ms \n", now_ms() - t0);
  GraphMat::graph_program_clear(ds_ts);

  int reachable_vertices = 0;
  G.applyReduceAllVertices(&reachable_vertices, reachable_or_not);
  printf("Reachable vertices = %d , buckets = %d \n", reachable_vertices, deltastep.bid);
  for (int i = 1; i <= std::min(10, G.getNumberOfVertices()); i++) {
    if (G.getVertexproperty(i).distance < gm_sssp::kMaxDist) printf("%d : distance = %u\n", i, G.getVertexproperty(i).distance);
    else printf("%d : distance = INF\n", i);
  }
  if (dump) {
    FILE* f = fopen(dump, "w");
    for (int i = 1; i <= G.getNumberOfVertices(); i++)
      fprintf(f, "%d %u %d\n", i, G.getVertexproperty(i).distance, G.getVertexproperty(i).bucket);
    fclose(f);
  }
}

int main(int argc, char* argv[]) {
  if (argc < 4) {
    printf("Correct format: %s A.mtx delta source_vertex (1-based index)\n", argv[0]);
    return 0;
  }
  run_deltastepping(argv[1], atoi(argv[2]), atoi(argv[3]), dump_path(argc, argv));
  return 0;
}
