// ApplyEdgesCheck: the reference's test/test_apply_edges.cpp:38-63 (Graph::applyToAllEdges followed by
// getEdgelist) on the device engine, plus a check that the DEVICE matrices carry the new values:
// SSSP over the rewritten weights must equal a host Bellman-Ford over the same weights.
// usage: ApplyEdgesCheck <N>      prints "apply_edges ok" on success
#include <unistd.h>

#include <algorithm>
#include <string>
#include <vector>

#include "GraphMatRuntime.h"
#include "GraphMat/programs/SSSP.h"
#include "common.h"

// src_vp.distance doubles as the "vertex id" payload of the reference test (V = int there)
void apply_edges_fn(int* edge_val, const SSSP_vertex_type& src_vp, const SSSP_vertex_type& dst_vp, void* vsp) {
  int s = *(int*)vsp;
  *edge_val = (int)src_vp.distance + s * (int)dst_vp.distance;
}

// the same three operations as GM_HD functors: evaluated on the DEVICE (gm_vertex_ops.cuh), nothing is pulled to the host
struct apply_edges_functor {
  int s;
  GM_HD void operator()(int* edge_val, const SSSP_vertex_type& src_vp, const SSSP_vertex_type& dst_vp) const {
    *edge_val = (int)src_vp.distance + s * (int)dst_vp.distance;
  }
};
struct set_id_functor {  // applyToAllVertices: distance := 3 * distance + 1
  GM_HD void operator()(const SSSP_vertex_type& in, SSSP_vertex_type* out) const { out->distance = 3 * in.distance + 1; }
};
struct sum_distance_map {  // applyReduceAllVertices: sum of the distances, as unsigned long long
  GM_HD void operator()(SSSP_vertex_type* v, unsigned long long* out) const { *out = v->distance; }
};
struct sum_reduce {
  GM_HD void operator()(const unsigned long long& a, const unsigned long long& b, unsigned long long* c) const { *c = a + b; }
};

// device-side variants: same inputs as check(), results compared with the host formulas
int check_device(GraphMat::edgelist_t<int> E) {
  GraphMat::Graph<SSSP_vertex_type, int> G;
  GraphMat::edgelist_t<int> E0(E.m, E.n, E.nnz);
  for (int i = 0; i < E.nnz; i++) E0.edges[i] = E.edges[i];
  G.ReadEdgelist(E);
  const int n = G.getNumberOfVertices();
  for (int i = 1; i <= n; i++) {
    SSSP_vertex_type v;
    v.distance = i;
    G.setVertexproperty(i, v);
  }
  G.applyToAllVertices(set_id_functor());  // distance(i) = 3 i + 1
  unsigned long long sum = 0;
  G.applyReduceAllVertices(&sum, sum_distance_map(), sum_reduce());
  unsigned long long expect = 0;
  for (int i = 1; i <= n; i++) expect += 3ull * i + 1;
  if (sum != expect) return 11;
  for (int i = 1; i <= n; i += 97)
    if (G.getVertexproperty(i).distance != 3u * i + 1) return 12;
  apply_edges_functor f;
  f.s = 2;
  G.applyToAllEdges(f);  // w(u, v) = (3u + 1) + 2 (3v + 1), in both device matrices
  SSSP_vertex_type inf, zero;
  zero.distance = 0;
  G.setAllVertexproperty(inf);
  G.setAllInactive();
  G.setVertexproperty(1, zero);
  G.setActive(1);
  SSSP<int> prog;
  GraphMat::run_graph_program(&prog, G, GraphMat::UNTIL_CONVERGENCE);
  std::vector<unsigned> dist(n + 1, gm_sssp::kMaxDist);
  dist[1] = 0;
  for (int it = 0; it < n; it++) {
    bool ch = false;
    for (int i = 0; i < E0.nnz; i++) {
      const auto& e = E0.edges[i];
      const unsigned w = (3u * e.src + 1) + 2u * (3u * e.dst + 1);
      if (dist[e.src] != gm_sssp::kMaxDist && dist[e.src] + w < dist[e.dst]) {
        dist[e.dst] = dist[e.src] + w;
        ch = true;
      }
    }
    if (!ch) break;
  }
  for (int i = 1; i <= n; i++)
    if (G.getVertexproperty(i).distance != dist[i]) return 13;
  E0.clear();
  return 0;
}

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int check(GraphMat::edgelist_t<int> E) {
  GraphMat::Graph<SSSP_vertex_type, int> G;
  G.ReadEdgelist(E);
  const int n = G.getNumberOfVertices();
  for (int i = 1; i <= n; i++) {
    SSSP_vertex_type v;
    v.distance = i;
    G.setVertexproperty(i, v);
  }
  int s = 2;
  G.applyToAllEdges(apply_edges_fn, (void*)&s);
  GraphMat::edgelist_t<int> E2;
  G.getEdgelist(E2);
  if (E2.nnz != E.nnz) return 1;
  for (int i = 0; i < E2.nnz; i++)
    if (E2.edges[i].val != E2.edges[i].src + s * E2.edges[i].dst) return 2;  // test_apply_edges.cpp:62-65
  // the device matrices: SSSP from vertex 1 over the new weights vs host Bellman-Ford
  SSSP_vertex_type inf, zero;
  zero.distance = 0;
  G.setAllVertexproperty(inf);
  G.setAllInactive();
  G.setVertexproperty(1, zero);
  G.setActive(1);
  SSSP<int> prog;
  GraphMat::run_graph_program(&prog, G, GraphMat::UNTIL_CONVERGENCE);
  std::vector<unsigned> dist(n + 1, gm_sssp::kMaxDist);
  dist[1] = 0;
  for (int it = 0; it < n; it++) {
    bool ch = false;
    for (int i = 0; i < E2.nnz; i++) {
      const auto& e = E2.edges[i];
      if (dist[e.src] != gm_sssp::kMaxDist && dist[e.src] + (unsigned)e.val < dist[e.dst]) {
        dist[e.dst] = dist[e.src] + (unsigned)e.val;
        ch = true;
      }
    }
    if (!ch) break;
  }
  for (int i = 1; i <= n; i++)
    if (G.getVertexproperty(i).distance != dist[i]) return 3;
  E2.clear();
  return 0;
}

int main(int argc, char* argv[]) {
  const int N = argc > 1 ? atoi(argv[1]) : 500;
  // generate_identity_edgelist / generate_random_edgelist of test/generator.h
  GraphMat::edgelist_t<int> I(N, N, N);
  for (int i = 0; i < N; i++) I.edges[i] = GraphMat::edge_t<int>(i + 1, i + 1, 1);
  int rc = check(I);
  if (rc) { printf("apply_edges identity FAILED (%d)\n", rc); return 1; }
  I.clear();
  const int nnz = N * 16;
  GraphMat::edgelist_t<int> R(N, N, nnz);
  unsigned seed = 7;
  for (int i = 0; i < nnz; i++) R.edges[i] = GraphMat::edge_t<int>(1 + lcg(seed) % N, 1 + lcg(seed) % N, 1 + lcg(seed) % 9);
  rc = check(R);
  if (rc) { printf("apply_edges random FAILED (%d)\n", rc); return 1; }
  rc = check_device(R);
  if (rc) { printf("apply_edges device functors FAILED (%d)\n", rc); return 1; }
  printf("device functors ok\n");
  {  // snapshot round trip (Graph.h:152-208 re-specified): same graph back, SSSP distances equal
    GraphMat::edgelist_t<int> R2(N, N, nnz);
    for (int i = 0; i < nnz; i++) R2.edges[i] = R.edges[i];
    GraphMat::Graph<SSSP_vertex_type, int> G1, G2;
    G1.ReadEdgelist(R2);
    char path[256];
    snprintf(path, sizeof path, "/tmp/gm_snapshot_%d_", (int)getpid());
    G1.WriteGraphMatBin(path);
    G2.ReadGraphMatBin(path);
    remove((std::string(path) + "0").c_str());
    if (G2.getNumberOfVertices() != N || G2.nnz != nnz) { printf("snapshot FAILED (shape)\n"); return 1; }
    SSSP_vertex_type zero;
    zero.distance = 0;
    GraphMat::Graph<SSSP_vertex_type, int>* gs[2] = {&G1, &G2};
    for (auto* g : gs) {
      g->setAllInactive();
      g->setVertexproperty(1, zero);
      g->setActive(1);
      SSSP<int> prog;
      GraphMat::run_graph_program(&prog, *g, GraphMat::UNTIL_CONVERGENCE);
    }
    for (int i = 1; i <= N; i++)
      if (G1.getVertexproperty(i).distance != G2.getVertexproperty(i).distance) { printf("snapshot FAILED (%d)\n", i); return 1; }
    printf("snapshot ok\n");
  }
  R.clear();
  printf("apply_edges ok\n");
  return 0;
}
