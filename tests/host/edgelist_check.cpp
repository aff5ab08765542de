// edgelist_check.cpp -- TEST INFRASTRUCTURE.  One driver, compiled twice: against this repository's
// GraphMat/edgelist.h (the product's host mirror) and, with -DUSE_REFERENCE, against the unmodified reference's
// include/GMDP/utils/edgelist.h + edgelist_transformation.h (oracle/Makefile -> oracle/_ref/edgelist_check).
// It runs the edge-list helpers that graph_converter does not reach and prints what they return on lines that
// start with '|' (the libraries' own chatter differs); those lines must be the same text (tests/test_graph_converter.py, golden in tests/golden/edgelist_check.txt).
// usage: edgelist_check <text mtx prefix>
#ifdef USE_REFERENCE
#include <mpi.h>
#include "GMDP/gmdp.h"
#else
#include "GraphMat/edgelist.h"
#endif
#include <cstdio>

using GraphMat::edge_t;
using GraphMat::edgelist_t;

static void print(const char* what, const edgelist_t<int>& e) {
  printf("| %s: %d x %d, %d edges\n", what, e.m, e.n, e.nnz);
  for (int i = 0; i < e.nnz; i++) printf("| %d %d %d\n", e.edges[i].src, e.edges[i].dst, e.edges[i].val);
}

static edgelist_t<int> copy_of(const edgelist_t<int>& e) {
  edgelist_t<int> c(e.m, e.n, e.nnz);
  for (int i = 0; i < e.nnz; i++) c.edges[i] = e.edges[i];
  return c;
}

static bool heavy_and_forward(edge_t<int> e, void* param) { return e.val >= *(int*)param && e.src < e.dst; }

int main(int argc, char** argv) {
#ifdef USE_REFERENCE
  MPI_Init(&argc, &argv);
#endif
  if (argc < 2) return 2;
  edgelist_t<int> in;
  GraphMat::ReadEdges<int>(&in, argv[1], false, true, true, false);
  print("read", in);

  int mm = 0, nn = 0;
  GraphMat::get_dimensions<int>(in.edges, in.nnz, mm, nn);
  printf("| dimensions: %d %d\n", mm, nn);

  edgelist_t<int> a = copy_of(in);
  int* kept = nullptr;
  GraphMat::remove_empty_columns<int>(&a, &kept);
  print("remove_empty_columns", a);
  printf("| remaining:");
  for (int i = 0; i < a.n; i++) printf(" %d", kept[i]);
  printf("\n");

  edgelist_t<int> b = copy_of(in);
  GraphMat::filter_edges_by_row<int>(&b, in.m / 4, in.m / 2);
  print("filter_edges_by_row", b);

  int threshold = 40;
  edgelist_t<int> c = GraphMat::filter_edges<int>(&in, heavy_and_forward, &threshold);
  print("filter_edges", c);

  edgelist_t<int> d = copy_of(in);
  srand(11);
  GraphMat::randomize_edge_direction<int>(&d);
  print("randomize_edge_direction", d);

  edgelist_t<int> e;
  GraphMat::ReadEdges<int>(&e, argv[1], false, true, true, true);  // with the vertex relabelling
  print("read randomized", e);
  return 0;
}
