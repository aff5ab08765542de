// Graph.h -- host-side mirror of GraphMat::Graph<V,E> (narayanan2004/GraphMat
// include/Graph.h:58-107) over the C ABI: same members, argument meaning and ownership.
// The adjacency matrices, vertex properties and active set live in HBM behind a gm_graph
// handle; per-vertex accessors (the apps call them in loops over all vertices,
// src/BFS.cpp:114-119, src/SGD.cpp:176-184) go through a lazily synchronised host mirror
// so they cost O(1) each instead of one PCIe round trip each.
//
// Differences a maintainer should know: E must be a 4-byte type; the
// GraphMat-binary snapshot (ReadGraphMatBin / WriteGraphMatBin) is re-specified without Boost;
// applyToAllEdges & co. evaluate host function pointers on the host and GM_HD functors on the device.
#ifndef GRAPHMAT_B200_GRAPH_H
#define GRAPHMAT_B200_GRAPH_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "edgelist.h"
#include "graphmat_b200.h"
#ifdef __CUDACC__
#include <type_traits>

#include "gm_vertex_ops.cuh"
#endif

namespace GraphMat {

template <class T>
void AddFn(const T& a, const T& b, T* c, void* vsp) {
  *c = a + b;
}

namespace detail {
inline void check(int rc, const char* what) {
  if (rc != 0) {
    printf("graphmat_b200: %s failed: %s\n", what, gm_last_error());
    exit(1);
  }
}
// the vertex-property array as the host sees it (public-id order), shared by graphs that
// share their vertex property (shareVertexProperty)
template <class V>
struct vp_mirror {
  std::vector<V> host;
  bool host_valid = false;   // host copy equals the device copy
  bool host_dirty = false;   // host copy has writes the device has not seen
};
}  // namespace detail

template <class V, class E = int>
class Graph {
 public:
  int nvertices;
  long long int nnz;
  bool vertexpropertyowner;
  int tiles_per_dim;
  int num_threads;   // fixes the reference's vertex permutation (Graph.h:117); env GM_REF_THREADS, default 4
  gm_graph* handle;  // A, AT, vertexproperty, active (device)

 private:
  std::shared_ptr<detail::vp_mirror<V> > mirror;
  std::vector<int> e_src, e_dst;  // kept so that shareVertexProperty can rebuild in the owner's placement
  std::vector<E> e_val;
  bool edges_on_host_stale = false;

  static int ref_threads() {
    const char* e = getenv("GM_REF_THREADS");
    int t = e ? atoi(e) : 4;
    return t > 0 ? t : 4;
  }
  void build(const gm_graph* like) {
    static_assert(sizeof(E) == 4, "graphmat_b200: edge values must be 4 bytes (int / float / unsigned)");
    static_assert(sizeof(V) % 4 == 0, "graphmat_b200: sizeof(V) must be a multiple of 4");
    gm_graph_opts o;
    memset(&o, 0, sizeof o);
    o.ref_threads = num_threads;
    o.order_like = like;
    if (handle) gm_graph_destroy(handle);
    handle = nullptr;
    detail::check(gm_graph_create(&handle, nvertices, nnz, e_src.data(), e_dst.data(), e_val.data(), (int)sizeof(E),
                                  (int)sizeof(V), &o), "gm_graph_create");
  }

 public:
  Graph()
      : nvertices(0), nnz(0), vertexpropertyowner(true), tiles_per_dim(1), num_threads(ref_threads()), handle(nullptr),
        mirror(new detail::vp_mirror<V>()) {}
  Graph(const Graph&) = delete;
  Graph& operator=(const Graph&) = delete;
  ~Graph() {
    if (handle) gm_graph_destroy(handle);
  }

  // Graph.h:210-246: permute ids, build A and AT, vertexproperty = V(), active = false
  void ReadEdgelist(GraphMat::edgelist_t<E> A_edges) {
    if (A_edges.m != A_edges.n) {
      printf("graphmat_b200: ReadEdgelist needs a square matrix (got %d x %d)\n", A_edges.m, A_edges.n);
      exit(1);
    }
    nvertices = A_edges.m;
    nnz = A_edges.nnz;
    e_src.resize(nnz);
    e_dst.resize(nnz);
    e_val.resize(nnz);
    for (long long i = 0; i < nnz; i++) {
      e_src[i] = A_edges.edges[i].src;
      e_dst[i] = A_edges.edges[i].dst;
      e_val[i] = A_edges.edges[i].val;
    }
    build(nullptr);
    V v0;
    mirror->host.assign(nvertices, v0);
    mirror->host_valid = true;
    mirror->host_dirty = true;
    vertexpropertyowner = true;
  }
  // Graph.h:248-260
  void ReadMTX(const char* filename) {
    GraphMat::edgelist_t<E> A_edges;
    GraphMat::load_edgelist(filename, &A_edges, true, true, true);
    if (A_edges.m != A_edges.n) {
      int maxn = std::max(A_edges.m, A_edges.n);
      A_edges.m = maxn;
      A_edges.n = maxn;
    }
    ReadEdgelist(A_edges);
    A_edges.clear();
  }
  // Graph.h:152-208.  The reference's snapshot is a Boost binary archive of its DCSC tiles (A, AT) and the OpenMP
  // thread count; Boost is not part of this build and the tiles are not this engine's layout, so the snapshot is
  // re-specified: little-endian  "GMB200\0\1" | int nvertices | int num_threads | int sizeof(E) |
  // long long nnz | nnz x (int src, int dst, E val), public ids.  Same contract as the reference: one file per rank
  // (<filename>0), vertex properties and the active set are NOT part of it (a loaded graph starts from V() /
  // inactive), and loading under a different thread count is refused because it would change the fold order.
  void WriteGraphMatBin(const char* filename) {
    const std::string fn = std::string(filename) + "0";
    std::cout << "Writing file " << fn << std::endl;
    std::ofstream out(fn.c_str(), std::ios::out | std::ios::binary);
    if (!out) { printf("graphmat_b200: cannot write %s\n", fn.c_str()); exit(1); }
    if (edges_on_host_stale) unsupported("WriteGraphMatBin after a device-side applyToAllEdges");
    const char magic[8] = {'G', 'M', 'B', '2', '0', '0', 0, 1};
    const int se = (int)sizeof(E);
    out.write(magic, 8);
    out.write((const char*)&nvertices, 4);
    out.write((const char*)&num_threads, 4);
    out.write((const char*)&se, 4);
    out.write((const char*)&nnz, 8);
    for (long long i = 0; i < nnz; i++) {
      out.write((const char*)&e_src[i], 4);
      out.write((const char*)&e_dst[i], 4);
      out.write((const char*)&e_val[i], sizeof(E));
    }
  }
  void ReadGraphMatBin(const char* filename) {
    const std::string fn = std::string(filename) + "0";
    std::cout << "Reading file " << fn << std::endl;
    std::ifstream in(fn.c_str(), std::ios::in | std::ios::binary);
    char magic[8] = {0};
    int nv = 0, nt = 0, se = 0;
    long long nz = 0;
    in.read(magic, 8);
    in.read((char*)&nv, 4);
    in.read((char*)&nt, 4);
    in.read((char*)&se, 4);
    in.read((char*)&nz, 8);
    if (!in || memcmp(magic, "GMB200", 6) != 0 || se != (int)sizeof(E) || nv <= 0 || nz < 0) {
      std::cout << "Error reading file - not a graphmat_b200 snapshot of this edge type" << std::endl;
      exit(1);
    }
    if (nt != num_threads) {
      std::cout << "Error reading file - mismatch in number of OpenMP threads used in load vs save graph" << std::endl;
      exit(1);
    }
    GraphMat::edgelist_t<E> E_(nv, nv, (int)nz);
    for (long long i = 0; i < nz; i++) {
      in.read((char*)&E_.edges[i].src, 4);
      in.read((char*)&E_.edges[i].dst, 4);
      in.read((char*)&E_.edges[i].val, sizeof(E));
    }
    if (!in) { std::cout << "Error reading file - truncated snapshot" << std::endl; exit(1); }
    ReadEdgelist(E_);
    E_.clear();
  }
  void getEdgelist(GraphMat::edgelist_t<E>& out) {
    if (edges_on_host_stale) unsupported("getEdgelist after a device-side applyToAllEdges");
    out = GraphMat::edgelist_t<E>(nvertices, nvertices, (int)nnz);
    for (long long i = 0; i < nnz; i++) out.edges[i] = GraphMat::edge_t<E>(e_src[i], e_dst[i], e_val[i]);
  }

  void setAllActive() { detail::check(gm_graph_set_all_active(handle), "gm_graph_set_all_active"); }
  void setAllInactive() { detail::check(gm_graph_set_all_inactive(handle), "gm_graph_set_all_inactive"); }
  void setActive(int v) { detail::check(gm_graph_set_active(handle, v), "gm_graph_set_active"); }
  void setInactive(int v) { detail::check(gm_graph_set_inactive(handle, v), "gm_graph_set_inactive"); }

  void setAllVertexproperty(const V& val) {
    mirror->host.assign(nvertices, val);
    mirror->host_valid = true;
    mirror->host_dirty = true;
  }
  void setVertexproperty(int v, const V& val) {
    pull();
    mirror->host[v - 1] = val;
    mirror->host_dirty = true;
  }
  V getVertexproperty(int v) const {
    const_cast<Graph*>(this)->pull();
    return mirror->host[v - 1];
  }
  bool vertexNodeOwner(const int v) const { return gm_graph_vertex_owner(handle, v) == 0; }
  void saveVertexproperty(std::string fname, bool includeHeader = true) const {
    const_cast<Graph*>(this)->pull();
    // the reference's text format (DenseSegment::save, include/GMDP/vectors/DenseSegment.h): "<id> <value>" lines
    // through V's operator<<, behind an "<n> <count>" header; one file per rank, this build is rank 0
    std::ofstream out((fname + "0").c_str());
    if (includeHeader) out << nvertices << " " << nvertices << "\n";
    for (int i = 0; i < nvertices; i++) out << (i + 1) << " " << mirror->host[i] << "\n";
  }
  void reset() {
    setAllInactive();
    V v;
    setAllVertexproperty(v);
  }
  // Graph.h:300-305.  The two graphs must agree on where each vertex lives, so this graph is
  // rebuilt in g's placement before it adopts g's vertex-property storage.
  void shareVertexProperty(Graph<V, E>& g) {
    g.push();
    build(g.handle);
    detail::check(gm_graph_share_vertexproperty(handle, g.handle), "gm_graph_share_vertexproperty");
    mirror = g.mirror;
    vertexpropertyowner = false;
  }
  int getNumberOfVertices() const { return nvertices; }

  // Graph.h:371-381: host-side map / map-reduce over all vertex properties (the apps use them
  // between run_graph_program calls, outside the timed path)
  void applyToAllVertices(void (*ApplyFn)(const V&, V*, void*), void* param = nullptr) {
    pull();
    for (int i = 0; i < nvertices; i++) {
      V in = mirror->host[i];
      ApplyFn(in, &mirror->host[i], param);
    }
    mirror->host_dirty = true;
  }
  template <class T>
  void applyReduceAllVertices(T* val, void (*ApplyFn)(V*, T*, void*),
                              void (*ReduceFn)(const T&, const T&, T*, void*) = AddFn<T>, void* param = nullptr) {
    pull();
    bool first = true;
    for (int i = 0; i < nvertices; i++) {
      T t;
      ApplyFn(&mirror->host[i], &t, param);
      if (first) { *val = t; first = false; }
      else { T a = *val; ReduceFn(a, t, val, param); }
    }
  }
  // Graph.h:389-402 / GMDP/singlenode/applyedges.h:38-76: ApplyFn(&edge, vp[src], vp[dst], param) for every
  // edge, on both A and AT.  ApplyFn is host code (the reference's signature), so it is evaluated here
  // on the host mirror and the device matrices are refilled in place (gm_graph_set_edge_values).
  void applyToAllEdges(void (*ApplyFn)(E*, const V&, const V&, void*), void* param = nullptr) {
    pull();
    for (long long i = 0; i < nnz; i++)
      ApplyFn(&e_val[i], mirror->host[e_src[i] - 1], mirror->host[e_dst[i] - 1], param);
    push();
    detail::check(gm_graph_set_edge_values(handle, nnz, e_src.data(), e_dst.data(), e_val.data()),
                  "gm_graph_set_edge_values");
  }

#ifdef __CUDACC__
  // ---- the same three, on the DEVICE, for GM_HD functors (gm_vertex_ops.cuh): no pull of the vertex array, the
  //      functor's state replaces `param`.  Function pointers keep resolving to the host overloads above. ----
  template <class F, class = typename std::enable_if<!std::is_pointer<F>::value && !std::is_function<F>::value>::type>
  void applyToAllVertices(F f) {
    push();
    detail::check(gm::map_vertices<V, F>(handle, f), "applyToAllVertices (device)");
    invalidate();
  }
  template <class T, class M, class R,
            class = typename std::enable_if<!std::is_pointer<M>::value && !std::is_function<M>::value>::type>
  void applyReduceAllVertices(T* val, M map, R reduce) {
    push();
    detail::check(gm::map_reduce_vertices<V, T, M, R>(handle, val, map, reduce), "applyReduceAllVertices (device)");
  }
  template <class F, class = typename std::enable_if<!std::is_pointer<F>::value && !std::is_function<F>::value>::type>
  void applyToAllEdges(F f) {
    push();
    detail::check(gm::apply_edges<V, E, F>(handle, f), "applyToAllEdges (device)");
    edges_on_host_stale = true;  // e_val no longer mirrors the device matrices (getEdgelist refuses)
  }
#endif

  // ---- used by run_graph_program ----
  void push() {  // host writes -> device
    if (mirror->host_dirty && handle) {
      detail::check(gm_graph_set_vertexproperties(handle, mirror->host.data()), "gm_graph_set_vertexproperties");
      mirror->host_dirty = false;
      mirror->host_valid = true;
    }
  }
  void invalidate() { mirror->host_valid = false; }  // the device copy changed

 private:
  void pull() {  // device -> host, once per run
    if (!mirror->host_valid && handle) {
      mirror->host.resize(nvertices);
      detail::check(gm_graph_get_vertexproperties(handle, mirror->host.data()), "gm_graph_get_vertexproperties");
      mirror->host_valid = true;
      mirror->host_dirty = false;
    }
  }
  static void unsupported(const char* what) {
    printf("graphmat_b200: %s is outside the hot-path build (see INTEGRATION.md)\n", what);
    exit(1);
  }
};

}  // namespace GraphMat
#endif
