// TopologicalSort.h -- the InDegree and TopSort vertex programs of the reference app
// (narayanan2004/GraphMat src/TopologicalSort.cpp:36-130), annotated GM_HD so the device engine can call them.
// TopSort peels the DAG level by level: a vertex whose in-degree just reached zero tells its out-neighbours,
// which subtract what they hear; vertices on a cycle keep topsort_order == MAX_DIST.
#ifndef GRAPHMAT_B200_PROGRAMS_TOPOLOGICALSORT_H
#define GRAPHMAT_B200_PROGRAMS_TOPOLOGICALSORT_H
#include <cassert>
#include "../GraphProgram.h"

namespace gm_topsort {
typedef unsigned int depth_type;
static const depth_type kMaxDist = 0xffffffffu;  // src/TopologicalSort.cpp:37
}

class TopSortVertex {  // Vertex_type, :39-58
 public:
  gm_topsort::depth_type topsort_order;
  int in_degree;
  GM_HD TopSortVertex() : topsort_order(gm_topsort::kMaxDist), in_degree(0) {}
  GM_HD bool operator!=(const TopSortVertex& p) const { return topsort_order != p.topsort_order; }
};

template <class V, class E = int>
class InDegree : public GraphMat::GraphProgram<int, int, V, E> {  // :60-87
 public:
  static const bool gm_reorderable = true;  // integer +
  GM_HD InDegree() {
    this->activity = GraphMat::ALL_VERTICES;
    this->order = GraphMat::OUT_EDGES;
    this->process_message_requires_vertexprop = false;
  }
  GM_HD bool send_message(const V& vertex, int& message) const {
    message = 1;
    return true;
  }
  GM_HD void process_message(const int& message, const E edge_value, const V& vertex, int& result) const { result = message; }
  GM_HD void reduce_function(int& a, const int& b) const { a += b; }
  GM_HD void apply(const int& message_out, V& vertex) { vertex.in_degree = message_out; }
};

class TopSort : public GraphMat::GraphProgram<bool, int, TopSortVertex> {  // :90-130
 public:
  gm_topsort::depth_type current_topsort_order;
  static const bool gm_reorderable = true;  // integer +
  GM_HD TopSort() {
    current_topsort_order = 1;
    this->order = GraphMat::OUT_EDGES;
    this->process_message_requires_vertexprop = false;
  }
  GM_HD void reduce_function(int& a, const int& b) const { a += b; }
  GM_HD void process_message(const bool& message, const int edge_val, const TopSortVertex& vertex, int& res) const {
    res = (message == true) ? (1) : (0);
  }
  GM_HD bool send_message(const TopSortVertex& vertex, bool& message) const {
    message = (vertex.in_degree == 0) ? true : false;
    return true;
  }
  GM_HD void apply(const int& message_out, TopSortVertex& vertex) {
    vertex.in_degree -= message_out;
    if (vertex.in_degree == 0) vertex.topsort_order = current_topsort_order;
  }
  void do_every_iteration(int iteration_number) { current_topsort_order++; }
};
#endif
