// SSSP.h -- the SSSP and DeltaStepping vertex programs of the reference apps
// (narayanan2004/GraphMat src/SSSP.cpp:36-100, src/DeltaStepping.cpp:37-122), annotated GM_HD.
#ifndef GRAPHMAT_B200_PROGRAMS_SSSP_H
#define GRAPHMAT_B200_PROGRAMS_SSSP_H
#include "../GraphProgram.h"

namespace gm_sssp {
typedef unsigned int distance_type;
static const distance_type kMaxDist = 0xffffffffu;  // src/SSSP.cpp:41
GM_HD inline distance_type umin(distance_type a, distance_type b) { return a < b ? a : b; }
}

class SSSP_vertex_type {  // src/SSSP.cpp:43-60
 public:
  gm_sssp::distance_type distance;
  GM_HD SSSP_vertex_type() : distance(gm_sssp::kMaxDist) {}
  GM_HD bool operator!=(const SSSP_vertex_type& p) const { return distance != p.distance; }
};

template <class edge_type>
class SSSP : public GraphMat::GraphProgram<gm_sssp::distance_type, gm_sssp::distance_type, SSSP_vertex_type, edge_type> {
 public:
  typedef gm_sssp::distance_type distance_type;
  static const bool gm_reorderable = true;  // min
  static const bool gm_atomic_min = true;   // ... of 32-bit unsigned values: sparse passes fold with atomicMin
  GM_HD SSSP() {
    this->order = GraphMat::OUT_EDGES;
    this->process_message_requires_vertexprop = false;
  }
  GM_HD void reduce_function(distance_type& a, const distance_type& b) const { a = (a <= b) ? a : b; }
  GM_HD void process_message(const distance_type& message, const edge_type edge_val, const SSSP_vertex_type& vertexprop,
                             distance_type& res) const {
    res = message + edge_val;
  }
  GM_HD bool send_message(const SSSP_vertex_type& vertexprop, distance_type& message) const {
    message = vertexprop.distance;
    return true;
  }
  GM_HD void apply(const distance_type& message_out, SSSP_vertex_type& vertexprop) {
    vertexprop.distance = gm_sssp::umin(vertexprop.distance, message_out);
  }
};

class DeltaSteppingDS {  // src/DeltaStepping.cpp:42-62
 public:
  gm_sssp::distance_type distance;
  int bucket;
  GM_HD DeltaSteppingDS() : distance(gm_sssp::kMaxDist), bucket(0x7fffffff) {}
  GM_HD bool operator!=(const DeltaSteppingDS& p) const { return distance != p.distance; }
};

class DeltaStepping : public GraphMat::GraphProgram<gm_sssp::distance_type, gm_sssp::distance_type, DeltaSteppingDS> {
 public:
  typedef gm_sssp::distance_type distance_type;
  int delta;
  int bid;
  static const bool gm_reorderable = true;  // min
  static const bool gm_atomic_min = true;   // ... of 32-bit unsigned values: sparse passes fold with atomicMin
  GM_HD DeltaStepping(int d = 1) {
    delta = d;
    bid = 0;
    this->order = GraphMat::OUT_EDGES;
    this->process_message_requires_vertexprop = false;
  }
  GM_HD void reduce_function(distance_type& a, const distance_type& b) const { a = (a <= b) ? a : b; }
  GM_HD void process_message(const distance_type& message, const int edge_val, const DeltaSteppingDS& vertex,
                             distance_type& res) const {
    res = (message < gm_sssp::kMaxDist) ? (message + edge_val) : gm_sssp::kMaxDist;
  }
  // a vertex outside the current bucket sends MAX_DIST, which process_message hands on as MAX_DIST and apply
  // ignores (distance > MAX_DIST is never true): the engine may drop such messages (gm_engine.cuh, has_null_message)
  static GM_HD bool gm_null_message(const distance_type& m) { return m == gm_sssp::kMaxDist; }
  GM_HD bool send_message(const DeltaSteppingDS& vertex, distance_type& message) const {
    message = (vertex.bucket == bid) ? vertex.distance : gm_sssp::kMaxDist;
    return true;
  }
  GM_HD void apply(const distance_type& message_out, DeltaSteppingDS& vertex) {
    if (vertex.distance > message_out) {
      vertex.distance = message_out;
      vertex.bucket = (int)(message_out / delta);
    }
  }
};
#endif
