// PageRank.h -- the PageRank and Degree vertex programs of the reference app
// (narayanan2004/GraphMat src/PageRank.cpp:34-112), annotated GM_HD so the device
// engine can call them.  Operator bodies compute exactly what the reference's do.
#ifndef GRAPHMAT_B200_PROGRAMS_PAGERANK_H
#define GRAPHMAT_B200_PROGRAMS_PAGERANK_H
#include <cmath>
#include "../GraphProgram.h"

// vertex property (src/PageRank.cpp:34-52): rank starts at 0.3, out-degree at 0
class PR {
 public:
  float pagerank;
  int degree;
  GM_HD PR() : pagerank(0.3f), degree(0) {}
  // "changed" = moved by more than 1e-5 (:44-46); the float difference is compared in double
  GM_HD int operator!=(const PR& p) const { return fabs((double)(p.pagerank - pagerank)) > 1e-5; }
};

// src/PageRank.cpp:54-79: every vertex sends 1 along its in-edges, so a vertex receives its out-degree
template <class V, class E = int>
class Degree : public GraphMat::GraphProgram<int, int, V, E> {
 public:
  static const bool gm_reorderable = true;  // integer +
  GM_HD Degree() {
    this->order = GraphMat::IN_EDGES;
    this->process_message_requires_vertexprop = false;
  }
  GM_HD bool send_message(const V& vertexprop, int& message) const {
    message = 1;
    return true;
  }
  GM_HD void process_message(const int& message, const E edge_value, const V& vertexprop, int& result) const {
    result = message;
  }
  GM_HD void reduce_function(int& a, const int& b) const { a += b; }
  GM_HD void apply(const int& message_out, V& vertexprop) { vertexprop.degree = message_out; }
};

// src/PageRank.cpp:81-112
template <class E>
class PageRank : public GraphMat::GraphProgram<float, float, PR, E> {
 public:
  float alpha;
  static const bool gm_fadd32_exact = true;  // fp32 +, messages >= 0: long rows may use the exact emulation

  GM_HD PageRank(float a = 0.3f) {
    alpha = a;
    this->activity = GraphMat::ALL_VERTICES;
    this->process_message_requires_vertexprop = false;
  }
  GM_HD void reduce_function(float& a, const float& b) const { a += b; }
  GM_HD void process_message(const float& message, const E edge_val, const PR& vertexprop, float& res) const {
    res = message;
  }
  GM_HD bool send_message(const PR& vertexprop, float& message) const {
    if (vertexprop.degree == 0) message = 0.0f;
    else message = vertexprop.pagerank / (float)vertexprop.degree;
    return true;
  }
  GM_HD void apply(const float& message_out, PR& vertexprop) {
    // (1.0 - alpha) and the product/sum are double in the reference (:109), then narrowed
    vertexprop.pagerank = (float)((double)alpha + (1.0 - (double)alpha) * (double)message_out);
  }
};
#endif
