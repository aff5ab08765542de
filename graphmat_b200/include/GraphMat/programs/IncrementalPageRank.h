// IncrementalPageRank.h -- the delta-PageRank vertex program of the reference app
// (narayanan2004/GraphMat src/IncrementalPageRank.cpp:33-123), annotated GM_HD so the device engine can
// call it.  ACTIVE_ONLY + fp64: only vertices whose rank moved by more than 1e-8 keep sending their delta.
// The fp64 sum is order-sensitive and no trait is declared, so every row is folded in the reference's order.
#ifndef GRAPHMAT_B200_PROGRAMS_INCREMENTALPAGERANK_H
#define GRAPHMAT_B200_PROGRAMS_INCREMENTALPAGERANK_H
#include <cmath>
#include "../GraphProgram.h"

class dPR {  // src/IncrementalPageRank.cpp:33-50
 public:
  double delta;
  double pagerank;
  int degree;
  GM_HD dPR() : delta(0.3), pagerank(0.3), degree(0) {}
  GM_HD int operator!=(const dPR& p) const { return fabs(p.pagerank - pagerank) > 1e-8; }
};

class DeltaPageRank : public GraphMat::GraphProgram<double, double, dPR> {  // :80-123
 public:
  double alpha;
  int iter;
  GM_HD DeltaPageRank(double a = 0.3) {
    alpha = a;
    iter = 0;
    this->order = GraphMat::OUT_EDGES;
    this->activity = GraphMat::ACTIVE_ONLY;
    this->process_message_requires_vertexprop = false;
  }
  GM_HD void reduce_function(double& a, const double& b) const { a += b; }
  GM_HD void process_message(const double& message, const int edge_val, const dPR& vertexprop, double& res) const {
    res = message;
  }
  GM_HD bool send_message(const dPR& vertexprop, double& message) const {
    if (vertexprop.degree == 0) message = 0.0;
    else message = vertexprop.delta / (double)vertexprop.degree;
    return true;
  }
  GM_HD void apply(const double& message_out, dPR& vertexprop) {
    if (fabs(vertexprop.delta) > 1e-8) vertexprop.delta = 0.0;
    vertexprop.delta += (1.0 - alpha) * message_out;
    if (fabs(vertexprop.delta) > 1e-8) vertexprop.pagerank += vertexprop.delta;
  }
  void do_every_iteration(int iteration_number) { iter++; }
};
#endif
