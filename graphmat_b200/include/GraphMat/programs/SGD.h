// SGD.h -- the collaborative-filtering programs of the reference app
// (narayanan2004/GraphMat src/SGD.cpp:36-161), annotated GM_HD.
#ifndef GRAPHMAT_B200_PROGRAMS_SGD_H
#define GRAPHMAT_B200_PROGRAMS_SGD_H
#include <cmath>
#include "../GraphProgram.h"

template <unsigned int K>
class LatentVector {  // src/SGD.cpp:36-75; the constructor leaves the members uninitialised, as there
 public:
  double lv[K];
  double sqerr;
  GM_HD LatentVector() {}
  GM_HD bool operator!=(const LatentVector<K>& p) const {
    bool result = false;
    for (unsigned int i = 0; i < K; i++)
      if (fabs(p.lv[i] - lv[i]) > 1e-7) result = true;
    return result;
  }
};

template <unsigned int K>
class SGDProgram : public GraphMat::GraphProgram<LatentVector<K>, LatentVector<K>, LatentVector<K> > {  // :77-121
 public:
  double lambda;
  double step;
  // fp64 vector +: re-association moves results by ~1e-16 relative, far inside the 1e-6 parity bound
  static const bool gm_reorderable = true;
  GM_HD SGDProgram(double l = 0.001, double s = 0.00000035) {
    lambda = l;
    step = s;
    this->order = GraphMat::ALL_EDGES;
    this->activity = GraphMat::ALL_VERTICES;
  }
  GM_HD void reduce_function(LatentVector<K>& v, const LatentVector<K>& w) const {
    for (unsigned int i = 0; i < K; i++) v.lv[i] += w.lv[i];
  }
  GM_HD void process_message(const LatentVector<K>& message, const int edge_val, const LatentVector<K>& vertexprop,
                             LatentVector<K>& res) const {
    double estimate = 0;
    for (unsigned int i = 0; i < K; i++) estimate += message.lv[i] * vertexprop.lv[i];
    double error = edge_val - estimate;
    for (unsigned int i = 0; i < K; i++) res.lv[i] = message.lv[i] * error;
  }
  GM_HD bool send_message(const LatentVector<K>& vertexprop, LatentVector<K>& message) const {
    message = vertexprop;
    return true;
  }
  GM_HD void apply(const LatentVector<K>& message_out, LatentVector<K>& vertexprop) {
    for (unsigned int i = 0; i < K; i++) vertexprop.lv[i] += step * (-lambda * vertexprop.lv[i] + message_out.lv[i]);
  }
};

template <unsigned int K>
class RMSEProgram : public GraphMat::GraphProgram<LatentVector<K>, double, LatentVector<K> > {  // :123-156
 public:
  static const bool gm_reorderable = true;
  GM_HD RMSEProgram() { this->order = GraphMat::IN_EDGES; }
  GM_HD void reduce_function(double& v, const double& w) const { v += w; }
  GM_HD void process_message(const LatentVector<K>& message, const int edge_val, const LatentVector<K>& vertexprop,
                             double& res) const {
    double est = 0;
    for (unsigned int i = 0; i < K; i++) est += message.lv[i] * vertexprop.lv[i];
    double error = edge_val - est;
    res = error * error;
  }
  GM_HD bool send_message(const LatentVector<K>& vertexprop, LatentVector<K>& message) const {
    message = vertexprop;
    return true;
  }
  GM_HD void apply(const double& message_out, LatentVector<K>& vertexprop) { vertexprop.sqerr = message_out; }
};
#endif
