// GraphProgram.h -- host-side mirror of the reference's vertex-program surface
// (narayanan2004/GraphMat include/GraphProgram.h:34-101): same names, template
// parameters, protected knobs, getters and the five virtuals with the same
// const-ness.  An un-overridden operator prints and exits like the reference's
// (:73-96) on the host and traps on the device.
#ifndef GRAPHMAT_B200_GRAPHPROGRAM_H
#define GRAPHMAT_B200_GRAPHPROGRAM_H
#include <cstdio>
#include <cstdlib>
#include "gm_hd.h"

namespace GraphMat {

enum edge_direction { OUT_EDGES, IN_EDGES, ALL_EDGES };
enum activity_type { ACTIVE_ONLY, ALL_VERTICES };

namespace detail {
GM_HD inline void null_operator(const char* what) {
#if defined(__CUDA_ARCH__)
  printf("Trying to use default (null) %s\n", what);
  __trap();
#else
  printf("Trying to use default (null) %s\n", what);
  exit(1);
#endif
}
}  // namespace detail

// T: message, U: reduced message, V: vertex property, E: edge value
template <class T, class U, class V, class E = int>
class GraphProgram {
 public:
  typedef T message_type;
  typedef U message_reduction_type;
  typedef V vertex_property_type;
  typedef E edge_type;

 protected:
  edge_direction order;
  activity_type activity;
  bool process_message_requires_edge_value;  // "strictly for optimization only"
  bool process_message_requires_vertexprop;

 public:
  GM_HD GraphProgram() {
    order = OUT_EDGES;
    activity = ACTIVE_ONLY;
    process_message_requires_edge_value = true;
    process_message_requires_vertexprop = true;
  }
  GM_HD edge_direction getOrder() const { return order; }
  GM_HD activity_type getActivity() const { return activity; }
  GM_HD bool getProcessMessageRequiresVertexprop() const { return process_message_requires_vertexprop; }

  GM_HD virtual void reduce_function(U& v, const U& w) const { detail::null_operator("reduce_function"); }
  GM_HD virtual void process_message(const T& message, const E edge_val, const V& vertexprop, U& res) const {
    detail::null_operator("process_message");
  }
  GM_HD virtual bool send_message(const V& vertexprop, T& message) const {
    detail::null_operator("send_message");
    return true;
  }
  GM_HD virtual void apply(const U& message_out, V& vertexprop) { detail::null_operator("apply"); }
  virtual void do_every_iteration(int iteration_number) {}
  virtual ~GraphProgram() {}
};

}  // namespace GraphMat
#endif
