// GraphMatRuntime.h -- drop-in entry points of the reference runtime
// (narayanan2004/GraphMat include/GraphMatRuntime.h:51-279) over the device engine:
// graph_program_init / graph_program_clear / run_graph_program with the same arguments and
// defaults.  Compile the translation unit that defines the vertex program with nvcc
// (-x cu), annotate the program's operators with GM_HD, link libgraphmat_b200.so.
#ifndef GRAPHMAT_B200_RUNTIME_H
#define GRAPHMAT_B200_RUNTIME_H
#include <cstdio>

#include "GraphMat/Graph.h"
#include "GraphMat/GraphProgram.h"
#include "GraphMat/gm_engine.cuh"

namespace GraphMat {

const int UNTIL_CONVERGENCE = -1;

template <class T, class U, class V>
struct run_graph_program_temp_structure {
  gm_vectors* vectors;  // x (message vector, all-gathered) and y (reduced messages), in HBM
};

template <class T, class U, class V, class E>
run_graph_program_temp_structure<T, U, V> graph_program_init(const GraphProgram<T, U, V, E>& gp, const Graph<V, E>& g) {
  run_graph_program_temp_structure<T, U, V> r;
  r.vectors = nullptr;
  detail::check(gm_vectors_create(&r.vectors, g.handle, (int)sizeof(T), (int)sizeof(U)), "gm_vectors_create");
  return r;
}

template <class T, class U, class V>
void graph_program_clear(run_graph_program_temp_structure<T, U, V>& rgpts) {
  gm_vectors_destroy(rgpts.vectors);
  rgpts.vectors = nullptr;
}

// iterations = -1 ==> until convergence.  More specialised than the reference's
// GraphProgram<T,U,V,E>* signature so that the concrete program type reaches the kernels.
template <class P, class V, class E>
void run_graph_program(P* gp, Graph<V, E>& g, int iterations = 1,
                       run_graph_program_temp_structure<typename P::message_type, typename P::message_reduction_type, V>*
                           rgpts = NULL) {
  g.push();
  gm_run_stats st;
  if (gm::engine<P>::run(*gp, g.handle, iterations, rgpts ? rgpts->vectors : nullptr, &st)) {
    printf("graphmat_b200: run_graph_program failed: %s\n", gm_last_error());
    exit(1);
  }
  g.invalidate();
#ifdef __TIMING
  printf("run_graph_program: %.3f ms on device (%.3f ms in SpMSpV), %lld kernel launches\n", st.ms_total, st.ms_spmv,
         st.kernel_launches);
#endif
  printf("Completed %d iterations \n", st.iterations);
}

}  // namespace GraphMat
#endif
