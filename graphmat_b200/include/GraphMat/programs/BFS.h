// BFS.h -- the BFS vertex program of the reference app (narayanan2004/GraphMat
// src/BFS.cpp:36-108), annotated GM_HD.
#ifndef GRAPHMAT_B200_PROGRAMS_BFS_H
#define GRAPHMAT_B200_PROGRAMS_BFS_H
#include "../GraphProgram.h"

namespace gm_bfs {
typedef unsigned int depth_type;
static const depth_type kMaxDist = 0xffffffffu;  // std::numeric_limits<depth_type>::max(), src/BFS.cpp:38
}

// src/BFS.cpp:40-59; "changed" looks at depth only
class BFSD2 {
 public:
  gm_bfs::depth_type depth;
  unsigned long long int parent;
  unsigned long long int id;
  GM_HD BFSD2() : depth(gm_bfs::kMaxDist), parent(~0ull), id(~0ull) {}
  GM_HD bool operator!=(const BFSD2& p) const { return depth != p.depth; }
};

// src/BFS.cpp:61-99
class BFS2 : public GraphMat::GraphProgram<unsigned long long int, unsigned long long int, BFSD2> {
 public:
  gm_bfs::depth_type current_depth;
  static const bool gm_reorderable = true;  // a = b keeps the LAST contribution: associative, order kept
  static const bool gm_last_writer = true;  // ... so a sparse pass may scan each row from its end (gm_engine.cuh)

  GM_HD BFS2() {
    current_depth = 1;
    this->order = GraphMat::OUT_EDGES;
    this->process_message_requires_vertexprop = false;
  }
  GM_HD void reduce_function(unsigned long long int& a, const unsigned long long int& b) const { a = b; }
  GM_HD void process_message(const unsigned long long int& message, const int edge_val, const BFSD2& vertexprop,
                             unsigned long long int& res) const {
    res = message;
  }
  GM_HD bool send_message(const BFSD2& vertexprop, unsigned long long int& message) const {
    message = vertexprop.id;
    return (vertexprop.depth == current_depth - 1);  // ignored by the runtime (GraphMatRuntime.h:79-85)
  }
  GM_HD void apply(const unsigned long long int& message_out, BFSD2& vertexprop) {
    if (vertexprop.depth == gm_bfs::kMaxDist) {
      vertexprop.depth = current_depth;
      vertexprop.parent = message_out;
    }
  }
  void do_every_iteration(int iteration_number) { current_depth++; }
};
#endif
