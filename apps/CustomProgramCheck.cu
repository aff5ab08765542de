// CustomProgramCheck: a vertex program that is NOT one of the reference's apps and declares NO engine
// trait, through the drop-in surface (GraphProgram / Graph / run_graph_program).  Its reduce is a plain
// fp32 sum, i.e. order-sensitive, its activity is ACTIVE_ONLY and process_message reads the edge value and
// the destination's vertex property, so the run exercises: the serial exact fold of long rows, the
// sparse-frontier (push) path with an arbitrary reduce_function, SpMSpV3-style vertex-property access.
// The result must equal, bit for bit, a host evaluation that follows the reference's definition: per
// destination, fold the active in-neighbours' contributions left to right in ascending NATIVE column id
// (include/GMDP/singlenode/spmspv.h:55-77, include/Graph.h:111-130), apply only where a message arrived
// (include/GraphMatRuntime.h:184-226).
// usage: CustomProgramCheck [iterations]        prints "custom program ok"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>
#include <vector>

#include "GraphMatRuntime.h"
#include "common.h"

struct Heat {
  float h;
  int hits;
  GM_HD Heat() : h(0.f), hits(0) {}
  GM_HD bool operator!=(const Heat& o) const { return h != o.h; }
};

class Diffuse : public GraphMat::GraphProgram<float, float, Heat, int> {
 public:
  float keep;
  GM_HD Diffuse() {
    keep = 0.25f;
    this->order = GraphMat::OUT_EDGES;
    this->activity = GraphMat::ACTIVE_ONLY;
    this->process_message_requires_vertexprop = true;
  }
  GM_HD bool send_message(const Heat& v, float& m) const { m = v.h; return true; }
  GM_HD void process_message(const float& m, const int w, const Heat& dst, float& res) const {
    res = m * (float)w * 0.125f + (float)(dst.hits & 3) * 0.0625f;
  }
  GM_HD void reduce_function(float& a, const float& b) const { a = a + b; }
  GM_HD void apply(const float& y, Heat& v) {
    v.h = keep * v.h + (1.0f - keep) * y * 0.01f;
    v.hits++;
  }
};

static unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

static int to_native1(int v, int n, int threads) {  // include/Graph.h:111-130, nsegments = 1
  int np = threads * 16, h = n / np, v0 = v - 1;
  if (v0 >= h * np) return v;
  return v0 / np + (v0 % np) * h + 1;
}

int main(int argc, char* argv[]) {
  const int iters = argc > 1 ? atoi(argv[1]) : 6;
  const int n = 3000;
  unsigned seed = 11;
  std::set<std::pair<int, int>> uniq;
  std::vector<GraphMat::edge_t<int>> ev;
  auto add = [&](int s, int d) {
    if (uniq.insert({s, d}).second) ev.push_back(GraphMat::edge_t<int>(s, d, 1 + (int)(lcg(seed) % 7)));
  };
  for (int i = 0; i < 36000; i++) add(1 + lcg(seed) % n, 1 + lcg(seed) % n);
  for (int s = 1; s <= 1800; s++) add(s, 7);    // a long row (in-degree 1800+): the heavy-row kernels
  for (int s = 600; s <= 1000; s++) add(s, 11);
  GraphMat::edgelist_t<int> E(n, n, (int)ev.size());
  for (size_t i = 0; i < ev.size(); i++) E.edges[i] = ev[i];

  GraphMat::Graph<Heat, int> G;
  G.ReadEdgelist(E);
  std::vector<Heat> vp(n + 1);
  std::vector<char> active(n + 1, 0);
  G.setAllInactive();
  for (int v = 1; v <= n; v++) {
    vp[v].h = 1.0f + (float)(v % 13) * 0.37f;
    vp[v].hits = v % 5;
    G.setVertexproperty(v, vp[v]);
  }
  for (int v = 1; v <= n; v += 97) { active[v] = 1; G.setActive(v); }   // 31 sources: a sparse first frontier

  Diffuse prog;
  auto tmp = GraphMat::graph_program_init(prog, G);
  GraphMat::run_graph_program(&prog, G, iters, &tmp);
  GraphMat::graph_program_clear(tmp);

  // host evaluation in the reference's order
  const int threads = G.num_threads;
  std::vector<std::vector<std::pair<int, int>>> in(n + 1);  // dst -> (native(src), edge index)
  for (size_t i = 0; i < ev.size(); i++) in[ev[i].dst].push_back({to_native1(ev[i].src, n, threads), (int)i});
  for (int d = 1; d <= n; d++) std::sort(in[d].begin(), in[d].end());
  for (int it = 0; it < iters; it++) {
    std::vector<float> x(n + 1, 0.f);
    for (int v = 1; v <= n; v++) if (active[v]) prog.send_message(vp[v], x[v]);
    std::vector<char> next(n + 1, 0);
    std::vector<Heat> nvp = vp;
    for (int d = 1; d <= n; d++) {
      bool have = false;
      float acc = 0.f;
      for (auto& pr : in[d]) {
        const auto& e = ev[pr.second];
        if (!active[e.src]) continue;
        float t;
        prog.process_message(x[e.src], e.val, vp[d], t);
        if (have) prog.reduce_function(acc, t); else { acc = t; have = true; }
      }
      if (have) {
        Heat old = nvp[d];
        prog.apply(acc, nvp[d]);
        if (old != nvp[d]) next[d] = 1;
      }
    }
    vp = nvp;
    active = next;
  }
  int bad = 0;
  for (int v = 1; v <= n; v++) {
    Heat g = G.getVertexproperty(v);
    if (memcmp(&g.h, &vp[v].h, 4) != 0 || g.hits != vp[v].hits) {
      if (bad < 5) printf("vertex %d: device (%.9g, %d) host (%.9g, %d)\n", v, g.h, g.hits, vp[v].h, vp[v].hits);
      bad++;
    }
  }
  if (bad) { printf("custom program FAILED: %d vertices differ\n", bad); return 1; }
  printf("custom program ok\n");
  return 0;
}
