// graph_converter: the edge-list conversion tool of the reference (src/graph_converter.cpp:38-338),
// same options, same defaults, same order of transformations, one process.
// usage: graph_converter [options] <input file prefix> <output file prefix>
// Formats 0 and 1 (binary / text mtx) are byte-compatible with the reference's; format 2 is this
// repository's GraphMat-binary snapshot (Graph.h WriteGraphMatBin: re-specified without Boost, see
// DESIGN.md), which builds the device graph and therefore needs a GPU.  Formats 0 and 1 run on the host alone.
#include <getopt.h>

#include "GraphMatRuntime.h"
#include "common.h"

struct converter_options {
  int selfloops = 0;
  int duplicatededges = 0;
  int uppertriangular = 0;
  int bidirectional = 0;
  int inputformat = 1;
  int outputformat = 0;
  int inputheader = 1;
  int outputheader = 1;
  int inputedgeweights = 1;
  int outputedgeweights = 1;
  int edgeweighttype = 0;
  int nvertices = 0;
  int random_range = 128;
  int nsplits = 1;
  int randomizeID = 0;
};

// graph_converter.cpp:38-66
static bool options_valid(const converter_options& o) {
  bool ok = true;
  if (o.selfloops != 0 && o.selfloops != 1) { printf("selfloops must be 0 or 1 \n"); ok = false; }
  if (o.uppertriangular == 1 && o.bidirectional == 1) {
    printf("Cannot be both uppertriangular and bidirectional\n");
    ok = false;
  }
  if (o.inputedgeweights == 0 && o.outputedgeweights == 1) {
    printf("No input edge weights and want output edge weights\n");
    ok = false;
  }
  if (o.nsplits < 0) { printf("Cannot split into negative number of pieces\n"); ok = false; }
  if (o.nsplits != 1) {
    printf("Split functionality is deprecated. Call with \"mpirun -np <nsplits> ... \" instead\n");
    ok = false;
  }
  if (!ok) printf("Error in validating options\n");
  return ok;
}

static void print_options(const converter_options& o) {
  printf("Options -- \n");
  printf("Selfloops = %d \n", o.selfloops);
  printf("Duplicated edges = %d \n", o.duplicatededges);
  printf("Uppertriangular = %d \n", o.uppertriangular);
  printf("Bidirectional = %d \n", o.bidirectional);
  printf("Input format = %d \n", o.inputformat);
  printf("Output format = %d \n", o.outputformat);
  printf("Input header = %d \n", o.inputheader);
  printf("Output header = %d \n", o.outputheader);
  printf("Input edge weights = %d \n", o.inputedgeweights);
  printf("Output edge weights = %d \n", o.outputedgeweights);
  printf("Edge weight type = %d \n", o.edgeweighttype);
  printf("Range of random edge weights = %d \n", o.random_range);
  printf("Number of vertices = %d \n", o.nvertices);
  printf("Randomize vertex IDs = %d \n", o.randomizeID);
}

static void print_help(const char* argv0) {
  printf("Usage: %s [options] <input mtx file prefix> <output mtx file prefix> \n", argv0);
  printf("Options:\n"
         "\t--help Print help message and exit.\n"
         "\t--selfloops\n\t\t0: Remove all self loops (default)\n\t\t1: Retain self loops\n"
         "\t--duplicatededges\n\t\t0: Remove all duplicated edges (default)\n\t\t1: Retain duplicated edges\n"
         "\t--uppertriangular\tAll edges (u,v), leave edge unchanged if u <= v, and swap u & v if u > v\n"
         "\t--bidirectional\tFor all edges (u,v), add (v,u)\n"
         "\t--inputformat\n\t\t0: Binary mtx input\n\t\t1: Text mtx input (default)\n"
         "\t\t2: GraphMat-binary snapshot of this library (needs a GPU)\n"
         "\t--outputformat\n\t\t0: Binary mtx output (default)\n\t\t1: Text mtx output\n"
         "\t\t2: GraphMat-binary snapshot of this library (needs a GPU)\n"
         "\t--inputheader\n\t\t0: no header (can provide nvertices through --nvertices or we take the max as nvertices)\n"
         "\t\t1: (n,n,nnz) (default)\n"
         "\t--outputheader\n\t\t0: no header\n\t\t1: (n,n,nnz) (default)\n"
         "\t--inputedgeweights\n\t\t0: no weights\n\t\t1: weights present (default)\n"
         "\t--outputedgeweights\n\t\t0: no weights\n\t\t1: weights present (default)\n\t\t2: create unit weights\n"
         "\t\t3: create random weights in range [1,r) (specify r with --r option, default r=128) \n"
         "\t--edgeweighttype\n\t\t0: int (default)\n\t\t1: double\n\t\t2: float\n"
         "\t--r [number] range of random edge weights created (use only with \"--outputedgeweights 3\")\n"
         "\t--nvertices [number] (use only with \"--inputheader 0\")\n"
         "\t--randomizeID\tUsing this flag would randomize the vertex IDs from the input file\n");
}

static int no_snapshot_for_type() {
  printf("graphmat_b200: the device engine holds 4-byte edge values; use --edgeweighttype 0 or 2 with format 2\n");
  return 1;
}

// graph_converter.cpp:162-222: read, then weights -> self loops -> bidirectional -> dag -> duplicates ->
// relabel, then write.
template <typename T>
static int process_graph(const char* in, const char* out, const converter_options& o) {
  GraphMat::edgelist_t<T> edgelist;
  if (o.inputformat == 0 || o.inputformat == 1) {
    GraphMat::load_edgelist<T>(in, &edgelist, o.inputformat == 0, o.inputheader == 1, o.inputedgeweights == 1);
    int nv = std::max(edgelist.m, edgelist.n);  // the tool works on square matrices
    if (o.nvertices > 0) {
      if (o.nvertices < nv) {
        printf("--nvertices %d is smaller than the largest vertex id %d\n", o.nvertices, nv);
        return 1;
      }
      nv = o.nvertices;
    }
    edgelist.m = edgelist.n = nv;
  } else if (o.inputformat == 2) {
    if constexpr (sizeof(T) == 4) {
      GraphMat::Graph<int, T> G;
      G.ReadGraphMatBin(in);
      G.getEdgelist(edgelist);
    } else {
      return no_snapshot_for_type();
    }
  } else {
    printf("Invalid input format: %d\n", o.inputformat);
    return 1;
  }
  if (o.outputedgeweights == 3) GraphMat::random_edge_weights(&edgelist, o.random_range);
  if (o.selfloops == 0) GraphMat::remove_selfedges(&edgelist);
  if (o.bidirectional == 1) GraphMat::create_bidirectional_edges(&edgelist);
  if (o.uppertriangular == 1) GraphMat::convert_to_dag(&edgelist);
  if (o.duplicatededges == 0) GraphMat::remove_duplicate_edges(&edgelist);
  if (o.randomizeID == 1) GraphMat::randomize_edgelist_square(&edgelist);

  if (o.outputformat == 0 || o.outputformat == 1) {
    // option 2 ("unit weights") writes the weights as they stand, as the reference does (graph_converter.cpp:207)
    GraphMat::write_edgelist<T>(out, edgelist, o.outputformat == 0, o.outputheader == 1, o.outputedgeweights != 0);
  } else if (o.outputformat == 2) {
    if constexpr (sizeof(T) == 4) {
      GraphMat::Graph<int, T> G;
      G.ReadEdgelist(edgelist);
      G.WriteGraphMatBin(out);
    } else {
      return no_snapshot_for_type();
    }
  } else {
    printf("Invalid output format: %d\n", o.outputformat);
    return 1;
  }
  edgelist.clear();
  return 0;
}

int main(int argc, char* argv[]) {
  converter_options o;
  const struct option long_options[] = {{"uppertriangular", no_argument, &o.uppertriangular, 1},
                                        {"bidirectional", no_argument, &o.bidirectional, 1},
                                        {"randomizeID", no_argument, &o.randomizeID, 1},
                                        {"selfloops", required_argument, 0, 's'},
                                        {"duplicatededges", required_argument, 0, 'd'},
                                        {"inputformat", required_argument, 0, 'i'},
                                        {"outputformat", required_argument, 0, 'o'},
                                        {"inputheader", required_argument, 0, 'n'},
                                        {"outputheader", required_argument, 0, 'u'},
                                        {"inputedgeweights", required_argument, 0, 'e'},
                                        {"outputedgeweights", required_argument, 0, 'w'},
                                        {"edgeweighttype", required_argument, 0, 't'},
                                        {"r", required_argument, 0, 'r'},
                                        {"nvertices", required_argument, 0, 'v'},
                                        {"split", required_argument, 0, 'p'},
                                        {"help", no_argument, 0, 'h'},
                                        {0, 0, 0, 0}};
  for (;;) {
    int c = getopt_long(argc, argv, "hs:d:i:o:n:u:e:w:t:v:r:p:", long_options, nullptr);
    if (c == -1) break;
    int* target = nullptr;
    switch (c) {
      case 0: break;  // a flag option, already stored
      case 'h': print_help(argv[0]); return 0;
      case 's': target = &o.selfloops; break;
      case 'd': target = &o.duplicatededges; break;
      case 'i': target = &o.inputformat; break;
      case 'o': target = &o.outputformat; break;
      case 'n': target = &o.inputheader; break;
      case 'u': target = &o.outputheader; break;
      case 'e': target = &o.inputedgeweights; break;
      case 'w': target = &o.outputedgeweights; break;
      case 't': target = &o.edgeweighttype; break;
      case 'v': target = &o.nvertices; break;
      case 'r': target = &o.random_range; break;
      case 'p': target = &o.nsplits; break;
      default: break;  // unknown option: getopt has printed the complaint
    }
    if (target) sscanf(optarg, "%i", target);
  }
  if (optind != argc - 2) {
    print_help(argv[0]);
    return 0;
  }
  if (!options_valid(o)) return 1;
  print_options(o);
  const char* in = argv[optind];
  const char* out = argv[optind + 1];
  switch (o.edgeweighttype) {
    case 0: return process_graph<unsigned int>(in, out, o);
    case 1: return process_graph<double>(in, out, o);
    case 2: return process_graph<float>(in, out, o);
    default: printf("Invalid edge type: %d\n", o.edgeweighttype); return 1;
  }
}
