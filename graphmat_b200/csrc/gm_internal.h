// gm_internal.h -- private state behind the opaque handles of include/graphmat_b200.h.
#ifndef GM_INTERNAL_H
#define GM_INTERNAL_H
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "graphmat_b200.h"

#define GM_DEFAULT_HEAVY_THRESHOLD 4096
#define GM_DEFAULT_COOP_THRESHOLD 16384
#define GM_SEG_LEN 2048
#define GM_DEFAULT_LONG_THRESHOLD 32768
#define GM_PUSH_BIG_COL 2048  /* == GM_PUSH_BIG of gm_engine.cuh */

void gm_set_error(const std::string& s);

// a buffer every rank allocates with the same size, with the other ranks' copies mapped here
struct gm_sym {
  void* local = nullptr;
  void* peer[GM_MAX_WORLD] = {};  // indexed by rank; peer[my rank] == local
  bool opened[GM_MAX_WORLD] = {};  // mapped through cudaIpcOpenMemHandle (to be closed)
  size_t bytes = 0;
};

// one operand matrix (device memory owned here); see gm_matrix_view for the meaning
struct gm_matrix {
  int n_slots = 0, n_heavy = 0, n_slices = 0, identity = 0, n_coop = 0, n_slices_wide = 0;
  int* slot_vertex = nullptr;
  int* row_len = nullptr;
  long long* h_ptr = nullptr;
  int* h_col = nullptr;
  void* h_val = nullptr;
  long long* slice_ptr = nullptr;
  int* s_col = nullptr;
  void* s_val = nullptr;
  long long nnz = 0;
  long long s_entries = 0;  // padded sliced-ELL entries
  int n_segs = 0, seg_len = 0;
  int* seg_ptr = nullptr;
  int* seg_row = nullptr;
  // column-major companion for sparse frontiers (built on first use, gm_graph_push_ready)
  long long* c_ptr = nullptr;  // n_full + 1
  int* c_row = nullptr;        // row slot of each entry
  int* c_rank = nullptr;       // position of the entry in its row's fold order
  void* c_val = nullptr;       // edge value
  int rank_bits = 0;           // bits needed for c_rank
  int n_big_cols = 0;          // columns above GM_PUSH_BIG_COL entries
  int* big_cols = nullptr;
  int n_long = 0;              // rows longer than gm_graph::long_threshold (a prefix: rows are stored longest first)
  long long long_entries = 0;  // h_ptr[n_long]
  bool push_built = false;
};

struct gm_graph {
  int n = 0, n_local = 0, n_pad = 0, n_full = 0;
  int rank = 0, world = 1, ref_threads = 4, heavy_threshold = GM_DEFAULT_HEAVY_THRESHOLD, coop_threshold = GM_DEFAULT_COOP_THRESHOLD;
  int sizeof_V = 0, sizeof_E = 4;
  long long nnz = 0;
  int first_source = 0;
  bool heavy_auto = true;
  int* d_xidx = nullptr;    // native id -> x index (owner * n_pad + local)
  std::vector<int> h_xidx;  // lazily mirrored for single-vertex accessors
  void* vp = nullptr;
  bool vp_owner = true;
  unsigned* active = nullptr;
  gm_matrix A, AT;
  int* d_flags = nullptr;
  int* h_flags = nullptr;
  void* staging = nullptr;
  size_t staging_bytes = 0;
  cudaStream_t stream = nullptr, aux_stream = nullptr, aux_stream2 = nullptr, aux_stream3 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr, ev_join3 = nullptr;
  int hot_limit = -1;
  int long_threshold = GM_DEFAULT_LONG_THRESHOLD;
  void* push_scratch = nullptr;  // triples + sort buffers of the sparse-frontier path, grown geometrically
  size_t push_scratch_bytes = 0;
  int push_divisor = 16;
  long long push_min_nnz = 1ll << 18;
  gm_allgather_fn allgather = nullptr;
  gm_allreduce_or_fn allreduce_or = nullptr;
  void* xctx = nullptr;
  // peer memory (gm_peer.cu): the other ranks' buffers mapped into this process
  bool peers_on = false;
  bool peers_same_process = false;  // some peer is a rank of this process (test harness): barriers meet on the host
  gm_allgather_host_fn host_gather = nullptr;
  void* host_ctx = nullptr;
  gm_sym sync;              // 2 * GM_MAX_WORLD 64-bit words per rank: barrier flags (parity-alternated)
  unsigned barrier_round = 0;
  gm_sym sym_staging;       // public-order staging area for the *_slice accessors (peers only)
  std::vector<void*> retired;  // outgrown scratch blocks, freed with the graph
};

struct gm_vectors {
  int sizeof_T = 0, sizeof_U = 0, n_full = 0, n_pad = 0;
  gm_graph* g = nullptr;    // the graph these vectors were created for (stream, peers)
  void* x_val = nullptr;
  unsigned* x_bits = nullptr;
  void* y_val = nullptr;
  unsigned* y_bits = nullptr;
  void* x_alt = nullptr;
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* aux = nullptr;      // persistent second block (gm_vectors_aux)
  size_t aux_bytes = 0;
  std::vector<void*> retired;  // outgrown scratch blocks, freed with the vectors
  bool sym = false;         // x_val / x_bits / x_alt are symmetric allocations mapped on every rank
  gm_sym s_val, s_bits, s_alt;
};

// gm_peer.cu
int gm_sym_alloc(gm_graph* g, size_t bytes, gm_sym* out);  // collective
int gm_sym_free(gm_graph* g, gm_sym* s);                   // collective
int gm_peer_copy_slice(gm_graph* g, gm_sym* s, size_t offset, size_t bytes);  // my [offset, +bytes) -> every peer, same offset
#endif
