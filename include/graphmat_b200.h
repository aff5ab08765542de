/* graphmat_b200.h -- C ABI of the B200-native GraphMat hot path.
 *
 * The reference (narayanan2004/GraphMat) has no C ABI: its plugin surface is the
 * C++ template API (GraphProgram<T,U,V,E>, Graph<V,E>, run_graph_program), which
 * it lowers internally to C function pointers + void* (include/SPMV.h:41-59,
 * include/GraphMatRuntime.h:79-91, include/GMDP/multinode/spmspv.h:43-44).  This
 * header is the boundary a maintainer would bind instead of those internals; each
 * entry cites the reference interface it replaces (paths relative to the
 * reference tree).  Plain pointers and sizes only; all status returns are
 * 0 = ok, non-zero = error (text via gm_last_error()).  The reference's own
 * convention is printf + exit(1) (include/GraphProgram.h:73-96); the C++ mirror
 * in graphmat_b200/include keeps that behaviour on top of these status codes.
 *
 * Vertex ids at this boundary are the reference's PUBLIC ids: 1-based, before
 * Graph::vertexToNative (include/Graph.h:111-130).
 */
#ifndef GRAPHMAT_B200_H
#define GRAPHMAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gm_graph gm_graph;     /* replaces GraphMat::Graph<V,E> (include/Graph.h:58-107) */
typedef struct gm_vectors gm_vectors; /* replaces run_graph_program_temp_structure (include/GraphMatRuntime.h:53-57) */

/* include/GraphProgram.h:34,36 */
enum { GM_OUT_EDGES = 0, GM_IN_EDGES = 1, GM_ALL_EDGES = 2 };
enum { GM_ACTIVE_ONLY = 0, GM_ALL_VERTICES = 1 };
#define GM_UNTIL_CONVERGENCE (-1) /* include/GraphMatRuntime.h:51 */

/* Layout options.  ref_threads reproduces the reference's thread-count dependent
 * vertex permutation (include/Graph.h:117: npartitions = num_threads*16*nranks),
 * which fixes the per-row fold order; results equal a 1-rank reference run with
 * OMP_NUM_THREADS == ref_threads whatever (rank, world) the rows are sharded on. */
typedef struct gm_graph_opts {
  int ref_threads;      /* default 4 when 0 */
  int rank, world;      /* tile-row sharding: this process owns 1/world of the rows; default 0,1 */
  int heavy_threshold;  /* rows longer than this are stored row-contiguous and folded by a warp; 0 = default */
  int coop_threshold;   /* rows longer than this are folded by a whole thread block; 0 = default */
  int edges_on_device;  /* src/dst/val are device pointers */
  const gm_graph* order_like; /* adopt this graph's vertex placement (needed before gm_graph_share_vertexproperty) */
  int build_mask;       /* bit0: A (IN_EDGES operand), bit1: AT (OUT_EDGES operand); 0 = both (include/Graph.h:226-227) */
} gm_graph_opts;

/* One operand matrix as the kernels see it (device pointers).  Rows ("slots") are
 * sorted by decreasing length; the first n_heavy rows are stored row-contiguous
 * (CSR), the rest as 32-row sliced-ELL.  Within a row entries are in ascending
 * NATIVE column id -- the reference's fold order (include/GMDP/singlenode/spmspv.h:55-77). */
typedef struct gm_matrix_view {
  int n_slots, n_heavy, n_slices, identity;
  int n_coop;                   /* the first n_coop (longest) heavy rows get one thread block each */
  int n_slices_wide;            /* leading slices whose rows hold >= 32 entries (one warp each) */
  const int* slot_vertex;       /* slot -> local vertex (unused when identity) */
  const int* row_len;           /* n_slots */
  const long long* h_ptr;       /* n_heavy + 1 */
  const int* h_col;             /* x index per entry */
  const void* h_val;            /* edge value per entry */
  const long long* slice_ptr;   /* n_slices + 1, entry offsets (32 * width each) */
  const int* s_col;
  const void* s_val;
  long long nnz;                /* entries owned by this rank */
  int n_segs, seg_len;          /* heavy rows cut into segments of seg_len entries (associative programs) */
  const int* seg_ptr;           /* n_heavy + 1: first segment of each heavy row */
  const int* seg_row;           /* n_segs: heavy row of each segment */
  /* column-major companion for sparse frontiers (NULL until gm_graph_push_ready): the reference's
   * DCSC walks only the ACTIVE columns (include/GMDP/singlenode/spmspv.h:55-63); so does the push path */
  const long long* c_ptr;       /* n_full + 1: entries of x index c are [c_ptr[c], c_ptr[c+1]) */
  const int* c_row;             /* row slot per entry */
  const int* c_rank;            /* position of the entry in its row's fold order (ascending native column) */
  const void* c_val;            /* edge value per entry */
  int rank_bits;                /* c_rank < 2^rank_bits */
  int n_big_cols;               /* x indices whose column holds more than 2048 entries (walked by many blocks) */
  const int* big_cols;
  /* the n_long longest rows (more than long_threshold entries) hold long_entries entries, [0, long_entries) of
   * h_col: fp32-sum programs gather these into a staging buffer with the whole GPU before one block folds them */
  int n_long;
  long long long_entries;
} gm_matrix_view;

typedef struct gm_graph_view {
  int nvertices, n_local, n_local_pad, n_full, rank, world, ref_threads;
  int sizeof_V, sizeof_E;
  long long nnz;                /* whole graph */
  void* vertexproperty;         /* device, n_local_pad * sizeof_V (include/Graph.h:70) */
  unsigned int* active_bits;    /* device, n_local_pad / 32 words (include/Graph.h:71) */
  gm_matrix_view A, AT;         /* include/Graph.h:68-69 */
  int* d_flags;                 /* device scratch: [0] = "some vertex changed" */
  int* h_flags;                 /* pinned host mirror */
  void* stream;                 /* cudaStream_t all work is ordered on */
  void* aux_stream;             /* second stream: heavy rows run beside the sliced-ELL rows */
  void* ev_fork; void* ev_join; /* cudaEvent_t pair ordering the two streams */
  void* aux_stream2; void* aux_stream3; /* more of the same: warp-per-row heavy rows, narrow sliced-ELL tail */
  void* ev_join2; void* ev_join3;
  int hot_limit;                /* x indices below this are gathered with an L1-resident hint */
  gm_graph* owner;              /* the handle this view was taken from */
  int push_divisor;             /* sparse-frontier path when frontier entries * divisor <= nnz; 0 = never */
  long long push_min_nnz;       /* ... and the matrix holds at least this many entries */
} gm_graph_view;

#define GM_MAX_WORLD 16
typedef struct gm_vectors_view {
  int sizeof_T, sizeof_U;
  void* x_val; unsigned int* x_bits;  /* n_full entries: the all-gathered message vector */
  void* y_val; unsigned int* y_bits;  /* n_local_pad entries */
  void* x_alt;                        /* second message buffer of the fused apply+send pass (NULL until
                                         gm_vectors_need_alt): the pass gathers from one and writes the other */
  int n_peers;                        /* > 0: the message buffers of the other ranks are mapped into this
                                         process (gm_graph_enable_peers) and the kernels store into them */
  void* peer_x_val[GM_MAX_WORLD - 1];
  void* peer_x_alt[GM_MAX_WORLD - 1];
  unsigned int* peer_x_bits[GM_MAX_WORLD - 1];
} gm_vectors_view;

typedef struct gm_run_stats {
  int iterations;        /* iterations executed ("Completed %d iterations", GraphMatRuntime.h:277) */
  int converged;
  float ms_total;        /* device time of the whole call (CUDA events) */
  float ms_spmv;         /* device time inside the SpMSpV kernels */
  long long kernel_launches;
  long long edges_processed; /* matrix entries swept */
  long long push_passes;     /* SpMSpV passes that took the sparse-frontier (push) path */
} gm_run_stats;

const char* gm_last_error(void);
/* sizeof of the structs above as this library was compiled, in the order gm_graph_opts, gm_matrix_view, gm_graph_view,
 * gm_vectors_view, gm_run_stats, gm_push_plan: lets a foreign-language binding verify its mirror of the layouts */
int gm_abi_struct_sizes(int out[6]);
int gm_set_device(int device);
int gm_device_count(void);

/* ---- graph: replaces Graph::ReadEdgelist / ReadMTX (include/Graph.h:210-260) ---- */
/* src/dst: nnz public 1-based ids; val: nnz edge values of sizeof_E bytes (NULL = all-ones int);
 * nvertices = max(m, n) (the squaring of Graph.h:253-257 is the caller's job). */
int gm_graph_create(gm_graph** out, int nvertices, long long nnz, const int* src, const int* dst, const void* val,
                    int sizeof_E, int sizeof_V, const gm_graph_opts* opts);
/* Synthetic RMAT (a,b,c = .57,.19,.19, duplicates and self loops kept) generated on the device;
 * weight_max == 0 -> all weights 1, else uniform int in [1, weight_max]. */
int gm_graph_create_rmat(gm_graph** out, int scale, int edge_factor, unsigned long long seed, int weight_max,
                         unsigned long long weight_seed, int sizeof_V, const gm_graph_opts* opts);
/* the same generator on the host (no GPU needed): fills nnz = edge_factor << scale edges, public ids */
int gm_rmat_edges_host(int scale, int edge_factor, unsigned long long seed, int weight_max,
                       unsigned long long weight_seed, int* src, int* dst, int* val);
int gm_graph_destroy(gm_graph* g);
/* Graph::applyToAllEdges (include/Graph.h:389-402; GMDP/singlenode/applyedges.h:38-76): new value of every edge,
 * in the order and with the (src, dst) the graph was created from (host arrays).  Both operand matrices are
 * refilled in place; vertex properties, the active set and the vertex placement are kept. */
int gm_graph_set_edge_values(gm_graph* g, long long nnz, const int* src, const int* dst, const void* val);
int gm_graph_edges_changed(gm_graph* g);  /* edge values were rewritten in place on the device (device-side
                                             applyToAllEdges): drop the column-major companion, it is rebuilt on demand */
int gm_graph_view_get(const gm_graph* g, gm_graph_view* out);
int gm_graph_synchronize(const gm_graph* g);

/* include/Graph.h:263-292 */
int gm_graph_set_all_active(gm_graph* g);
int gm_graph_set_all_inactive(gm_graph* g);
int gm_graph_set_active(gm_graph* g, int v);
int gm_graph_set_inactive(gm_graph* g, int v);
int gm_graph_set_active_array(gm_graph* g, const unsigned char* flags); /* n flags, public-id order (host): the whole active set */
/* include/Graph.h:300-364; values are sizeof_V bytes each, arrays are in public-id order (index v-1) */
int gm_graph_set_all_vertexproperty(gm_graph* g, const void* value);
int gm_graph_set_vertexproperty(gm_graph* g, int v, const void* value);
int gm_graph_get_vertexproperty(const gm_graph* g, int v, void* value);
int gm_graph_set_vertexproperties(gm_graph* g, const void* values);   /* whole array, host */
int gm_graph_get_vertexproperties(const gm_graph* g, void* values);   /* whole array, host; on world>1 only owned entries are written */
int gm_graph_share_vertexproperty(gm_graph* g, gm_graph* owner);      /* g must have been created order_like = owner */
int gm_graph_vertex_owner(const gm_graph* g, int v);                  /* Graph::vertexNodeOwner, returns owning rank */
int gm_graph_out_degree_source(const gm_graph* g, int* v);            /* first public id with an out-edge (bench source) */

/* ---- x / y: graph_program_init / graph_program_clear (include/GraphMatRuntime.h:59-76) ---- */
int gm_vectors_create(gm_vectors** out, const gm_graph* g, int sizeof_T, int sizeof_U);
int gm_vectors_destroy(gm_vectors* v);
int gm_vectors_view_get(const gm_vectors* v, gm_vectors_view* out);
int gm_vectors_scratch(gm_vectors* v, long long bytes, void** out);   /* device scratch, grown on demand */
int gm_vectors_aux(gm_vectors* v, long long bytes, void** out);       /* second, persistent block, handed out filled with
                                                                         0xff bytes (the "no winner yet" table of the atomic push) */

/* ---- multi-GPU exchange (replaces the MPI sends of include/GMDP/multinode/spmspv.h:61-116 and the
 *      Allreduce of include/GraphMatRuntime.h:226).  The library calls these between send and SpMSpV /
 *      after apply when world > 1; the host language supplies them (NCCL through torch.distributed, MPI, ...). */
typedef int (*gm_allgather_fn)(void* ctx, void* buf, long long bytes_per_rank, void* stream);
typedef int (*gm_allreduce_or_fn)(void* ctx, int* host_flag);
int gm_graph_set_exchange(gm_graph* g, gm_allgather_fn allgather, gm_allreduce_or_fn allreduce_or, void* ctx);
int gm_graph_exchange_x(gm_graph* g, gm_vectors* v);     /* all-gather x values + bit words in place */
int gm_graph_exchange_x_parts(gm_graph* g, gm_vectors* v, int values, int bits); /* ... either half alone: an
                                                            ALL_VERTICES program's bit words never change after iteration 0 */
int gm_graph_allreduce_or(gm_graph* g, int* flag);       /* "some vertex changed" across ranks */
int gm_graph_exchange_buffer(gm_graph* g, void* buf, long long bytes_per_rank); /* all-gather of any buffer of
                                                            world equal slices (the second message buffer) */

/* ---- peer memory: the exchange without a library call per iteration.  Every rank maps the message
 *      buffers, the vertex-property staging area and a small flag array of every other rank (CUDA IPC
 *      between processes, plain pointers between ranks of one process) and the kernels store straight
 *      into them over NVLink: the fused apply+send pass writes each new message to all ranks as it
 *      finishes a row, ACTIVE_ONLY programs push the bit words and only the ACTIVE values of their slice
 *      (the (index,value) wire format of include/GMDP/vectors/DenseSegment.h:532-538,665-700, with the bit
 *      words as the index list), and a one-block barrier kernel (release/acquire flags in peer memory)
 *      orders the iterations and ORs the "changed" flag.  The host language only supplies a blocking
 *      all-gather of small HOST blobs for the one-time handle exchange (torch.distributed, MPI_Allgather).
 *      Collective: every rank must call gm_graph_enable_peers, and afterwards gm_vectors_create /
 *      gm_vectors_destroy / gm_run_program / gm_graph_{set,get}_vertexproperties_slice, in the same order.
 *      Returns non-zero (and leaves the graph on the callback exchange) if a peer cannot be mapped. */
typedef int (*gm_allgather_host_fn)(void* ctx, const void* mine, void* all, int bytes_per_rank);
int gm_graph_enable_peers(gm_graph* g, gm_allgather_host_fn allgather_host, void* ctx);
int gm_graph_peers_enabled(const gm_graph* g);
int gm_graph_detach_host(gm_graph* g);  /* this process is going down alone: later destroys skip the collective rendezvous */
int gm_graph_peer_barrier(gm_graph* g, int or_changed_flag); /* enqueue the barrier kernel on the graph's stream;
                                                            or_changed_flag: d_flags[0] becomes its OR over ranks */
int gm_graph_push_x(gm_graph* g, gm_vectors* v, int dense); /* store this rank's slice of x (bit words + values,
                                                            dense != 0: every value, else only where the bit is set)
                                                            into every peer's x; no barrier */
int gm_vectors_need_alt(gm_vectors* v);                  /* allocate x_alt (collective when peers are enabled) */
/* Distributed Graph::setVertexproperty / getVertexproperty for all vertices: rank r passes / receives only the
 * contiguous public ids [slice_begin(r), slice_begin(r+1)) of gm_graph_slice_begin (host memory), 1/world of the PCIe traffic of the
 * whole-array calls; the redistribution to the owning ranks runs over peer memory. */
long long gm_graph_slice_begin(const gm_graph* g, int rank);
int gm_graph_set_vertexproperties_slice(gm_graph* g, const void* slice_values);
int gm_graph_get_vertexproperties_slice(gm_graph* g, void* slice_values);

/* ---- sparse frontiers: push SpMSpV over the active columns only.  The reference's my_spmspv visits only
 *      columns whose x bit is set (include/GMDP/singlenode/spmspv.h:55-63), so its work is proportional to
 *      the frontier; the row-major kernels sweep every entry.  For ACTIVE_ONLY programs the engine counts the
 *      frontier's entries each iteration and, when they are few, expands them into (row, fold position, value)
 *      triples, sorts the triples and folds each row in the reference's order.  which: 0 = A, 1 = AT. */
typedef struct gm_push_plan {
  int n_active;               /* active columns with at least one owned entry */
  long long n_entries;        /* entries of those columns */
  int* f_col;                 /* device, n_active: active x indices */
  long long* f_off;           /* device, n_active + 1: first triple of each active column */
  unsigned long long* keys;   /* device, n_entries: (row slot << rank_bits) | fold position */
  unsigned int* order;        /* device, n_entries: index into vals */
  void* vals;                 /* device, n_entries * sizeof_U: process_message results */
  unsigned long long* keys_alt; unsigned int* order_alt; void* sort_tmp; long long sort_tmp_bytes; /* sort buffers */
  int key_bits;
} gm_push_plan;
int gm_graph_push_ready(gm_graph* g, int which);   /* build the column-major companion (idempotent) */
int gm_graph_set_push_policy(gm_graph* g, int divisor, long long min_nnz); /* defaults 16, 2^18; divisor 0 = never push */
int gm_push_count(gm_graph* g, int which, const gm_vectors* v, int* n_active, long long* n_entries); /* blocking */
int gm_push_prepare(gm_graph* g, int which, gm_vectors* v, int n_active, long long n_entries, gm_push_plan* plan);
int gm_push_sort(gm_graph* g, gm_push_plan* plan);  /* sorts (keys, order) by key; keys/order point at the result */

/* ---- the five vertex programs BASELINE.json names, compiled into the library ----
 * run == run_graph_program(&program, G, iterations, &tmp) (include/GraphMatRuntime.h:93-279).
 * `state` is the program's own state block, read and written back (do_every_iteration mutates it). */
enum {
  GM_PROG_DEGREE = 1,        /* src/PageRank.cpp:54-79   V = PR                 */
  GM_PROG_PAGERANK = 2,      /* src/PageRank.cpp:81-112  V = PR                 */
  GM_PROG_BFS = 3,           /* src/BFS.cpp:61-99        V = BFSD2              */
  GM_PROG_SSSP = 4,          /* src/SSSP.cpp:62-90       V = SSSP_vertex_type   */
  GM_PROG_DELTASTEPPING = 5, /* src/DeltaStepping.cpp:64-98 V = DeltaSteppingDS */
  GM_PROG_SGD20 = 6,         /* src/SGD.cpp:77-121 K=20  V = LatentVector<20>   */
  GM_PROG_RMSE20 = 7,        /* src/SGD.cpp:123-156 K=20                        */
  GM_PROG_SGD32 = 8,
  GM_PROG_RMSE32 = 9,
  GM_PROG_SGD4 = 10,
  GM_PROG_RMSE4 = 11,
  /* the remaining POD vertex programs */
  GM_PROG_DEGREE_DPR = 12,     /* src/IncrementalPageRank.cpp:53-78  V = dPR              */
  GM_PROG_DELTAPAGERANK = 13,  /* src/IncrementalPageRank.cpp:80-123 V = dPR              */
  GM_PROG_INDEGREE = 14,       /* src/TopologicalSort.cpp:60-87      V = Vertex_type      */
  GM_PROG_TOPSORT = 15,        /* src/TopologicalSort.cpp:90-130     V = Vertex_type      */
  GM_PROG_LDAINIT20 = 16,      /* src/LDA.cpp:69-111  K = 20         V = LatentVector<20> (the LDA one) */
  GM_PROG_LDA20 = 17,          /* src/LDA.cpp:127-196: global_N is recomputed on the device before the run and in
                                  every do_every_iteration, and handed back in the state */
  GM_PROG_LDALL20 = 18         /* src/LDA.cpp:198-250 */
};
typedef struct gm_lda_state { double alpha, eta, vocab_size; double global_N[20]; } gm_lda_state;   /* src/LDA.cpp:129-132 */
typedef struct gm_ldall_state { double N_k[20]; double eta; int nterms; } gm_ldall_state;          /* :201-203, N_k unsmoothed */
typedef struct gm_deltapagerank_state { double alpha; int iter; } gm_deltapagerank_state; /* src/IncrementalPageRank.cpp:82-83 */
typedef struct gm_topsort_state { unsigned int current_topsort_order; } gm_topsort_state;   /* src/TopologicalSort.cpp:93 */
typedef struct gm_pagerank_state { float alpha; } gm_pagerank_state;                  /* src/PageRank.cpp:84 */
typedef struct gm_bfs_state { unsigned int current_depth; } gm_bfs_state;             /* src/BFS.cpp:64 */
typedef struct gm_deltastepping_state { int delta, bid; } gm_deltastepping_state;     /* src/DeltaStepping.cpp:67-68 */
typedef struct gm_sgd_state { double lambda, step; } gm_sgd_state;                    /* src/SGD.cpp:79-80 */

int gm_program_sizes(int program, int* sizeof_T, int* sizeof_U, int* sizeof_V, int* sizeof_state);
int gm_run_program(gm_graph* g, int program, void* state, int iterations, gm_vectors* tmp, gm_run_stats* stats);
/* the three steps of one iteration, separately (tests; include/GraphMatRuntime.h:145,160-176,184-226) */
int gm_step_send(gm_graph* g, int program, const void* state, gm_vectors* tmp);
int gm_step_spmspv(gm_graph* g, int program, const void* state, gm_vectors* tmp);
int gm_step_apply(gm_graph* g, int program, void* state, gm_vectors* tmp, int* changed);

/* Graph::applyReduceAllVertices (include/Graph.h:377-381) for the app drivers' three map functions */
/* GM_REDUCE_REACHABLE also serves src/TopologicalSort.cpp:132-138 ("unreachable" = nvertices - reachable) */
enum { GM_REDUCE_REACHABLE = 1,      /* src/BFS.cpp:101-108 etc.: count of vertices whose first uint field < UINT_MAX */
       GM_REDUCE_BUCKET_NOT_EMPTY = 2, /* src/DeltaStepping.cpp:109-111, param = bid */
       GM_REDUCE_SQERR = 3 };        /* src/SGD.cpp:158-161: sum of the trailing double (sqerr); also src/LDA.cpp:268-271,
                                        338-339 (token_loglik is LatentVector's trailing double) */
int gm_graph_reduce(const gm_graph* g, int what, int param, double* result);

/* ---- test hooks: the bit-exact parallel fp32 fold used for long PageRank rows (gm_fadd32.cuh).
 * Both return the value of the serial left fold a[0] + a[1] + ... in fp32; status 2 = empty. */
int gm_debug_fold_f32_host(const float* a, long long n, float* out);
int gm_debug_fold_f32_device(const float* a, long long n, int warps, int offset, float* out);

#ifdef __cplusplus
}
#endif
#endif
