"""ctypes binding of include/graphmat_b200.h (the stub INTEGRATION.md shows).

No compute happens here and nothing falls back to the CPU: every call goes to
libgraphmat_b200.so, and a missing library or a non-zero status raises.
"""
import ctypes as C
import os
import subprocess
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgraphmat_b200.so")
_lib = None

# program ids / enums (include/graphmat_b200.h)
OUT_EDGES, IN_EDGES, ALL_EDGES = 0, 1, 2
UNTIL_CONVERGENCE = -1
PROG_DEGREE, PROG_PAGERANK, PROG_BFS, PROG_SSSP, PROG_DELTASTEPPING = 1, 2, 3, 4, 5
PROG_SGD20, PROG_RMSE20, PROG_SGD32, PROG_RMSE32, PROG_SGD4, PROG_RMSE4 = 6, 7, 8, 9, 10, 11
PROG_DEGREE_DPR, PROG_DELTAPAGERANK, PROG_INDEGREE, PROG_TOPSORT = 12, 13, 14, 15
PROG_LDAINIT20, PROG_LDA20, PROG_LDALL20 = 16, 17, 18
REDUCE_REACHABLE, REDUCE_BUCKET_NOT_EMPTY, REDUCE_SQERR = 1, 2, 3
SGD_PROGRAMS = {20: (PROG_SGD20, PROG_RMSE20), 32: (PROG_SGD32, PROG_RMSE32), 4: (PROG_SGD4, PROG_RMSE4)}

# vertex property layouts of the five apps (programs/*.h)
PR_DTYPE = np.dtype([("pagerank", np.float32), ("degree", np.int32)])
BFS_DTYPE = np.dtype([("depth", np.uint32), ("parent", np.uint64), ("id", np.uint64)], align=True)
SSSP_DTYPE = np.dtype([("distance", np.uint32)])
DS_DTYPE = np.dtype([("distance", np.uint32), ("bucket", np.int32)])
DPR_DTYPE = np.dtype([("delta", np.float64), ("pagerank", np.float64), ("degree", np.int32)], align=True)
TOPSORT_DTYPE = np.dtype([("topsort_order", np.uint32), ("in_degree", np.int32)])
LDA_DTYPE = np.dtype({"names": ["N", "type", "token_loglik"], "formats": [(np.float64, (20,)), "S1", np.float64],
                      "offsets": [0, 160, 168], "itemsize": 176})  # LatentVector<20> of src/LDA.cpp:36-47


def latent_dtype(K):
    return np.dtype([("lv", np.float64, (K,)), ("sqerr", np.float64)])


class GraphOpts(C.Structure):
    _fields_ = [("ref_threads", C.c_int), ("rank", C.c_int), ("world", C.c_int), ("heavy_threshold", C.c_int),
                ("coop_threshold", C.c_int), ("edges_on_device", C.c_int), ("order_like", C.c_void_p), ("build_mask", C.c_int)]


class RunStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("converged", C.c_int), ("ms_total", C.c_float), ("ms_spmv", C.c_float),
                ("kernel_launches", C.c_longlong), ("edges_processed", C.c_longlong),
                ("push_passes", C.c_longlong)]


class MatrixView(C.Structure):
    _fields_ = [("n_slots", C.c_int), ("n_heavy", C.c_int), ("n_slices", C.c_int), ("identity", C.c_int),
                ("n_coop", C.c_int), ("n_slices_wide", C.c_int), ("slot_vertex", C.c_void_p), ("row_len", C.c_void_p), ("h_ptr", C.c_void_p), ("h_col", C.c_void_p),
                ("h_val", C.c_void_p), ("slice_ptr", C.c_void_p), ("s_col", C.c_void_p), ("s_val", C.c_void_p),
                ("nnz", C.c_longlong), ("n_segs", C.c_int), ("seg_len", C.c_int), ("seg_ptr", C.c_void_p),
                ("seg_row", C.c_void_p), ("c_ptr", C.c_void_p), ("c_row", C.c_void_p), ("c_rank", C.c_void_p),
                ("c_val", C.c_void_p), ("rank_bits", C.c_int), ("n_big_cols", C.c_int), ("big_cols", C.c_void_p),
                ("n_long", C.c_int), ("long_entries", C.c_longlong)]


class GraphView(C.Structure):
    _fields_ = [("nvertices", C.c_int), ("n_local", C.c_int), ("n_local_pad", C.c_int), ("n_full", C.c_int),
                ("rank", C.c_int), ("world", C.c_int), ("ref_threads", C.c_int), ("sizeof_V", C.c_int),
                ("sizeof_E", C.c_int), ("nnz", C.c_longlong), ("vertexproperty", C.c_void_p),
                ("active_bits", C.c_void_p), ("A", MatrixView), ("AT", MatrixView), ("d_flags", C.c_void_p),
                ("h_flags", C.c_void_p), ("stream", C.c_void_p), ("aux_stream", C.c_void_p), ("ev_fork", C.c_void_p),
                ("ev_join", C.c_void_p), ("aux_stream2", C.c_void_p), ("aux_stream3", C.c_void_p), ("ev_join2", C.c_void_p),
                ("ev_join3", C.c_void_p), ("hot_limit", C.c_int), ("owner", C.c_void_p),
                ("push_divisor", C.c_int), ("push_min_nnz", C.c_longlong)]


MAX_WORLD = 16


class VectorsView(C.Structure):
    _fields_ = [("sizeof_T", C.c_int), ("sizeof_U", C.c_int), ("x_val", C.c_void_p), ("x_bits", C.c_void_p),
                ("y_val", C.c_void_p), ("y_bits", C.c_void_p), ("x_alt", C.c_void_p), ("n_peers", C.c_int),
                ("peer_x_val", C.c_void_p * (MAX_WORLD - 1)), ("peer_x_alt", C.c_void_p * (MAX_WORLD - 1)),
                ("peer_x_bits", C.c_void_p * (MAX_WORLD - 1))]


class PushPlan(C.Structure):
    _fields_ = [("n_active", C.c_int), ("n_entries", C.c_longlong), ("f_col", C.c_void_p), ("f_off", C.c_void_p),
                ("keys", C.c_void_p), ("order", C.c_void_p), ("vals", C.c_void_p), ("keys_alt", C.c_void_p),
                ("order_alt", C.c_void_p), ("sort_tmp", C.c_void_p), ("sort_tmp_bytes", C.c_longlong),
                ("key_bits", C.c_int)]


class PageRankState(C.Structure):
    _fields_ = [("alpha", C.c_float)]


class BFSState(C.Structure):
    _fields_ = [("current_depth", C.c_uint)]


class DeltaSteppingState(C.Structure):
    _fields_ = [("delta", C.c_int), ("bid", C.c_int)]


class DeltaPageRankState(C.Structure):
    _fields_ = [("alpha", C.c_double), ("iter", C.c_int)]


class TopSortState(C.Structure):
    _fields_ = [("current_topsort_order", C.c_uint)]


class LDAState(C.Structure):
    _fields_ = [("alpha", C.c_double), ("eta", C.c_double), ("vocab_size", C.c_double), ("global_N", C.c_double * 20)]


class LDALLState(C.Structure):
    _fields_ = [("N_k", C.c_double * 20), ("eta", C.c_double), ("nterms", C.c_int)]


class SGDState(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("step", C.c_double)]


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p)
ALLREDUCE_OR_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int))
ALLGATHER_HOST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int)

# every symbol include/graphmat_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "gm_last_error", "gm_set_device", "gm_device_count", "gm_graph_create", "gm_graph_create_rmat",
    "gm_rmat_edges_host", "gm_graph_destroy", "gm_graph_view_get", "gm_graph_synchronize", "gm_graph_set_all_active",
    "gm_graph_set_all_inactive", "gm_graph_set_active", "gm_graph_set_inactive", "gm_graph_set_all_vertexproperty",
    "gm_graph_set_vertexproperty", "gm_graph_get_vertexproperty", "gm_graph_set_vertexproperties",
    "gm_graph_get_vertexproperties", "gm_graph_share_vertexproperty", "gm_graph_vertex_owner",
    "gm_graph_out_degree_source", "gm_vectors_create", "gm_vectors_destroy", "gm_vectors_view_get", "gm_vectors_scratch",
    "gm_graph_set_exchange", "gm_graph_exchange_x", "gm_graph_allreduce_or", "gm_program_sizes", "gm_run_program",
    "gm_step_send", "gm_step_spmspv", "gm_step_apply", "gm_graph_reduce", "gm_debug_fold_f32_host",
    "gm_debug_fold_f32_device", "gm_graph_push_ready", "gm_graph_set_push_policy", "gm_push_count", "gm_push_prepare",
    "gm_push_sort", "gm_graph_set_edge_values", "gm_graph_exchange_x_parts", "gm_abi_struct_sizes",
    "gm_graph_exchange_buffer", "gm_graph_enable_peers", "gm_graph_peers_enabled", "gm_graph_peer_barrier",
    "gm_graph_push_x", "gm_vectors_need_alt", "gm_graph_slice_begin", "gm_graph_set_vertexproperties_slice",
    "gm_graph_get_vertexproperties_slice", "gm_vectors_aux", "gm_graph_set_active_array", "gm_graph_edges_changed", "gm_graph_detach_host",
]


def build_library(force=False):
    """Compile csrc/ into libgraphmat_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j2"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError("libgraphmat_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                               "there is no CPU fallback")
        L = C.CDLL(_LIB_PATH)
        L.gm_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, (lib().gm_last_error() or b"").decode()))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def rmat_edges(scale, edge_factor=16, seed=1, weight_max=0, weight_seed=2):
    """Host twin of the device RMAT generator (same edges, public 1-based ids)."""
    nnz = edge_factor << scale
    src = np.empty(nnz, np.int32)
    dst = np.empty(nnz, np.int32)
    val = np.empty(nnz, np.int32)
    lib().gm_rmat_edges_host(C.c_int(scale), C.c_int(edge_factor), C.c_ulonglong(seed), C.c_int(weight_max),
                             C.c_ulonglong(weight_seed), _p(src), _p(dst), _p(val))
    return 1 << scale, src, dst, val


class Graph:
    """GraphMat::Graph<V,E> behind the C ABI (E = int)."""

    def __init__(self, handle, vdtype):
        self.h = handle
        self.vdtype = np.dtype(vdtype)
        self._keep = []
        self._vectors = weakref.WeakSet()  # gm_vectors point back at their graph: they go first

    @staticmethod
    def _opts(threads, rank, world, heavy_threshold, order_like, build_mask, on_device=False, coop_threshold=0):
        o = GraphOpts()
        o.ref_threads, o.rank, o.world, o.heavy_threshold = threads, rank, world, heavy_threshold
        o.coop_threshold = coop_threshold
        o.edges_on_device = 1 if on_device else 0
        o.order_like = order_like.h if order_like is not None else None
        o.build_mask = build_mask
        return o

    @classmethod
    def from_edges(cls, n, src, dst, val, vdtype, threads=4, rank=0, world=1, heavy_threshold=0, order_like=None,
                   build_mask=0, coop_threshold=0):
        src, dst = _i32(src), _i32(dst)
        val = _i32(val) if val is not None else None
        h = C.c_void_p()
        o = cls._opts(threads, rank, world, heavy_threshold, order_like, build_mask, coop_threshold=coop_threshold)
        _check(lib().gm_graph_create(C.byref(h), C.c_int(n), C.c_longlong(len(src)), _p(src), _p(dst),
                                     _p(val) if val is not None else None, C.c_int(4),
                                     C.c_int(np.dtype(vdtype).itemsize), C.byref(o)), "gm_graph_create")
        return cls(h, vdtype)

    @classmethod
    def from_device_edges(cls, n, nnz, src_ptr, dst_ptr, val_ptr, vdtype, threads=4, rank=0, world=1, heavy_threshold=0,
                          order_like=None, build_mask=0):
        """edge list already in device memory (int32 arrays, public 1-based ids); the arrays are only read"""
        h = C.c_void_p()
        o = cls._opts(threads, rank, world, heavy_threshold, order_like, build_mask, on_device=True)
        _check(lib().gm_graph_create(C.byref(h), C.c_int(n), C.c_longlong(nnz), C.c_void_p(src_ptr), C.c_void_p(dst_ptr),
                                     C.c_void_p(val_ptr) if val_ptr else None, C.c_int(4),
                                     C.c_int(np.dtype(vdtype).itemsize), C.byref(o)), "gm_graph_create")
        return cls(h, vdtype)

    @classmethod
    def rmat(cls, scale, vdtype, edge_factor=16, seed=1, weight_max=0, weight_seed=2, threads=4, rank=0, world=1,
             heavy_threshold=0, build_mask=0, coop_threshold=0):
        h = C.c_void_p()
        o = cls._opts(threads, rank, world, heavy_threshold, None, build_mask, coop_threshold=coop_threshold)
        _check(lib().gm_graph_create_rmat(C.byref(h), C.c_int(scale), C.c_int(edge_factor), C.c_ulonglong(seed),
                                          C.c_int(weight_max), C.c_ulonglong(weight_seed),
                                          C.c_int(np.dtype(vdtype).itemsize), C.byref(o)), "gm_graph_create_rmat")
        return cls(h, vdtype)

    def close(self):
        if self.h:
            for v in list(self._vectors):
                v.close()
            lib().gm_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.h:  # garbage collection, not an orderly close(): never wait for the other ranks here
                lib().gm_graph_detach_host(self.h)
            self.close()
        except Exception:
            pass

    def view(self):
        v = GraphView()
        _check(lib().gm_graph_view_get(self.h, C.byref(v)), "gm_graph_view_get")
        return v

    @property
    def nvertices(self):
        return self.view().nvertices

    @property
    def nnz(self):
        return self.view().nnz

    def set_all_active(self):
        _check(lib().gm_graph_set_all_active(self.h), "gm_graph_set_all_active")

    def set_all_inactive(self):
        _check(lib().gm_graph_set_all_inactive(self.h), "gm_graph_set_all_inactive")

    def set_active(self, v):
        _check(lib().gm_graph_set_active(self.h, C.c_int(v)), "gm_graph_set_active")

    def set_active_many(self, ids):
        """the active set = exactly these public ids"""
        flags = np.zeros(self.nvertices, np.uint8)
        flags[np.asarray(ids, dtype=np.int64) - 1] = 1
        _check(lib().gm_graph_set_active_array(self.h, _p(flags)), "gm_graph_set_active_array")

    def set_inactive(self, v):
        _check(lib().gm_graph_set_inactive(self.h, C.c_int(v)), "gm_graph_set_inactive")

    def set_all_vertexproperty(self, value):
        a = np.zeros(1, self.vdtype)
        a[0] = value
        _check(lib().gm_graph_set_all_vertexproperty(self.h, _p(a)), "gm_graph_set_all_vertexproperty")

    def set_vertexproperty(self, v, value):
        a = np.zeros(1, self.vdtype)
        a[0] = value
        _check(lib().gm_graph_set_vertexproperty(self.h, C.c_int(v), _p(a)), "gm_graph_set_vertexproperty")

    def get_vertexproperty(self, v):
        a = np.zeros(1, self.vdtype)
        _check(lib().gm_graph_get_vertexproperty(self.h, C.c_int(v), _p(a)), "gm_graph_get_vertexproperty")
        return a[0]

    def set_vertexproperties(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.vdtype)
        assert len(arr) == self.nvertices
        _check(lib().gm_graph_set_vertexproperties(self.h, _p(arr)), "gm_graph_set_vertexproperties")

    def set_vertexproperties_ptr(self, ptr):
        _check(lib().gm_graph_set_vertexproperties(self.h, C.c_void_p(ptr)), "gm_graph_set_vertexproperties")

    def get_vertexproperties(self, out=None):
        if out is None:
            out = np.zeros(self.nvertices, self.vdtype)
        _check(lib().gm_graph_get_vertexproperties(self.h, _p(out)), "gm_graph_get_vertexproperties")
        return out

    def get_vertexproperties_ptr(self, ptr):
        _check(lib().gm_graph_get_vertexproperties(self.h, C.c_void_p(ptr)), "gm_graph_get_vertexproperties")

    def share_vertexproperty(self, owner):
        _check(lib().gm_graph_share_vertexproperty(self.h, owner.h), "gm_graph_share_vertexproperty")
        self._keep.append(owner)

    def set_edge_values(self, src, dst, val):
        """Graph::applyToAllEdges, host-evaluated: new value of every edge (same src/dst as at creation)."""
        src, dst, val = _i32(src), _i32(dst), _i32(val)
        _check(lib().gm_graph_set_edge_values(self.h, C.c_longlong(len(src)), _p(src), _p(dst), _p(val)),
               "gm_graph_set_edge_values")

    def push_ready(self, which=1):
        """Build the column-major companion used for sparse frontiers (0 = A, 1 = AT) ahead of the first run."""
        _check(lib().gm_graph_push_ready(self.h, C.c_int(which)), "gm_graph_push_ready")

    def set_push_policy(self, divisor=16, min_nnz=1 << 18):
        """Sparse-frontier path when frontier entries * divisor <= nnz and nnz >= min_nnz; divisor 0 = never."""
        _check(lib().gm_graph_set_push_policy(self.h, C.c_int(divisor), C.c_longlong(min_nnz)), "gm_graph_set_push_policy")

    def first_source(self):
        v = C.c_int()
        _check(lib().gm_graph_out_degree_source(self.h, C.byref(v)), "gm_graph_out_degree_source")
        return v.value

    def reduce(self, what, param=0):
        r = C.c_double()
        _check(lib().gm_graph_reduce(self.h, C.c_int(what), C.c_int(param), C.byref(r)), "gm_graph_reduce")
        return r.value

    def set_exchange(self, allgather, allreduce_or):
        self._keep += [allgather, allreduce_or]
        _check(lib().gm_graph_set_exchange(self.h, allgather, allreduce_or, None), "gm_graph_set_exchange")

    def enable_peers(self, allgather_host):
        """Map the other ranks' buffers (collective).  Returns False when the devices cannot reach each other."""
        self._keep.append(allgather_host)
        rc = lib().gm_graph_enable_peers(self.h, allgather_host, None)
        return rc == 0

    def peers_enabled(self):
        return bool(lib().gm_graph_peers_enabled(self.h))

    def slice_range(self, rank):
        f = lib().gm_graph_slice_begin
        f.restype = C.c_longlong
        return int(f(self.h, C.c_int(rank))), int(f(self.h, C.c_int(rank + 1)))

    def set_vertexproperties_slice_ptr(self, ptr):
        _check(lib().gm_graph_set_vertexproperties_slice(self.h, C.c_void_p(ptr)), "gm_graph_set_vertexproperties_slice")

    def get_vertexproperties_slice_ptr(self, ptr):
        _check(lib().gm_graph_get_vertexproperties_slice(self.h, C.c_void_p(ptr)), "gm_graph_get_vertexproperties_slice")

    def set_vertexproperties_slice(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.vdtype)
        self.set_vertexproperties_slice_ptr(arr.ctypes.data)

    def get_vertexproperties_slice(self, rank):
        lo, hi = self.slice_range(rank)
        out = np.zeros(hi - lo, self.vdtype)
        self.get_vertexproperties_slice_ptr(out.ctypes.data)
        return out

    def synchronize(self):
        _check(lib().gm_graph_synchronize(self.h), "gm_graph_synchronize")

    def run(self, program, state=None, iterations=1, vectors=None):
        """run_graph_program(&program, G, iterations, &tmp)"""
        st = RunStats()
        _check(lib().gm_run_program(self.h, C.c_int(program), C.byref(state) if state is not None else None,
                                    C.c_int(iterations), vectors.h if vectors is not None else None, C.byref(st)),
               "gm_run_program")
        return st


class Vectors:
    """graph_program_init / graph_program_clear"""

    def __init__(self, graph, program):
        sT, sU = C.c_int(), C.c_int()
        _check(lib().gm_program_sizes(C.c_int(program), C.byref(sT), C.byref(sU), None, None), "gm_program_sizes")
        self.h = C.c_void_p()
        self.graph = graph
        _check(lib().gm_vectors_create(C.byref(self.h), graph.h, sT, sU), "gm_vectors_create")
        graph._vectors.add(self)

    def view(self):
        v = VectorsView()
        _check(lib().gm_vectors_view_get(self.h, C.byref(v)), "gm_vectors_view_get")
        return v

    def close(self):
        if self.h:
            lib().gm_vectors_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.h and self.graph is not None and self.graph.h:  # garbage collection: see Graph.__del__
                lib().gm_graph_detach_host(self.graph.h)
            self.close()
        except Exception:
            pass


def fold_f32_host(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = C.c_float()
    rc = lib().gm_debug_fold_f32_host(_p(a), C.c_longlong(len(a)), C.byref(out))
    assert rc in (0, 2)
    return np.float32(out.value)


def fold_f32_device(a, warps=1, offset=0):
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = C.c_float()
    rc = lib().gm_debug_fold_f32_device(_p(a), C.c_longlong(len(a)), C.c_int(warps), C.c_int(offset), C.byref(out))
    if rc not in (0, 2):
        _check(rc, "gm_debug_fold_f32_device")
    return np.float32(out.value)
