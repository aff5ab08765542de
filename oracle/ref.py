"""ctypes bindings for oracle/_ref/libgm_ref_<app>.so -- TEST INFRASTRUCTURE ONLY.

The libraries are the UNMODIFIED reference (narayanan2004/GraphMat) compiled
single-rank by oracle/Makefile from /root/reference; see oracle/ref_driver.cpp.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def available(app="pagerank"):
    return os.path.exists(os.path.join(_HERE, "_ref", "libgm_ref_%s.so" % app))


def _cpu_has_avx512():
    try:
        flags = open("/proc/cpuinfo").read().split("flags", 1)[1].split("\n", 1)[0].split()
        return all(f in flags for f in ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"))
    except Exception:
        return False


def build_flavour(app="pagerank"):
    """which build of the reference library this host gets (oracle/Makefile)"""
    v4 = os.path.join(_HERE, "_ref", "libgm_ref_%s_v4.so" % app)
    if os.path.exists(v4) and _cpu_has_avx512() and not os.environ.get("GM_REF_NO_AVX512"):
        return "x86-64-v4", v4
    return "x86-64-v3", os.path.join(_HERE, "_ref", "libgm_ref_%s.so" % app)


def _lib(app):
    if app not in _libs:
        path = build_flavour(app)[1]
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        _libs[app] = C.CDLL(path)
    return _libs[app]


def rmat_edges(scale, edge_factor=16, seed=1, weight_max=0, weight_seed=2):
    """The synthetic RMAT input of SURVEY 8(d) from oracle/librmat.so (no product code involved)."""
    path = os.path.join(_HERE, "librmat.so")
    if "rmat" not in _libs:
        _libs["rmat"] = C.CDLL(path)
    nnz = edge_factor << scale
    src = np.empty(nnz, np.int32)
    dst = np.empty(nnz, np.int32)
    val = np.empty(nnz, np.int32)
    _libs["rmat"].gmo_rmat_edges(C.c_int(scale), C.c_int(edge_factor), C.c_ulonglong(seed), C.c_int(weight_max),
                                 C.c_ulonglong(weight_seed), _p(src), _p(dst), _p(val))
    return 1 << scale, src, dst, val


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def max_threads():
    return _lib("pagerank").gm_ref_max_threads()


def pagerank(n, src, dst, val=None, threads=4, iterations=-1):
    """-> (pagerank f32[n], degree i32[n], iterations, ms)."""
    src, dst = _i32(src), _i32(dst)
    val = _i32(val) if val is not None else np.ones(len(src), np.int32)
    pr = np.empty(n, np.float32)
    deg = np.empty(n, np.int32)
    ms = C.c_double()
    it = _lib("pagerank").gm_ref_pagerank(C.c_int(threads), C.c_int(n), C.c_int(n), C.c_int(len(src)), _p(src),
                                          _p(dst), _p(val), C.c_int(iterations), _p(pr), _p(deg), C.byref(ms))
    return pr, deg, it, ms.value


def bfs(n, src, dst, source, val=None, threads=4):
    """-> (depth u32[n], parent u64[n], iterations, reachable, ms)."""
    src, dst = _i32(src), _i32(dst)
    val = _i32(val) if val is not None else np.ones(len(src), np.int32)
    depth = np.empty(n, np.uint32)
    parent = np.empty(n, np.uint64)
    ms = C.c_double()
    reach = C.c_int()
    it = _lib("bfs").gm_ref_bfs(C.c_int(threads), C.c_int(n), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst),
                                _p(val), C.c_int(source), _p(depth), _p(parent), C.byref(reach), C.byref(ms))
    return depth, parent, it, reach.value, ms.value


def sssp(n, src, dst, val, source, threads=4):
    """-> (distance u32[n], iterations, reachable, ms)."""
    src, dst, val = _i32(src), _i32(dst), _i32(val)
    dist = np.empty(n, np.uint32)
    ms = C.c_double()
    reach = C.c_int()
    it = _lib("sssp").gm_ref_sssp(C.c_int(threads), C.c_int(n), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst),
                                  _p(val), C.c_int(source), _p(dist), C.byref(reach), C.byref(ms))
    return dist, it, reach.value, ms.value


def deltastepping(n, src, dst, val, delta, source, threads=4):
    """-> (distance u32[n], bucket i32[n], buckets_processed, reachable, ms)."""
    src, dst, val = _i32(src), _i32(dst), _i32(val)
    dist = np.empty(n, np.uint32)
    bucket = np.empty(n, np.int32)
    ms = C.c_double()
    reach = C.c_int()
    nb = _lib("deltastepping").gm_ref_deltastepping(C.c_int(threads), C.c_int(n), C.c_int(n), C.c_int(len(src)),
                                                    _p(src), _p(dst), _p(val), C.c_int(delta), C.c_int(source),
                                                    _p(dist), _p(bucket), C.byref(reach), C.byref(ms))
    return dist, bucket, nb, reach.value, ms.value


def sgd(m, n, src, dst, val, K=20, iterations=10, lam=0.001, step=0.00000035, threads=4):
    """-> (lv f64[max(m,n),K], rmse_before, rmse_after, ms)."""
    src, dst, val = _i32(src), _i32(dst), _i32(val)
    nv = max(m, n)
    lv = np.empty((nv, K), np.float64)
    rmse = np.empty(2, np.float64)
    ms = C.c_double()
    r = _lib("sgd").gm_ref_sgd(C.c_int(threads), C.c_int(K), C.c_int(m), C.c_int(n), C.c_int(len(src)), _p(src),
                               _p(dst), _p(val), C.c_int(iterations), C.c_double(lam), C.c_double(step), _p(lv),
                               _p(rmse), C.byref(ms))
    if r < 0:
        raise ValueError("reference SGD built for K in {4, 20, 32}")
    return lv, rmse[0], rmse[1], ms.value


def incremental_pagerank(n, src, dst, val=None, threads=4, iterations=-1):
    """-> (pagerank f64[n], delta f64[n], degree i32[n], iterations, ms)."""
    src, dst = _i32(src), _i32(dst)
    val = _i32(val) if val is not None else np.ones(len(src), np.int32)
    pr, de, deg = np.empty(n, np.float64), np.empty(n, np.float64), np.empty(n, np.int32)
    ms = C.c_double()
    it = _lib("incrementalpagerank").gm_ref_incremental_pagerank(
        C.c_int(threads), C.c_int(n), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst), _p(val), C.c_int(iterations),
        _p(pr), _p(de), _p(deg), C.byref(ms))
    return pr, de, deg, it, ms.value


def topsort(n, src, dst, val=None, threads=4):
    """-> (order u32[n], in_degree i32[n], iterations, unreachable, ms)."""
    src, dst = _i32(src), _i32(dst)
    val = _i32(val) if val is not None else np.ones(len(src), np.int32)
    order, indeg = np.empty(n, np.uint32), np.empty(n, np.int32)
    un, ms = C.c_int(), C.c_double()
    it = _lib("topologicalsort").gm_ref_topsort(C.c_int(threads), C.c_int(n), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst),
                                                _p(val), _p(order), _p(indeg), C.byref(un), C.byref(ms))
    return order, indeg, it, un.value, ms.value


def lda(ndoc, nterms, src, dst, val, iterations=10, alpha=1.0, eta=5.0, threads=4):
    """-> (N f64[ndoc+nterms, 20], global_N f64[20], total log-likelihood, ms)."""
    src, dst, val = _i32(src), _i32(dst), _i32(val)
    n = ndoc + nterms
    N = np.empty((n, 20), np.float64)
    gN = np.empty(20, np.float64)
    ll, ms = C.c_double(), C.c_double()
    _lib("lda").gm_ref_lda(C.c_int(threads), C.c_int(ndoc), C.c_int(nterms), C.c_int(len(src)), _p(src), _p(dst), _p(val),
                           C.c_int(iterations), C.c_double(alpha), C.c_double(eta), _p(N), _p(gN), C.byref(ll), C.byref(ms))
    return N, gN, ll.value, ms.value


class PageRankSession:
    """Build once, time run_graph_program per call (bench.py reference arm / cpu_baseline)."""

    def __init__(self, n, src, dst, val=None, threads=4):
        src, dst = _i32(src), _i32(dst)
        val = _i32(val) if val is not None else np.ones(len(src), np.int32)
        L = _lib("pagerank")
        L.gm_ref_pagerank_open.restype = C.c_void_p
        self.threads = threads
        self.n = n
        self.nnz = len(src)
        self.h = C.c_void_p(L.gm_ref_pagerank_open(C.c_int(threads), C.c_int(n), C.c_int(n), C.c_int(len(src)), _p(src),
                                                   _p(dst), _p(val)))

    def run(self, iterations):
        """-> (iterations run, ms of run_graph_program)"""
        ms = C.c_double()
        it = _lib("pagerank").gm_ref_pagerank_run(self.h, C.c_int(self.threads), C.c_int(iterations), C.byref(ms))
        return it, ms.value

    def get(self):
        """-> (pagerank f32[n], degree i32[n]) as the last run left them"""
        pr = np.empty(self.n, np.float32)
        deg = np.empty(self.n, np.int32)
        _lib("pagerank").gm_ref_pagerank_get(self.h, _p(pr), _p(deg))
        return pr, deg

    def close(self):
        if self.h:
            _lib("pagerank").gm_ref_pagerank_close(self.h)
            self.h = None
