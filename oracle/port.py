"""ctypes bindings for oracle/liboracle.so (gm_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Same call shapes as oracle/ref.py (minus timing) so tests can swap one for the
other.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this module; the product (graphmat_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(path)
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def vertex_to_native(v, n, threads):
    return lib().gmo_vertex_to_native(C.c_int(v), C.c_int(n), C.c_int(threads))


def pagerank(n, src, dst, val=None, threads=4, iterations=-1):
    src, dst = _i32(src), _i32(dst)
    val = _i32(val) if val is not None else np.ones(len(src), np.int32)
    pr = np.empty(n, np.float32)
    deg = np.empty(n, np.int32)
    it = lib().gmo_pagerank(C.c_int(threads), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst), _p(val),
                            C.c_int(iterations), _p(pr), _p(deg))
    return pr, deg, it


def bfs(n, src, dst, source, val=None, threads=4):
    src, dst = _i32(src), _i32(dst)
    val = _i32(val) if val is not None else np.ones(len(src), np.int32)
    depth = np.empty(n, np.uint32)
    parent = np.empty(n, np.uint64)
    reach = C.c_int()
    it = lib().gmo_bfs(C.c_int(threads), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst), _p(val), C.c_int(source),
                       _p(depth), _p(parent), C.byref(reach))
    return depth, parent, it, reach.value


def sssp(n, src, dst, val, source, threads=4):
    src, dst, val = _i32(src), _i32(dst), _i32(val)
    dist = np.empty(n, np.uint32)
    reach = C.c_int()
    it = lib().gmo_sssp(C.c_int(threads), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst), _p(val), C.c_int(source),
                        _p(dist), C.byref(reach))
    return dist, it, reach.value


def deltastepping(n, src, dst, val, delta, source, threads=4):
    src, dst, val = _i32(src), _i32(dst), _i32(val)
    dist = np.empty(n, np.uint32)
    bucket = np.empty(n, np.int32)
    reach = C.c_int()
    nb = lib().gmo_deltastepping(C.c_int(threads), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst), _p(val),
                                 C.c_int(delta), C.c_int(source), _p(dist), _p(bucket), C.byref(reach))
    return dist, bucket, nb, reach.value


def sgd(m, n, src, dst, val, K=20, iterations=10, lam=0.001, step=0.00000035, threads=4):
    src, dst, val = _i32(src), _i32(dst), _i32(val)
    nv = max(m, n)
    lv = np.empty((nv, K), np.float64)
    rmse = np.empty(2, np.float64)
    lib().gmo_sgd(C.c_int(threads), C.c_int(K), C.c_int(m), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst),
                  _p(val), C.c_int(iterations), C.c_double(lam), C.c_double(step), _p(lv), _p(rmse))
    return lv, rmse[0], rmse[1]


def incremental_pagerank(n, src, dst, val=None, threads=4, iterations=-1):
    """src/IncrementalPageRank.cpp:128-175 -> (pagerank f64[n], delta f64[n], degree i32[n], iterations)"""
    src, dst = _i32(src), _i32(dst)
    val = _i32(val) if val is not None else np.ones(len(src), np.int32)
    pr, de, deg = np.empty(n, np.float64), np.empty(n, np.float64), np.empty(n, np.int32)
    it = lib().gmo_incremental_pagerank(C.c_int(threads), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst), _p(val),
                                        C.c_int(iterations), _p(pr), _p(de), _p(deg))
    return pr, de, deg, it


def topsort(n, src, dst, val=None, threads=4):
    """src/TopologicalSort.cpp:141-190 -> (order u32[n], in_degree i32[n], iterations, unreachable)"""
    src, dst = _i32(src), _i32(dst)
    val = _i32(val) if val is not None else np.ones(len(src), np.int32)
    order, indeg = np.empty(n, np.uint32), np.empty(n, np.int32)
    un = C.c_int()
    it = lib().gmo_topsort(C.c_int(threads), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst), _p(val), _p(order), _p(indeg),
                           C.byref(un))
    return order, indeg, it, un.value


def lda(ndoc, nterms, src, dst, val, iterations=10, alpha=1.0, eta=5.0, threads=4):
    """src/LDA.cpp:274-341 (K = 20) -> (N f64[ndoc+nterms, 20], global_N f64[20], total log-likelihood)"""
    src, dst, val = _i32(src), _i32(dst), _i32(val)
    n = ndoc + nterms
    N = np.empty((n, 20), np.float64)
    gN = np.empty(20, np.float64)
    ll = C.c_double()
    lib().gmo_lda(C.c_int(threads), C.c_int(ndoc), C.c_int(nterms), C.c_int(len(src)), _p(src), _p(dst), _p(val),
                  C.c_int(iterations), C.c_double(alpha), C.c_double(eta), _p(N), _p(gN), C.byref(ll))
    return N, gN, ll.value


def rowblock_sum_f32(n, src, dst, row_begin, row_end, x, xbit, y, ybit, threads=4):
    src, dst = _i32(src), _i32(dst)
    lib().gmo_rowblock_sum_f32(C.c_int(threads), C.c_int(n), C.c_int(len(src)), _p(src), _p(dst), C.c_int(row_begin),
                               C.c_int(row_end), _p(x), _p(xbit), _p(y), _p(ybit))
