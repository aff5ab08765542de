"""The five app drivers of the reference, restated over the C ABI.

Each function follows the reference's run_* driver line by line (cited) but takes
the edge list from memory and returns the FULL vertex-property arrays, in the
same call shapes as oracle/port.py so the parity tests can swap one for the other.
"""
import numpy as np

from . import capi as gm


def pagerank(n, src, dst, val=None, threads=4, iterations=-1, graph=None, **kw):
    """src/PageRank.cpp:115-161 -> (pagerank f32[n], degree i32[n], iterations)"""
    G = graph or gm.Graph.from_edges(n, src, dst, val, gm.PR_DTYPE, threads=threads, **kw)
    init = np.zeros(1, gm.PR_DTYPE)
    init["pagerank"], init["degree"] = 0.3, 0
    G.set_all_vertexproperty(init[0])  # V() of Graph.h:232-234
    G.set_all_active()
    G.run(gm.PROG_DEGREE, None, 1)
    G.set_all_active()
    st = G.run(gm.PROG_PAGERANK, gm.PageRankState(0.3), iterations if iterations > 0 else gm.UNTIL_CONVERGENCE)
    vp = G.get_vertexproperties()
    return vp["pagerank"].copy(), vp["degree"].copy(), st.iterations


def bfs(n, src, dst, source, val=None, threads=4, graph=None, **kw):
    """src/BFS.cpp:110-156 -> (depth u32[n], parent u64[n], iterations, reachable)"""
    G = graph or gm.Graph.from_edges(n, src, dst, val, gm.BFS_DTYPE, threads=threads, **kw)
    vp = np.zeros(G.nvertices, gm.BFS_DTYPE)
    vp["depth"] = 0xFFFFFFFF
    vp["parent"] = np.uint64(0xFFFFFFFFFFFFFFFF)
    vp["id"] = np.arange(1, G.nvertices + 1, dtype=np.uint64)  # :114-119
    vp["depth"][source - 1] = 0                                 # :125-127
    G.set_vertexproperties(vp)
    G.set_all_inactive()
    G.set_active(source)
    state = gm.BFSState(1)
    st = G.run(gm.PROG_BFS, state, gm.UNTIL_CONVERGENCE)
    reach = int(G.reduce(gm.REDUCE_REACHABLE))                  # :143
    out = G.get_vertexproperties()
    return out["depth"].copy(), out["parent"].copy(), st.iterations, reach


def sssp(n, src, dst, val, source, threads=4, graph=None, **kw):
    """src/SSSP.cpp:102-142 -> (distance u32[n], iterations, reachable)"""
    G = graph or gm.Graph.from_edges(n, src, dst, val, gm.SSSP_DTYPE, threads=threads, **kw)
    inf = np.zeros(1, gm.SSSP_DTYPE)
    inf["distance"] = 0xFFFFFFFF
    G.set_all_vertexproperty(inf[0])
    G.set_all_inactive()
    zero = np.zeros(1, gm.SSSP_DTYPE)
    G.set_vertexproperty(source, zero[0])
    G.set_active(source)
    st = G.run(gm.PROG_SSSP, None, gm.UNTIL_CONVERGENCE)
    reach = int(G.reduce(gm.REDUCE_REACHABLE))
    out = G.get_vertexproperties()
    return out["distance"].copy(), st.iterations, reach


def deltastepping(n, src, dst, val, delta, source, threads=4, **kw):
    """src/DeltaStepping.cpp:124-198 -> (distance u32[n], bucket i32[n], buckets, reachable)"""
    src, dst, val = np.asarray(src), np.asarray(dst), np.asarray(val)
    light = val <= delta   # filter_edges(less_than_delta), :136-137
    heavy = ~light
    G = gm.Graph.from_edges(n, src[light], dst[light], val[light], gm.DS_DTYPE, threads=threads, **kw)
    G2 = gm.Graph.from_edges(n, src[heavy], dst[heavy], val[heavy], gm.DS_DTYPE, threads=threads, order_like=G, **kw)
    G2.share_vertexproperty(G)  # :142
    init = np.zeros(1, gm.DS_DTYPE)
    init["distance"], init["bucket"] = 0xFFFFFFFF, 0x7FFFFFFF
    G.set_all_vertexproperty(init[0])
    G.set_all_inactive()
    s = np.zeros(1, gm.DS_DTYPE)
    G.set_vertexproperty(source, s[0])
    G.set_active(source)
    state = gm.DeltaSteppingState(delta, 0)
    tmp = gm.Vectors(G, gm.PROG_DELTASTEPPING)
    while True:  # :166-177
        G.set_all_active()
        G.run(gm.PROG_DELTASTEPPING, state, gm.UNTIL_CONVERGENCE, tmp)
        G2.set_all_active()
        G2.run(gm.PROG_DELTASTEPPING, state, 1, tmp)
        state.bid += 1
        if int(G.reduce(gm.REDUCE_BUCKET_NOT_EMPTY, state.bid)) == 0:
            break
    reach = int(G.reduce(gm.REDUCE_REACHABLE))
    out = G.get_vertexproperties()
    tmp.close()
    G2.close()
    return out["distance"].copy(), out["bucket"].copy(), state.bid, reach


def sgd_init(nv, K):
    """src/SGD.cpp:176-184: lv[j] = rand_r(&seed = vertex id) / RAND_MAX (glibc rand_r)."""
    ids = np.arange(1, nv + 1, dtype=np.uint64)
    lv = np.empty((nv, K), np.float64)
    seed = ids.astype(np.uint32)
    M = np.uint32
    for j in range(K):
        # glibc rand_r: three LCG rounds
        seed = (seed * M(1103515245) + M(12345)).astype(np.uint32)
        r = ((seed // M(65536)) % M(2048)).astype(np.uint32)
        seed = (seed * M(1103515245) + M(12345)).astype(np.uint32)
        r = ((r << M(10)) ^ ((seed // M(65536)) % M(1024))).astype(np.uint32)
        seed = (seed * M(1103515245) + M(12345)).astype(np.uint32)
        r = ((r << M(10)) ^ ((seed // M(65536)) % M(1024))).astype(np.uint32)
        lv[:, j] = r.astype(np.float64) / 2147483647.0
    return lv


def sgd(m, n, src, dst, val, K=20, iterations=10, lam=0.001, step=0.00000035, threads=4, **kw):
    """src/SGD.cpp:163-224 -> (lv f64[max(m,n),K], rmse_before, rmse_after)"""
    nv = max(m, n)  # Graph.h:253-257
    dt = gm.latent_dtype(K)
    p_sgd, p_rmse = gm.SGD_PROGRAMS[K]
    G = gm.Graph.from_edges(nv, src, dst, val, dt, threads=threads, **kw)
    vp = np.zeros(nv, dt)
    vp["lv"] = sgd_init(nv, K)
    G.set_vertexproperties(vp)
    nnz = len(src)

    def rmse():
        G.set_all_active()
        G.run(p_rmse, None, 1)
        return float(np.sqrt(G.reduce(gm.REDUCE_SQERR) / nnz))

    before = rmse()
    G.set_all_active()
    G.run(p_sgd, gm.SGDState(lam, step), iterations)
    after = rmse()
    out = G.get_vertexproperties()
    return out["lv"].copy(), before, after


def incremental_pagerank(n, src, dst, val=None, threads=4, iterations=-1, **kw):
    """src/IncrementalPageRank.cpp:128-175 -> (pagerank f64[n], delta f64[n], degree i32[n], iterations)"""
    G = gm.Graph.from_edges(n, src, dst, val, gm.DPR_DTYPE, threads=threads, **kw)
    init = np.zeros(1, gm.DPR_DTYPE)
    init["delta"], init["pagerank"], init["degree"] = 0.3, 0.3, 0   # dPR(), :39-43
    G.set_all_vertexproperty(init[0])
    G.set_all_active()
    G.run(gm.PROG_DEGREE_DPR, None, 1)
    G.set_all_active()
    state = gm.DeltaPageRankState(0.3, 0)
    st = G.run(gm.PROG_DELTAPAGERANK, state, iterations if iterations > 0 else gm.UNTIL_CONVERGENCE)
    vp = G.get_vertexproperties()
    return vp["pagerank"].copy(), vp["delta"].copy(), vp["degree"].copy(), st.iterations


def topsort(n, src, dst, val=None, threads=4, **kw):
    """src/TopologicalSort.cpp:141-190 -> (order u32[n], in_degree i32[n], iterations, unreachable)"""
    G = gm.Graph.from_edges(n, src, dst, val, gm.TOPSORT_DTYPE, threads=threads, **kw)
    init = np.zeros(1, gm.TOPSORT_DTYPE)
    init["topsort_order"], init["in_degree"] = 0xFFFFFFFF, 0        # Vertex_type(), :44-47
    G.set_all_vertexproperty(init[0])
    G.run(gm.PROG_INDEGREE, None, 1)                                 # ALL_VERTICES: no setAllActive in the app (:153)
    vp = G.get_vertexproperties()
    roots = np.nonzero(vp["in_degree"] == 0)[0]                      # :156-167
    vp["topsort_order"][roots] = 0
    G.set_vertexproperties(vp)
    G.set_all_inactive()
    G.set_active_many(roots + 1)
    st = G.run(gm.PROG_TOPSORT, gm.TopSortState(1), gm.UNTIL_CONVERGENCE)
    unreachable = G.nvertices - int(G.reduce(gm.REDUCE_REACHABLE))   # :132-138, 177-178
    out = G.get_vertexproperties()
    return out["topsort_order"].copy(), out["in_degree"].copy(), st.iterations, unreachable


def lda(ndoc, nterms, src, dst, val, iterations=10, alpha=1.0, eta=5.0, threads=4, **kw):
    """src/LDA.cpp:274-341 (K = 20) -> (N f64[ndoc+nterms, 20], global_N f64[20], total log-likelihood)"""
    nv = ndoc + nterms
    G = gm.Graph.from_edges(nv, src, dst, val, gm.LDA_DTYPE, threads=threads, **kw)
    vp = np.zeros(nv, gm.LDA_DTYPE)
    vp["type"][:ndoc] = b"d"                                       # :285-293
    vp["type"][ndoc:] = b"w"
    G.set_vertexproperties(vp)
    G.set_all_active()
    G.run(gm.PROG_LDAINIT20, None, 1)                              # :296-298
    state = gm.LDAState(alpha, eta, float(nterms))
    G.set_all_active()
    G.run(gm.PROG_LDA20, state, iterations)                        # :300-313 (calcGlobalN before the run and per iteration)
    gN = np.array(state.global_N[:], np.float64)
    ll = gm.LDALLState()
    for i in range(20):
        ll.N_k[i] = gN[i]
    ll.eta, ll.nterms = eta, nterms
    G.set_all_active()
    G.run(gm.PROG_LDALL20, ll, 1)                                  # :335-337
    total = G.reduce(gm.REDUCE_SQERR)                              # sum of the trailing double = token_loglik, :338-339
    out = G.get_vertexproperties()
    return out["N"].copy(), gN, float(total)
