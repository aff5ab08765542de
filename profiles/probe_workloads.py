"""Probe timings of the non-headline configs (BFS / SSSP / DeltaStepping / SGD) on one GPU.
usage: python profiles/probe_workloads.py [bfs22] [sssp22] [ds22] [sgd]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from graphmat_b200 import apps, capi  # noqa: E402

what = sys.argv[1:] or ["bfs22", "sssp22", "ds22", "sgd"]


def timed(G, prog, state, iters, tmp=None):
    st = G.run(prog, state, iters, tmp)
    return st


if "bfs22" in what or "bfs24" in what:
    sc = 24 if "bfs24" in what else 22
    G = capi.Graph.rmat(sc, capi.BFS_DTYPE, seed=1, threads=4, build_mask=2)
    n = G.nvertices
    src0 = G.first_source()
    vp = np.zeros(n, capi.BFS_DTYPE)
    for rep in range(3):
        vp["depth"] = 0xFFFFFFFF
        vp["parent"] = np.uint64(0xFFFFFFFFFFFFFFFF)
        vp["id"] = np.arange(1, n + 1, dtype=np.uint64)
        vp["depth"][src0 - 1] = 0
        G.set_vertexproperties(vp)
        G.set_all_inactive()
        G.set_active(src0)
        st = G.run(capi.PROG_BFS, capi.BFSState(1), -1)
    print("BFS RMAT-%d: %d iterations %.3f ms (spmv %.3f) -> %.1f GTEPS, launches %d" % (
        sc, st.iterations, st.ms_total, st.ms_spmv, G.nnz / st.ms_total / 1e6, st.kernel_launches))
    G.close()

if "sssp22" in what or "ds22" in what or "sssp24" in what or "ds24" in what:
    sc = 24 if ("sssp24" in what or "ds24" in what) else 22
    n, s, d, v = capi.rmat_edges(sc, 16, seed=1, weight_max=127, weight_seed=2)
    src0 = int(s.min())
    if "sssp22" in what or "sssp24" in what:
        G = capi.Graph.from_edges(n, s, d, v, capi.SSSP_DTYPE, threads=4, build_mask=2)
        for rep in range(2):
            inf = np.zeros(1, capi.SSSP_DTYPE)
            inf["distance"] = 0xFFFFFFFF
            G.set_all_vertexproperty(inf[0])
            G.set_all_inactive()
            G.set_vertexproperty(src0, np.zeros(1, capi.SSSP_DTYPE)[0])
            G.set_active(src0)
            st = G.run(capi.PROG_SSSP, None, -1)
        print("SSSP RMAT-%d: %d iterations %.3f ms -> %.1f GTEPS, push passes %d, entries swept %d" % (
            sc, st.iterations, st.ms_total, len(s) / st.ms_total / 1e6, st.push_passes, st.edges_processed))
        G.close()
    if "ds22" in what or "ds24" in what:
        t0 = time.time()
        dist, bucket, nb, reach = apps.deltastepping(n, s, d, v, 16, src0, threads=4)
        print("DeltaStepping RMAT-%d delta 16: %d buckets, %d reachable, %.1f ms wall incl. build" % (sc, nb, reach, (time.time() - t0) * 1e3))

if "sgd" in what or "sgdbig" in what:
    rng = np.random.default_rng(3)
    nu, ni, nnz, K = 1000000, 100000, 20000000, 32
    if "sgdbig" in what:  # BASELINE.json config 4 shape (10 M x 1 M); nnz is a parameter there
        nu, ni, nnz = 10000000, 1000000, 200000000
    u = rng.integers(1, nu + 1, nnz).astype(np.int32)
    it = (np.floor(np.exp(rng.random(nnz) * np.log(ni))).astype(np.int64).clip(1, ni) + nu).astype(np.int32)
    r = rng.integers(1, 6, nnz).astype(np.int32)
    nv = nu + ni
    dt = capi.latent_dtype(K)
    G = capi.Graph.from_edges(nv, u, it, r, dt, threads=4)
    vp = np.zeros(nv, dt)
    vp["lv"] = rng.random((nv, K))
    G.set_vertexproperties(vp)
    G.set_all_active()
    st = G.run(capi.PROG_RMSE32, None, 1)
    print("RMSE pass: %.3f ms" % st.ms_total)
    G.set_all_active()
    st = G.run(capi.PROG_SGD32, capi.SGDState(0.001, 0.00000035), 3)
    print("SGD K=32 %d ratings: %.3f ms / iteration -> %.2f GTEPS (2 passes per iteration)" % (
        nnz, st.ms_total / 3, 2 * nnz / (st.ms_total / 3) / 1e6))
    G.close()
