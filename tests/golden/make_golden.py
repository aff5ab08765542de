"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container, where the reference
tree exists:   python tests/golden/make_golden.py
The inputs are the reference's own data files plus seeded graphs; outputs are the
full vertex-property arrays of the five apps.  OMP thread count = the `threads`
recorded in each file (it fixes the vertex permutation, SURVEY.md hazard 2).
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref  # noqa: E402
import util  # noqa: E402


def read_bin_mtx(path):
    """edgelist.h:242-334 binary format: int m, n, nnz then (int src, int dst, int val) records."""
    raw = open(path, "rb").read()
    m, n, nnz = struct.unpack("iii", raw[:12])
    rec = np.frombuffer(raw[12:12 + nnz * 12], dtype=np.int32).reshape(nnz, 3)  # header nnz is authoritative
    return m, n, rec[:, 0].copy(), rec[:, 1].copy(), rec[:, 2].copy()


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **kw)
    print("wrote", name)


def main():
    data = "/root/reference/data"
    # --- the reference's own fixtures ---
    m, n, s, d, v = read_bin_mtx(os.path.join(data, "test.bin.mtx"))
    assert (s == util.TEST_MTX["src"]).all() and (d == util.TEST_MTX["dst"]).all()
    for t in (1, 2, 4):
        pr, deg, it, _ = ref.pagerank(n, s, d, v, threads=t)
        depth, parent, bit, reach, _ = ref.bfs(n, s, d, 1, v, threads=t)
        dist, sit, sreach, _ = ref.sssp(n, s, d, v, 1, threads=t)
        ddist, dbucket, nb, dreach, _ = ref.deltastepping(n, s, d, v, 2, 1, threads=t)
        save("test_mtx_t%d" % t, threads=t, pagerank=pr, degree=deg, pr_iterations=it, depth=depth, parent=parent,
             bfs_iterations=bit, reachable=reach, sssp_distance=dist, sssp_iterations=sit, ds_distance=ddist,
             ds_bucket=dbucket, ds_buckets=nb)
    r = util.RATINGS7
    lv, r0, r1, _ = ref.sgd(r["m"], r["n"], r["src"], r["dst"], r["val"], K=20, threads=4)
    save("ratings7_t4", threads=4, lv=lv, rmse0=r0, rmse1=r1)
    m, n, s, d, v = read_bin_mtx(os.path.join(data, "2_10_upper_triangle.bin.mtx"))
    depth, parent, bit, reach, _ = ref.bfs(n, s, d, 1, v, threads=4)
    dist, sit, sreach, _ = ref.sssp(n, s, d, v, 1, threads=4)
    save("upper_triangle_t4", threads=4, n=n, src=s, dst=d, val=v, depth=depth, parent=parent, bfs_iterations=bit,
         reachable=reach, sssp_distance=dist, sssp_iterations=sit)
    # --- seeded RMAT, scale 12 (weights 1..127), every program ---
    for t in (1, 4):
        n, s, d, v = util.rmat_numpy(12, weight_max=127)
        src0 = util.first_source(s)
        ones = np.ones_like(v)
        pr, deg, it, _ = ref.pagerank(n, s, d, ones, threads=t)
        pr10, _, _, _ = ref.pagerank(n, s, d, ones, threads=t, iterations=10)
        depth, parent, bit, reach, _ = ref.bfs(n, s, d, src0, ones, threads=t)
        dist, sit, sreach, _ = ref.sssp(n, s, d, v, src0, threads=t)
        ddist, dbucket, nb, dreach, _ = ref.deltastepping(n, s, d, v, 16, src0, threads=t)
        save("rmat12_t%d" % t, threads=t, source=src0, pagerank=pr, pagerank10=pr10, degree=deg, pr_iterations=it,
             depth=depth, parent=parent, bfs_iterations=bit, reachable=reach, sssp_distance=dist,
             sssp_iterations=sit, ds_distance=ddist, ds_bucket=dbucket, ds_buckets=nb)
    # --- seeded ratings, K = 20 and 32 ---
    u, it_, r_ = util.ratings(300, 60, 4000)
    for K in (20, 32):
        lv, r0, r1, _ = ref.sgd(300, 360, u, it_, r_, K=K, threads=4)
        save("ratings_k%d_t4" % K, threads=4, lv=lv, rmse0=r0, rmse1=r1)
    # --- IncrementalPageRank and TopologicalSort (RMAT-12; a seeded DAG; the reference's fixture) ---
    for t in (1, 4):
        n, s, d, _ = util.rmat_numpy(12)
        dpr, ddelta, ddeg, dit, _ = ref.incremental_pagerank(n, s, d, None, threads=t)
        dpr5, ddelta5, _, _, _ = ref.incremental_pagerank(n, s, d, None, threads=t, iterations=5)
        order, indeg, tit, unreach, _ = ref.topsort(n, s, d, None, threads=t)
        nd, ds_, dd_ = util.random_dag(3000, 40000, seed=1)
        dorder, dindeg, dtit, dun, _ = ref.topsort(nd, ds_, dd_, None, threads=t)
        m = util.TEST_MTX
        mpr, mdelta, mdeg, mit, _ = ref.incremental_pagerank(m["n"], m["src"], m["dst"], m["val"], threads=t)
        morder, mindeg, mtit, mun, _ = ref.topsort(m["n"], m["src"], m["dst"], m["val"], threads=t)
        save("f3_t%d" % t, threads=t, dpr_pagerank=dpr, dpr_delta=ddelta, dpr_degree=ddeg, dpr_iterations=dit,
             dpr_pagerank5=dpr5, dpr_delta5=ddelta5, ts_order=order, ts_in_degree=indeg, ts_iterations=tit,
             ts_unreachable=unreach, dag_order=dorder, dag_in_degree=dindeg, dag_iterations=dtit, dag_unreachable=dun,
             mtx_dpr_pagerank=mpr, mtx_dpr_delta=mdelta, mtx_dpr_degree=mdeg, mtx_dpr_iterations=mit,
             mtx_ts_order=morder, mtx_ts_in_degree=mindeg, mtx_ts_iterations=mtit, mtx_ts_unreachable=mun)
    # --- LDA (K = 20): seeded document-term counts ---
    dd, tt, cc = util.doc_term_counts(300, 120, 4000)
    N, gN, ll, _ = ref.lda(300, 120, dd, tt, cc, iterations=10, threads=4)
    save("lda_t4", threads=4, N=N, global_N=gN, loglik=ll)


if __name__ == "__main__":
    main()
