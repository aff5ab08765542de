"""CPU tests: pin oracle/gm_oracle.c (the C restatement) against
  (1) the golden vectors generated from the unmodified reference (tests/golden/*.npz),
  (2) the reference itself (oracle/_ref) run live on seeded graphs, when its .so files are present,
  (3) the closed forms the reference's own tests assert (test/test_bfs.cpp:97-258) and the
      known answers decoded in SURVEY.md 8(c).
"""
import os

import numpy as np
import pytest

import util
from oracle import port, ref

G = util.GOLDEN


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


@pytest.mark.parametrize("t", [1, 2, 4])
def test_golden_test_mtx(t):
    g = load("test_mtx_t%d" % t)
    m = util.TEST_MTX
    pr, deg, it = port.pagerank(m["n"], m["src"], m["dst"], m["val"], threads=t)
    assert it == int(g["pr_iterations"])
    assert (deg == g["degree"]).all()
    np.testing.assert_allclose(pr, g["pagerank"], rtol=1e-6)
    depth, parent, bit, reach = port.bfs(m["n"], m["src"], m["dst"], 1, m["val"], threads=t)
    assert (depth == g["depth"]).all() and (parent == g["parent"]).all()
    assert bit == int(g["bfs_iterations"]) and reach == int(g["reachable"])
    dist, sit, _ = port.sssp(m["n"], m["src"], m["dst"], m["val"], 1, threads=t)
    assert (dist == g["sssp_distance"]).all() and sit == int(g["sssp_iterations"])
    ddist, dbucket, nb, _ = port.deltastepping(m["n"], m["src"], m["dst"], m["val"], 2, 1, threads=t)
    assert (ddist == g["ds_distance"]).all() and (dbucket == g["ds_bucket"]).all() and nb == int(g["ds_buckets"])


def test_survey_known_answers():
    """outputs of the reference apps on data/test.bin.mtx, OMP_NUM_THREADS=4."""
    m = util.TEST_MTX
    pr, deg, it = port.pagerank(m["n"], m["src"], m["dst"], m["val"], threads=4)
    assert it == 6
    assert list(deg) == [2, 2, 3, 3, 1, 2, 0, 0]
    np.testing.assert_allclose(pr, [0.300000, 0.405000, 0.546750, 0.569325, 0.432843, 0.560418, 0.931978, 0.623721],
                               atol=1e-6)
    depth, parent, bit, reach = port.bfs(m["n"], m["src"], m["dst"], 1, m["val"], threads=4)
    assert bit == 4 and reach == 8
    assert list(depth) == [0, 1, 1, 2, 3, 2, 3, 2]
    assert list(parent.astype(np.int64)) == [-1, 1, 1, 3, 4, 3, 6, 3]
    dist, _, _ = port.sssp(m["n"], m["src"], m["dst"], m["val"], 1, threads=4)
    assert list(dist) == [0, 1, 1, 2, 3, 2, 3, 2]
    ddist, _, nb, _ = port.deltastepping(m["n"], m["src"], m["dst"], m["val"], 2, 1, threads=4)
    assert list(ddist) == [0, 1, 1, 2, 3, 2, 3, 2] and nb == 2
    r = util.RATINGS7
    lv, r0, r1 = port.sgd(r["m"], r["n"], r["src"], r["dst"], r["val"], K=20, threads=4)
    assert abs(r0 - 2.508787) < 1e-6 and abs(r1 - 2.508619) < 1e-6


def test_golden_ratings7():
    g = load("ratings7_t4")
    r = util.RATINGS7
    lv, r0, r1 = port.sgd(r["m"], r["n"], r["src"], r["dst"], r["val"], K=20, threads=4)
    np.testing.assert_allclose(lv, g["lv"], rtol=1e-9)
    np.testing.assert_allclose([r0, r1], [g["rmse0"], g["rmse1"]], rtol=1e-9)


def test_golden_upper_triangle():
    g = load("upper_triangle_t4")
    n = int(g["n"])
    depth, parent, bit, reach = port.bfs(n, g["src"], g["dst"], 1, g["val"], threads=4)
    assert (depth == g["depth"]).all() and (parent == g["parent"]).all()
    assert bit == int(g["bfs_iterations"]) == 4 and reach == int(g["reachable"]) == 1024
    dist, sit, _ = port.sssp(n, g["src"], g["dst"], g["val"], 1, threads=4)
    assert (dist == g["sssp_distance"]).all() and sit == int(g["sssp_iterations"])


@pytest.mark.parametrize("t", [1, 4])
def test_golden_rmat12(t):
    g = load("rmat12_t%d" % t)
    n, s, d, v = util.rmat_numpy(12, weight_max=127)
    src0 = int(g["source"])
    ones = np.ones_like(v)
    pr, deg, it = port.pagerank(n, s, d, ones, threads=t)
    assert it == int(g["pr_iterations"]) and (deg == g["degree"]).all()
    assert (pr == g["pagerank"]).all()  # same fold order, same arithmetic: bit-identical
    pr10, _, _ = port.pagerank(n, s, d, ones, threads=t, iterations=10)
    assert (pr10 == g["pagerank10"]).all()
    depth, parent, bit, reach = port.bfs(n, s, d, src0, ones, threads=t)
    assert (depth == g["depth"]).all() and (parent == g["parent"]).all() and bit == int(g["bfs_iterations"])
    dist, sit, _ = port.sssp(n, s, d, v, src0, threads=t)
    assert (dist == g["sssp_distance"]).all() and sit == int(g["sssp_iterations"])
    ddist, dbucket, nb, _ = port.deltastepping(n, s, d, v, 16, src0, threads=t)
    assert (ddist == g["ds_distance"]).all() and (dbucket == g["ds_bucket"]).all() and nb == int(g["ds_buckets"])


@pytest.mark.parametrize("K", [20, 32])
def test_golden_ratings(K):
    g = load("ratings_k%d_t4" % K)
    u, it_, r_ = util.ratings(300, 60, 4000)
    lv, r0, r1 = port.sgd(300, 360, u, it_, r_, K=K, threads=4)
    np.testing.assert_allclose(lv, g["lv"], rtol=1e-9)
    np.testing.assert_allclose([r0, r1], [g["rmse0"], g["rmse1"]], rtol=1e-9)


# ---- the closed forms of the reference's own BFS tests (test/test_bfs.cpp) ----
@pytest.mark.parametrize("n", [100, 500])
@pytest.mark.parametrize("start", ["first", "mid"])
def test_bfs_closed_forms(n, start):
    s0 = 1 if start == "first" else n // 2
    # upper triangular (test_bfs.cpp:97-140): vertices below the start are unreachable, above it depth 1
    s, d = util.upper_triangular(n)
    depth, _, _, _ = port.bfs(n, s, d, s0, threads=4)
    for v in range(1, n + 1):
        exp = 0 if v == s0 else (1 if v > s0 else 0xFFFFFFFF)
        assert depth[v - 1] == exp
    # dense (:142-190): everything at depth 1
    s, d = util.dense(n)
    depth, _, _, _ = port.bfs(n, s, d, s0, threads=4)
    assert depth[s0 - 1] == 0 and (np.delete(depth, s0 - 1) == 1).all()
    # circular chain (:192-236): depth = distance along the ring
    s, d = util.circular_chain(n)
    depth, _, _, _ = port.bfs(n, s, d, s0, threads=4)
    for v in range(1, n + 1):
        assert depth[v - 1] == (v - s0) % n


# ---- live comparison with the unmodified reference build ----
needs_ref = pytest.mark.skipif(not ref.available("pagerank"), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("seed,t", [(11, 1), (12, 2), (13, 4), (14, 3)])
def test_port_vs_reference_random(seed, t):
    n, m = 700, 9000
    s, d, v = util.random_graph(n, m, seed, weight_max=50)
    src0 = util.first_source(s)
    a = port.pagerank(n, s, d, None, threads=t)
    b = ref.pagerank(n, s, d, None, threads=t)
    assert a[2] == b[2] and (a[1] == b[1]).all() and (a[0] == b[0]).all()
    a = port.bfs(n, s, d, src0, threads=t)
    b = ref.bfs(n, s, d, src0, threads=t)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2] and a[3] == b[3]
    a = port.sssp(n, s, d, v, src0, threads=t)
    b = ref.sssp(n, s, d, v, src0, threads=t)
    assert (a[0] == b[0]).all() and a[1] == b[1]
    a = port.deltastepping(n, s, d, v, 10, src0, threads=t)
    b = ref.deltastepping(n, s, d, v, 10, src0, threads=t)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2]


@needs_ref
def test_port_vs_reference_sgd():
    u, it_, r_ = util.ratings(120, 40, 1500, seed=5)
    a = port.sgd(120, 160, u, it_, r_, K=4, threads=2)
    b = ref.sgd(120, 160, u, it_, r_, K=4, threads=2)
    np.testing.assert_allclose(a[0], b[0], rtol=1e-9)
    np.testing.assert_allclose(a[1:], b[1:3], rtol=1e-9)


@needs_ref
def test_parent_depends_on_thread_count():
    """SURVEY.md hazard 2: graph {1->2, 1->17, 2->3, 17->3}, n = 64."""
    s = np.array([1, 1, 2, 17], np.int32)
    d = np.array([2, 17, 3, 3], np.int32)
    for impl in (port, ref):
        assert impl.bfs(64, s, d, 1, threads=1)[1][2] == 2
        assert impl.bfs(64, s, d, 1, threads=2)[1][2] == 17


def test_empty_and_isolated():
    """no edges at all; a source without out-edges"""
    s = np.array([], np.int32)
    pr, deg, it = port.pagerank(40, s, s, None, threads=1)
    assert it == 1 and (deg == 0).all() and np.allclose(pr, 0.3)
    depth, parent, bit, reach = port.bfs(40, np.array([2], np.int32), np.array([3], np.int32), 1, threads=1)
    assert reach == 1 and bit == 1 and depth[0] == 0


# ---- IncrementalPageRank and TopologicalSort ----
@pytest.mark.parametrize("t", [1, 4])
def test_golden_f3_programs(t):
    g = load("f3_t%d" % t)
    n, s, d, _ = util.rmat_numpy(12)
    pr, delta, deg, it = port.incremental_pagerank(n, s, d, None, threads=t)
    assert it == int(g["dpr_iterations"]) and (deg == g["dpr_degree"]).all()
    assert (pr == g["dpr_pagerank"]).all() and (delta == g["dpr_delta"]).all()      # fp64, same fold order: same bits
    pr5, delta5, _, _ = port.incremental_pagerank(n, s, d, None, threads=t, iterations=5)
    assert (pr5 == g["dpr_pagerank5"]).all() and (delta5 == g["dpr_delta5"]).all()
    order, indeg, tit, un = port.topsort(n, s, d, None, threads=t)
    assert (order == g["ts_order"]).all() and (indeg == g["ts_in_degree"]).all()
    assert tit == int(g["ts_iterations"]) and un == int(g["ts_unreachable"])
    nd, ds_, dd_ = util.random_dag(3000, 40000, seed=1)
    order, indeg, tit, un = port.topsort(nd, ds_, dd_, None, threads=t)
    assert (order == g["dag_order"]).all() and (indeg == g["dag_in_degree"]).all()
    assert tit == int(g["dag_iterations"]) and un == int(g["dag_unreachable"]) == 0
    m = util.TEST_MTX
    pr, delta, deg, it = port.incremental_pagerank(m["n"], m["src"], m["dst"], m["val"], threads=t)
    assert (pr == g["mtx_dpr_pagerank"]).all() and it == int(g["mtx_dpr_iterations"])
    order, indeg, tit, un = port.topsort(m["n"], m["src"], m["dst"], m["val"], threads=t)
    assert (order == g["mtx_ts_order"]).all() and tit == int(g["mtx_ts_iterations"]) and un == int(g["mtx_ts_unreachable"])


def test_topsort_is_a_topological_order():
    """property: on a DAG every edge goes from a lower to a strictly higher level"""
    nd, s, d = util.random_dag(2000, 30000, seed=7)
    order, indeg, _, un = port.topsort(nd, s, d, None, threads=2)
    assert un == 0 and (indeg == 0).all()
    assert (order[s - 1] < order[d - 1]).all()


def test_oracle_generator_is_the_products_generator():
    """oracle/librmat.so (bench.py's reference arm, the -m gpu parity tests) restates the product's RMAT generator"""
    from graphmat_b200 import capi
    a = ref.rmat_edges(14, 16, seed=1, weight_max=127, weight_seed=2)
    b = capi.rmat_edges(14, 16, seed=1, weight_max=127, weight_seed=2)
    assert a[0] == b[0] and all((x == y).all() for x, y in zip(a[1:], b[1:]))


def test_golden_lda():
    """LDA K = 20, 10 iterations: fp64, within north_star's 1e-6 of the unmodified reference (observed ~1e-15)"""
    g = load("lda_t4")
    dd, tt, cc = util.doc_term_counts(300, 120, 4000)
    N, gN, ll = port.lda(300, 120, dd, tt, cc, iterations=10, threads=4)
    np.testing.assert_allclose(N, g["N"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(gN, g["global_N"], rtol=1e-6)
    assert abs(ll - float(g["loglik"])) <= 1e-6 * abs(float(g["loglik"]))
    # conservation: every term count is spread over the 20 topics, once per endpoint
    assert abs(N[:300].sum() - cc.sum()) < 1e-6 * cc.sum() and abs(N[300:].sum() - cc.sum()) < 1e-6 * cc.sum()
