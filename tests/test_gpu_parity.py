"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
  * the committed golden vectors of the unmodified reference (tests/golden),
  * the C oracle (oracle/gm_oracle.c) on the same seeded inputs,
  * size-independent properties at larger sizes.
Bars: bit-exact for BFS depth/parent, SSSP/DeltaStepping distance/bucket, degrees and
iteration counts; PageRank and SGD within 1e-6 relative (north_star) -- PageRank is in
fact compared bit-for-bit because the engine keeps the reference's fold order.
"""
import os

import numpy as np
import pytest

import util
from graphmat_b200 import apps, capi
from oracle import port

pytestmark = pytest.mark.gpu
G = util.GOLDEN
REL = 1e-6  # north_star tolerance for floating-point programs


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def assert_rel(a, b, rel=REL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    assert err.max() <= rel, "max relative error %g" % err.max()


@pytest.mark.parametrize("t", [1, 2, 4])
def test_golden_test_mtx(t):
    g = load("test_mtx_t%d" % t)
    m = util.TEST_MTX
    pr, deg, it = apps.pagerank(m["n"], m["src"], m["dst"], m["val"], threads=t)
    assert it == int(g["pr_iterations"]) and (deg == g["degree"]).all()
    assert_rel(pr, g["pagerank"])
    depth, parent, bit, reach = apps.bfs(m["n"], m["src"], m["dst"], 1, m["val"], threads=t)
    assert (depth == g["depth"]).all() and (parent == g["parent"]).all()
    assert bit == int(g["bfs_iterations"]) and reach == int(g["reachable"])
    dist, sit, _ = apps.sssp(m["n"], m["src"], m["dst"], m["val"], 1, threads=t)
    assert (dist == g["sssp_distance"]).all() and sit == int(g["sssp_iterations"])
    ddist, dbucket, nb, _ = apps.deltastepping(m["n"], m["src"], m["dst"], m["val"], 2, 1, threads=t)
    assert (ddist == g["ds_distance"]).all() and (dbucket == g["ds_bucket"]).all() and nb == int(g["ds_buckets"])


def test_golden_ratings7():
    g = load("ratings7_t4")
    r = util.RATINGS7
    lv, r0, r1 = apps.sgd(r["m"], r["n"], r["src"], r["dst"], r["val"], K=20, threads=4)
    assert_rel(lv, g["lv"])
    assert_rel([r0, r1], [g["rmse0"], g["rmse1"]])


def test_golden_upper_triangle():
    g = load("upper_triangle_t4")
    n = int(g["n"])
    depth, parent, bit, reach = apps.bfs(n, g["src"], g["dst"], 1, g["val"], threads=4)
    assert (depth == g["depth"]).all() and (parent == g["parent"]).all()
    assert bit == int(g["bfs_iterations"]) and reach == int(g["reachable"])
    dist, sit, _ = apps.sssp(n, g["src"], g["dst"], g["val"], 1, threads=4)
    assert (dist == g["sssp_distance"]).all() and sit == int(g["sssp_iterations"])


@pytest.mark.parametrize("t", [1, 4])
@pytest.mark.parametrize("heavy", [0, 16])  # 16: push most rows through the row-cooperative kernel
def test_golden_rmat12(t, heavy):
    g = load("rmat12_t%d" % t)
    n, s, d, v = util.rmat_numpy(12, weight_max=127)
    src0 = int(g["source"])
    kw = dict(threads=t, heavy_threshold=heavy)
    pr, deg, it = apps.pagerank(n, s, d, None, **kw)
    assert it == int(g["pr_iterations"]) and (deg == g["degree"]).all()
    assert_rel(pr, g["pagerank"])
    assert (pr == g["pagerank"]).all(), "PageRank is expected bit-identical (same fold order)"
    pr10, _, _ = apps.pagerank(n, s, d, None, iterations=10, **kw)
    assert (pr10 == g["pagerank10"]).all()
    depth, parent, bit, reach = apps.bfs(n, s, d, src0, None, **kw)
    assert (depth == g["depth"]).all() and (parent == g["parent"]).all() and bit == int(g["bfs_iterations"])
    dist, sit, _ = apps.sssp(n, s, d, v, src0, **kw)
    assert (dist == g["sssp_distance"]).all() and sit == int(g["sssp_iterations"])
    ddist, dbucket, nb, _ = apps.deltastepping(n, s, d, v, 16, src0, **kw)
    assert (ddist == g["ds_distance"]).all() and (dbucket == g["ds_bucket"]).all() and nb == int(g["ds_buckets"])


@pytest.mark.parametrize("K", [20, 32])
def test_golden_ratings(K):
    g = load("ratings_k%d_t4" % K)
    u, it_, r_ = util.ratings(300, 60, 4000)
    lv, r0, r1 = apps.sgd(300, 360, u, it_, r_, K=K, threads=4)
    assert_rel(lv, g["lv"])
    assert_rel([r0, r1], [g["rmse0"], g["rmse1"]])


@pytest.mark.parametrize("n", [100, 500])
@pytest.mark.parametrize("start", ["first", "mid"])
def test_bfs_closed_forms(n, start):
    """the reference's own BFS tests (test/test_bfs.cpp:97-258)"""
    s0 = 1 if start == "first" else n // 2
    s, d = util.upper_triangular(n)
    depth, _, _, _ = apps.bfs(n, s, d, s0, threads=4)
    exp = np.where(np.arange(1, n + 1) > s0, 1, 0xFFFFFFFF).astype(np.uint32)
    exp[s0 - 1] = 0
    assert (depth == exp).all()
    s, d = util.dense(n)
    depth, _, _, _ = apps.bfs(n, s, d, s0, threads=4)
    assert depth[s0 - 1] == 0 and (np.delete(depth, s0 - 1) == 1).all()
    s, d = util.circular_chain(n)
    depth, _, _, _ = apps.bfs(n, s, d, s0, threads=4)
    assert (depth == (np.arange(1, n + 1) - s0) % n).all()


@pytest.mark.parametrize("seed,t,n,m", [(21, 1, 700, 9000), (22, 2, 1000, 40000), (23, 4, 5000, 30000),
                                        (24, 3, 33, 400), (25, 8, 4096, 100000)])
def test_vs_oracle_random(seed, t, n, m):
    s, d, v = util.random_graph(n, m, seed, weight_max=50)
    src0 = util.first_source(s)
    a = apps.pagerank(n, s, d, None, threads=t)
    b = port.pagerank(n, s, d, None, threads=t)
    assert a[2] == b[2] and (a[1] == b[1]).all()
    assert_rel(a[0], b[0])
    a = apps.bfs(n, s, d, src0, threads=t)
    b = port.bfs(n, s, d, src0, threads=t)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2] and a[3] == b[3]
    a = apps.sssp(n, s, d, v, src0, threads=t)
    b = port.sssp(n, s, d, v, src0, threads=t)
    assert (a[0] == b[0]).all() and a[1] == b[1] and a[2] == b[2]
    a = apps.deltastepping(n, s, d, v, 10, src0, threads=t)
    b = port.deltastepping(n, s, d, v, 10, src0, threads=t)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2]


def test_vs_oracle_rmat16_library_generator():
    """the library's own RMAT generator (device) and its host twin produce the same graph"""
    n, s, d, v = capi.rmat_edges(16, 16, seed=1)
    Gd = capi.Graph.rmat(16, capi.PR_DTYPE, seed=1, threads=4)
    a = apps.pagerank(n, None, None, graph=Gd, iterations=10)
    b = port.pagerank(n, s, d, None, threads=4, iterations=10)
    assert (a[1] == b[1]).all()
    assert_rel(a[0], b[0])
    assert (a[0] == b[0]).all()
    src0 = util.first_source(s)
    assert Gd.first_source() == src0
    Gb = capi.Graph.rmat(16, capi.BFS_DTYPE, seed=1, threads=4)
    a = apps.bfs(n, None, None, src0, graph=Gb)
    b = port.bfs(n, s, d, src0, threads=4)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2] and a[3] == b[3]


def test_sgd_vs_oracle():
    u, it_, r_ = util.ratings(500, 100, 20000, seed=9)
    a = apps.sgd(500, 600, u, it_, r_, K=20, threads=4)
    b = port.sgd(500, 600, u, it_, r_, K=20, threads=4)
    assert_rel(a[0], b[0])
    assert_rel(a[1:], b[1:])


def test_empty_and_edge_cases():
    e = np.array([], np.int32)
    pr, deg, it = apps.pagerank(40, e, e, None, threads=1)
    assert it == 1 and (deg == 0).all() and np.allclose(pr, 0.3)
    depth, parent, bit, reach = apps.bfs(40, np.array([2], np.int32), np.array([3], np.int32), 1, threads=1)
    assert reach == 1 and bit == 1 and depth[0] == 0
    # thread-count dependent parent (SURVEY.md hazard 2)
    s = np.array([1, 1, 2, 17], np.int32)
    d = np.array([2, 17, 3, 3], np.int32)
    assert apps.bfs(64, s, d, 1, threads=1)[1][2] == 2
    assert apps.bfs(64, s, d, 1, threads=2)[1][2] == 17
    # single-vertex accessors go through the id permutation (test/test_graph_basics.cpp:56-145)
    Gr = capi.Graph.from_edges(100, s, d, None, capi.SSSP_DTYPE, threads=2)
    for v in (1, 2, 50, 100):
        Gr.set_vertexproperty(v, (v * 7,))
    for v in (1, 2, 50, 100):
        assert Gr.get_vertexproperty(v)["distance"] == v * 7
    allv = Gr.get_vertexproperties()
    assert allv["distance"][49] == 350 and allv["distance"][0] == 7


@pytest.mark.parametrize("scale", [20, 22])
def test_properties_at_scale(scale):
    """RMAT-20 (16.8 M edges) and RMAT-22 (67 M edges, BASELINE.json's BFS configuration): size-independent checks
    instead of the oracle -- Degree == out-degree histogram, PageRank only moves vertices that received a message,
    Graph500-style BFS validation (tree edges exist, levels differ by one, no edge skips a level), SSSP distances
    satisfy every edge's triangle inequality and are attained by some in-edge."""
    Gd = capi.Graph.rmat(scale, capi.PR_DTYPE, seed=1, threads=4)
    n = Gd.nvertices
    pr, deg, it = apps.pagerank(n, None, None, graph=Gd, iterations=3)
    Gd.close()
    _, s, d, w = capi.rmat_edges(scale, 16, seed=1, weight_max=127)
    assert (deg == np.bincount(s - 1, minlength=n)).all()          # Degree pass == out-degree histogram
    indeg = np.bincount(d - 1, minlength=n)
    assert (pr[indeg == 0] == np.float32(0.3)).all()                # apply only where a message arrived
    assert np.isfinite(pr).all() and (pr >= np.float32(0.3) - 1e-6).all()
    src0 = util.first_source(s)
    Gb = capi.Graph.rmat(scale, capi.BFS_DTYPE, seed=1, threads=4, build_mask=2)
    depth, parent, bit, reach = apps.bfs(n, None, None, src0, graph=Gb)
    Gb.close()
    vis = depth < 0xFFFFFFFF
    assert reach == vis.sum() and depth[src0 - 1] == 0
    # every visited non-source vertex has a visited parent exactly one level up, and (parent, v) is an edge
    idx = np.nonzero(vis)[0]
    idx = idx[idx != src0 - 1]
    p = parent[idx].astype(np.int64)
    assert (depth[p - 1] + 1 == depth[idx]).all()
    key = s.astype(np.int64) * (n + 1) + d
    key.sort()
    q = p * (n + 1) + (idx + 1)
    pos = np.searchsorted(key, q)
    assert (key[np.minimum(pos, len(key) - 1)] == q).all()
    # no edge skips a level
    assert (depth[d - 1][vis[s - 1]] <= depth[s - 1][vis[s - 1]] + 1).all()
    del key
    # SSSP: fixed point of the relaxation
    Gs = capi.Graph.rmat(scale, capi.SSSP_DTYPE, seed=1, weight_max=127, threads=4, build_mask=2)
    dist, sit, sreach = apps.sssp(n, None, None, None, src0, graph=Gs)
    Gs.close()
    dd = dist.astype(np.int64)
    fin = dist < 0xFFFFFFFF
    assert sreach == fin.sum() == reach and dist[src0 - 1] == 0
    es = fin[s - 1]
    assert (dd[d - 1][es] <= dd[s - 1][es] + w[es]).all()          # no edge can still relax
    best = np.full(n, np.iinfo(np.int64).max)
    np.minimum.at(best, d[es] - 1, dd[s[es] - 1] + w[es])
    chk = fin.copy()
    chk[src0 - 1] = False
    assert (best[chk] == dd[chk]).all()                             # every distance is attained by an in-edge


# ---- sparse-frontier (push) path: every SpMSpV pass forced through it must still equal the oracle ----
def _bfs_run(Gb, src0):
    n = Gb.nvertices
    vp = np.zeros(n, capi.BFS_DTYPE)
    vp["depth"] = 0xFFFFFFFF
    vp["parent"] = np.uint64(0xFFFFFFFFFFFFFFFF)
    vp["id"] = np.arange(1, n + 1, dtype=np.uint64)
    vp["depth"][src0 - 1] = 0
    Gb.set_vertexproperties(vp)
    Gb.set_all_inactive()
    Gb.set_active(src0)
    st = Gb.run(capi.PROG_BFS, capi.BFSState(1), capi.UNTIL_CONVERGENCE)
    out = Gb.get_vertexproperties()
    return out["depth"].copy(), out["parent"].copy(), st


def _sssp_run(Gs, src0):
    inf = np.zeros(1, capi.SSSP_DTYPE)
    inf["distance"] = 0xFFFFFFFF
    Gs.set_all_vertexproperty(inf[0])
    Gs.set_all_inactive()
    Gs.set_vertexproperty(src0, np.zeros(1, capi.SSSP_DTYPE)[0])
    Gs.set_active(src0)
    st = Gs.run(capi.PROG_SSSP, None, capi.UNTIL_CONVERGENCE)
    return Gs.get_vertexproperties()["distance"].copy(), st


@pytest.mark.parametrize("sort_path", [False, True])
@pytest.mark.parametrize("case", ["rmat12", "rmat14_heavy16", "random", "test_mtx"])
def test_push_path_forced(case, sort_path, monkeypatch):
    """BFS (last writer) and SSSP (min) take the atomic push; GM_NO_ATOMIC_PUSH sends them through the sorted-triples
    path that programs without such a trait use -- both must reproduce the reference's fold"""
    if sort_path:
        monkeypatch.setenv("GM_NO_ATOMIC_PUSH", "1")
    t = 4
    heavy = 0
    if case == "rmat12":
        n, s, d, v = util.rmat_numpy(12, weight_max=127)
    elif case == "rmat14_heavy16":
        n, s, d, v = capi.rmat_edges(14, 16, seed=3, weight_max=127)
        heavy = 16
    elif case == "random":
        n = 3000
        s, d, v = util.random_graph(n, 50000, 31, weight_max=50)
    else:
        m = util.TEST_MTX
        n, s, d, v = m["n"], m["src"], m["dst"], m["val"]
    src0 = util.first_source(s)
    od, op, oit, _ = port.bfs(n, s, d, src0, threads=t)
    Gb = capi.Graph.from_edges(n, s, d, None, capi.BFS_DTYPE, threads=t, heavy_threshold=heavy)
    Gb.set_push_policy(1, 0)            # frontier entries <= nnz always holds: every pass is a push pass
    depth, parent, st = _bfs_run(Gb, src0)
    assert st.push_passes == st.iterations == oit
    assert (depth == od).all() and (parent == op).all()
    Gb.set_push_policy(0, 0)            # never push: the row-major kernels
    depth, parent, st = _bfs_run(Gb, src0)
    assert st.push_passes == 0 and (depth == od).all() and (parent == op).all()
    odist, osit, _ = port.sssp(n, s, d, v, src0, threads=t)
    Gs = capi.Graph.from_edges(n, s, d, v, capi.SSSP_DTYPE, threads=t, heavy_threshold=heavy)
    Gs.set_push_policy(1, 0)
    dist, st = _sssp_run(Gs, src0)
    assert st.push_passes == st.iterations == osit
    assert (dist == odist).all()
    Gs.set_push_policy(4, 0)            # mixed: push while the frontier holds <= 1/4 of the entries
    dist, st = _sssp_run(Gs, src0)
    assert (dist == odist).all() and st.iterations == osit


def test_push_path_default_policy_rmat18():
    """default policy on a graph large enough to use it: some passes push, some sweep; results as the oracle"""
    n, s, d, v = capi.rmat_edges(18, 16, seed=1, weight_max=127)
    src0 = util.first_source(s)
    Gb = capi.Graph.rmat(18, capi.BFS_DTYPE, seed=1, threads=4, build_mask=2)
    Gb.push_ready(1)
    depth, parent, st = _bfs_run(Gb, src0)
    od, op, oit, _ = port.bfs(n, s, d, src0, threads=4)
    assert 0 < st.push_passes < st.iterations == oit
    assert (depth == od).all() and (parent == op).all()
    Gs = capi.Graph.from_edges(n, s, d, v, capi.SSSP_DTYPE, threads=4, build_mask=2)
    dist, st = _sssp_run(Gs, src0)
    odist, osit, _ = port.sssp(n, s, d, v, src0, threads=4)
    assert st.push_passes > 0 and st.iterations == osit and (dist == odist).all()


def test_set_edge_values_refills_both_matrices():
    """Graph::applyToAllEdges path of the C ABI: new weights, same structure; vertex properties survive"""
    n = 2000
    s, d, v = util.random_graph(n, 30000, 41, weight_max=50)
    src0 = util.first_source(s)
    Gs = capi.Graph.from_edges(n, s, d, v, capi.SSSP_DTYPE, threads=4)
    dist, st = _sssp_run(Gs, src0)
    assert (dist == port.sssp(n, s, d, v, src0, threads=4)[0]).all()
    v2 = ((v.astype(np.int64) * 7 + s + 3 * d) % 61 + 1).astype(np.int32)
    before = Gs.get_vertexproperties().copy()
    Gs.set_edge_values(s, d, v2)
    assert (Gs.get_vertexproperties() == before).all()
    for policy in ((0, 0), (1, 0)):                 # row-major kernels, then the push path (companion rebuilt)
        Gs.set_push_policy(*policy)
        dist, st = _sssp_run(Gs, src0)
        odist, osit, _ = port.sssp(n, s, d, v2, src0, threads=4)
        assert (dist == odist).all() and st.iterations == osit
    # A (IN_EDGES operand) is refilled too: DeltaStepping-free check through SGD's ALL_EDGES is heavy; use the
    # weighted in-degree instead: Degree-like sums are unweighted, so compare SSSP on the transposed graph
    Gt = capi.Graph.from_edges(n, d, s, v, capi.SSSP_DTYPE, threads=4)
    Gt.set_edge_values(d, s, v2)
    src1 = util.first_source(d)
    dist, st = _sssp_run(Gt, src1)
    assert (dist == port.sssp(n, d, s, v2, src1, threads=4)[0]).all()


def test_staged_long_rows_pagerank(monkeypatch):
    """the longest rows' gathers hoisted into a staging array (k_stage_rows + k_heavy_fadd32<STAGED>): same bits"""
    n, s, d, _ = capi.rmat_edges(15, 16, seed=5)
    monkeypatch.setenv("GM_LONG_ROW", "700")
    Gd = capi.Graph.from_edges(n, s, d, None, capi.PR_DTYPE, threads=4, heavy_threshold=128, coop_threshold=300)
    monkeypatch.delenv("GM_LONG_ROW")
    v = Gd.view()
    assert v.AT.n_long > 0 and v.AT.n_long < v.AT.n_coop and v.AT.long_entries > 700 * v.AT.n_long
    a = apps.pagerank(n, None, None, graph=Gd, iterations=8)
    b = port.pagerank(n, s, d, None, threads=4, iterations=8)
    assert (a[1] == b[1]).all() and (a[0] == b[0]).all()


def test_pagerank_unfused_paths(monkeypatch):
    """The fused apply+send epilogue is the default for PageRank; the separate apply kernel (with and without the
    fused send) must give the same bits."""
    g = load("rmat12_t4")
    n, s, d, _ = util.rmat_numpy(12)
    for env in ("GM_NO_EPILOGUE", "GM_NO_FUSE"):
        monkeypatch.setenv(env, "1")
        pr10, deg, _ = apps.pagerank(n, s, d, None, iterations=10, threads=4, heavy_threshold=16)
        pr, _, it = apps.pagerank(n, s, d, None, threads=4)
        monkeypatch.delenv(env)
        assert (pr10 == g["pagerank10"]).all() and (deg == g["degree"]).all()
        assert (pr == g["pagerank"]).all() and it == int(g["pr_iterations"])


def test_bad_edge_ids_are_refused():
    """an id outside [1, n] must come back as an error, not as an out-of-bounds write on the device."""
    s = np.array([1, 2, 9], np.int32)
    d = np.array([2, 3, 1], np.int32)
    with pytest.raises(RuntimeError, match="outside"):
        capi.Graph.from_edges(8, s, d, None, capi.PR_DTYPE)
    with pytest.raises(RuntimeError, match="outside"):
        capi.Graph.from_edges(8, d, np.array([0, 1, 2], np.int32), None, capi.PR_DTYPE)


# ---- BASELINE.json's configurations at their size, against the UNMODIFIED reference (oracle/_ref) ----
from oracle import ref  # noqa: E402

needs_ref = pytest.mark.skipif(not ref.available("bfs"), reason="oracle/_ref is built where /root/reference exists")
REF_THREADS = 8  # the reference's OpenMP thread count fixes its vertex permutation; the engine is told the same


@needs_ref
@pytest.mark.parametrize("scale", [20, 22])
def test_pagerank_rmat_vs_reference(scale):
    """PageRank x10 on RMAT-20/22 (the rows above 100 K entries go through k_heavy_fadd32<16>): bit-identical"""
    n, s, d, _ = ref.rmat_edges(scale, 16, seed=1)
    rpr, rdeg, rit, _ = ref.pagerank(n, s, d, None, threads=REF_THREADS, iterations=10)
    pr, deg, it = apps.pagerank(n, s, d, None, threads=REF_THREADS, iterations=10)
    assert it == rit == 10 and (deg == rdeg).all()
    assert_rel(pr, rpr)
    assert (pr == rpr).all()


@needs_ref
def test_bfs_rmat22_vs_reference():
    """BASELINE config 2: BFS on RMAT scale-22, depth AND parent arrays bit-exact (src/BFS.cpp:110-156)"""
    n, s, d, _ = ref.rmat_edges(22, 16, seed=1)
    src0 = int(s.min())
    rd, rp, rit, rreach, _ = ref.bfs(n, s, d, src0, None, threads=REF_THREADS)
    depth, parent, it, reach = apps.bfs(n, s, d, src0, threads=REF_THREADS)
    assert it == rit and reach == rreach
    assert (depth == rd).all() and (parent == rp).all()


@needs_ref
def test_sssp_rmat22_vs_reference():
    n, s, d, w = ref.rmat_edges(22, 16, seed=1, weight_max=127, weight_seed=2)
    src0 = int(s.min())
    rdist, rit, rreach, _ = ref.sssp(n, s, d, w, src0, threads=REF_THREADS)
    dist, it, reach = apps.sssp(n, s, d, w, src0, threads=REF_THREADS)
    assert it == rit and reach == rreach and (dist == rdist).all()


@needs_ref
def test_deltastepping_rmat20_vs_reference():
    n, s, d, w = ref.rmat_edges(20, 16, seed=1, weight_max=127, weight_seed=2)
    src0 = int(s.min())
    rdist, rbucket, rnb, rreach, _ = ref.deltastepping(n, s, d, w, 16, src0, threads=REF_THREADS)
    dist, bucket, nb, reach = apps.deltastepping(n, s, d, w, 16, src0, threads=REF_THREADS)
    assert nb == rnb and reach == rreach and (dist == rdist).all() and (bucket == rbucket).all()


# ---- IncrementalPageRank (ACTIVE_ONLY, fp64, order-sensitive sum) and TopologicalSort (T = bool) ----
@pytest.mark.parametrize("t", [1, 4])
@pytest.mark.parametrize("heavy", [0, 16])
def test_golden_f3_programs_gpu(t, heavy):
    g = load("f3_t%d" % t)
    n, s, d, _ = util.rmat_numpy(12)
    kw = dict(threads=t, heavy_threshold=heavy)
    pr, delta, deg, it = apps.incremental_pagerank(n, s, d, None, **kw)
    assert it == int(g["dpr_iterations"]) and (deg == g["dpr_degree"]).all()
    assert_rel(pr, g["dpr_pagerank"])
    assert (pr == g["dpr_pagerank"]).all() and (delta == g["dpr_delta"]).all()      # every row folded in the reference's order
    pr5, delta5, _, _ = apps.incremental_pagerank(n, s, d, None, iterations=5, **kw)
    assert (pr5 == g["dpr_pagerank5"]).all() and (delta5 == g["dpr_delta5"]).all()
    order, indeg, tit, un = apps.topsort(n, s, d, None, **kw)
    assert (order == g["ts_order"]).all() and (indeg == g["ts_in_degree"]).all()
    assert tit == int(g["ts_iterations"]) and un == int(g["ts_unreachable"])
    nd, ds_, dd_ = util.random_dag(3000, 40000, seed=1)
    order, indeg, tit, un = apps.topsort(nd, ds_, dd_, None, **kw)
    assert (order == g["dag_order"]).all() and (indeg == g["dag_in_degree"]).all()
    assert tit == int(g["dag_iterations"]) and un == 0
    m = util.TEST_MTX
    pr, delta, deg, it = apps.incremental_pagerank(m["n"], m["src"], m["dst"], m["val"], threads=t)
    assert (pr == g["mtx_dpr_pagerank"]).all() and it == int(g["mtx_dpr_iterations"])
    order, indeg, tit, un = apps.topsort(m["n"], m["src"], m["dst"], m["val"], threads=t)
    assert (order == g["mtx_ts_order"]).all() and tit == int(g["mtx_ts_iterations"]) and un == int(g["mtx_ts_unreachable"])


@needs_ref
def test_f3_programs_rmat18_vs_reference(monkeypatch):
    """larger inputs, sparse-frontier passes included (IncrementalPageRank takes the sorted-triples path: its fp64
    sum has no trait), against the unmodified reference"""
    n, s, d, _ = ref.rmat_edges(18, 16, seed=1)
    rpr, rdelta, rdeg, rit, _ = ref.incremental_pagerank(n, s, d, None, threads=REF_THREADS)
    pr, delta, deg, it = apps.incremental_pagerank(n, s, d, None, threads=REF_THREADS)
    assert it == rit and (deg == rdeg).all()
    assert_rel(pr, rpr)
    assert (pr == rpr).all() and (delta == rdelta).all()
    nd, ds_, dd_ = util.random_dag(1 << 18, 4 << 18, seed=3)
    rorder, rindeg, rtit, run_, _ = ref.topsort(nd, ds_, dd_, None, threads=REF_THREADS)
    order, indeg, tit, un = apps.topsort(nd, ds_, dd_, None, threads=REF_THREADS)
    assert tit == rtit and un == run_ and (order == rorder).all() and (indeg == rindeg).all()


def test_golden_lda_gpu():
    """LDA (ALL_EDGES, 176-byte messages, rand_r inside process_message, global_N reduced on the device in every
    do_every_iteration) against the golden output of the unmodified reference"""
    g = load("lda_t4")
    dd, tt, cc = util.doc_term_counts(300, 120, 4000)
    N, gN, ll = apps.lda(300, 120, dd, tt, cc, iterations=10, threads=4)
    assert_rel(N + 1e-9, g["N"] + 1e-9)
    assert_rel(gN, g["global_N"])
    assert abs(ll - float(g["loglik"])) <= 1e-6 * abs(float(g["loglik"]))
    oN, ogN, oll = port.lda(300, 120, dd, tt, cc, iterations=3, threads=4)
    N3, gN3, ll3 = apps.lda(300, 120, dd, tt, cc, iterations=3, threads=4, heavy_threshold=16)
    assert_rel(N3 + 1e-9, oN + 1e-9)
    assert abs(ll3 - oll) <= 1e-6 * abs(oll)
