"""The sharded (multi-rank) path.

CPU (gloo, world_size 2): the exchange halves in graphmat_b200/exchange.py and the placement /
sharding rule in graphmat_b200/sharding.py, driven through a numpy model of one rank's iteration
(send -> all-gather x -> local rows in native column order -> apply -> OR of changed flags); the
union of the ranks' results must equal the single-rank oracle.

GPU: the real CUDA engine with world = 2 and 3, all ranks inside one process on one GPU
(exchange.LocalRanks), against the oracle: the logical layout does not depend on the sharding, so
BFS parents stay bit-exact and PageRank bit-identical.
"""
import os
import socket
import sys

import numpy as np
import pytest

import util
from graphmat_b200 import sharding
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_placement_rule():
    n, threads = 200, 2
    s, d, _ = util.random_graph(n, 3000, 5)
    nd = sharding.to_native0(d, n, threads)
    for world in (1, 2, 3):
        owner, local, xidx = sharding.placement(n, nd, world)
        assert sorted(xidx.tolist()) == sorted(set(xidx.tolist()))            # injective
        assert xidx.max() < world * sharding.n_pad(n, world)
        indeg = np.bincount(nd, minlength=n)
        for r in range(world):                                                 # rows of a rank: longest first
            mine = np.nonzero(owner == r)[0]
            mine = mine[np.argsort(local[mine])]
            assert (np.diff(indeg[mine]) <= 0).all()
        assert abs((owner == 0).sum() - n / world) <= 1                        # balanced
    # the id permutation matches the oracle's (Graph.h:111-130)
    for v in (1, 2, 31, 32, 33, 199, 200):
        assert sharding.to_native0(v, n, threads) + 1 == port.vertex_to_native(v, n, threads)


def _rank_main(rank, world, port_no, q):
    import torch
    import torch.distributed as dist
    from graphmat_b200 import exchange
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port_no, rank=rank, world_size=world)
    try:
        n, threads, iters = 512, 2, 6
        s, d, _ = util.random_graph(n, 6000, 17)
        ns, nd = sharding.to_native0(s, n, threads), sharding.to_native0(d, n, threads)
        owner, local, xidx = sharding.placement(n, nd, world)
        npad = sharding.n_pad(n, world)
        mine = np.nonzero(owner == rank)[0]                    # native ids owned here
        # owned rows of AT, entries in ascending native column (stable sort by (row, col))
        keep = owner[nd] == rank
        order = np.lexsort((ns[keep], nd[keep]))
        rows, cols = nd[keep][order], ns[keep][order]
        outdeg = np.bincount(ns, minlength=n)
        pr = np.full(n, np.float32(0.3))                       # only entries in `mine` are maintained
        x = torch.zeros(world * npad, dtype=torch.float32)
        for it in range(iters):
            msg = np.where(outdeg[mine] == 0, np.float32(0), pr[mine] / np.maximum(outdeg[mine], 1).astype(np.float32))
            x.zero_()
            x[rank * npad + torch.from_numpy(local[mine])] = torch.from_numpy(msg.astype(np.float32))
            exchange.allgather_inplace(x, rank, world, dist)   # every rank now holds all of x
            xv = x.numpy()
            y = np.zeros(n, np.float32)
            got = np.zeros(n, bool)
            first = np.ones(len(rows), bool)
            first[1:] = rows[1:] != rows[:-1]
            vals = xv[xidx[cols]]
            for r_, v_, f_ in zip(rows, vals, first):          # serial left fold per row
                y[r_] = v_ if f_ else np.float32(y[r_] + v_)
                got[r_] = True
            new = (np.float32(0.3) + (1.0 - np.float64(np.float32(0.3))) * y.astype(np.float64)).astype(np.float32)
            changed = got & (np.abs((new - pr).astype(np.float32).astype(np.float64)) > 1e-5)
            pr = np.where(got, new, pr)
            flag = exchange.allreduce_or(int(changed[mine].any()), dist)
            assert flag in (0, 1)
        q.put((rank, mine, pr[mine]))
    finally:
        dist.destroy_process_group()


def test_two_rank_pagerank_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port_no = 2, _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, threads = 512, 2
    s, d, _ = util.random_graph(n, 6000, 17)
    ref, _, _ = port.pagerank(n, s, d, None, threads=threads, iterations=6)
    full = np.zeros(n, np.float32)
    seen = np.zeros(n, bool)
    for rank, mine, vals in got:
        full[mine] = vals
        seen[mine] = True
    assert seen.all()
    pub = np.arange(1, n + 1)
    mine_by_pub = full[sharding.to_native0(pub, n, threads)]
    assert (mine_by_pub == ref).all()                         # same fold order -> bit-identical


@pytest.mark.gpu
@pytest.mark.parametrize("push", [False, True])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_engine_one_gpu(world, push):
    from graphmat_b200 import capi, exchange
    n, s, d, v = util.rmat_numpy(12, weight_max=127)
    src0 = util.first_source(s)
    threads = 4
    # ---- PageRank ----
    graphs = [capi.Graph.from_edges(n, s, d, None, capi.PR_DTYPE, threads=threads, rank=r, world=world,
                                    heavy_threshold=64, coop_threshold=512) for r in range(world)]
    vecs = [capi.Vectors(g, capi.PROG_PAGERANK) for g in graphs]
    dvecs = [capi.Vectors(g, capi.PROG_DEGREE) for g in graphs]
    init = np.zeros(1, capi.PR_DTYPE)
    init["pagerank"], init["degree"] = 0.3, 0
    lr = exchange.LocalRanks(graphs, vecs)

    def run_pr(r):
        g = graphs[r]
        g.set_all_vertexproperty(init[0])
        g.set_all_active()
        g.run(capi.PROG_DEGREE, None, 1, dvecs[r])
        g.set_all_active()
        st = g.run(capi.PROG_PAGERANK, capi.PageRankState(0.3), capi.UNTIL_CONVERGENCE, vecs[r])
        return st.iterations
    its = lr.run(run_pr)
    opr, odeg, oit = port.pagerank(n, s, d, None, threads=threads)
    assert its == [oit] * world
    out = np.zeros(n, capi.PR_DTYPE)
    for r in range(world):
        graphs[r].get_vertexproperties(out)   # each rank fills the entries it owns
    assert (out["degree"] == odeg).all()
    assert (out["pagerank"] == opr).all()
    # ---- BFS ----
    graphs = [capi.Graph.from_edges(n, s, d, None, capi.BFS_DTYPE, threads=threads, rank=r, world=world,
                                    heavy_threshold=64, coop_threshold=512) for r in range(world)]
    for g in graphs:
        g.set_push_policy(1, 0) if push else g.set_push_policy(0, 0)   # every pass / no pass on the sparse-frontier path
    vecs = [capi.Vectors(g, capi.PROG_BFS) for g in graphs]
    lr = exchange.LocalRanks(graphs, vecs)
    vp = np.zeros(n, capi.BFS_DTYPE)
    vp["depth"] = 0xFFFFFFFF
    vp["parent"] = np.uint64(0xFFFFFFFFFFFFFFFF)
    vp["id"] = np.arange(1, n + 1, dtype=np.uint64)
    vp["depth"][src0 - 1] = 0

    def run_bfs(r):
        g = graphs[r]
        g.set_vertexproperties(vp)
        g.set_all_inactive()
        g.set_active(src0)
        return g.run(capi.PROG_BFS, capi.BFSState(1), capi.UNTIL_CONVERGENCE, vecs[r]).iterations
    its = lr.run(run_bfs)
    od, op, oit, _ = port.bfs(n, s, d, src0, threads=threads)
    assert its == [oit] * world
    out = np.zeros(n, capi.BFS_DTYPE)
    for r in range(world):
        graphs[r].get_vertexproperties(out)
    assert (out["depth"] == od).all() and (out["parent"] == op).all()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_engine_peer_memory(world):
    """The peer-memory exchange (gm_peer.cu): fused apply+send stores into every rank's message buffer,
    sparse push of ACTIVE_ONLY frontiers, the barrier kernel with the OR of the changed flags, and the
    distributed (slice) vertex-property accessors -- ranks inside one process on one GPU."""
    from graphmat_b200 import apps, capi, exchange
    n, s, d, v = util.rmat_numpy(12, weight_max=127)
    src0 = util.first_source(s)
    threads = 4
    mk = lambda dt, val=None: [capi.Graph.from_edges(n, s, d, val, dt, threads=threads, rank=r, world=world,
                                                     heavy_threshold=64, coop_threshold=512) for r in range(world)]
    # ---- PageRank: fixed iterations (async loop) and until convergence ----
    graphs = mk(capi.PR_DTYPE)
    lr = exchange.LocalRanks.with_peers(graphs, lambda g: (capi.Vectors(g, capi.PROG_DEGREE), capi.Vectors(g, capi.PROG_PAGERANK)))
    assert all(g.peers_enabled() for g in graphs)
    init = np.zeros(n, capi.PR_DTYPE)
    init["pagerank"] = 0.3
    for iters in (10, capi.UNTIL_CONVERGENCE):
        def run_pr(r):
            g = graphs[r]
            lo, hi = g.slice_range(r)
            g.set_vertexproperties_slice(init[lo:hi])
            g.set_all_active()
            g.run(capi.PROG_DEGREE, None, 1, lr.vectors[r][0])
            g.set_all_active()
            st = g.run(capi.PROG_PAGERANK, capi.PageRankState(0.3), iters, lr.vectors[r][1])
            return st.iterations, g.get_vertexproperties_slice(r)
        res = lr.run(run_pr)
        opr, odeg, oit = port.pagerank(n, s, d, None, threads=threads, iterations=iters)
        assert [r[0] for r in res] == [oit] * world
        out = np.concatenate([r[1] for r in res])
        assert len(out) == n and (out["degree"] == odeg).all()
        assert (out["pagerank"] == opr).all()
    lr.close()
    # ---- BFS (sparse frontier: bit words + active values) and SSSP, pull and push passes ----
    for push in (False, True):
        graphs = mk(capi.BFS_DTYPE)
        for g in graphs:
            g.set_push_policy(1, 0) if push else g.set_push_policy(0, 0)
        lr = exchange.LocalRanks.with_peers(graphs, lambda g: capi.Vectors(g, capi.PROG_BFS))
        vp = np.zeros(n, capi.BFS_DTYPE)
        vp["depth"] = 0xFFFFFFFF
        vp["parent"] = np.uint64(0xFFFFFFFFFFFFFFFF)
        vp["id"] = np.arange(1, n + 1, dtype=np.uint64)
        vp["depth"][src0 - 1] = 0

        def prep_bfs(r):
            # building the column-major companion frees device memory (cudaFree waits for the whole device): with
            # all ranks in one process it must not overlap another rank's barrier kernel, so it is its own phase
            g = graphs[r]
            if push:
                g.push_ready(1)
            lo, hi = g.slice_range(r)
            g.set_vertexproperties_slice(vp[lo:hi])
            g.synchronize()

        def run_bfs(r):
            g = graphs[r]
            g.set_all_inactive()
            g.set_active(src0)
            it = g.run(capi.PROG_BFS, capi.BFSState(1), capi.UNTIL_CONVERGENCE, lr.vectors[r]).iterations
            return it, g.get_vertexproperties_slice(r)
        lr.run(prep_bfs)
        res = lr.run(run_bfs)
        od, op, oit, _ = port.bfs(n, s, d, src0, threads=threads)
        assert [r[0] for r in res] == [oit] * world
        out = np.concatenate([r[1] for r in res])
        assert (out["depth"] == od).all() and (out["parent"] == op).all()
        lr.close()
    graphs = mk(capi.SSSP_DTYPE, val=v)
    lr = exchange.LocalRanks.with_peers(graphs, lambda g: capi.Vectors(g, capi.PROG_SSSP))

    def run_sssp(r):
        g = graphs[r]
        inf = np.zeros(1, capi.SSSP_DTYPE)
        inf["distance"] = 0xFFFFFFFF
        g.set_all_vertexproperty(inf[0])
        g.set_all_inactive()
        g.set_vertexproperty(src0, np.zeros(1, capi.SSSP_DTYPE)[0])
        g.set_active(src0)
        it = g.run(capi.PROG_SSSP, None, capi.UNTIL_CONVERGENCE, lr.vectors[r]).iterations
        return it, g.get_vertexproperties_slice(r)
    res = lr.run(run_sssp)
    odist, osit, _ = port.sssp(n, s, d, v, src0, threads=threads)
    assert [r[0] for r in res] == [osit] * world
    assert (np.concatenate([r[1] for r in res])["distance"] == odist).all()
    lr.close()
    # ---- SGD (ALL_EDGES: separate apply kernel, dense push of 264-byte messages) ----
    u, it_, r_ = util.ratings(300, 60, 4000)
    nv, K = 360, 32
    dt = capi.latent_dtype(K)
    p_sgd, _ = capi.SGD_PROGRAMS[K]
    graphs = [capi.Graph.from_edges(nv, u, it_, r_, dt, threads=threads, rank=r, world=world) for r in range(world)]
    lr = exchange.LocalRanks.with_peers(graphs, lambda g: capi.Vectors(g, p_sgd))
    vp = np.zeros(nv, dt)
    vp["lv"] = apps.sgd_init(nv, K)

    def run_sgd(r):
        g = graphs[r]
        lo, hi = g.slice_range(r)
        g.set_vertexproperties_slice(vp[lo:hi])
        g.set_all_active()
        g.run(p_sgd, capi.SGDState(0.001, 0.00000035), 10, lr.vectors[r])
        return g.get_vertexproperties_slice(r)
    out = np.concatenate(lr.run(run_sgd))
    olv, _, _ = port.sgd(300, 360, u, it_, r_, K=K, iterations=10, threads=threads)
    err = np.abs(out["lv"] - olv) / np.maximum(np.abs(olv), 1e-300)
    assert err.max() <= 1e-6
    lr.close()
