"""Multi-GPU parity worker (one process per GPU), launched by tests/test_multigpu.py or by hand:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tests/multigpu_parity_check.py [--exchange peers|nccl] [--scale S]

Every rank builds its tile-row of the same inputs and the ranks run PageRank (fixed count and until
convergence), BFS, SSSP, DeltaStepping and SGD with the message vector exchanged either through peer
memory over NVLink (gm_peer.cu: stores from the kernels + the barrier kernel) or through NCCL
all-gathers; rank 0 reassembles the vertex properties and compares them with the CPU oracle
(oracle/port.py): bit for bit for everything but SGD (1e-6 relative, north_star's tolerance).
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import util
    from graphmat_b200 import apps, capi, exchange
    from oracle import port

    ap = argparse.ArgumentParser()
    ap.add_argument("--exchange", default="peers", choices=["peers", "nccl"])
    ap.add_argument("--scale", type=int, default=14)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    capi._check(capi.lib().gm_set_device(C.c_int(local)), "gm_set_device")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    threads = 4
    n, s, d, w = capi.rmat_edges(args.scale, 16, seed=1, weight_max=127)
    src0 = int(s.min())
    ok = True

    def attach(G):
        if args.exchange == "peers":
            if not exchange.attach_peers(G, dist):
                raise RuntimeError("peer memory cannot be mapped between these GPUs: " +
                                   (capi.lib().gm_last_error() or b"").decode())
        else:
            exchange.attach(G, None, dist)

    def collect(G, dtype, nv):
        return exchange.collect_vp(G, dtype, nv, dist, rank, world)

    def upload(G, vp):
        exchange.upload_vp(G, vp, rank)

    def report(what, same):
        nonlocal ok
        if rank == 0:
            print("%s on %d ranks (%s): %s" % (what, world, args.exchange, "identical to the oracle" if same else "MISMATCH"),
                  flush=True)
            ok &= bool(same)

    def allsum(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    mk = lambda dt, val=None, **kw: capi.Graph.from_edges(n, s, d, val, dt, threads=threads, rank=rank, world=world, **kw)

    # ---- PageRank: fixed count (the whole run enqueued at once) and until convergence ----
    G = mk(capi.PR_DTYPE)
    attach(G)
    tmp, dtmp = capi.Vectors(G, capi.PROG_PAGERANK), capi.Vectors(G, capi.PROG_DEGREE)
    init = np.zeros(n, capi.PR_DTYPE)
    init["pagerank"] = 0.3
    for iters in (10, capi.UNTIL_CONVERGENCE):
        upload(G, init)
        G.set_all_active()
        G.run(capi.PROG_DEGREE, None, 1, dtmp)
        G.set_all_active()
        st = G.run(capi.PROG_PAGERANK, capi.PageRankState(0.3), iters, tmp)
        got = collect(G, capi.PR_DTYPE, n)
        opr, odeg, oit = port.pagerank(n, s, d, None, threads=threads, iterations=iters)
        report("PageRank RMAT-%d, %d iterations" % (args.scale, st.iterations),
               (got["degree"] == odeg).all() and (got["pagerank"] == opr).all() and st.iterations == oit)
    tmp.close(); dtmp.close(); G.close()

    # ---- BFS: row-major passes only, then with the sparse-frontier path ----
    od, op, oit, _ = port.bfs(n, s, d, src0, threads=threads)
    for push in (0, 1):
        G = mk(capi.BFS_DTYPE)
        attach(G)
        G.set_push_policy(16 if push else 0, 0)
        tmp = capi.Vectors(G, capi.PROG_BFS)
        vp = np.zeros(n, capi.BFS_DTYPE)
        vp["depth"] = 0xFFFFFFFF
        vp["parent"] = np.uint64(0xFFFFFFFFFFFFFFFF)
        vp["id"] = np.arange(1, n + 1, dtype=np.uint64)
        vp["depth"][src0 - 1] = 0
        upload(G, vp)
        G.set_all_inactive()
        G.set_active(src0)
        st = G.run(capi.PROG_BFS, capi.BFSState(1), capi.UNTIL_CONVERGENCE, tmp)
        got = collect(G, capi.BFS_DTYPE, n)
        report("BFS RMAT-%d (push %d), %d iterations, %d push passes" % (args.scale, push, st.iterations, st.push_passes),
               (got["depth"] == od).all() and (got["parent"] == op).all() and st.iterations == oit)
        tmp.close(); G.close()

    # ---- SSSP ----
    G = mk(capi.SSSP_DTYPE, w)
    attach(G)
    tmp = capi.Vectors(G, capi.PROG_SSSP)
    inf = np.zeros(1, capi.SSSP_DTYPE)
    inf["distance"] = 0xFFFFFFFF
    G.set_all_vertexproperty(inf[0])
    G.set_all_inactive()
    G.set_vertexproperty(src0, np.zeros(1, capi.SSSP_DTYPE)[0])
    G.set_active(src0)
    st = G.run(capi.PROG_SSSP, None, capi.UNTIL_CONVERGENCE, tmp)
    got = collect(G, capi.SSSP_DTYPE, n)
    odist, osit, _ = port.sssp(n, s, d, w, src0, threads=threads)
    report("SSSP RMAT-%d, %d iterations" % (args.scale, st.iterations), (got["distance"] == odist).all() and st.iterations == osit)
    tmp.close(); G.close()

    # ---- DeltaStepping (two graphs sharing the vertex properties, src/DeltaStepping.cpp:124-198) ----
    delta = 16
    light = w <= delta
    G = capi.Graph.from_edges(n, s[light], d[light], w[light], capi.DS_DTYPE, threads=threads, rank=rank, world=world)
    G2 = capi.Graph.from_edges(n, s[~light], d[~light], w[~light], capi.DS_DTYPE, threads=threads, rank=rank, world=world,
                               order_like=G)
    G2.share_vertexproperty(G)
    attach(G)
    attach(G2)
    tmp = capi.Vectors(G, capi.PROG_DELTASTEPPING)
    init1 = np.zeros(1, capi.DS_DTYPE)
    init1["distance"], init1["bucket"] = 0xFFFFFFFF, 0x7FFFFFFF
    G.set_all_vertexproperty(init1[0])
    G.set_all_inactive()
    G.set_vertexproperty(src0, np.zeros(1, capi.DS_DTYPE)[0])
    G.set_active(src0)
    state = capi.DeltaSteppingState(delta, 0)
    while True:
        G.set_all_active()
        G.run(capi.PROG_DELTASTEPPING, state, capi.UNTIL_CONVERGENCE, tmp)
        G2.set_all_active()
        G2.run(capi.PROG_DELTASTEPPING, state, 1, tmp)
        state.bid += 1
        if allsum(G.reduce(capi.REDUCE_BUCKET_NOT_EMPTY, state.bid)) == 0:
            break
    got = collect(G, capi.DS_DTYPE, n)
    odd, odb, onb, _ = port.deltastepping(n, s, d, w, delta, src0, threads=threads)
    report("DeltaStepping RMAT-%d, %d buckets" % (args.scale, state.bid),
           (got["distance"] == odd).all() and (got["bucket"] == odb).all() and state.bid == onb)
    tmp.close(); G2.close(); G.close()

    # ---- SGD K=32 (ALL_EDGES, 264-byte messages) ----
    m_users, n_items, K = 3000, 400, 32
    u, it_, r_ = util.ratings(m_users, n_items, 60000)
    nv = m_users + n_items
    dt = capi.latent_dtype(K)
    p_sgd, p_rmse = capi.SGD_PROGRAMS[K]
    G = capi.Graph.from_edges(nv, u, it_, r_, dt, threads=threads, rank=rank, world=world)
    attach(G)
    tmp = capi.Vectors(G, p_sgd)
    vp = np.zeros(nv, dt)
    vp["lv"] = apps.sgd_init(nv, K)
    upload(G, vp)
    G.set_all_active()
    G.run(p_sgd, capi.SGDState(0.001, 0.00000035), 10, tmp)
    got = collect(G, dt, nv)
    olv, _, _ = port.sgd(m_users, nv, u, it_, r_, K=K, iterations=10, threads=threads)
    err = np.abs(got["lv"] - olv) / np.maximum(np.abs(olv), 1e-300)
    report("SGD K=%d, %d ratings, max relative error %.2e" % (K, len(u), err.max()), err.max() <= 1e-6)
    tmp.close(); G.close()

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
