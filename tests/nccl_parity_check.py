"""Real-NCCL parity check of the sharded engine (one process per GPU), run by hand / by the GPU job:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tests/nccl_parity_check.py [scale]

Every rank builds its tile-row of the same RMAT graph, the ranks run Degree + PageRank (10
iterations) and BFS with the x all-gather over NCCL, rank 0 reassembles the vertex properties and
compares them with the CPU oracle (oracle/port.py) -- bit for bit, as in tests/test_multirank.py
(which covers the same path with in-process ranks on one GPU and with gloo on CPU).
Not collected by pytest (no test_ prefix): it needs torchrun and N GPUs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ctypes as C

    import torch
    import torch.distributed as dist

    from graphmat_b200 import capi, exchange
    from oracle import port

    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    capi._check(capi.lib().gm_set_device(C.c_int(local)), "gm_set_device")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    threads = 4
    n, s, d, _ = capi.rmat_edges(scale, 16, seed=1)
    src0 = int(s.min())

    def gather_vp(G, dtype):
        out = np.zeros(n, dtype)
        G.get_vertexproperties(out)                      # fills the entries this rank owns
        t = torch.from_numpy(out.view(np.uint8).copy()).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)         # owned entries are disjoint, others are 0
        return t.cpu().numpy().view(dtype)

    # ---- PageRank ----
    G = capi.Graph.from_edges(n, s, d, None, capi.PR_DTYPE, threads=threads, rank=rank, world=world)
    tmp = capi.Vectors(G, capi.PROG_PAGERANK)
    dtmp = capi.Vectors(G, capi.PROG_DEGREE)
    exchange.attach(G, tmp, dist)
    init = np.zeros(1, capi.PR_DTYPE)
    init["pagerank"], init["degree"] = 0.3, 0
    G.set_all_vertexproperty(init[0])
    G.set_all_active()
    G.run(capi.PROG_DEGREE, None, 1, dtmp)
    G.set_all_active()
    st = G.run(capi.PROG_PAGERANK, capi.PageRankState(0.3), 10, tmp)
    got = gather_vp(G, capi.PR_DTYPE)
    ok = True
    if rank == 0:
        opr, odeg, oit = port.pagerank(n, s, d, None, threads=threads, iterations=10)
        same = bool((got["degree"] == odeg).all() and (got["pagerank"] == opr).all())
        print("PageRank RMAT-%d on %d ranks (NCCL): iterations %d, bit-identical to the oracle: %s" % (scale, world, st.iterations, same))
        ok &= same
    tmp.close(); dtmp.close(); G.close()

    # ---- BFS ----
    G = capi.Graph.from_edges(n, s, d, None, capi.BFS_DTYPE, threads=threads, rank=rank, world=world)
    tmp = capi.Vectors(G, capi.PROG_BFS)
    exchange.attach(G, tmp, dist)
    vp = np.zeros(n, capi.BFS_DTYPE)
    vp["depth"] = 0xFFFFFFFF
    vp["parent"] = np.uint64(0xFFFFFFFFFFFFFFFF)
    vp["id"] = np.arange(1, n + 1, dtype=np.uint64)
    vp["depth"][src0 - 1] = 0
    G.set_vertexproperties(vp)
    G.set_all_inactive()
    G.set_active(src0)
    st = G.run(capi.PROG_BFS, capi.BFSState(1), capi.UNTIL_CONVERGENCE, tmp)
    out = np.zeros(n, capi.BFS_DTYPE)
    G.get_vertexproperties(out)
    # depth/parent of unowned entries are 0 in `out`; MAX over ranks restores the owner's value
    # only when it is the largest, so mask with ownership instead: use a SUM of (value+1)*owned
    dep = torch.from_numpy(out["depth"].astype(np.int64)).cuda()
    par = torch.from_numpy(out["parent"].view(np.int64).copy()).cuda()
    dist.all_reduce(dep, op=dist.ReduceOp.SUM)
    dist.all_reduce(par, op=dist.ReduceOp.SUM)
    if rank == 0:
        od, op, oit, oreach = port.bfs(n, s, d, src0, threads=threads)
        same = bool((dep.cpu().numpy().astype(np.uint32) == od).all() and
                    (par.cpu().numpy().view(np.uint64) == op).all() and st.iterations == oit)
        print("BFS RMAT-%d on %d ranks (NCCL): iterations %d, depth and parent bit-identical to the oracle: %s" % (
            scale, world, st.iterations, same))
        ok &= same
    tmp.close(); G.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
