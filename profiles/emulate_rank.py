"""One rank of an N-GPU PageRank run on ONE GPU (no-op exchange): the per-kernel times of that rank's pass.
usage: python profiles/emulate_rank.py [world] [rank] [scale]   (run it under ncu for the launch list)"""
import sys
import numpy as np
sys.path.insert(0, ".")
from graphmat_b200 import capi  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
scale = int(sys.argv[3]) if len(sys.argv) > 3 else 26
G = capi.Graph.rmat(scale, capi.PR_DTYPE, seed=1, threads=4, rank=rank, world=world)
G.set_exchange(capi.ALLGATHER_FN(lambda ctx, buf, nbytes, stream: 0), capi.ALLREDUCE_OR_FN(lambda ctx, flag: 0))
tmp, dtmp = capi.Vectors(G, capi.PROG_PAGERANK), capi.Vectors(G, capi.PROG_DEGREE)
init = np.zeros(1, capi.PR_DTYPE)
init["pagerank"] = 0.3
G.set_all_vertexproperty(init[0])
G.set_all_active()
G.run(capi.PROG_DEGREE, None, 1, dtmp)
for rep in range(3):
    G.set_all_active()
    st = G.run(capi.PROG_PAGERANK, capi.PageRankState(0.3), 10, tmp)
v = G.view()
print("rank %d of %d, RMAT-%d: %.3f ms per iteration (pass %.3f ms), %d local entries, heavy %d coop %d long %d rows" % (
    rank, world, scale, st.ms_total / 10, st.ms_spmv / 10, v.AT.nnz, v.AT.n_heavy, v.AT.n_coop, v.AT.n_long))
