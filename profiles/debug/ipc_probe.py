"""torchrun -N probe of the peer-memory setup: prints why mapping fails, if it does."""
import ctypes as C, os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from graphmat_b200 import capi, exchange
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
capi._check(capi.lib().gm_set_device(C.c_int(local)), "gm_set_device")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, s, d, _ = capi.rmat_edges(12, 16, seed=1)
G = capi.Graph.from_edges(n, s, d, None, capi.PR_DTYPE, threads=4, rank=rank, world=world)
ok = exchange.attach_peers(G, dist)
print("rank", rank, "attach_peers:", ok, "| last error:", (capi.lib().gm_last_error() or b"").decode(), flush=True)
print("rank", rank, "can access peer:", [torch.cuda.can_device_access_peer(local, q) for q in range(world) if q != local], flush=True)
if ok:
    tmp = capi.Vectors(G, capi.PROG_PAGERANK)
    dtmp = capi.Vectors(G, capi.PROG_DEGREE)
    init = np.zeros(1, capi.PR_DTYPE); init["pagerank"] = 0.3
    G.set_all_vertexproperty(init[0]); G.set_all_active(); G.run(capi.PROG_DEGREE, None, 1, dtmp)
    G.set_all_active(); st = G.run(capi.PROG_PAGERANK, capi.PageRankState(0.3), 10, tmp)
    print("rank", rank, "ran", st.iterations, "iterations in", st.ms_total, "ms", flush=True)
    tmp.close(); dtmp.close()
G.close()
dist.destroy_process_group()
