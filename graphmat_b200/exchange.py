"""The per-iteration exchange of the sharded engine, supplied to the C ABI
through gm_graph_set_exchange:

  * all-gather of the message vector x (values, then bit words), in place: rank r owns the
    r-th slice of the buffer -- replaces the column broadcast of
    include/GMDP/multinode/spmspv.h:61-116 of the reference;
  * logical OR of the "some vertex changed" flag -- replaces the MPI_Allreduce(LAND) on
    "converged" of include/GraphMatRuntime.h:226.

`attach` wires torch.distributed (NCCL on GPUs).  `LocalRanks` runs several ranks inside ONE
process on ONE GPU (a thread per rank, device-to-device copies) so that the real sharded CUDA
path can be tested on a single-GPU box.  `allgather_inplace` / `allreduce_or` are the pure
torch.distributed halves, shared with the gloo CPU tests.
"""
import ctypes as C
import threading

from . import capi


class _DevPtr:
    """zero-copy view of device memory for torch.as_tensor"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def allgather_inplace(buf, rank, world, dist, group=None):
    """buf: 1-D tensor of world equal slices; slice `rank` is valid on entry, all are on exit."""
    per = buf.numel() // world
    dist.all_gather_into_tensor(buf, buf[rank * per:(rank + 1) * per].clone() if buf.device.type == "cpu"
                                else buf[rank * per:(rank + 1) * per], group=group)
    return buf


def allreduce_or(flag, dist, group=None):
    """flag: python int -> OR over ranks"""
    import torch
    t = torch.tensor([1 if flag else 0], dtype=torch.int32)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def attach(G, tmp, dist):
    """Register NCCL-backed exchange functions on graph G (one process per GPU)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    cache = {}

    def _tensor(ptr, nbytes):
        key = (ptr, nbytes)
        if key not in cache:
            cache[key] = torch.as_tensor(_DevPtr(ptr, nbytes), device="cuda")
        return cache[key]

    def ag(ctx, buf, bytes_per_rank, stream):
        try:
            t = _tensor(buf, bytes_per_rank * world)
            with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                allgather_inplace(t, rank, world, dist)
            return 0
        except Exception as e:  # surfaces as a non-zero status in the C caller
            print("graphmat_b200.exchange: all-gather failed:", e)
            return 1

    def ar(ctx, flag):
        try:
            flag[0] = allreduce_or(flag[0], dist)
            return 0
        except Exception as e:
            print("graphmat_b200.exchange: all-reduce failed:", e)
            return 1

    G.set_exchange(capi.ALLGATHER_FN(ag), capi.ALLREDUCE_OR_FN(ar))


def attach_peers(G, dist):
    """Map every rank's buffers into every process (CUDA IPC over NVLink): after this the per-iteration
    exchange is device stores + a barrier kernel, with no NCCL call and no Python in the loop.  The only
    thing torch.distributed does is the one-time all-gather of the handle blobs.  Must be called before
    the Vectors are created.  Returns False (graph unchanged) when peer mapping is not possible."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    on_gpu = dist.get_backend() == "nccl"

    def gather(ctx, mine, all_, nbytes):
        try:
            src = torch.frombuffer((C.c_char * nbytes).from_address(mine), dtype=torch.uint8).clone()
            out = torch.empty(world * nbytes, dtype=torch.uint8)
            if on_gpu:
                src, out = src.cuda(), out.cuda()
            dist.all_gather_into_tensor(out, src)
            host = out.cpu().numpy()  # keep the array alive while its bytes are copied out
            C.memmove(all_, host.ctypes.data, world * nbytes)
            return 0
        except Exception as e:
            print("graphmat_b200.exchange: host all-gather failed:", e)
            return 1

    return G.enable_peers(capi.ALLGATHER_HOST_FN(gather))


def upload_vp(G, vp, rank):
    """whole public-order array `vp` (identical on every rank) -> the sharded graph: each rank moves only its slice
    over PCIe when peers are mapped, else every rank uploads the whole array and keeps what it owns"""
    if G.peers_enabled():
        lo, hi = G.slice_range(rank)
        G.set_vertexproperties_slice(vp[lo:hi])
    else:
        G.set_vertexproperties(vp)


def collect_vp(G, dtype, nv, dist, rank, world):
    """the whole vertex-property array in public order, on every rank (tests / parity checks)"""
    import numpy as np
    import torch
    if dist is None or world == 1:
        return G.get_vertexproperties()
    if G.peers_enabled():
        mine = G.get_vertexproperties_slice(rank)
        per = (nv + world - 1) // world
        buf = np.zeros(per, dtype)
        buf[:len(mine)] = mine
        t = torch.from_numpy(buf.view(np.uint8).copy()).cuda()
        allt = torch.empty(world * t.numel(), dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allt, t)
        return allt.cpu().numpy().view(dtype)[:nv]       # slice q starts at q * per: the padding lies beyond nv
    out = np.zeros(nv, dtype)
    G.get_vertexproperties(out)                          # owned entries; the others stay 0
    t = torch.from_numpy(out.view(np.int32).copy()).cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)             # owners are disjoint
    return t.cpu().numpy().view(dtype)


class LocalRanks:
    """`world` ranks of one sharded graph inside one process / one GPU (test harness).

    graphs[r], vectors[r] belong to rank r.  run(fn) calls fn(r) on a thread per rank; the
    exchange callbacks rendezvous on barriers and copy slices device-to-device."""

    def __init__(self, graphs, vectors):
        import torch
        self.torch = torch
        self.world = len(graphs)
        self.graphs, self.vectors = graphs, vectors
        self.barrier = threading.Barrier(self.world)
        self.flags = [0] * self.world
        self.bufs = [None] * self.world
        for r, g in enumerate(graphs):
            g.set_exchange(capi.ALLGATHER_FN(self._make_ag(r)), capi.ALLREDUCE_OR_FN(self._make_ar(r)))

    @classmethod
    def with_peers(cls, graphs, make_vectors):
        """The peer-memory exchange (gm_peer.cu) between ranks of ONE process: every rank's buffers are plain
        device pointers for the others, the stores and the barrier kernel are the ones a multi-GPU run uses.
        make_vectors(graph) -> Vectors is called on a thread per rank (vector creation is collective)."""
        self = cls.__new__(cls)
        import torch
        self.torch = torch
        self.world = len(graphs)
        self.graphs = graphs
        self.barrier = threading.Barrier(self.world)
        self.blobs = [None] * self.world

        def make_gather(r):
            def gather(ctx, mine, all_, nbytes):
                self.blobs[r] = C.string_at(mine, nbytes)
                self.barrier.wait()
                C.memmove(all_, b"".join(self.blobs), nbytes * self.world)
                self.barrier.wait()
                return 0
            return gather
        ok = self.run(lambda r: graphs[r].enable_peers(capi.ALLGATHER_HOST_FN(make_gather(r))))
        if not all(ok):
            raise RuntimeError("peer mapping failed inside one process")
        self.vectors = self.run(lambda r: make_vectors(graphs[r]))
        return self

    def _make_ag(self, r):
        torch = self.torch

        def ag(ctx, buf, bytes_per_rank, stream):
            n = bytes_per_rank * self.world
            mine = torch.as_tensor(_DevPtr(buf, n), device="cuda")
            self.bufs[r] = mine
            ext = torch.cuda.ExternalStream(stream)
            ext.synchronize()            # my slice is written
            self.barrier.wait()          # everybody's slice is written
            with torch.cuda.stream(ext):
                for q in range(self.world):
                    if q != r:
                        mine[q * bytes_per_rank:(q + 1) * bytes_per_rank].copy_(
                            self.bufs[q][q * bytes_per_rank:(q + 1) * bytes_per_rank])
            ext.synchronize()
            self.barrier.wait()          # nobody overwrites a slice that is still being read
            return 0
        return ag

    def _make_ar(self, r):
        def ar(ctx, flag):
            self.flags[r] = flag[0]
            self.barrier.wait()
            v = 1 if any(self.flags) else 0
            self.barrier.wait()
            flag[0] = v
            return 0
        return ar

    def close(self):
        """destroy the vectors, then the graphs (after every rank's work is done: see gm_sym_free)"""
        for g in self.graphs:
            g.synchronize()
        for v in self.vectors:
            for one in (v if isinstance(v, (tuple, list)) else (v,)):
                one.close()
        for g in self.graphs:
            g.close()

    def run(self, fn):
        out = [None] * self.world
        err = []

        def work(r):
            try:
                out[r] = fn(r)
            except Exception as e:  # pragma: no cover
                err.append(e)
                self.barrier.abort()
        ts = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if err:
            raise err[0]
        return out
