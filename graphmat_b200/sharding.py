"""Host-side statement of the engine's vertex layout (mirrors csrc/gm_core.cu).

  public id (1-based) --vertexToNative--> native id (0-based)       [reference, Graph.h:111-130]
  native id --placement--> p : rank in (in-degree desc, native id asc)  [engine: hot columns first]
  p --sharding--> owner = p % world, local = p // world             [one tile-row per GPU, SURVEY 8e]
  x index = owner * n_pad + local                                    [layout of the all-gathered x]

Used by the CPU tests of the multi-rank logic and to reassemble per-rank results.
"""
import numpy as np


def to_native0(pub1, n, threads):
    """Graph::vertexToNative with nsegments = 1, returned 0-based (vectorised)."""
    v = np.asarray(pub1, dtype=np.int64) - 1
    npart = threads * 16
    height = n // npart
    vmax = height * npart
    col = v % npart
    row = v // npart
    out = np.where(v >= vmax, v, row + col * height) if height > 0 else v
    return out.astype(np.int64)


def n_pad(n, world):
    per = (n + world - 1) // world
    return max(32, (per + 31) // 32 * 32)


def placement(n, native_dst, world=1):
    """-> (owner[n], local[n], xidx[n]) indexed by native id."""
    indeg = np.bincount(native_dst, minlength=n).astype(np.int64)
    order = np.lexsort((np.arange(n), -indeg))  # primary: in-degree descending, then native id ascending
    p = np.empty(n, np.int64)
    p[order] = np.arange(n)
    owner = p % world
    local = p // world
    return owner, local, owner * n_pad(n, world) + local


def owned_public_ids(n, src, dst, threads, rank, world):
    """Public ids (1-based) whose vertex property lives on `rank`."""
    nd = to_native0(dst, n, threads)
    owner, _, _ = placement(n, nd, world)
    pub = np.arange(1, n + 1)
    return pub[owner[to_native0(pub, n, threads)] == rank]
