"""Seeded inputs shared by the tests (all ids public, 1-based)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# /root/reference/data/test.bin.mtx (header 8 8 13), decoded in SURVEY.md 8(c)
TEST_MTX = dict(n=8,
                src=np.array([1, 1, 2, 2, 3, 3, 3, 4, 4, 4, 5, 6, 6], np.int32),
                dst=np.array([2, 3, 3, 4, 4, 6, 8, 5, 6, 7, 7, 7, 8], np.int32),
                val=np.ones(13, np.int32))
# /root/reference/data/ratings7.bin.mtx truncated to its header nnz = 7 (hazard 6)
RATINGS7 = dict(m=4, n=7,
                src=np.array([1, 1, 2, 2, 3, 3, 4], np.int32),
                dst=np.array([5, 7, 5, 7, 6, 7, 7], np.int32),
                val=np.array([1, 2, 2, 4, 2, 3, 3], np.int32))


def rmat_numpy(scale, edge_factor=16, seed=1, a=0.57, b=0.19, c=0.19, weight_max=0, weight_seed=2):
    """numpy RMAT (independent of the library's generator), duplicates and self loops kept."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = n * edge_factor
    src = np.zeros(m, np.int64)
    dst = np.zeros(m, np.int64)
    for bit in range(scale):
        r = rng.random(m)
        sb = (r >= a + b).astype(np.int64)
        db = (((r >= a) & (r < a + b)) | (r >= a + b + c)).astype(np.int64)
        src |= sb << bit
        dst |= db << bit
    if weight_max:
        val = np.random.default_rng(weight_seed).integers(1, weight_max + 1, m).astype(np.int32)
    else:
        val = np.ones(m, np.int32)
    return n, (src + 1).astype(np.int32), (dst + 1).astype(np.int32), val


def random_graph(n, m, seed, weight_max=0):
    rng = np.random.default_rng(seed)
    src = rng.integers(1, n + 1, m).astype(np.int32)
    dst = rng.integers(1, n + 1, m).astype(np.int32)
    val = rng.integers(1, weight_max + 1, m).astype(np.int32) if weight_max else np.ones(m, np.int32)
    return src, dst, val


def upper_triangular(n):
    """test/generator.h:107-127: all (i, j) with i < j."""
    i, j = np.triu_indices(n, k=1)
    return (i + 1).astype(np.int32), (j + 1).astype(np.int32)


def dense(n):
    """test/generator.h:129-149: every ordered pair incl. the diagonal."""
    i, j = np.meshgrid(np.arange(1, n + 1), np.arange(1, n + 1), indexing="ij")
    return i.ravel().astype(np.int32), j.ravel().astype(np.int32)


def circular_chain(n):
    """test/generator.h:151-167: i -> i+1, n -> 1."""
    i = np.arange(1, n + 1)
    return i.astype(np.int32), (i % n + 1).astype(np.int32)


def first_source(src):
    return int(np.min(src))


def ratings(n_users, n_items, nnz, seed=3):
    """users uniform, items Zipf(1.0), rating uniform 1..5; items offset by n_users."""
    rng = np.random.default_rng(seed)
    u = rng.integers(1, n_users + 1, nnz)
    p = 1.0 / np.arange(1, n_items + 1)
    p /= p.sum()
    it = rng.choice(n_items, nnz, p=p) + 1 + n_users
    r = rng.integers(1, 6, nnz)
    return u.astype(np.int32), it.astype(np.int32), r.astype(np.int32)


def random_dag(n, m, seed=1):
    """seeded DAG: edges u -> v with u < v (public 1-based ids), duplicates kept"""
    rng = np.random.default_rng(seed)
    u = rng.integers(1, n, m)
    v = rng.integers(1, n + 1, m)
    keep = u < v
    return n, u[keep].astype(np.int32), v[keep].astype(np.int32)


def doc_term_counts(ndoc, nterms, nnz, seed=9):
    """seeded bipartite document-term count graph for LDA (src/LDA.cpp): documents are ids 1..ndoc, terms
    ndoc+1..ndoc+nterms, edge value = term count 1..9; every vertex gets at least one edge"""
    rng = np.random.default_rng(seed)
    d = np.concatenate([np.arange(1, ndoc + 1), rng.integers(1, ndoc + 1, nnz - ndoc)])
    t = np.concatenate([rng.integers(1, nterms + 1, ndoc), np.arange(1, nterms + 1), rng.integers(1, nterms + 1, nnz - ndoc - nterms)])[:len(d)]
    c = rng.integers(1, 10, len(d))
    return d.astype(np.int32), (t + ndoc).astype(np.int32), c.astype(np.int32)
