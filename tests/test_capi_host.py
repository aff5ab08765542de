"""CPU tests of the boundary: the C-ABI library loads without a GPU and exports every
symbol include/graphmat_b200.h declares; the host-side helpers (RMAT twin, id mapping
inputs, rand_r restatement) behave.  No compute call is made here."""
import ctypes as C
import os
import re

import numpy as np

from graphmat_b200 import apps, capi
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(gmlib):
    header = open(os.path.join(ROOT, "include", "graphmat_b200.h")).read()
    declared = set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", header))
    declared -= {"gm_allgather_fn", "gm_allreduce_or_fn", "gm_allgather_host_fn"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for name in declared:
        assert hasattr(gmlib, name), name


def test_binding_mirrors_the_struct_layouts(gmlib):
    """capi.py's ctypes mirrors have the sizes the library was compiled with (a silent mismatch would shift fields)"""
    out = (C.c_int * 6)()
    assert gmlib.gm_abi_struct_sizes(out) == 0
    mirrors = [capi.GraphOpts, capi.MatrixView, capi.GraphView, capi.VectorsView, capi.RunStats, capi.PushPlan]
    assert list(out) == [C.sizeof(m) for m in mirrors]


def test_struct_sizes_match_programs(gmlib):
    for prog, dt in [(capi.PROG_PAGERANK, capi.PR_DTYPE), (capi.PROG_DEGREE, capi.PR_DTYPE),
                     (capi.PROG_BFS, capi.BFS_DTYPE), (capi.PROG_SSSP, capi.SSSP_DTYPE),
                     (capi.PROG_DELTASTEPPING, capi.DS_DTYPE), (capi.PROG_SGD20, capi.latent_dtype(20)),
                     (capi.PROG_RMSE32, capi.latent_dtype(32)), (capi.PROG_SGD4, capi.latent_dtype(4))]:
        sV = C.c_int()
        assert gmlib.gm_program_sizes(C.c_int(prog), None, None, C.byref(sV), None) == 0
        assert sV.value == dt.itemsize


def test_unknown_program_is_an_error(gmlib):
    assert gmlib.gm_program_sizes(C.c_int(99), None, None, None, None) != 0
    assert b"unknown program" in gmlib.gm_last_error()


def test_rmat_host_generator():
    n, s, d, v = capi.rmat_edges(10, 16, seed=1)
    n2, s2, d2, _ = capi.rmat_edges(10, 16, seed=1)
    assert n == 1024 and len(s) == 16384 and (s == s2).all() and (d == d2).all()
    assert s.min() >= 1 and s.max() <= n and d.min() >= 1 and d.max() <= n and (v == 1).all()
    # quadrant a dominates: low ids are hot
    assert (s <= n // 2).mean() > 0.7 and (d <= n // 2).mean() > 0.7
    _, s3, _, w = capi.rmat_edges(10, 16, seed=7, weight_max=127)
    assert not (s3 == s).all() and w.min() >= 1 and w.max() <= 127


def test_rand_r_restatement():
    """apps.sgd_init restates glibc rand_r (src/SGD.cpp:176-184); check against libc itself."""
    libc = C.CDLL(None)
    lv = apps.sgd_init(50, 6)
    for i in (1, 2, 17, 50):
        seed = C.c_uint(i)
        for j in range(6):
            assert lv[i - 1, j] == libc.rand_r(C.byref(seed)) / 2147483647.0


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "_LIB_PATH", "/nonexistent/libgraphmat_b200.so")
    try:
        capi.lib()
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("expected a RuntimeError")
