"""The C++ drop-in surface (graphmat_b200/include: GraphProgram / Graph / run_graph_program) through
the five app drivers in apps/, which mirror the reference's src/*.cpp mains: run the binaries on the
reference's own fixtures (written in its binary mtx format) and compare the dumped vertex properties
with the golden vectors of the unmodified reference."""
import os
import struct
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "apps", "bin")


def write_mtx(prefix, m, n, src, dst, val):
    """edgelist.h:208-240 binary format; the loader reads <prefix>0."""
    with open(prefix + "0", "wb") as f:
        f.write(struct.pack("iii", m, n, len(src)))
        rec = np.stack([src, dst, val], axis=1).astype(np.int32)
        f.write(rec.tobytes())


def run(app, *args, threads=4, extra_env=None):
    exe = os.path.join(BIN, app)
    if not os.path.exists(exe):
        pytest.fail("apps/bin/%s is not built (python -c 'import __graft_entry__ as g; g.build()')" % app)
    env = dict(os.environ, GM_REF_THREADS=str(threads))
    env.update(extra_env or {})
    out = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    return out.stdout


def test_apps_on_test_mtx(tmp_path):
    g = np.load(os.path.join(util.GOLDEN, "test_mtx_t4.npz"))
    m = util.TEST_MTX
    prefix = str(tmp_path / "test.bin.mtx")
    write_mtx(prefix, m["n"], m["n"], m["src"], m["dst"], m["val"])
    d = str(tmp_path / "dump.txt")
    out = run("PageRank", prefix, "--dump", d)
    assert "Completed 6 iterations" in out
    a = np.loadtxt(d)
    assert (a[:, 1] == g["degree"]).all()
    assert np.max(np.abs(a[:, 2] - g["pagerank"]) / g["pagerank"]) <= 1e-6
    out = run("BFS", prefix, 1, "--dump", d)
    assert "Reachable vertices = 8" in out and "Completed 4 iterations" in out
    a = np.loadtxt(d, dtype=np.uint64)
    assert (a[:, 1] == g["depth"]).all() and (a[:, 2] == g["parent"]).all()
    out = run("SSSP", prefix, 1, "--dump", d)
    a = np.loadtxt(d, dtype=np.uint64)
    assert (a[:, 1] == g["sssp_distance"]).all()
    out = run("DeltaStepping", prefix, 2, 1, "--dump", d)
    a = np.loadtxt(d, dtype=np.int64)
    assert (a[:, 1] == g["ds_distance"]).all() and (a[:, 2] == g["ds_bucket"]).all()
    assert "buckets = %d" % int(g["ds_buckets"]) in out


def test_apps_on_rmat12(tmp_path):
    g = np.load(os.path.join(util.GOLDEN, "rmat12_t4.npz"))
    n, s, dd, v = util.rmat_numpy(12, weight_max=127)
    src0 = int(g["source"])
    prefix = str(tmp_path / "rmat12.bin.mtx")
    d = str(tmp_path / "dump.txt")
    write_mtx(prefix, n, n, s, dd, np.ones_like(v))
    run("PageRank", prefix, "--dump", d)
    a = np.loadtxt(d)
    assert (a[:, 1] == g["degree"]).all()
    assert (a[:, 2].astype(np.float32) == g["pagerank"]).all()
    run("BFS", prefix, src0, "--dump", d)
    a = np.loadtxt(d, dtype=np.uint64)
    assert (a[:, 1] == g["depth"]).all() and (a[:, 2] == g["parent"]).all()
    write_mtx(prefix, n, n, s, dd, v)
    run("SSSP", prefix, src0, "--dump", d)
    a = np.loadtxt(d, dtype=np.uint64)
    assert (a[:, 1] == g["sssp_distance"]).all()
    run("DeltaStepping", prefix, 16, src0, "--dump", d)
    a = np.loadtxt(d, dtype=np.int64)
    assert (a[:, 1] == g["ds_distance"]).all() and (a[:, 2] == g["ds_bucket"]).all()


def test_sgd_app_on_ratings7(tmp_path):
    g = np.load(os.path.join(util.GOLDEN, "ratings7_t4.npz"))
    r = util.RATINGS7
    prefix = str(tmp_path / "ratings7.bin.mtx")
    write_mtx(prefix, r["m"], r["n"], r["src"], r["dst"], r["val"])
    d = str(tmp_path / "dump.txt")
    out = run("SGD", prefix, "--dump", d)
    rm = [float(l.split("=")[1].split()[0]) for l in out.splitlines() if l.startswith("RMSE error")]
    assert abs(rm[0] - float(g["rmse0"])) < 1e-5 and abs(rm[1] - float(g["rmse1"])) < 1e-5
    a = np.loadtxt(d)
    assert np.max(np.abs(a[:, 1:] - g["lv"]) / np.abs(g["lv"])) <= 1e-6


@pytest.mark.parametrize("n", [5, 500])
def test_apply_edges_cpp(n):
    """test/test_apply_edges.cpp of the reference (applyToAllEdges + getEdgelist) on the C++ mirror, and SSSP over
    the rewritten weights (the device matrices carry them)"""
    out = run("ApplyEdgesCheck", n)
    assert "apply_edges ok" in out and "device functors ok" in out and "snapshot ok" in out


@pytest.mark.parametrize("policy", ["default", "always_push", "never_push"])
@pytest.mark.parametrize("threads", [1, 4])
def test_custom_program_cpp(policy, threads):
    """a user-defined, trait-less, order-sensitive (fp32 sum) ACTIVE_ONLY program through the C++ surface equals
    the reference's definition evaluated on the host, bit for bit, on the row-major and on the push path"""
    env = {"default": {}, "always_push": {"GM_PUSH_DIVISOR": "1", "GM_PUSH_MIN_NNZ": "0"},
           "never_push": {"GM_PUSH_DIVISOR": "0"}}[policy]
    assert "custom program ok" in run("CustomProgramCheck", 6, threads=threads, extra_env=env)


def test_f3_apps_on_fixtures(tmp_path):
    """IncrementalPageRank and TopologicalSort through the C++ mirror"""
    g = np.load(os.path.join(util.GOLDEN, "f3_t4.npz"))
    m = util.TEST_MTX
    prefix = str(tmp_path / "test.bin.mtx")
    write_mtx(prefix, m["n"], m["n"], m["src"], m["dst"], m["val"])
    d = str(tmp_path / "dump.txt")
    out = run("IncrementalPageRank", prefix, "--dump", d)
    assert "Completed %d iterations" % int(g["mtx_dpr_iterations"]) in out
    a = np.loadtxt(d)
    assert (a[:, 1] == g["mtx_dpr_degree"]).all() and (a[:, 2] == g["mtx_dpr_pagerank"]).all()
    out = run("TopologicalSort", prefix, "--dump", d)
    a = np.loadtxt(d, dtype=np.int64)
    assert (a[:, 1] == g["mtx_ts_order"]).all() and (a[:, 2] == g["mtx_ts_in_degree"]).all()
    nd, s, dd = util.random_dag(3000, 40000, seed=1)
    prefix = str(tmp_path / "dag.bin.mtx")
    write_mtx(prefix, nd, nd, s, dd, np.ones(len(s), np.int32))
    out = run("TopologicalSort", prefix, "--dump", d)
    a = np.loadtxt(d, dtype=np.int64)
    assert (a[:, 1] == g["dag_order"]).all() and "Top Sort order 1 :" in out


def test_lda_app(tmp_path):
    """LDA through the C++ mirror: do_every_iteration reduces over all vertices on the device (functor overload)"""
    g = np.load(os.path.join(util.GOLDEN, "lda_t4.npz"))
    dd, tt, cc = util.doc_term_counts(300, 120, 4000)
    prefix = str(tmp_path / "docs.bin.mtx")
    write_mtx(prefix, 420, 420, dd, tt, cc)
    d = str(tmp_path / "dump.txt")
    out = run("LDA", prefix, 300, 120, 10, "--dump", d)
    a = np.loadtxt(d)[:, 1:]
    assert np.max(np.abs(a - g["N"]) / np.maximum(np.abs(g["N"]), 1e-9)) <= 1e-6
    ll = float(out.split("Total Loglikelihood =")[1].split()[0])
    assert abs(ll - float(g["loglik"])) <= 1e-6 * abs(float(g["loglik"]))
