"""graph_converter (apps/graph_converter.cu) against the reference's tool (src/graph_converter.cpp:162-338):
for every option set of converter_cases.py the output file must be byte-identical to what the UNMODIFIED
reference wrote -- compared with the committed digests (tests/golden/converter.json, made by
tests/golden/make_converter_golden.py) and, when oracle/_ref/graph_converter was built here, with a fresh
run of it.  Formats 0/1 are host work: no GPU needed."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

import converter_cases as cc
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MINE = os.path.join(ROOT, "apps", "bin", "graph_converter")
REF = os.path.join(ROOT, "oracle", "_ref", "graph_converter")
GOLD = json.load(open(os.path.join(util.GOLDEN, "converter.json")))
ROUNDTRIP = ["--inputformat", "0", "--outputformat", "1", "--selfloops", "1", "--duplicatededges", "1"]


def convert(exe, args, src, dst, ok=True):
    if not os.path.exists(exe):
        pytest.fail("%s is not built (python -c 'import __graft_entry__ as g; g.build()')" % exe)
    out = subprocess.run([exe] + list(args) + [src, dst], capture_output=True, text=True, timeout=120)
    if ok:
        assert out.returncode == 0, out.stdout + out.stderr
    return out


def digest(path):
    b = open(path, "rb").read()
    return {"sha256": hashlib.sha256(b).hexdigest(), "bytes": len(b)}


@pytest.mark.parametrize("name", cc.ALL_CASES)
def test_converter_matches_reference_output(name, tmp_path):
    args = cc.case_args(name)
    src = str(tmp_path / "in")
    cc.write_input(src, name)
    convert(MINE, args, src, str(tmp_path / "mine"))
    assert digest(str(tmp_path / "mine0")) == GOLD[name]
    if os.path.exists(REF):  # the live reference, when it was built in this container
        convert(REF, args, src, str(tmp_path / "ref"))
        assert open(str(tmp_path / "mine0"), "rb").read() == open(str(tmp_path / "ref0"), "rb").read()


def test_binary_mtx_round_trip(tmp_path):
    src = str(tmp_path / "in")
    cc.write_input(src, "default_to_binary")
    convert(MINE, cc.CASES["default_to_binary"][3], src, str(tmp_path / "bin"))
    convert(MINE, ROUNDTRIP, str(tmp_path / "bin"), str(tmp_path / "text"))
    assert digest(str(tmp_path / "text0")) == GOLD["binary_to_text"]
    # the text of the round trip is the text the converter writes directly
    assert GOLD["binary_to_text"] == GOLD["dedupe_text"]


def test_transformations_properties(tmp_path):
    """Size-independent properties on a larger input (no golden): the default conversion leaves a sorted,
    duplicate-free, loop-free edge set equal to the input's; --bidirectional output is symmetric;
    --uppertriangular has src <= dst; --randomizeID is a relabelling (degree multiset kept)."""
    r = np.random.default_rng(3)
    n, nnz = 5000, 200000
    s = r.integers(1, n + 1, nnz)
    d = r.integers(1, n + 1, nnz)
    w = cc.weight(s, d, "int")
    cc.write_text(str(tmp_path / "big0"), s, d, w, n=n)

    def run(args):
        convert(MINE, ["--outputformat", "1"] + args, str(tmp_path / "big"), str(tmp_path / "o"))
        a = np.loadtxt(str(tmp_path / "o0"), dtype=np.int64, skiprows=1)
        hdr = open(str(tmp_path / "o0")).readline().split()
        assert [int(x) for x in hdr] == [n, n, len(a)]
        return a

    want = np.unique(np.stack([s, d], 1)[s != d], axis=0)
    a = run([])
    assert np.array_equal(a[:, :2], want)                       # sorted by (src, dst), unique, no loops
    assert np.array_equal(a[:, 2], cc.weight(a[:, 0], a[:, 1], "int"))
    b = run(["--bidirectional"])
    both = np.unique(np.concatenate([want, want[:, ::-1]]), axis=0)
    assert np.array_equal(b[:, :2], both)
    u = run(["--uppertriangular"])
    assert (u[:, 0] < u[:, 1]).all()
    assert np.array_equal(u[:, :2], np.unique(np.sort(want, axis=1), axis=0))
    p = run(["--randomizeID"])
    assert len(p) == len(want) and len(np.unique(p[:, :2], axis=0)) == len(want)
    deg = lambda e: np.sort(np.bincount(e[:, 0], minlength=n + 1))  # noqa: E731
    assert np.array_equal(deg(p), deg(want))
    k = run(["--selfloops", "1", "--duplicatededges", "1"])
    assert np.array_equal(k, np.stack([s, d, w], 1))            # nothing asked: the file comes back as it was


def test_binary_reader_trusts_the_file_not_a_wrong_header(tmp_path):
    """The header's count is authoritative when the file holds more (records beyond it are ignored: the reference
    overruns its buffer there, SURVEY hazard 6); when the file holds fewer, what is there is read."""
    src, dst = cc.edges()
    w = cc.weight(src, dst, "int")
    keep = ["--inputformat", "0", "--outputformat", "1", "--selfloops", "1", "--duplicatededges", "1"]

    def text(path):
        return np.loadtxt(path, dtype=np.int64, skiprows=1)

    cc.write_binary(str(tmp_path / "a0"), src, dst, w)
    with open(str(tmp_path / "a0"), "ab") as f:
        f.write(b"\x07" * 40)                                  # three records of garbage and a bit
    convert(MINE, keep, str(tmp_path / "a"), str(tmp_path / "ao"))
    assert np.array_equal(text(str(tmp_path / "ao0")), np.stack([src, dst, w], 1))

    cc.write_binary(str(tmp_path / "b0"), src, dst, w)
    with open(str(tmp_path / "b0"), "r+b") as f:
        f.truncate(12 + 12 * 1000 + 5)                          # 1000 whole records and a torn one
    convert(MINE, keep, str(tmp_path / "b"), str(tmp_path / "bo"))
    assert np.array_equal(text(str(tmp_path / "bo0")), np.stack([src, dst, w], 1)[:1000])
    assert open(str(tmp_path / "bo0")).readline().split()[2] == "1000"


def test_converter_rejects_bad_options(tmp_path):
    src = str(tmp_path / "in")
    cc.write_input(src, "identity_text")
    out = convert(MINE, ["--uppertriangular", "--bidirectional"], src, str(tmp_path / "o"), ok=False)
    assert out.returncode != 0 and "Cannot be both uppertriangular and bidirectional" in out.stdout
    out = convert(MINE, ["--inputedgeweights", "0"], src, str(tmp_path / "o"), ok=False)
    assert out.returncode != 0 and "No input edge weights and want output edge weights" in out.stdout
    out = convert(MINE, ["--split", "2"], src, str(tmp_path / "o"), ok=False)
    assert out.returncode != 0 and "deprecated" in out.stdout
    out = convert(MINE, ["--inputheader", "0", "--nvertices", "10"], src, str(tmp_path / "o"), ok=False)
    assert out.returncode != 0 and not os.path.exists(str(tmp_path / "o0"))
    out = convert(MINE, ["--edgeweighttype", "1", "--outputformat", "2"], src, str(tmp_path / "o"), ok=False)
    assert out.returncode != 0 and "4-byte edge values" in out.stdout


def test_edgelist_helpers_match_reference(tmp_path):
    """remove_empty_columns, filter_edges_by_row, get_dimensions, filter_edges, randomize_edge_direction and
    ReadEdges(randomize) of GraphMat/edgelist.h: the same driver (tests/host/edgelist_check.cpp) built against the
    product's header and against the reference's prints the same results."""
    exe = os.path.join(ROOT, "apps", "bin", "edgelist_check")
    if not os.path.exists(exe):
        pytest.fail("apps/bin/edgelist_check is not built (python -c 'import __graft_entry__ as g; g.build()')")
    src = str(tmp_path / "helpers")
    cc.write_helper_input(src)
    mine = subprocess.run([exe, src], capture_output=True, text=True, timeout=60)
    assert mine.returncode == 0, mine.stderr
    assert cc.result_lines(mine.stdout) == open(os.path.join(util.GOLDEN, "edgelist_check.txt")).read()
    ref = os.path.join(ROOT, "oracle", "_ref", "edgelist_check")
    if os.path.exists(ref):
        live = subprocess.run([ref, src], capture_output=True, text=True, timeout=60)
        assert cc.result_lines(live.stdout) == cc.result_lines(mine.stdout)


@pytest.mark.gpu
def test_snapshot_format_round_trip(tmp_path):
    """Format 2 (this library's GraphMat-binary snapshot) -> text gives the same edge set as text -> text."""
    src = str(tmp_path / "in")
    cc.write_input(src, "dedupe_text")
    convert(MINE, ["--outputformat", "2"], src, str(tmp_path / "snap"))
    convert(MINE, ["--inputformat", "2", "--outputformat", "1"], str(tmp_path / "snap"), str(tmp_path / "back"))
    assert digest(str(tmp_path / "back0")) == GOLD["dedupe_text"]
