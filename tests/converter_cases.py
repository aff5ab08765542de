"""Inputs and option sets for the graph_converter parity test (reference: src/graph_converter.cpp:162-338).
Shared by tests/test_graph_converter.py and tests/golden/make_converter_golden.py."""
import numpy as np

N_VERTICES = 300


def edges(seed=7, n=N_VERTICES, nnz=5000):
    """Random edges with self loops and duplicates.  A duplicate carries the same weight as its twin
    (weight = f(src, dst) = f(dst, src)): which twin survives --duplicatededges 0 is unspecified in the reference
    (unstable parallel sort), so the fixtures do not depend on it."""
    r = np.random.default_rng(seed)
    src = r.integers(1, n + 1, nnz)
    dst = r.integers(1, n + 1, nnz)
    dst[::97] = src[::97]                      # self loops
    k = len(src[1::53])                        # duplicates of the first k edges
    src[1::53], dst[1::53] = src[:k].copy(), dst[:k].copy()
    src[-1], dst[-1] = n, n - 1                # the largest id occurs
    return src.astype(np.int64), dst.astype(np.int64)


def weight(src, dst, kind):
    w = ((src + dst) * 31 + src * dst * 17) % 97 + 1  # symmetric: (u, v) and (v, u) tie after --bidirectional too
    if kind == "int":
        return w
    return w / 8.0 + 0.1                       # fractional: exercises %.8f / %.15lf


def write_text(path, src, dst, w=None, header=True, n=N_VERTICES, kind="int"):
    with open(path, "w") as f:
        if header:
            f.write("%d %d %d\n" % (n, n, len(src)))
        for i in range(len(src)):
            if w is None:
                f.write("%d %d\n" % (src[i], dst[i]))
            elif kind == "int":
                f.write("%d %d %d\n" % (src[i], dst[i], w[i]))
            else:
                f.write("%d %d %.6f\n" % (src[i], dst[i], w[i]))


# name -> (input kind, input has header, input has weights, converter arguments)
CASES = {
    "default_to_binary": ("int", True, True, []),
    "identity_text": ("int", True, True, ["--outputformat", "1", "--selfloops", "1", "--duplicatededges", "1"]),
    "dedupe_text": ("int", True, True, ["--outputformat", "1"]),
    "keep_selfloops": ("int", True, True, ["--outputformat", "1", "--selfloops", "1"]),
    "bidirectional": ("int", True, True, ["--outputformat", "1", "--bidirectional"]),
    "uppertriangular": ("int", True, True, ["--outputformat", "1", "--uppertriangular"]),
    "randomize_ids": ("int", True, True, ["--outputformat", "1", "--randomizeID"]),
    "random_weights": ("int", True, True, ["--outputformat", "1", "--outputedgeweights", "3", "--r", "64",
                                           "--duplicatededges", "1"]),
    "double_weights": ("real", True, True, ["--outputformat", "1", "--edgeweighttype", "1"]),
    "float_weights": ("real", True, True, ["--outputformat", "1", "--edgeweighttype", "2"]),
    "float_to_binary": ("real", True, True, ["--edgeweighttype", "2", "--bidirectional"]),
    "no_header_in": ("int", False, True, ["--outputformat", "1", "--inputheader", "0", "--nvertices", "320"]),
    "no_header_in_max": ("int", False, True, ["--outputformat", "1", "--inputheader", "0"]),
    "no_header_out": ("int", True, True, ["--outputformat", "1", "--outputheader", "0"]),
    "no_weights": ("int", True, False, ["--outputformat", "1", "--inputedgeweights", "0", "--outputedgeweights", "0"]),
    "unit_weights": ("int", True, False, ["--outputformat", "1", "--inputedgeweights", "0", "--outputedgeweights", "2"]),
    "drop_weights_binary": ("int", True, True, ["--outputedgeweights", "0", "--uppertriangular"]),
}


# binary mtx inputs: name -> (input kind, header, weights, number of input files, converter arguments)
BIN_CASES = {
    "bin_in": ("int", True, True, 1, ["--inputformat", "0", "--outputformat", "1"]),
    "bin_in_two_files": ("int", True, True, 2, ["--inputformat", "0", "--outputformat", "1"]),
    "bin_in_no_weights": ("int", True, False, 1, ["--inputformat", "0", "--outputformat", "1", "--inputedgeweights", "0",
                                                  "--outputedgeweights", "0"]),
    "bin_in_no_header": ("int", False, True, 1, ["--inputformat", "0", "--outputformat", "1", "--inputheader", "0"]),
    "bin_in_no_header_no_weights": ("int", False, False, 1, ["--inputformat", "0", "--outputformat", "0", "--inputheader", "0",
                                                             "--inputedgeweights", "0", "--outputedgeweights", "0",
                                                             "--outputheader", "0"]),
    "bin_in_double": ("real", True, True, 1, ["--inputformat", "0", "--outputformat", "1", "--edgeweighttype", "1"]),
    "bin_in_float": ("real", True, True, 1, ["--inputformat", "0", "--outputformat", "1", "--edgeweighttype", "2"]),
    "bin_out_double": ("real", True, True, 1, ["--inputformat", "0", "--edgeweighttype", "1", "--bidirectional"]),
}
ALL_CASES = sorted(CASES) + sorted(BIN_CASES)


def case_args(name):
    return CASES[name][3] if name in CASES else BIN_CASES[name][4]


def write_binary(path, src, dst, w=None, header=True, n=N_VERTICES, wtype=np.uint32):
    """edgelist.h:208-240 of the reference: int m, n, nnz, then (int src, int dst[, T val]) records"""
    with open(path, "wb") as f:
        if header:
            f.write(np.array([n, n, len(src)], np.int32).tobytes())
        if w is None:
            f.write(np.stack([src, dst], 1).astype(np.int32).tobytes())
        else:
            rec = np.zeros(len(src), dtype=[("s", np.int32), ("d", np.int32), ("v", wtype)])
            rec["s"], rec["d"], rec["v"] = src, dst, w
            f.write(rec.tobytes())


def write_input(prefix, name):
    src, dst = edges()
    if name in CASES:
        kind, header, weights, _ = CASES[name]
        w = weight(src, dst, kind) if weights else None
        write_text(prefix + "0", src, dst, w, header=header, kind=kind)
        return
    kind, header, weights, nfiles, args = BIN_CASES[name]
    wtype = np.uint32
    if "--edgeweighttype" in args:
        wtype = {"1": np.float64, "2": np.float32}[args[args.index("--edgeweighttype") + 1]]
    w = weight(src, dst, kind) if weights else None
    cut = np.linspace(0, len(src), nfiles + 1).astype(int)
    for k in range(nfiles):                    # one rank reads <prefix>0, <prefix>1, ... (edgelist.h:250-253)
        lo, hi = cut[k], cut[k + 1]
        write_binary(prefix + str(k), src[lo:hi], dst[lo:hi], None if w is None else w[lo:hi], header=header, wtype=wtype)


def write_helper_input(prefix):
    """small edge list with empty columns (odd ids never occur as a destination) for tests/host/edgelist_check.cpp"""
    r = np.random.default_rng(5)
    n, nnz = 40, 160
    src = r.integers(1, n + 1, nnz)
    dst = r.integers(1, n // 2 + 1, nnz) * 2
    write_text(prefix + "0", src, dst, weight(src, dst, "int"), n=n)


def result_lines(stdout):
    return "".join(line + "\n" for line in stdout.splitlines() if line.startswith("|"))
