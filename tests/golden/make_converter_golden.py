"""Writes tests/golden/converter.json: sha256 and size of what the UNMODIFIED reference's graph_converter
(oracle/_ref/graph_converter, built by `make -C oracle ref` from /root/reference/src/graph_converter.cpp)
writes for every case of tests/converter_cases.py, plus the binary->text round trip; and
tests/golden/edgelist_check.txt: what the reference's edge-list helpers return (oracle/_ref/edgelist_check).
Run in the build container:  python tests/golden/make_converter_golden.py"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import converter_cases as cc  # noqa: E402

REF = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "graph_converter")


def digest(path):
    b = open(path, "rb").read()
    return {"sha256": hashlib.sha256(b).hexdigest(), "bytes": len(b)}


def convert(exe, args, src, dst):
    out = subprocess.run([exe] + args + [src, dst], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr


def main():
    gold = {}
    with tempfile.TemporaryDirectory() as d:
        for name in cc.ALL_CASES:
            args = cc.case_args(name)
            cc.write_input(os.path.join(d, name + ".in"), name)
            convert(REF, args, os.path.join(d, name + ".in"), os.path.join(d, name + ".out"))
            gold[name] = digest(os.path.join(d, name + ".out0"))
        # binary back to text: the reader of format 0
        convert(REF, ["--inputformat", "0", "--outputformat", "1", "--selfloops", "1", "--duplicatededges", "1"],
                os.path.join(d, "default_to_binary.out"), os.path.join(d, "roundtrip"))
        gold["binary_to_text"] = digest(os.path.join(d, "roundtrip0"))
        # the helpers graph_converter does not reach, through tests/host/edgelist_check.cpp built against the reference
        cc.write_helper_input(os.path.join(d, "helpers"))
        out = subprocess.run([os.path.join(os.path.dirname(REF), "edgelist_check"), os.path.join(d, "helpers")],
                             capture_output=True, text=True, timeout=60)
        assert out.returncode == 0, out.stderr
        with open(os.path.join(HERE, "edgelist_check.txt"), "w") as f:
            f.write(cc.result_lines(out.stdout))
    with open(os.path.join(HERE, "converter.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", len(gold), "digests")


if __name__ == "__main__":
    main()
