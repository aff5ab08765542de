"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (ms)."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        full = r[ki]
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", full))
        name = re.sub(r"<(\w+)<[^>]*>", r"<\1", name)
        t = float(r[vi].replace(",", ""))
        t = t / 1e6 if r[ui] in ("ns", "nsecond") else (t / 1e3 if r[ui] in ("us", "usecond") else t)
        agg.setdefault(name, []).append(t)
    tot = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print("%-110s n=%3d avg=%8.3f ms  share=%5.1f%%" % (k[:110], len(v), sum(v) / len(v), 100 * sum(v) / tot))


if __name__ == "__main__":
    main(sys.argv[1])
