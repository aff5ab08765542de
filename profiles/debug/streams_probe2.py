import os, sys, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
if len(sys.argv) > 2:
    from graphmat_b200 import apps
    from oracle import ref
    n, s, d, _ = ref.rmat_edges(20, 16, seed=1)
    pr, deg, it = apps.pagerank(n, s, d, None, threads=8, iterations=int(sys.argv[2]))
    np.save(sys.argv[1], pr)
else:
    from oracle import ref
    n, s, d, _ = ref.rmat_edges(20, 16, seed=1)
    indeg = np.bincount(d - 1, minlength=n)
    for iters in (2, 3, 10):
        f = "/tmp/pr_i%d.npy" % iters
        subprocess.check_call(["timeout", "100", sys.executable, __file__, f, str(iters)], env=dict(os.environ, GM_AUX_STREAMS="2"))
        a = np.load(f)
        rpr, _, _, _ = ref.pagerank(n, s, d, None, threads=8, iterations=iters)
        bad = np.nonzero(a != rpr)[0]
        print("iterations", iters, "mismatches", len(bad), "in-degree of mismatching vertices: min", indeg[bad].min() if len(bad) else None,
              "max", indeg[bad].max() if len(bad) else None, "count >16384:", int((indeg[bad] > 16384).sum()), "of", int((indeg > 16384).sum()),
              "count 256..16384:", int(((indeg[bad] > 256) & (indeg[bad] <= 16384)).sum()), "of", int(((indeg > 256) & (indeg <= 16384)).sum()))
        if len(bad): print("   e.g.", [(int(v), int(indeg[v]), float(a[v]), float(rpr[v])) for v in bad[:4]])
