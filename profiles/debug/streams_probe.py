import os, sys, subprocess
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
if len(sys.argv) > 1:
    from graphmat_b200 import apps, capi
    from oracle import ref
    n, s, d, _ = ref.rmat_edges(20, 16, seed=1)
    pr, deg, it = apps.pagerank(n, s, d, None, threads=8, iterations=10)
    np.save(sys.argv[1], np.stack([pr, deg.astype(np.float32)]))
else:
    outs = {}
    combos = [("2", {}), ("2 no-epilogue", {"GM_NO_EPILOGUE": "1"}), ("2 hot_limit=0", {"GM_HOT_LIMIT": "0"}),
              ("2 h1 after h16", {"GM_DBG_SERIALIZE_H1": "1"})]
    for i, (k, extra) in enumerate(combos):
        f = "/tmp/pr_%d.npy" % i
        subprocess.check_call(["timeout", "100", sys.executable, __file__, f], env=dict(os.environ, GM_AUX_STREAMS=k.split()[0], **extra))
        outs[k] = np.load(f)
    from oracle import ref
    n, s, d, _ = ref.rmat_edges(20, 16, seed=1)
    rpr, rdeg, _, _ = ref.pagerank(n, s, d, None, threads=8, iterations=10)
    for k, a in outs.items():
        bad = np.nonzero(a[0] != rpr)[0]
        print("aux streams", k, ": pagerank mismatches", len(bad), "degree mismatches", int((a[1] != rdeg).sum()), bad[:8])
