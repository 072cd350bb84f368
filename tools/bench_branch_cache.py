"""Single-branch shortcut on the north-star workload: time of one probe (new matrix for ONE branch in every class ->
lnL) through hb2_branch_cache_evaluate vs a partial re-evaluation (updateNodes = [branch]) vs a full evaluation.
Wall clock around the public API calls (host buffers), median of `reps`."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hyphy_b200 import synth, LikelihoodFunction

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
w = synth.codon_workload(200, 2000, 4)
Qt = w.Qt()
lf = LikelihoodFunction(w)
lf.set_all_matrices(Qt)
base = lf.compute()
t = w.tree
L, I = t.n_leaves, t.n_internal
par = np.asarray(t.flat_parents)
depth = np.zeros(L + I, dtype=int)
for n in range(L + I - 2, -1, -1):
    depth[n] = depth[L + par[n]] + 1
out = {"workload": "MG94xREV 200x2000, 4 classes", "lnL": base, "probes": {}}
for label, node in (("deepest_leaf", int(np.argmax(depth[:L]))), ("median_depth_leaf", int(np.argsort(depth[:L])[L // 2])), ("root_child", int(next(n for n in range(L + I - 1) if par[n] == I - 1)))):
    def set_branch(scale):
        for c in range(w.C):
            lf.part.set_matrices(c, [node], (Qt[c, node] * scale)[None])
    lf.part.branch_cache_build(node, w.pi)          # untimed: first call allocates the cache and loads the kernels
    tb = []
    for k in range(5):
        t0 = time.perf_counter(); lf.part.branch_cache_build(node, w.pi); tb.append(time.perf_counter() - t0)
    t_build = float(np.median(tb))
    tp, tu = [], []
    vals = []
    for k in range(reps):
        s = 1.0 + 0.01 * (k % 7)
        t0 = time.perf_counter(); set_branch(s); v = lf.part.branch_cache_evaluate(w.class_weights); tp.append(time.perf_counter() - t0)
        vals.append((s, v))
    for k in range(reps):
        s = 1.0 + 0.01 * (k % 7)
        t0 = time.perf_counter(); set_branch(s); v2 = lf.compute(update_nodes=[node]); tu.append(time.perf_counter() - t0)
        ref = dict(vals)[s]
        assert abs(v2 - ref) <= 2e-7 * abs(ref), (v2, ref)
    set_branch(1.0); lf.compute(update_nodes=[node])
    out["probes"][label] = {"node": node, "depth": int(depth[node]), "build_ms": t_build * 1e3, "probe_ms": float(np.median(tp)) * 1e3,
                            "partial_update_ms": float(np.median(tu)) * 1e3}
tf = []
for k in range(20):
    t0 = time.perf_counter(); lf.set_all_matrices(Qt); lf.compute(); tf.append(time.perf_counter() - t0)
out["full_evaluation_dense_matrices_ms"] = float(np.median(tf)) * 1e3
lf.close()
print(json.dumps(out))
