"""Race hunt: evaluates the north-star workload repeatedly (fresh partition each round, and repeated full evaluations
within a partition) and requires bit-identical lnL and per-pattern results; then a pattern permutation with doubled
frequencies must give exactly twice the value up to fp64 summation order.  Usage: python tools/stress_determinism.py [rounds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hyphy_b200 import engine
from tests import golden_cases as gc

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
w, g = gc.load("ns_mg94_200x2000_c4")
Qt = w.Qt()

def run(wl, reps):
    lf = engine.LikelihoodFunction(wl)
    lf.set_all_matrices(Qt)
    out = []
    for _ in range(reps):
        lnl, site, scc = lf.compute(want_sites=True)
        out.append((lnl, site.copy(), scc.copy()))
    lf.close()
    return out

base = None
bad = 0
for r in range(rounds):
    res = run(w, 12)
    for k, (lnl, site, scc) in enumerate(res):
        if base is None:
            base = (lnl, site, scc)
            print("base lnL", repr(lnl), "golden rel", abs(lnl - g["lnL"]) / abs(g["lnL"]))
            continue
        if lnl != base[0] or not np.array_equal(site, base[1]) or not np.array_equal(scc, base[2]):
            d = np.nonzero((site != base[1]) | (scc != base[2]))[0]
            rel = np.abs(site[d] / base[1][d] - 1) if len(d) else []
            print(f"MISMATCH round {r} rep {k}: lnL diff {lnl - base[0]:.3e}; {len(d)} patterns differ, idx {d[:10]}, rel {np.asarray(rel)[:10]}")
            bad += 1
print("identical-input mismatches:", bad)
# permutation
w2 = gc.CASES["ns_mg94_200x2000_c4"]()
for seed in range(rounds):
    perm = np.random.default_rng(seed).permutation(w.S)
    w2.leaf_states = np.ascontiguousarray(w.leaf_states[:, perm])
    w2.pattern_freq = w.pattern_freq[perm] * 2
    res = run(w2, 3)
    for k, (lnl, site, scc) in enumerate(res):
        same = np.array_equal(site, base[1][perm]) and np.array_equal(scc, base[2][perm])
        if not same:
            d = np.nonzero((site != base[1][perm]) | (scc != base[2][perm]))[0]
            print(f"PERM MISMATCH seed {seed} rep {k}: lnL-2a {lnl - 2 * base[0]:.3e}; {len(d)} patterns differ: positions {d[:10]} orig {perm[d][:10]} rel {np.abs(site[d]/base[1][perm][d]-1)[:10]}")
            bad += 1
print("total mismatches:", bad)
sys.exit(1 if bad else 0)
