"""Debug aid: repeats tests/test_gpu_parity.py::test_mixture_matrices_bsrel's evaluation and prints lnL and the worst branch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyphy_b200 import synth, LikelihoodFunction
from oracle import port

w = synth.codon_workload(10, 40, 1, seed=11)
comps = [synth.mg94_rev_Q(om) for om in (0.1, 1.0, 4.0)]
wk = np.array([0.6, 0.3, 0.1])
nb = w.tree.n_branches
M = np.stack([np.stack([Qk * w.tree.t[b] for Qk in comps]) for b in range(nb)])
P = np.stack([sum(wk[k] * port.expm(M[b, k], True) for k in range(3)) for b in range(nb)])
oL, oS = port.prune(w, P)
ref = (w.pattern_freq * (np.log(oL) - 64 * np.log(2.0) * oS)).sum()
for flags in (1, 0):
    for rep in range(6):
        lf = LikelihoodFunction(w, flags=flags)
        lf.part.set_mixture_matrices(0, np.arange(nb), M, np.tile(wk, (nb, 1)))
        got = lf.compute()
        errs = [np.abs(lf.part.read_transition(0, b) - P[b]).max() for b in range(nb)]
        lf.close()
        print(flags, rep, got, ref, got - ref, "worst branch", int(np.argmax(errs)), max(errs), flush=True)
