"""Golden fixtures: outputs of the UNMODIFIED reference binary on seeded workloads (tools/make_golden.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden import CASES, checksum  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
_cache = {}


def load(name):
    """Returns (workload, golden dict) and proves the regenerated input is the one the reference saw."""
    if name not in _cache:
        w = CASES[name]()
        g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        assert str(g["checksum"]) == checksum(w), f"synthetic workload {name} no longer matches its golden fixture"
        _cache[name] = (w, {"lnL": float(g["lnL"]), "site_lnL": g["site_lnL"]})
    return _cache[name]


SMALL = ["c1_hky85_8x500", "nuc_300x200_scaling", "mg94_8x60_c1", "mg94_8x60_c4_ambig", "mg94_30x100_c4_ambig",
         "mg94_200x64_c4_scaling"]
MEDIUM = ["c2_mg94_50x1000_c1"]
FULL = ["ns_mg94_200x2000_c4"]
# explicit-form mixtures (BS-REL / BUSTED shape; the workload carries mix_Q / mix_weights): config c3 and a small sibling
MIXTURE_SMALL = ["bsrel_12x80_k3"]
MIXTURE_FULL = ["c3_bsrel_100x1500_k3"]
# config c5 (aBSREL size: 500 taxa x 5000 codons x 4 classes, 5.2 GB of conditionals): GPU against the reference's output
# directly -- the scalar oracle would need minutes for it, so the CPU suite pins the oracle on the smaller cases only
HUGE = ["c5_mg94_500x5000_c4"]


def load_smallcodon():
    """The reference's own test tests/hbltests/SimpleOptimizations/SmallCodon.bf (golden lnL at :37) at the parameter
    values the reference binary fitted (tools/make_smallcodon_fixture.py).  Returns (workload, Qt[1,B,61,61], fixture)."""
    from hyphy_b200 import synth
    g = np.load(os.path.join(GOLDEN_DIR, "smallcodon_fit.npz"))
    L = int(g["n_leaves"])
    fp = g["flat_parents"]
    I = len(fp) - L
    tree = synth.FlatTree(L, I, fp, [f"n{k}" for k in range(L + I)], np.zeros(L + I), "")
    w = synth.Workload("smallcodon_fit", tree, 61, g["pi"], [np.eye(61)], np.array([1.0]), g["leaf_states"], g["ambig"].reshape(-1, 61),
                       g["pattern_freq"], np.zeros(0, dtype=np.int64))
    return w, g["Qt"][None], {"lnL_reference_run": float(g["lnL_reference_run"]), "lnL_golden": float(g["lnL_golden"])}
