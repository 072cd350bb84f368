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
