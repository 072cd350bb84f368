"""Nucleotide / protein path measurement (D=4 and D=20 are HBM-bandwidth bound): one full evaluation of a
large synthetic alignment through the single-launch small-state walk kernel, reported as achieved GB/s of algorithmic
traffic against the measured HBM peak.

    python tools/bench_small.py --states 4 --taxa 256 --sites 400000 [--iters 10]
Algorithmic bytes per evaluation and class (fp64 conditionals, each written once and read once, int32 exponents, int32
leaf codes):  (2I-1)*S*Dp*8 + (2I-1)*S*4 + L*S*4.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyphy_b200 import synth, LikelihoodFunction  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--states", type=int, default=4)
    ap.add_argument("--taxa", type=int, default=256)
    ap.add_argument("--sites", type=int, default=400000)
    ap.add_argument("--classes", type=int, default=1)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--check", action="store_true", help="compare lnL with the oracle (slow for big inputs)")
    args = ap.parse_args()
    t0 = time.time()
    if args.states == 4:
        w = synth.nucleotide_workload(args.taxa, args.sites, mean_t=0.1)
    else:
        w = synth.generic_workload(args.states, args.taxa, args.sites, args.classes, mean_t=0.1)
    gen_s = time.time() - t0
    lf = LikelihoodFunction(w)
    lf.set_template()
    lf.set_all_compiled()
    lnl = lf.compute()
    ms, stage, lnl2 = lf.part.time_resident(w.class_weights, w.pi, iters=3)
    ms, stage, lnl2 = lf.part.time_resident(w.class_weights, w.pi, iters=args.iters)
    L, I, S, C = w.tree.n_leaves, w.tree.n_internal, w.S, w.C
    Dp = {2: 4, 4: 4}.get(w.D, (w.D + 7) // 8 * 8)
    byts = C * ((2 * I - 1) * S * Dp * 8 + (2 * I - 1) * S * 4 + L * S * 4)
    flops = C * ((I - 1) * S * 2 * w.D * w.D + (L + L + I - 1) * S * w.D)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured"
    except Exception:
        peak, src = 6650.0, "fallback"
    out = {"workload": w.name, "states": w.D, "taxa": L, "patterns": S, "classes": C, "lnL": lnl, "ms_per_eval": ms,
           "stage_ms": {"expm": stage[0], "pruning": stage[1], "root": stage[2]}, "algorithmic_bytes": byts,
           "achieved_gbs_pruning": byts / (stage[1] * 1e-3) / 1e9, "hbm_peak_gbs": peak, "peak_source": src,
           "frac": byts / (stage[1] * 1e-3) / 1e9 / peak, "gflops_pruning": flops / (stage[1] * 1e-3) / 1e9, "generate_s": gen_s}
    if args.check:
        from oracle import port
        ref, _ = port.lnl(w)
        out["oracle_lnL"] = ref
        out["rel_err"] = abs(lnl - ref) / abs(ref)
    lf.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
