"""Bring-up aid: per-step clock stamps of one walk-kernel CTA (HB2_WALK_TRACE) on the north-star workload."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cta = sys.argv[1] if len(sys.argv) > 1 else "0"
os.environ["HB2_WALK_TRACE"] = os.path.join(ROOT, "gpurun_out", f"walk_trace_cta{cta}.txt")
os.environ["HB2_WALK_TRACE_CTA"] = cta
from hyphy_b200 import synth, LikelihoodFunction  # noqa: E402

w = synth.codon_workload(200, 2000, 4)
from hyphy_b200.engine import FLAG_DEFAULT, FLAG_FORCE_FP64  # noqa: E402
lf = LikelihoodFunction(w, flags=FLAG_FORCE_FP64 if os.environ.get("HB2_TRACE_FP64") == "1" else FLAG_DEFAULT)    # HB2_TRACE_FP64=1: the fp64 lanes kernel
lf.set_template()
for k in range(3):
    lf.set_all_compiled(w.compiled_values(perturb=1e-4 * k))
    print(lf.compute())
lf.close()
