"""CPU suite: the C-ABI library builds, loads and exports every symbol include/hyphy_b200.h declares;
without a GPU it fails loudly instead of falling back."""
import os
import re

import pytest

from hyphy_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hyphy_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(engine_lib):
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(engine_lib, name), f"{name} declared in include/hyphy_b200.h but not exported"
    assert sorted(n for n, _, _ in engine.ABI) == declared


def test_abi_version(engine_lib):
    assert engine_lib.hb2_abi_version() == 1


def test_no_cpu_fallback(engine_lib):
    """No GPU => create must fail with an explicit message (the host treats it as fatal)."""
    if engine.device_count() > 0:
        pytest.skip("a GPU is visible here")
    from hyphy_b200 import synth, LikelihoodFunction, EngineError
    with pytest.raises(EngineError, match="no CPU fallback"):
        LikelihoodFunction(synth.nucleotide_workload(4, 10))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under hyphy_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hyphy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("SURVEY", ""), f"{f} mentions the oracle"
