"""GPU suite (pytest -m gpu): the CUDA path, called through the C ABI, against the oracle and the reference's
golden vectors.  Tolerance: north_star's |dlnL| < 1e-6*|lnL| is the contract; the fp64 kernels are held to 1e-10."""
import numpy as np
import pytest

from hyphy_b200 import synth, LikelihoodFunction, engine
from oracle import port
from tests import golden_cases as gc

pytestmark = pytest.mark.gpu

RTOL_CONTRACT = 1e-6      # north_star
RTOL_FP64 = 1e-10         # what the fp64 kernels actually have to deliver


@pytest.fixture(scope="module", autouse=True)
def _need_gpu(engine_lib):
    assert engine.device_count() > 0, "GPU tests need a CUDA device (no fallback exists)"


def _site_lnl(sl, ss):
    return np.log(sl) - 64.0 * np.log(2.0) * ss


@pytest.mark.parametrize("name", gc.SMALL + gc.MEDIUM + gc.FULL)
def test_lnl_matches_reference_golden(name):
    w, g = gc.load(name)
    lf = LikelihoodFunction(w)
    lf.set_all_matrices()
    lnl, sl, ss = lf.compute(want_sites=True)
    lf.close()
    assert abs(lnl - g["lnL"]) <= RTOL_FP64 * abs(g["lnL"]), (lnl, g["lnL"])
    site = _site_lnl(sl, ss)
    np.testing.assert_allclose(site[w.site_to_pattern], g["site_lnL"], rtol=0, atol=1e-8)
    # a checksum of checksums: frequency-weighted per-pattern values must re-add to lnL
    assert abs((site * w.pattern_freq).sum() - lnl) <= 1e-9 * abs(lnl)


@pytest.mark.parametrize("name", ["mg94_8x60_c4_ambig", "c1_hky85_8x500", "mg94_200x64_c4_scaling"])
def test_per_class_compute_block_matches_oracle(name):
    """ComputeBlock semantics: one rate class at a time with (L, scaler count) outputs, combined on the host the
    way PopulateConditionalProbabilities does (likefunc2.cpp:828-859) == fused device path == oracle."""
    w, g = gc.load(name)
    lf = LikelihoodFunction(w)
    lf.set_all_matrices()
    per_class = []
    for c in range(w.C):
        lnl_c, sl, ss = lf.compute_block(c, want_sites=True)
        P = np.stack([port.expm(w.Q_classes[c] * w.tree.t[b], w.D > 20) for b in range(w.tree.n_branches)])
        oL, oS = port.prune(w, P)
        np.testing.assert_allclose(_site_lnl(sl, ss), _site_lnl(oL, oS), rtol=0, atol=1e-9)
        assert np.all(sl > 2.0 ** -64) and np.all(sl <= 1.0)
        assert abs(lnl_c - (w.pattern_freq * _site_lnl(oL, oS)).sum()) <= RTOL_FP64 * abs(lnl_c)
        per_class.append(np.log(w.class_weights[c]) + _site_lnl(sl, ss))
    fused = lf.compute()
    lf.close()
    host_combined = (w.pattern_freq * np.logaddexp.reduce(np.stack(per_class), axis=0)).sum()
    assert abs(fused - host_combined) <= RTOL_FP64 * abs(fused)
    assert abs(fused - g["lnL"]) <= RTOL_FP64 * abs(g["lnL"])


def test_expm_matches_oracle():
    w, _ = gc.load("mg94_30x100_c4_ambig")
    lf = LikelihoodFunction(w)
    lf.set_all_matrices()
    for c in (0, 3):
        for b in (0, 7, w.tree.n_branches - 1):
            P = lf.part.read_transition(c, b)
            Po = port.expm(w.Q_classes[c] * w.tree.t[b], True)
            np.testing.assert_allclose(P, Po, rtol=0, atol=2e-14)
            np.testing.assert_allclose(P.sum(axis=1), 1.0, atol=1e-14)
    lf.close()


def test_expm_large_rates_and_zero():
    """Long branches need several squarings; a zero matrix must give the identity."""
    w = synth.codon_workload(6, 20, 1, seed=3)
    lf = LikelihoodFunction(w)
    Q = w.Q_classes[0]
    ts = np.array([0.0, 1e-6, 0.3, 2.0, 25.0, 400.0, 3.0, 0.01, 1.0])[: w.tree.n_branches]
    lf.part.set_matrices(0, np.arange(len(ts)), Q[None] * ts[:, None, None])
    for b, t in enumerate(ts):
        P = lf.part.read_transition(0, b)
        np.testing.assert_allclose(P, port.expm(Q * t, True), rtol=0, atol=5e-13 if t < 5 else 1e-10)
    np.testing.assert_array_equal(lf.part.read_transition(0, 0), np.eye(61))
    np.testing.assert_allclose(lf.part.read_transition(0, 5), np.tile(w.pi, (61, 1)), atol=1e-9)
    lf.close()


def test_host_transition_matrices_path():
    """HB2_MATRIX_TRANS: host-exponentiated P (GetCompExp()->theData) gives the same lnL."""
    w, g = gc.load("mg94_8x60_c1")
    lf = LikelihoodFunction(w)
    P = np.stack([port.expm(w.Q_classes[0] * w.tree.t[b], True) for b in range(w.tree.n_branches)])
    lf.part.set_matrices_ptrs(0, lf.all_nodes, list(P), engine.MATRIX_TRANS)
    lnl = lf.compute()
    lf.close()
    assert abs(lnl - g["lnL"]) <= 1e-12 * abs(g["lnL"])


def test_partial_update_equals_full_recompute():
    """DetermineNodesForUpdate semantics (tree.cpp:3117): change one branch, pass only that node; the engine must
    re-prune its ancestors and reuse every other cached conditional."""
    w, _ = gc.load("mg94_30x100_c4_ambig")
    lf = LikelihoodFunction(w)
    lf.set_all_matrices()
    base = lf.compute()
    rng = np.random.default_rng(1)
    for node in [0, 5, w.tree.n_leaves + 2, w.tree.n_branches - 1]:
        w.tree.t[node] *= 1.7
        Qt = w.Qt()
        for c in range(w.C):
            lf.part.set_matrices(c, [node], Qt[c, node:node + 1])
        got = lf.compute(update_nodes=[node])
        ref, _ = port.lnl(w)
        assert abs(got - ref) <= RTOL_FP64 * abs(ref)
        assert got != base
    # idempotence: nothing changed, empty update list -> same value, and a forced full recompute agrees
    again = lf.compute(update_nodes=[])
    full = lf.compute(update_nodes=None)
    assert again == got and abs(full - got) <= 1e-12 * abs(got)
    lf.close()


def test_root_frequencies_and_weights_only_change():
    w, _ = gc.load("mg94_8x60_c4_ambig")
    lf = LikelihoodFunction(w)
    lf.set_all_matrices()
    lf.compute()
    w2 = np.array([0.1, 0.2, 0.3, 0.4])
    got = lf.compute(update_nodes=[], weights=w2)
    ref, _ = port.lnl(w, weights=w2)
    lf.close()
    assert abs(got - ref) <= RTOL_FP64 * abs(ref)


def test_mixture_matrices_bsrel():
    """Explicit-form models P = sum_k w_k Exp(Q_k) per branch (tree.cpp:3047-3089)."""
    w = synth.codon_workload(10, 40, 1, seed=11)
    K = 3
    comps = [synth.mg94_rev_Q(om) for om in (0.1, 1.0, 4.0)]
    wk = np.array([0.6, 0.3, 0.1])
    nb = w.tree.n_branches
    M = np.stack([np.stack([Qk * w.tree.t[b] for Qk in comps]) for b in range(nb)])      # [nb, K, D, D]
    lf = LikelihoodFunction(w)
    lf.part.set_mixture_matrices(0, np.arange(nb), M, np.tile(wk, (nb, 1)))
    got = lf.compute()
    P = np.stack([sum(wk[k] * port.expm(M[b, k], True) for k in range(K)) for b in range(nb)])
    oL, oS = port.prune(w, P)
    ref = (w.pattern_freq * _site_lnl(oL, oS)).sum()
    np.testing.assert_allclose(lf.part.read_transition(0, 3), P[3], atol=1e-14)
    lf.close()
    assert abs(got - ref) <= RTOL_FP64 * abs(ref)


@pytest.mark.parametrize("D,taxa,sites,C", [(20, 12, 300, 1), (20, 40, 200, 4), (2, 9, 64, 1), (16, 7, 100, 2), (29, 6, 50, 1), (62, 9, 70, 1)])
def test_other_state_counts(D, taxa, sites, C):
    """Protein-sized (20), binary, dinucleotide (16) and non-61 codon tables (60-63 -> padded 64) state spaces."""
    w = synth.generic_workload(D, taxa, sites, C)
    lf = LikelihoodFunction(w)
    lf.set_all_matrices()
    got = lf.compute()
    lf.close()
    ref, _ = port.lnl(w, sparse_storage=False)
    assert abs(got - ref) <= RTOL_FP64 * abs(ref)


def test_conditionals_readback_matches_oracle():
    w, _ = gc.load("mg94_8x60_c1")
    lf = LikelihoodFunction(w)
    lf.set_all_matrices()
    lf.compute()
    P = np.stack([port.expm(w.Q_classes[0] * w.tree.t[b], True) for b in range(w.tree.n_branches)])
    _, _, ocond = port.prune(w, P, want_cond=True)
    for inode in range(w.tree.n_internal):
        cond, e = lf.part.read_conditionals(0, inode)
        np.testing.assert_allclose(cond * np.exp2(e)[:, None], ocond[inode], rtol=1e-12, atol=0)
        assert np.all(cond.max(axis=1) >= 0.5) and np.all(cond.max(axis=1) <= 1.0)
    lf.close()


def test_impossible_pattern_gives_minus_infinity():
    """tree_evaluator.cpp:4094-4112: a pattern with likelihood 0 makes the block -inf."""
    w = synth.nucleotide_workload(5, 30, seed=2)
    lf = LikelihoodFunction(w)
    P = np.tile(np.eye(4), (w.tree.n_branches, 1, 1))        # zero-length branches: differing leaves are impossible
    lf.part.set_matrices(0, lf.all_nodes, P, engine.MATRIX_TRANS)
    assert lf.compute() == -np.inf
    lf.close()


def test_nan_is_propagated():
    w = synth.nucleotide_workload(5, 30, seed=2)
    lf = LikelihoodFunction(w)
    Qt = w.Qt()
    Qt[0, 2, 1, 1] = np.nan
    lf.set_all_matrices(Qt)
    assert np.isnan(lf.compute())
    lf.close()


def test_errors_are_reported_not_swallowed():
    w = synth.nucleotide_workload(5, 30, seed=2)
    lf = LikelihoodFunction(w)
    with pytest.raises(engine.EngineError, match="no matrix was ever set"):
        lf.compute()
    with pytest.raises(engine.EngineError):
        lf.part.set_matrices(0, [99], np.zeros((1, 4, 4)))
    lf.close()
    with pytest.raises(engine.EngineError, match="not supported"):
        engine.Partition(4, 70, 3, 1, 1, [0, 0, 0, -1], np.zeros((3, 4), dtype=np.int64), None, np.ones(4, dtype=np.int64))


def test_full_size_properties():
    """At BASELINE.json's full size: golden lnL, linearity in pattern frequencies (doubling every frequency doubles
    lnL exactly up to rounding), and invariance under a permutation of the patterns."""
    w, g = gc.load("ns_mg94_200x2000_c4")
    lf = LikelihoodFunction(w)
    lf.set_all_matrices()
    a = lf.compute()
    lf.close()
    assert abs(a - g["lnL"]) <= RTOL_FP64 * abs(g["lnL"])
    perm = np.random.default_rng(0).permutation(w.S)
    w.leaf_states = np.ascontiguousarray(w.leaf_states[:, perm])
    w.pattern_freq = w.pattern_freq[perm] * 2
    lf = LikelihoodFunction(w)
    lf.set_all_matrices()
    b = lf.compute()
    lf.close()
    assert abs(b - 2 * a) <= 1e-11 * abs(b)
