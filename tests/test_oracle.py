"""CPU suite: the oracle (oracle/hb2_oracle.c) against the reference's golden vectors, and its own invariants."""
import numpy as np
import pytest

from hyphy_b200 import synth
from oracle import port
from tests import golden_cases as gc


@pytest.mark.parametrize("name", gc.SMALL + gc.MEDIUM)
def test_oracle_matches_reference_golden(name):
    w, g = gc.load(name)
    lnl, site = port.lnl(w)
    assert abs(lnl - g["lnL"]) <= 1e-11 * abs(g["lnL"])
    np.testing.assert_allclose(site[w.site_to_pattern], g["site_lnL"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("name", gc.MIXTURE_SMALL + gc.MIXTURE_FULL)
def test_oracle_mixture_matches_reference_golden(name):
    """Explicit-form models (P_b = sum_k w_k Exp(Q_k t_b), reference tree.cpp:3047-3089): the fixture was produced by the
    unmodified binary from `Model M = ("Exp(Q1)*bw1+...", freqs, EXPLICIT_FORM_MATRIX_EXPONENTIAL)`."""
    w, g = gc.load(name)
    lnl, site = port.lnl_mixture(w)
    assert abs(lnl - g["lnL"]) <= 1e-11 * abs(g["lnL"])
    np.testing.assert_allclose(site[w.site_to_pattern], g["site_lnL"], rtol=0, atol=1e-10)


def test_oracle_forced_states_marginalise():
    """setBranch / setBranchTo restatement (tree_evaluator.cpp:3624,173-181,585-592,4059): pinning a node to each state in
    turn and summing recovers the unpinned likelihood -- for a leaf with an ambiguous observation, an internal node and
    the root; pinning a leaf to its observed state changes nothing."""
    w, _ = gc.load("mg94_8x60_c4_ambig")
    L, I = w.tree.n_leaves, w.tree.n_internal
    P = np.stack([port.expm(q, True) for q in w.Qt()[1]])
    sl, ss = port.prune(w, P)
    base = sl * 2.0 ** (-64.0 * ss)
    for node in (L + 1, L + I - 1):
        tot = np.zeros(w.S)
        for st in range(w.D):
            a, b = port.prune_forced(w, P, node, np.full(w.S, st))
            tot += a * 2.0 ** (-64.0 * b)
        np.testing.assert_allclose(tot, base, rtol=1e-12)
    leaf = int(np.argmax((w.leaf_states < 0).sum(axis=1)))             # the leaf with most ambiguities
    tot = np.zeros(w.S)
    for st in range(w.D):
        a, b = port.prune_forced(w, P, leaf, np.full(w.S, st))
        allowed = np.array([w.ambig[-c - 1][st] if c < 0 else float(c == st) for c in w.leaf_states[leaf]])
        tot += allowed * a * 2.0 ** (-64.0 * b)
    np.testing.assert_allclose(tot, base, rtol=1e-12)


def test_oracle_expm_is_a_transition_matrix_and_semigroup():
    Q = synth.mg94_rev_Q(0.7)
    P1 = port.expm(Q * 0.05, sparse_storage=True)
    P2 = port.expm(Q * 0.10, sparse_storage=True)
    assert np.all(P1 >= -1e-18)
    np.testing.assert_allclose(P1.sum(axis=1), 1.0, atol=1e-14)
    np.testing.assert_allclose(P1 @ P1, P2, atol=1e-13)
    # dense and sparse storage scalings agree to rounding
    np.testing.assert_allclose(port.expm(Q * 0.05, sparse_storage=False), P1, atol=1e-14)


def test_oracle_expm_zero_and_large():
    Z = np.zeros((4, 4))
    np.testing.assert_array_equal(port.expm(Z), np.eye(4))
    Q = synth.hky85_Q(2.0, [0.3, 0.22, 0.24, 0.24])
    Pinf = port.expm(Q * 500.0)
    np.testing.assert_allclose(Pinf, np.tile([0.3, 0.22, 0.24, 0.24], (4, 1)), atol=1e-9)


def test_oracle_scaler_convention():
    """Deep tree: scaling must trigger and (L, count) must satisfy L_true = L * 2^(-64 count)."""
    w, g = gc.load("mg94_200x64_c4_scaling")
    P = np.stack([port.expm(w.Q_classes[0] * w.tree.t[b], True) for b in range(w.tree.n_branches)])
    L, cnt = port.prune(w, P)
    assert cnt.max() >= 1
    assert np.all(L > 0) and np.all(L < 2.0 ** 64)


def test_oracle_ambiguity_all_ones_equals_pruned_leaf():
    """A fully missing leaf (all-ones vector) must not change the site likelihood versus any resolution sum."""
    w = synth.nucleotide_workload(6, 40, seed=5)
    P = np.stack([port.expm(w.Q_classes[0] * w.tree.t[b]) for b in range(w.tree.n_branches)])
    base, _ = port.prune(w, P)
    w2 = synth.nucleotide_workload(6, 40, seed=5)
    w2.ambig = np.ones((1, 4))
    total = np.zeros(w.S)
    for st in range(4):
        w3 = synth.nucleotide_workload(6, 40, seed=5)
        w3.leaf_states = w.leaf_states.copy()
        w3.leaf_states[0, :] = st
        total += port.prune(w3, P)[0]
    w2.leaf_states = w.leaf_states.copy()
    w2.leaf_states[0, :] = -1
    amb, _ = port.prune(w2, P)
    np.testing.assert_allclose(amb, total, rtol=1e-13)


def test_oracle_reproduces_reference_own_golden_smallcodon():
    """SmallCodon.bf:37 expects -3189.516375 +- 2*OPTIMIZATION_PRECISION (0.002); at the parameters the reference binary
    itself fitted, the oracle must give the reference's lnL (fixed-parameter parity on real HIV-1 RT data)."""
    w, Qt, g = gc.load_smallcodon()
    lnl, _ = port.lnl(w, Qt=Qt)
    assert abs(lnl - g["lnL_reference_run"]) <= 1e-11 * abs(lnl)
    assert abs(lnl - g["lnL_golden"]) < 0.002


def test_single_branch_factorisation_identity():
    """The algebra behind hb2_branch_cache_* (DESIGN 4.6), on the CPU with the oracle's matrices: for every branch b,
    L_s = sum_a rest_b[s][a] * (P_b below_b[s])[a] with rest propagated from the root through TRANSPOSED matrices
    (no re-rooting, no reversibility).  Plain numpy restatement; small cases only."""
    import numpy as np
    from tests import golden_cases as gc
    from oracle import port
    for name in ("mg94_8x60_c4_ambig", "c1_hky85_8x500"):
        w, _ = gc.load(name)
        t = w.tree
        L, I, D, S = t.n_leaves, t.n_internal, w.D, min(w.S, 40)
        par = np.asarray(t.flat_parents)
        Qt = w.Qt()
        c = Qt.shape[0] - 1
        P = np.stack([port.expm(Qt[c, b]) for b in range(L + I - 1)])          # P[b][parent state][child state]
        children = [[] for _ in range(I)]
        for n in range(L + I - 1):
            children[par[n]].append(n)

        def leafvec(n):
            v = np.zeros((S, D))
            for s in range(S):
                code = w.leaf_states[n, s]
                if code >= 0:
                    v[s, code] = 1.0
                else:
                    v[s] = w.ambig[-code - 1]
            return v

        cond = [None] * I
        below = lambda n: leafvec(n) if n < L else cond[n - L]
        for i in range(I):
            v = np.ones((S, D))
            for ch in children[i]:
                v *= below(ch) @ P[ch].T
            cond[i] = v
        full = cond[I - 1] @ w.pi
        for b in range(L + I - 1):
            path = []
            u = par[b]
            while u >= 0:
                path.append(u)
                u = par[L + u]
            out = np.tile(w.pi, (S, 1))
            rest = None
            for i, u in enumerate(reversed(path)):
                on_path = L + path[len(path) - 2 - i] if i + 1 < len(path) else b
                rest = out.copy()
                for ch in children[u]:
                    if ch != on_path:
                        rest *= below(ch) @ P[ch].T
                if i + 1 < len(path):
                    out = rest @ P[on_path]                                   # out'[a'] = sum_a rest[a] P[a][a']
            got = np.sum(rest * (below(b) @ P[b].T), axis=1)
            assert np.abs(got / full - 1.0).max() < 1e-12, (name, b)
