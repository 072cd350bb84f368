"""Seeded synthetic workloads for the likelihood hot path.

Everything here is *input generation* for tests and bench.py: a random unrooted binary tree in the
reference's flat layout, MG94xREV / HKY85 rate matrices in the reference's state order, sequences
evolved down the tree, and pattern compression.  No likelihood arithmetic lives here.

Reference conventions mirrored (so the same numbers can be fed to the reference binary):
  * nucleotide order ACGT; codon index 16*n1+4*n2+n3; Universal-code stops TAA/TAG/TGA removed,
    leaving 61 sense states in increasing index order (dataset_filter.cpp:1586-1614).
  * flat tree: leaves 0..L-1 in post-order, internals L..L+I-1 in post-order, root last;
    flatParents[node] = parent's *internal* index (0..I-1), root = -1 (tree.cpp:722-766).
  * leaf codes: >=0 state index, <0 -> -(k+1) indexes a 0/1 ambiguity vector k (likefunc.cpp:4299).
  * rate matrices row = from-state, column = to-state; the engine receives Q*t with the diagonal
    already set to minus the off-diagonal row sum (matrix.cpp:1546 MultByFreqs).
"""
from __future__ import annotations

import dataclasses
import numpy as np

NUC = "ACGT"
_AA = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"  # index 16a+4b+c, ACGT order
STOPS = [i for i, a in enumerate(_AA) if a == "*"]           # [48, 50, 56]
SENSE = [i for i in range(64) if i not in STOPS]              # 61 sense codons
SENSE_INDEX = {c: k for k, c in enumerate(SENSE)}


def codon_str(c64: int) -> str:
    return NUC[c64 >> 4] + NUC[(c64 >> 2) & 3] + NUC[c64 & 3]


# ----------------------------------------------------------------------------------------------
# tree
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class FlatTree:
    n_leaves: int
    n_internal: int
    flat_parents: np.ndarray       # int64 [L+I], reference layout (see module docstring)
    names: list                    # names of the L+I nodes in flat order
    t: np.ndarray                  # float64 [L+I]; branch parameter of each non-root node (root entry 0)
    newick: str                    # topology with named internal nodes, no lengths

    @property
    def n_branches(self) -> int:
        return self.n_leaves + self.n_internal - 1

    def children(self):
        L = self.n_leaves
        ch = [[] for _ in range(self.n_internal)]
        for node in range(L + self.n_internal - 1):
            ch[self.flat_parents[node]].append(node)
        return ch


def random_tree(n_leaves: int, seed: int, mean_t: float = 0.05) -> FlatTree:
    """Random unrooted binary topology (root trifurcating) by random sequential leaf attachment."""
    assert n_leaves >= 3
    rng = np.random.default_rng(seed)
    # nodes as dicts: children list; start with star of 3 leaves
    children = {0: [1, 2, 3]}
    parent = {1: 0, 2: 0, 3: 0}
    is_leaf = {0: False, 1: True, 2: True, 3: True}
    nxt = 4
    for _ in range(n_leaves - 3):
        # choose a random branch (identified by its child end) and split it
        cands = list(parent.keys())
        b = cands[rng.integers(len(cands))]
        p = parent[b]
        mid, leaf = nxt, nxt + 1
        nxt += 2
        children[p][children[p].index(b)] = mid
        children[mid] = [b, leaf]
        parent[mid] = p
        parent[b] = mid
        parent[leaf] = mid
        is_leaf[mid] = False
        is_leaf[leaf] = True
    # post-order traversal from root 0
    order = []
    stack = [(0, 0)]
    while stack:
        node, ci = stack.pop()
        if is_leaf[node]:
            order.append(node)
            continue
        if ci < len(children[node]):
            stack.append((node, ci + 1))
            stack.append((children[node][ci], 0))
        else:
            order.append(node)
    leaves = [n for n in order if is_leaf[n]]
    internals = [n for n in order if not is_leaf[n]]
    L, I = len(leaves), len(internals)
    assert L == n_leaves and I == n_leaves - 2
    flat_id = {}
    for k, n in enumerate(leaves):
        flat_id[n] = k
    for k, n in enumerate(internals):
        flat_id[n] = L + k
    flat_parents = np.full(L + I, -1, dtype=np.int64)
    for n in order:
        if n in parent:
            flat_parents[flat_id[n]] = flat_id[parent[n]] - L
    names = [None] * (L + I)
    for k in range(L):
        names[k] = f"T{k + 1}"
    for k in range(I):
        names[L + k] = f"N{k + 1}"

    def nwk(n):
        if is_leaf[n]:
            return names[flat_id[n]]
        inner = ",".join(nwk(c) for c in children[n])
        return f"({inner})" + ("" if n == 0 else names[flat_id[n]])

    newick = nwk(0)
    t = rng.exponential(mean_t, size=L + I)
    t = np.maximum(t, 1e-4)
    t[L + I - 1] = 0.0
    return FlatTree(L, I, flat_parents, names, t, newick)


# ----------------------------------------------------------------------------------------------
# substitution models (per unit branch parameter t)
# ----------------------------------------------------------------------------------------------
_THETA_IDX = {("A", "C"): 0, ("A", "G"): 1, ("A", "T"): 2, ("C", "G"): 3, ("C", "T"): 4, ("G", "T"): 5}
DEFAULT_THETA = np.array([0.25, 1.0, 0.25, 0.3, 1.1, 0.3])      # AC AG AT CG CT GT
DEFAULT_POSFREQ = np.array([[0.30, 0.20, 0.30, 0.20],             # CF3x4: position-specific ACGT
                            [0.32, 0.23, 0.17, 0.28],
                            [0.22, 0.27, 0.24, 0.27]])
DEFAULT_OMEGAS = np.array([0.05, 0.3, 1.0, 2.5])
DEFAULT_OMEGA_WEIGHTS = np.array([0.5, 0.3, 0.15, 0.05])


def theta_of(x: int, y: int, theta) -> float:
    a, b = (x, y) if x < y else (y, x)
    return float(theta[_THETA_IDX[(NUC[a], NUC[b])]])


def codon_frequencies(posfreq) -> np.ndarray:
    """CF3x4-style product frequencies renormalised over the 61 sense codons."""
    pi = np.array([posfreq[0][c >> 4] * posfreq[1][(c >> 2) & 3] * posfreq[2][c & 3] for c in SENSE])
    return pi / pi.sum()


def mg94_entries(theta, posfreq):
    """List of (i, j, theta_xy * pi^{pos}_y, is_nonsyn, theta_name) for all one-nucleotide codon changes (sense only)."""
    out = []
    for i, ci in enumerate(SENSE):
        for pos in range(3):
            shift = (4, 2, 0)[pos]
            x = (ci >> shift) & 3
            for y in range(4):
                if y == x:
                    continue
                cj = (ci & ~(3 << shift)) | (y << shift)
                if cj in STOPS:
                    continue
                j = SENSE_INDEX[cj]
                nonsyn = _AA[ci] != _AA[cj]
                a, b = (x, y) if x < y else (y, x)
                out.append((i, j, float(posfreq[pos][y]), nonsyn, NUC[a] + NUC[b]))
    return out


def mg94_rev_Q(omega: float, theta=DEFAULT_THETA, posfreq=DEFAULT_POSFREQ) -> np.ndarray:
    """MG94xREV rate matrix per unit t: q_ij = theta_xy * pi^{pos}_y * (omega if nonsynonymous)."""
    Q = np.zeros((61, 61))
    th = dict(zip(["AC", "AG", "AT", "CG", "CT", "GT"], theta))
    for i, j, pf, nonsyn, name in mg94_entries(theta, posfreq):
        Q[i, j] = th[name] * pf * (omega if nonsyn else 1.0)
    Q[np.diag_indices(61)] = -Q.sum(axis=1)
    return Q


def hky85_Q(kappa: float, freqs) -> np.ndarray:
    Q = np.zeros((4, 4))
    for i in range(4):
        for j in range(4):
            if i != j:
                transition = (i + j) in (2, 4) and abs(i - j) == 2     # A<->G (0,2), C<->T (1,3)
                Q[i, j] = (kappa if transition else 1.0) * freqs[j]
    Q[np.diag_indices(4)] = -Q.sum(axis=1)
    return Q


def gtr_like_Q(D: int, seed: int):
    """Random reversible D-state model (used for the 20-state protein-shaped cases)."""
    rng = np.random.default_rng(seed)
    pi = rng.dirichlet(np.full(D, 5.0))
    R = rng.gamma(1.0, 1.0, size=(D, D))
    R = (R + R.T) / 2
    Q = R * pi[None, :]
    Q[np.diag_indices(D)] = 0
    Q[np.diag_indices(D)] = -Q.sum(axis=1)
    scale = -(pi * np.diag(Q)).sum()
    return Q / scale, pi


def _expm_np(A: np.ndarray) -> np.ndarray:
    """Plain scaling-and-squaring Taylor expm, only used to SIMULATE sequences (not a likelihood path)."""
    n = A.shape[0]
    nrm = np.abs(A).sum(axis=1).max()
    s = max(0, int(np.ceil(np.log2(max(nrm, 1e-300)))) + 4)
    As = A / (2.0 ** s)
    R = np.eye(n)
    T = np.eye(n)
    for k in range(1, 20):
        T = T @ As / k
        R = R + T
    for _ in range(s):
        R = R @ R
    return R


# ----------------------------------------------------------------------------------------------
# alignments
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Workload:
    name: str
    tree: FlatTree
    D: int
    pi: np.ndarray                 # root/equilibrium frequencies [D]
    Q_classes: list                # C matrices [D,D], per unit t
    class_weights: np.ndarray      # [C]
    leaf_states: np.ndarray        # int64 [L, S] over unique patterns; <0 = -(k+1) ambiguity
    ambig: np.ndarray              # float64 [nAmb, D] of 0/1
    pattern_freq: np.ndarray       # int64 [S]
    site_to_pattern: np.ndarray    # int64 [sites]
    site_chars: list | None = None  # per leaf, the character string (for FASTA export to the reference)
    meta: dict = dataclasses.field(default_factory=dict)
    _tmpl: tuple | None = None
    # explicit-form mixture (BS-REL / BUSTED shape, reference tree.cpp:3047-3089): every branch's transition matrix is
    # sum_k mix_weights[k] * Exp(mix_Q[k] * t_b); None for ordinary models
    mix_Q: list | None = None
    mix_weights: np.ndarray | None = None

    @property
    def S(self):
        return self.leaf_states.shape[1]

    @property
    def C(self):
        return len(self.Q_classes)

    # -- compact ("compiled") form of the same matrices: the static template and per-evaluation formula values,
    #    the shape in which the reference itself holds a model matrix (_CompiledMatrixData, matrix.h:69-80)
    def _formulas(self):
        """Returns (entry_index[nnz], entry_formula[nnz], per-class base value of each formula [C, nF], col_freq|None);
        the formula VALUE for branch b is base * t_b."""
        if getattr(self, "_tmpl", None) is not None:
            return self._tmpl
        kind = self.meta.get("kind")
        D = self.D
        if kind == "codon":
            th = dict(zip(["AC", "AG", "AT", "CG", "CT", "GT"], self.meta["theta"]))
            keys, ei, ef = {}, [], []
            for i, j, pf, nonsyn, nm in mg94_entries(self.meta["theta"], np.array(self.meta["posfreq"])):
                key = (nm, nonsyn, pf)
                if key not in keys:
                    keys[key] = len(keys)
                ei.append(i * D + j)
                ef.append(keys[key])
            omegas = self.meta["omegas"]
            base = np.array([[th[nm] * pf * (om if nonsyn else 1.0) for (nm, nonsyn, pf) in keys] for om in omegas])
            cf = None
        elif kind == "nuc":
            ei, ef = [], []
            for i in range(4):
                for j in range(4):
                    if i != j:
                        ei.append(i * 4 + j)
                        ef.append(1 if ((i + j) in (2, 4) and abs(i - j) == 2) else 0)
            base = np.array([[1.0, self.meta["kappa"]]])
            cf = np.asarray(self.pi, dtype=np.float64)
        else:                                   # generic: every off-diagonal entry is its own formula
            ei = [i * D + j for i in range(D) for j in range(D) if i != j]
            ef = list(range(len(ei)))
            base = np.array([[Q[i, j] for i in range(D) for j in range(D) if i != j] for Q in self.Q_classes])
            cf = None
        self._tmpl = (np.array(ei, dtype=np.int64), np.array(ef, dtype=np.int64), base, cf)
        return self._tmpl

    def compiled_template(self):
        ei, ef, base, cf = self._formulas()
        return ei, ef, base.shape[1], cf

    def compiled_values(self, perturb: float = 0.0) -> np.ndarray:
        """Formula values [C, B, nF] (what the host evaluates per branch: cmd->formulaValues, matrix.cpp:3131)."""
        _, _, base, _ = self._formulas()
        nb = self.tree.n_branches
        return base[:, None, :] * (self.tree.t[:nb, None] * (1.0 + perturb))[None, :, :]

    def mixture_Qt(self, perturb: float = 0.0):
        """Explicit-form mixture inputs: (M [B, K, D, D] component rate matrices Q_k * t_b, weights [B, K])."""
        assert self.mix_Q is not None
        nb = self.tree.n_branches
        t = self.tree.t[:nb] * (1.0 + perturb)
        M = np.stack([np.stack([Qk * tb for Qk in self.mix_Q]) for tb in t])
        return M, np.tile(np.asarray(self.mix_weights, dtype=np.float64), (nb, 1))

    def Qt(self, perturb: float = 0.0) -> np.ndarray:
        """Dense Q*t for every (class, branch node): float64 [C, L+I-1, D, D].  `perturb` scales all
        off-diagonal rates by (1+perturb) -- stands in for 'one global parameter moved', so that every
        matrix must be re-exponentiated (SURVEY §8d evaluation stream)."""
        nb = self.tree.n_branches
        out = np.empty((self.C, nb, self.D, self.D))
        for c, Q in enumerate(self.Q_classes):
            out[c] = Q[None, :, :] * (self.tree.t[:nb, None, None] * (1.0 + perturb))
        return out


def _simulate(tree: FlatTree, Ps_by_class, pi, site_class, rng) -> np.ndarray:
    """Evolve states down the tree.  Ps_by_class[c][node] = transition matrix of the branch above node."""
    L, I = tree.n_leaves, tree.n_internal
    n_sites = len(site_class)
    D = len(pi)
    state = np.zeros((L + I, n_sites), dtype=np.int64)
    root = L + I - 1
    state[root] = rng.choice(D, size=n_sites, p=pi)
    # parents come after children in flat order, so walk internals in reverse post-order
    order = list(range(L + I - 2, -1, -1))
    order.sort(key=lambda n: -(n if n >= L else -1))  # internals (descending) first, leaves last
    for node in order:
        par = tree.flat_parents[node] + L
        for c in range(len(Ps_by_class)):
            idx = np.nonzero(site_class == c)[0]
            if idx.size == 0:
                continue
            P = Ps_by_class[c][node]
            cdf = np.cumsum(P, axis=1)
            cdf /= cdf[:, -1:]
            u = rng.random(idx.size)
            ps = state[par, idx]
            state[node, idx] = (u[:, None] > cdf[ps]).sum(axis=1)
    return state[:L]


def compress(leaf_cols: np.ndarray):
    """Unique alignment columns (patterns), their multiplicities and the site->pattern map."""
    cols = np.ascontiguousarray(leaf_cols.T)
    uniq, inverse, counts = np.unique(cols, axis=0, return_inverse=True, return_counts=True)
    # keep first-appearance order (closer to the reference's pattern ids)
    first = np.full(len(uniq), len(cols), dtype=np.int64)
    np.minimum.at(first, inverse, np.arange(len(cols)))
    perm = np.argsort(first)
    rank = np.empty_like(perm)
    rank[perm] = np.arange(len(perm))
    return uniq[perm].T.copy(), counts[perm].astype(np.int64), rank[inverse].astype(np.int64)


def codon_workload(n_taxa: int, n_codons: int, n_classes: int = 1, seed: int = 20260924,
                   ambig_frac: float = 0.0, mean_t: float = 0.05, name: str | None = None) -> Workload:
    """MG94xREV (CF3x4) alignment; sites drawn from 4 omega classes; the *model* has `n_classes`
    omega categories (1 -> single omega 0.3; 4 -> the generating mixture)."""
    rng = np.random.default_rng(seed + 7919 * n_taxa + n_codons)
    tree = random_tree(n_taxa, seed + n_taxa, mean_t)
    pi = codon_frequencies(DEFAULT_POSFREQ)
    gen_Q = [mg94_rev_Q(w) for w in DEFAULT_OMEGAS]
    # normalise so mean_t is expected substitutions per codon under the mixture
    rate = sum(wt * -(pi * np.diag(Q)).sum() for wt, Q in zip(DEFAULT_OMEGA_WEIGHTS, gen_Q))
    tree.t[:tree.n_branches] /= rate
    site_class = rng.choice(4, size=n_codons, p=DEFAULT_OMEGA_WEIGHTS)
    nb = tree.n_branches
    Ps = [[_expm_np(Q * tree.t[b]) for b in range(nb)] for Q in gen_Q]
    states = _simulate(tree, Ps, pi, site_class, rng)                # [L, sites] sense-codon indices
    # characters + optional ambiguity injection
    chars = [[codon_str(SENSE[s]) for s in row] for row in states]
    codes = states.copy()
    amb_rows: list = []
    amb_index: dict = {}
    if ambig_frac > 0:
        n_amb = int(round(ambig_frac * states.size))
        cells = rng.choice(states.size, size=n_amb, replace=False)
        for cell in cells:
            l, s = divmod(int(cell), n_codons)
            kind = rng.integers(3)
            c3 = list(chars[l][s])
            if kind == 0:
                c3 = list("---")
            elif kind == 1:
                c3[rng.integers(3)] = "N"
            else:
                p = rng.integers(3)
                c3[p] = "R" if c3[p] in "AG" else "Y"
            chars[l][s] = "".join(c3)
            vec = resolve_codon("".join(c3))
            if vec.sum() == 1:
                codes[l, s] = int(np.argmax(vec))
            else:
                key = vec.tobytes()
                if key not in amb_index:
                    amb_index[key] = len(amb_rows)
                    amb_rows.append(vec)
                codes[l, s] = -(amb_index[key] + 1)
    leaf_states, freq, s2p = compress(codes)
    if n_classes == 1:
        Qc, w = [mg94_rev_Q(0.3)], np.array([1.0])
    else:
        assert n_classes == 4
        Qc, w = gen_Q, DEFAULT_OMEGA_WEIGHTS.copy()
    ambig = np.array(amb_rows, dtype=np.float64).reshape(len(amb_rows), 61)
    return Workload(name or f"mg94_{n_taxa}x{n_codons}_c{n_classes}", tree, 61, pi, Qc, w, leaf_states, ambig, freq, s2p,
                    ["".join(r) for r in chars],
                    {"kind": "codon", "theta": DEFAULT_THETA.tolist(), "posfreq": DEFAULT_POSFREQ.tolist(),
                     "omegas": (DEFAULT_OMEGAS.tolist() if n_classes == 4 else [0.3]), "seed": seed})


def bsrel_workload(n_taxa: int, n_codons: int, omegas=(0.1, 1.0, 4.0), weights=(0.6, 0.3, 0.1), seed: int = 20260924,
                   mean_t: float = 0.05, name: str | None = None) -> Workload:
    """BUSTED / BS-REL shaped model (config c3): ONE rate class whose per-branch transition matrix is the explicit-form
    mixture sum_k w_k Exp(Q(omega_k) t_b) (random-effects over omega on every branch).  The alignment is the codon_workload
    alignment of the same size and seed; only the MODEL differs."""
    w = codon_workload(n_taxa, n_codons, 1, seed=seed, mean_t=mean_t)
    w.name = name or f"bsrel_{n_taxa}x{n_codons}_k{len(omegas)}"
    w.mix_Q = [mg94_rev_Q(om) for om in omegas]
    w.mix_weights = np.asarray(weights, dtype=np.float64)
    w.Q_classes = [sum(wt * Q for wt, Q in zip(weights, w.mix_Q))]     # placeholder (never used for likelihoods of this model)
    w.meta = dict(w.meta, mixture={"omegas": list(map(float, omegas)), "weights": list(map(float, weights))})
    return w


_IUPAC = {"A": "A", "C": "C", "G": "G", "T": "T", "R": "AG", "Y": "CT", "N": "ACGT", "-": "ACGT", "?": "ACGT"}


def resolve_codon(c3: str) -> np.ndarray:
    """0/1 vector over the 61 sense codons compatible with an (ambiguous) codon string
    (dataset_filter.cpp:1594-1632 Translate2Frequencies: product of per-position sets, stops dropped,
    all-ones if nothing is compatible)."""
    vec = np.zeros(61)
    for a in _IUPAC[c3[0]]:
        for b in _IUPAC[c3[1]]:
            for c in _IUPAC[c3[2]]:
                idx = 16 * NUC.index(a) + 4 * NUC.index(b) + NUC.index(c)
                if idx in SENSE_INDEX:
                    vec[SENSE_INDEX[idx]] = 1.0
    if vec.sum() == 0:
        vec[:] = 1.0
    return vec


def nucleotide_workload(n_taxa: int, n_sites: int, seed: int = 20260924, kappa: float = 2.0,
                        ambig_frac: float = 0.0, mean_t: float = 0.05, name: str | None = None) -> Workload:
    rng = np.random.default_rng(seed + 104729 * n_taxa + n_sites)
    tree = random_tree(n_taxa, seed + n_taxa, mean_t)
    pi = np.array([0.30, 0.22, 0.24, 0.24])
    Q = hky85_Q(kappa, pi)
    nb = tree.n_branches
    Ps = [[_expm_np(Q * tree.t[b]) for b in range(nb)]]
    states = _simulate(tree, Ps, pi, np.zeros(n_sites, dtype=np.int64), rng)
    chars = [[NUC[s] for s in row] for row in states]
    codes = states.copy()
    amb_rows, amb_index = [], {}
    if ambig_frac > 0:
        cells = rng.choice(states.size, size=int(round(ambig_frac * states.size)), replace=False)
        for cell in cells:
            l, s = divmod(int(cell), n_sites)
            ch = ["-", "N", "R" if chars[l][s] in "AG" else "Y"][rng.integers(3)]
            chars[l][s] = ch
            vec = np.array([1.0 if n in _IUPAC[ch] else 0.0 for n in NUC])
            key = vec.tobytes()
            if key not in amb_index:
                amb_index[key] = len(amb_rows)
                amb_rows.append(vec)
            codes[l, s] = -(amb_index[key] + 1)
    leaf_states, freq, s2p = compress(codes)
    ambig = np.array(amb_rows, dtype=np.float64).reshape(len(amb_rows), 4)
    return Workload(name or f"hky85_{n_taxa}x{n_sites}", tree, 4, pi, [Q], np.array([1.0]), leaf_states, ambig, freq, s2p,
                    ["".join(r) for r in chars], {"kind": "nuc", "kappa": kappa, "freqs": pi.tolist(), "seed": seed})


def generic_workload(D: int, n_taxa: int, n_sites: int, n_classes: int = 1, seed: int = 20260924,
                     mean_t: float = 0.05, name: str | None = None) -> Workload:
    """Random reversible D-state model (D=20 stands in for the protein path); classes are rate multipliers."""
    rng = np.random.default_rng(seed + 31 * D + 1009 * n_taxa + n_sites)
    tree = random_tree(n_taxa, seed + n_taxa, mean_t)
    Q, pi = gtr_like_Q(D, seed + D)
    rates = np.array([1.0]) if n_classes == 1 else np.linspace(0.25, 2.5, n_classes)
    w = np.full(n_classes, 1.0 / n_classes)
    rates = rates / (rates * w).sum()
    Qc = [Q * r for r in rates]
    nb = tree.n_branches
    Ps = [[_expm_np(Qk * tree.t[b]) for b in range(nb)] for Qk in Qc]
    states = _simulate(tree, Ps, pi, rng.choice(n_classes, size=n_sites), rng)
    leaf_states, freq, s2p = compress(states)
    return Workload(name or f"rev{D}_{n_taxa}x{n_sites}_c{n_classes}", tree, D, pi, Qc, w, leaf_states,
                    np.zeros((0, D)), freq, s2p, None, {"kind": "generic", "seed": seed})
