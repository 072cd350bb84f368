"""Pins the oracle against one of the reference's OWN golden vectors:
tests/hbltests/SimpleOptimizations/SmallCodon.bf:37 expects lnL = -3189.516375 (+-0.002) after optimisation.

The script runs the unmodified reference binary (oracle/_ref/hyphy) on a scratch copy of that test with a few dump lines
appended (fitted global parameters, every branch's `synRate`, the lnL it reached), rebuilds the numeric rate matrices
from the formulas in the test file at those fitted values, and stores everything the engine/oracle need in
tests/golden/smallcodon_fit.npz.  Nothing from the reference is copied into the repository: the fixture holds numbers
(the alignment as integer codon states, the tree as a parent array, Q*t per branch).
Run here (needs /root/reference):  python tools/make_smallcodon_fixture.py
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyphy_b200 import synth            # noqa: E402

REF = "/root/reference/tests/hbltests"
BIN = os.path.join(ROOT, "oracle", "_ref", "hyphy")
DUMP = r'''
fprintf (stdout, "\nHB2DUMP lnL ", Format (res[1][0], 30, 16), "\n");
fprintf (stdout, "HB2DUMP AC ", Format (AC, 30, 16), "\n");
fprintf (stdout, "HB2DUMP AT ", Format (AT, 30, 16), "\n");
fprintf (stdout, "HB2DUMP CT ", Format (CT, 30, 16), "\n");
fprintf (stdout, "HB2DUMP R ", Format (R, 30, 16), "\n");
hb2bn = BranchName (givenTree, -1);
for (hb2k = 0; hb2k < Columns (hb2bn) - 1; hb2k += 1) {
    ExecuteCommands ("hb2v = givenTree." + hb2bn[hb2k] + ".synRate;");
    fprintf (stdout, "HB2DUMP BRANCH ", hb2bn[hb2k], " ", Format (hb2v, 30, 16), "\n");
}
'''


def parse_newick(s):
    """Returns (children dict, names dict, root id)."""
    s = s.strip().rstrip(";")
    pos = [0]
    children, names = {}, {}
    counter = [0]

    def node():
        nid = counter[0]
        counter[0] += 1
        children[nid] = []
        if s[pos[0]] == "(":
            pos[0] += 1
            while True:
                children[nid].append(node())
                if s[pos[0]] == ",":
                    pos[0] += 1
                    continue
                assert s[pos[0]] == ")"
                pos[0] += 1
                break
        m = re.match(r"[A-Za-z0-9_.]*", s[pos[0]:])
        names[nid] = m.group(0)
        pos[0] += len(m.group(0))
        return nid

    root = node()
    return children, names, root


def main():
    src = open(os.path.join(REF, "SimpleOptimizations", "SmallCodon.bf")).read()
    tmp = tempfile.mkdtemp(prefix="hb2sc_")
    os.makedirs(os.path.join(tmp, "SimpleOptimizations"))
    os.makedirs(os.path.join(tmp, "Shared"))
    shutil.copy(os.path.join(REF, "Shared", "TestInstrumentation.bf"), os.path.join(tmp, "Shared"))
    assert "Optimize 			(res,lf);" in src
    open(os.path.join(tmp, "SimpleOptimizations", "SmallCodon.bf"), "w").write(src.replace("Optimize 			(res,lf);", "Optimize 			(res,lf);\n" + DUMP))
    out = subprocess.run([BIN, os.path.join(tmp, "SimpleOptimizations", "SmallCodon.bf")], cwd=os.path.join(tmp, "SimpleOptimizations"),
                         stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=600).stdout
    shutil.rmtree(tmp)
    assert "[TEST PASSED]" in out, out[-2000:]
    vals, branches = {}, {}
    for line in out.splitlines():
        if line.startswith("HB2DUMP BRANCH"):
            _, _, nm, v = line.split()
            branches[nm] = float(v)
        elif line.startswith("HB2DUMP"):
            _, k, v = line.split()
            vals[k] = float(v)
    golden = float(re.search(r"_expectedLL\s*=\s*(-[0-9.]+)", src).group(1))
    print("reference reached", vals["lnL"], "golden", golden, "branches", len(branches))
    # --- model: formulas and frequencies from the test text, evaluated at the fitted values
    env = {"AC": vals["AC"], "AT": vals["AT"], "CT": vals["CT"], "R": vals["R"], "CG": vals["AT"], "GT": vals["AT"]}
    forms = re.findall(r"MG94custom\[(\d+)\]\[(\d+)\]:=([^;]+);", src)
    freq_txt = src[src.index("vectorOfFrequencies={") + len("vectorOfFrequencies={"):]
    freq_txt = freq_txt[:freq_txt.index(";")]
    pi = np.array([float(x) for x in re.findall(r"\{\s*([0-9.eE+-]+)\s*\}", freq_txt)])
    assert pi.shape == (61,) and abs(pi.sum() - 1) < 1e-6

    def Q_of(syn):
        Q = np.zeros((61, 61))
        e = dict(env, synRate=syn)
        for i, j, expr in forms:
            Q[int(i), int(j)] = eval(expr, {"__builtins__": {}}, e)
        Q[np.diag_indices(61)] = -Q.sum(axis=1)
        return Q
    # --- tree
    tree_txt = re.search(r"Tree givenTree=([^;]+);", src).group(1)
    children, names, root = parse_newick(tree_txt)
    order = []

    def post(n):
        for c in children[n]:
            post(c)
        order.append(n)
    post(root)
    leaves = [n for n in order if not children[n]]
    internals = [n for n in order if children[n]]
    L, I = len(leaves), len(internals)
    fid = {n: k for k, n in enumerate(leaves)}
    fid.update({n: L + k for k, n in enumerate(internals)})
    parent = {c: p for p in children for c in children[p]}
    flat_parents = np.full(L + I, -1, dtype=np.int64)
    t = np.zeros(L + I)
    for n in order:
        if n in parent:
            flat_parents[fid[n]] = fid[parent[n]] - L
            t[fid[n]] = branches[names[n]]
    # --- data: NEXUS matrix -> sense-codon states per leaf
    seqs = dict(re.findall(r"'([A-Za-z0-9_]+)'\s+([ACGTacgtNn?-]{100,})", src))
    states = np.zeros((L, 440), dtype=np.int64)
    amb_rows, amb_index = [], {}
    for n in leaves:
        sq = seqs[names[n]].upper()
        assert len(sq) == 1320
        for s in range(440):
            vec = synth.resolve_codon(sq[3 * s:3 * s + 3].replace("?", "N"))
            if vec.sum() == 1:
                states[fid[n], s] = int(np.argmax(vec))
            else:
                key = vec.tobytes()
                if key not in amb_index:
                    amb_index[key] = len(amb_rows)
                    amb_rows.append(vec)
                states[fid[n], s] = -(amb_index[key] + 1)
    leaf_states, freq, s2p = synth.compress(states)
    Qt = np.stack([Q_of(t[b]) for b in range(L + I - 1)])
    ambig = np.array(amb_rows).reshape(len(amb_rows), 61)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "smallcodon_fit.npz"), flat_parents=flat_parents, leaf_states=leaf_states,
                        pattern_freq=freq, ambig=ambig, Qt=Qt, pi=pi, lnL_reference_run=np.float64(vals["lnL"]), lnL_golden=np.float64(golden),
                        n_leaves=np.int64(L))
    # sanity: the oracle at these numbers
    from oracle import port
    tree = synth.FlatTree(L, I, flat_parents, [names[n] for n in leaves + internals], t, tree_txt)
    w = synth.Workload("smallcodon_fit", tree, 61, pi, [np.eye(61)], np.array([1.0]), leaf_states, ambig, freq, s2p)
    lnl, _ = port.lnl(w, Qt=Qt[None])
    print("oracle", lnl, "diff to reference run", lnl - vals["lnL"], "patterns", leaf_states.shape[1], "ambiguity rows", len(amb_rows))


if __name__ == "__main__":
    main()
