"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/hb2_oracle.c (the CPU restatement of the reference path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hb2_oracle.c")
LIB = os.path.join(HERE, "libhb2_oracle.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.hb2o_expm.restype = C.c_int
        _lib.hb2o_expm.argtypes = [_dp, C.c_int, C.c_int, _dp]
        _lib.hb2o_prune.restype = C.c_int
        _lib.hb2o_prune.argtypes = [C.c_int64, C.c_int, C.c_int64, C.c_int64, _ip, _ip, _dp, C.c_int64, _dp, _dp, _dp, _ip, _dp]
        _lib.hb2o_prune_forced.restype = C.c_int
        _lib.hb2o_prune_forced.argtypes = [C.c_int64, C.c_int, C.c_int64, C.c_int64, _ip, _ip, _dp, C.c_int64, _dp, _dp, _dp, _ip, C.c_int64, _ip]
        _lib.hb2o_combine.restype = None
        _lib.hb2o_combine.argtypes = [C.c_int64, C.c_int64, _dp, _dp, _ip, _dp, _ip]
        _lib.hb2o_sum.restype = C.c_double
        _lib.hb2o_sum.argtypes = [C.c_int64, _dp, _ip, _ip, _dp]
        _lib.hb2o_lnl.restype = C.c_double
        _lib.hb2o_lnl.argtypes = [C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_int64, _ip, _ip, _dp, C.c_int64, _ip,
                                  _dp, C.c_int, _dp, _dp, _dp]
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_ip)


def expm(A: np.ndarray, sparse_storage: bool = False) -> np.ndarray:
    A, pA = _d(A)
    D = A.shape[0]
    P = np.empty((D, D))
    rc = lib().hb2o_expm(pA, D, int(sparse_storage), P.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError(f"hb2o_expm failed rc={rc}")
    return P


def prune(w, P: np.ndarray, want_cond: bool = False):
    """One rate class.  P: [B, D, D].  Returns (siteL[S], siteScale[S][, cond[I,S,D]])."""
    t = w.tree
    S, D = w.S, w.D
    fp, pfp = _i(t.flat_parents)
    ls, pls = _i(w.leaf_states)
    am, pam = _d(w.ambig if len(w.ambig) else np.zeros((1, D)))
    Pm, pP = _d(P)
    pi, ppi = _d(w.pi)
    sl = np.empty(S)
    ss = np.empty(S, dtype=np.int64)
    cond = np.empty((t.n_internal, S, D)) if want_cond else None
    rc = lib().hb2o_prune(S, D, t.n_leaves, t.n_internal, pfp, pls, pam, len(w.ambig), pP, ppi,
                          sl.ctypes.data_as(_dp), ss.ctypes.data_as(_ip),
                          cond.ctypes.data_as(_dp) if want_cond else None)
    if rc:
        raise RuntimeError(f"hb2o_prune failed rc={rc}")
    return (sl, ss, cond) if want_cond else (sl, ss)


def prune_forced(w, P: np.ndarray, forced_node: int, forced_states):
    """One rate class with ONE node pinned (flat node id: leaves 0..L-1, internals L..).  Returns (siteL[S], siteScale[S])."""
    t = w.tree
    S, D = w.S, w.D
    fp, pfp = _i(t.flat_parents)
    ls, pls = _i(w.leaf_states)
    am, pam = _d(w.ambig if len(w.ambig) else np.zeros((1, D)))
    Pm, pP = _d(P)
    pi, ppi = _d(w.pi)
    fs, pfs = _i(forced_states)
    set_branch = forced_node - t.n_leaves if forced_node >= t.n_leaves else t.n_internal + forced_node   # reference convention
    sl = np.empty(S)
    ss = np.empty(S, dtype=np.int64)
    rc = lib().hb2o_prune_forced(S, D, t.n_leaves, t.n_internal, pfp, pls, pam, len(w.ambig), pP, ppi,
                                 sl.ctypes.data_as(_dp), ss.ctypes.data_as(_ip), set_branch, pfs)
    if rc:
        raise RuntimeError(f"hb2o_prune_forced failed rc={rc}")
    return sl, ss


def lnl(w, Qt: np.ndarray | None = None, weights=None, sparse_storage: bool | None = None):
    """Full likelihood of workload `w` (all rate classes).  Returns (lnL, per-pattern lnL[S])."""
    t = w.tree
    S, D = w.S, w.D
    if Qt is None:
        Qt = w.Qt()
    if sparse_storage is None:
        sparse_storage = D > 20
    fp, pfp = _i(t.flat_parents)
    ls, pls = _i(w.leaf_states)
    am, pam = _d(w.ambig if len(w.ambig) else np.zeros((1, D)))
    fr, pfr = _i(w.pattern_freq)
    Q, pQ = _d(Qt)
    wt, pwt = _d(w.class_weights if weights is None else weights)
    pi, ppi = _d(w.pi)
    site = np.empty(S)
    v = lib().hb2o_lnl(S, D, t.n_leaves, t.n_internal, Q.shape[0], pfp, pls, pam, len(w.ambig), pfr, pQ,
                       int(sparse_storage), pwt, ppi, site.ctypes.data_as(_dp))
    return float(v), site


def mixture_P(w, perturb: float = 0.0) -> np.ndarray:
    """Explicit-form mixture transition matrices [B, D, D]: sum_k w_k Exp(Q_k t_b), each Exp by the reference's algorithm
    (tree.cpp:3047-3089 evaluates the formula on the exponentiated components)."""
    M, wk = w.mixture_Qt(perturb)
    return np.stack([sum(wk[b, k] * expm(M[b, k], True) for k in range(M.shape[1])) for b in range(M.shape[0])])


def lnl_mixture(w, perturb: float = 0.0):
    """(lnL, per-pattern log-likelihoods) of an explicit-form mixture workload (one rate class)."""
    sl, ss = prune(w, mixture_P(w, perturb))
    site = np.log(sl) - 64.0 * np.log(2.0) * ss
    return float((w.pattern_freq * site).sum()), site
