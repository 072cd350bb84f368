"""CPU suite, world_size 2 over gloo: the host-side logic of the multi-GPU path -- pattern sharding, the id exchange
bench.py uses to bootstrap the engine's NCCL communicator, and the identity `sum over shards of partial lnL == lnL`
that makes one scalar all-reduce per evaluation sufficient.  The partial likelihoods come from the oracle here (there
is no GPU); on the GPU box bench.py --gpus N runs the same plumbing over NCCL."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from hyphy_b200.sharding import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    for S in (2, 7, 250, 1990, 1991):
        for world in (1, 2, 3, 4, 8):
            if S < world:
                with pytest.raises(ValueError):
                    shard_bounds(S, world, 0)
                continue
            b = [shard_bounds(S, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == S
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys, dataclasses
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["HB2_ROOT"])
    from hyphy_b200 import synth
    from hyphy_b200.sharding import shard_bounds, exchange_unique_id
    from oracle import port
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    w = synth.codon_workload(10, 48, 4, ambig_frac=0.02, seed=5)
    full, _ = port.lnl(w)
    lo, hi = shard_bounds(w.S, world, rank)
    ws = dataclasses.replace(w, leaf_states=np.ascontiguousarray(w.leaf_states[:, lo:hi]), pattern_freq=w.pattern_freq[lo:hi])
    part, _ = port.lnl(ws)
    t = torch.tensor([part], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)             # the engine does this with one fp64 ncclAllReduce
    uid = exchange_unique_id(dist, rank, lambda: bytes(range(128)))
    ms = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)            # bench.py: time = max over ranks
    ok = abs(t.item() - full) <= 1e-11 * abs(full) and uid == bytes(range(128)) and ms.item() == float(world)
    print(f"RANK{rank} ok={ok} sum={t.item()!r} full={full!r}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
""")


def test_world_size_2_gloo_partial_sums(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_ = s.getsockname()[1]
    env = dict(os.environ, HB2_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port_), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "RANK0 ok=True" in r.stdout and "RANK1 ok=True" in r.stdout


def test_class_group_layout():
    from hyphy_b200.sharding import class_groups, layout
    assert [class_groups(n, 4) for n in (1, 2, 3, 4, 6, 8)] == [1, 2, 1, 4, 2, 4]
    assert class_groups(8, 3) == 1 and class_groups(6, 3) == 3 and class_groups(4, 1) == 1
    for world, C, S in ((8, 4, 1990), (4, 4, 1990), (2, 4, 7), (6, 3, 100), (3, 4, 50)):
        lays = [layout(world, r, C, S) for r in range(world)]
        G = lays[0]["groups"]
        cover = np.zeros((C, S), dtype=int)
        for r, l in enumerate(lays):
            assert l["group"] == r % G and l["shard"] == r // G and l["shards"] * G == world
            (lo, hi), (c0, c1) = l["patterns"], l["classes"]
            cover[c0:c1, lo:hi] += 1
            # the G ranks of a shard share one pattern slice
            assert l["patterns"] == lays[(r // G) * G]["patterns"]
        assert (cover == 1).all()          # every (class, pattern) cell is pruned by exactly one rank


CG_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["HB2_ROOT"])
    from hyphy_b200 import synth
    from hyphy_b200.sharding import layout
    from oracle import port
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    w = synth.codon_workload(10, 48, 4, ambig_frac=0.02, seed=5)
    full, _ = port.lnl(w)
    lay = layout(world, rank, w.C, w.S)
    assert lay["groups"] == 2 and lay["shards"] == 1
    # what hb2_comm_class_groups does: per-pattern partial over the owned classes as (value, exponent) ...
    Qt = w.Qt()
    c0, c1 = lay["classes"]
    val = np.zeros(w.S); ex = np.full(w.S, -10**9)
    parts = []
    for c in range(c0, c1):
        P = np.stack([port.expm(Qt[c, b]) for b in range(Qt.shape[1])])
        sl, ss = port.prune(w, P)                      # L = sl * 2^(-64*ss)
        parts.append((w.class_weights[c] * sl, -64 * ss))
    emax = np.max([e for _, e in parts], axis=0)
    val = sum(m * np.exp2((e - emax).astype(float)) for m, e in parts)
    # ... one all-gather of both arrays, merge, and each pattern counted by exactly one rank of the shard
    mine = torch.tensor(np.stack([val, emax.astype(float)]))
    got = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(got, mine)
    E = np.max([g[1].numpy() for g in got], axis=0)
    tot = sum(g[0].numpy() * np.exp2(g[1].numpy() - E) for g in got)
    lnl_s = np.log(tot) + E * np.log(2.0)
    keep = (np.arange(w.S) % lay["groups"]) == lay["group"]
    t = torch.tensor([float(np.sum(w.pattern_freq[keep] * lnl_s[keep]))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ok = abs(t.item() - full) <= 1e-11 * abs(full)
    print(f"RANK{rank} ok={ok} sum={t.item()!r} full={full!r}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
""")


def test_world_size_2_gloo_class_groups(tmp_path):
    """The exchange hb2_comm_class_groups adds (class partials as value+exponent, all-gather, merge, dedup by pattern
    index, then the usual sum) reproduces the full likelihood; partials from the oracle, transport over gloo."""
    script = tmp_path / "worker_cg.py"
    script.write_text(CG_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port_ = s.getsockname()[1]
    env = dict(os.environ, HB2_ROOT=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port_), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "RANK0 ok=True" in r.stdout and "RANK1 ok=True" in r.stdout
