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
