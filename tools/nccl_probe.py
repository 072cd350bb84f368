"""Bring-up probe for the engine's own NCCL communicator under torchrun: prints a rank-tagged line after every stage."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
def say(*a):
    print(f"[rank {rank} t={time.time() % 1000:.2f}]", *a, file=sys.stderr, flush=True)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
say("pg up")
from hyphy_b200 import synth, LikelihoodFunction, Partition
from hyphy_b200.sharding import shard_bounds, exchange_unique_id
w = synth.codon_workload(12, 600, 4)
lo, hi = shard_bounds(w.S, world, rank)
lf = LikelihoodFunction(w, device=lr, pattern_slice=slice(lo, hi))
say("partition created", lo, hi)
lf.set_template(); lf.set_all_compiled()
part = lf.compute()
say("local partial lnL", part)
uid = exchange_unique_id(dist, rank, Partition.comm_unique_id)
say("uid exchanged", uid[:8].hex())
lf.part.comm_init(world, rank, uid)
say("comm_init done")
tot = lf.compute()
say("summed lnL", tot)
t = torch.tensor([part], dtype=torch.float64, device="cuda"); dist.all_reduce(t)
say("torch sum", t.item(), "engine sum", tot, "diff", abs(t.item() - tot))
ms, st, l2 = lf.part.time_resident(w.class_weights, w.pi, iters=5)
say("time_resident ok", ms, l2)
lf.close()
say("closed")
dist.barrier(); dist.destroy_process_group()
say("done")
