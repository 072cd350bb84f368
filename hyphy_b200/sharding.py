"""Host-side pattern sharding for the multi-GPU path: patterns are independent given the transition
matrices, so the S unique patterns are cut into `world` contiguous shards balanced by count; pattern frequencies travel
with the shard, every rank holds the full tree and all matrices, and the only exchange per evaluation is one sum of
partial log-likelihoods (ncclAllReduce inside the engine, hb2_comm_init)."""
from __future__ import annotations


def shard_bounds(n_patterns: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous pattern shard; sizes differ by at most one and cover 0..n_patterns exactly."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    if n_patterns < world:
        raise ValueError(f"{n_patterns} patterns cannot be sharded over {world} ranks")
    return rank * n_patterns // world, (rank + 1) * n_patterns // world


def class_groups(world: int, n_classes: int) -> int:
    """Number of class groups for `world` ranks: the largest divisor of n_classes that also divides world.  Classes are
    the cheaper axis to shard (each rank exponentiates only its own classes' matrices), so they are used first; the
    remaining world // groups factor shards the patterns."""
    g = 1
    for d in range(1, n_classes + 1):
        if n_classes % d == 0 and world % d == 0:
            g = d
    return g


def layout(world: int, rank: int, n_classes: int, n_patterns: int) -> dict:
    """Rank's place in the (pattern shards) x (class groups) grid of hb2_comm_class_groups: rank r is class group
    r % G of pattern shard r // G; the G consecutive ranks of a shard are created with the same pattern slice."""
    G = class_groups(world, n_classes)
    shards = world // G
    shard = rank // G
    lo, hi = shard_bounds(n_patterns, shards, shard)
    per = n_classes // G
    return {"groups": G, "group": rank % G, "shards": shards, "shard": shard, "patterns": (lo, hi),
            "classes": (rank % G * per, (rank % G + 1) * per)}


def exchange_unique_id(dist, rank: int, make_id) -> bytes:
    """Rank 0 creates the 128-byte communicator id (hb2_comm_unique_id) and broadcasts it over the host transport
    (`dist` = torch.distributed; in HyPhy this would be MPISendString, batchlan.cpp:189)."""
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("communicator id must be 128 bytes")
    return bytes(uid)
