"""Read sharding for multi-process runs (one process per GPU): the path has no cross-read dependence
(src/pml_query.cpp:74-86), so a batch is cut into `world` contiguous read ranges of ~equal bases, every rank
traverses its own range against its own replica of the index, and the results are concatenated host-side in input
order.  No collective touches the data path; `gather_ordered` is only the final host-side hand-over to rank 0."""
from __future__ import annotations

import numpy as np


def shard_bounds(offsets: np.ndarray, world: int) -> np.ndarray:
    """world+1 read indices b with shard r = reads [b[r], b[r+1]); shards are contiguous and balanced by bases."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n_reads = offsets.size - 1
    total = int(offsets[-1] - offsets[0])
    targets = int(offsets[0]) + (np.arange(world + 1, dtype=np.float64) * total / world).astype(np.uint64)
    b = np.searchsorted(offsets, targets, side="left").astype(np.int64)
    b[0], b[-1] = 0, n_reads
    return np.maximum.accumulate(np.clip(b, 0, n_reads))


def local_shard(seqs: np.ndarray, offsets: np.ndarray, rank: int, world: int):
    """(seqs_r, offsets_r rebased to 0, first_read, first_base) of this rank's shard."""
    b = shard_bounds(offsets, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    base = int(offsets[lo])
    return seqs[base:int(offsets[hi])], (offsets[lo:hi + 1] - offsets[lo]).astype(np.uint64), lo, base


def gather_ordered(pml: np.ndarray, cid: np.ndarray, group=None, dst: int = 0):
    """Host-side ordered gather: rank `dst` gets the concatenation of all ranks' arrays in rank order (= input
    order, because shards are contiguous); other ranks get None.  Uses the process group's object gather (gloo or
    NCCL-backed groups both work; the arrays live in host memory)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    out = [None] * world if rank == dst else None
    dist.gather_object((pml, cid), out, dst=dst, group=group)
    if rank != dst:
        return None
    return np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out])
