"""CPU, world_size 2 over gloo: the multi-process read sharding + ordered host-side gather (no data-path collective).
The per-shard compute is stood in for by the oracle (no GPU here); what is under test is that shards are contiguous,
balanced, cover every read once, and that the gathered result equals the single-process result in input order."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, path, seqs, off, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    from col_bwt_b200 import sharding
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s, o, first_read, first_base = sharding.local_shard(seqs, off, rank, world)
    pml, cid = oracle.Oracle(path).query_batch(s, o)          # stand-in for tbl.query(s, o) on this rank's GPU
    got = sharding.gather_ordered(pml, cid)
    if rank == 0:
        ret["pml"], ret["cid"] = got
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_and_gather_in_input_order(small_index):
    import oracle
    from col_bwt_b200 import sharding
    seqs, off = small_index["seqs"][:60000], small_index["off"][:401]
    b = sharding.shard_bounds(off, 2)
    assert b[0] == 0 and b[-1] == 400 and 150 < b[1] < 250
    want = oracle.Oracle(small_index["path"]).query_batch(seqs, off)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, small_index["path"], seqs, off, ret), nprocs=2, join=True)
    assert np.array_equal(ret["pml"], want[0]) and np.array_equal(ret["cid"], want[1])


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_bounds_cover_everything_once(world):
    from col_bwt_b200 import sharding
    rng = np.random.default_rng(world)
    lens = rng.integers(0, 500, 1000)
    off = np.concatenate(([0], np.cumsum(lens))).astype(np.uint64)
    b = sharding.shard_bounds(off, world)
    assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all()
    bases = np.diff(off[b].astype(np.int64))
    assert bases.sum() == lens.sum() and bases.max() - bases.min() <= 2 * 500
    seqs = np.zeros(int(off[-1]), np.uint8)
    tot = 0
    for r in range(world):
        s, o, fr, fb = sharding.local_shard(seqs, off, r, world)
        assert o[0] == 0 and fr == b[r] and fb == off[b[r]] and s.size == o[-1]
        tot += s.size
    assert tot == seqs.size
